cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "merged_crop" 2>&1 | grep -E "assert|Error|passed|failed|rel\(" | tail -8
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q 2>&1 | tail -5
for v in "" "--no-ln-tail" "" "--no-ln-tail"; do timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e $v > gpurun_out/tmp_lnt.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/tmp_lnt.json')); print('[$v]', round(d['value']), d['ms_per_step'], d['roofline']['launches_per_step'], round(d['roofline']['gemm_ms_per_step'],2), d['gpu_launches'], d['clocks'])"; done
