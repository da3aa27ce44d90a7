cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -x -q -k "attention or other_baseline or frozen or embed" 2>&1 | tail -3
timeout 300 python bench.py --config 5a --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_config5a_paired.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_config5a_paired.json')); print('5a', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['gpu_launches'])"
