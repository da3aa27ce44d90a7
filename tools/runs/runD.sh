cd $GRAFT_REPO_ROOT
timeout 300 python tests/gpu_checks/attn_check.py --bench 2>&1 | grep -v Warn | tail -12
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in 5b 5a; do timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c${c}_stream.json 2> gpurun_out/bench_c${c}_stream.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c${c}_stream.json')); print('$c', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"; done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_b.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_c2_b.json')); print('2', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
