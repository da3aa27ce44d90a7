cd $GRAFT_REPO_ROOT
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8_config2.json 2> gpurun_out/r2_bench_n8_config2.err; echo "rc=$?" >> gpurun_out/r2_bench_n8_config2.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 8 --config 4 --steps 10 --warmup 3 --no-e2e > gpurun_out/r2_bench_n8_config4.json 2> gpurun_out/r2_bench_n8_config4.err; echo "rc=$?" >> gpurun_out/r2_bench_n8_config4.err
timeout 120 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_n1_samebox8.json 2>/dev/null
for f in r2_bench_n8_config2 r2_bench_n8_config4 r2_bench_n1_samebox8; do python -c "
import json,sys
try:
    txt=open('gpurun_out/$f.json').read(); line=[l for l in txt.splitlines() if l.startswith('{')][-1]
    d=json.loads(line); print('$f', d['value'], d['ms_per_step'], d['step_api'], d['clocks'], d.get('e2e'))
except Exception as e: print('$f', 'ERR', e)"; tail -2 gpurun_out/$f.err 2>/dev/null | cut -c1-300; done
