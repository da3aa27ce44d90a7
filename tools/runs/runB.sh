cd $GRAFT_REPO_ROOT
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/gpu_checks/ddp_check.py > gpurun_out/r2_ddp_check_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2_ddp_check_2gpu.log
grep -n "rank\|rc=" gpurun_out/r2_ddp_check_2gpu.log | cut -c1-300
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2_graphnccl.json 2> gpurun_out/r2_bench_n2_graphnccl.err; echo "rc=$?" >> gpurun_out/r2_bench_n2_graphnccl.err
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 10 --warmup 3 --no-nccl-graph --no-e2e > gpurun_out/r2_bench_n2_eagernccl.json 2> gpurun_out/r2_bench_n2_eagernccl.err; echo "rc=$?" >> gpurun_out/r2_bench_n2_eagernccl.err
timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_n1_samebox.json 2>/dev/null
for f in r2_bench_n2_graphnccl r2_bench_n2_eagernccl r2_bench_n1_samebox; do python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/$f.json')); print('$f', d['value'], d['ms_per_step'], d['step_api'], d['clocks'])
except Exception as e: print('$f', 'ERR', e)"; tail -3 gpurun_out/$f.err 2>/dev/null | cut -c1-300; done
timeout 280 python -m pytest tests -m gpu -x -q -k "ddp_two_ranks" > gpurun_out/r2_ddp_two_ranks.log 2>&1; echo "rc=$?" >> gpurun_out/r2_ddp_two_ranks.log; tail -3 gpurun_out/r2_ddp_two_ranks.log
