set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q -k "ddp_two_ranks" > gpurun_out/r2_ddp_two_ranks.log 2>&1; echo "rc=$?" >> gpurun_out/r2_ddp_two_ranks.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/gpu_checks/ddp_check.py > gpurun_out/r2_ddp_check_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2_ddp_check_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2_graphnccl.json 2> gpurun_out/r2_bench_n2_graphnccl.err; echo "rc=$?" >> gpurun_out/r2_bench_n2_graphnccl.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 10 --warmup 3 --no-nccl-graph --no-e2e > gpurun_out/r2_bench_n2_eagernccl.json 2> gpurun_out/r2_bench_n2_eagernccl.err; echo "rc=$?" >> gpurun_out/r2_bench_n2_eagernccl.err
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_n1_samebox.json 2>/dev/null
tail -5 gpurun_out/r2_ddp_two_ranks.log; tail -12 gpurun_out/r2_ddp_check_2gpu.log
for f in r2_bench_n2_graphnccl r2_bench_n2_eagernccl r2_bench_n1_samebox; do python -c "
import json,sys
d=json.load(open('gpurun_out/$f.json')); print('$f', d['value'], d['ms_per_step'], d['step_api'], d['clocks'])"; done
