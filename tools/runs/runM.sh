set -x
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k ddp_two_ranks > gpurun_out/r2_ddp_two_ranks_final.log 2>&1; tail -n 3 gpurun_out/r2_ddp_two_ranks_final.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2_final.json 2> gpurun_out/r2_bench_n2_final.err; cut -c1-300 gpurun_out/r2_bench_n2_final.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n1_samebox_final.json 2>/dev/null; cut -c1-300 gpurun_out/r2_bench_n1_samebox_final.json
