cd /root/repo
timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2_pytest_gpu_s4.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu_s4.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_s4.json 2> gpurun_out/r2_bench_s4.err; cat gpurun_out/r2_bench_s4.json | cut -c1-600
