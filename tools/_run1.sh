cd /root/repo
timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2_pytest_gpu_s5.log 2>&1; tail -n 3 gpurun_out/r2_pytest_gpu_s5.log
timeout 120 python tests/gpu_checks/attn_roles.py > gpurun_out/attn_roles_new.log 2>&1; grep "bwd B\|issuer" gpurun_out/attn_roles_new.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_s5.json 2> gpurun_out/r2_bench_s5.err; cut -c1-330 gpurun_out/r2_bench_s5.json
