cd /root/repo
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/pytest_kernels.log 2>&1; tail -n 2 gpurun_out/pytest_kernels.log
timeout 300 python tests/gpu_checks/layer_gemm_bench.py > gpurun_out/layer_gemm_new6.log 2>&1
grep "gelu\|TOTAL" gpurun_out/layer_gemm_new6.log | sed 's/narrow.*wide/wide/'
timeout 120 python tests/gpu_checks/fc1_epilogue_probe.py 2>&1 | tail -8
