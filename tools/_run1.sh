cd /root/repo
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/pytest_kernels.log 2>&1; tail -n 2 gpurun_out/pytest_kernels.log
timeout 300 python tests/gpu_checks/layer_gemm_bench.py > gpurun_out/layer_gemm_new5.log 2>&1
cat gpurun_out/layer_gemm_new5.log | sed 's/narrow.*wide/wide/'
