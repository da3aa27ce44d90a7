set -x
cd /root/repo
python tests/gpu_checks/gemm_prof.py > gpurun_out/gemm_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 7 -c 7 -o gpurun_out/r2_gemm_final -f python tests/gpu_checks/gemm_prof.py > gpurun_out/ncu_gemm.log 2>&1
tail -n 3 gpurun_out/ncu_gemm.log
