set -x
cd /root/repo
python tests/gpu_checks/attn_sanitize.py > gpurun_out/attn_sanitize_plain.log 2>&1 && \
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tests/gpu_checks/attn_sanitize.py > gpurun_out/r2_attn_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -n 6 gpurun_out/r2_attn_memcheck.log
