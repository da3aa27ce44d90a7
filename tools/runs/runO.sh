set -x
cd /root/repo
timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2_pytest_gpu_final3.log 2>&1; tail -n 3 gpurun_out/r2_pytest_gpu_final3.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_final.log 2>&1; echo "smoke rc=$?"; tail -n 4 gpurun_out/r2_smoke_final.log
timeout 900 python bench.py > gpurun_out/r2_bench_default_final.json 2> gpurun_out/r2_bench_default_final.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2_bench_default_final.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_final.json 2>/dev/null; echo "ref rc=$?"; cut -c1-400 gpurun_out/r2_bench_reference_final.json
