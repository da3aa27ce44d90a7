set -x
cd /root/repo
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_head.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/r2_bench_default_head.json 2> gpurun_out/bench_head.err; echo "bench rc=$?"
python bench.py --ncu-step --warmup 2 > gpurun_out/ncu_step_plain.log 2>&1 && \
timeout 170 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_launches_step_v3.csv python bench.py --ncu-step --warmup 2 > gpurun_out/ncu_step.log 2>&1
echo "ncu rc=$?"
tail -n 3 gpurun_out/r2_smoke_head.log; cut -c1-400 gpurun_out/r2_bench_default_head.json; wc -l gpurun_out/r2_launches_step_v3.csv
