set -x
cd $GRAFT_REPO_ROOT
for c in 4 5a 5b 1; do
  python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err; echo "rc=$?" >> gpurun_out/bench_c$c.err
done
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_launches_step_v1.csv python bench.py --ncu-step --warmup 2 > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log
for c in 4 5a 5b 1; do tail -c 1500 gpurun_out/bench_c$c.json; tail -2 gpurun_out/bench_c$c.err; done
