set -x
cd /root/repo
for c in 4 5a 5b 1; do
  python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2_bench_config${c}_final.json 2> gpurun_out/bench_c$c.err; echo "rc=$?" >> gpurun_out/bench_c$c.err
done
python bench.py --e2e-tiles --steps 10 --warmup 3 > gpurun_out/r2_bench_config2_e2e_tiles_final.json 2> gpurun_out/bench_e2e_tiles.err
python bench.py --ncu-step --warmup 2 > gpurun_out/ncu_step_plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_launches_step_v2.csv python bench.py --ncu-step --warmup 2 > gpurun_out/ncu_step.log 2>&1
tail -n 2 gpurun_out/ncu_step.log
python tests/gpu_checks/attn_prof.py 512 197 6 > gpurun_out/attn_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention_ -s 2 -c 2 -o gpurun_out/r2_attn_n197_final -f python tests/gpu_checks/attn_prof.py 512 197 6 > gpurun_out/ncu_attn197.log 2>&1
python tests/gpu_checks/attn_prof.py 2560 37 6 > gpurun_out/attn_prof_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention_ -s 2 -c 2 -o gpurun_out/r2_attn_n37_final -f python tests/gpu_checks/attn_prof.py 2560 37 6 > gpurun_out/ncu_attn37.log 2>&1
tail -n 2 gpurun_out/ncu_attn197.log gpurun_out/ncu_attn37.log
for c in 4 5a 5b 1; do cut -c1-260 gpurun_out/r2_bench_config${c}_final.json; tail -n 1 gpurun_out/bench_c$c.err; done
