cd $GRAFT_REPO_ROOT
timeout 100 python -m pytest tests/test_augment.py -m gpu -x -q 2>&1 | tail -2
timeout 100 python tests/gpu_checks/augment_bench.py 2>&1 | tail -4
timeout 200 ncu --set full --clock-control none --import-source on -k regex:attention_fwd_stream -s 1 -c 1 -o gpurun_out/r2_attn_stream_n785 -f python tests/gpu_checks/attn_prof.py 64 785 6 > gpurun_out/ncu_attn_stream.log 2>&1; tail -2 gpurun_out/ncu_attn_stream.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:multicrop_augment -s 42 -c 1 -o gpurun_out/r2_augment_jitter -f python tests/gpu_checks/augment_bench.py 256 > gpurun_out/ncu_augment.log 2>&1; tail -2 gpurun_out/ncu_augment.log
ls -la gpurun_out/*.ncu-rep
