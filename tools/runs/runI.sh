cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_final.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_gpu_final.log; tail -3 gpurun_out/r2_pytest_gpu_final.log
timeout 300 python bench.py > gpurun_out/r2_bench_config2_final.json 2> gpurun_out/r2_bench_config2_final.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_config2_final.json')); print('final', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['clocks'], d['cpu_baseline'])"
timeout 300 python tests/gpu_checks/proj_gemm_ab.py 2>&1 | grep "dgrad block_n=  0"
