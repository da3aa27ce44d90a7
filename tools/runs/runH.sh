cd $GRAFT_REPO_ROOT
timeout 200 python -m pytest tests/test_augment.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python __graft_entry__.py smoke 2>&1 | grep -v Warn | tail -4
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-tiles > gpurun_out/bench_c2_e2e_tiles.json 2> gpurun_out/bench_c2_e2e_tiles.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c2_e2e_tiles.json')); print('tiles e2e', d['value'], d['ms_per_step'], d['e2e'], d['clocks'])"; tail -2 gpurun_out/bench_c2_e2e_tiles.err
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_e2e_crops.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_c2_e2e_crops.json')); print('crops e2e', d['value'], d['ms_per_step'], d['e2e'], d['clocks'])"
