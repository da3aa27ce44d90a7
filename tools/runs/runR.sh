cd /root/repo
# same-box A/B: programmatic dependent launch inside the captured step (off by default) on the final kernels
for i in 1 2; do
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('default', round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"
python bench.py --pdl --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pdl    ', round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"
done
