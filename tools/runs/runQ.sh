cd /root/repo
# same-box A/B of the whole step: the library as it was at the start of this half of the round (commit 28aac48) vs HEAD
for i in 1 2 3; do
B200SSL_LIB=/root/repo/gipmed-project-self-supervised-vit_b200/build/libb200ssl_start.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('start', round(d['ms_per_step'],2), round(d['value'],1), d['clocks']['sm_mhz'], round(d['roofline']['gemm_ms_per_step'],2))"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('head ', round(d['ms_per_step'],2), round(d['value'],1), d['clocks']['sm_mhz'], round(d['roofline']['gemm_ms_per_step'],2))"
done
