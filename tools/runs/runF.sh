cd $GRAFT_REPO_ROOT
i=0
for v in "--bucket-mb 100000" "--bucket-mb 100000 --grad-compress bf16" "--grad-compress bf16"; do
  i=$((i+1))
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29630+i)) bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e $v > gpurun_out/r2_n8_var$i.json 2> gpurun_out/r2_n8_var$i.err; echo "rc=$? [$v]" >> gpurun_out/r2_n8_var$i.err
done
timeout 120 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r2_n8_var0.json 2>/dev/null
for f in r2_n8_var0 r2_n8_var1 r2_n8_var2 r2_n8_var3; do python -c "
import json,sys
try:
    txt=open('gpurun_out/$f.json').read(); line=[l for l in txt.splitlines() if l.startswith('{')][-1]
    d=json.loads(line); print('$f', round(d['value']), d['ms_per_step'], d['step_api'][40:], d['clocks'])
except Exception as e: print('$f', 'ERR', e)"; tail -1 gpurun_out/$f.err 2>/dev/null | cut -c1-200; done
