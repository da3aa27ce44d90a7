cd $GRAFT_REPO_ROOT
for m in 1; do
  echo "=== stream mode $m"; timeout 300 python tests/gpu_checks/attn_check.py --stream $m --bench 2>&1 | grep -v Warn | tail -28
done
