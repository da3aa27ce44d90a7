cd $GRAFT_REPO_ROOT
for m in 0 -1; do echo "=== stream mode $m"; timeout 300 python tests/gpu_checks/attn_check.py --stream $m --bench 2>&1 | grep -v Warn | grep "N=257\|N=785\|N=325\|N=1025\|bench\|ALL\|FAIL" | tail -16; done
