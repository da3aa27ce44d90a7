cd $GRAFT_REPO_ROOT
echo "=== stream mode 1"; timeout 300 python tests/gpu_checks/attn_check.py --stream 1 --bench 2>&1 | grep -v Warn | tail -28
timeout 200 python tests/gpu_checks/attn_stream_roles.py 2>&1 | grep -v Warn | tail
