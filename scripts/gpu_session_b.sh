set -x
cd $GRAFT_REPO_ROOT
for w in 8 1; do timeout 200 python scripts/tail_probe.py 2 $w 2>&1 | tail -4; done
timeout 200 python scripts/tail_probe.py 3 8 2>&1 | tail -4
