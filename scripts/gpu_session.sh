# round 2, last look at the final binaries: smoke() and the golden / hand-off tests
set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()"
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 -k "golden or tail_handoff or super or ring or packers" 2>&1 | tail -3
