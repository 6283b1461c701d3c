# round 2, session r: hand-off to the burst walker for stragglers only (a warp gives its rays up once it holds at most N of them)
set -x
cd $GRAFT_REPO_ROOT
M="OCLR_HANDOFF_MAX_PATHS=4000000 OCLR_HANDOFF_MODE=1"
( timeout 400 python scripts/share_sweep.py 2 8 "OCLR_X=off" "$M OCLR_HANDOFF_AFTER=0 OCLR_HANDOFF_LANES=1" "$M OCLR_HANDOFF_AFTER=0 OCLR_HANDOFF_LANES=2" "$M OCLR_HANDOFF_AFTER=0 OCLR_HANDOFF_LANES=4" "$M OCLR_HANDOFF_AFTER=0 OCLR_HANDOFF_LANES=8" "$M OCLR_HANDOFF_AFTER=2 OCLR_HANDOFF_LANES=2" "$M OCLR_HANDOFF_AFTER=2 OCLR_HANDOFF_LANES=4" "OCLR_X=off"
  timeout 300 python scripts/share_sweep.py 2 16 "OCLR_X=off" "$M OCLR_HANDOFF_AFTER=0 OCLR_HANDOFF_LANES=2" "$M OCLR_HANDOFF_AFTER=0 OCLR_HANDOFF_LANES=4"
  timeout 300 python scripts/share_sweep.py 3 8 "OCLR_X=off" "$M OCLR_HANDOFF_AFTER=0 OCLR_HANDOFF_LANES=2" "$M OCLR_HANDOFF_AFTER=0 OCLR_HANDOFF_LANES=4" ) 2>&1 | tee gpurun_out/r02r_share.log
timeout 600 python -m pytest tests -m gpu -x -q --timeout 600 -k "tail_handoff" 2>&1 | tail -2
