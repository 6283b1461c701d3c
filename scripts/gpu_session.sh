# round 2, session t: hand-off by the AGE of a ray (outer iterations it has been with its warp) once the queue is dry: only long walks go
# to the burst walker
set -x
cd $GRAFT_REPO_ROOT
M="OCLR_HANDOFF_MAX_PATHS=4000000 OCLR_HANDOFF_MODE=1"
( timeout 400 python scripts/share_sweep.py 2 8 "OCLR_X=off" "$M OCLR_HANDOFF_AFTER=2" "$M OCLR_HANDOFF_AFTER=4" "$M OCLR_HANDOFF_AFTER=6" "$M OCLR_HANDOFF_AFTER=8" "$M OCLR_HANDOFF_AFTER=12" "$M OCLR_HANDOFF_AFTER=16" "$M OCLR_HANDOFF_AFTER=24" "OCLR_X=off"
  timeout 300 python scripts/share_sweep.py 2 64 "OCLR_X=off" "$M OCLR_HANDOFF_AFTER=4" "$M OCLR_HANDOFF_AFTER=8" "$M OCLR_HANDOFF_AFTER=16"
  timeout 300 python scripts/share_sweep.py 3 8 "OCLR_X=off" "$M OCLR_HANDOFF_AFTER=4" "$M OCLR_HANDOFF_AFTER=8" "$M OCLR_HANDOFF_AFTER=16" ) 2>&1 | tee gpurun_out/r02t_share.log
timeout 600 python -m pytest tests -m gpu -x -q --timeout 600 -k "tail_handoff" 2>&1 | tail -2
