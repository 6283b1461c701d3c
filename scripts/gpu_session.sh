# round 2, session z2: brick-plane burst walker that reads a brick's record before it does anything else for it
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 -k "tail_handoff" > gpurun_out/r02z3_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02z3_tests.log | cut -c1-300
M="OCLR_HANDOFF_MAX_PATHS=4000000 OCLR_HANDOFF_MODE=1"
( timeout 400 python scripts/share_sweep.py 2 64 "OCLR_X=off" "$M OCLR_HANDOFF_AFTER=2" "$M OCLR_HANDOFF_AFTER=4" "$M OCLR_HANDOFF_AFTER=6" "$M OCLR_HANDOFF_AFTER=8" "$M OCLR_HANDOFF_AFTER=12"
  timeout 400 python scripts/share_sweep.py 2 8 "OCLR_X=off" "$M OCLR_HANDOFF_AFTER=8" "$M OCLR_HANDOFF_AFTER=12" "$M OCLR_HANDOFF_AFTER=16" "$M OCLR_HANDOFF_AFTER=24" "OCLR_X=off"
  timeout 400 python scripts/share_sweep.py 2 16 "OCLR_X=off" "$M OCLR_HANDOFF_AFTER=8" "$M OCLR_HANDOFF_AFTER=12" ) 2>&1 | tee gpurun_out/r02z3_share.log
