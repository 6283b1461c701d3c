# round 2, session p: hand-off mode 2 -- rays given up by thinly filled warps go back into a queue and a second pass packs them densely
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 -k "tail_handoff" > gpurun_out/r02p_tests.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02p_tests.log
M="OCLR_HANDOFF_MAX_PATHS=4000000 OCLR_HANDOFF_MODE=2"
( timeout 400 python scripts/share_sweep.py 2 8 "OCLR_X=off" "$M OCLR_HANDOFF_AFTER=0 OCLR_HANDOFF_LANES=8" "$M OCLR_HANDOFF_AFTER=0 OCLR_HANDOFF_LANES=16" "$M OCLR_HANDOFF_AFTER=0 OCLR_HANDOFF_LANES=24" "$M OCLR_HANDOFF_AFTER=1 OCLR_HANDOFF_LANES=16" "$M OCLR_HANDOFF_AFTER=2 OCLR_HANDOFF_LANES=16" "$M OCLR_HANDOFF_AFTER=0 OCLR_HANDOFF_LANES=32" "OCLR_X=off"
  timeout 300 python scripts/share_sweep.py 2 4 "OCLR_X=off" "$M OCLR_HANDOFF_AFTER=0 OCLR_HANDOFF_LANES=16" "$M OCLR_HANDOFF_AFTER=1 OCLR_HANDOFF_LANES=12"
  timeout 300 python scripts/share_sweep.py 3 8 "OCLR_X=off" "$M OCLR_HANDOFF_AFTER=0 OCLR_HANDOFF_LANES=16" "$M OCLR_HANDOFF_AFTER=1 OCLR_HANDOFF_LANES=12"
  timeout 300 python scripts/share_sweep.py 2 1 "OCLR_X=off" "$M OCLR_HANDOFF_AFTER=0 OCLR_HANDOFF_LANES=16" ) 2>&1 | tee gpurun_out/r02p_share.log
