# round 2, session l: refill inside a walk burst once enough lanes sit idle (OCLR_REFILL_BURST; 0 = between bursts only)
set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 -k "golden or whole_frame or config3 or sliced" > gpurun_out/r02z_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02z_tests.log
( timeout 400 python scripts/sweep_env.py 2 2 "OCLR_REFILL_BURST=0" "OCLR_REFILL_BURST=4" "OCLR_REFILL_BURST=6" "OCLR_REFILL_BURST=8" "OCLR_REFILL_BURST=12" "OCLR_REFILL_BURST=16" "OCLR_REFILL_BURST=0" "OCLR_REFILL_BURST=8"
  timeout 400 python scripts/sweep_env.py 3 2 "OCLR_REFILL_BURST=0" "OCLR_REFILL_BURST=4" "OCLR_REFILL_BURST=8" "OCLR_REFILL_BURST=12"
  timeout 400 python scripts/sweep_env.py 5 2 "OCLR_REFILL_BURST=0" "OCLR_REFILL_BURST=8" ) > gpurun_out/r02z_ab.log 2>&1
grep -E "^---|frame|walk util|walk-iteration" gpurun_out/r02z_ab.log
( timeout 300 python scripts/share_sweep.py 2 8 "OCLR_REFILL_BURST=0" "OCLR_REFILL_BURST=4" "OCLR_REFILL_BURST=8" "OCLR_REFILL_BURST=16" ) 2>&1 | tee gpurun_out/r02z_share.log
