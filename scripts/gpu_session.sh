# round 2, session k: both level switches in one instruction stream (pwalk_switch) against the two-switch build; switch threshold on top
set -x
cd $GRAFT_REPO_ROOT
P=$GRAFT_REPO_ROOT/opencl_render_b200/libopencl_render_b200
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 -k "golden or super or whole_frame or config3" > gpurun_out/r02y_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02y_tests.log
( timeout 400 python scripts/sweep_env.py 2 2 "OCLR_LIB=${P}_splitsw.so" "OCLR_X=default" "OCLR_SWITCH_MIN=4" "OCLR_SWITCH_MIN=8" "OCLR_SWITCH_MIN=12" "OCLR_LIB=${P}_splitsw.so" "OCLR_X=default"
  timeout 400 python scripts/sweep_env.py 3 2 "OCLR_LIB=${P}_splitsw.so" "OCLR_X=default" "OCLR_SWITCH_MIN=8" "OCLR_SWITCH_MIN=12"
  timeout 400 python scripts/sweep_env.py 5 2 "OCLR_LIB=${P}_splitsw.so" "OCLR_X=default" ) > gpurun_out/r02y_ab.log 2>&1
grep -E "^---|frame|switch util" gpurun_out/r02y_ab.log
( timeout 300 python scripts/share_sweep.py 2 8 "OCLR_LIB=${P}_splitsw.so" "OCLR_X=default" "OCLR_SWITCH_MIN=8" ) 2>&1 | tee gpurun_out/r02y_share.log
