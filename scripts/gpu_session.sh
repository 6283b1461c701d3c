# round 2, session g: stepped-axis ray components from the shared ray table instead of six registers; 9 / 10 trace CTAs per SM on top
set -x
cd $GRAFT_REPO_ROOT
P=$GRAFT_REPO_ROOT/opencl_render_b200/libopencl_render_b200
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 -k "golden or ring or super or packers or whole_frame or split" > gpurun_out/r02u_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02u_tests.log
( timeout 400 python scripts/sweep_env.py 2 2 "OCLR_LIB=${P}_rayreg.so" "OCLR_X=default" "OCLR_LIB=${P}_ctas9.so" "OCLR_LIB=${P}_ctas10.so" "OCLR_LIB=${P}_rayreg.so" "OCLR_X=default"
  timeout 400 python scripts/sweep_env.py 3 2 "OCLR_LIB=${P}_rayreg.so" "OCLR_X=default" "OCLR_LIB=${P}_ctas9.so" "OCLR_LIB=${P}_ctas10.so" ) > gpurun_out/r02u_ab.log 2>&1
grep -E "^---|frame|walk util" gpurun_out/r02u_ab.log
( timeout 200 python scripts/share_sweep.py 2 8 "OCLR_LIB=${P}_rayreg.so" "OCLR_X=default" "OCLR_LIB=${P}_ctas9.so" ) 2>&1 | tee gpurun_out/r02u_share.log
