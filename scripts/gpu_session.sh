# round 2, session h: the super-brick level again, now that the loops of the trace kernel are free of spills either way
set -x
cd $GRAFT_REPO_ROOT
P=$GRAFT_REPO_ROOT/opencl_render_b200/libopencl_render_b200
( timeout 400 python scripts/sweep_env.py 2 2 "OCLR_X=default" "OCLR_LIB=${P}_super.so OCLR_HIERARCHICAL=2" "OCLR_LIB=${P}_super.so OCLR_HIERARCHICAL=1" "OCLR_X=default" "OCLR_LIB=${P}_super.so OCLR_HIERARCHICAL=2"
  timeout 400 python scripts/sweep_env.py 3 2 "OCLR_X=default" "OCLR_LIB=${P}_super.so OCLR_HIERARCHICAL=2" "OCLR_LIB=${P}_super.so OCLR_HIERARCHICAL=1" "OCLR_X=default" "OCLR_LIB=${P}_super.so OCLR_HIERARCHICAL=2"
  timeout 400 python scripts/sweep_env.py 5 2 "OCLR_X=default" "OCLR_LIB=${P}_super.so OCLR_HIERARCHICAL=2" ) > gpurun_out/r02v_ab.log 2>&1
grep -E "^---|frame|coarse" gpurun_out/r02v_ab.log
