# round 2, session i: full GPU suite on the final kernels, logic-kernel occupancy / outline variants, the default bench line,
# and the round's ncu evidence (profile_round.sh r02b)
set -x
cd $GRAFT_REPO_ROOT
P=$GRAFT_REPO_ROOT/opencl_render_b200/libopencl_render_b200
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/r02w_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02w_tests.log
( timeout 400 python scripts/sweep_env.py 2 2 "OCLR_X=default" "OCLR_LIB=${P}_l6.so" "OCLR_LIB=${P}_l4.so" "OCLR_LIB=${P}_out1.so" "OCLR_LIB=${P}_out3.so" "OCLR_X=default"
  timeout 400 python scripts/sweep_env.py 3 2 "OCLR_X=default" "OCLR_LIB=${P}_l4.so" "OCLR_LIB=${P}_out1.so" "OCLR_LIB=${P}_out3.so" ) > gpurun_out/r02w_ab.log 2>&1
grep -E "^---|frame" gpurun_out/r02w_ab.log
timeout 900 python bench.py > gpurun_out/r02w_bench_n1.json 2> gpurun_out/r02w_bench_n1.err; echo "bench rc=$?"; cat gpurun_out/r02w_bench_n1.json | cut -c1-1500
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02w_bench_ref.json 2> gpurun_out/r02w_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/r02w_bench_ref.json | cut -c1-600
timeout 900 bash scripts/profile_round.sh r02b
