set -x
cd $GRAFT_REPO_ROOT
P=$GRAFT_REPO_ROOT/opencl_render_b200
( for v in "OCLR_NONE=1" "OCLR_LIB=$P/libopencl_render_b200_out1.so" "OCLR_LIB=$P/libopencl_render_b200_out2.so" "OCLR_LIB=$P/libopencl_render_b200_out3.so" "OCLR_LIB=$P/libopencl_render_b200_out3c6.so"; do
  echo "== $v"
  env $v OCLR_NO_BUILD=1 timeout 120 python scripts/ncu_target_band.py 2 1 8 2>&1 | tail -1
  env $v OCLR_NO_BUILD=1 timeout 120 python scripts/ncu_target_band.py 2 8 8 2>&1 | tail -1
  env $v OCLR_NO_BUILD=1 timeout 120 python scripts/ncu_target_band.py 3 1 4 2>&1 | tail -1
  env $v OCLR_NO_BUILD=1 timeout 120 python scripts/ncu_target_band.py 5 1 4 2>&1 | tail -1
done ) > gpurun_out/r02o_outline.log 2>&1
grep -E "==|BEST" gpurun_out/r02o_outline.log
for l in 1 3; do OCLR_LIB=$P/libopencl_render_b200_out$l.so OCLR_NO_BUILD=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 -k "golden or whole_frame" 2>&1 | tail -2; done
