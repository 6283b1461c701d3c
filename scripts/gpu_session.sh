set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L; nproc
python scripts/opencl_probe.py > gpurun_out/r02_opencl_probe.log 2>&1
export OCLR_SPLIT_MIN=0
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 --timeout 600 --durations=15 > gpurun_out/r02a_gpu_tests.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/r02a_gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err; echo "bench rc=$?"
tail -5 gpurun_out/r02a_bench_n1.err; cat gpurun_out/r02a_bench_n1.json
# run-time split of long walks: parity first (short leash), then A/B on the whole frame and on a 1/8 band share
export OCLR_SPLIT_MIN=48
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 -k "golden or whole_frame or sliced or ahead" > gpurun_out/r02a_split_tests.log 2>&1; echo "split pytest rc=$?"
tail -15 gpurun_out/r02a_split_tests.log
for s in 0 32 48 96; do for p in 8 16 32; do
  [ $s = 0 ] && [ $p != 16 ] && continue
  echo "== split_min=$s part=$p"; OCLR_SPLIT_MIN=$s OCLR_SPLIT_PART=$p timeout 120 python scripts/ncu_target_band.py 2 1 6 2>&1 | tail -1
  OCLR_SPLIT_MIN=$s OCLR_SPLIT_PART=$p timeout 120 python scripts/ncu_target_band.py 2 8 6 2>&1 | tail -1
done; done > gpurun_out/r02a_split_ab.log 2>&1
cat gpurun_out/r02a_split_ab.log
