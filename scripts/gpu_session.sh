set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l; nproc
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 --timeout 600 > gpurun_out/r02g_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02g_gpu_tests.log
OCLR_TRACE=1 timeout 200 python scripts/e2e_trace.py 2 1 12 > gpurun_out/r02g_e2e_trace_1.log 2>&1; grep -E "x1:" gpurun_out/r02g_e2e_trace_1.log; grep "RaytraceAll dev" gpurun_out/r02g_e2e_trace_1.log | sed -n 10,12p; grep "RaytraceAll dev" gpurun_out/r02g_e2e_trace_1.log | tail -2
OCLR_TRACE=1 timeout 200 python scripts/e2e_trace.py 2 2 12 > gpurun_out/r02g_e2e_trace_2.log 2>&1; grep -E "x2:" gpurun_out/r02g_e2e_trace_2.log; grep "RaytraceAll dev" gpurun_out/r02g_e2e_trace_2.log | sed -n 19,22p
OCLR_BLOCK_CACHE=0 timeout 200 python scripts/e2e_trace.py 2 2 12 2>&1 | grep -E "x2:"
timeout 200 python scripts/e2e_trace.py 3 2 8 2>&1 | grep -E "x2:"
timeout 200 python scripts/e2e_trace.py 5 2 8 2>&1 | grep -E "x2:"
