set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l; nproc
timeout 600 python -m pytest tests/test_gpu_full_configs.py tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "all_devices or two_gpus or e2e_probe or bad_caller or by_value" > gpurun_out/r02n_tests.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02n_tests.log
timeout 300 python scripts/cache_soak.py 2 48 2>&1 | grep -v "RaytraceAll dev" | tail -12
timeout 300 python scripts/cache_soak.py 1 32 2>&1 | tail -6
bash scripts/gpu_session_b.sh
