# round 2, closing two-GPU check of the final tree: the tests that need two GPUs
set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 -k "all_devices or two_gpus or raytrace_all" 2>&1 | tail -3
