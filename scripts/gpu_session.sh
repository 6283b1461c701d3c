# round 2, two-GPU sanity session on the final code: the tests that need two GPUs, and the bench line at N = 2
set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 -k "all_devices or two_gpus or raytrace_all" > gpurun_out/r02c_tests_n2.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02c_tests_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02c_bench_n2.json 2> gpurun_out/r02c_bench_n2.err; echo "bench rc=$?"; cut -c1-1200 gpurun_out/r02c_bench_n2.json; tail -3 gpurun_out/r02c_bench_n2.err
