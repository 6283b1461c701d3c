# round 2, eight-GPU run of the final code: the bench line at N = 8 (config 4 -- 10 M triangles, 7680x4320 -- is measured only here)
set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02c_bench_n8.json 2> gpurun_out/r02c_bench_n8.err; echo "bench rc=$?"; cut -c1-700 gpurun_out/r02c_bench_n8.json; tail -3 gpurun_out/r02c_bench_n8.err
