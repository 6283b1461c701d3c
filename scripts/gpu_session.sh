# round 2, final session on one GPU: the whole GPU suite, the default bench line, the reference arm, and the round's ncu evidence (r02c)
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/r02c_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02c_tests.log
timeout 900 python bench.py > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err; echo "bench rc=$?"; cut -c1-900 gpurun_out/r02c_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02c_bench_ref.json 2> gpurun_out/r02c_bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r02c_bench_ref.json
timeout 900 bash scripts/profile_round.sh r02c
python -c "import __graft_entry__ as g; g.smoke()"
