# round 2, closing session: the whole GPU suite, smoke() and a short bench line on the final tree
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/r02end_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02end_tests.log
python -c "import __graft_entry__ as g; g.smoke()"
OCLR_BENCH_NO_E2E=1 timeout 600 python bench.py --steps 10 --warmup 3 --extra none > gpurun_out/r02end_bench.json 2> gpurun_out/r02end_bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r02end_bench.json
