# The round's ncu evidence (run through gpurun on ONE GPU; every ncu command only after the same command exited 0 without ncu).
# usage: bash scripts/profile_round.sh r02
TAG=${1:-r02}
set -x
OCLR_BENCH_NO_E2E=1 python bench.py --steps 2 --warmup 3 --extra none > gpurun_out/${TAG}_bench_plain.json 2> gpurun_out/${TAG}_bench_plain.err || exit 1
OCLR_BENCH_NO_E2E=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_bench_launches.csv python bench.py --steps 2 --warmup 3 --extra none > gpurun_out/${TAG}_ncu_bench.log 2>&1
python scripts/ncu_target.py 2 2 1 > gpurun_out/${TAG}_target_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:wf_pipe -c 3 -f -o gpurun_out/${TAG}_wf_pipe_cfg2 python scripts/ncu_target.py 2 2 1 > gpurun_out/${TAG}_ncu_pipe.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:wf_setup|wf_logic" -c 7 -f -o gpurun_out/${TAG}_logic_setup_cfg2 python scripts/ncu_target.py 2 2 1 > gpurun_out/${TAG}_ncu_logic.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
