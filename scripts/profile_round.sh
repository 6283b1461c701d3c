set -x
python bench.py --steps 2 --warmup 3 > gpurun_out/r01e_bench_plain.json 2> gpurun_out/r01e_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01e_bench_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r01e_ncu_bench.log 2>&1
python scripts/ncu_target.py 2 2 1 > gpurun_out/r01e_target_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:wf_pipe -c 3 -f -o gpurun_out/r01e_wf_pipe_cfg2 python scripts/ncu_target.py 2 2 1 > gpurun_out/r01e_ncu_pipe.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:wf_setup|wf_logic" -c 7 -f -o gpurun_out/r01e_logic_setup_cfg2 python scripts/ncu_target.py 2 2 1 > gpurun_out/r01e_ncu_logic.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
