set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l; nproc
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 --timeout 600 > gpurun_out/r02k_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02k_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02k_bench_n1.json 2> gpurun_out/r02k_bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/r02k_bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r02k_bench_n1.json'))
print('N=1 value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'issue', d['roofline'].get('issue',{}).get('frac'), 'parity', d['parity'], 'e2e', {k: d['e2e'].get(k) for k in ('value','ms_per_call','ms_median','spread','value_pageable_host_arrays','error')})
for k, v in d['configs'].items(): print(k, {a: v.get(a) for a in ('value', 'ms_per_step', 'parity', 'frames_per_s', 'error')}, {a: (v.get('e2e') or {}).get(a) for a in ('value','ms_per_call','value_pageable_host_arrays','error')})
print(d['cpu_baseline'])
"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02k_bench_ref.json 2> gpurun_out/r02k_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/r02k_bench_ref.json | cut -c1-600
