set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l; nproc
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02f_bench_n8.json 2> gpurun_out/r02f_bench_n8.err; echo "bench8 rc=$?"
tail -3 gpurun_out/r02f_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r02f_bench_n4.json 2> gpurun_out/r02f_bench_n4.err; echo "bench4 rc=$?"
tail -3 gpurun_out/r02f_bench_n4.err
for n in 8 4; do python -c "
import json; d=json.load(open('gpurun_out/r02f_bench_n$n.json'))
print('N=$n value', d['value'], 'ms', d['ms_per_step'], 'parity', d['parity'], 'e2e', {k: d['e2e'].get(k) for k in ('value','ms_per_call','spread','value_pageable_host_arrays','error')})
for k, v in d['configs'].items(): print(k, {a: v.get(a) for a in ('value', 'ms_per_step', 'parity', 'frames_per_s', 'error')}, {a: (v.get('e2e') or {}).get(a) for a in ('value','ms_per_call','value_pageable_host_arrays','error')})
"; done
OCLR_TRACE=1 timeout 200 python scripts/e2e_trace.py 2 8 10 > gpurun_out/r02f_e2e_trace_8.log 2>&1; grep -E "x8:" gpurun_out/r02f_e2e_trace_8.log; grep "RaytraceAll dev" gpurun_out/r02f_e2e_trace_8.log | sed -n 41,48p
