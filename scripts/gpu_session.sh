set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l; nproc
for n in 2 4 8; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r02m_bench_n$n.json 2> gpurun_out/r02m_bench_n$n.err; echo "bench$n rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02m_bench_n$n.json'))
print('N=$n value', d['value'], 'ms', d['ms_per_step'], 'parity', d['parity'], 'e2e', {k: d['e2e'].get(k) for k in ('value','ms_per_call','ms_median','spread','value_pageable_host_arrays','error')})
for k, v in d['configs'].items(): print(k, {a: v.get(a) for a in ('value', 'ms_per_step', 'parity', 'frames_per_s', 'error')}, {a: (v.get('e2e') or {}).get(a) for a in ('value','ms_per_call','value_pageable_host_arrays','error')})
"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus 8 --steps 5 --warmup 1 > gpurun_out/r02m_bench_ref_n8.json 2> gpurun_out/r02m_bench_ref_n8.err; echo "ref8 rc=$?"; cut -c1-300 gpurun_out/r02m_bench_ref_n8.json
