set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L; nproc
for v in "OCLR_NONE=1" "OCLR_SHARD_VERIFY=1"; do
  echo "== cfg3 x2 with $v"
  env $v timeout 300 python -m opencl_render_b200.e2e_probe 3 2 8 0 > gpurun_out/r02e_cfg3_x2_$v.log 2>&1; echo "rc=$?"
  grep -E "Error|error|ms_per_call|differ, first" gpurun_out/r02e_cfg3_x2_$v.log | cut -c1-400 | head -5
done
OCLR_TRACE=1 timeout 200 python scripts/e2e_trace.py 2 2 10 > gpurun_out/r02e_e2e_trace_2.log 2>&1; grep -E "x2:" gpurun_out/r02e_e2e_trace_2.log; grep "RaytraceAll dev" gpurun_out/r02e_e2e_trace_2.log | sed -n 9,12p; grep "RaytraceAll dev" gpurun_out/r02e_e2e_trace_2.log | tail -2
for kb in 128 256 512 1024; do echo "== staging block $kb KB"; OCLR_STAGING_BLOCK_KB=$kb timeout 200 python scripts/e2e_trace.py 2 1 10 2>&1 | grep -E "pageable x1" ; done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02e_bench_n2.json 2> gpurun_out/r02e_bench_n2.err; echo "bench2 rc=$?"
tail -3 gpurun_out/r02e_bench_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r02e_bench_n2.json'))
print('N=2 value', d['value'], 'ms', d['ms_per_step'], 'parity', d['parity'], 'e2e', {k: d['e2e'].get(k) for k in ('value','ms_per_call','spread','value_pageable_host_arrays','error')})
for k, v in d['configs'].items(): print(k, {a: v.get(a) for a in ('value', 'ms_per_step', 'parity', 'frames_per_s', 'error')}, {a: (v.get('e2e') or {}).get(a) for a in ('value','ms_per_call','value_pageable_host_arrays','error')})
"
