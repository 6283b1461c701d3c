set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L; nproc
timeout 300 python bench.py --steps 10 --warmup 3 --extra none > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('gpurun_out/r02b_bench_n1.json')); print('N=1 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'])"
OCLR_TRACE=1 timeout 200 python scripts/e2e_trace.py 2 1 10 > gpurun_out/r02b_e2e_trace_1.log 2>&1; tail -4 gpurun_out/r02b_e2e_trace_1.log
# split policies (SPLIT instantiation): late only / early, on the whole frame and on a 1/8 share
( for cfgs in "0 16 0" "64 16 0" "64 16 2" "64 16 6" "128 16 2" "128 32 4" "96 24 12"; do set -- $cfgs
  echo "== split_min=$1 part=$2 early=$3"
  OCLR_SPLIT_MIN=$1 OCLR_SPLIT_PART=$2 OCLR_SPLIT_EARLY=$3 timeout 120 python scripts/ncu_target_band.py 2 1 6 2>&1 | tail -2
  OCLR_SPLIT_MIN=$1 OCLR_SPLIT_PART=$2 OCLR_SPLIT_EARLY=$3 timeout 120 python scripts/ncu_target_band.py 2 8 6 2>&1 | tail -2
done ) > gpurun_out/r02b_split_ab.log 2>&1
grep -E "==|BEST|split:" gpurun_out/r02b_split_ab.log
OCLR_SPLIT_MIN=64 OCLR_SPLIT_EARLY=2 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 -k "golden or whole_frame" > gpurun_out/r02b_split_tests.log 2>&1; echo "early split pytest rc=$?"; tail -3 gpurun_out/r02b_split_tests.log
# ---- two GPUs: RaytraceAll(all devices) with the shared upload; the multi-GPU tests; the bench under torchrun
OCLR_TRACE=1 timeout 200 python scripts/e2e_trace.py 2 2 10 > gpurun_out/r02b_e2e_trace_2.log 2>&1; echo "e2e x2 rc=$?"; tail -30 gpurun_out/r02b_e2e_trace_2.log
timeout 300 python -m pytest tests/test_gpu_full_configs.py -m gpu -q --timeout 200 -k "two_gpus or all_devices" > gpurun_out/r02b_multi_tests.log 2>&1; echo "multi pytest rc=$?"; tail -5 gpurun_out/r02b_multi_tests.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02b_bench_n2.json 2> gpurun_out/r02b_bench_n2.err; echo "bench2 rc=$?"
tail -5 gpurun_out/r02b_bench_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r02b_bench_n2.json'))
print('N=2 value', d['value'], 'ms', d['ms_per_step'], 'parity', d['parity'], 'e2e', d['e2e'])
for k, v in d['configs'].items(): print(k, {a: v.get(a) for a in ('value', 'ms_per_step', 'parity', 'e2e', 'frames_per_s', 'error')})
"
