set -x
cd $GRAFT_REPO_ROOT
( for cfgs in "0 16 0" "48 12 0" "64 16 0" "96 16 0" "96 32 0" "160 32 0"; do set -- $cfgs
  echo "== split_min=$1 part=$2 early=$3"
  OCLR_SPLIT_MIN=$1 OCLR_SPLIT_PART=$2 OCLR_SPLIT_EARLY=$3 timeout 120 python scripts/ncu_target_band.py 2 8 8 2>&1 | tail -2
  OCLR_SPLIT_MIN=$1 OCLR_SPLIT_PART=$2 OCLR_SPLIT_EARLY=$3 timeout 120 python scripts/ncu_target_band.py 2 1 6 2>&1 | tail -2
  OCLR_SPLIT_MIN=$1 OCLR_SPLIT_PART=$2 OCLR_SPLIT_EARLY=$3 timeout 120 python scripts/ncu_target_band.py 3 8 6 2>&1 | tail -2
done ) > gpurun_out/r02p_split2.log 2>&1
grep -E "==|BEST|split:" gpurun_out/r02p_split2.log
OCLR_SPLIT_MIN=64 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 -k "golden or whole_frame" 2>&1 | tail -2
