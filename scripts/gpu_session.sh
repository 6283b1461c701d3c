set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l; nproc
python scripts/sweep_env.py 2 2 "OCLR_NONE=1" "OCLR_REFILL_MIN=2" "OCLR_REFILL_MIN=8" "OCLR_REFILL_MIN=12" "OCLR_DRAIN_MIN=32" "OCLR_DRAIN_MIN=64" "OCLR_WALK_MIN3=4" "OCLR_WALK_MIN3=12" "OCLR_WALK_MIN3=16" "OCLR_SWITCH_MIN=3" "OCLR_SWITCH_MIN=10" "OCLR_SWITCH_MIN=16" "OCLR_TAIL_DRAIN=4" "OCLR_TAIL_DRAIN=16" "OCLR_TRACE_CTAS=6" "OCLR_TRACE_CTAS=7" > gpurun_out/r02i_knobs.log 2>&1
grep -E "^---|variant 2:" gpurun_out/r02i_knobs.log
bash scripts/profile_round.sh r02 > gpurun_out/r02_profile_round.log 2>&1; tail -5 gpurun_out/r02_profile_round.log
