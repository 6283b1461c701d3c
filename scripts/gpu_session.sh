# round 2, session m: slices whose trace grids fit the SM together (OCLR_SLICES x OCLR_TRACE_CTAS <= 8): do the tails of one slice's
# trace launches overlap the other slice's logic / trace on a 1/8 share?
set -x
cd $GRAFT_REPO_ROOT
( timeout 400 python scripts/share_sweep.py 2 8 "OCLR_X=default" "OCLR_SLICES=2 OCLR_TRACE_CTAS=4" "OCLR_SLICES=2 OCLR_TRACE_CTAS=5" "OCLR_SLICES=2 OCLR_TRACE_CTAS=6" "OCLR_SLICES=3 OCLR_TRACE_CTAS=3" "OCLR_SLICES=4 OCLR_TRACE_CTAS=2" "OCLR_SLICES=2" "OCLR_X=default"
  timeout 300 python scripts/share_sweep.py 2 4 "OCLR_X=default" "OCLR_SLICES=2 OCLR_TRACE_CTAS=4" "OCLR_SLICES=2 OCLR_TRACE_CTAS=5"
  timeout 300 python scripts/share_sweep.py 3 8 "OCLR_X=default" "OCLR_SLICES=2 OCLR_TRACE_CTAS=4" "OCLR_SLICES=2 OCLR_TRACE_CTAS=5"
  timeout 300 python scripts/share_sweep.py 2 1 "OCLR_X=default" "OCLR_SLICES=2 OCLR_TRACE_CTAS=4" "OCLR_SLICES=2 OCLR_TRACE_CTAS=5" ) 2>&1 | tee gpurun_out/r02m_slices.log
