# round 2, session q: band height of the multi-GPU partition -- every rank's share timed on one GPU (the N-GPU step is the slowest rank's)
set -x
cd $GRAFT_REPO_ROOT
( timeout 300 python scripts/band_probe.py 2 8 16,8,32,24
  timeout 300 python scripts/band_probe.py 2 4 16,8,32
  timeout 300 python scripts/band_probe.py 3 8 16,8,32 ) 2>&1 | tee gpurun_out/r02q_bands.log
