#!/bin/bash
# where does the error against the fp64 goldens come from: engine variants
TAG=${1:-r02z}
O=gpurun_out
for v in "SCANN_CHAIN2=1" "SCANN_CHAIN2=0" "SCANN_ENGINE=simt" "SCANN_LA_FWD=simt" "SCANN_DENSE=simt" "SCANN_LA_PIPE=0" "SCANN_LA_PIPE=0 SCANN_TILE_STRIDE=64"; do
  env $v python tools/golden_err.py 2>&1 | grep -v Warning | sed "s/^/[$v] /"
done | tee $O/${TAG}_golden_err.log
