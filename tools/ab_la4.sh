#!/bin/bash
# A/B of the four-warp-group local-attention kernels: parity tests with the mask on, then step time per mask.
# usage: tools/ab_la4.sh MASK [TAG]
MASK=${1:-15}
TAG=${2:-ab}
O=gpurun_out
mkdir -p $O
SCANN_LA4=$MASK python -m pytest tests/test_gpu_parity.py -x -q -k "golden or per_tensor or dropout or edge_cases or full_size" 2>&1 | tail -4
for m in 0 $MASK; do
  echo "== SCANN_LA4=$m"
  SCANN_LA4=$m python tools/floor_time.py 2>&1 | grep "B= 128\|B= 512\|B=  64"
done
SCANN_LA4=$MASK python tools/dbg_clocks.py 2>&1 | head -9
