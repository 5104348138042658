#!/bin/bash
# quick GPU visit: pipelined-kernel tests + step time vs batch
TAG=${1:-r02q}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "pipelined or golden" > $O/${TAG}_pytest_pipe.log 2>&1; echo "pipe rc=$?" | tee -a $O/${TAG}_pytest_pipe.log
tail -8 $O/${TAG}_pytest_pipe.log
timeout 600 python tools/floor_time.py > $O/${TAG}_floor.log 2>&1; cat $O/${TAG}_floor.log
