#!/bin/bash
# launch lists (per-kernel durations, cold cache) of a train step: pipelined kernels vs the round-1 kernels on the same plan
TAG=${1:-r02l}; B=${2:-128}
O=gpurun_out
mkdir -p $O
for P in 15 0; do
SCANN_LA_PIPE=$P python tools/infer_loop.py $B 3 train > $O/${TAG}_plain_$P.log 2>&1 && \
SCANN_LA_PIPE=$P ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches_pipe$P.csv python tools/infer_loop.py $B 3 train > $O/${TAG}_ncu_$P.log 2>&1
echo "launch list pipe=$P rc=$?"
python tools/launch_summary.py $O/${TAG}_launches_pipe$P.csv 3 | head -16
done
