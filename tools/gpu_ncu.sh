#!/bin/bash
# launch list + one ncu --set full capture of the pipelined kernels (inference loop, no graphs)
TAG=${1:-r02n}; B=${2:-128}; MODE=${3:-infer}; KRE=${4:-la_.*_pipe}
O=gpurun_out
mkdir -p $O
python tools/infer_loop.py $B 3 $MODE > $O/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches.csv python tools/infer_loop.py $B 3 $MODE > $O/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"; tail -3 $O/${TAG}_ncu1.log
python tools/infer_loop.py $B 3 $MODE > $O/${TAG}_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s 4 -c 4 -o $O/${TAG}_prof python tools/infer_loop.py $B 3 $MODE > $O/${TAG}_ncu2.log 2>&1
echo "ncu full rc=$?"; tail -3 $O/${TAG}_ncu2.log
ls -la $O | grep ${TAG}
