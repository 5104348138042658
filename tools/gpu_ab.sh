#!/bin/bash
# One GPU-box visit: whole GPU test suite, A/B of library builds under scann_b200/ab/ (train / inference step, CUDA events),
# default bench line.   usage: tools/gpu_ab.sh TAG "old A B" [workloads]
TAG=${1:-rXX}
LIBS=${2:-"old B"}
WLS=${3:-"qm9 mp2018"}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
tail -3 $O/${TAG}_pytest.log
for rep in 1 2; do
  for wl in $WLS; do
    for l in $LIBS; do
      echo -n "$l " | tee -a $O/${TAG}_ab.log
      timeout 300 python tools/ab_time.py $wl SCANN_B200_LIB=$PWD/scann_b200/ab/$l.so 2>&1 | tee -a $O/${TAG}_ab.log
    done
  done
done
timeout 900 python bench.py > $O/${TAG}_bench_qm9.json 2> $O/${TAG}_bench_qm9.err; echo "bench rc=$?"; tail -2 $O/${TAG}_bench_qm9.err
cat $O/${TAG}_bench_qm9.json
