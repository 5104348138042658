#!/bin/bash
# Round 2, first GPU visit: new pipelined forward kernels, full-configuration parity, tf32 issue rates, sanitizer.
TAG=${1:-r02a}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.log 2>&1
echo "== pipelined kernel tests"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "pipelined" > $O/${TAG}_pytest_pipe.log 2>&1; echo "pipe rc=$?" | tee -a $O/${TAG}_pytest_pipe.log
tail -25 $O/${TAG}_pytest_pipe.log
echo "== full suite (default engine)"
timeout 1500 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; RC=$?; echo "pytest rc=$RC" | tee -a $O/${TAG}_pytest.log
tail -40 $O/${TAG}_pytest.log
if [ $RC -ne 0 ]; then
  echo "== full suite with the round-1 kernels (SCANN_LA_PIPE=0, 64-row plan)"
  SCANN_LA_PIPE=0 timeout 1500 python -m pytest tests -m gpu -q -k "not pipelined" > $O/${TAG}_pytest_old.log 2>&1; echo "old rc=$?" | tee -a $O/${TAG}_pytest_old.log
  tail -15 $O/${TAG}_pytest_old.log
fi
echo "== tcgen05 issue rates"
timeout 300 python tools/tc_time.py > $O/${TAG}_tc_time.log 2>&1; cat $O/${TAG}_tc_time.log
echo "== step time vs batch, pipelined forward / round-1 kernels on the 32-row plan / round-1 64-row plan"
timeout 600 python tools/floor_time.py > $O/${TAG}_floor_pipe.log 2>&1; cat $O/${TAG}_floor_pipe.log
SCANN_LA_PIPE=0 SCANN_TILE_STRIDE=32 timeout 600 python tools/floor_time.py > $O/${TAG}_floor_old32.log 2>&1; cat $O/${TAG}_floor_old32.log
SCANN_LA_PIPE=0 timeout 600 python tools/floor_time.py > $O/${TAG}_floor_old64.log 2>&1; cat $O/${TAG}_floor_old64.log
echo "== bench"
timeout 900 python bench.py > $O/${TAG}_bench_qm9.json 2> $O/${TAG}_bench_qm9.err; echo "bench rc=$?"; cat $O/${TAG}_bench_qm9.json; tail -5 $O/${TAG}_bench_qm9.err
echo "== compute-sanitizer memcheck on the smoke case"
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py --smoke > $O/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -15 $O/${TAG}_memcheck.log
