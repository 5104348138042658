#!/bin/bash
# chain2 visit: phase clocks / launch times (dev-probe build), parity of the golden / variant tests, A/B step time
TAG=${1:-r02u}
O=gpurun_out
mkdir -p $O
timeout 300 python tools/chain_time.py qm9 > $O/${TAG}_chain_time.log 2>&1; cat $O/${TAG}_chain_time.log
for c in 0 1; do SCANN_CHAIN2=$c python tools/golden_err.py 2>&1 | sed "s/^/chain2=$c /"; done | tee $O/${TAG}_golden_err.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "golden or forward_matches or gradients or dropout" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
tail -15 $O/${TAG}_pytest.log
timeout 600 python tools/ab_time.py qm9 SCANN_CHAIN2=0 SCANN_CHAIN2=1 > $O/${TAG}_ab.log 2>&1; cat $O/${TAG}_ab.log
