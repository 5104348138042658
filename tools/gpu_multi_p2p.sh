#!/bin/bash
# multi-GPU visit (gpurun --gpus N): data-parallel correctness + the bench with the peer-memory exchange (the default)
N=${1:-4}; TAG=${2:-r02m}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/dp_check.py > $O/${TAG}_dp_check_${N}gpu.log 2>&1; echo "dp_check rc=$?"; grep -E "world|DP_CHECK" $O/${TAG}_dp_check_${N}gpu.log
for rep in 1 2; do
timeout 900 $TR --master-port 2951$rep bench.py --gpus $N --no-cpu --steps 30 > $O/${TAG}_bench_${N}gpu_p2p.json 2> $O/${TAG}_bench_${N}gpu_p2p.err; echo "bench p2p rc=$?"
python - <<PY
import json
try:
    d = json.load(open("$O/${TAG}_bench_${N}gpu_p2p.json"))
    print("p2p qm9 ms/step", round(d["ms_per_step"], 4), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "| mp2018 ms/step", round(d["mp2018"]["ms_per_step"], 4), "value", round(d["mp2018"]["value"]), "e2e", round(d["mp2018"]["e2e"]["value"]))
except Exception as e:
    print("failed", e)
PY
done
