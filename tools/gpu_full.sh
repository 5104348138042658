#!/bin/bash
# full GPU visit: whole test suite, bench (both arms), other workloads
TAG=${1:-r02x}
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
tail -30 $O/${TAG}_pytest.log
timeout 900 python bench.py > $O/${TAG}_bench_qm9.json 2> $O/${TAG}_bench_qm9.err; echo "bench rc=$?"; cat $O/${TAG}_bench_qm9.json; tail -3 $O/${TAG}_bench_qm9.err
[ "$2" = quick ] && exit 0
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; echo "ref rc=$?"; cat $O/${TAG}_bench_ref.json
for w in mp2018 fullerene ptgp; do
  timeout 600 python bench.py --workload $w --no-cpu > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err; cat $O/${TAG}_bench_$w.json; tail -2 $O/${TAG}_bench_$w.err
done
