#!/bin/bash
# One GPU-box visit: parity tests, bench (both arms), launch list and one ncu --set full capture.
# usage: tools/gpu_round.sh TAG [quick]      (outputs under gpurun_out/TAG_*)
TAG=${1:-rXX}
MODE=${2:-full}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
tail -3 $O/${TAG}_pytest.log
python bench.py > $O/${TAG}_bench_qm9.json 2> $O/${TAG}_bench_qm9.err; echo "bench rc=$?"
cat $O/${TAG}_bench_qm9.json
python tools/floor_time.py > $O/${TAG}_floor.log 2>&1; cat $O/${TAG}_floor.log
[ "$MODE" = quick ] && exit 0
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; echo "ref rc=$?"
cat $O/${TAG}_bench_ref.json
for w in mp2018 fullerene ptgp; do
  python bench.py --workload $w --no-cpu > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err; cat $O/${TAG}_bench_$w.json
done
python tools/dbg_clocks.py > $O/${TAG}_dbg_clocks.log 2>&1; cat $O/${TAG}_dbg_clocks.log
# launch list (graphs off so that every kernel is its own launch); plain run of the same command first
SCANN_GRAPHS=0 python bench.py --steps 2 --warmup 3 --no-cpu > $O/${TAG}_plain.log 2>&1 && \
SCANN_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 290 -c 150 --csv \
  --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
SCANN_GRAPHS=0 ncu --set full --clock-control none --import-source on -k regex:'la_.*_tc|wgrad_batch|dense_chain|plan_chain|geom_init' -s 60 -c 36 \
  -o $O/${TAG}_la_tc python bench.py --steps 2 --warmup 3 --no-cpu > $O/${TAG}_ncu2.log 2>&1
echo "ncu full rc=$?"
ls -la $O | tail -20
