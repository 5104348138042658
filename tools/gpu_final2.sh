#!/bin/bash
# Final GPU visit (short form of gpu_final.sh): whole GPU test suite, bench (both arms, all workloads), in-graph group costs,
# launch list of the train step (ncu run behind a plain run of the same command).
TAG=${1:-r02z}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
tail -3 $O/${TAG}_pytest.log
timeout 600 python bench.py > $O/${TAG}_bench_qm9.json 2> $O/${TAG}_bench_qm9.err; echo "bench rc=$?"; tail -2 $O/${TAG}_bench_qm9.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; echo "ref rc=$?"; cat $O/${TAG}_bench_ref.json
for w in mp2018 fullerene ptgp; do
  timeout 400 python bench.py --workload $w --no-cpu > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err; echo "$w rc=$?"; tail -2 $O/${TAG}_bench_$w.err
done
timeout 400 python tools/skip_time.py > $O/${TAG}_skip_time.log 2>&1; cat $O/${TAG}_skip_time.log
python tools/infer_loop.py 128 3 train > $O/${TAG}_plain.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches.csv python tools/infer_loop.py 128 3 train > $O/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"; python tools/launch_summary.py $O/${TAG}_launches.csv 3 > $O/${TAG}_launches.md; head -24 $O/${TAG}_launches.md
python tools/wl_loop.py ptgp 2 train > $O/${TAG}_ptgp_plain.log 2>&1 && \
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches_ptgp.csv python tools/wl_loop.py ptgp 2 train > $O/${TAG}_ncu_ptgp.log 2>&1
python tools/launch_summary.py $O/${TAG}_launches_ptgp.csv 2 > $O/${TAG}_launches_ptgp.md; head -14 $O/${TAG}_launches_ptgp.md
