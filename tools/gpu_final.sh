#!/bin/bash
# Final GPU visit of a round: whole test suite, bench (both arms, all workloads), in-graph group costs, step time vs batch,
# launch list and one `ncu --set full` capture of the dominant kernels (each ncu run behind a plain run of the same command).
TAG=${1:-r02f}
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
tail -4 $O/${TAG}_pytest.log
timeout 900 python bench.py > $O/${TAG}_bench_qm9.json 2> $O/${TAG}_bench_qm9.err; echo "bench rc=$?"; tail -2 $O/${TAG}_bench_qm9.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; echo "ref rc=$?"; cat $O/${TAG}_bench_ref.json
for w in mp2018 fullerene ptgp; do
  timeout 600 python bench.py --workload $w --no-cpu > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err; tail -2 $O/${TAG}_bench_$w.err
done
timeout 600 python tools/skip_time.py > $O/${TAG}_skip_time.log 2>&1; cat $O/${TAG}_skip_time.log
timeout 600 python tools/floor_time.py > $O/${TAG}_floor_time.log 2>&1; cat $O/${TAG}_floor_time.log
timeout 300 python tools/chain_time.py qm9 > $O/${TAG}_chain_time.log 2>&1; grep "first 5\|first 4\|first 3 steps (5" $O/${TAG}_chain_time.log
timeout 300 python tools/la_kernel_time.py 128 > $O/${TAG}_la_kernel_time.log 2>&1; timeout 300 python tools/la_kernel_time.py 512 >> $O/${TAG}_la_kernel_time.log 2>&1; cat $O/${TAG}_la_kernel_time.log
python tools/infer_loop.py 128 3 train > $O/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches.csv python tools/infer_loop.py 128 3 train > $O/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"; python tools/launch_summary.py $O/${TAG}_launches.csv 3 > $O/${TAG}_launches.md; cat $O/${TAG}_launches.md
python tools/infer_loop.py 128 2 train > $O/${TAG}_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'la_.*_pipe|dense_chain2|wgrad_batch' -s 30 -c 14 -o $O/${TAG}_prof python tools/infer_loop.py 128 2 train > $O/${TAG}_ncu2.log 2>&1
echo "ncu full rc=$?"; tail -2 $O/${TAG}_ncu2.log
