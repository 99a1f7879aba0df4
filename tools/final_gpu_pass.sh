#!/bin/bash
# One GPU pass that produces every number / profile of a round (run through gpurun; outputs under gpurun_out/).
# usage: tools/final_gpu_pass.sh <round tag>
TAG=${1:-r02}
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/${TAG}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.txt 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference_n1.json 2>/dev/null
python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
for w in cfg3 cfg4 sinc cfg5shard cfg5; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 > $O/${TAG}_bench_$w.json 2>/dev/null
done
for w in cfg2 cfg5shard; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_$w.csv \
      python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:skeleton_kernel -s 1 -c 1 -f -o $O/${TAG}_skeleton_cfg2 \
    python bench.py --workload cfg2 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
for w in cfg2 cfg5shard; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:replay_kernel -s 5 -c 1 -f -o $O/${TAG}_replay_$w \
      python bench.py --workload $w --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mix_fx_kernel -s 5 -c 1 -f -o $O/${TAG}_mixfx_cfg2 \
    python bench.py --workload cfg2 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
PB200_NO_PERSISTENT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:skeleton_kernel -s 3 -c 1 -f -o $O/${TAG}_skeleton_cfg5shard \
    python bench.py --workload cfg5shard --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
# the reports are too big to travel back (64 MiB limit): keep their raw pages as CSV
for rep in $O/${TAG}_*.ncu-rep; do
  ncu -i $rep --page raw --csv > ${rep%.ncu-rep}_raw.csv 2>/dev/null
  rm -f $rep
done
cat $O/${TAG}_pytest_gpu.txt $O/${TAG}_smoke.txt
