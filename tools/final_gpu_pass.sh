#!/bin/bash
# One GPU pass that produces every number / profile of a round (run through gpurun; outputs under gpurun_out/).
# usage: tools/final_gpu_pass.sh <round tag>
TAG=${1:-r01f}
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/${TAG}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.txt 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference_n1.json 2>/dev/null
python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
for w in cfg3 cfg4 sinc cfg5shard; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 > $O/${TAG}_bench_$w.json 2>/dev/null
done
for w in cfg2 cfg5shard; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_$w.csv \
      python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
done
for w in cfg2 cfg5shard; do
  ncu --set full --clock-control none --import-source on -k regex:replay_kernel -s 5 -c 1 -f -o $O/${TAG}_replay_$w \
      python bench.py --workload $w --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:mix_fx_kernel -s 5 -c 1 -f -o $O/${TAG}_mixfx_cfg2 \
    python bench.py --workload cfg2 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
cat $O/${TAG}_pytest_gpu.txt $O/${TAG}_smoke.txt
