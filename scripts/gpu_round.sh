#!/bin/bash
# One GPU-box visit: smoke, gpu tests, bench, ncu launch list + full capture.  Everything lands in gpurun_out/.
# usage: scripts/gpu_round.sh [tag]
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
nproc >> $OUT/gpu.txt; lscpu | grep "Model name" >> $OUT/gpu.txt
echo "== build+smoke"; timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2>> $OUT/bench.err; cat $OUT/bench_ref.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > $OUT/ncu_launches.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:siddon -s 1 -c 3 -o $OUT/prof_siddon \
    python bench.py --steps 1 --warmup 1 --batch 16 --no-cpu-baseline --no-graph > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT
