#!/bin/bash
# Round evidence in one GPU call: full GPU test suite, the driver's bench command, the reference arm, the ncu launch list
# of the same bench command.  usage: bash tools/evidence.sh <tag>   (outputs under gpurun_out/<tag>_*)
tag=${1:-r2c}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${tag}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${tag}_bench.log 2>gpurun_out/${tag}_bench.err; tail -c 600 gpurun_out/${tag}_bench.log
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${tag}_bench_reference.log 2>gpurun_out/${tag}_bench_reference.err; tail -c 400 gpurun_out/${tag}_bench_reference.log
# (the ncu pass serialises ~5 300 launches: about 15 minutes of box time - skip it with NO_NCU=1 when the budget is short)
[ -n "$NO_NCU" ] && exit 0
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch-obs 0 > gpurun_out/${tag}_ncu_bench.log 2>&1
grep -c . gpurun_out/${tag}_launches.csv
