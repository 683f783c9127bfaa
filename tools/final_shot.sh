#!/bin/bash
# Last GPU call of round 2 (2 GPU-minutes left): memcheck + racecheck of the kernels added after r2c_sanitizer.log (grouped
# decode attention, cached state key), then the whole GPU suite at the final commit.  Outputs under gpurun_out/r2o_*.
mkdir -p gpurun_out
L=gpurun_out/r2o_sanitizer.log
: > $L
K='grouped or cached_state or denoise_attention_tcgen05'
echo "== compute-sanitizer --tool memcheck python -m pytest tests/test_attention_gpu.py -m gpu -q -x -k '$K'" >> $L
timeout 45 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_attention_gpu.py -m gpu -q -x -k "$K" 2>&1 | grep -v "^$" | tail -6 >> $L
echo "== compute-sanitizer --tool racecheck python -m pytest tests/test_attention_gpu.py -m gpu -q -x -k grouped" >> $L
timeout 45 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_attention_gpu.py -m gpu -q -x -k "grouped" 2>&1 | grep -v "^$" | tail -6 >> $L
tail -20 $L
timeout 80 python -m pytest tests -m gpu -q -x > gpurun_out/r2o_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2o_pytest_gpu.log
