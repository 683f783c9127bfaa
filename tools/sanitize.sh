#!/bin/bash
# compute-sanitizer memcheck over the small-shape GPU tests (eager launches: CVB_GRAPH=0), racecheck over the kernels that
# exchange data through (distributed) shared memory.  usage: bash tools/sanitize.sh <tag>
tag=${1:-r2}
export CVB_GRAPH=0
out=gpurun_out/${tag}_sanitizer.log
: > $out
run() { echo "== $*" >> $out; timeout 1500 "$@" 2>&1 | grep -v -i "warn\|^$" | tail -6 >> $out; }
run compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gemm_gpu.py tests/test_splitk_gpu.py tests/test_sgemm_gpu.py -m gpu -q -x
run compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_attention_gpu.py -m gpu -q -x
run compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_pi0_gpu.py tests/test_verifier_gpu.py tests/test_cover_gpu.py tests/test_batch_gpu.py -m gpu -q -x -k "TINY or tiny or select or policy"
run compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_verifier_gpu.py -m gpu -q -x -k "heads and VTINY"
cat $out
