#!/bin/bash
# Tensor-pipe / DRAM / L2 figures of the denoise-loop kernels in the BATCHED throughput mode (B observations, M = 160 B
# suffix rows): a metric list instead of --set full, ~100 launches from the middle of the loop.
# usage: bash tools/ncu_batched_denoise.sh <tag>  ->  gpurun_out/<tag>_batched_denoise.csv (summarise: tools/ncu_csv_by_kernel.py)
tag=${1:-r2j}
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed
B=${B:-8} timeout 900 ncu --metrics $M --clock-control none --profile-from-start off \
  -k "regex:^gemm_bf16_tcgen05$|attn_decode_umma_kernel|rmsnorm_warp_kernel" --launch-skip ${SKIP:-700} -c ${COUNT:-120} --csv \
  --log-file gpurun_out/${tag}_batched_denoise.csv python tools/batch_profile.py > gpurun_out/${tag}_batched_denoise.log 2>&1
grep -c . gpurun_out/${tag}_batched_denoise.csv
