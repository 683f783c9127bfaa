#!/bin/bash
# `ncu --set full` captures of the kernels added / changed in round 2 (one eager decision, tools/one_step.py), plus the
# roofline kernel.  Reports -> gpurun_out/<tag>_*.ncu-rep; summarise here with tools/ncu_traffic.py.
tag=${1:-r2d}
cap() {  # name regex count
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$2" -c $3 -f \
    -o gpurun_out/${tag}_$1 python tools/one_step.py > gpurun_out/${tag}_$1.log 2>&1
  tail -1 gpurun_out/${tag}_$1.log
}
cap verifier_heads "pool_chain_kernel|it_finalize_kernel|fuse_score_kernel" 4
cap long_mha "attn_mha_long_umma_kernel" 3
cap ln_reduce "layernorm_reduce_kernel|rmsnorm_reduce_kernel" 6
cap gateup "gemm_bf16_tcgen05_2sm" 6
cap denoise "attn_decode_umma_kernel|gemm_splitk_partial_tcgen05" 8
ls -la gpurun_out/${tag}_*.ncu-rep
