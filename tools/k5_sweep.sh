#!/bin/bash
# K5 per layer shape of C2: cp.async kernels (VCM_K5_TMA=0) against the TMA-staged kernel and its knobs.
for layer in ${LAYERS:-dec3 dec4 dec2 dec1 enc1 enc2 enc3 enc4}; do
  VCM_K5_TMA=0 python tools/kernel_bench.py --layer $layer --kernel k5 | sed 's/^/old      /'
  for cfg in ${CFGS:-"4 2 64" "2 3 64" "2 2 64" "3 2 64" "4 2 32" "4 3 32"}; do
    set -- $cfg
    VCM_K5_WARPS=$1 VCM_K5_STAGES=$2 VCM_K5_CW=$3 python tools/kernel_bench.py --layer $layer --kernel k5 | sed "s/^/w$1 s$2 cw$3 /"
  done
done
