#!/bin/bash
# One `ncu --set full` capture per hot kernel on the dec3 layer shape (unpool3: 344 -> 1811 vertices, I = 128, O = 64,
# B = 16), each after the same command has exited 0 without ncu. Reports land in gpurun_out/r2_ncu_<kernel>.ncu-rep,
# the plain runs (with the algorithmic byte counts) in gpurun_out/r2_ncu_plain.txt.
set -u
out=gpurun_out
: > $out/r2_ncu_plain.txt
cap() {  # kernel_bench kernel, regex, tag
  python tools/kernel_bench.py --layer ${LAYER:-dec3} --kernel $1 --iters 3 >> $out/r2_ncu_plain.txt 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -f -o $out/r2_ncu_$3 \
      python tools/kernel_bench.py --layer ${LAYER:-dec3} --kernel $1 --iters 3 > $out/r2_ncu_$3.log 2>&1
}
for k in ${KERNELS:-k5 k1 fwd dz dw k6}; do
  case $k in
    k5) cap k5 coeff_grad_tma k5 ;;
    k1) cap k1 gather_fwd_tile k1 ;;
    fwd) cap fwd tc_gemm gemm_fwd ;;
    dz) cap dz tc_gemm gemm_dz ;;
    dw) cap dw tc_gemm gemm_dw ;;
    k6) cap k6 scatter_rev k6 ;;
  esac
done
cat $out/r2_ncu_plain.txt
