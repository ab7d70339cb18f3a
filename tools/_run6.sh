for l in enc2 enc4 enc3; do
  timeout 100 python tools/kernel_bench.py --layer $l --kernel k5 --iters 10
  VCM_K5_MMA=1 timeout 100 python tools/kernel_bench.py --layer $l --kernel k5 --iters 10
done
