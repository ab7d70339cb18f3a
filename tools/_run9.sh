timeout 300 python -m pytest tests/test_gpu_k5_tc.py tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_train.py -x -q 2>&1 | tail -4
for l in dec4 enc3 dec2; do for v in 1 0; do echo "buckets=$v"; VCM_K5_BUCKETS=$v timeout 100 python tools/kernel_bench.py --layer $l --kernel k5 --iters 10; done; done
one() { timeout 200 python bench.py --steps 40 --no-cpu-baseline --no-fp32-mode --no-profile 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 %.4f ms/step' % d['ms_per_step'])"; }
one buckets; VCM_K5_BUCKETS=0 one nobuckets; one buckets; VCM_K5_BUCKETS=0 one nobuckets
