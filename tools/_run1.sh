mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
one() { timeout 200 python bench.py --steps 40 --no-cpu-baseline --no-fp32-mode --no-profile 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 %.4f ms/step e2e %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step']))"; }
one base_new
VCM_PREP_STREAM=0 one no_prep
VCM_EPS_OUTSIDE=0 one eps_inside
VCM_PREP_STREAM=0 VCM_EPS_OUTSIDE=0 one neither
one base_new
for l in dec3 dec1 dec4 enc2; do timeout 100 python tools/kernel_bench.py --layer $l --kernel k6 --iters 10; done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:scatter_rev -c 1 -s 3 -o gpurun_out/k6_dec3 python tools/kernel_bench.py --layer dec3 --kernel k6 --iters 2 > gpurun_out/k6_ncu.log 2>&1
tail -2 gpurun_out/k6_ncu.log
