mkdir -p gpurun_out
timeout 400 python bench.py --breakdown-file gpurun_out/breakdown_r1.txt > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err; echo bench rc=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_r1_n1.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['entry_point_all_shapes']['frac'], d['fp32_mode']['ms_per_step'], d['cpu_baseline']['value'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r1_step_tf32.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile --no-graph --no-fp32-mode > gpurun_out/launches.log 2>&1; echo launches rc=$?
timeout 200 ncu --set full --clock-control none --import-source on -k regex:coeff_grad_mma -c 1 -s 3 -o gpurun_out/k5_dec3 python tools/kernel_bench.py --layer dec3 --kernel k5 --iters 2 > gpurun_out/k5_ncu.log 2>&1; echo k5 rc=$?
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gather_contract_fwd -c 1 -s 3 -o gpurun_out/k1_dec3 python tools/kernel_bench.py --layer dec3 --kernel k1 --iters 2 > gpurun_out/k1_ncu.log 2>&1; echo k1 rc=$?
for k in k1 k5 k6 fwd dz dw; do timeout 100 python tools/kernel_bench.py --layer dec3 --kernel $k --iters 10; done
