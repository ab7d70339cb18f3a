mkdir -p gpurun_out
timeout -k 5 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 40 --no-cpu-baseline --no-fp32-mode > gpurun_out/bench_r1_n4.json 2> gpurun_out/n4.err; echo "n4 rc=$?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_r1_n4.json') if l.startswith('{')][-1]); print('n4', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'])" || tail -5 gpurun_out/n4.err
timeout 400 python bench.py --breakdown-file gpurun_out/breakdown_r1.txt > gpurun_out/bench_r1_n1.json 2> gpurun_out/n1.err; echo "n1 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_r1_n1.json')); print('n1', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'], [(s['shape'], round(s['frac'],2)) for s in d['roofline']['shapes']])"
