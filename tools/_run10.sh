mkdir -p gpurun_out
timeout -k 5 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 40 --no-cpu-baseline --no-fp32-mode --no-profile > gpurun_out/bench_r1_n8.json 2> gpurun_out/n8.err; echo "n8 rc=$?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_r1_n8.json') if l.startswith('{')][-1]); print('n8', d['value'], d['ms_per_step'], d['e2e']['value'])" || tail -5 gpurun_out/n8.err
