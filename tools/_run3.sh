mkdir -p gpurun_out
run() { timeout -k 5 $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 2 --steps 40 --no-cpu-baseline --no-fp32-mode --no-profile > gpurun_out/n2_$1.json 2> gpurun_out/n2_$1.err; echo "$1 rc=$?"; python -c "import json,sys; d=json.loads(open('gpurun_out/n2_$1.json').read()); print('$1 %.4f ms/step value %.0f e2e %.4f' % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step']))" || tail -5 gpurun_out/n2_$1.err; }
run default 120 29511
