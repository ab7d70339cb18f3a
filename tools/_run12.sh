mkdir -p gpurun_out
run() { timeout -k 5 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $NG --steps 40 --no-cpu-baseline --no-fp32-mode --no-profile > gpurun_out/ng_$1.json 2> gpurun_out/ng_$1.err; echo "$1 rc=$?"; python -c "
import json; d=json.loads([l for l in open('gpurun_out/ng_$1.json') if l.startswith('{')][-1]); print('$1 N=%d %.4f ms/step value %.0f e2e %.0f' % (d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value']))" || grep -v "^$" gpurun_out/ng_$1.err | tail -8; }
VCM_SYMM_AR=two_shot run two_shot 29561
VCM_SYMM_AR=multimem run multimem 29562
run nccl 29563
VCM_SYMM_AR=two_shot run two_shot 29564
