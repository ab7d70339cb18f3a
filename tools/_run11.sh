timeout -k 5 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 tools/symm_probe.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -30
