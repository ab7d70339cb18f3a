#!/bin/bash
# Bench lines on N GPUs of one box (weak scaling, BASELINE configs C2 / C3 / C4), the 55.6 MB reduction probe and the
# N-rank vs one-process equivalence run. Usage: tools/multi_gpu_lines.sh N "c2 c3 c4" [extra tag]
N=$1; CFGS=${2:-c2}; out=gpurun_out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
port=29520
for c in $CFGS; do
  port=$((port+1))
  run $port bench.py --config $c --gpus $N --steps 30 --warmup 5 --no-profile --no-other-modes > $out/r2_bench_${c}_n$N.json 2> $out/r2_bench_${c}_n$N.err
  echo "bench $c n$N rc=$?"; head -c 260 $out/r2_bench_${c}_n$N.json; echo
done
if [ "${TWO_SHOT:-0}" = "1" ]; then
  VCM_SYMM_AR=two_shot run 29540 bench.py --config c2 --gpus $N --steps 30 --warmup 5 --no-profile --no-other-modes > $out/r2_bench_c2_n${N}_two_shot.json 2> $out/r2_bench_c2_n${N}_two_shot.err
  echo "bench c2 two_shot n$N rc=$?"; head -c 260 $out/r2_bench_c2_n${N}_two_shot.json; echo
fi
run 29541 tools/collective_probe.py > $out/r2_collective_probe_n$N.txt 2>&1; grep -v "^\*\|OMP_NUM\|^$" $out/r2_collective_probe_n$N.txt | tail -8
run 29542 tests/multi_gpu_equiv.py --steps 3 --out $out/r2_multi_gpu_equiv_n$N.json > $out/r2_equiv_n$N.log 2>&1; echo "equiv n$N rc=$?"; tail -1 $out/r2_equiv_n$N.log | cut -c1-600
