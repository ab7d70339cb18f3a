#!/bin/bash
# One-GPU evidence pass of a round: GPU tests, the bench line of every single-GPU config with its breakdown, the
# 2000-step line, the graph timeline and the ncu launch list (after the same command exited 0 without ncu).
# Everything lands in gpurun_out/ev/; copy what should be judged into profiles/.
set -u
o=gpurun_out/ev
mkdir -p $o
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $o/pytest_gpu.txt
timeout 400 python bench.py --steps 20 --warmup 5 --breakdown-file $o/breakdown_c2.txt > $o/bench_c2.json 2> $o/bench_c2.err
timeout 300 python bench.py --config c1 --steps 20 --warmup 5 --breakdown-file $o/breakdown_c1.txt > $o/bench_c1.json 2> $o/bench_c1.err
timeout 400 python bench.py --config c5 --steps 20 --warmup 5 --breakdown-file $o/breakdown_c5.txt > $o/bench_c5.json 2> $o/bench_c5.err
timeout 500 python bench.py --config c3 --steps 20 --warmup 5 --breakdown-file $o/breakdown_c3.txt > $o/bench_c3.json 2> $o/bench_c3.err
timeout 200 python bench.py --steps 2000 --warmup 10 --no-profile --no-cpu-baseline --no-other-modes > $o/bench_c2_steps2000.json 2> $o/bench_c2_steps2000.err
timeout 200 python tools/graph_gaps.py $o/timeline.tsv > $o/graph_gaps.txt 2>&1
timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile --no-graph --no-other-modes > $o/plain_step.json 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $o/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile --no-graph --no-other-modes > $o/ncu_launch.log 2>&1
cat $o/pytest_gpu.txt; for c in c2 c1 c5 c3 c2_steps2000; do python -c "
import json,sys
d=json.loads(open('$o/bench_$c.json').read().strip().splitlines()[-1])
print('$c', d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), (d.get('roofline') or {}).get('kernel'), (d.get('roofline') or {}).get('frac'), d.get('clocks'))"; done
