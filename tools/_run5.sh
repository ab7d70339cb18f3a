mkdir -p gpurun_out
tools/ab.sh run 2
timeout 200 python tools/graph_gaps.py gpurun_out/timeline_r1_graph.tsv > gpurun_out/graph_gaps_r1.txt 2>&1; head -9 gpurun_out/graph_gaps_r1.txt | tail -6
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-400
timeout 300 python bench.py --vertices 5023 --batch 64 --no-cpu-baseline --no-fp32-mode --breakdown-file gpurun_out/breakdown_r1_c3.txt > gpurun_out/bench_r1_c3.json 2> gpurun_out/c3.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_c3.json')); print('C3', d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'])" || tail -5 gpurun_out/c3.err
timeout 500 python bench.py --vertices 150000 --batch 8 --no-cpu-baseline --no-fp32-mode --breakdown-file gpurun_out/breakdown_r1_c4.txt > gpurun_out/bench_r1_c4.json 2> gpurun_out/c4.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_c4.json')); print('C4', d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'])" || tail -5 gpurun_out/c4.err
