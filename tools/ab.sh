#!/bin/bash
# Same-box A/B of two builds of libvcmesh.so: A = git HEAD sources, B = working tree (already built in-tree).
# Usage (from the repo root, before gpurun): tools/ab.sh prepare ; then on the GPU box: tools/ab.sh run [reps]
set -e
case "$1" in
prepare)
    rm -rf /tmp/ab_src && mkdir -p /tmp/ab_src abtest
    for f in $(git ls-tree --name-only HEAD vcmeshconv_b200/csrc/); do git show HEAD:$f > /tmp/ab_src/$(basename $f); done
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
         --expt-relaxed-constexpr -shared -I include -DVCM_BUILDING /tmp/ab_src/*.cu -o abtest/libvcmesh_A.so
    ;;
run)
    reps=${2:-2}
    one() { timeout 200 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-fp32-mode --no-profile 2>/dev/null | \
            python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 %.4f ms/step' % d['ms_per_step'])"; }
    for i in $(seq $reps); do VCM_LIBRARY=$PWD/abtest/libvcmesh_A.so one A_head; one B_tree; done
    ;;
esac
