#!/bin/bash
# A/B of build variants (build/mkvariant.sh NAME "N list" "-D...") against the main library on the bench kernels.
cd $GRAFT_REPO_ROOT
show() { python -c "
import json,sys
d=json.load(open('$1'))
print('$2', 'value %.3e ms/step %.2f kern_ms %.3f bwd_ms %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline'].get('backward_ms',0)), 'fp64', d['roofline']['frac'], 'e2e %.3e' % d['e2e']['value'])"; }
run() { SYMPA_B200_LIB=$2 python bench.py ${@:3} --steps 5 --warmup 3 --no-cpu-baseline --no-extras --no-n10 > gpurun_out/r2v_$1.json 2>>gpurun_out/r2v_err.log; show gpurun_out/r2v_$1.json "$1"; }
M=$GRAFT_REPO_ROOT/sympa_b200/libsympa_b200.so
V=$GRAFT_REPO_ROOT/build/variants
run n4 $M --n 4
run n3 $M --n 3
for v in $V/libsympa_*.so; do name=$(basename $v .so); name=${name#libsympa_}; run n4_$name $v --n 4; run n3_$name $v --n 3; done
tail -3 gpurun_out/r2v_err.log
