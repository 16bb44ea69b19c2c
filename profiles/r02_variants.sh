#!/bin/bash
# A/B of build variants (profiles/mkvariant.sh NAME "N list" "-D...") against the main library.
# usage: bash profiles/r02_variants.sh "name|bench args" ...   (every variant .so under build/variants runs every spec)
cd $GRAFT_REPO_ROOT
show() { python -c "
import json,sys
d=json.load(open('$1'))
print('$2', 'value %.3e ms/step %.2f kern_ms %.3f bwd_ms %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline'].get('backward_ms',0)), 'frac', d['roofline']['frac'], 'e2e %.3e' % d['e2e']['value'])"; }
run() { SYMPA_B200_LIB=$2 python bench.py ${@:3} --steps 5 --warmup 3 --no-cpu-baseline --no-extras --no-n10 > gpurun_out/r2v_$1.json 2>>gpurun_out/r2v_err.log; show gpurun_out/r2v_$1.json "$1"; }
M=$GRAFT_REPO_ROOT/sympa_b200/libsympa_b200.so
V=$GRAFT_REPO_ROOT/build/variants
for spec in "$@"; do
  name=${spec%%|*}; args=${spec#*|}
  run ${name}_main $M $args
  for v in $V/libsympa_*.so; do vn=$(basename $v .so); vn=${vn#libsympa_}; run ${name}_$vn $v $args; done
done
tail -3 gpurun_out/r2v_err.log
