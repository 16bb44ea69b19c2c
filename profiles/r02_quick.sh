#!/bin/bash
# quick check after a kernel change: GPU tests + a few bench lines (args: "name|bench args" ...)
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
show() { python -c "
import json,sys
d=json.load(open('$1'))
print('$2', 'value %.3e ms/step %.2f kern_ms %.3f bwd_ms %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline'].get('backward_ms',0)), 'fp64', d['roofline']['frac'], 'e2e %.3e' % d['e2e']['value'])"; }
for spec in "$@"; do name=${spec%%|*}; args=${spec#*|}; python bench.py $args --steps 5 --warmup 3 --no-cpu-baseline --no-extras --no-n10 > gpurun_out/r2q_$name.json 2>>gpurun_out/r2q_err.log; show gpurun_out/r2q_$name.json "$name"; done
tail -3 gpurun_out/r2q_err.log
