#!/bin/bash
# re-measure selected sweep configurations (args: "name|bench args" ...) into gpurun_out/r02_f_<name>.json
cd $GRAFT_REPO_ROOT
P="--steps 5 --warmup 3 --no-cpu-baseline --no-extras --no-n10"
for spec in "$@"; do name=${spec%%|*}; args=${spec#*|}
  python bench.py $args $P > gpurun_out/r02_f_$name.json 2>>gpurun_out/r02_err.log
  python -c "
import json
d=json.load(open('gpurun_out/r02_f_$name.json')); r=d['roofline']
print('$name', 'value %.3e e2e %.3e kern_ms %.3f bwd_ms %.3f frac %.4f' % (d['value'], d['e2e']['value'], r['kernel_ms'], r['backward_ms'], r['frac']))"
done
