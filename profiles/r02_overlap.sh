#!/bin/bash
# sweep of the SM partition of the overlapped scatter (ops.ACC_SCATTER_SMS): GPU tests, then bench lines per setting
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
show() { python -c "
import json,sys
d=json.load(open('$1'))
print('$2', 'value %.3e ms/step %.2f kern_ms %.3f bwd_ms %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline'].get('backward_ms',0)), 'frac', d['roofline']['frac'], 'e2e %.3e' % d['e2e']['value'])"; }
for spec in "$@"; do name=${spec%%|*}; args=${spec#*|}
  for sms in ${SMS_LIST:-0 32 48 64 96}; do
    SYMPA_ACC_SCATTER_SMS=$sms python bench.py $args --steps 5 --warmup 3 --no-cpu-baseline --no-extras --no-n10 > gpurun_out/r2o_${name}_$sms.json 2>>gpurun_out/r2o_err.log; show gpurun_out/r2o_${name}_$sms.json "${name}_sms$sms"; done; done
tail -3 gpurun_out/r2o_err.log
