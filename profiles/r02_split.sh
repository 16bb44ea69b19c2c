#!/bin/bash
# the three-kernel path of the cooperative sizes: per-kernel durations (ncu launch list) and the bench line
cd $GRAFT_REPO_ROOT
N=${1:-10}
Q="--steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-n10"
SYMPA_SPLIT_PATH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_split_launches_n$N.csv python bench.py --n $N --metric fmin $Q --pairs 524288 > gpurun_out/r02_split_ncu.log 2>&1
python profiles/summarize_launches.py gpurun_out/r02_split_launches_n$N.csv "split path n=$N" | head -12
for v in 0 1; do SYMPA_SPLIT_PATH=$v python bench.py --n $N --metric fmin --pairs 4194304 --steps 5 --warmup 3 --no-cpu-baseline --no-extras --no-n10 > gpurun_out/r02_split_$v.json 2>>gpurun_out/r02_split_err.log
python -c "
import json
d=json.load(open('gpurun_out/r02_split_$v.json')); r=d['roofline']
print('split=$v', 'value %.3e kern_ms %.3f frac %.4f' % (d['value'], r['kernel_ms'], r['frac']))"; done
tail -2 gpurun_out/r02_split_err.log
