#!/bin/bash
# chunk-size experiment: a rank's share of the 8-GPU run (2^23 pairs per step) in 1 / 2 / 4 / 8 API calls, and the full batch
cd $GRAFT_REPO_ROOT
show() { python -c "
import json,sys
d=json.load(open('$1'))
print('$2', 'value %.3e ms/step %.3f kern_ms %.3f bwd_ms %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline'].get('backward_ms',0)), 'frac', d['roofline']['frac'], 'e2e %.3e (%.3f ms)' % (d['e2e']['value'], d['e2e']['ms_per_step']))"; }
for c in 23 22 21 20; do SYMPA_BENCH_CHUNK_LOG2=$c python bench.py --n 4 --pairs 8388608 --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-n10 > gpurun_out/r2c_r8_$c.json 2>>gpurun_out/r2c_err.log; show gpurun_out/r2c_r8_$c.json rank8share_chunk$c; done
for c in 23 22; do SYMPA_BENCH_CHUNK_LOG2=$c python bench.py --n 4 --steps 5 --warmup 3 --no-cpu-baseline --no-extras --no-n10 > gpurun_out/r2c_$c.json 2>>gpurun_out/r2c_err.log; show gpurun_out/r2c_$c.json full_chunk$c; done
tail -3 gpurun_out/r2c_err.log
