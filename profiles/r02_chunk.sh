cd $GRAFT_REPO_ROOT
show() { python -c "
import json,sys
d=json.load(open('$1'))
print('$2', 'value %.3e ms/step %.2f kern_ms %.3f bwd_ms %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline'].get('backward_ms',0)), 'frac', d['roofline']['frac'], 'e2e %.3e' % d['e2e']['value'])"; }
for c in 23 22 21; do SYMPA_BENCH_CHUNK_LOG2=$c python bench.py --n 4 --steps 5 --warmup 3 --no-cpu-baseline --no-extras --no-n10 > gpurun_out/r2c_$c.json 2>>gpurun_out/r2c_err.log; show gpurun_out/r2c_$c.json chunk$c; done
SYMPA_BENCH_CHUNK_LOG2=22 python bench.py --n 4 --pairs 8388608 --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-n10 > gpurun_out/r2c_r8a.json 2>>gpurun_out/r2c_err.log; show gpurun_out/r2c_r8a.json rank8_chunk22
SYMPA_BENCH_CHUNK_LOG2=23 python bench.py --n 4 --pairs 8388608 --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-n10 > gpurun_out/r2c_r8b.json 2>>gpurun_out/r2c_err.log; show gpurun_out/r2c_r8b.json rank8_chunk23
SYMPA_BENCH_CHUNK_LOG2=21 python bench.py --n 4 --pairs 8388608 --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-n10 > gpurun_out/r2c_r8c.json 2>>gpurun_out/r2c_err.log; show gpurun_out/r2c_r8c.json rank8_chunk21
tail -3 gpurun_out/r2c_err.log
