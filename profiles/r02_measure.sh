#!/bin/bash
# Round-2 measurement run on one B200 (gpurun -- 'bash profiles/r02_measure.sh'): GPU tests, smoke, the default
# bench line, the reference arm, a sweep over matrix sizes, and the ncu captures summarised under profiles/.
# Numbers printed by anything running under ncu are never bench values.
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
( time python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err ) 2>&1 | grep real
( time python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_n1.json 2>> gpurun_out/r02_bench_n1.err ) 2>&1 | grep real
show() { python -c "
import json,sys
d=json.load(open('$1'))
r=d['roofline']
print('$2', 'value %.3e fused %.3e e2e %.3e kern_ms %.3f bwd_ms %.3f' % (d['value'], d['fused_step']['pairs_per_s'], d['e2e']['value'], r['kernel_ms'], r['backward_ms']), r['bound'], r['frac'], 'other', r['other_bound']['frac'], 'step fp64 %.3f hbm %.3f' % (r['whole_step']['fp64_frac'], r['whole_step']['hbm_frac']))"; }
rm -f gpurun_out/r02_f_*.json
P="--steps 5 --warmup 3 --no-cpu-baseline --no-extras --no-n10"
for n in 2 3 4; do python bench.py --n $n $P > gpurun_out/r02_f_upper$n.json 2>>gpurun_out/r02_err.log; show gpurun_out/r02_f_upper$n.json "upper n=$n"; done
for n in 5 6; do python bench.py --n $n --pairs 16777216 $P > gpurun_out/r02_f_upper$n.json 2>>gpurun_out/r02_err.log; show gpurun_out/r02_f_upper$n.json "upper n=$n"; done
for n in 7 8 9 10; do python bench.py --n $n --metric fmin --pairs 4194304 $P > gpurun_out/r02_f_upper$n.json 2>>gpurun_out/r02_err.log; show gpurun_out/r02_f_upper$n.json "upper n=$n"; done
for n in 3 4; do python bench.py --kind bounded --metric fone --n $n $P > gpurun_out/r02_f_bounded$n.json 2>>gpurun_out/r02_err.log; show gpurun_out/r02_f_bounded$n.json "bounded n=$n"; done
for n in 6 7 10; do python bench.py --kind bounded --metric fone --n $n --pairs 4194304 $P > gpurun_out/r02_f_bounded$n.json 2>>gpurun_out/r02_err.log; show gpurun_out/r02_f_bounded$n.json "bounded n=$n"; done
for n in 4 7 10; do python bench.py --kind spd --n $n --pairs 16777216 $P > gpurun_out/r02_f_spd$n.json 2>>gpurun_out/r02_err.log; show gpurun_out/r02_f_spd$n.json "spd n=$n"; done
tail -3 gpurun_out/r02_err.log
Q="--steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-n10"
ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 2 -c 1 -o gpurun_out/r02_pair_n4 -f python bench.py $Q --pairs 1048576 > gpurun_out/r02_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:coop_kernel -s 2 -c 1 -o gpurun_out/r02_coop_n10 -f python bench.py --n 10 --metric fmin $Q --pairs 131072 > gpurun_out/r02_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 2 -c 1 -o gpurun_out/r02_pair_n6 -f python bench.py --n 6 $Q --pairs 262144 > gpurun_out/r02_ncu3.log 2>&1
ncu --set full --clock-control none -k regex:scatter_packed -s 2 -c 1 -o gpurun_out/r02_scatter_n4 -f python bench.py $Q --pairs 1048576 > gpurun_out/r02_ncu4.log 2>&1
ncu --set full --clock-control none -k regex:expand_table -s 2 -c 1 -o gpurun_out/r02_expand_n4 -f python bench.py $Q --pairs 1048576 > gpurun_out/r02_ncu5.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_n4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-n10 > gpurun_out/r02_ncu6.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_smoke.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_ncu7.log 2>&1
tail -2 gpurun_out/r02_ncu6.log
ls -la gpurun_out/r02_*
