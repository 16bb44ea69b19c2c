#!/bin/bash
# compute-sanitizer over the GPU parity tests (1 B200): gpurun -- 'bash profiles/r02_sanitizer.sh'
cd $GRAFT_REPO_ROOT
export PATH=/usr/local/cuda/bin:$PATH
export SYMPA_UNDER_SANITIZER=1   # tests skip their allocator-growth assertion (the tool keeps freed blocks accounted)
S1='table_path and (upper-4 or bounded-3 or upper-3 or upper-7 or upper-10 or bounded-8 or spd-4) or edge_cases or workspace or sync_grad'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "$S1" > gpurun_out/r02_san_memcheck.log 2>&1; echo "memcheck rc=$?"; grep "ERROR SUMMARY\|passed\|failed" gpurun_out/r02_san_memcheck.log | tail -3
S2='table_path and (upper-4 or bounded-3 or upper-3) or workspace'
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "$S2" > gpurun_out/r02_san_racecheck.log 2>&1; echo "racecheck rc=$?"; grep "RACECHECK SUMMARY\|passed\|failed" gpurun_out/r02_san_racecheck.log | tail -3
S3='golden and (n3_spread or n4_mid or n6_init or n7_mid or n10_mid) or rsgd or distance_matrix or distortion_loss or bounded_by_rows'
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py tests/test_optim_golden.py -m gpu -q -x -k "$S3" > gpurun_out/r02_san_racecheck2.log 2>&1; echo "racecheck2 rc=$?"; grep "RACECHECK SUMMARY\|passed\|failed" gpurun_out/r02_san_racecheck2.log | tail -3
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py tests/test_optim_golden.py -m gpu -q -x -k "$S3 or epoch or feeder or far_apart" > gpurun_out/r02_san_memcheck2.log 2>&1; echo "memcheck2 rc=$?"; grep "ERROR SUMMARY\|passed\|failed" gpurun_out/r02_san_memcheck2.log | tail -3
# kernels added late in round 2: SM-partitioned ticket scatter (side stream), accumulated backward, check_points, the
# one-directional ring exchange of the cooperative Jacobi (n = 8..10 through the split-path test)
S5='accumulator or check_points or scatter_rows or split_path'
if [ "$1" = "late" ] || [ -z "$1" ]; then
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "$S5" > gpurun_out/r02_san_memcheck3.log 2>&1; echo "memcheck3 rc=$?"; grep "ERROR SUMMARY\|passed\|failed" gpurun_out/r02_san_memcheck3.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "$S5" > gpurun_out/r02_san_racecheck3.log 2>&1; echo "racecheck3 rc=$?"; grep "RACECHECK SUMMARY\|passed\|failed" gpurun_out/r02_san_racecheck3.log | tail -3
fi
