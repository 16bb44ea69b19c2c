cd $GRAFT_REPO_ROOT
export PATH=/usr/local/cuda/bin:$PATH
export SYMPA_UNDER_SANITIZER=1   # tests skip their allocator-growth assertion (the tool keeps freed blocks accounted)
S5='accumulator or check_points or scatter_rows or split_path'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "$S5" > gpurun_out/r02_san_memcheck3.log 2>&1; echo "memcheck3 rc=$?"; grep "ERROR SUMMARY\|passed\|failed" gpurun_out/r02_san_memcheck3.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "$S5" > gpurun_out/r02_san_racecheck3.log 2>&1; echo "racecheck3 rc=$?"; grep "RACECHECK SUMMARY\|passed\|failed" gpurun_out/r02_san_racecheck3.log | tail -3
