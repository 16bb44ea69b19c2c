#!/bin/bash
# Jacobi stopping threshold of the cooperative sizes (SY_STOP_LOG2 variants under build/variants): parity tests and bench lines per variant
cd $GRAFT_REPO_ROOT
for v in build/variants/libsympa_*.so; do echo "== $v"; SYMPA_B200_LIB=$GRAFT_REPO_ROOT/$v python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3; done
bash profiles/r02_variants.sh "n10|--n 10 --metric fmin --pairs 4194304" "n8|--n 8 --metric fmin --pairs 4194304"
