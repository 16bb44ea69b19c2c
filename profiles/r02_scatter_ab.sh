#!/bin/bash
# A/B of the packed-table scatter kernel: GPU tests on the main library, then bench lines for every build/variants/*.so
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
bash profiles/r02_variants.sh "$@"
