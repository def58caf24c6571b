#!/bin/bash
# last sanity pass after the final edits: smoke + GPU tests
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final5_smoke.log 2>&1
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/final5_pytest.log
echo finished
