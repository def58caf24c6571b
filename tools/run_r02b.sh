#!/bin/bash
# second GPU session: device-built tables -- A/B against the host builders, tests, cfg3 at full size vs the oracle
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
(timeout 1500 python tools/fullsize_parity.py --workload cfg3 > gpurun_out/r02b_fullsize_cfg3.log 2>&1) &
timeout 600 python tools/ab_tables.py > gpurun_out/r02b_ab_tables.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r02b_pytest.log
M3B_HOST_TABLES=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "a549 or multi_tile or unit_split or ragged" 2>&1 | tail -5 > gpurun_out/r02b_pytest_hosttables.log
wait
timeout 600 python bench.py --workload cfg3 --steps 3 --warmup 3 > gpurun_out/r02b_cfg3_n1.json 2> gpurun_out/r02b_cfg3_n1.err
timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_cfg2_n1.json 2> gpurun_out/r02b_cfg2_n1.err
echo finished
