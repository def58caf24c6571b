#!/bin/bash
# twelfth GPU session (1 GPU): projection kernel v2 variants at cfg5 (A/B), cfg4 e2e breakdown
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "transform or lsi" 2>&1 | tail -5 > gpurun_out/r02l_pytest.log
timeout 600 python bench.py --workload cfg5 --steps 3 --warmup 2 > gpurun_out/r02l_cfg5_cpw4.json 2> gpurun_out/r02l_cfg5_cpw4.err
env M3B_PROJECT_CW27=1 timeout 600 python bench.py --workload cfg5 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02l_cfg5_cw27.json 2> gpurun_out/r02l_cfg5_cw27.err
env M3B_PROJECT_CPW8=1 timeout 600 python bench.py --workload cfg5 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02l_cfg5_cpw8.json 2> gpurun_out/r02l_cfg5_cpw8.err
env M3B_PROJECT_V1=1 timeout 600 python bench.py --workload cfg5 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02l_cfg5_v1.json 2> gpurun_out/r02l_cfg5_v1.err
echo finished
