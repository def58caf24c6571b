#!/bin/bash
# thirteenth GPU session (1 GPU): K10 32-warp variant, reference arm at cfg4, cfg2 / cfg3 lines with the final code
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r02m_pytest.log
env M3B_PROJECT_CPW2=1 timeout 600 python bench.py --workload cfg5 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02m_cfg5_cpw2.json 2> gpurun_out/r02m_cfg5_cpw2.err
env M3B_PROJECT_CW27=1 timeout 600 python bench.py --workload cfg5 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02m_cfg5_cw27.json 2> gpurun_out/r02m_cfg5_cw27.err
( time timeout 900 python bench.py --impl reference ) > gpurun_out/r02m_ref_cfg4.json 2> gpurun_out/r02m_ref_cfg4.err
timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 > gpurun_out/r02m_cfg2_n1.json 2> gpurun_out/r02m_cfg2_n1.err
timeout 600 python bench.py --workload cfg3 --steps 3 --warmup 3 > gpurun_out/r02m_cfg3_n1.json 2> gpurun_out/r02m_cfg3_n1.err
echo finished
