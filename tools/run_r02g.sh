#!/bin/bash
# seventh GPU session: ring variant of the product kernel (A/B + ncu), cfg4 on ONE GPU (3.2e9 entries, CSCBlocks)
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
( free -g; nproc; nvidia-smi -L; df -h /dev/shm | tail -1 ) > gpurun_out/r02g_box.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02g_pytest.log
timeout 300 python tools/spmv_sweep.py --cells 42000 --genes 20000 --density 0.08 --num-dim 100 --tile-w 20480,16384 --variants 1,3,4 > gpurun_out/r02g_sweep_cfg2.log 2>&1
timeout 400 python tools/spmv_sweep.py --cells 250000 --genes 30000 --density 0.06 --num-dim 50 --maxit 4 --tile-w 20480,16384,12288 --variants 1,3,4 > gpurun_out/r02g_sweep_cfg3.log 2>&1
for v in 3 1; do
  env M3B_FLAT_VARIANT=$v timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_flat_dot -s 12 -c 2 -o gpurun_out/r02g_full_flat_v${v}_cfg3 -f python tools/profile_step.py --cells 250000 --genes 30000 --density 0.06 --num-dim 50 --maxit 2 > gpurun_out/r02g_ncu_v$v.log 2>&1
done
timeout 1200 python bench.py --workload cfg4 --steps 2 --warmup 1 --write-expected gpurun_out/expected_cfg4.json > gpurun_out/r02g_cfg4_n1.json 2> gpurun_out/r02g_cfg4_n1.err
echo finished
