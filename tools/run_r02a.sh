#!/bin/bash
# first GPU session of round 2: tests (incl. sharded parity at world 2), full-size parity, bench lines
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,memory.total --format=csv | head -3 > gpurun_out/r02a_gpus.txt; nproc >> gpurun_out/r02a_gpus.txt; free -g >> gpurun_out/r02a_gpus.txt
(timeout 1200 python tools/fullsize_parity.py --workload cfg2 > gpurun_out/r02a_fullsize_cfg2.log 2>&1) &
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r02a_pytest.log
wait
timeout 600 python bench.py --workload cfg3 --steps 3 --warmup 3 --write-expected gpurun_out/expected_cfg3_bench.json > gpurun_out/r02a_cfg3_n1.json 2> gpurun_out/r02a_cfg3_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --workload cfg3 --steps 3 --warmup 3 > gpurun_out/r02a_cfg3_n2.json 2> gpurun_out/r02a_cfg3_n2.err
timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 > gpurun_out/r02a_cfg2_n1.json 2> gpurun_out/r02a_cfg2_n1.err
timeout 900 python bench.py --impl reference --workload cfg3 --steps 1 --warmup 0 > gpurun_out/r02a_cfg3_ref.json 2> gpurun_out/r02a_cfg3_ref.err
echo finished
