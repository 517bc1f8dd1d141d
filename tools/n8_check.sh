#!/bin/bash
# 8-GPU validation of the sharded pipelines + the bench line (run through gpurun --gpus 8)
N=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tools/check_sharded.py 16 4x 2>&1 | grep -E "SHARDED|Error|error" | head -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 tools/check_sharded.py 16 8x3 2>&1 | grep -E "SHARDED|Error|error" | head -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02e_bench_n$N.json 2> gpurun_out/r02e_bench_n$N.err; echo rc=$?; tail -3 gpurun_out/r02e_bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29554 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/r02e_bench_ref_n$N.json 2>/dev/null; echo ref_rc=$?
