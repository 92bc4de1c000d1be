#!/bin/bash
# multi-GPU session: NCCL ring parity, then ring (c5), sharded (c4) and weak (c2) benches at N GPUs
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
mkdir -p gpurun_out
$TR tools/ring_check.py 8192 2>&1 | grep -E "ring_check|RING_CHECK|Error|error" | tee gpurun_out/ring_check_$N.log
$TR bench.py --gpus $N --steps 5 --warmup 3 --workload c5 2>gpurun_out/bench_c5_$N.err | tee gpurun_out/bench_c5_$N.json
$TR bench.py --gpus $N --steps 5 --warmup 3 --workload c4 2>gpurun_out/bench_c4_$N.err | tee gpurun_out/bench_c4_$N.json | cut -c1-600
$TR bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/bench_c2_$N.err | tee gpurun_out/bench_c2_$N.json | cut -c1-600
tail -3 gpurun_out/bench_c5_$N.err gpurun_out/bench_c4_$N.err
