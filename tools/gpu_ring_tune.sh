#!/bin/bash
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621"
for cfg in "default" "NCCL_MIN_P2P_NCHANNELS=8 NCCL_MAX_P2P_NCHANNELS=8" "NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32"; do
  echo "=== $cfg"
  if [ "$cfg" = "default" ]; then
    $TR bench.py --gpus $N --steps 4 --warmup 3 --workload c5 2>/dev/null | grep '^{' | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['clocks'])"
  else
    env $cfg $TR bench.py --gpus $N --steps 4 --warmup 3 --workload c5 2>/dev/null | grep '^{' | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['clocks'])"
  fi
done
