#!/bin/bash
# 8-GPU box: contract bench at N=8, and the config-3 stream (600 frames, strong scaling) at N=8, 4, 2
mkdir -p gpurun_out
N=8 STEPS=60 STREAM=600 bash tools/gpu_multi.sh
for n in 4 2; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n"
  timeout 600 $TR bench.py --gpus $n --stream 600 --warmup 2 > gpurun_out/stream_n$n.json 2> gpurun_out/stream_n$n.err; echo "stream N=$n rc=$?"
  python -c "
import json; d=json.loads(open('gpurun_out/stream_n$n.json').read().strip().splitlines()[-1]); print('stream N=$n', d['value'], 'frames/s elapsed', d['elapsed_ms'], 'sha', d['answers_sha256'][:16], 'golden', d['golden_frames_equal'], d['device_map'].get('map'))"
done
