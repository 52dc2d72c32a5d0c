#!/bin/bash
# multi-GPU call: the contract bench at N ranks (device map + e2e per N) and the config-3 stream, strong scaling
mkdir -p gpurun_out
N=${N:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
timeout 500 $TR bench.py --gpus $N --steps ${STEPS:-30} --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
    print('N=$N value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'h2d/gpu', round(d['e2e']['h2d_gb_per_s_per_gpu'], 1),
          'score-map e2e', round(d['e2e_score_map_mode']['value'], 1), 'mismatch', d['result_mismatches_vs_first_pass'])
    print('device_map', d['device_map'])
except Exception as e:
    print('parse failed', e)
PY
tail -n 4 gpurun_out/bench_n$N.err
if [ -n "$STREAM" ]; then
  timeout 600 $TR bench.py --gpus $N --stream $STREAM --warmup 2 > gpurun_out/stream_n$N.json 2> gpurun_out/stream_n$N.err; echo "stream rc=$?"
  python -c "
import json; d=json.loads(open('gpurun_out/stream_n$N.json').read().strip().splitlines()[-1]); print('stream N=$N', d['value'], 'frames/s elapsed', d['elapsed_ms'], 'sha', d['answers_sha256'][:16], 'golden', d['golden_frames_equal'], d['device_map'].get('map'))"
  tail -n 3 gpurun_out/stream_n$N.err
fi
