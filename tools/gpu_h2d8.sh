#!/bin/bash
# 8-GPU call: host->device copy ceiling per placement (tools/h2d_probe.py), then the bench at N=8 with NUMA binding
mkdir -p gpurun_out
N=${N:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR tools/h2d_probe.py > gpurun_out/h2d_probe_n$N.json 2> gpurun_out/h2d_probe_n$N.err; echo "probe rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/h2d_probe_n$N.json"))
for k, v in d["summary"].items():
    print(k, v["aggregate"], v["min"])
print(d["locality"])
PY
timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 3 --skip-cpu-baseline ${BENCH_ARGS} > gpurun_out/bench_n${N}_bind.json 2> gpurun_out/bench_n${N}_bind.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_n${N}_bind.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e'], 'score', d.get('e2e_score_map_mode'), d['host_binding'])"
