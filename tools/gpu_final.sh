#!/bin/bash
# what the driver runs at round end, in its order: the GPU suite in ONE process, smoke, the reference arm, the default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu_one_process.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu_one_process.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -n 3 gpurun_out/bench_default.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'score', d['e2e_score_map_mode']['value'], 'cpu', d['cpu_baseline']['value'])
print('roofline', d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['issue_roofline'])
print('golden', {k: v for k, v in d['golden_check_batch0'].items() if k != 'per_frame'}, 'mismatch', d['result_mismatches_vs_first_pass'])
PY
