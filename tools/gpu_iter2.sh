#!/bin/bash
# round-2 iteration: full GPU suite (one process per file), dev bench with kernel timing, optional launch list
mkdir -p gpurun_out
for f in ${TESTS:-tests/test_gpu_*.py}; do
  timeout 900 python -m pytest $f -q -m gpu -x --timeout 600 > gpurun_out/$(basename $f .py).log 2>&1; rc=$?
  echo "$f rc=$rc $(tail -n 1 gpurun_out/$(basename $f .py).log)"
  if [ $rc -ne 0 ]; then tail -n 30 gpurun_out/$(basename $f .py).log; fi
done
timeout 600 python bench.py --steps ${STEPS:-100} --warmup 10 --skip-cpu-baseline ${BENCH_ARGS:---skip-e2e} > gpurun_out/bench_dev.json 2> gpurun_out/bench_dev.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_dev.json').read().strip().splitlines()[-1])
    print('value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 4), 'mismatch', d['result_mismatches_vs_first_pass'],
          'lat', d['batch_latency_ms']['mean'])
    print('stage_ms', {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d['stage_ms'].items() if k != 'note'})
    print('roofline frac', d['roofline']['frac'], 'pixel frac', d['roofline_pixel']['frac'], 'path frac', d['roofline_path']['frac'])
    for k in ('config4', 'config5'):
        if k in d: print(k, d[k].get('value'), d[k].get('unit'), d[k].get('ms'))
    print('golden', {k: v for k, v in (d.get('golden_check_batch0') or {}).items() if k != 'per_frame'})
    if 'e2e' in d: print('e2e', d['e2e']['value'], 'score', d.get('e2e_score_map_mode', {}).get('value'))
except Exception as e:
    print('bench parse failed', e)
PY
tail -n 5 gpurun_out/bench_dev.err
if [ -n "$NCU_LIST" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
     python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --skip-e2e --no-graph --slots 1 --batches 1 > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
  python tools/ncu_summary.py gpurun_out/launches.csv | head -30
fi
