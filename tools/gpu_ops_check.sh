#!/bin/bash
# after the fused path moved behind torch.ops.sd_fusion: the tests that cross it + a short bench
mkdir -p gpurun_out
for f in tests/test_gpu_fuse.py tests/test_gpu_scores.py tests/test_gpu_fcn_head.py tests/test_gpu_frame_processor.py tests/test_gpu_sanitizer.py tests/test_gpu_fullsize.py; do
  timeout 600 python -m pytest $f -q -m gpu -x --timeout 600 > gpurun_out/$(basename $f .py).log 2>&1; rc=$?
  echo "$f rc=$rc $(tail -n 1 gpurun_out/$(basename $f .py).log)"
  if [ $rc -ne 0 ]; then tail -n 40 gpurun_out/$(basename $f .py).log; fi
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --steps 100 --warmup 10 --skip-cpu-baseline --skip-configs --no-kernel-timing > gpurun_out/bench_ops.json 2> gpurun_out/bench_ops.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_ops.json').read().strip().splitlines()[-1])
print('value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'score', round(d['e2e_score_map_mode']['value'], 1), 'mismatch', d['result_mismatches_vs_first_pass'], 'golden', d['golden_check_batch0']['all_stage_counts_equal'])
PY
