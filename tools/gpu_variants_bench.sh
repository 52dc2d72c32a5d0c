#!/bin/bash
# plain bench value for each library variant under semantic_depth_b200/variants (developer experiments; overlap effects
# do not show in a serialised ncu launch list)
mkdir -p gpurun_out
cp semantic_depth_b200/libsd_fusion.so /tmp/libsd_orig.so
for v in semantic_depth_b200/variants/libsd_fusion_*.so; do
  name=$(basename $v .so); name=${name#libsd_fusion_}
  cp $v semantic_depth_b200/libsd_fusion.so
  timeout 200 python bench.py --steps ${STEPS:-150} --warmup 8 --skip-cpu-baseline --skip-e2e > gpurun_out/vbench_$name.json 2> gpurun_out/vbench_$name.err
  echo "== variant $name: $(python -c "import json,sys; d=json.loads(open('gpurun_out/vbench_$name.json').read().strip().splitlines()[-1]); print(d['value'], 'frames/s  knn_ms', d['roofline']['kernel_ms'], 'isolated', d['roofline']['isolated']['kernel_ms'])")"
done
cp /tmp/libsd_orig.so semantic_depth_b200/libsd_fusion.so
