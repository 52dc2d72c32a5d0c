#!/bin/bash
# run the launch list for each library variant under semantic_depth_b200/variants (developer experiments)
mkdir -p gpurun_out
cp semantic_depth_b200/libsd_fusion.so /tmp/libsd_orig.so
for v in semantic_depth_b200/variants/libsd_fusion_*.so; do
  name=$(basename $v .so); name=${name#libsd_fusion_}
  cp $v semantic_depth_b200/libsd_fusion.so
  for cs in ${CELL_SCALES:-1.0}; do
    SD_KNN_CELL_SCALE=$cs ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${name}_cs$cs.csv \
      python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --skip-e2e --no-graph --slots 1 --batches 1 > gpurun_out/ncu_list_$name.log 2>&1
    echo "== variant $name cell scale $cs"; python tools/ncu_summary.py gpurun_out/launches_${name}_cs$cs.csv | grep -E "knn_|radius_kernel|TOTAL"
  done
done
cp /tmp/libsd_orig.so semantic_depth_b200/libsd_fusion.so
