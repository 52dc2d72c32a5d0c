#!/bin/bash
# build several library variants (developer experiments): tools/build_variants.sh name1="flags" name2="flags" ...
mkdir -p semantic_depth_b200/variants
for spec in "$@"; do
  name="${spec%%=*}"; flags="${spec#*=}"
  SD_EXTRA_NVCC_FLAGS="$flags" python -m semantic_depth_b200.build --force > /dev/null || exit 1
  cp semantic_depth_b200/libsd_fusion.so semantic_depth_b200/variants/libsd_fusion_$name.so
  echo "built $name ($flags)"
done
python -m semantic_depth_b200.build --force > /dev/null
