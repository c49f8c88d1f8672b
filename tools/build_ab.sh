#!/bin/bash
# usage: tools/build_ab.sh name "NVCC_EXTRA" [name "NVCC_EXTRA" ...]  -- builds rayito_b200/csrc/_ab/lib_<name>.so (CPU container)
# and restores the default library afterwards
mkdir -p rayito_b200/csrc/_ab
while [ $# -ge 2 ]; do
  name=$1; extra=$2; shift 2
  RT_NVCC_EXTRA="$extra" python -c "from rayito_b200 import build; build.build_core(force=True)" || exit 1
  cp rayito_b200/csrc/librayito_b200.so rayito_b200/csrc/_ab/lib_$name.so
  echo "built _ab/lib_$name.so [$extra]"
done
python -c "from rayito_b200 import build; build.build_core(force=True)"
