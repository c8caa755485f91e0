#!/bin/bash
# Builds experiment variants of the library side by side: scripts/build_exp.sh name "-Dflag ..." [name2 "flags2" ...]
# -> pragma_dsp_b200/exp/lib_<name>.so (objects under pragma_dsp_b200/exp/obj_<name>); timed with scripts/ab_tune.py --lib
set -u
cd "$(dirname "$0")/.."
mkdir -p pragma_dsp_b200/exp
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  PDSP_EXTRA_NVCC_FLAGS="$flags" PDSP_OBJ_DIR=pragma_dsp_b200/exp/obj_$name PDSP_LIB_OUT=pragma_dsp_b200/exp/lib_$name.so \
    python -m pragma_dsp_b200.build > pragma_dsp_b200/exp/build_$name.log 2>&1
  echo "$name rc=$? $(ls -la pragma_dsp_b200/exp/lib_$name.so 2>/dev/null | awk '{print $5}')"
done
