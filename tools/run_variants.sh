#!/bin/bash
# usage: tools/run_variants.sh tag [profile_align args...]  -- runs tools/profile_align.py against the in-tree library and every build/variants/*.so
tag=$1; shift
python tools/profile_align.py "$@" > gpurun_out/${tag}_default.log 2>&1
for v in build/variants/*.so; do
  n=$(basename $v .so)
  ICPB_LIB=$PWD/$v python tools/profile_align.py "$@" > gpurun_out/${tag}_$n.log 2>&1
done
grep -H "^align:" gpurun_out/${tag}_*.log
