#!/bin/bash
# A/B call: ab_bench over the given libs, then the phase profile of the prof build(s).  usage: scripts/gpu_ab.sh TAG "lib1 lib2" ["proflib1 proflib2"] [W]
TAG=$1; LIBS=$2; PROFS=$3
mkdir -p gpurun_out
python scripts/ab_bench.py $LIBS 2>&1 | grep -v Warning | tee gpurun_out/${TAG}_ab.log
for p in $PROFS; do
  echo "== phase profile $p"; RSB_PROF_LIB=$PWD/$p RSB_PROF_WARP=$4 RSB_PROF2=1 timeout 300 python -u scripts/phase_profile.py 4096 240 0 2 > gpurun_out/${TAG}_phase_$(basename $p .so).log 2>&1; tail -15 gpurun_out/${TAG}_phase_$(basename $p .so).log | cut -c1-150
done
