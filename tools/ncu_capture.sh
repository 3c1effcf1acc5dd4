#!/bin/bash
# ncu --set full captures of every persistent-kernel flavour (tools/profile_kernels.py), exported as raw-page CSV under
# gpurun_out/ (the .ncu-rep files are deleted: gpurun copies back at most 64 MiB).  usage: tools/ncu_capture.sh [tag] [parts, default "tc ffma ffjord"]
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
parts=${2:-"tc ffma ffjord"}
for part in $parts; do
  n=4; [ $part = ffjord ] && n=2
  src=""; [ $part = tc ] && src="--import-source on"
  timeout 600 ncu --set full --clock-control none $src -k regex:solve_kernel -c $n -o $out/ncu_${part}_$tag -f python tools/profile_kernels.py $part > $out/ncu_${part}_$tag.log 2>&1
  tail -2 $out/ncu_${part}_$tag.log
  ncu -i $out/ncu_${part}_$tag.ncu-rep --page raw --csv > $out/solve_kernels_${tag}_${part}_raw.csv 2>/dev/null
  if [ $part = tc ]; then
    ncu -i $out/ncu_${part}_$tag.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip -9 > $out/solve_kernels_${tag}_tc_source_sass.csv.gz
  fi
  rm -f $out/ncu_${part}_$tag.ncu-rep
done
ls -la $out | tail -12
