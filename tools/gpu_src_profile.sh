#!/bin/bash
# Source-level (per SASS / per line) instruction counts of the extractor kernels: one full-set capture per kernel with
# --import-source on, exported as the source page CSV.  TAG names the outputs under gpurun_out/.
mkdir -p gpurun_out
TAG=${TAG:-src}
KERNELS=${KERNELS:-"fast_cells orient_describe blur7_level quadtree"}
for k in $KERNELS; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 1 -f -o gpurun_out/${TAG}_$k \
    python bench.py --steps 1 --warmup 2 --pairs 128 --no-cpu-baseline --no-ba > gpurun_out/${TAG}_ncu_$k.log 2>&1; echo "$k rc=$?"
  ncu -i gpurun_out/${TAG}_$k.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_${k}_sass.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_$k.ncu-rep --page source --csv --print-source cuda > gpurun_out/${TAG}_${k}_cuda.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_$k.ncu-rep --page raw --csv > gpurun_out/${TAG}_${k}_raw.csv 2>/dev/null
  rm -f gpurun_out/${TAG}_$k.ncu-rep
done
ls -la gpurun_out | grep ${TAG}_
