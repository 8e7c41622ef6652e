#!/bin/bash
# Full-set ncu capture of the kernels only the masked stream runs: erosion, the pyramid instance that carries the mask level, masked FAST.
mkdir -p gpurun_out
TAG=${TAG:-r2j}
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:'erode10_tile_kernel|pyr_resize_strip_kernel<\(bool\)1>|fast_cells_warp_kernel<.*\(bool\)1>' -c 12 \
  -f -o gpurun_out/${TAG}_masked python bench.py --steps 1 --warmup 3 --pairs 128 --no-cpu-baseline --no-ba > gpurun_out/${TAG}_ncu_masked.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/${TAG}_masked.ncu-rep --page raw --csv > gpurun_out/${TAG}_masked_raw.csv 2>/dev/null
grep -c . gpurun_out/${TAG}_masked_raw.csv; tail -3 gpurun_out/${TAG}_ncu_masked.log | cut -c1-300
ls -la gpurun_out/${TAG}_masked*
