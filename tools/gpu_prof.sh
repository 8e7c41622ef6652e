#!/bin/bash
# one full-set ncu capture of selected extractor kernels (bench process, 128 frames per launch)
mkdir -p gpurun_out
K=${KERNELS:-fast_cells}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -s ${SKIP:-2} -c ${COUNT:-1} \
  -f -o gpurun_out/${TAG:-prof} python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ba > gpurun_out/ncu_${TAG:-prof}.log 2>&1
tail -2 gpurun_out/ncu_${TAG:-prof}.log; ls -la gpurun_out/*.ncu-rep
