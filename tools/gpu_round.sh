#!/bin/bash
# One GPU visit: parity tests, then the bench with the measurement switches.  Logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for v in default ADB_FAST_TWO_PHASE ADB_PYR_SIMPLE; do
  if [ $v = default ]; then timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ba > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  else env $v=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ba > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; fi
  echo "== $v"; tail -c 1500 gpurun_out/bench_$v.json; tail -3 gpurun_out/bench_$v.err
done
