#!/bin/bash
# Round-2 closing GPU visit (after the reference pins; the kernels are those of r2c except the Huber float-dsqr line): parity tests, smoke,
# the bench line (default flags) and the CPU arm on the same box.
mkdir -p gpurun_out
TAG=${TAG:-r2d}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_line.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; echo "ref rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/${TAG}_bench_line.json") if l.startswith("{")][-1])
print("value %.1f M  ms %.2f  e2e %.1f M  masked %.1f M  ba %.0f M  dyn %.0f M  search %.0f M" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["masked"]["value"]/1e6, d["ba"]["value"]/1e6, d["ba_dynamic"]["value"]/1e6, d["search"]["value"]/1e6))
PY
