#!/bin/bash
# Closing GPU visit of a round: parity tests, smoke, the default bench line and the reference arm, then initcheck over the
# erosion / masked-extraction tests.  Everything lands in gpurun_out/ under ${TAG}.
mkdir -p gpurun_out
TAG=${TAG:-r2j}
T0=$SECONDS
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$? t=$((SECONDS - T0))s"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$? t=$((SECONDS - T0))s"; tail -3 gpurun_out/${TAG}_smoke.log
timeout 400 python bench.py > gpurun_out/${TAG}_bench_line.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$? t=$((SECONDS - T0))s"
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; echo "ref rc=$? t=$((SECONDS - T0))s"
TAG=$TAG TOOLS=initcheck LIMIT=100 SEL='test_erosion_constant_tile or test_edge_cases or test_distinctive_descriptors_match_oracle' bash tools/gpu_initcheck.sh > /dev/null 2>&1
head -3 gpurun_out/${TAG}_sanitizer_initcheck.log; echo "t=$((SECONDS - T0))s"
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_line.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "clocks", d.get("clocks"))
print("masked", json.dumps(d.get("masked"))[:1200])
PY
