#!/bin/bash
# One GPU visit: parity tests, then the bench.  Logs under gpurun_out/.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $BENCH_ARGS > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"]); print(d["roofline"]["stage_ms"]); print(d.get("ba",{}))
PY
tail -3 gpurun_out/bench_default.err
