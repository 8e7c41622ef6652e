#!/bin/bash
# quick GPU check: extractor / matcher parity tests + a short bench line (TAG names the outputs)
TAG=${TAG:-q}
python -m pytest tests/test_orb_gpu.py tests/test_match_gpu.py -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 5 --warmup 3 --no-ba --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value %.1f M  ms %.2f  e2e %.1f M" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6))
print({k: round(v,3) for k,v in d["roofline"]["stage_ms"].items()})
m=d.get("masked")
if m: print("masked %.1f M ratio %.3f e2e %.1f M" % (m["value"]/1e6, m["time_vs_unmasked"], m["e2e"]["value"]/1e6))
PY
