#!/bin/bash
# Round-end GPU visit: parity tests, smoke, the bench (default flags), the ncu launch list of the same command and one
# full-set capture of every kernel family.  Everything lands in gpurun_out/; tools/summarize_ncu.py turns it into profiles/.
mkdir -p gpurun_out
TAG=${TAG:-r1b}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_line.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --pairs 128 --no-cpu-baseline > gpurun_out/${TAG}_ncu_list.log 2>&1; echo "list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fast_cells|pyr_resize|orient_describe|quadtree|stereo_' -s 22 -c 15 \
  -f -o gpurun_out/${TAG}_orb python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ba > gpurun_out/${TAG}_ncu_orb.log 2>&1; echo "orb rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:'ba_|chol_|pose_opt' -s 200 -c 20 \
  -f -o gpurun_out/${TAG}_ba python bench_ba.py > gpurun_out/${TAG}_ncu_ba.log 2>&1; echo "ba rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:'proj_search|project_last' -s 6 -c 2 \
  -f -o gpurun_out/${TAG}_search python bench_search.py > gpurun_out/${TAG}_ncu_search.log 2>&1; echo "search rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:'bow_search' -s 2 -c 1 \
  -f -o gpurun_out/${TAG}_bow python bench_search.py > gpurun_out/${TAG}_ncu_bow.log 2>&1; echo "bow rc=$?"
# gpurun copies back at most 64 MiB: keep the raw-metric CSV of every report, the .ncu-rep (source view) only for the extractor
for r in orb ba search bow; do
  ncu -i gpurun_out/${TAG}_$r.ncu-rep --page raw --csv > gpurun_out/${TAG}_${r}_raw.csv 2>/dev/null
done
rm -f gpurun_out/${TAG}_ba.ncu-rep gpurun_out/${TAG}_search.ncu-rep gpurun_out/${TAG}_bow.ncu-rep
ls -la gpurun_out/ | tail -20
