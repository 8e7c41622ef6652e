#!/bin/bash
# Round-2 evidence run on one B200: bench line, reference arm, ncu launch list of the bench command, full-set captures of every
# kernel family (+ FP64 / tensor pipe counters for the BA and Cholesky kernels), cuSOLVER comparison.  Everything lands in gpurun_out/;
# tools/summarize_ncu.py turns it into profiles/.
mkdir -p gpurun_out
TAG=${TAG:-r2}
EXTRA="sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum"
ncu --query-metrics 2>/dev/null | grep -i -E "dmma|pipe_tensor|pipe_fp64" | cut -c1-110 > gpurun_out/${TAG}_metric_names.txt
timeout 900 python bench.py > gpurun_out/${TAG}_bench_line.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; echo "ref rc=$?"
timeout 300 python tools/chol_bench.py > gpurun_out/${TAG}_chol_vs_cusolver.jsonl 2>> gpurun_out/${TAG}_bench.err; echo "chol rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_list.log 2>&1; echo "list rc=$?"
timeout 900 ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:'fast_cells|pyr_resize|orient_describe|quadtree|stereo_|erode|blur7' -s 66 -c 28 \
  -f -o gpurun_out/${TAG}_orb python bench.py --steps 1 --warmup 3 --pairs 128 --no-cpu-baseline --no-ba > gpurun_out/${TAG}_ncu_orb.log 2>&1; echo "orb rc=$?"
timeout 900 ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:'ba_|chol_|lm_|schur_' -s 120 -c 36 \
  -f -o gpurun_out/${TAG}_ba python bench_ba.py > gpurun_out/${TAG}_ncu_ba.log 2>&1; echo "ba rc=$?"
timeout 900 ncu --set full --metrics $EXTRA --clock-control none -k regex:'chol_cluster' -s 40 -c 2 \
  -f -o gpurun_out/${TAG}_ba5 python bench_ba.py --dynamic > gpurun_out/${TAG}_ncu_ba5.log 2>&1; echo "ba5 rc=$?"
timeout 900 ncu --set full --metrics $EXTRA --clock-control none -k regex:'proj_search|project_last|bow_search|hamming_best2' -s 6 -c 5 \
  -f -o gpurun_out/${TAG}_search python bench_search.py > gpurun_out/${TAG}_ncu_search.log 2>&1; echo "search rc=$?"
for r in orb ba ba5 search; do
  ncu -i gpurun_out/${TAG}_$r.ncu-rep --page raw --csv > gpurun_out/${TAG}_${r}_raw.csv 2>/dev/null
done
# gpurun copies back at most 64 MiB: keep the .ncu-rep (source view) only for the Cholesky capture
rm -f gpurun_out/${TAG}_orb.ncu-rep gpurun_out/${TAG}_ba.ncu-rep gpurun_out/${TAG}_search.ncu-rep
ls -la gpurun_out/ | grep ${TAG}_ | tail -20
