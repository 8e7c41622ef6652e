#!/bin/bash
# round 2, visit A: new cluster Cholesky in isolation + the BA tests with config 5 (old solver still wired) + timing vs cuSOLVER
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_chol_gpu.py -x -q -k "294 or 33 or flagged" > gpurun_out/r2a_chol_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r2a_chol_memcheck.log
timeout 600 python -m pytest tests/test_chol_gpu.py -x -q > gpurun_out/r2a_chol_test.log 2>&1; echo "chol rc=$?"; tail -15 gpurun_out/r2a_chol_test.log
timeout 300 python tools/chol_bench.py > gpurun_out/r2a_chol_bench.jsonl 2> gpurun_out/r2a_chol_bench.err; echo "bench rc=$?"; cat gpurun_out/r2a_chol_bench.jsonl; tail -3 gpurun_out/r2a_chol_bench.err
timeout 900 python -m pytest tests/test_ba_gpu.py -x -q > gpurun_out/r2a_ba_test.log 2>&1; echo "ba rc=$?"; tail -5 gpurun_out/r2a_ba_test.log
timeout 600 python tools/gpu_check_ba.py > gpurun_out/r2a_check_ba.log 2>&1; tail -12 gpurun_out/r2a_check_ba.log
