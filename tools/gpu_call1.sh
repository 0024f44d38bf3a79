#!/bin/bash
# round 2, GPU call 1: full GPU suite, A/B of the prepared variants, the bench line with parity
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
tail -5 gpurun_out/c1_pytest.log
timeout 900 bash tools/ab_optin.sh > gpurun_out/c1_ab_optin.txt 2>&1
tail -40 gpurun_out/c1_ab_optin.txt
timeout 400 python bench.py > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/c1_bench.json; tail -5 gpurun_out/c1_bench.err
