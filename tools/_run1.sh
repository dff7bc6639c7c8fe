#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/r02e_bench_1gpu.json 2> gpurun_out/r02e_bench_1gpu.err; echo "bench1 rc=$?"
cut -c1-400 gpurun_out/r02e_bench_1gpu.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02e_reference_1gpu_3steps.json 2> gpurun_out/r02e_reference.err; echo "ref rc=$?"
cat gpurun_out/r02e_reference_1gpu_3steps.json | cut -c1-600
timeout 600 python bench.py --to-rtol 512 --grid-depth 512 --alg SMSM_GLOBAL --s 10 > gpurun_out/r02e_ttr_512_smsm_global_s10_1gpu.json 2> gpurun_out/r02e_ttr.err; echo "ttr rc=$?"
cat gpurun_out/r02e_ttr_512_smsm_global_s10_1gpu.json | cut -c1-600
