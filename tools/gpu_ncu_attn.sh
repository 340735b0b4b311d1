#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_tc_kernel" -s 4 -c 1 -o gpurun_out/prof_attn_tc -f \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --images 8 > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn_tc exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_kernel" -s 30 -c 1 -o gpurun_out/prof_attn_win -f \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --images 8 > gpurun_out/ncu_attn2.log 2>&1; echo "ncu attn_win exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc" -s 140 -c 4 -o gpurun_out/prof_gemm2 -f \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --images 8 > gpurun_out/ncu_gemm2.log 2>&1; echo "ncu gemm exit $?"
