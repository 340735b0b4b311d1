#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k1_" -s 6 -c 2 -o gpurun_out/prof_k1 -f \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --images 8 > gpurun_out/ncu_k1.log 2>&1; echo "ncu k1 exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_tc|attn_|k1_|rmsnorm|transpose_v|gather" -s 705 -c 240 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --images 8 > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?"
