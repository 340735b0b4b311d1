#!/bin/bash
# ncu captures of the K1 kernels: config-2 shape (5x downscale, 23 taps) and a zoom-crop batch (scale ~ 1)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k1_" -s 6 -c 2 -o gpurun_out/prof_k1 -f \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --images 8 > gpurun_out/ncu_k1.log 2>&1; echo "ncu k1 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k1_" -s 4 -c 4 -o gpurun_out/prof_k1_zoom -f \
  python tools/bench_configs.py --config 4 > gpurun_out/ncu_k1_zoom.log 2>&1; echo "ncu k1 zoom exit $?"
