#!/bin/bash
# full-size (64 images) capture of the dominant GEMM launches: DRAM traffic per launch for bench.py's roofline.traffic
mkdir -p gpurun_out
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none \
  -k regex:"gemm_tc" -s 133 -c 5 --csv --log-file gpurun_out/traffic_gemm_64img.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --images 64 > gpurun_out/ncu_traffic.log 2>&1; echo "exit $?"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  -k regex:"k1_|attn_" -s 8 -c 8 --csv --log-file gpurun_out/traffic_other_64img.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --images 64 > gpurun_out/ncu_traffic2.log 2>&1; echo "exit $?"
