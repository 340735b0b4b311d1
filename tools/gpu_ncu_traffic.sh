#!/bin/bash
# full-size (64 images) captures for bench.py's roofline.traffic (tools/ncu_traffic.py turns them into profiles/ncu_traffic.json):
# DRAM bytes, duration and tensor-pipe activity per launch of one whole block's four GEMMs (QKV, proj, SwiGLU, down), of the
# window / full attention kernels, and of the three K1 launches of one step.  bench.py --steps 1 --warmup 1 runs 3 warm-up steps
# + 1 timed step of 131 GEMM launches each.
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-latency --images 64"
timeout 900 ncu --metrics $M --clock-control none -k regex:"gemm_tc" -s 132 -c 4 --csv --log-file gpurun_out/traffic_gemm_64img.csv $B > gpurun_out/ncu_traffic.log 2>&1; echo "exit $?"
timeout 900 ncu --metrics $M --clock-control none -k regex:"attn_" -s 38 -c 8 --csv --log-file gpurun_out/traffic_attn_64img.csv $B > gpurun_out/ncu_traffic2.log 2>&1; echo "exit $?"
timeout 900 ncu --metrics $M --clock-control none -k regex:"k1_" -s 3 -c 3 --csv --log-file gpurun_out/traffic_k1_64img.csv $B > gpurun_out/ncu_traffic3.log 2>&1; echo "exit $?"
