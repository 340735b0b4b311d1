#!/bin/bash
mkdir -p gpurun_out
./tools/micro/imma_rate | tee gpurun_out/imma_rate.txt
