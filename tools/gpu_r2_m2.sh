#!/bin/bash
python tools/win_trace.py zoomearth_b200/_variants/libzoomvit_trace.so 2>&1 | tail -17 | awk '$1 % 2 == 0 {print $1, "Oready", $13, "Oloaded", $3, "pre-wait", $4, "post-wait", $5, "fenced", $6, "stored", $14}'
