#!/bin/bash
O=gpurun_out/r2_s9
mkdir -p $O
timeout 200 python tools/knock_timeline.py > $O/knock_timeline.txt 2>&1
RIFE_B200_KS=1 timeout 200 python tools/knock_timeline.py > $O/knock_timeline_ks1.txt 2>&1
cat $O/knock_timeline.txt
