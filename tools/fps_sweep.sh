#!/bin/bash
# developer sweep of FPS launch shapes on the product library (uninstrumented)
for cfg in "0 0" "16 512" "14 512" "12 512" "10 512" "8 512" "16 1024" "8 1024"; do
  set -- $cfg
  if [ "$1" = "0" ]; then echo -n "auto:        "; python tools/run_fps_once.py 40000 2048 8 3 | tail -1
  else echo -n "cluster=$1 threads=$2: "; PN2_FPS_CLUSTER=$1 PN2_FPS_THREADS=$2 python tools/run_fps_once.py 40000 2048 8 3 | tail -1; fi
done
echo -n "b=7 cluster 16: "; PN2_FPS_CLUSTER=16 python tools/run_fps_once.py 40000 2048 7 3 | tail -1
echo -n "sa2 512thr: "; python tools/run_fps_once.py 2048 1024 8 3 | tail -1
echo -n "sa2 1024thr: "; PN2_FPS_THREADS=1024 python tools/run_fps_once.py 2048 1024 8 3 | tail -1
echo -n "sa3: "; python tools/run_fps_once.py 1024 512 8 3 | tail -1
echo -n "sa4: "; python tools/run_fps_once.py 512 256 8 3 | tail -1
