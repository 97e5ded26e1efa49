#!/bin/bash
# Second evidence session: ncu --set full of one bench-shaped launch (64 utterances x 16000 samples), cluster-fit sweep, smoke.
mkdir -p gpurun_out/final2
O=gpurun_out/final2
timeout 120 python __graft_entry__.py smoke > $O/smoke.txt 2>&1
KERNELS=lat timeout 300 python tools/lat_quick.py 64 768 784 1184 > $O/lat_sweep.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "teacher and lat" 2>&1 | tail -3 > $O/pytest_lat.txt
N=16000 KERNELS=lat timeout 1200 ncu --set full --clock-control none --import-source on -k regex:wn_lat2_kernel -s 1 -c 1 -f -o $O/lat2_full_16000 python tools/lat_quick.py 64 > $O/ncu_full.txt 2>&1
tail -2 $O/smoke.txt; cut -c1-150 $O/lat_sweep.txt; cat $O/pytest_lat.txt; tail -3 $O/ncu_full.txt; ls -la $O
