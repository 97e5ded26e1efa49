#!/bin/bash
# One GPU session that regenerates the evidence under profiles/ (run through gpurun from the repo root; outputs in gpurun_out/final/).
mkdir -p gpurun_out/final
O=gpurun_out/final
timeout 1200 python -m pytest tests/ -q -m gpu 2>&1 | tail -25 > $O/pytest_gpu.txt
timeout 120 python __graft_entry__.py smoke > $O/smoke.txt 2>&1
timeout 600 python bench.py > $O/bench_C3.json 2> $O/bench_C3.err
timeout 400 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err
timeout 400 python bench.py --config C2 > $O/bench_C2.json 2> $O/bench_C2.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --samples 2000 --no-extra --no-cpu > $O/bench_under_ncu.txt 2>&1
N=1200 KERNELS=lat timeout 900 ncu --set full --clock-control none --import-source on -k regex:wn_lat2_kernel -s 1 -c 1 -f -o $O/lat2_full python tools/lat_quick.py 64 > $O/ncu_full.txt 2>&1
KERNELS=lat timeout 200 python tools/lat_quick.py 64 128 256 784 > $O/lat_sweep.txt 2>&1
timeout 100 python tools/lat_trace.py 64 600 > $O/trace_b64_sample600.txt 2>&1
tail -3 $O/pytest_gpu.txt; cat $O/bench_C3.json | cut -c1-600; ls -la $O
