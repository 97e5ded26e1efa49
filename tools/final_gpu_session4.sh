#!/bin/bash
# Last evidence session of the round at HEAD: the whole GPU suite, smoke, the headline line, the fp32 lines of the C3 model.
mkdir -p gpurun_out/final4
O=gpurun_out/final4
timeout 1200 python -m pytest tests/ -q -m gpu 2>&1 | tail -8 > $O/pytest_gpu.txt
timeout 120 python __graft_entry__.py smoke > $O/smoke.txt 2>&1
timeout 600 python bench.py > $O/bench_C3.json 2> $O/bench_C3.err
timeout 400 python bench.py --dtype fp32 --samples 8000 --steps 3 --warmup 3 --no-cpu > $O/bench_C3_fp32.json 2> $O/bench_C3_fp32.err
timeout 400 python bench.py --dtype fp32fast --samples 8000 --steps 3 --warmup 3 --no-cpu > $O/bench_C3_fp32fast.json 2> $O/bench_C3_fp32fast.err
tail -3 $O/pytest_gpu.txt; tail -1 $O/smoke.txt | cut -c1-200; for f in bench_C3 bench_C3_fp32 bench_C3_fp32fast; do tail -1 $O/$f.json | cut -c1-260; done
