#!/bin/bash
# Third evidence session (2 GPUs): the N=2 bench line, then the single-GPU headline again (now with roofline.traffic from the committed capture).
mkdir -p gpurun_out/final3
O=gpurun_out/final3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > $O/bench_C3_2gpu.json 2> $O/bench_C3_2gpu.err
timeout 600 python bench.py > $O/bench_C3.json 2> $O/bench_C3.err
KERNELS=lat timeout 300 python tools/lat_quick.py 64 512 640 704 > $O/lat_sweep.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -3 > $O/pytest_multi.txt
tail -1 $O/bench_C3_2gpu.json | cut -c1-300; tail -1 $O/bench_C3.json | cut -c1-200; cut -c1-150 $O/lat_sweep.txt; cat $O/pytest_multi.txt
