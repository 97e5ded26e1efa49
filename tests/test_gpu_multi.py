"""Two GPUs, two processes (NCCL): rank 0 owns the weights, the packed weight blob is broadcast once (the product's only
collective), every rank generates its batch shard; each shard must equal the same utterances of a single-GPU run.
Skipped on boxes with fewer than two GPUs."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHAPE = (64, 256, 256, 20, 40, 300, 512)          # R, S, A, L, B, N, maxDil: C3 model, 40 utterances -> shards of 20 (ragged 16-tiles)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, dtype_name, out_dir):
    import torch
    import torch.distributed as dist
    import nv_wavenet_b200 as nw
    from nv_wavenet_b200 import sharding
    from tests import refgen
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    R, S, A, L, B, N, md = SHAPE
    full = refgen.lively_inputs(9, R, S, A, L, B, N)
    lo, hi = sharding.shard_range(B, rank, world)
    Lh, sel = sharding.shard_inputs(full["Lh"], full["selectors"], rank, world)
    e = nw.NVWavenetInfer(L, md, hi - lo, N, R=R, S=S, A=A, dtype=nw.FP16 if dtype_name == "fp16" else nw.FP32)
    if rank == 0:
        e.load(full)                                                    # only rank 0 uploads weights
    ptr, nbytes = e.weight_blob()

    class _Blob:
        __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
    blob = torch.as_tensor(_Blob(), device=torch.device("cuda", rank))
    dist.broadcast(blob, 0)                                             # NCCL over NVLink
    e.weights_updated()
    e.set_inputs(Lh, sel)
    y = np.zeros((hi - lo, N), np.int32)
    e.run(N, hi - lo, y); e.synchronize()
    np.save(os.path.join(out_dir, f"y_rank{rank}.npy"), y)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("dtype_name", ["fp16", "fp32"])
def test_two_gpu_shards_equal_single_gpu_run(dtype_name, tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import nv_wavenet_b200 as nw
    from nv_wavenet_b200 import sharding
    from tests import refgen
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), dtype_name, str(tmp_path)), nprocs=world, join=True)
    R, S, A, L, B, N, md = SHAPE
    full = refgen.lively_inputs(9, R, S, A, L, B, N)
    e = nw.NVWavenetInfer(L, md, B, N, R=R, S=S, A=A, dtype=nw.FP16 if dtype_name == "fp16" else nw.FP32)
    e.load(full); e.set_inputs(full["Lh"], full["selectors"])
    y = np.zeros((B, N), np.int32)
    e.run(N, B, y); e.synchronize()
    assert len(np.unique(y)) > 16
    for r in range(world):
        lo, hi = sharding.shard_range(B, r, world)
        assert np.array_equal(np.load(tmp_path / f"y_rank{r}.npy"), y[lo:hi]), f"rank {r} differs from the single-GPU run"
