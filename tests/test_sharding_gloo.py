"""N>1 path on CPU: world_size-2 gloo processes.  Rank 0 owns the weights and broadcasts them, every rank runs its
batch shard (here with the CPU oracle standing in for the GPU kernel -- the sharding logic is what is under test),
results are gathered and must equal the unsharded run: utterances never interact."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from nv_wavenet_b200 import sharding


def test_shard_ranges_cover_batch():
    for batch in (1, 2, 7, 64, 65, 512):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, shape, out_dir):
    import torch.distributed as dist
    from oracle import pyoracle as po
    from tests import refgen
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    R, S, A, L, B, N, md = shape
    full = refgen.lively_inputs(5, R, S, A, L, B, N)            # every rank can regenerate the inputs ...
    keys = [k for k in full if k not in ("Lh", "selectors")]
    w = {k: (full[k] if rank == 0 else np.zeros_like(full[k])) for k in keys}   # ... but only rank 0 has the weights
    w = sharding.broadcast_weights(w, src=0)
    Lh, sel = sharding.shard_inputs(full["Lh"], full["selectors"], rank, world)
    lo, hi = sharding.shard_range(B, rank, world)
    o = po.Oracle(L, hi - lo, N, R, S, A, md, math=po.MATH_PORTABLE)
    o.load(w); o.set_inputs(Lh, sel)
    y = o.run(N, hi - lo)
    y_all = sharding.gather_outputs(y, B, world)
    if rank == 0:
        np.save(os.path.join(out_dir, "y_all.npy"), y_all)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_equals_unsharded(world, tmp_path):
    from oracle import pyoracle as po
    from tests import refgen
    shape = (32, 128, 256, 4, 5, 12, 4)
    R, S, A, L, B, N, md = shape
    mp.spawn(_worker, args=(world, _free_port(), shape, str(tmp_path)), nprocs=world, join=True)
    y_all = np.load(tmp_path / "y_all.npy")
    full = refgen.lively_inputs(5, R, S, A, L, B, N)
    o = po.Oracle(L, B, N, R, S, A, md, math=po.MATH_PORTABLE)
    o.load(full); o.set_inputs(full["Lh"], full["selectors"])
    assert np.array_equal(y_all, o.run(N, B))
