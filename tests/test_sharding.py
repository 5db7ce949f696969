"""N > 1 path on CPU: world_size-2/3 gloo runs of the sharding + gather plumbing.  The per-rank
"compute" here is the CPU oracle (tests may use it); the product's compute is CUDA-only."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from planet_b200.sharding import gather_patches, shard_range


def test_shard_range_tiles_the_leaf_range():
    for n in (0, 1, 5, 96, 98304, 16384):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(98304, 3, 8) == (36864, 49152)        # C3 on 8 GPUs: 12 288 quads each
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_quads, dim, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.bindings import FBM, PortOracle, height_params
    orc = PortOracle()
    quads = np.concatenate([orc.uniform_quads(f, 2) for f in range(6)])[:n_quads]
    lo, hi = shard_range(n_quads, rank, world)
    hp = height_params(kind=FBM, gain=0.5, fixed_octaves=4)
    local = torch.from_numpy(orc.generate_height_maps(quads[lo:hi], dim, 18, hp))
    full = gather_patches(local, n_units=n_quads)
    if rank == 0:
        np.save(os.path.join(out_dir, f"gathered_{world}_{n_quads}.npy"), full.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_quads", [(2, 96), (2, 37), (3, 50)])
def test_gloo_gather_equals_unsharded(tmp_path, world, n_quads):
    """Sharded + gathered buffer is byte-identical to the 1-rank buffer (equal and ragged shards)."""
    from oracle.bindings import FBM, PortOracle, height_params
    dim = 8
    mp.spawn(_worker, args=(world, _free_port(), n_quads, dim, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / f"gathered_{world}_{n_quads}.npy")
    orc = PortOracle()
    quads = np.concatenate([orc.uniform_quads(f, 2) for f in range(6)])[:n_quads]
    want = orc.generate_height_maps(quads, dim, 18, height_params(kind=FBM, gain=0.5, fixed_octaves=4))
    assert got.tobytes() == want.tobytes()
