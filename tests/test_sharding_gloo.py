"""World-size-2 run of the frequency-sharding + all-gather plumbing on CPU (gloo): the same code
path bench.py and the multi-GPU callers use, with a stand-in local sweep."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fastfp_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, F, D, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    freqs = torch.from_numpy(np.linspace(2e-9, 3e-7, F))
    fp = parallel.sharded_sweep(lambda f: f * 3.0 + 1.0, freqs)
    draws = torch.arange(D, dtype=torch.float64)
    nm = parallel.sharded_sweep(lambda f: draws[:, None] * 10.0 + f[None, :] * 1e9, freqs, lead_shape=(D,))
    Dd = D + 2  # odd draw count: ragged last shard
    dd = parallel.sharded_draws(
        lambda lo, hi: torch.arange(lo, hi, dtype=torch.float64)[:, None] * 10.0 + freqs[None, :] * 1e9, Dd)
    np.save(os.path.join(out_dir, f"dd_{rank}.npy"), dd.numpy())
    np.save(os.path.join(out_dir, f"fp_{rank}.npy"), fp.numpy())
    np.save(os.path.join(out_dir, f"nm_{rank}.npy"), nm.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("F", [9, 16])
def test_sharded_sweep_world2(tmp_path, F):
    world, D = 2, 3
    mp.spawn(_worker, args=(world, _free_port(), F, D, str(tmp_path)), nprocs=world, join=True)
    freqs = np.linspace(2e-9, 3e-7, F)
    for r in range(world):
        np.testing.assert_allclose(np.load(tmp_path / f"fp_{r}.npy"), freqs * 3.0 + 1.0, rtol=0, atol=0)
        want = np.arange(D)[:, None] * 10.0 + freqs[None, :] * 1e9
        np.testing.assert_array_equal(np.load(tmp_path / f"nm_{r}.npy"), want)
        want_d = np.arange(D + 2)[:, None] * 10.0 + freqs[None, :] * 1e9
        np.testing.assert_array_equal(np.load(tmp_path / f"dd_{r}.npy"), want_d)


def test_single_process_is_passthrough():
    f = torch.arange(5, dtype=torch.float64)
    assert torch.equal(parallel.sharded_sweep(lambda x: x + 1, f), f + 1)
    assert parallel.sharded_draws(lambda lo, hi: torch.zeros(hi - lo, 4), 7).shape == (7, 4)


def test_tile_blocks_cover_the_grid_with_even_blocks():
    # two-dimensional nmfp sharding: tiles of 32 frequencies, an even number per rank, all tiles covered once
    for F, world in [(1, 1), (1, 8), (31, 2), (32, 2), (33, 2), (1000, 2), (10_000, 8), (10_000, 3), (64, 1)]:
        nt, per = parallel.tile_blocks(F, world)
        assert nt == -(-F // 32) and per % 2 == 0 and per * world >= nt and (per - 2) * world < nt
        owners = [t // per for t in range(nt)]
        assert max(owners) < world
    with pytest.raises(ValueError):
        parallel.tile_blocks(0, 2)
