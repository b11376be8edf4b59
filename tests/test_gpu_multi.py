"""Two-GPU check of the sharded sweeps (skipped with fewer than 2 devices): one process per GPU over NCCL, the
same code path bench.py uses (fastfp_b200.parallel), and the gathered result compared BIT FOR BIT with what one
GPU computes for the whole grid / the whole draw batch (the multi-GPU result was never checked)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    import fastfp_b200
    from fastfp_b200 import NMFP, CURN_container, RN_container, parallel, synth

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    pta = synth.make_pta(5, [1500, 2203, 997, 3000, 1801], n_tm=[8, 12, 10, 12, 6], ncomps=30, seed=123)
    F = 10_007  # odd: the last shard is one bin short
    freqs = torch.from_numpy(synth.fp_freqs(F)).to(dev)
    fp = fastfp_b200.FastFp(pta.psrs, device=rank)
    mats = (pta.Nvecs, pta.Ts, pta.sigmas)
    gathered = parallel.sharded_sweep(lambda f: fp.calculate_Fp(f, *mats), freqs)
    single = fp.calculate_Fp(freqs, *mats)
    np.save(os.path.join(out_dir, f"fp_gathered_{rank}.npy"), gathered.cpu().numpy())
    np.save(os.path.join(out_dir, f"fp_single_{rank}.npy"), single.cpu().numpy())
    # noise-marginalised: draws sharded
    curn = CURN_container(pta.Ffreqs)
    sigs = [RN_container(q, Ffreqs=pta.Ffreqs, add_curn=True, curn_container=curn) for q in pta.psrs]
    nm = NMFP(pta.psrs, sigs, device=rank)
    nmats = (pta.Nvecs, pta.Ts, pta.TNTs)
    D = 37
    samples = synth.draw_samples(pta, D)
    fn = torch.from_numpy(synth.nmfp_freqs(150, pta.Tspan) * 1.003).to(dev)
    lo, hi, _ = parallel.shard_bounds(D, rank, world)
    mine = {k: v[lo:hi] for k, v in samples.items()}
    g2 = parallel.sharded_draws(lambda a, b: nm(fn, mine, *nmats), D)
    s2 = nm(fn, samples, *nmats)
    # two-dimensional form: the draw-independent stage computed in frequency slices and all-gathered
    g3 = parallel.sharded_draws(lambda a, b: nm.calculate_nmfp_2d(fn, mine, *nmats), D)
    np.save(os.path.join(out_dir, f"nm2d_gathered_{rank}.npy"), g3.cpu().numpy())
    # fewer draws than ranks and a grid of two tiles: rank 1 has no draw and only padding tiles besides its own
    one = {k: v[:1] for k, v in samples.items()}
    lo1, hi1, _ = parallel.shard_bounds(1, rank, world)
    mine1 = {k: v[lo1:hi1] for k, v in one.items()}
    g4 = parallel.sharded_draws(lambda a, b: nm.calculate_nmfp_2d(fn[:40], mine1, *nmats), 1)
    np.save(os.path.join(out_dir, f"nm2d_small_{rank}.npy"), g4.cpu().numpy())
    np.save(os.path.join(out_dir, f"nm_small_single_{rank}.npy"), nm(fn[:40], one, *nmats).cpu().numpy())
    np.save(os.path.join(out_dir, f"nm_gathered_{rank}.npy"), g2.cpu().numpy())
    np.save(os.path.join(out_dir, f"nm_single_{rank}.npy"), s2.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_gather_equals_single_gpu_sweep(tmp_path):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices (run with gpurun --gpus 2)")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    ref = np.load(tmp_path / "fp_single_0.npy")
    assert ref.shape == (10_007,) and np.all(np.isfinite(ref)) and ref.min() > 0
    refn = np.load(tmp_path / "nm_single_0.npy")
    assert refn.shape == (37, 150) and np.all(np.isfinite(refn))
    for r in range(world):
        np.testing.assert_array_equal(np.load(tmp_path / f"fp_gathered_{r}.npy"), ref)
        np.testing.assert_array_equal(np.load(tmp_path / f"fp_single_{r}.npy"), ref)  # the two GPUs agree
        np.testing.assert_array_equal(np.load(tmp_path / f"nm_gathered_{r}.npy"), refn)
        np.testing.assert_array_equal(np.load(tmp_path / f"nm_single_{r}.npy"), refn)
        np.testing.assert_array_equal(np.load(tmp_path / f"nm2d_gathered_{r}.npy"), refn)  # 2-D sharding: same bits
        small = np.load(tmp_path / f"nm_small_single_{r}.npy")
        assert small.shape == (1, 40)
        np.testing.assert_array_equal(np.load(tmp_path / f"nm2d_small_{r}.npy"), small)
