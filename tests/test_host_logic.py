"""Host-side logic that needs no GPU: the vmap shim, the get_mats_* collectors, argument
validation, frequency sharding, and loud failure without a device."""
import numpy as np
import pytest

import fastfp_b200
from fastfp_b200 import _cabi, parallel, synth, vmap


def test_constants_match_reference_values():
    from fastfp_b200 import constants as c

    assert c.yr == 365.25 * 86400.0 and c.day == 86400.0 and c.fyr == 1.0 / c.yr


def test_get_mats_orders_and_sigma(monkeypatch):
    pta = synth.make_pta(2, [40, 50], n_tm=[3, 4], ncomps=5)
    Nvecs, Ts, sigmas = fastfp_b200.get_mats_fp(pta, pta.noise)
    TNTs, Nvecs2, Ts2 = fastfp_b200.get_mats_nmfp(pta, pta.noise)  # note the different order
    assert Nvecs is Nvecs2 and Ts is Ts2
    for p in range(2):
        np.testing.assert_array_equal(sigmas[p], TNTs[p] + np.diag(1.0 / pta.phis[p]))
        assert Ts[p].shape == (pta.psrs[p].toas.size, pta.n_tm[p] + 10)


def test_vmap_shim_forwards_batched_arguments():
    calls = []

    def fake(fgw, a, b, c):
        calls.append(np.shape(fgw))
        return np.asarray(fgw) * 2

    fn = vmap(fake, in_axes=(0, None, None, None))
    out = fn(np.arange(5.0), 1, 2, 3)
    np.testing.assert_array_equal(out, np.arange(5.0) * 2)
    assert calls == [(5,)]  # one batched call, not a loop
    with pytest.raises(ValueError):
        vmap(fake, in_axes=(1, None, None, None))
    with pytest.raises(ValueError):
        vmap(fake, in_axes=(None, None, 0, None))
    g = vmap(vmap(lambda f, s, a, b, c: (np.shape(f), sorted(s)), in_axes=(0, None, None, None, None)),
             in_axes=(None, 0, None, None, None))
    assert g.batched == (0, 1)
    assert g(np.zeros(3), {"x": np.zeros(2)}, 0, 0, 0) == ((3,), ["x"])
    with pytest.raises(TypeError):
        fn(1, 2)


def test_shape_validation_happens_before_any_device_work():
    pta = synth.make_pta(2, [30, 40], n_tm=3, ncomps=4)
    fp = fastfp_b200.FastFp(pta.psrs)
    bad_N = [np.eye(30), pta.Nvecs[1]]  # a block / dense N is not supported (utils.py:29-31)
    with pytest.raises(ValueError):
        fp(1e-8, bad_N, pta.Ts, pta.sigmas)
    with pytest.raises(ValueError):
        fp(1e-8, pta.Nvecs, pta.Ts, [pta.sigmas[0][:5, :5], pta.sigmas[1]])
    with pytest.raises(ValueError):
        fp(1e-8, pta.Nvecs[:1], pta.Ts, pta.sigmas)
    with pytest.raises(ValueError):
        fastfp_b200.get_xCy(pta.Nvecs[0], pta.Ts[0], pta.sigmas[0], np.zeros(3), np.zeros(30))


def test_hot_path_fails_loudly_without_a_gpu():
    if _cabi.load().fastfp_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    pta = synth.make_pta(1, 32, n_tm=3, ncomps=2)
    with pytest.raises(_cabi.FastFpError, match="no CUDA device"):
        fastfp_b200.FastFp(pta.psrs)(1e-8, pta.Nvecs, pta.Ts, pta.sigmas)
    with pytest.raises(_cabi.FastFpError, match="no CUDA device"):
        fastfp_b200.get_xCy(pta.Nvecs[0], pta.Ts[0], pta.sigmas[0], pta.psrs[0].toas, pta.psrs[0].residuals)


def test_initialize_pta_is_out_of_scope():
    from fastfp_b200.utils import initialize_pta

    with pytest.raises(NotImplementedError):
        initialize_pta([], {})


@pytest.mark.parametrize("F,world", [(10, 1), (10, 3), (7, 8), (0, 2), (1000003, 8)])
def test_shard_bounds_cover_the_grid(F, world):
    pieces = [parallel.shard_bounds(F, r, world) for r in range(world)]
    assert pieces[0][0] == 0 and pieces[-1][1] == F
    for (lo, hi, per), nxt in zip(pieces, pieces[1:]):
        assert hi == nxt[0] and 0 <= hi - lo <= per
    assert sum(hi - lo for lo, hi, _ in pieces) == F


def test_synthetic_configs_have_the_survey_shapes():
    pta = synth.make_config("C1")
    assert pta.P == 10 and pta.Ts[0].shape == (1000, 12) and pta.Ffreqs is None
    small = synth.make_pta(2, 64, n_tm=4, ncomps=3)
    assert small.Ts[0].shape == (64, 10) and small.sigmas[0].shape == (10, 10)
    again = synth.make_pta(2, 64, n_tm=4, ncomps=3)
    np.testing.assert_array_equal(small.psrs[1].residuals, again.psrs[1].residuals)  # seeded
    assert synth.fp_freqs(200)[0] == 2e-9 and synth.fp_freqs(200)[-1] == 3e-7
    d = synth.draw_samples(small, 5)
    assert d["gw_gamma"].shape == (5,) and len(d) == 2 * 2 + 2


def test_compat_package_resolves_reference_import_paths():
    import importlib
    import os
    import sys

    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "compat")
    sys.path.insert(0, root)
    try:
        for mod, names in (("fastfp.fastfp", ["FastFp"]), ("fastfp.nmfp", ["NMFP", "RN_container", "CURN_container", "GPEcorr_container"]),
                           ("fastfp.utils", ["get_xCy", "get_mats_fp", "get_mats_nmfp"]), ("fastfp.constants", ["yr", "fyr"])):
            m = importlib.import_module(mod)
            for n in names:
                assert hasattr(m, n)
        assert importlib.import_module("fastfp.fastfp").FastFp is fastfp_b200.FastFp
    finally:
        sys.path.remove(root)
        for k in [k for k in sys.modules if k == "fastfp" or k.startswith("fastfp.")]:
            del sys.modules[k]


def test_vmap_output_axes_follow_the_nesting_order_and_axes_are_validated():
    """jax.vmap semantics around the natively batched call: the reference's nesting (frequencies inside,
    draws outside; examples/run_nmfp.py:265-266) gives (D, F), the opposite nesting (F, D); an argument marked
    0 needs a leading axis, one marked None must not have one."""
    def native(fgw, samples, a, b, c):  # what NMFP.calculate_nmfp returns for batched input: (D, F)
        D = len(next(iter(samples.values())))
        return np.arange(D)[:, None] * 100.0 + np.asarray(fgw)[None, :]

    f, s = np.arange(3.0), {"x": np.zeros(2), "y": np.zeros(2)}
    vf = vmap(native, in_axes=(0, None, None, None, None))
    vg = vmap(vf, in_axes=(None, 0, None, None, None))
    assert vg(f, s, 0, 0, 0).shape == (2, 3)
    wf = vmap(native, in_axes=(None, 0, None, None, None))
    wg = vmap(wf, in_axes=(0, None, None, None, None))
    out = wg(f, s, 0, 0, 0)
    assert out.shape == (3, 2)
    np.testing.assert_array_equal(out, vg(f, s, 0, 0, 0).T)
    with pytest.raises(ValueError, match="no leading axis"):
        vg(1e-8, s, 0, 0, 0)
    with pytest.raises(ValueError, match="no leading axis"):
        vg(f, {"x": 1.0, "y": 2.0}, 0, 0, 0)
    with pytest.raises(ValueError, match="unbatched"):
        vmap(native, in_axes=(None, 0, None, None, None))(f, s, 0, 0, 0)  # frequency array, not mapped
    with pytest.raises(ValueError, match="disagree"):
        vg(f, {"x": np.zeros(2), "y": np.zeros(3)}, 0, 0, 0)
    with pytest.raises(ValueError, match="already mapped"):
        vmap(vf, in_axes=(0, None, None, None, None))


def test_pack_cache_key_sees_every_byte():
    """the pack cache key must change for ANY in-place edit of the caller's arrays."""
    from fastfp_b200.blockn import BlockNvec
    from fastfp_b200.fastfp import _fingerprint

    pta = synth.make_pta(2, [301, 257], n_tm=3, ncomps=4)
    lists = (pta.Nvecs, pta.Ts, pta.sigmas)
    k0 = _fingerprint(lists)
    assert _fingerprint(lists) == k0  # deterministic
    assert _fingerprint(([a.copy() for a in pta.Nvecs], pta.Ts, pta.sigmas)) == k0  # content, not identity
    rng = np.random.default_rng(0)
    for arr in (pta.Nvecs[1], pta.Ts[0], pta.sigmas[1]):
        flat = arr.reshape(-1)
        for _ in range(8):  # single-element edits at random positions, each one must be seen
            i = int(rng.integers(flat.size))
            old = flat[i]
            flat[i] = np.nextafter(old, np.inf)
            assert _fingerprint(lists) != k0, i
            flat[i] = old
    assert _fingerprint(lists) == k0
    # block-N objects: the diagonal part, the epoch variances and the slices all enter
    B = BlockNvec(pta.Nvecs[0].copy(), [slice(0, 4), slice(8, 11)], np.array([1e-13, 2e-13]))
    kb = _fingerprint(([B, pta.Nvecs[1]], pta.Ts, pta.sigmas))
    assert kb != k0
    B.jvec[1] *= 1.0000001
    assert _fingerprint(([B, pta.Nvecs[1]], pta.Ts, pta.sigmas)) != kb
    # the hash itself: thread-count independent block structure, seed and length sensitive
    big = rng.standard_normal(3_000_017)
    h = _cabi.hash64(big)
    assert h == _cabi.hash64(big.copy()) and h != _cabi.hash64(big, seed=1) and h != _cabi.hash64(big[:-1])
    swapped = big.copy()
    swapped[[5, 2_000_000]] = swapped[[2_000_000, 5]]
    assert _cabi.hash64(swapped) != h
    # many buffers in one call (what _fingerprint uses): identical to the per-buffer values, empty buffers included
    arrs = [big, big[:1000].copy(), np.zeros(0), rng.standard_normal((700, 9))]
    assert _cabi.hash64_many(arrs, [0, 5, 6, 7]) == [_cabi.hash64(a, seed=s) for a, s in zip(arrs, [0, 5, 6, 7])]
    assert _cabi.hash64_many([], []) == []
