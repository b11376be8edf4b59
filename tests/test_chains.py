"""Chain -> samples-dict plumbing of the nmfp driver (reference examples/run_nmfp.py:174-186, 250-261)."""
import numpy as np
import pytest

from fastfp_b200 import chains


def test_map_params_scalar_and_batched():
    names = ["a_log10_A", "a_gamma", "gw_log10_A"]
    one = chains.map_params(names, np.array([1.0, 2.0, 3.0]))
    assert one["a_gamma"] == 2.0 and np.ndim(one["a_gamma"]) == 0
    xs = np.arange(12.0).reshape(3, 4)
    many = chains.map_params(names, xs)
    np.testing.assert_array_equal(many["gw_log10_A"], xs[2])
    with pytest.raises(ValueError):
        chains.map_params(names, np.zeros(2))


def test_draws_from_chain_respects_burn_in_and_metadata_columns(tmp_path):
    names = ["p0", "p1", "p2"]
    rng = np.random.default_rng(0)
    full = {n: rng.standard_normal(40) for n in names}
    path = tmp_path / "chain_1.txt"
    chains.write_chain(path, full, names)
    raw = np.loadtxt(path)
    assert raw.shape == (40, 3 + chains.N_META_COLUMNS)
    samples, idxs = chains.draws_from_chain(str(path), names, 12, rng=np.random.default_rng(1))
    assert len(set(idxs.tolist())) == 12 and idxs.min() >= 10  # distinct rows, first quarter discarded
    for k, n in enumerate(names):
        np.testing.assert_allclose(samples[n], raw[idxs, k], rtol=0, atol=0)
    with pytest.raises(ValueError):
        chains.draws_from_chain(raw, names, 31)
    with pytest.raises(ValueError):
        chains.draws_from_chain(raw, names[:2], 5)


def test_draw_batches_cover_all_draws_in_order():
    s = {"a": np.arange(10.0), "b": np.arange(10.0) * 2}
    parts = list(chains.draw_batches(s, 4))
    assert [len(p["a"]) for p in parts] == [4, 4, 2]
    np.testing.assert_array_equal(np.concatenate([p["b"] for p in parts]), s["b"])
