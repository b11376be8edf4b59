"""A ``jax.vmap``-shaped shim so caller scripts change one import.

The reference batches by wrapping the per-frequency methods:
``jax.vmap(Fp_obj.calculate_Fp, in_axes=(0, None, None, None))`` (``examples/run_fp.py:63``) and
``vmap_f = jax.vmap(nmfp, in_axes=(0, None, None, None, None))``,
``vmap_g = jax.vmap(vmap_f, in_axes=(None, 0, None, None, None))`` (``examples/run_nmfp.py:265-266``).
The engine's methods take the batched arguments natively (an array of frequencies, a dict of
``(D,)`` arrays), so ``vmap`` only has to check ``in_axes`` and forward: the whole batch is one
kernel launch, not a loop.
"""
from __future__ import annotations


class _Batched:
    def __init__(self, fn, in_axes, batched):
        self.fn, self.in_axes, self.batched = fn, in_axes, batched

    def __call__(self, *args):
        if len(args) != len(self.in_axes):
            raise TypeError(f"expected {len(self.in_axes)} positional arguments, got {len(args)}")
        return self.fn(*args)


def vmap(fn, in_axes=0):
    """Batch ``fn`` (a ``FastFp`` / ``NMFP`` object or bound method, or a previous ``vmap`` of
    one) over leading axes. Only ``0`` / ``None`` entries are meaningful: axis 0 of ``fgw``
    (frequencies) and of the ``samples`` dict values (noise draws)."""
    if isinstance(in_axes, int):
        raise TypeError("in_axes must be a tuple with one entry per positional argument")
    in_axes = tuple(in_axes)
    for ax in in_axes:
        if ax not in (0, None):
            raise ValueError("only in_axes entries 0 and None are supported")
    inner = fn.batched if isinstance(fn, _Batched) else ()
    batched = tuple(sorted(set(inner) | {i for i, ax in enumerate(in_axes) if ax == 0}))
    target = fn.fn if isinstance(fn, _Batched) else fn
    if any(i >= 2 for i in batched):
        raise ValueError("only fgw (argument 0) and samples (argument 1) can be batched")
    return _Batched(target, in_axes, batched)
