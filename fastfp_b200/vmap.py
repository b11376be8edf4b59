"""A ``jax.vmap``-shaped shim so caller scripts change one import.

The reference batches by wrapping the per-frequency methods:
``jax.vmap(Fp_obj.calculate_Fp, in_axes=(0, None, None, None))`` (``examples/run_fp.py:63``) and
``vmap_f = jax.vmap(nmfp, in_axes=(0, None, None, None, None))``,
``vmap_g = jax.vmap(vmap_f, in_axes=(None, 0, None, None, None))`` (``examples/run_nmfp.py:265-266``).
The engine's methods take the batched arguments natively (an array of frequencies, a dict of
``(D,)`` arrays) and the whole batch is one sweep on the device, not a loop; what ``vmap`` adds is
``jax.vmap``'s contract around that call:

* the argument marked ``0`` must actually carry a leading axis, an argument marked ``None`` must not
  (a frequency array passed where ``in_axes`` says "unbatched" is an error, as it is in the reference,
  whose per-frequency body cannot take one);
* the output axes follow the nesting order, outermost ``vmap`` first: the reference's nesting
  (frequencies inside, draws outside) gives ``(D, F)``; the opposite nesting gives ``(F, D)``.
"""
from __future__ import annotations

import numpy as np


def _leading(x):
    """leading-axis length of a batched argument (array / tensor, or a dict of them), or None"""
    if isinstance(x, dict):
        lens = {_leading(v) for v in x.values()}
        lens.discard(None)
        if len(lens) > 1:
            raise ValueError(f"batched dict entries disagree on the leading axis: {sorted(lens)}")
        return lens.pop() if lens else None
    shape = getattr(x, "shape", None)
    if shape is None:
        shape = np.shape(x)
    return int(shape[0]) if len(shape) >= 1 else None


class _Batched:
    """``order``: batched positional arguments from the INNERMOST ``vmap`` to the outermost."""

    def __init__(self, fn, nargs, order):
        self.fn, self.nargs, self.order = fn, nargs, tuple(order)

    @property
    def batched(self):
        return tuple(sorted(self.order))

    def __call__(self, *args):
        if len(args) != self.nargs:
            raise TypeError(f"expected {self.nargs} positional arguments, got {len(args)}")
        for i in (0, 1)[: min(2, self.nargs)]:
            if i == 1 and not isinstance(args[1], dict):
                continue  # FastFp: argument 1 is the Nvecs list, never batched
            lead = _leading(args[i])
            name = "fgw" if i == 0 else "samples"
            if i in self.order and lead is None:
                raise ValueError(f"vmap: {name} is mapped over axis 0 but has no leading axis")
            if i not in self.order and lead is not None:
                raise ValueError(f"vmap: {name} has a leading axis but in_axes marks it as unbatched")
        out = self.fn(*args)
        # native layout with both axes batched is (D, F) (draw-major, the reference's nesting); if the
        # frequency vmap is the OUTER one, jax would return (F, D)
        if len(self.order) == 2 and self.order[-1] == 0:
            out = out.T if hasattr(out, "T") else np.asarray(out).T
        return out


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
    mine = [i for i, ax in enumerate(in_axes) if ax == 0]
    if any(i >= 2 for i in mine):
        raise ValueError("only fgw (argument 0) and samples (argument 1) can be batched")
    if len(mine) > 1:
        raise ValueError("one vmap maps one argument here (nest two vmaps like examples/run_nmfp.py:265-266)")
    inner = fn.order if isinstance(fn, _Batched) else ()
    if isinstance(fn, _Batched) and fn.nargs != len(in_axes):
        raise ValueError("in_axes must have one entry per positional argument of the wrapped function")
    if set(inner) & set(mine):
        raise ValueError("argument already mapped by an inner vmap")
    target = fn.fn if isinstance(fn, _Batched) else fn
    return _Batched(target, len(in_axes), tuple(inner) + tuple(mine))
