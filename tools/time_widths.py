"""Sweep time against the basis width m (P pulsars x n TOAs x F frequencies): which kernel family serves which width,
and what a (frequency, TOA, basis column) costs there. usage: time_widths.py [P n F]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fastfp_b200
from fastfp_b200 import synth

P, n, F = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (16, 5000, 4096)
fr = torch.tensor(synth.fp_freqs(F), dtype=torch.float64, device="cuda")
WIDTHS = [(12, 30), (7, 60), (30, 60), (120, 60), (190, 60), (380, 60), (500, 60)]
if os.environ.get("WIDTHS"):
    WIDTHS = [tuple(int(x) for x in w.split(":")) for w in os.environ["WIDTHS"].split(",")]
for n_tm, ncomps in WIDTHS:
    pta = synth.make_pta(P, n, n_tm=n_tm, ncomps=ncomps, seed=5)
    m = pta.Ts[0].shape[1]
    for path in ("auto", "fp64"):
        fp = fastfp_b200.FastFp(pta.psrs, path=path)
        a = (pta.Nvecs, pta.Ts, pta.sigmas)
        used = fp.prepare(*a).path
        if path == "fp64" and used != "fp64":
            continue
        if path == "auto" and used == "fp64":
            continue  # the fp64 pass below covers it
        fp(fr, *a); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): fp(fr, *a)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"m={m:4d} kernel {used:5s}: {ms:9.3f} ms   {ms * 1e-3 * 148 * 1.965e9 / (P * F * n):7.2f} SM-cycles per (freq, TOA)   "
              f"{ms * 1e-3 * 148 * 1.965e9 / (P * F * n * m) * 1e3:7.2f} per 1000 (freq, TOA, column)", flush=True)
