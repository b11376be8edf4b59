"""Clocks/power under sustained fp64 load: DFMA microbench loop and repeated C2 sweeps."""
import os, subprocess, sys, time, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fastfp_b200
from fastfp_b200 import synth, _cabi

Q = "clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.hw_thermal_slowdown"
def sample(tag, dur):
    p = subprocess.Popen(["nvidia-smi", f"--query-gpu={Q}", "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
    time.sleep(dur)
    p.terminate()
    lines = [l.strip() for l in p.stdout.read().strip().splitlines() if l.strip()]
    print(tag, "samples:", len(lines))
    for l in lines[:: max(1, len(lines) // 12)]: print("   ", l)

def load_dfma(sec):
    t0 = time.time()
    while time.time() - t0 < sec:
        tf, ms = _cabi.fp64_peak(0, 200000)
        print(f"   dfma {tf:.2f} TFLOP/s ({ms:.1f} ms)", flush=True)

th = threading.Thread(target=sample, args=("DFMA loop", 5.0)); th.start()
load_dfma(4.5); th.join()
time.sleep(3)
pta = synth.make_config("C2")
fp = fastfp_b200.FastFp(pta.psrs)
fr = torch.tensor(synth.fp_freqs(10000), dtype=torch.float64, device="cuda")
fp(fr, pta.Nvecs, pta.Ts, pta.sigmas); torch.cuda.synchronize()
time.sleep(3)
def load_sweep(sec):
    t0 = time.time(); k = 0
    while time.time() - t0 < sec:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4): fp(fr, pta.Nvecs, pta.Ts, pta.sigmas)
        e1.record(); torch.cuda.synchronize()
        if k % 3 == 0: print(f"   sweep {e0.elapsed_time(e1)/4:.2f} ms/sweep", flush=True)
        k += 1
th = threading.Thread(target=sample, args=("C2 sweep loop", 5.0)); th.start()
load_sweep(4.5); th.join()
