"""Turn an .ncu-rep of the sweep kernel into the markdown summary kept under profiles/.
usage: ncu_summary.py REPORT.ncu-rep OUT.md "capture description" """
import collections, csv, io, subprocess, sys
rep, out_path, desc = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
M = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
keys = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.max',
        'sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'lts__t_bytes.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'sm__ops_path_tensor_src_fp64.sum']
out = ["# ncu summary -- round 1, fp_sweep_kernel (warp-specialised, fp64 MMA)", "", desc, "",
       "| metric | value | unit |", "|---|---|---|"]
for k in keys:
    if k in M:
        out.append(f"| `{k}` | {M[k][0]} | {M[k][1]} |")
d = {}
for h, (v, u) in M.items():
    if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h:
        try:
            d[h.split('stalled_')[1]] = float(v)
        except ValueError:
            pass
tot = sum(d.values())
out += ["", "Warp-state samples (all warps, issue-stall reasons):", "", "| reason | share |", "|---|---|"]
for k, v in sorted(d.items(), key=lambda kv: -kv[1])[:9]:
    out.append(f"| {k} | {v / tot * 100:.1f}% |")
rows = list(csv.reader(io.StringIO(src)))
h2 = rows[1]; H = {h: i for i, h in enumerate(h2)}; data = rows[2:]
def g(r, k):
    try:
        return float(r[H[k]])
    except ValueError:
        return 0.0
agg = collections.defaultdict(lambda: [0, 0])
for r in data:
    t = r[H['Source']].split()
    op = (t[1] if t and t[0].startswith('@') and len(t) > 1 else (t[0] if t else '')).split('.')[0]
    agg[op][0] += g(r, '# Samples'); agg[op][1] += g(r, 'Instructions Executed')
ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
out += ["", "SASS opcode mix (warp-level instructions executed) and where the samples sit:", "",
        "| opcode | executed | share of instructions | share of samples |", "|---|---|---|---|"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    out.append(f"| {k} | {v[1]:.3g} | {v[1] / ti * 100:.1f}% | {v[0] / ts * 100:.1f}% |")
f = lambda k: float(M[k][0])
out += ["", "Reading: DMMA (`mma.sync.m8n8k4.f64`) and DFMA share one pipe (`sm__pipe_shared`); it is busy "
        f"{f('sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active'):.1f}% of the time, "
        f"{f('sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active'):.1f}% with the contraction and "
        f"{f('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'):.1f}% with the sincos/weighted-sum work of the producers. "
        f"DRAM traffic is {f('dram__bytes_read.sum'):.1f} {M['dram__bytes_read.sum'][1]} read + {f('dram__bytes_write.sum'):.1f} "
        f"{M['dram__bytes_write.sum'][1]} written for the whole launch: the packed pulsar arrays are read from HBM once and then "
        "served from L2 (compulsory traffic only). SASS evidence of the Blackwell path: `UBLKCP` (TMA bulk copy), "
        "`SYNCS.*` (mbarrier), `USETMAXREG`, `DMMA.8x8x4`."]
open(out_path, "w").write("\n".join(out) + "\n")
print("\n".join(out[6:26]))
