"""Summarise an ncu report of solve_pass: headline metrics, stall reasons, and stall samples per
kernel phase (phases are delimited by BAR.SYNC in the SASS).  Usage: python tools/ncu_phases.py rep.ncu-rep"""
import csv
import subprocess
import sys


def f(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return 0.0


rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed_op_shared_ld.sum",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print("%-72s %-8s %s" % (w, units[i], [r[i] for r in data]))
for i, hn in enumerate(hdr):
    if "average_warps_issue_stalled" in hn and hn.endswith("_per_issue_active.ratio") and "not_issued" not in hn:
        v = [f(r[i]) for r in data]
        if max(v) > 0.15:
            print("%-72s %s" % (hn.replace("smsp__average_warps_issue_stalled_", "stall ").replace("_per_issue_active.ratio", ""), v))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h2 = rows[hi]
recs = [dict(zip(h2, r)) for r in rows[hi + 1:] if len(r) == len(h2)]
# first kernel instance only
first = []
seen = set()
for r in recs:
    if r["Address"] in seen:
        break
    seen.add(r["Address"])
    first.append(r)
recs = first
tot = sum(f(r["# Samples"]) for r in recs) or 1
stall_cols = [c for c in h2 if c.startswith("stall_") and "(Not Issued)" not in c]
seg, segs = None, []
def new(i):
    return {"start": i, "n": 0, "samples": 0.0, "inst": 0.0, **{c: 0.0 for c in stall_cols}}
seg = new(0)
for i, r in enumerate(recs):
    seg["n"] += 1
    seg["samples"] += f(r["# Samples"])
    seg["inst"] += f(r["Instructions Executed"])
    for c in stall_cols:
        seg[c] += f(r[c])
    if "BAR.SYNC" in r["Source"]:
        seg["end"] = i
        segs.append(seg)
        seg = new(i + 1)
seg["end"] = len(recs)
segs.append(seg)
print("\nphase (SASS range)      static   samples   share   warp-inst   top stalls")
for s in segs:
    top = sorted(((s[c], c) for c in stall_cols), reverse=True)[:4]
    print("[%5d..%5d] %8d %9.0f %6.1f%% %11.0f   %s" % (s["start"], s["end"], s["n"], s["samples"], 100 * s["samples"] / tot, s["inst"],
          ", ".join("%s %.0f" % (c.replace("stall_", ""), v) for v, c in top if v > 0)))
print("\ntop instructions by samples")
for r in sorted(recs, key=lambda r: -f(r["# Samples"]))[:25]:
    top = sorted(((f(r[c]), c) for c in stall_cols), reverse=True)[:2]
    print("%6s %-58s %6.0f exec=%8.0f  %s" % (r["Address"][-5:], r["Source"][:58], f(r["# Samples"]), f(r["Instructions Executed"]),
          ", ".join("%s %.0f" % (c.replace("stall_", ""), v) for v, c in top if v > 0)))
