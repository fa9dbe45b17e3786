"""Summarise an .ncu-rep (read here, no GPU): headline metrics of each profiled launch and the stall samples of the first
one aggregated by SASS opcode.  usage: python tools/ncu_summary.py file.ncu-rep"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "sm__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    print(r[idx["Kernel Name"]][:90])
    for w in want:
        if w in idx:
            print(f"   {w:70s} {r[idx[w]]:>16s} {units[idx[w]]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = list(csv.reader(io.StringIO(src)))
hi = [i for i, l in enumerate(lines) if l and l[0] == "Address"]
if hi:
    h = lines[hi[0]]
    body = lines[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(lines))]
    ci, si, ei = h.index("Warp Stall Sampling (All Samples)"), h.index("Source"), h.index("Instructions Executed")
    cls, ex = collections.Counter(), collections.Counter()
    for b in body:
        op = [o for o in b[si].strip().split() if not o.startswith("@")]
        op = op[0].split(".")[0] if op else ""
        cls[op] += int(b[ci] or 0)
        ex[op] += int(b[ei] or 0)
    tot = sum(cls.values())
    print("stall samples by opcode (first launch):")
    for k, v in cls.most_common(16):
        print(f"   {k:12s} {v:7d} {100 * v / tot:5.1f}%   executed {ex[k]}")
