#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) per kernel: duration, pipe utilisation, DRAM bytes,
stall mix and the hottest opcodes.  Usage: python tools/ncu_summary.py report.ncu-rep [--ops]"""
import collections
import csv
import subprocess
import sys


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def src(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    ks, cur = [], None
    for r in csv.reader(out.splitlines()):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            ks.append(cur)
        elif r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    return ks


WANT = [("gpu__time_duration.sum", "time"), ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
        ("smsp__inst_executed.sum", "inst"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]


def main():
    rep = sys.argv[1]
    hdr, units, rows = raw(rep)
    for r in rows:
        print("=" * 100)
        print(r[hdr.index("Kernel Name")])
        print("  " + "  ".join(f"{n}={r[hdr.index(k)]}{units[hdr.index(k)] if n in ('time','dram_rd','dram_wr') else ''}"
                               for k, n in WANT if k in hdr))
        st = [(h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(r[i]))
              for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("ratio")]
        print("  stalls/issue: " + "  ".join(f"{k}={v:.2f}" for k, v in sorted(st, key=lambda kv: -kv[1])[:7]))
    if "--ops" in sys.argv:
        for k in src(rep):
            h = k["hdr"]
            iS, iN, iE = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
            cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
            tot = sum(int(r[iN]) for r in k["rows"]) or 1
            agg, ex, st = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
            for r in k["rows"]:
                toks = [o for o in r[iS].strip().split() if not o.startswith("@")]
                op = toks[0].split(".")[0] if toks else "?"
                agg[op] += int(r[iN])
                ex[op] += int(r[iE])
                for c in cols:
                    st[op][c] += int(r[h.index(c)])
            print("-" * 100)
            print(k["name"], "total executed %.1fM" % (sum(ex.values()) / 1e6))
            for op, n in agg.most_common(12):
                top = ", ".join(f"{c[6:]}={v / max(1, n):.2f}" for c, v in st[op].most_common(3))
                print(f"  {op:10s} samples {n / tot:6.3f}  executed {ex[op] / 1e6:9.1f}M   {top}")


if __name__ == "__main__":
    main()
