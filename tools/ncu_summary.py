#!/usr/bin/env python3
"""Compact per-kernel summary of an .ncu-rep (run here on the CPU box): python tools/ncu_summary.py rep [--source]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("sm__inst_executed.sum", "warp_inst"), ("smsp__inst_executed.sum", "warp_inst"),
        ("sm__inst_executed_pipe_alu.sum", "alu"), ("sm__inst_executed_pipe_lsu.sum", "lsu"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_registers", "lim_reg"),
        ("launch__occupancy_limit_shared_mem", "lim_smem"), ("launch__occupancy_limit_warps", "lim_warps"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "gld_sectors"),
        ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "gld_requests"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst")]
STALLS = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
for r in rows[2:]:
    print("==", r[idx["Kernel Name"]][:60], r[idx.get("Grid Size", 0)])
    print("   " + "  ".join(f"{n}={r[idx[k]]}{units[idx[k]] if units[idx[k]] not in ('', '%') else ''}" for k, n in KEYS if k in idx))
    st = sorted(((float(r[idx[h]].replace(",", "") or 0), h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for h in STALLS), reverse=True)
    tot = sum(v for v, _ in st) or 1
    print("   stalls: " + "  ".join(f"{n}={100*v/tot:.0f}%" for v, n in st[:8]))
if "--source" in sys.argv:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    agg, h, fname, pos = [], None, "", {}
    for r in rows:
        if len(r) >= 2 and r[0] == "File Name":
            fname = r[1].split("/")[-1]
        elif len(r) > 5 and r[0] == "Line No":
            h = r
            ci = h.index("# Samples"); ii = h.index("Instructions Executed")
        elif h and len(r) == len(h) and r[0].isdigit():
            try:
                sm, ie = float(r[ci] or 0), float(r[ii] or 0)
            except ValueError:
                continue
            if sm or ie:
                key = (fname, r[0])
                if key in pos:
                    a = agg[pos[key]]; agg[pos[key]] = (a[0] + sm, a[1] + ie, a[2], a[3], a[4] or r[1].strip()[:100])
                else:
                    pos[key] = len(agg); agg.append((sm, ie, fname, r[0], r[1].strip()[:100]))
    tot = sum(a[0] for a in agg) or 1
    toti = sum(a[1] for a in agg) or 1
    print(f"  total samples {tot:.0f}, warp instructions {toti:.0f}")
    for sm, ie, fn, ln, txt in sorted(agg, reverse=True)[:int(sys.argv[sys.argv.index("--source") + 1]) if sys.argv[-1].isdigit() else 25]:
        print(f"  {100*sm/tot:5.1f}% smp {100*ie/toti:5.1f}% inst  {fn}:{ln}: {txt}")
if "--ops" in sys.argv:
    import re
    sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(sass)))
    h = next(r for r in rows if "Instructions Executed" in r)
    ii, si = h.index("Instructions Executed"), h.index("Source")
    ops, tot = collections.Counter(), 0.0
    for r in rows:
        if len(r) != len(h) or r is h:
            continue
        try:
            n = float(r[ii] or 0)
        except ValueError:
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[si])
        ops[m.group(2).split(".")[0] if m else "?"] += n
        tot += n
    print(f"  dynamic warp instructions {tot:.0f}; by opcode: " + "  ".join(f"{o}={100*n/tot:.1f}%" for o, n in ops.most_common(18)))
