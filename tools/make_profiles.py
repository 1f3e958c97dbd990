#!/usr/bin/env python3
"""Build the tracked profiles/ artefacts from what tools/refresh_profiles.sh left in gpurun_out/ (run here, on the
CPU box):  python tools/make_profiles.py [round-prefix]"""
import csv, io, json, shutil, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
G, P = ROOT / "gpurun_out", ROOT / "profiles"
pre = sys.argv[1] if len(sys.argv) > 1 else "r02"
SM_CLOCK_GHZ, SMS = 1.965, 148  # measured (MEASURED_PEAKS.json sm_max_mhz; clocks.sm during the bench)
SHARED_PEAK_GBPS = SMS * 128 * SM_CLOCK_GHZ  # one 128-byte shared-memory wavefront per SM per clock = 37.2 TB/s
KERNELS = {"fast": "k_fast_cells", "describe": "k_describe", "pyramid_level1": "k_resize_level", "quadtree": "k_quadtree",
           "stereo_match": "k_stereo_match", "ingest": "k_gray_repitch"}

for src, dst in (("bench.json", f"{pre}_bench.json"), ("bench_reference.json", f"{pre}_bench_reference.json"),
                 ("launches_bench.csv", f"{pre}_launches_bench.csv"), ("launches_cvorb.csv", f"{pre}_launches_cvorb.csv"),
                 ("cvorb_bench.json", f"{pre}_cvorb_bench.json")):
    if (G / src).exists() and "--no-bench" not in sys.argv:
        shutil.copy(G / src, P / dst)

# the cv::ORB row's kernels: text summaries only (16 images per launch: a chunk of a 64-image batch)
for k in ("k_cv_fast_score", "k_cv_nms_collect", "k_cv_harris", "k_cv_describe", "k_cv_resize_exact", "k_cv_select"):
    rep = G / f"prof_{k}.ncu-rep"
    if rep.exists():
        txt = subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_summary.py"), str(rep), "--source", "20", "--ops"],
                             capture_output=True, text=True).stdout
        (P / f"{pre}_ncu_{k}.txt").write_text(txt)

ncu_kernels = {"note": "ncu --set full --clock-control none, one launch of each kernel over a 125-pair sub-batch of the bench "
               "workload (tools/prof.sh); pct = % of peak sustained; shared_GBps = shared-memory wavefronts x 128 B / duration "
               "against 148 SM x 128 B/clk x 1.965 GHz; bound = the busiest of {dram -> hbm, l1tex/shared -> shared, issue slots -> "
               "issue, integer ALU pipe -> alu}, or latency when none reaches 50 %", "kernels": {}}
traffic = {"batch": 125, "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch from ncu --set full "
           "(tools/prof.sh), sub-batch of 125 pairs = 250 images", "kernels": {}}
for name, k in KERNELS.items():
    rep = G / f"prof_{k}.ncu-rep"
    txt = subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_summary.py"), str(rep), "--source", "25", "--ops"],
                         capture_output=True, text=True).stdout
    short = {"fast": "k_fast", "describe": "k_describe", "pyramid_level1": "k_resize", "quadtree": "k_quadtree",
             "stereo_match": "k_stereo_match", "ingest": "k_gray_repitch"}[name]
    (P / f"{pre}_ncu_{short}.txt").write_text(txt)
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, r = rows[0], rows[1], rows[2]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(key):
        v, u = float(r[ix[key]].replace(",", "")), units[ix[key]]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    def pct(key):
        return float(r[ix[key]].replace(",", "")) if key in ix and r[ix[key]] not in ("", "n/a") else None
    t_us = val("gpu__time_duration.sum") * {"usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(units[ix["gpu__time_duration.sum"]], 1.0)
    wf = pct("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") or 0.0
    d = {"dram_pct": pct("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
         "l1tex_pct": pct("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
         "lts_pct": pct("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
         "issue_pct": pct("smsp__issue_active.avg.pct_of_peak_sustained_active"),
         "sm_pct": pct("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
         "alu_pipe_pct": pct("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
         "fma_pipe_pct": pct("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
         "lsu_data_pipe_pct": pct("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
         "occupancy_pct": pct("sm__warps_active.avg.pct_of_peak_sustained_active"),
         "registers": pct("launch__registers_per_thread"),
         "duration_us": t_us, "capture_pairs": 125,
         "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
         "shared_wavefronts": wf, "shared_bank_conflicts": pct("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
         "shared_GBps": wf * 128 / (t_us * 1e-6) / 1e9 if t_us else None, "shared_peak_GBps": SHARED_PEAK_GBPS}
    d["shared_frac"] = d["shared_GBps"] / SHARED_PEAK_GBPS if d["shared_GBps"] else None
    cands = {"hbm": d["dram_pct"] or 0, "shared": d["l1tex_pct"] or 0, "issue": d["issue_pct"] or 0, "alu": d["alu_pipe_pct"] or 0}
    best = max(cands, key=cands.get)
    d["bound"] = best if cands[best] >= 50 else "latency"
    ncu_kernels["kernels"][k] = d
    traffic["kernels"][name] = {"dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                                "ncu_time": f"{r[ix['gpu__time_duration.sum']]} {units[ix['gpu__time_duration.sum']]}"}
(P / "dram_traffic.json").write_text(json.dumps(traffic, indent=1))
(P / "ncu_kernels.json").write_text(json.dumps(ncu_kernels, indent=1))

if not (G / "launches_bench.csv").exists() or "--no-bench" in sys.argv:
    sys.exit(0)
# kernel shares of the step from the launch list
rows = list(csv.reader(open(G / "launches_bench.csv")))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
tot = {}
for r in rows[h + 1:]:
    if len(r) > 10:
        n = r[4].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        tot[n] = tot.get(n, 0.0) + float(r[-1].replace(",", ""))
s = sum(tot.values())
print("kernel shares under ncu:", ", ".join(f"{k} {100 * v / s:.0f} %" for k, v in sorted(tot.items(), key=lambda kv: -kv[1])))
b = json.load(open(G / "bench.json"))
ev = {k: v["ms_per_launch"] for k, v in b["stages"].items() if k not in ("ingest_copy", "result_copy")}
se = sum(ev.values())
print("kernel shares by CUDA events:", ", ".join(f"{k} {100 * v / se:.0f} %" for k, v in sorted(ev.items(), key=lambda kv: -kv[1])))
print("value", round(b["value"]), "e2e", round(b["e2e"]["value"]), "pcie", b["e2e"].get("pcie"), "cpu", b["cpu_baseline"]["value"],
      b["cpu_baseline"]["as_shipped_2_threads"]["value"])
print("roofline", {k: b["roofline"][k] for k in ("kernel", "achieved", "frac", "traffic", "ms_per_launch")})
