#!/usr/bin/env python3
"""Build the tracked profiles/ artefacts from what tools/refresh_profiles.sh left in gpurun_out/ (run here, on the
CPU box):  python tools/make_profiles.py [round-prefix]"""
import csv, io, json, shutil, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
G, P = ROOT / "gpurun_out", ROOT / "profiles"
pre = sys.argv[1] if len(sys.argv) > 1 else "r01"
KERNELS = {"fast": "k_fast_cells", "describe": "k_describe", "pyramid_level1": "k_resize_level", "quadtree": "k_quadtree",
           "stereo_match": "k_stereo_match", "ingest": "k_gray_repitch"}

shutil.copy(G / "bench.json", P / f"{pre}_bench.json")
shutil.copy(G / "bench_reference.json", P / f"{pre}_bench_reference.json")
shutil.copy(G / "launches_bench.csv", P / f"{pre}_launches_bench.csv")

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
    traffic["kernels"][name] = {"dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                                "ncu_time": f"{r[ix['gpu__time_duration.sum']]} {units[ix['gpu__time_duration.sum']]}"}
(P / "dram_traffic.json").write_text(json.dumps(traffic, indent=1))

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
