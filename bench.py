#!/usr/bin/env python3
"""bench.py — stereo pairs/sec (ORB both views + stereo match) on synthetic KITTI-shaped 1242x375 pairs.

  python bench.py [--gpus N] [--steps K] [--warmup W]           our CUDA path (default N=1)
  python bench.py --impl reference [...]                        the reference's CPU path (oracle port)
  python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...   one rank per GPU

Workload (BASELINE.json configs[1]): a 1000-frame synthetic KITTI-tracking-shaped stereo sequence with
PointSLOT's Examples/Stereo/0000-0013.yaml extractor settings ORBextractor(1000, 1.2, 8, 20, 5),
bf = 384.38148, fx = 721.5377.  One "step" = the whole sequence once through the hot path (both
extractions + ComputeStereoMatches per pair), processed in sub-batches.  With N GPUs every rank
processes its own 1000-pair shard (weak scaling, no collective on the data path).

Printed JSON (rank 0): `value` = pairs/s with inputs resident in HBM; `e2e` = pairs/s through the C ABI
with pinned HOST input and output buffers (H2D + D2H inside the timed region); `roofline` = the
dominant kernel's algorithmic bytes / its CUDA-event duration vs the measured HBM peak;
`cpu_baseline` = the oracle (CPU port of the reference) timed on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "stereo pairs/sec (ORB both views + stereo match) KITTI 1242x375"
UNIT = "pairs/s"
W, H = 1242, 375
PARAMS = (1000, 1.2, 8, 20, 5)  # Examples/Stereo/0000-0013.yaml:41-54
BF, FX = 384.38148, 721.5377    # 0000-0013.yaml:26, :8
WORKLOAD = "1000-frame synthetic KITTI-tracking-shaped stereo sequence, ORBextractor(1000,1.2,8,20,5) [0000-0013.yaml]"


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def max_disparity() -> float:
    mb = np.float32(BF) / np.float32(FX)
    return float(np.float32(BF) / mb)


# ------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi during the timed regions
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.rows = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.idx)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, windows):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for t, line in self.rows:
            if not any(a <= t <= b for a, b in windows):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1])); smax = float(f[2])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (port of the reference's CPU path) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_pairs_per_sec(Ls, Rs, threads: int, seconds: float, per_pair_threads: int = 1):
    """Time the oracle on the given pairs with `threads` worker threads (ctypes releases the GIL).
    per_pair_threads=2 reproduces the reference as shipped (left/right extraction on two threads,
    Frame.cc:161-164; stereo matching serial)."""
    import oracle as O

    O.build()
    maxD = max_disparity()
    n = len(Ls)
    lock = threading.Lock()
    state = {"next": 0, "done": 0}
    t_end = [0.0]
    deadline = time.time() + seconds

    def worker():
        eL, eR = O.OracleExtractor(*PARAMS), O.OracleExtractor(*PARAMS)
        while True:
            with lock:
                if time.time() >= deadline:
                    return
                i = state["next"]
                state["next"] += 1
            j = i % n
            if per_pair_threads == 2:
                box = {}
                th = threading.Thread(target=lambda: box.__setitem__("r", eR.extract(Rs[j])))
                th.start()
                kL, dL = eL.extract(Ls[j])
                th.join()
                kR, dR = box["r"]
            else:
                kL, dL = eL.extract(Ls[j]); kR, dR = eR.extract(Rs[j])
            O.stereo_match(eL, eR, kL, dL, kR, dR, BF, maxD)
            with lock:
                state["done"] += 1
                t_end[0] = time.time()

    t0 = time.time()
    ths = [threading.Thread(target=worker) for _ in range(threads)]
    for t in ths: t.start()
    for t in ths: t.join()
    el = t_end[0] - t0
    return state["done"] / el, state["done"], el


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path (oracle port: the reference itself
    cannot be built here — no OpenCV C++ headers) on all host cores."""
    if rank != 0:
        return
    from pointslot_b200.synth import synth_batch_parallel

    cores = host_cores()
    sample = max(8 * cores, 64)  # ~0.3 s of work per step on all cores
    Ls, Rs = synth_batch_parallel(W, H, range(sample))
    import oracle as O
    O.build()
    maxD = max_disparity()

    def step():
        lock = threading.Lock(); nxt = [0]

        def worker():
            eL, eR = O.OracleExtractor(*PARAMS), O.OracleExtractor(*PARAMS)
            while True:
                with lock:
                    i = nxt[0]; nxt[0] += 1
                if i >= sample:
                    return
                kL, dL = eL.extract(Ls[i]); kR, dR = eR.extract(Rs[i])
                O.stereo_match(eL, eR, kL, dL, kR, dR, BF, maxD)

        ths = [threading.Thread(target=worker) for _ in range(cores)]
        for t in ths: t.start()
        for t in ths: t.join()

    for _ in range(args.warmup):
        step()
    t0 = time.time()
    for _ in range(args.steps):
        step()
    el = time.time() - t0
    v = sample * args.steps / el
    desc = f"{sample} pairs/step of the same synthetic sequence, {cores} threads, 1 pair per thread"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_step": sample, "host_threads": cores},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch

    from pointslot_b200 import binding as B
    from pointslot_b200.shard import max_over_ranks, shard_range, sub_batches
    from pointslot_b200.synth import synth_batch_parallel

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"  # rank 0's stdout must carry exactly one JSON line (no NCCL version banner)
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    P, Bsz = args.pairs, args.batch
    nsub = (P + Bsz - 1) // Bsz
    maxD = max_disparity()
    t_gen = time.time()
    procs = max(1, host_cores() // max(world, 1))
    # weak scaling: the global batch is world * P pairs, rank r owns the contiguous block r (no collective)
    Ls, Rs = synth_batch_parallel(W, H, shard_range(world * P, rank, world), procs=min(procs, 64))
    t_gen = time.time() - t_gen

    # pinned host copies (e2e) and device-resident copies (value)
    hL = torch.empty((P, H, W), dtype=torch.uint8, pin_memory=True); hL.numpy()[:] = Ls
    hR = torch.empty((P, H, W), dtype=torch.uint8, pin_memory=True); hR.numpy()[:] = Rs
    dL = hL.cuda(non_blocking=False); dR = hR.cuda(non_blocking=False)
    img_bytes = W * H

    main = torch.cuda.current_stream()
    nctx = max(1, args.streams)
    streams = [torch.cuda.Stream() for _ in range(nctx)]
    ctxs = [B.StereoBatch(PARAMS, W, H, Bsz, device=local_rank, stream=s.cuda_stream) for s in streams]
    outs = [c.alloc_out(Bsz, pinned=True) for c in ctxs]
    copy_stream = torch.cuda.Stream()  # one upload stream for all contexts: H2D transfers at the full PCIe rate
    if not args.no_copy_stream:
        for c in ctxs:
            c.set_copy_stream(copy_stream.cuda_stream)
    K = ctxs[0].K
    for c in ctxs:
        c.set_profiling(True)

    stage_acc = {False: [np.zeros(len(B.StereoBatch.STAGES), np.float64), 0],
                 True: [np.zeros(len(B.StereoBatch.STAGES), np.float64), 0]}  # keyed by host-input mode

    subs = sub_batches(P, Bsz)

    def sub(i):
        return subs[i]

    def one_step(host: bool, collect: bool):
        pending = [False] * nctx
        for i in range(nsub):
            c = i % nctx
            ctx = ctxs[c]
            if pending[c]:
                ctx.wait()
                if collect:
                    stage_acc[host][0] += ctx.stage_times_ms(); stage_acc[host][1] += 1
            lo, n = sub(i)
            if host:
                ctx.submit_ptr(hL.data_ptr() + lo * img_bytes, hR.data_ptr() + lo * img_bytes, W, img_bytes, n, True,
                               BF, maxD, outs[c])
            else:
                ctx.submit_ptr(dL.data_ptr() + lo * img_bytes, dR.data_ptr() + lo * img_bytes, W, img_bytes, n, False,
                               BF, maxD, None)
            pending[c] = True
        for c in range(nctx):
            if pending[c]:
                ctxs[c].wait()
                if collect:
                    stage_acc[host][0] += ctxs[c].stage_times_ms(); stage_acc[host][1] += 1

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(host: bool, collect: bool):
        """W warm-up steps, then exactly K steps between CUDA events (fork/join over the batch streams)."""
        for _ in range(args.warmup):
            one_step(host, False)
        barrier()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        start.record(main)
        for s in streams + [copy_stream]:
            s.wait_event(start)
        for _ in range(args.steps):
            one_step(host, collect)
        for s in streams:
            e = torch.cuda.Event(); e.record(s); main.wait_event(e)
        stop.record(main)
        stop.synchronize()
        t1 = time.time()
        barrier()
        ms = start.elapsed_time(stop)
        ms = max_over_ranks(ms, dist, "cuda")  # a multi-GPU number is the slowest rank's device time
        return ms, (t0, t1)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ms_dev, win1 = timed(host=False, collect=False)
    # Per-kernel durations for the roofline: the same submits on ONE stream, so that the CUDA events around each
    # kernel (recorded by the library on that stream) are not stretched by kernels of the other streams.
    prof_steps = max(1, min(args.steps, 3))
    for _ in range(prof_steps):
        for i in range(nsub):
            lo, n = sub(i)
            ctxs[0].submit_ptr(dL.data_ptr() + lo * img_bytes, dR.data_ptr() + lo * img_bytes, W, img_bytes, n, False,
                               BF, maxD, None)
            ctxs[0].wait()
            stage_acc[False][0] += ctxs[0].stage_times_ms(); stage_acc[False][1] += 1
    stats = ctxs[0].stats()  # work counters of one sub-batch (off the timed path)
    stats_pairs = sub(nsub - 1)[1]
    launches_per_submit = ctxs[0].kernel_launches  # of the device-resident submits just timed
    ms_e2e, win2 = timed(host=True, collect=not args.no_e2e_stages)
    clocks = sampler.stop([win1, win2]) if rank == 0 else None

    # PCIe reference for the e2e number: the same pinned images uploaded alone on the copy stream (nothing else running)
    with torch.cuda.stream(copy_stream):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dL.copy_(hL, non_blocking=True); dR.copy_(hR, non_blocking=True)  # warm
        e0.record()
        for _ in range(3):
            dL.copy_(hL, non_blocking=True); dR.copy_(hR, non_blocking=True)
        e1.record()
    e1.synchronize()
    h2d_alone_gbs = 3 * 2 * P * img_bytes / (e0.elapsed_time(e1) * 1e-3) / 1e9

    # sanity: the e2e outputs of the last sub-batch on ctx 0 are real features
    n0 = int(outs[0]["counts"][0])
    assert 0 < n0 <= K, "bench produced no keypoints"

    if rank != 0:
        return
    total_pairs = P * world * args.steps
    value = total_pairs / (ms_dev * 1e-3)
    e2e = total_pairs / (ms_e2e * 1e-3)
    h2d = 2 * P * img_bytes
    d2h = sum(int(np.prod(outs[0][k].shape)) * outs[0][k].dtype.itemsize for k in ("counts", "kps", "desc", "u_right", "depth"))
    d2h = d2h * nsub  # every sub-batch returns its full-capacity arrays

    # ---- roofline of the dominant kernel ----------------------------------------------------------
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    stage_ms = stage_acc[False][0] / max(stage_acc[False][1], 1)  # average per submit (one sub-batch of Bsz pairs)
    e2e_stage_ms = stage_acc[True][0] / max(stage_acc[True][1], 1)
    # algorithmic bytes per PAIR of each stage (SURVEY.md §8d / DESIGN.md), from the measured counters
    from pointslot_b200.binding import lib as _lib  # noqa: F401
    sp = max(stats_pairs, 1)
    n_c = stats["candidates"] / (2 * sp)   # per image
    n_k = stats["keypoints"] / (2 * sp)    # per image
    N_m = stats["sad_evals"] / sp          # per pair
    lv = level_sizes()
    SP, P0, PL = sum(lv), lv[0], lv[-1]
    c_row, c_bar = 58.0, 10.0  # row-band candidates / Hamming evaluations per left keypoint (measured there)
    bytes_pair = {
        "ingest_copy": 2 * 2 * P0,
        "pyramid": 2 * ((SP - PL) + (SP - P0)),
        "fast": 2 * (SP + 8 * n_c),
        "quadtree": 2 * (8 * n_c + 8 * n_k),
        "describe": 2 * n_k * (1849 + 4 + 28 + 32),
        "stereo_match": 60 * 2 * n_k + n_k * (8 * c_row + 32 * c_bar) + 352 * N_m + 8 * n_k + 8 * N_m,
        "stereo_reject": 12 * n_k,
        "result_copy": 0,
    }
    stages = {}
    for name, ms in zip(B.StereoBatch.STAGES, stage_ms):
        gbs = bytes_pair[name] * Bsz / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        stages[name] = {"ms_per_launch": float(ms), "algorithmic_MB_per_launch": bytes_pair[name] * Bsz / 1e6,
                        "GBps": gbs, "frac_hbm": gbs / peak}
    kernels_only = {k: v for k, v in stages.items() if k not in ("ingest_copy", "result_copy")}
    dom = max(kernels_only, key=lambda k: kernels_only[k]["ms_per_launch"])
    traffic = None
    tpath = ROOT / "profiles" / "dram_traffic.json"
    if tpath.exists():
        tj = json.loads(tpath.read_text())
        if dom in tj.get("kernels", {}):  # ncu capture of a 125-pair launch; traffic is linear in the batch size
            traffic = tj["kernels"][dom]["dram_bytes_per_launch"] * Bsz / float(tj.get("batch", Bsz))
    roofline = {"bound": "hbm", "kernel": dom, "achieved": stages[dom]["GBps"], "peak": peak, "unit": "GB/s",
                "frac": stages[dom]["GBps"] / peak, "traffic": traffic, "peak_source": peak_src,
                "launch_pairs": Bsz, "ms_per_launch": stages[dom]["ms_per_launch"],
                "note": "integer-ALU / issue-bound kernel (ncu: issue slots ~66 % busy, DRAM traffic == algorithmic bytes); durations from a single-stream pass of the same submits inside this run; HBM fraction reported as the contract asks"}

    # whole path: compulsory bytes of all stages per pair x pairs/s, against the same HBM peak (per GPU)
    path_bytes = float(sum(bytes_pair.values()))
    path_gbs = path_bytes * (value / world) / 1e9
    roofline["path"] = {"algorithmic_MB_per_pair": path_bytes / 1e6, "achieved_GBps": path_gbs, "frac": path_gbs / peak,
                        "note": "all kernels of the path at the multi-stream `value` rate, per GPU"}

    # ---- CPU baseline (bounded sample, N=1 only) ---------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        ns = min(P, max(64, 4 * cores))
        v_all, done_all, el_all = cpu_pairs_per_sec(Ls[:ns], Rs[:ns], cores, args.cpu_seconds)
        v_ship, done_s, el_s = cpu_pairs_per_sec(Ls[:ns], Rs[:ns], 1, min(4.0, args.cpu_seconds), per_pair_threads=2)
        cpu = {"value": v_all, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{done_all} pairs of the same sequence in {el_all:.1f} s, {cores} threads x 1 pair each",
               "as_shipped_2_threads": {"value": v_ship, "sample": f"{done_s} pairs in {el_s:.1f} s, 2 threads per pair "
                                        "(Frame.cc:161-164), 1 pair in flight"}}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_step_per_gpu": P, "sub_batch_pairs": Bsz, "streams": nctx,
                   "image": f"{W}x{H}", "bf": BF, "maxD": maxD,
                   "l2": f"inputs larger than L2: {2 * P * img_bytes / 1e6:.0f} MB of images per step vs 126 MB L2",
                   "keypoints_per_image": n_k, "fast_candidates_per_image": n_c, "sad_refinements_per_pair": N_m,
                   "matches_per_pair": stats["matches"] / sp, "synth_seconds": t_gen},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                "ms_per_step": ms_e2e / args.steps,
                "pcie": {"h2d_GBps_in_run": h2d / (ms_e2e / args.steps * 1e-3) / 1e9, "h2d_GBps_alone": h2d_alone_gbs,
                         "note": "per GPU; e2e is bounded by the upload: ms_per_step >= h2d_bytes / h2d_GBps_alone"},
                "stage_ms_per_submit": {n: float(v) for n, v in zip(B.StereoBatch.STAGES, e2e_stage_ms)}},
        "gpu_launches": launches_per_submit * nsub * args.steps * world,
        "roofline": roofline, "stages": stages, "cpu_baseline": cpu, "clocks": clocks,
    }
    print(json.dumps(line), flush=True)


def level_sizes():
    """Pixel counts of the 8 pyramid levels of a 1242x375 image."""
    sf, out = np.float32(1.0), []
    for l in range(PARAMS[2]):
        if l:
            sf = np.float32(float(sf) * float(np.float32(PARAMS[1])))
        inv = np.float32(1.0) / sf
        out.append(int(np.rint(np.float64(np.float32(W) * inv))) * int(np.rint(np.float64(np.float32(H) * inv))))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=1000, help="pairs per step per GPU")
    ap.add_argument("--batch", type=int, default=50, help="pairs per sub-batch (one submit)")
    ap.add_argument("--streams", type=int, default=4, help="batch contexts / CUDA streams used round-robin")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-copy-stream", action="store_true", help="per-context uploads instead of one shared upload stream")
    ap.add_argument("--no-e2e-stages", action="store_true", help="do not read the per-stage events in the e2e pass")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
