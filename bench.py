#!/usr/bin/env python3
"""bench.py — stereo pairs/sec (ORB both views + stereo match) of PointSLOT's per-frame front end on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config NAME]      our CUDA path (default N=1, kitti1000)
  python bench.py --impl reference [...]                                   the reference's own CPU code (oracle/_ref)
  python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...   one rank per GPU

Workloads (BASELINE.json `configs`):
  kitti1000  configs[1]: 1000-frame synthetic KITTI-tracking-shaped sequence, ORBextractor(1000,1.2,8,20,5)
             [Examples/Stereo/0000-0013.yaml]  - THE METRIC'S CONFIGURATION, the default main line
  kitti2000  configs[0]: 1242x375, 2000 features, FAST 20/7 (the reference's CPU-runnable case): batch throughput and the
             single-pair latency through the reference-shaped API
  euroc      configs[2]: 752x480, 1200 features, batched
  hd1080     configs[3]: 1920x1080, 4000 features
  shard8000  configs[4]: 8000 kitti2000 pairs partitioned over the N GPUs (strong scaling, no collective)
The main line measures --config; the other configurations are measured briefly in the same run and reported under
`extra_configs` (N = 1; shard8000 at every N).

One "step" of the main line = the whole sequence `passes` times through the hot path (both extractions +
ComputeStereoMatches per pair), in sub-batches round-robin over a few streams; with N GPUs every rank processes its
own shard (weak scaling, no collective on the data path).

Printed JSON (rank 0): `value` = pairs/s with inputs resident in HBM; `e2e` = pairs/s through the C ABI with pinned
HOST input and output buffers (H2D + D2H inside the timed region); `roofline` = the dominant kernel's algorithmic bytes
/ its CUDA-event duration vs the measured HBM peak, with every kernel's bound and ncu pipe percentages beside it;
`cpu_baseline` = the reference's own CPU code (oracle/_ref, compiled from /root/reference) on this box's host cores.
"""
from __future__ import annotations

import argparse
import hashlib
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

UNIT = "pairs/s"
CONFIGS = {
    # bf / fx: Examples/Stereo/0000-0013.yaml:26, :8 for KITTI; EuRoC-like and a 1080p rig otherwise (they only set maxD)
    "kitti1000": dict(W=1242, H=375, params=(1000, 1.2, 8, 20, 5), bf=384.38148, fx=721.5377, pairs=1000, batch=50, passes=8,
                      workload="1000-frame synthetic KITTI-tracking-shaped stereo sequence, ORBextractor(1000,1.2,8,20,5) [0000-0013.yaml]"),
    "kitti2000": dict(W=1242, H=375, params=(2000, 1.2, 8, 20, 7), bf=384.38148, fx=721.5377, pairs=1000, batch=50, passes=1,
                      workload="synthetic KITTI-shaped 1242x375 stereo pairs, 2000 ORB features, 8 levels, scale 1.2, FAST 20/7"),
    "euroc": dict(W=752, H=480, params=(1200, 1.2, 8, 20, 7), bf=47.9, fx=435.2, pairs=512, batch=64, passes=1,
                  workload="EuRoC-shaped 752x480 stereo, 1200 features, 8 levels, batched frames"),
    "hd1080": dict(W=1920, H=1080, params=(4000, 1.2, 8, 20, 7), bf=756.0, fx=1400.0, pairs=64, batch=16, passes=1,
                   workload="1920x1080 stereo, 4000 features, 8 levels"),
    "shard8000": dict(W=1242, H=375, params=(2000, 1.2, 8, 20, 7), bf=384.38148, fx=721.5377, pairs=1000, batch=50, passes=8,
                      total_pairs=8000,
                      workload="8000-pair synthetic KITTI-shaped batch (2000 features, FAST 20/7) sharded across the GPUs"),
}
METRIC = "stereo pairs/sec (ORB both views + stereo match) KITTI 1242x375"


def metric_name(cfg):
    return f"stereo pairs/sec (ORB both views + stereo match) {cfg['W']}x{cfg['H']}" if (cfg["W"], cfg["H"]) != (1242, 375) else METRIC


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def max_disparity(cfg) -> float:
    """mb = mbf/fx, maxD = mbf/mb in float32 (Frame.cc:197, 2172-2174; DESIGN.md parity rule 2)."""
    mb = np.float32(cfg["bf"]) / np.float32(cfg["fx"])
    return float(np.float32(cfg["bf"]) / mb)


def level_sizes(cfg):
    """Pixel counts of the pyramid levels (ORBextractor.cc:421-429, 1111-1112)."""
    nfeat, scale, nlevels = cfg["params"][:3]
    sf, out = np.float32(1.0), []
    for l in range(nlevels):
        if l:
            sf = np.float32(float(sf) * float(np.float32(scale)))
        inv = np.float32(1.0) / sf
        out.append(int(np.rint(np.float64(np.float32(cfg["W"]) * inv))) * int(np.rint(np.float64(np.float32(cfg["H"]) * inv))))
    return out


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
# CPU arm: the reference's own code (oracle/_ref/libref_orb.so: src/ORBextractor.cc + Frame::ComputeStereoMatches compiled
# from /root/reference, OpenCV primitives served by the pinned C restatements) or, if that library was not built, the
# oracle port.  Test infrastructure timed as the CPU baseline; never on the product path.
# ------------------------------------------------------------------------------------------------
class CpuArm:
    def __init__(self, cfg, prefer_reference=True):
        self.cfg = cfg
        self.kind = "port"
        import oracle as O
        O.build()
        self.O = O
        self.maxD = max_disparity(cfg)
        self.mb = float(np.float32(cfg["bf"]) / np.float32(cfg["fx"]))
        if prefer_reference:
            try:
                from oracle import ref_path
                if ref_path.available("stock"):
                    ref_path._load("stock")
                    ref_path._set_primitives("stock", "native")  # per library, set once before the worker threads start
                    self.ref_path = ref_path
                    self.kind = "reference"
            except Exception:
                self.kind = "port"

    def describe(self):
        if self.kind == "reference":
            return ("the reference's own src/ORBextractor.cc + Frame::ComputeStereoMatches compiled from /root/reference "
                    "(oracle/_ref/libref_orb.so, -O3, stock allocator; OpenCV primitives = the pinned C restatements, OpenCV "
                    "C++ is not linkable in this image)")
        return "C port of the reference path (oracle/orb_oracle.c); oracle/_ref was not built"

    def make_pair_fn(self):
        """-> a function(L, R) that processes one stereo pair on the calling thread (its own extractor objects)."""
        cfg = self.cfg
        if self.kind == "reference":
            rp = self.ref_path
            eL = rp.ReferenceExtractor(*cfg["params"], variant="stock", primitives="native")
            eR = rp.ReferenceExtractor(*cfg["params"], variant="stock", primitives="native")
            lib = eL._lib
            cap = eL.cap + 4096
            kL = np.zeros(cap, rp.KEYPOINT_DTYPE); dL = np.zeros((cap, 32), np.uint8)
            kR = np.zeros(cap, rp.KEYPOINT_DTYPE); dR = np.zeros((cap, 32), np.uint8)
            ur = np.zeros(cap, np.float32); dp = np.zeros(cap, np.float32)

            def run(L, R, two_threads=False):
                box = {}

                def right():
                    box["n"] = lib.ref_extract(eR._h, R.ctypes.data, R.shape[1], R.shape[0], R.strides[0], kR.ctypes.data, dR.ctypes.data, cap)
                if two_threads:
                    th = threading.Thread(target=right); th.start()
                nl = lib.ref_extract(eL._h, L.ctypes.data, L.shape[1], L.shape[0], L.strides[0], kL.ctypes.data, dL.ctypes.data, cap)
                if two_threads:
                    th.join()
                else:
                    right()
                nr = box["n"]
                assert nl >= 0 and nr >= 0
                lib.reff_stereo(eL._h, eR._h, kL.ctypes.data, nl, dL.ctypes.data, kR.ctypes.data, nr, dR.ctypes.data,
                                cfg["bf"], self.mb, ur.ctypes.data, dp.ctypes.data)
                return nl
            run.keep = (eL, eR)
            return run
        O = self.O
        eL, eR = O.OracleExtractor(*cfg["params"]), O.OracleExtractor(*cfg["params"])

        def run(L, R, two_threads=False):
            if two_threads:
                box = {}
                th = threading.Thread(target=lambda: box.__setitem__("r", eR.extract(R)))
                th.start()
                kL, dL = eL.extract(L)
                th.join()
                kR, dR = box["r"]
            else:
                kL, dL = eL.extract(L); kR, dR = eR.extract(R)
            O.stereo_match(eL, eR, kL, dL, kR, dR, cfg["bf"], self.maxD)
            return len(kL)
        return run

    def pairs_per_sec(self, Ls, Rs, threads, seconds, per_pair_threads=1):
        """Pairs/s with `threads` worker threads for about `seconds` (ctypes releases the GIL).  per_pair_threads=2
        reproduces the reference as shipped: left/right extraction on two threads (Frame.cc:161-164), stereo serial."""
        n = len(Ls)
        lock = threading.Lock()
        state = {"next": 0, "done": 0}
        t_end = [0.0]
        deadline = time.time() + seconds

        def worker():
            fn = self.make_pair_fn()
            while True:
                with lock:
                    if time.time() >= deadline:
                        return
                    i = state["next"]; state["next"] += 1
                fn(Ls[i % n], Rs[i % n], per_pair_threads == 2)
                with lock:
                    state["done"] += 1
                    t_end[0] = time.time()

        t0 = time.time()
        ths = [threading.Thread(target=worker) for _ in range(threads)]
        for t in ths: t.start()
        for t in ths: t.join()
        el = max(t_end[0] - t0, 1e-9)
        return state["done"] / el, state["done"], el


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path on all host cores, bounded sample per step."""
    if rank != 0:
        return
    from pointslot_b200.synth import synth_batch_parallel

    cfg = CONFIGS[args.config]
    cores = host_cores()
    sample = max(8 * cores, 64)  # ~0.3-0.6 s of work per step on all cores
    Ls, Rs = synth_batch_parallel(cfg["W"], cfg["H"], range(sample))
    arm = CpuArm(cfg)
    fns = [arm.make_pair_fn() for _ in range(cores)]

    def step():
        lock = threading.Lock(); nxt = [0]

        def worker(fn):
            while True:
                with lock:
                    i = nxt[0]; nxt[0] += 1
                if i >= sample:
                    return
                fn(Ls[i], Rs[i])

        ths = [threading.Thread(target=worker, args=(fn,)) for fn in fns]
        for t in ths: t.start()
        for t in ths: t.join()

    for _ in range(args.warmup):
        step()
    t0 = time.time()
    for _ in range(args.steps):
        step()
    el = time.time() - t0
    v = sample * args.steps / el
    desc = f"{sample} pairs/step of the same synthetic sequence, {cores} threads, 1 pair per thread; {arm.describe()}"
    print(json.dumps({
        "impl": "reference", "metric": metric_name(cfg), "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        # the GPU arm's workload (the static keys of its `config`, same names and values); what this arm ran of it is `sample`
        "config": {"workload": cfg["workload"], "name": args.config, "image": f"{cfg['W']}x{cfg['H']}", "params": list(cfg["params"]),
                   "bf": cfg["bf"], "maxD": max_disparity(cfg)},
        "sample": {"pairs_per_step": sample, "host_threads": cores,
                   "note": "bounded sample of the GPU arm's workload: same images, same extractor parameters, fewer pairs per step"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": arm.kind, "sample": desc},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
STAGE_KERNEL = {"pyramid": "k_resize_level", "fast": "k_fast_cells", "quadtree": "k_quadtree", "describe": "k_describe",
                "stereo_match": "k_stereo_match", "stereo_reject": "k_stereo_reject", "ingest_copy": "k_gray_repitch"}


def gpu_cpu_affinity(local_rank: int):
    """CPU set nvidia-smi reports next to this GPU (NUMA locality of its PCIe root), or None."""
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        for line in out.splitlines():
            f = line.replace("\x1b[4m", "").replace("\x1b[0m", "").split("\t")
            if f and f[0].strip() == f"GPU{local_rank}":
                for tok in f[1:]:
                    tok = tok.strip()
                    if tok and tok[0].isdigit() and all(c in "0123456789-," for c in tok):
                        cpus = set()
                        for part in tok.split(","):
                            a, _, b = part.partition("-")
                            cpus.update(range(int(a), int(b or a) + 1))
                        return cpus
    except Exception:
        pass
    return None


class Workload:
    """One configuration on one rank: pinned host images, device-resident copies, batch contexts on streams."""

    def __init__(self, cfg, rank, world, local_rank, args, dist, pairs=None, seed0=0, images=None):
        import torch
        from pointslot_b200 import binding as B
        from pointslot_b200.shard import shard_range, sub_batches
        from pointslot_b200.synth import synth_batch_parallel

        self.torch, self.B, self.cfg, self.args, self.dist = torch, B, cfg, args, dist
        self.rank, self.world, self.local_rank = rank, world, local_rank
        W, H = cfg["W"], cfg["H"]
        self.P = P = pairs if pairs is not None else cfg["pairs"]  # distinct pairs held by this rank
        self.Bsz = Bsz = min(args.batch or cfg["batch"], P)
        self.passes = cfg.get("passes", 1)
        self.maxD = max_disparity(cfg)
        t0 = time.time()
        if images is not None:
            Ls, Rs = images
        else:
            procs = max(1, host_cores() // max(world, 1))
            # the global batch is world * P pairs; rank r owns the contiguous block r (no collective)
            seeds = [seed0 + s for s in shard_range(world * P, rank, world)]
            Ls, Rs = synth_batch_parallel(W, H, seeds, procs=min(procs, 64))
        self.t_gen = time.time() - t0
        self.Ls, self.Rs = Ls, Rs
        self.img_bytes = W * H
        # pinned host copies (e2e) and device-resident copies (value)
        if args.pin != "default":
            self._pin = [B.PinnedBuffer((P, H, W), np.uint8, mode=args.pin) for _ in range(2)]
            self._pin[0].array[:] = Ls; self._pin[1].array[:] = Rs
            self.hL_ptr, self.hR_ptr = self._pin[0].ptr, self._pin[1].ptr
            # torch sees the pages as pinned (cudaPointerGetAttributes), so copy_(non_blocking=True) is a direct async DMA
            self.hL, self.hR = torch.from_numpy(self._pin[0].array), torch.from_numpy(self._pin[1].array)
        else:
            self.hL = torch.empty((P, H, W), dtype=torch.uint8, pin_memory=True); self.hL.numpy()[:] = Ls
            self.hR = torch.empty((P, H, W), dtype=torch.uint8, pin_memory=True); self.hR.numpy()[:] = Rs
            self.hL_ptr, self.hR_ptr = self.hL.data_ptr(), self.hR.data_ptr()
        self.dL = torch.from_numpy(Ls).cuda(); self.dR = torch.from_numpy(Rs).cuda()
        self.main = torch.cuda.current_stream()
        self.nctx = max(1, args.streams)
        self.streams = [torch.cuda.Stream() for _ in range(self.nctx)]
        self.ctxs = [B.StereoBatch(cfg["params"], W, H, Bsz, device=local_rank, stream=s.cuda_stream) for s in self.streams]
        self.outs = [c.alloc_out(Bsz, pinned=True) for c in self.ctxs]
        self.copy_stream = torch.cuda.Stream()  # one upload stream for all contexts: H2D transfers at the full PCIe rate
        for c in self.ctxs:
            c.set_copy_stream(self.copy_stream.cuda_stream)
            c.set_profiling(True)
        self.K = self.ctxs[0].K
        self.subs = sub_batches(P, Bsz) * self.passes  # one step = `passes` times through the sequence
        self.nsub = len(self.subs)
        self.e2e_nsub = self.nsub  # sub-batches this rank uploads per end-to-end step (rate-proportional shards change it)
        nst = len(B.StereoBatch.STAGES)
        self.stage_acc = {False: [np.zeros(nst, np.float64), 0], True: [np.zeros(nst, np.float64), 0]}

    def pairs_per_step(self):
        return self.P * self.passes

    def _submit(self, ctx, i, host, out):
        lo, n = self.subs[i % self.nsub]
        ib = self.img_bytes
        if host:
            ctx.submit_ptr(self.hL_ptr + lo * ib, self.hR_ptr + lo * ib, self.cfg["W"], ib, n, True, self.cfg["bf"], self.maxD, out)
        else:
            ctx.submit_ptr(self.dL.data_ptr() + lo * ib, self.dR.data_ptr() + lo * ib, self.cfg["W"], ib, n, False,
                           self.cfg["bf"], self.maxD, None)

    def one_step(self, host, collect):
        pending = [False] * self.nctx
        for i in range(self.e2e_nsub if host else self.nsub):
            c = i % self.nctx
            ctx = self.ctxs[c]
            if pending[c]:
                ctx.wait()
                if collect:
                    self.stage_acc[host][0] += ctx.stage_times_ms(); self.stage_acc[host][1] += 1
            self._submit(ctx, i, host, self.outs[c] if host else None)
            pending[c] = True
        for c in range(self.nctx):
            if pending[c]:
                self.ctxs[c].wait()
                if collect:
                    self.stage_acc[host][0] += self.ctxs[c].stage_times_ms(); self.stage_acc[host][1] += 1

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def timed(self, host, steps, warmup, collect=False, local=False):
        """`warmup` untimed steps, then exactly `steps` steps between CUDA events (fork/join over the batch streams);
        -> (ms of the slowest rank - or of this rank when local=True -, wall-clock window)."""
        from pointslot_b200.shard import max_over_ranks
        torch = self.torch
        for _ in range(warmup):
            self.one_step(host, False)
        self.barrier()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        start.record(self.main)
        for s in self.streams + [self.copy_stream]:
            s.wait_event(start)
        for _ in range(steps):
            self.one_step(host, collect)
        for s in self.streams:
            e = torch.cuda.Event(); e.record(s); self.main.wait_event(e)
        stop.record(self.main)
        stop.synchronize()
        t1 = time.time()
        self.barrier()
        ms = start.elapsed_time(stop)
        if not local:
            ms = max_over_ranks(ms, self.dist, "cuda")  # a multi-GPU number is the slowest rank's device time
        return ms, (t0, t1)

    def gather(self, x: float):
        """-> the value of every rank, in rank order."""
        if self.dist is None:
            return [x]
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        g = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(g, t)
        return [float(v.item()) for v in g]

    def profile_pass(self, reps=2):
        """Per-kernel durations for the roofline: the same submits on ONE stream, so that the CUDA events around each
        kernel (recorded by the library on that stream) are not stretched by kernels of the other streams; plus the
        work counters of the run (candidates, keypoints, SAD evaluations, matches, stereo search counters)."""
        ctx = self.ctxs[0]
        ctx.stereo_counters()  # reset
        nsub1 = self.nsub // self.passes
        tot = dict(candidates=0, keypoints=0, sad_evals=0, matches=0, pairs=0)
        for r in range(reps):
            for i in range(nsub1):
                self._submit(ctx, i, False, None)
                ctx.wait()
                self.stage_acc[False][0] += ctx.stage_times_ms(); self.stage_acc[False][1] += 1
                if r == 0:
                    st = ctx.stats()
                    for k in ("candidates", "keypoints", "sad_evals", "matches"):
                        tot[k] += st[k]
                    tot["pairs"] += self.subs[i][1]
        cnt = ctx.stereo_counters()
        tot["row_candidates"] = cnt["row_candidates"] / reps
        tot["hamming_evals"] = cnt["hamming_evals"] / reps
        self.launches_per_submit = ctx.kernel_launches
        return tot

    def h2d_only(self, reps=3):
        """The same pinned images uploaded with nothing else running on this GPU -> GB/s of this rank (call between
        barriers to get the all-ranks-concurrent figure)."""
        torch = self.torch
        if self.hL is None:
            return None
        with torch.cuda.stream(self.copy_stream):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.dL.copy_(self.hL, non_blocking=True); self.dR.copy_(self.hR, non_blocking=True)  # warm
            e0.record()
            for _ in range(reps):
                self.dL.copy_(self.hL, non_blocking=True); self.dR.copy_(self.hR, non_blocking=True)
            e1.record()
        e1.synchronize()
        return reps * 2 * self.P * self.img_bytes / (e0.elapsed_time(e1) * 1e-3) / 1e9

    def h2d_window(self, seconds=0.6):
        """Upload sub-batch-sized chunks of the pinned images back to back for a fixed wall-clock window, nothing else running
        on this GPU -> bytes moved.  Started on all ranks behind a barrier, the sum over ranks / window is what the host
        delivers while EVERY link is busy for the whole window (h2d_only lets a fast rank finish early, after which the
        slow ranks speed up: the sum of those per-rank rates overstates the concurrent aggregate)."""
        torch = self.torch
        if self.hL is None:
            return None
        ib, moved = self.img_bytes, 0
        hL, hR = self.hL.view(-1), self.hR.view(-1)
        dL, dR = self.dL.view(-1), self.dR.view(-1)
        with torch.cuda.stream(self.copy_stream):
            t_end = time.perf_counter() + seconds
            i, inflight = 0, []
            while time.perf_counter() < t_end:
                lo, n = self.subs[i % self.nsub]
                a, b = lo * ib, (lo + n) * ib
                dL[a:b].copy_(hL[a:b], non_blocking=True); dR[a:b].copy_(hR[a:b], non_blocking=True)
                ev = torch.cuda.Event(); ev.record()
                inflight.append(ev)
                if len(inflight) > 2:  # two chunks queued behind the running one keep the copy engine busy
                    inflight.pop(0).synchronize()
                moved += 2 * n * ib
                i += 1
            self.copy_stream.synchronize()
        return moved

    def d2h_bytes_per_step(self):
        o = self.outs[0]
        per = sum(int(np.prod(o[k].shape)) * o[k].dtype.itemsize for k in ("counts", "kps", "desc", "u_right", "depth"))
        return per * self.nsub  # every sub-batch returns its full-capacity arrays

    def close(self):
        for c in self.ctxs:
            c.close()
        self.ctxs = []
        self.dL = self.dR = self.hL = self.hR = None
        self.outs = []
        self._pin = []
        self.torch.cuda.empty_cache()


def proportional_shards(total: int, rates):
    """Split `total` sub-batches over the ranks in proportion to `rates` (largest-remainder rounding, every rank >= 1)."""
    rates = [max(float(r), 1e-9) for r in rates]
    s = sum(rates)
    raw = [total * r / s for r in rates]
    out = [max(1, int(x)) for x in raw]
    while sum(out) < total:
        i = max(range(len(out)), key=lambda k: raw[k] - out[k])
        out[i] += 1
    while sum(out) > total:
        i = max((k for k in range(len(out)) if out[k] > 1), key=lambda k: out[k] - raw[k])
        out[i] -= 1
    return out


def stage_bytes_per_pair(cfg, tot):
    """Algorithmic (compulsory) global bytes of each stage per PAIR (SURVEY.md §8d / DESIGN.md §4), evaluated with the
    counters measured in this run."""
    sp = max(tot["pairs"], 1)
    n_c = tot["candidates"] / (2 * sp)      # FAST candidates per image
    n_k = tot["keypoints"] / (2 * sp)       # keypoints per image
    N_m = tot["sad_evals"] / sp             # left keypoints reaching the SAD refinement, per pair
    c_row = tot["row_candidates"] / max(tot["keypoints"] / 2, 1)   # row-band candidates scanned per left keypoint
    c_bar = tot["hamming_evals"] / max(tot["keypoints"] / 2, 1)    # descriptors compared per left keypoint
    lv = level_sizes(cfg)
    SP, P0, PL = sum(lv), lv[0], lv[-1]
    b = {
        "ingest_copy": 2 * 2 * P0,
        "pyramid": 2 * ((SP - PL) + (SP - P0)),
        "fast": 2 * (SP + 8 * n_c),
        "quadtree": 2 * (8 * n_c + 8 * n_k),
        "describe": 2 * n_k * (1849 + 4 + 28 + 32),
        "stereo_match": 60 * 2 * n_k + n_k * (8 * c_row + 32 * c_bar) + 352 * N_m + 8 * n_k + 8 * N_m,
        "stereo_reject": 12 * n_k,
        "result_copy": 0,
    }
    counts = {"keypoints_per_image": n_k, "fast_candidates_per_image": n_c, "sad_refinements_per_pair": N_m,
              "matches_per_pair": tot["matches"] / sp, "row_candidates_per_left_keypoint": c_row,
              "hamming_evals_per_left_keypoint": c_bar}
    return b, counts


def determinism_check(cfg, rank, world, local_rank, dist, n=50):
    """the same sub-batch on every GPU must give byte-identical results.  Every rank runs the same `n` pairs,
    hashes every output array, and the digests are gathered; rank 0 also checks a sample of pairs against the oracle."""
    import torch
    from pointslot_b200 import binding as B
    from pointslot_b200.synth import synth_batch_parallel
    W, H = cfg["W"], cfg["H"]
    Ls, Rs = synth_batch_parallel(W, H, range(90000, 90000 + n), procs=max(1, min(16, host_cores() // max(world, 1))))
    sb = B.StereoBatch(cfg["params"], W, H, n, device=local_rank)
    out = sb.alloc_out(n)
    sb.submit(Ls, Rs, cfg["bf"], max_disparity(cfg), out)
    sb.wait()
    h = hashlib.sha256()
    for k in ("counts", "kps", "desc", "u_right", "depth"):
        h.update(np.ascontiguousarray(out[k]).tobytes())
    digest = h.digest()
    digests = [digest]
    if dist is not None:
        t = torch.tensor(list(digest), dtype=torch.uint8, device="cuda")
        g = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        digests = [bytes(x.cpu().tolist()) for x in g]
    res = None
    if rank == 0:
        import oracle as O
        O.build()
        checked, ok = 0, True
        for i in (0, n // 2, n - 1)[:3]:
            oL, oR = O.OracleExtractor(*cfg["params"]), O.OracleExtractor(*cfg["params"])
            okL, odL = oL.extract(Ls[i]); okR, odR = oR.extract(Rs[i])
            our, odp = O.stereo_match(oL, oR, okL, odL, okR, odR, cfg["bf"], max_disparity(cfg))
            nl, nr = int(out["counts"][i]), int(out["counts"][n + i])
            ok &= (nl == len(okL) and nr == len(okR) and out["kps"][i, :nl].tobytes() == okL.tobytes()
                   and out["kps"][n + i, :nr].tobytes() == okR.tobytes() and np.array_equal(out["desc"][i, :nl], odL)
                   and np.array_equal(out["desc"][n + i, :nr], odR) and out["u_right"][i, :nl].tobytes() == our.tobytes()
                   and out["depth"][i, :nl].tobytes() == odp.tobytes())
            checked += 1
        res = {"pairs": n, "ranks": world, "identical_across_ranks": all(d == digests[0] for d in digests),
               "sha256": digests[0].hex(), "oracle_checked_pairs": checked, "equal_to_oracle": bool(ok)}
    sb.close()
    return res


def measure_brief(cfg, name, rank, world, local_rank, args, dist, steps, warmup, pairs=None, images=None):
    """value + e2e of a secondary configuration with few steps -> dict (rank 0) / None."""
    wl = Workload(cfg, rank, world, local_rank, args, dist, pairs=pairs, images=images)
    ms_dev, _ = wl.timed(False, steps, warmup)
    tot = wl.profile_pass(reps=1)
    ms_e2e, _ = wl.timed(True, steps, warmup)
    res = None
    if rank == 0:
        pps = wl.pairs_per_step() * world
        bpp, counts = stage_bytes_per_pair(cfg, tot)
        stage_ms = wl.stage_acc[False][0] / max(wl.stage_acc[False][1], 1)
        peak = hbm_peak()[0]
        stages = {}
        for nme, ms in zip(wl.B.StereoBatch.STAGES, stage_ms):
            gbs = bpp[nme] * wl.Bsz / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
            stages[nme] = {"ms_per_launch": float(ms), "GBps": gbs, "frac_hbm": gbs / peak}
        res = {"workload": cfg["workload"], "image": f"{cfg['W']}x{cfg['H']}", "params": list(cfg["params"]),
               "n_gpus": world, "pairs_per_step": pps, "sub_batch_pairs": wl.Bsz, "steps": steps, "warmup": warmup,
               "value": pps * steps / (ms_dev * 1e-3), "e2e": pps * steps / (ms_e2e * 1e-3), "unit": UNIT,
               "ms_per_step": ms_dev / steps, "e2e_ms_per_step": ms_e2e / steps,
               "h2d_bytes_per_step": 2 * wl.pairs_per_step() * wl.img_bytes * world, "d2h_bytes_per_step": wl.d2h_bytes_per_step() * world,
               "algorithmic_MB_per_pair": float(sum(bpp.values())) / 1e6, "counts": counts, "stages": stages}
    wl.close()
    return res


def single_pair_latency(cfg, local_rank, reps=30):
    """configs[0] as the reference runs it: ONE pair through the reference-shaped API (two extractor contexts, both views in
    flight at once as Frame.cc:161-164 arranges with two threads, then ComputeStereoMatches), host images in, host results out
    -> ms per pair."""
    from pointslot_b200 import binding as B
    from pointslot_b200.synth import synth_pair
    L, R = synth_pair(cfg["W"], cfg["H"], 0)
    eL, eR = B.OrbExtractor(*cfg["params"], device=local_rank), B.OrbExtractor(*cfg["params"], device=local_rank)
    maxD = max_disparity(cfg)

    def once():
        # one host thread keeps both views in flight (pslot_orb_extract_begin / _end), as the drop-in Frame constructor does
        eL.extract_begin(L); eR.extract_begin(R)
        kL, dL = eL.extract_end(); kR, dR = eR.extract_end()
        # ... and ComputeStereoMatches runs on the feature sets where they lie (pslot_stereo_match_resident)
        B.stereo_match_resident(eL, eR, B.extract_token(eL), B.extract_token(eR), cfg["bf"], maxD)
    for _ in range(5):
        once()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); once(); ts.append(time.perf_counter() - t0)
    return float(np.median(ts) * 1e3)


def object_frame_latency(cfg, local_rank, reps=30):
    """The device work of the reference's object-mode Frame constructor (SLOT mode 4, Frame.cc:660-757) for ONE stereo pair:
    ExtractORB left / right on two host threads + ExtractObjORB (cv::ORB on both views inside the object masks, here one
    batched call) on a third (:706-713), then ComputeStereoMatches and ComputeObjStereoMatches (:720-726) -> ms per frame,
    host images in, host results out."""
    from pointslot_b200 import binding as B
    from pointslot_b200.synth import synth_pair
    W, H = cfg["W"], cfg["H"]
    L, R = synth_pair(W, H, 0)
    mask = np.zeros((H, W), np.uint8)
    mask[H // 4:3 * H // 4, W // 6:W // 2] = 255; mask[H // 8:H // 2, 2 * W // 3:W - 40] = 255
    eL, eR = B.OrbExtractor(*cfg["params"], device=local_rank), B.OrbExtractor(*cfg["params"], device=local_rank)
    orb = B.CvOrb(1000, 1.2, 8, 19, device=local_rank)
    maxD = max_disparity(cfg)

    def once():
        eL.extract_begin(L); eR.extract_begin(R)  # in flight while this thread runs the (synchronous) cv::ORB pair
        (okL, odL), (okR, odR) = orb.detect_and_compute_batch([L, R], [mask, mask])
        kL, dL = eL.extract_end(); kR, dR = eR.extract_end()
        B.stereo_match_resident(eL, eR, B.extract_token(eL), B.extract_token(eR), cfg["bf"], maxD)
        B.stereo_match(eL, eR, okL, odL, okR, odR, cfg["bf"], maxD)  # ComputeObjStereoMatches: same routine, object keypoints
        return len(kL), len(okL)
    for _ in range(5):
        n = once()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); once(); ts.append(time.perf_counter() - t0)
    orb.close()
    return {"ms_per_frame": float(np.median(ts) * 1e3), "keypoints_left": n[0], "object_keypoints_left": n[1]}


def measure_cvorb(Ls, Rs, W, H, local_rank, seconds=1.0):
    """SURVEY §8f row 4 (next row): OpenCV's own ORB for the dense object features, reference Frame.cc:2623-2665 - the pair
    (left + right view with object masks, one batched call: what Frame::ExtractObjORB does per frame) and a 64-image batch,
    through the C ABI from page-locked host images, wall clock; cv2.ORB_create on one host core beside it when importable."""
    from pointslot_b200 import binding as B
    n = min(64, len(Ls))
    pin = B.PinnedBuffer((n, H, W)); pm = B.PinnedBuffer((H, W))
    for i in range(n):
        pin.array[i] = Ls[i] if i % 2 == 0 else Rs[i]
    pm.array[:] = 0
    pm.array[H // 4:3 * H // 4, W // 6:W // 2] = 255; pm.array[H // 8:H // 2, 2 * W // 3:W - 40] = 255
    orb = B.CvOrb(1000, 1.2, 8, 19, device=local_rank)
    out = {"what": "cv::ORB::create(1000,1.2,8,19)->detectAndCompute(image, objectMask), host images in, host keypoints + descriptors out"}
    for label, nb in (("pair", 2), ("batch64", n)):
        batch = [pin.array[i] for i in range(nb)]
        masks = [pm.array] * nb
        for _ in range(3):
            orb.detect_and_compute_batch(batch, masks, copy=False)
        reps, t0, abi = 0, time.perf_counter(), 0.0
        while time.perf_counter() - t0 < seconds / 2:
            res = orb.detect_and_compute_batch(batch, masks, copy=False)  # views into the binding's result arrays
            abi += float(orb.last_timing().sum())
            reps += 1
        dt = time.perf_counter() - t0
        out[label] = {"images": nb, "ms_per_call": dt / reps * 1e3, "images_per_s": nb * reps / dt,
                      "ms_per_call_inside_the_abi": abi / reps, "images_per_s_inside_the_abi": nb * reps / (abi * 1e-3),
                      "keypoints_per_image": float(np.mean([len(k) for k, _ in res])),
                      "abi_phases_ms": [round(float(v), 4) for v in orb.last_timing()]}
    try:
        import cv2
        cv2.setNumThreads(1)
        c = cv2.ORB_create(1000, 1.2, 8, 19)
        m = np.array(pm.array)
        t0, reps = time.perf_counter(), 0
        while time.perf_counter() - t0 < seconds / 2:
            c.detectAndCompute(np.asarray(Ls[reps % n]), m)
            reps += 1
        out["cv2_one_core_images_per_s"] = reps / (time.perf_counter() - t0)
    except ImportError:
        pass
    orb.close(); pin.close(); pm.close()
    return out


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_ours(args, rank, world, local_rank):
    import torch

    from pointslot_b200 import binding as B

    torch.cuda.set_device(local_rank)
    affinity = None
    full_affinity = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    if args.affinity == "auto":
        affinity = gpu_cpu_affinity(local_rank)
        if affinity:
            try:
                os.sched_setaffinity(0, affinity & os.sched_getaffinity(0) or os.sched_getaffinity(0))
            except Exception:
                affinity = None
    dist = None
    if world > 1:
        # NCCL's own log level is left alone (the driver reads its INFO lines to confirm the rank count); rank 0's stdout
        # carries the JSON line, NCCL writes to stderr / NCCL_DEBUG_FILE
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cfg = CONFIGS[args.config]
    pairs = args.pairs or cfg["pairs"]
    if "total_pairs" in cfg:  # strong scaling: the global batch is fixed, every rank holds total/world/passes distinct pairs
        pairs = max(cfg["total_pairs"] // world // cfg["passes"], 1)
    wl = Workload(cfg, rank, world, local_rank, args, dist, pairs=pairs)
    P, Bsz = wl.P, wl.Bsz

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ms_dev, win1 = wl.timed(False, args.steps, args.warmup)
    tot = wl.profile_pass(reps=2)
    ms_e2e, win2 = wl.timed(True, args.steps, args.warmup, collect=not args.no_e2e_stages)
    e2e_equal = None
    shard_subs = [wl.nsub] * world
    if world > 1 and not args.equal_shards:
        # The PCIe paths of the GPUs of one box are not equally fast when every GPU uploads at once (tools/h2d_probe.py:
        # GPUs 0-3 of these hosts share about 115 GB/s, GPUs 4-7 get about 55 GB/s each).  End to end the upload is the
        # bottleneck, so the contiguous blocks of the batch are sized by what each GPU's path delivers instead of equally:
        # first from the concurrent upload-only rates, then corrected by the step times measured with that split.
        # Same total number of pairs per step; no collective on the data path.
        e2e_equal = {"value": None, "ms_per_step": ms_e2e / args.steps}
        wl.barrier()
        rates = wl.gather(wl.h2d_only() or 1.0)
        wl.barrier()
        total_subs = wl.nsub * world
        shard_subs = proportional_shards(total_subs, rates)
        wl.e2e_nsub = shard_subs[rank]
        # a rank that finishes early leaves its share of the host's bandwidth to the others, so one correction
        # under-shoots: repeat it a few times (each pass = 2 steps with the current split)
        for _ in range(3):
            t_local, _ = wl.timed(True, 2, 1, local=True)
            times = wl.gather(t_local)
            shard_subs = proportional_shards(total_subs, [n / max(t, 1e-6) for n, t in zip(shard_subs, times)])
            wl.e2e_nsub = shard_subs[rank]
        ms_e2e, win3 = wl.timed(True, args.steps, 2)
        win2 = (win2[0], win3[1])
        t_local, _ = wl.timed(True, 2, 1, local=True)
        e2e_equal["ms_per_step_by_rank"] = [t / 2 for t in wl.gather(t_local)]
    clocks = sampler.stop([win1, win2]) if rank == 0 else None

    # PCIe references for the e2e number: this rank's pinned images uploaded (1) with every rank uploading at the same time
    # (barrier on both sides: the ceiling the host can feed N GPUs at) and (2) by rank 0 alone
    wl.barrier()
    h2d_all = wl.h2d_only()
    wl.barrier()
    h2d_min = h2d_sum = h2d_all
    if dist is not None and h2d_all is not None:
        t = torch.tensor([h2d_all, -h2d_all, h2d_all], dtype=torch.float64, device="cuda")
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        h2d_min, h2d_sum = -float(tmax[1].item()), float(tsum[2].item())
    h2d_alone = wl.h2d_only() if rank == 0 else None
    wl.barrier()
    # the same ceiling over a COMMON window: every rank uploads for 0.6 s, bytes summed over ranks
    h2d_win = None
    if dist is not None and wl.hL is not None:
        wl.barrier()
        t0w = time.perf_counter()
        moved = wl.h2d_window(0.6)
        dt = time.perf_counter() - t0w
        wl.barrier()
        parts = wl.gather(moved / dt / 1e9)
        h2d_win = {"aggregate_GBps": sum(parts), "GBps_by_rank": parts, "window_s": 0.6}

    n0 = int(wl.outs[0]["counts"][0])
    assert 0 < n0 <= wl.K, "bench produced no keypoints"

    det = None if args.no_determinism else determinism_check(cfg, rank, world, local_rank, dist)

    pairs_step = wl.pairs_per_step()
    e2e_pairs_step_all = sum(shard_subs) * wl.Bsz          # pairs of one end-to-end step over all ranks
    h2d = 2 * pairs_step * wl.img_bytes
    d2h = wl.d2h_bytes_per_step()
    stage_acc, launches_per_submit, nsub, t_gen = wl.stage_acc, wl.launches_per_submit, wl.nsub, wl.t_gen
    Ls, Rs = wl.Ls, wl.Rs
    wl.close()

    # ---- the other BASELINE configurations, briefly --------------------------------------------------------------
    extras = {}
    if not args.no_extra:
        es, ew = args.extra_steps, 3
        for name in ("kitti2000", "euroc", "hd1080", "shard8000"):
            if name == args.config:
                continue
            c = CONFIGS[name]
            if name == "shard8000":
                pr = max(c["total_pairs"] // world // c["passes"], 1)
                imgs = (Ls[:pr], Rs[:pr]) if (c["W"], c["H"]) == (cfg["W"], cfg["H"]) and len(Ls) >= pr else None
                r = measure_brief(c, name, rank, world, local_rank, args, dist, max(es // 2, 2), 3, pairs=pr, images=imgs)
                if r is not None:
                    r["scaling"] = "strong (8000 pairs in total at every N)"
            elif world == 1:
                imgs = (Ls, Rs) if (c["W"], c["H"]) == (cfg["W"], cfg["H"]) and len(Ls) >= c["pairs"] else None
                if imgs is not None:
                    imgs = (Ls[:c["pairs"]], Rs[:c["pairs"]])
                r = measure_brief(c, name, rank, world, local_rank, args, dist, es, ew, images=imgs)
                if r is not None and name == "kitti2000":
                    r["single_pair_latency_ms"] = single_pair_latency(c, local_rank)
                    r["single_pair_pairs_per_s"] = 1e3 / r["single_pair_latency_ms"]
            else:
                continue
            if r is not None:
                extras[name] = r

    next_rows = {}
    if rank == 0 and world == 1 and not args.no_extra:
        try:
            next_rows["cvorb"] = measure_cvorb(Ls, Rs, cfg["W"], cfg["H"], local_rank)
            next_rows["object_mode_frame"] = object_frame_latency(cfg, local_rank)
        except Exception as e:  # a secondary row must not take the headline line down
            next_rows["cvorb"] = {"error": repr(e)}

    if rank != 0:
        return
    total_pairs = pairs_step * world * args.steps
    value = total_pairs / (ms_dev * 1e-3)
    e2e = e2e_pairs_step_all * args.steps / (ms_e2e * 1e-3)
    if e2e_equal is not None:
        e2e_equal["value"] = pairs_step * world / (e2e_equal["ms_per_step"] * 1e-3)

    # ---- roofline --------------------------------------------------------------------------------------------------
    peak, peak_src = hbm_peak()
    stage_ms = stage_acc[False][0] / max(stage_acc[False][1], 1)  # average per submit (one sub-batch of Bsz pairs)
    e2e_stage_ms = stage_acc[True][0] / max(stage_acc[True][1], 1)
    bytes_pair, counts = stage_bytes_per_pair(cfg, tot)
    ncu = {}
    npath = ROOT / "profiles" / "ncu_kernels.json"
    if npath.exists():
        ncu = json.loads(npath.read_text())
    stages = {}
    for name, ms in zip(B.StereoBatch.STAGES, stage_ms):
        gbs = bytes_pair[name] * Bsz / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        st = {"ms_per_launch": float(ms), "algorithmic_MB_per_launch": bytes_pair[name] * Bsz / 1e6, "GBps": gbs,
              "frac_hbm": gbs / peak}
        k = ncu.get("kernels", {}).get(STAGE_KERNEL.get(name, ""), None)
        if k:  # ncu --set full capture of this build (profiles/ncu_kernels.json; tools/make_profiles.py)
            st.update({"bound": k.get("bound"), "ncu": {m: k[m] for m in ("dram_pct", "l1tex_pct", "lts_pct", "issue_pct", "sm_pct",
                                                                       "shared_GBps", "shared_peak_GBps", "shared_frac", "dram_bytes_per_launch",
                                                                       "capture_pairs") if m in k}})
        stages[name] = st
    kernels_only = {k: v for k, v in stages.items() if k not in ("ingest_copy", "result_copy")}
    dom = max(kernels_only, key=lambda k: kernels_only[k]["ms_per_launch"])
    traffic = None
    kd = ncu.get("kernels", {}).get(STAGE_KERNEL.get(dom, ""), None)
    if kd and "dram_bytes_per_launch" in kd:  # traffic is linear in the batch size
        traffic = kd["dram_bytes_per_launch"] * Bsz / float(kd.get("capture_pairs", Bsz))
    roofline = {"bound": "hbm", "kernel": dom, "achieved": stages[dom]["GBps"], "peak": peak, "unit": "GB/s",
                "frac": stages[dom]["GBps"] / peak, "traffic": traffic, "peak_source": peak_src,
                "launch_pairs": Bsz, "ms_per_launch": stages[dom]["ms_per_launch"],
                "kernel_bound": stages[dom].get("bound"),
                "note": "HBM fraction of the longest kernel as the contract asks; what actually limits each kernel (`stages[*].bound`: "
                        "hbm / shared / issue / latency) and its ncu pipe percentages are listed per stage; durations from a "
                        "single-stream pass of the same submits inside this run"}
    path_bytes = float(sum(bytes_pair.values()))
    path_gbs = path_bytes * (value / world) / 1e9
    roofline["path"] = {"algorithmic_MB_per_pair": path_bytes / 1e6, "achieved_GBps": path_gbs, "frac": path_gbs / peak,
                        "note": "all kernels of the path at the multi-stream `value` rate, per GPU"}

    # ---- CPU baseline (bounded sample, N=1 only) ---------------------------------------------------------------------
    cpu = None
    if full_affinity:
        try:
            os.sched_setaffinity(0, full_affinity)  # the CPU arm gets every host core again
        except Exception:
            pass
    if world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        ns = min(len(Ls), max(64, 4 * cores))
        arm = CpuArm(cfg)
        v_all, done_all, el_all = arm.pairs_per_sec(Ls[:ns], Rs[:ns], cores, args.cpu_seconds)
        v_ship, done_s, el_s = arm.pairs_per_sec(Ls[:ns], Rs[:ns], 1, min(4.0, args.cpu_seconds), per_pair_threads=2)
        cpu = {"value": v_all, "unit": UNIT, "cores": cores, "kind": arm.kind,
               "sample": f"{done_all} pairs of the same sequence in {el_all:.1f} s, {cores} threads x 1 pair each; {arm.describe()}",
               "as_shipped_2_threads": {"value": v_ship, "sample": f"{done_s} pairs in {el_s:.1f} s, 2 threads per pair "
                                        "(Frame.cc:161-164), 1 pair in flight"}}
        if arm.kind == "reference":  # the C port beside it, for continuity with round 1
            port = CpuArm(cfg, prefer_reference=False)
            v_port, done_p, el_p = port.pairs_per_sec(Ls[:ns], Rs[:ns], cores, min(5.0, args.cpu_seconds))
            cpu["port"] = {"value": v_port, "sample": f"{done_p} pairs in {el_p:.1f} s, {cores} threads (oracle/orb_oracle.c)"}

    bytes_per_pair_h2d = 2 * wl.img_bytes
    pcie = {"h2d_GBps_in_run": 2 * e2e_pairs_step_all * wl.img_bytes / world / (ms_e2e / args.steps * 1e-3) / 1e9, "h2d_GBps_alone": h2d_alone,
            "h2d_GBps_all_ranks": {"per_rank_min": h2d_min, "aggregate": h2d_sum},
            "h2d_GBps_common_window": h2d_win,
            "e2e_frac_of_common_window_ceiling": (e2e * bytes_per_pair_h2d / (h2d_win["aggregate_GBps"] * 1e9)) if h2d_win else None,
            "e2e_ceiling_pairs_per_s": (h2d_sum * 1e9 / bytes_per_pair_h2d) if h2d_sum else None,
            "e2e_frac_of_ceiling": (e2e * bytes_per_pair_h2d / (h2d_sum * 1e9)) if h2d_sum else None,
            "note": "h2d_GBps_all_ranks: every rank uploads its pinned images at the same time between barriers, no kernels - the "
                    "rate this host can feed N GPUs at; e2e cannot exceed e2e_ceiling_pairs_per_s (raw 8-bit images, 2 per pair)"}
    line = {
        "metric": metric_name(cfg), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "strong" if "total_pairs" in cfg else "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": cfg["workload"], "name": args.config, "pairs_per_step_per_gpu": pairs_step,
                   "sequence_pairs_per_gpu": P, "sequence_passes_per_step": wl.passes, "sub_batch_pairs": Bsz,
                   "streams": wl.nctx, "image": f"{cfg['W']}x{cfg['H']}", "params": list(cfg["params"]), "bf": cfg["bf"],
                   "maxD": wl.maxD, "pinned_input": args.pin, "cpu_affinity": sorted(affinity) if affinity else None,
                   "l2": f"inputs larger than L2: {2 * P * wl.img_bytes / 1e6:.0f} MB of distinct images per pass vs 126 MB L2",
                   **counts, "synth_seconds": t_gen},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 2 * e2e_pairs_step_all * wl.img_bytes,
                "d2h_bytes_per_step": d2h // max(nsub, 1) * sum(shard_subs),
                "ms_per_step": ms_e2e / args.steps, "pairs_per_step": e2e_pairs_step_all,
                "partition": {"kind": "equal contiguous blocks" if e2e_equal is None else
                              "contiguous blocks sized by each GPU's measured upload rate (same total; DESIGN.md §6)",
                              "pairs_per_step_by_rank": [n * wl.Bsz for n in shard_subs],
                              "ms_per_step_by_rank": None if e2e_equal is None else e2e_equal.pop("ms_per_step_by_rank", None)},
                "equal_shards": e2e_equal, "pcie": pcie,
                "stage_ms_per_submit": {n: float(v) for n, v in zip(B.StereoBatch.STAGES, e2e_stage_ms)}},
        "gpu_launches": launches_per_submit * nsub * args.steps * world,
        "roofline": roofline, "stages": stages, "cpu_baseline": cpu, "clocks": clocks, "determinism": det,
        "extra_configs": extras, "next_rows": next_rows,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="kitti1000", choices=sorted(CONFIGS))
    ap.add_argument("--pairs", type=int, default=0, help="distinct pairs per GPU (default: the configuration's)")
    ap.add_argument("--batch", type=int, default=0, help="pairs per sub-batch (one submit; default: the configuration's)")
    ap.add_argument("--streams", type=int, default=4, help="batch contexts / CUDA streams used round-robin")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--extra-steps", type=int, default=6, help="timed steps of each secondary configuration")
    ap.add_argument("--pin", default="default", choices=["default", "wc", "thp"],
                    help="pinned input buffers: cudaHostAlloc default, write-combined, or huge-page backed + cudaHostRegister")
    ap.add_argument("--affinity", default="auto", choices=["auto", "none"], help="bind each rank to the CPUs next to its GPU before allocating")
    ap.add_argument("--equal-shards", action="store_true", help="N > 1: keep equal shards in the end-to-end pass (no rate-proportional split)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary configurations")
    ap.add_argument("--no-determinism", action="store_true")
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
