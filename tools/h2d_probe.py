#!/usr/bin/env python3
"""How fast can this host feed its GPUs?  Run under torchrun with one rank per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/h2d_probe.py

Every rank owns pinned host buffers of the bench's size (2 x 466 MB = 1000 KITTI-shaped pairs) and uploads them to its GPU
with nothing else running.  Measured, device time by CUDA events, max over the participating ranks:
  * which ranks upload at the same time: one alone, pairs, halves, all  -> shows where the shared links are;
  * how the buffers were pinned: cudaHostAlloc default / write-combined / huge-page backed + cudaHostRegister;
  * one upload stream or two (two halves in flight), 32 MB chunks or whole buffers;
  * with the rank bound to the CPUs nvidia-smi lists next to its GPU before allocating, or not.
Output: one JSON document on rank 0 (gpurun_out/h2d_probe.json when that directory exists)."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import gpu_cpu_affinity  # noqa: E402
from pointslot_b200 import binding as B  # noqa: E402

BYTES = 1000 * 1242 * 375  # one view of the bench sequence


def main():
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    full = os.sched_getaffinity(0)
    aff = gpu_cpu_affinity(lr)
    dev = [torch.empty(BYTES, dtype=torch.uint8, device="cuda") for _ in range(2)]
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()

    def reduce(x):
        if world == 1:
            return x, x, x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        mx, mn, sm = t.clone(), t.clone(), t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(mn, op=dist.ReduceOp.MIN); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        return float(mn), float(mx), float(sm)

    def upload(host, active, two_streams=False, chunk=0, reps=3):
        """-> (GB/s of the slowest active rank, aggregate GB/s over the active ranks)"""
        barrier()
        gbs = 0.0
        if rank in active:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(torch.cuda.current_stream())
            s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
            for _ in range(reps):
                for v in range(2):
                    st = (s2 if v else s1) if two_streams else s1
                    with torch.cuda.stream(st):
                        if chunk:
                            for lo in range(0, BYTES, chunk):
                                dev[v][lo:lo + chunk].copy_(host[v][lo:lo + chunk], non_blocking=True)
                        else:
                            dev[v].copy_(host[v], non_blocking=True)
            torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
            e1.record(torch.cuda.current_stream())
            e1.synchronize()
            gbs = reps * 2 * BYTES / (e0.elapsed_time(e1) * 1e-3) / 1e9
        barrier()
        if world == 1:
            return gbs, gbs
        t = torch.tensor([gbs if rank in active else 1e30], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        s = torch.tensor([gbs], dtype=torch.float64, device="cuda")
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        return float(t), float(s)

    res = {"world": world, "bytes_per_rank_per_rep": 2 * BYTES, "affinity_of_gpu": {}, "runs": []}
    if world > 1:
        allaff = [None] * world
        dist.all_gather_object(allaff, sorted(aff) if aff else None)
        res["affinity_of_gpu"] = {str(i): (f"{a[0]}-{a[-1]} ({len(a)} cpus)" if a else None) for i, a in enumerate(allaff)}
    everyone = set(range(world))
    subsets = [("rank0_alone", {0}), ("all", everyone)]
    if world >= 2:
        subsets += [("ranks_0_1", {0, 1})]
    if world >= 4:
        subsets += [("ranks_0_2", {0, 2}), ("ranks_0_3", {0, 3}), ("ranks_0-3", {0, 1, 2, 3})]
    if world >= 8:
        subsets += [("ranks_0_4", {0, 4}), ("ranks_0_7", {0, 7}), ("ranks_4-7", {4, 5, 6, 7}), ("ranks_0_2_4_6", {0, 2, 4, 6})]
    for bind in (False, True):
        if bind and aff:
            try:
                os.sched_setaffinity(0, aff & full or full)
            except Exception:
                pass
        for mode in ("default", "wc", "thp"):
            try:
                bufs = [B.PinnedBuffer((BYTES,), np.uint8, mode=mode) for _ in range(2)]
            except Exception as e:  # e.g. cudaHostRegister refused
                if rank == 0:
                    res["runs"].append({"pin": mode, "bound_to_gpu_cpus": bind, "error": str(e)})
                continue
            for b in bufs:
                b.array[:] = 7  # first touch under the current affinity
            host = [torch.from_numpy(b.array) for b in bufs]
            upload(host, everyone, reps=1)  # warm
            for name, act in (subsets if (mode == "default" and not bind) else [("rank0_alone", {0}), ("all", everyone)]):
                mn, agg = upload(host, act)
                if rank == 0:
                    res["runs"].append({"pin": mode, "bound_to_gpu_cpus": bind, "who": name, "n_active": len(act),
                                        "GBps_slowest_rank": mn, "GBps_aggregate": agg})
            for two, chunk in ((True, 0), (False, 32 << 20), (True, 32 << 20)):
                mn, agg = upload(host, everyone, two_streams=two, chunk=chunk)
                if rank == 0:
                    res["runs"].append({"pin": mode, "bound_to_gpu_cpus": bind, "who": "all", "n_active": world,
                                        "two_streams": two, "chunk_MB": chunk >> 20, "GBps_slowest_rank": mn, "GBps_aggregate": agg})
            del host
            for b in bufs:
                b.close()
        os.sched_setaffinity(0, full)
    if rank == 0:
        txt = json.dumps(res, indent=1)
        print(txt)
        out = ROOT / "gpurun_out"
        if out.exists():
            (out / f"h2d_probe_n{world}.json").write_text(txt)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
