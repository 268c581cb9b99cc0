#!/usr/bin/env python
"""bench.py -- candidate windows/sec of the JDA detect path on the VGA 3-octave pyramid.

    python bench.py --gpus N --steps K --warmup W              (ours; N>1 under torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K ...    (the reference's own c/jda.c on host cores)

A step = one pass of the hot path (jdaB200DetectBatch: scan + cascade + hit read-back + host NMS)
over one batch of B synthetic 640x480 frames per GPU, args (scale 1.25, min 24, max 192, th 0.0):
BASELINE.json configs[1], 169,236 candidate windows per frame.  K steps x B frames ~ the config's
10k frames (K=20, B=512).  The batch (157 MB) is larger than L2 (126 MB) and two different batches
alternate between steps.

value : whole-job windows/s with the frames already resident in HBM (device pointers in).
e2e   : same metric through the reference-facing call with HOST frames (pinned): H2D of every frame
        and D2H of the hit records happen inside the timed region.
Prints ONE JSON line on rank 0.

    --workload cfg4   BASELINE.json configs[3]: 2845 FDDB-shaped frames of mixed sizes, sharded over the ranks in
                      contiguous blocks (STRONG scaling: the job is fixed), jdaB200DetectMixed per rank + one NCCL
                      all-gather of the detection records; every rank checks that it holds the same job-wide table.
    --workload cfg5   BASELINE.json configs[4]: hard-negative mining scan over 100,000 VGA backgrounds (sigma in
                      {0,1,2,4,6} by seed mod 5), truncated cascade (--mine-t / --mine-k = Validate's
                      current_stage_idx / current_cart_idx + 1), every survivor emitted and all-gathered (strong scaling).
The default workload (vga) is the headline the driver runs; cfg4 / cfg5 print the same JSON contract.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
MODEL = os.path.join(ROOT, "tests", "golden", "jda_shipped_f32.model")
W, H = 640, 480
ARGS = dict(scale=1.25, min_size=24, max_size=192, th=0.0)
WINDOWS_PER_FRAME = 169236
METRIC = "candidate windows/sec, VGA 3-octave pyramid (scale 1.25, min 24, max 192), full cascade"
METRIC_BY_WORKLOAD = {"vga": METRIC,
                      "cfg4": "candidate windows/sec, 2845 FDDB-shaped frames (longest side 450), full pyramid, full cascade",
                      "cfg5": "candidate windows/sec, mining scan over 100k VGA backgrounds (3-octave pyramid, truncated cascade, every survivor emitted)"}


def frame_pool(n_distinct, dist, seed0):
    from jda_b200 import synth
    return synth.make_frames(dist, n_distinct, W, H, seed0=seed0)


def tile_batch(pool, batch, shift):
    idx = (np.arange(batch) + shift) % len(pool)
    return np.ascontiguousarray(pool[idx])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        in_region = [r for r in self.rows if t0 - 0.05 <= r[0] <= t1 + 0.05]
        if not in_region:  # a very short timed region: take the samples of the warm-up just before it (same load)
            in_region = [r for r in self.rows if t0 - 1.5 <= r[0] <= t1 + 0.05]
        keep = set(id(r) for r in in_region)
        for row in self.rows:
            ts, line = row
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                mx = float(f[1])
                if id(row) in keep:
                    sm.append(float(f[0]))
                    for nm, v in zip(names, f[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(nm)
            except ValueError:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


def work_stats(c, pool_sample):
    """avg carts evaluated per window, for the algorithmic-bytes figure: counted by the library's own per-window
    trace (jdaB200Trace runs the same kernels with a trace store), not by the CPU oracle."""
    carts = wins = 0
    for img in pool_sample:
        tn, _, _ = c.trace(img, scale=ARGS["scale"], min_size=ARGS["min_size"], max_size=ARGS["max_size"])
        carts += int(tn.sum())
        wins += len(tn)
    return carts / wins


def cpu_reference_run(frames, threads, use_ref=True):
    """The reference's own jdaDetect (oracle/_ref) -- or the oracle port -- on `threads` host threads,
    one shared read-only cascador (the function is re-entrant).  Returns seconds."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pyoracle
    kind = "port"
    if use_ref and os.path.exists(pyoracle.REF_SO):
        lib = pyoracle.RefLib()
        h = lib.load(MODEL, double=False)
        kind = "reference"
    else:
        lib = pyoracle.Oracle()
        h = lib.load(MODEL, double=False)

    def one(i):
        lib.detect(h, frames[i], ARGS["scale"], 0.1, ARGS["min_size"], ARGS["max_size"], ARGS["th"])

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(one, range(len(frames))))
    dt = time.perf_counter() - t0
    lib.release(h)
    return dt, kind


def cpu_frames_run(frames, threads, kw):
    """the reference's jdaDetect over a list of frames of any sizes on `threads` host threads; returns seconds, kind"""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pyoracle
    if os.path.exists(pyoracle.REF_SO):
        lib, kind = pyoracle.RefLib(), "reference"
    else:
        lib, kind = pyoracle.Oracle(), "port"
    h = lib.load(MODEL, double=False)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda f: lib.detect(h, f, kw["scale"], 0.1, kw["min_size"], kw["max_size"], kw["th"]), frames))
    dt = time.perf_counter() - t0
    lib.release(h)
    return dt, kind


CFG4_FRAMES, CFG4_DISTINCT = 2845, 48
CFG4_ARGS = dict(scale=1.25, min_size=24, max_size=-1, th=0.0)
CFG5_FRAMES, CFG5_DISTINCT = 100000, 40


def cfg4_pool():
    from jda_b200 import synth
    return [synth.facemix_frame(11000 + i, *synth.fddb_shape(i)) for i in range(CFG4_DISTINCT)]


def cfg4_windows(api_or_oracle_count, frames):
    return sum(api_or_oracle_count(f.shape[1], f.shape[0], 1.25, 24, -1) for f in frames)


def cfg5_pool():
    from jda_b200 import synth
    return np.stack([synth.mining_background(21000 + i) for i in range(CFG5_DISTINCT)])


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = max(8, min(2 * cores, 256))  # two frames per host thread so every core stays busy for the whole step
    if a.workload in ("cfg4", "cfg5"):
        # the reference's jdaDetect on a bounded sample of the same frames; it has no truncated-cascade entry point
        # in its C API (Validate lives in the C++ tree), so the mining scan is timed as full detects of the backgrounds
        # -- on face-free frames nearly every window dies inside stage 0 either way
        from oracle import pyoracle
        if a.workload == "cfg4":
            pool = cfg4_pool()
            frames = [pool[i % len(pool)] for i in range(n)]
            kw, name = CFG4_ARGS, "fddb_2845_mixed_sizes"
            wins = cfg4_windows(pyoracle.Oracle().count_windows, frames)
        else:
            pool = cfg5_pool()
            frames = [pool[i % len(pool)] for i in range(n)]
            kw, name = ARGS, "mining_100k_backgrounds"
            wins = n * WINDOWS_PER_FRAME
        for _ in range(a.warmup):
            cpu_frames_run(frames[:max(2, n // 4)], cores, kw)
        tot, kind = 0.0, "port"
        for _ in range(a.steps):
            dt, kind = cpu_frames_run(frames, cores, kw)
            tot += dt
        val = a.steps * wins / tot
        print(json.dumps({"impl": "reference", "metric": METRIC_BY_WORKLOAD[a.workload], "value": val, "unit": "windows/s",
                          "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * tot / a.steps,
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32+u8",
                          "data": "synthetic", "config": {"workload": name, "frames_per_step": n, **kw},
                          "cpu_baseline": {"value": val, "unit": "windows/s", "cores": cores, "kind": kind,
                                           "sample": "%d frames of the workload per step, %d threads over jdaDetect" % (n, cores)},
                          "e2e": {"value": val, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}), flush=True)
        return
    pool = frame_pool(min(n, 96), a.dist, 0)
    pool = pool[np.arange(n) % len(pool)]
    for _ in range(a.warmup):
        cpu_reference_run(pool[:max(2, n // 4)], cores)
    tot = 0.0
    kind = "port"
    for _ in range(a.steps):
        dt, kind = cpu_reference_run(pool, cores)
        tot += dt
    val = a.steps * n * WINDOWS_PER_FRAME / tot
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": "windows/s", "n_gpus": a.gpus,
           "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * tot / a.steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32+u8", "data": "synthetic",
           "config": {"workload": "vga_3oct_%s" % a.dist, "frames_per_step": n, "frame": [W, H], **ARGS},
           "cpu_baseline": {"value": val, "unit": "windows/s", "cores": cores, "kind": kind,
                            "sample": "%d %s VGA frames per step, %d threads over jdaDetect" % (n, a.dist, cores)},
           "e2e": {"value": val, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="frames per step per GPU")
    ap.add_argument("--dist", default="mix", choices=["mix", "noise", "blur6", "facemix"])
    ap.add_argument("--distinct", type=int, default=96, help="distinct synthetic frames tiled into a batch")
    ap.add_argument("--workload", default="vga", choices=["vga", "cfg4", "cfg5"])
    ap.add_argument("--mine-t", type=int, default=2, help="cfg5: full stages of the truncated cascade (current_stage_idx)")
    ap.add_argument("--mine-k", type=int, default=0, help="cfg5: carts of the unfinished stage (current_cart_idx + 1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-breakdown", action="store_true")
    a = ap.parse_args()
    if a.impl == "reference":
        return run_reference_arm(a)

    import torch
    from jda_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the detect path has no CPU fallback")
    torch.cuda.set_device(local)
    dist_on = world > 1
    if dist_on:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    c = api.Cascador(MODEL, double=False, device=local)
    stream = torch.cuda.current_stream()
    c.set_stream(stream.cuda_stream)

    B = a.batch
    if a.workload == "vga":
        pool = frame_pool(a.distinct, a.dist, 100000 * rank)
        host = [torch.from_numpy(tile_batch(pool, B, s)).pin_memory() for s in (0, 37)]
        dev = [h.cuda(non_blocking=False) for h in host]
        torch.cuda.synchronize()

    from jda_b200 import shard

    gather = shard.RecordGather(5 + 2 * c.L, device="cuda") if dist_on else None
    exchange = {"pool": None, "fut": None}
    if dist_on:
        from concurrent.futures import ThreadPoolExecutor
        exchange["pool"] = ThreadPoolExecutor(1)
        exchange["pool"].submit(torch.cuda.set_device, local).result()

    def _exchange(res):
        table = gather.finish()
        gather.start(shard.pack_records_flat(*res, frame0=rank * B))
        return table

    def exchange_wait():
        """collect what is still in flight: the helper thread's work, then the last all-gather"""
        if exchange["fut"] is not None:
            exchange["fut"].result()
            exchange["fut"] = None
        if gather is not None:
            gather.finish()

    def gather_records(res):
        """the one exchange step of the path: a single NCCL all-gather of the fixed-stride detection records
        (frame, x, y, size, score, 54 landmark floats) so every rank holds the job-wide table.  Packing, the
        asynchronous launch and the collection of the previous batch's table run on a helper thread while the main
        thread is already inside the next batch's detect call (which releases the GIL); whatever is still in
        flight at the end of a timed region is collected before the region ends."""
        if not dist_on:
            return
        if exchange["fut"] is not None:
            exchange["fut"].result()
        exchange["fut"] = exchange["pool"].submit(_exchange, res)

    # Steps are pipelined through jdaB200Submit / jdaB200Collect (two batches in flight on the handle): step i submits
    # batch i and then collects batch i - 1, so the host -> device copy of a batch and the host's NMS / relocation of
    # the one before run beside the scan; the last batch is collected before the timed region ends (finish_pending).
    pending = []
    zero_stats = {k: 0 for k in ("ms_scan", "ms_cascade", "ms_h2d", "ms_d2h", "ms_host", "raw_hits", "stage0_survivors",
                                 "detections", "scan_launches", "cascade_launches", "resize_launches")}

    def collect_oldest():
        res = c.collect(pending.pop(0))
        gather_records(res)
        return c.last_stats

    def finish_pending():
        tot = dict(zero_stats)
        while pending:
            st = collect_oldest()
            for k in tot:
                tot[k] += st[k]
        return tot

    def step_resident(i):
        d = dev[i & 1]
        pending.append(c.submit(None, device_ptr=d.data_ptr(), shape=(B, H, W), **ARGS))
        return collect_oldest() if len(pending) == 2 else zero_stats

    def step_e2e(i):
        pending.append(c.submit(host[i & 1].numpy(), **ARGS))
        return collect_oldest() if len(pending) == 2 else zero_stats

    def step_e2e_sync(i):   # the one-call form (jdaB200DetectBatchFlat), for comparison
        res = c.detect_batch(host[i & 1].numpy(), flat=True, **ARGS)
        gather_records(res)
        return c.last_stats

    def timed(step_fn, steps, warmup, sample_clocks=False, warm_fn=None):
        # the sampler starts before the warm-up: nvidia-smi needs ~0.5 s before its first line
        sampler = ClockSampler(local) if sample_clocks else None
        for i in range(warmup):
            (warm_fn or step_fn)(i)
        if a.workload == "vga":
            finish_pending()
        exchange_wait()
        torch.cuda.synchronize()
        if dist_on:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record(stream)
        acc = {"ms_scan": 0.0, "ms_cascade": 0.0, "ms_h2d": 0.0, "ms_d2h": 0.0, "ms_host": 0.0,
               "raw_hits": 0, "stage0_survivors": 0, "detections": 0, "launches": 0}
        for i in range(steps):
            st = step_fn(i)
            for k in ("ms_scan", "ms_cascade", "ms_h2d", "ms_d2h", "ms_host", "raw_hits", "stage0_survivors",
                      "detections"):
                acc[k] += st[k]
            acc["launches"] += st["scan_launches"] + st["cascade_launches"] + st["resize_launches"]
        if a.workload == "vga":   # the batch still in flight is collected inside the timed region
            st = finish_pending()
            for k in ("ms_scan", "ms_cascade", "ms_h2d", "ms_d2h", "ms_host", "raw_hits", "stage0_survivors", "detections"):
                acc[k] += st[k]
            acc["launches"] += st["scan_launches"] + st["cascade_launches"] + st["resize_launches"]
        exchange_wait()           # the last batch's exchange completes inside the timed region
        e1.record(stream)
        torch.cuda.synchronize()
        if dist_on:
            dist.barrier()
        torch.cuda.synchronize()
        t1 = time.time()
        ms = e0.elapsed_time(e1)
        if dist_on:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        clocks = sampler.stop(t0, t1) if sampler else None
        return ms, acc, clocks

    def table_check(table, n_local):
        """every rank must hold the same job-wide table: compare a checksum and the record count across ranks"""
        chk = int(np.ascontiguousarray(table, np.float32).view(np.uint32).astype(np.uint64).sum()) & ((1 << 62) - 1)
        mine = torch.tensor([chk, len(table), n_local], dtype=torch.int64, device="cuda")
        if dist_on:
            allv = torch.empty(world * 3, dtype=torch.int64, device="cuda")
            dist.all_gather_into_tensor(allv, mine)
            allv = allv.cpu().numpy().reshape(world, 3)
        else:
            allv = mine.cpu().numpy().reshape(1, 3)
        fid = table[:, 0]
        return {"ranks_agree": bool((allv[:, 0] == allv[0, 0]).all() and (allv[:, 1] == allv[0, 1]).all()),
                "records": int(len(table)), "records_equal_sum_of_ranks": bool(int(allv[:, 2].sum()) == len(table)),
                "frame_order": bool((np.diff(fid) >= 0).all()) if len(fid) > 1 else True}

    def finish_line(out, metric_windows_per_step, cpu_fn):
        """clocks / cpu_baseline / print, shared by the cfg4 and cfg5 workloads"""
        if rank == 0:
            if not a.no_cpu_baseline and world == 1:
                out["cpu_baseline"] = cpu_fn()
            print(json.dumps(out), flush=True)
        c.close()
        if dist_on:
            exchange["pool"].shutdown()
            dist.destroy_process_group()

    # ------------------------------------------------------------------ config 4: 2845 FDDB-shaped frames, sharded
    if a.workload == "cfg4":
        poolm = cfg4_pool()
        lo, hi = shard.shard_range(CFG4_FRAMES, rank, world)
        sizes = [poolm[i % CFG4_DISTINCT].size for i in range(lo, hi)]
        pin = torch.empty(max(sum(sizes), 1), dtype=torch.uint8).pin_memory().numpy()
        fr, o = [], 0
        for i in range(lo, hi):
            f = poolm[i % CFG4_DISTINCT]
            v = pin[o:o + f.size].reshape(f.shape)
            v[:] = f
            fr.append(v)
            o += f.size
        per_distinct = [api.count_windows(f.shape[1], f.shape[0], 1.25, 24, -1) for f in poolm]
        job_windows = sum(per_distinct[i % CFG4_DISTINCT] for i in range(CFG4_FRAMES))
        gsync = shard.RecordGather(5 + 2 * c.L, device="cuda") if dist_on else None
        last = {}

        def step4(i):
            res = c.detect_many(fr, **CFG4_ARGS) if fr else []
            st = dict(c.last_stats) if fr else {k: 0 for k in ("ms_scan", "ms_cascade", "ms_h2d", "ms_d2h", "ms_host", "raw_hits", "stage0_survivors", "detections", "scan_launches", "cascade_launches", "resize_launches")}
            rec = shard.pack_records(res, frame0=lo, landmark_n=c.L)
            if dist_on:
                gsync.start(rec)
                last["table"] = gsync.finish()
            else:
                last["table"] = rec
            last["n_local"] = len(rec)
            return st

        ms, acc, clocks = timed(step4, a.steps, a.warmup, sample_clocks=True)
        value = a.steps * job_windows / (ms * 1e-3)
        check = table_check(last["table"], last["n_local"])
        h2d = sum(sizes)
        if dist_on:
            t = torch.tensor([h2d], dtype=torch.int64, device="cuda")
            dist.all_reduce(t)
            h2d = int(t.item())
        out = {"metric": METRIC_BY_WORKLOAD["cfg4"], "value": value, "unit": "windows/s", "n_gpus": world, "steps": a.steps,
               "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": "f32+u8", "data": "synthetic",
               "config": {"workload": "fddb_2845_mixed_sizes", "frames": CFG4_FRAMES, "distinct_frames": CFG4_DISTINCT,
                          "windows_per_step": job_windows, **CFG4_ARGS, "l2": "inputs larger than L2 (425 MB of frames per step)",
                          "parallelism": "contiguous frame blocks over %d rank(s), jdaB200DetectMixed per rank, one NCCL "
                                         "all-gather of detection records per step (inside the timed region)" % world,
                          "note": "host frames (pinned) in: every step copies its frames host->device, so value IS the "
                                  "end-to-end figure; e2e repeats it"},
               "e2e": {"value": value, "unit": "windows/s", "h2d_bytes_per_step": h2d,
                       "d2h_bytes_per_step": int(check["records"]) * (6 + 2 * c.L) * 4},
               "gpu_launches": acc["launches"], "clocks": clocks,
               "kernel_ms_per_step_rank0": {"k2_scan": acc["ms_scan"] / a.steps, "k3_cascade": acc["ms_cascade"] / a.steps,
                                            "h2d": acc["ms_h2d"] / a.steps, "host_nms": acc["ms_host"] / a.steps},
               "gather_check": check}

        def cpu4():
            cores = os.cpu_count() or 1
            n = max(8, min(2 * cores, 256))
            sample = [poolm[i % CFG4_DISTINCT] for i in range(n)]
            dt, kind = cpu_frames_run(sample, cores, CFG4_ARGS)
            wins = sum(per_distinct[i % CFG4_DISTINCT] for i in range(n))
            return {"value": wins / dt, "unit": "windows/s", "cores": cores, "kind": kind,
                    "sample": "%d of the 2845 frames, %d threads over jdaDetect (%.1f s)" % (n, cores, dt)}
        return finish_line(out, job_windows, cpu4)

    # ------------------------------------------------------------------ config 5: mining scan over 100k backgrounds
    if a.workload == "cfg5":
        pool5 = cfg5_pool()
        lo, hi = shard.shard_range(CFG5_FRAMES, rank, world)
        # two resident batches of B frames alternate (each larger than L2); frame g of the job is pool5[g % 40]
        host5 = [torch.from_numpy(tile_batch(pool5, B, s)).pin_memory() for s in (0, 17)]
        dev5 = [h.cuda(non_blocking=False) for h in host5]
        torch.cuda.synchronize()
        mine_kw = dict(scale=1.25, min_size=24, max_size=192, th=0.0, t_limit=a.mine_t, k_limit=a.mine_k,
                       flags=api.RAW_HITS | api.NO_FINAL_TH)
        keys = ("ms_scan", "ms_cascade", "ms_h2d", "ms_d2h", "ms_host", "raw_hits", "stage0_survivors", "detections",
                "scan_launches", "cascade_launches", "resize_launches")
        last = {"table": None, "n_local": 0}

        def run5(n_frames, from_host):
            """the rank's frames in batches of B, two batches in flight (jdaB200Submit / jdaB200Collect)"""
            tot = {k: 0 for k in keys}
            inflight = []       # (ticket, first frame of the batch)

            def collect_one():
                t, f0 = inflight.pop(0)
                res = c.collect(t)
                for k in keys:
                    tot[k] += c.last_stats[k]
                last["n_local"] = len(res[2])
                if dist_on:
                    if exchange["fut"] is not None:
                        last["table"] = exchange["fut"].result()
                    exchange["fut"] = exchange["pool"].submit(_exchange5, res, lo + f0)
                else:
                    last["table"] = shard.pack_records_flat(*res, frame0=lo + f0)

            f = 0
            bi = 0
            while f < n_frames:
                nb = min(B, n_frames - f)
                if from_host:
                    t = c.submit(host5[bi & 1].numpy()[:nb], **mine_kw)
                else:
                    t = c.submit(None, device_ptr=dev5[bi & 1].data_ptr(), shape=(nb, H, W), **mine_kw)
                inflight.append((t, f))
                if len(inflight) == 2:
                    collect_one()
                f += nb
                bi += 1
            while inflight:
                collect_one()
            return tot

        def _exchange5(res, frame0):
            table = gather.finish()
            gather.start(shard.pack_records_flat(*res, frame0=frame0))
            return table

        def exchange_wait5():
            if exchange["fut"] is not None:
                exchange["fut"].result()
                exchange["fut"] = None
            if gather is not None:
                t = gather.finish()
                if t is not None:
                    last["table"] = t
        exchange_wait_vga = exchange_wait

        def step5(i):
            st = run5(hi - lo, False)
            exchange_wait5()
            return st

        def warm5(i):
            st = run5(min(2 * B, hi - lo), False)
            exchange_wait5()
            return st

        ms, acc, clocks = timed(step5, a.steps, a.warmup, sample_clocks=True, warm_fn=warm5)
        value = a.steps * CFG5_FRAMES * WINDOWS_PER_FRAME / (ms * 1e-3)
        check = table_check(last["table"], last["n_local"])     # the table of the job's LAST batch on every rank
        # e2e: the same loop from pinned host batches over a bounded slice of the share (1/8, at least 4 batches)
        n_e = min(hi - lo, max(4 * B, (hi - lo) // 8))

        def step5e(i):
            st = run5(n_e, True)
            exchange_wait5()
            return st
        ms_e, acc_e, _ = timed(step5e, 1, 1, warm_fn=lambda i: (run5(min(B, hi - lo), True), exchange_wait5())[0])
        ne_all = n_e
        if dist_on:
            t = torch.tensor([n_e], dtype=torch.int64, device="cuda")
            dist.all_reduce(t)
            ne_all = int(t.item())
        e2e_value = ne_all * WINDOWS_PER_FRAME / (ms_e * 1e-3)
        out = {"metric": METRIC_BY_WORKLOAD["cfg5"], "value": value, "unit": "windows/s", "n_gpus": world, "steps": a.steps,
               "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": "f32+u8", "data": "synthetic",
               "config": {"workload": "mining_100k_backgrounds", "frames": CFG5_FRAMES, "distinct_frames": CFG5_DISTINCT,
                          "sigmas": [0, 1, 2, 4, 6], "frame": [W, H], "windows_per_frame": WINDOWS_PER_FRAME,
                          "batch": B, "full_stages": a.mine_t, "carts_of_unfinished_stage": a.mine_k,
                          "l2": "inputs larger than L2 (157 MB batches, two alternate)",
                          "warmup_step": "two batches per rank (a timed step is the rank's whole share)",
                          "parallelism": "contiguous frame blocks over %d rank(s); per batch one NCCL all-gather of the "
                                         "survivor records (frame, x, y, size, score, 54 shape floats), launched under "
                                         "the next batch's scan" % world},
               "e2e": {"value": e2e_value, "unit": "windows/s", "h2d_bytes_per_step": ne_all * W * H,
                       "d2h_bytes_per_step": int(acc_e["raw_hits"]) * (6 + 2 * c.L) * 4 * world,
                       "frames": ne_all, "note": "pinned host batches in, over 1/8 of the job"},
               "gpu_launches": acc["launches"], "clocks": clocks,
               "kernel_ms_per_step_rank0": {"k2_scan": acc["ms_scan"] / a.steps, "k3_cascade": acc["ms_cascade"] / a.steps,
                                            "d2h": acc["ms_d2h"] / a.steps, "host": acc["ms_host"] / a.steps},
               "survivors_per_step_rank0": acc["raw_hits"] / a.steps, "gather_check_last_batch": check}

        def cpu5():
            cores = os.cpu_count() or 1
            n = max(8, min(2 * cores, 256))
            sample = [pool5[i % CFG5_DISTINCT] for i in range(n)]
            dt, kind = cpu_frames_run(sample, cores, ARGS)
            return {"value": n * WINDOWS_PER_FRAME / dt, "unit": "windows/s", "cores": cores, "kind": kind,
                    "sample": "%d of the backgrounds through full jdaDetect (the C API has no truncated cascade), "
                              "%d threads (%.1f s)" % (n, cores, dt)}
        return finish_line(out, 0, cpu5)

    ms, acc, clocks = timed(step_resident, a.steps, a.warmup, sample_clocks=True)
    total_windows = a.steps * B * WINDOWS_PER_FRAME * world
    value = total_windows / (ms * 1e-3)

    e2e_steps = max(3, a.steps // 2)
    ms_e, acc_e, _ = timed(step_e2e, e2e_steps, 2)
    e2e_value = e2e_steps * B * WINDOWS_PER_FRAME * world / (ms_e * 1e-3)
    # the same call from plain pageable memory (what a caller that never heard of cudaHostAlloc passes)
    pageable = [np.array(h.numpy(), copy=True) for h in host]

    def step_e2e_pageable(i):
        pending.append(c.submit(pageable[i & 1], **ARGS))
        return collect_oldest() if len(pending) == 2 else zero_stats
    ms_p, _, _ = timed(step_e2e_pageable, 4, 2)   # (two warm-up steps: each of the handle's two scratch sets grows its pinned staging once)
    e2e_pageable = 4 * B * WINDOWS_PER_FRAME * world / (ms_p * 1e-3)
    del pageable
    ms_s, _, _ = timed(step_e2e_sync, 3, 1)
    e2e_sync = 3 * B * WINDOWS_PER_FRAME * world / (ms_s * 1e-3)
    rec_bytes = (6 + 2 * c.L) * 4
    d2h = int(acc_e["raw_hits"] / e2e_steps) * rec_bytes + 22 * 4
    # the reference's own program shape: host threads over jdaDetect, one frame per call (what `--impl reference` times
    # on the CPU); concurrent calls are coalesced into mixed-size batches inside the library.  Wall clock, one rank.
    threads_fig = None
    if world == 1:
        import ctypes as C
        from concurrent.futures import ThreadPoolExecutor
        L_ = api.lib()
        fr = host[0].numpy()
        nthr, nfr = 32, min(B, 512)
        ptrs = [fr[i].ctypes.data_as(C.POINTER(C.c_ubyte)) for i in range(nfr)]

        def worker(t):      # a host thread with its own share of the frames, one frame per call
            for i in range(t, nfr, nthr):
                L_.jdaResultRelease(L_.jdaDetect(c._h, ptrs[i], W, H, ARGS["scale"], 0.1, ARGS["min_size"], ARGS["max_size"], ARGS["th"]))
        with ThreadPoolExecutor(nthr) as ex:
            list(ex.map(worker, range(nthr)))      # warm-up
            c0 = c.coalescing_stats()
            t0 = time.perf_counter()
            for _ in range(3):
                list(ex.map(worker, range(nthr)))
            dt = time.perf_counter() - t0
            c1 = c.coalescing_stats()
        threads_fig = {"value": 3 * nfr * WINDOWS_PER_FRAME / dt, "unit": "windows/s", "host_threads": nthr,
                       "frames_per_call": 1, "calls": c1[0] - c0[0], "device_batches": c1[1] - c0[1], "largest_batch": c1[2],
                       "timing": "host wall clock over 3 x %d calls" % nfr}

    out = {"metric": METRIC, "value": value, "unit": "windows/s", "n_gpus": world, "steps": a.steps,
           "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32+u8", "data": "synthetic",
           "config": {"workload": "vga_3oct_%s" % a.dist, "frames_per_step_per_gpu": B, "frame": [W, H],
                      "windows_per_frame": WINDOWS_PER_FRAME, **ARGS, "distinct_frames": a.distinct,
                      "l2": "inputs larger than L2 (157 MB batch, two batches alternate)",
                      "model": "shipped T=5 K=540 L=27 (tests/golden/jda_shipped_f32.model)",
                      "parallelism": "frames sharded, %d rank(s), NCCL all-gather of detections" % world},
           "e2e": {"value": e2e_value, "unit": "windows/s", "h2d_bytes_per_step": B * W * H * world,
                   "d2h_bytes_per_step": d2h * world, "steps": e2e_steps, "ms_per_step": ms_e / e2e_steps,
                   "host_memory": "pinned", "api": "jdaB200Submit / jdaB200Collect, two batches in flight",
                   "pageable_host_memory_value": e2e_pageable, "one_call_api_value": e2e_sync,
                   "jdaDetect_from_host_threads": threads_fig},
           "gpu_launches": acc["launches"], "clocks": clocks,
           "kernel_ms_per_step": {"k2_scan": acc["ms_scan"] / a.steps, "k3_cascade": acc["ms_cascade"] / a.steps,
                                  "d2h": acc["ms_d2h"] / a.steps, "host_nms": acc["ms_host"] / a.steps},
           "per_step": {"stage0_survivors": acc["stage0_survivors"] / a.steps, "raw_hits": acc["raw_hits"] / a.steps,
                        "detections": acc["detections"] / a.steps}}

    if rank == 0:
        # roofline of the dominant kernel (k2_scan): SURVEY.md 8(d) figure (A), logical bytes touched
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        carts_pw = work_stats(c, pool[:6])
        bytes_pw = 118.0 * carts_pw + 216.0      # per window inside k2 (stage 0); survivors' stages are k3's
        k2_s = acc["ms_scan"] / a.steps * 1e-3
        achieved = B * WINDOWS_PER_FRAME * bytes_pw / k2_s / 1e9
        # counters of the dominant kernel come from the committed ncu capture (profiles/k2_capture.json, written by
        # tools/ncu_digest.py on the GPU box) and are tied to the tree by the hash of csrc/: a capture taken on other
        # kernel sources is reported as stale instead of being quoted as if it were this tree's
        from jda_b200 import buildinfo
        cap, sha = None, buildinfo.source_sha()
        try:
            cap = json.load(open(os.path.join(ROOT, "profiles", "k2_capture.json")))
        except Exception:
            pass
        fresh = bool(cap) and cap.get("source_sha") == sha
        if cap and not fresh:
            print("bench.py: profiles/k2_capture.json was taken on csrc %s, this tree is %s -- re-run tools/gpu_round.sh"
                  % (cap.get("source_sha"), sha), file=sys.stderr)
        traffic = (cap["dram_bytes_read"] + cap["dram_bytes_write"]) / cap["frames"] * B if cap else None
        out["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                           "frac": achieved / peak,
                           "traffic": traffic, "kernel": "k2_scan",
                           # physical DRAM bytes per launch (ncu) over this run's launch time, against the same peak
                           "dram_frac": (traffic / k2_s / 1e9 / peak) if traffic else None,
                           "capture": {"source_sha": cap.get("source_sha") if cap else None, "tree_sha": sha,
                                       "matches_tree": fresh, "frames": cap.get("frames") if cap else None,
                                       "file": "profiles/k2_capture.json"},
                           "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                           "note": "frac: logical (algorithmic touched) bytes, 118 B x carts/window + 216 B, served from "
                                   "shared memory -- may exceed 1; dram_frac: the bytes that really cross HBM "
                                   "(frames + tables + survivor records, ~2 B/window); see DESIGN.md",
                           "carts_per_window": carts_pw, "algorithmic_bytes_per_window": bytes_pw,
                           "k2_windows_per_s": B * WINDOWS_PER_FRAME / k2_s}
        # The unit that actually binds k2_scan is the shared-memory data pipe: 1 wavefront / clk / SM (measured:
        # tools/probes/lds_probe.cu -> profiles/r1g_lds_probe.txt).  Wavefronts per window from the capture above,
        # the rate from this run.
        if cap:
            wf_per_window = cap["smem_wavefronts"] / (cap["frames"] * WINDOWS_PER_FRAME)
            sm_hz = float((clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0) * 1e6
            wf_rate = wf_per_window * B * WINDOWS_PER_FRAME / k2_s
            out["roofline_onchip"] = {"bound": "shared-memory data pipe (LSU wavefronts)", "kernel": "k2_scan",
                                      "achieved": wf_rate / 1e12, "peak": 148 * sm_hz / 1e12, "unit": "Twavefronts/s",
                                      "frac": wf_rate / (148 * sm_hz), "wavefronts_per_window": wf_per_window,
                                      "bank_conflict_share": cap["smem_bank_conflicts"] / cap["smem_wavefronts"],
                                      "capture_matches_tree": fresh,
                                      "ncu_pipe_pct_in_capture": cap.get("lsu_data_pipe_pct"),  # shared + global wavefronts of the whole launch
                                      "peak_source": "1 wavefront/clk/SM measured by tools/probes/lds_probe.cu x 148 SMs x SM clock"}
        if not a.no_cpu_baseline and world == 1:  # the reported CPU baseline is an N=1 item
            cores = os.cpu_count() or 1
            n = max(8, min(2 * cores, 256))
            dt, kind = cpu_reference_run(pool[np.arange(n) % len(pool)], cores)
            out["cpu_baseline"] = {"value": n * WINDOWS_PER_FRAME / dt, "unit": "windows/s", "cores": cores,
                                   "kind": kind, "sample": "%d %s VGA frames, %d threads over jdaDetect (%.1f s)"
                                   % (n, a.dist, cores, dt)}
    if not a.no_breakdown and not dist_on:
        # the three SURVEY.md 8(d) distributions separately (short runs, resident frames)
        by = {}
        for d in ("noise", "blur6", "facemix"):
            p = frame_pool(32, d, 5000)
            t = torch.from_numpy(tile_batch(p, B, 0)).cuda()
            torch.cuda.synchronize()

            def fn(i, t=t):
                c.detect_batch(None, device_ptr=t.data_ptr(), shape=(B, H, W), unpack=False, **ARGS)
                return c.last_stats
            m, ac, _ = timed(fn, 5, 2)
            by[d] = {"windows_per_s": 5 * B * WINDOWS_PER_FRAME / (m * 1e-3), "k2_ms": ac["ms_scan"] / 5,
                     "k3_ms": ac["ms_cascade"] / 5}
            del t
        out["by_distribution"] = by
    if not a.no_breakdown and not dist_on:
        # the other BASELINE.json configs, as context (not the headline): config 3 = one 1080p frame per
        # plain jdaDetect call (5-octave pyramid), config 5 = mining scan (first stage only, every survivor)
        from jda_b200 import synth
        ex = {}
        hd = [synth.facemix_frame(7000 + i, 1920, 1080) for i in range(4)]
        for i in range(10):
            c.detect(hd[i % 4], 1.25, 0.1, 24, 768, 0.0)
        lat = []
        for i in range(100):
            t0 = time.perf_counter()
            c.detect(hd[i % 4], 1.25, 0.1, 24, 768, 0.0)
            lat.append((time.perf_counter() - t0) * 1e3)
        ex["cfg3_1080p_5oct_jdaDetect_ms"] = {"p50": float(np.percentile(lat, 50)), "p99": float(np.percentile(lat, 99)),
                                              "windows_per_frame": 1245202}
        if not a.no_cpu_baseline:   # the reference's own jdaDetect on the same frames, one host thread per frame
            dt, kind = cpu_frames_run(hd, min(4, os.cpu_count() or 1), dict(scale=1.25, min_size=24, max_size=768, th=0.0))
            ex["cfg3_1080p_5oct_jdaDetect_ms"]["cpu_%s_ms_per_frame" % kind] = dt * 1e3 / len(hd) * min(4, os.cpu_count() or 1)
        vga = synth.noise_frame(1)
        for i in range(10):
            c.detect(vga, 1.25, 0.1, 24, 192, 0.0)
        lat = []
        for i in range(100):
            t0 = time.perf_counter()
            c.detect(vga, 1.25, 0.1, 24, 192, 0.0)
            lat.append((time.perf_counter() - t0) * 1e3)
        ex["vga_3oct_single_frame_jdaDetect_ms"] = {"p50": float(np.percentile(lat, 50)), "p99": float(np.percentile(lat, 99))}
        p = frame_pool(32, "noise", 9000)
        t = torch.from_numpy(tile_batch(p, B, 0)).cuda()
        torch.cuda.synchronize()

        def mine(i, t=t):
            c.detect_batch(None, device_ptr=t.data_ptr(), shape=(B, H, W), unpack=False, scale=1.25, min_size=24,
                           max_size=192, th=0.0, t_limit=1, flags=api.RAW_HITS | api.NO_FINAL_TH)
            return c.last_stats
        m, ac, _ = timed(mine, 5, 2)
        ex["cfg5_mining_stage1_noise"] = {"windows_per_s": 5 * B * WINDOWS_PER_FRAME / (m * 1e-3),
                                          "survivors_per_step": ac["raw_hits"] / 5}
        del t
        # config 4: 2845 FDDB-shaped frames (longest side 450) of mixed sizes through jdaB200DetectMixed -- host
        # frames (pinned) in, H2D inside the timed region, one launch per kernel over a common canvas; next to it
        # the older scheme (one batch call per distinct shape)
        shapes = [synth.fddb_shape(s) for s in range(48)]
        poolm = [synth.facemix_frame(11000 + i, *shapes[i]) for i in range(48)]
        nfr = 2845
        pin = torch.empty(sum(poolm[i % 48].size for i in range(nfr)), dtype=torch.uint8).pin_memory().numpy()
        fr, o = [], 0
        for i in range(nfr):
            f = poolm[i % 48]
            v = pin[o:o + f.size].reshape(f.shape)
            v[:] = f
            fr.append(v)
            o += f.size
        wins4 = sum(api.count_windows(f.shape[1], f.shape[0], 1.25, 24, -1) for f in fr)

        def run4(group):
            t0 = time.perf_counter()
            c.detect_many(fr, group=group, th=0.0, unpack=False) if not group else c.detect_many(fr, group=True, th=0.0)
            return time.perf_counter() - t0
        run4(False)
        dt = min(run4(False) for _ in range(3))
        st4 = dict(c.last_stats)
        dtg = run4(True)
        ex["cfg4_fddb_2845_mixed_sizes"] = {"windows": wins4, "windows_per_s": wins4 / dt, "ms": dt * 1e3,
                                            "k2_ms": st4["ms_scan"], "k3_ms": st4["ms_cascade"],
                                            "h2d_ms": st4["ms_h2d"], "host_nms_ms": st4["ms_host"],
                                            "h2d_bytes": int(o), "per_shape_batches_windows_per_s": wins4 / dtg,
                                            "distinct_shapes": len({f.shape for f in fr})}
        # the reference's double-precision C++ detector (JoinCascador::Detect, fddb.method 1,
        # config.json's fddb settings: min 20, step 5, scale 1.2) -- host frames in, 256 VGA frames per call; the
        # CPU figure beside it is the oracle's C restatement on one core (the C++ detector itself cannot be built here)
        fcpp = tile_batch(frame_pool(32, "mix", 13000), 256, 0)
        wcpp = api.count_windows_cpp(W, H)
        c.detect_cpp(fcpp[:8])
        c.detect_cpp(fcpp)
        t0 = time.perf_counter()
        for _ in range(3):
            c.detect_cpp(fcpp)
        dtc = (time.perf_counter() - t0) / 3
        stc = dict(c.last_stats)
        from oracle import pyoracle
        oc = pyoracle.OracleCpp()
        hc = oc.load(MODEL, double=False)
        t0 = time.perf_counter()
        for i in range(4):
            oc.detect(hc, fcpp[i])
        dto = (time.perf_counter() - t0) / 4
        oc.release(hc)
        ex["cpp_f64_detector_vga_256_frames"] = {"windows_per_frame": wcpp, "windows_per_s": 256 * wcpp / dtc,
                                                 "ms_per_call": dtc * 1e3, "k2_prefilter_ms": stc["ms_scan"],
                                                 "k4_f64_ms": stc["ms_cascade"],
                                                 "prefilter_survivors": stc["stage0_survivors"], "faces_pre_nms": stc["raw_hits"],
                                                 "cpu_port_1core_windows_per_s": wcpp / dto}
        out["other_configs"] = ex
    if rank == 0:
        print(json.dumps(out), flush=True)
    c.close()
    if dist_on:
        exchange["pool"].shutdown()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
