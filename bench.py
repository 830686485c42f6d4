#!/usr/bin/env python
"""bench.py -- sources/sec of the ELBO(+gradient) hot path on synthetic 5-band SDSS-shaped fields.

    python bench.py --gpus N --steps K --warmup W            # CUDA library (one rank per GPU; torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the CPU path (oracle port), rank 0 only

Workload (BASELINE.json configs[3]): a synthetic stripe of F fields x 1000 sources, 5 bands, 2048 x 1489
px/band; every source is one (ElboArgs, vp) task = the source + its find_neighbors neighbours, active_sources
= [1] (ParallelRun.jl:236-253).  The stripe's tasks are sharded round-robin-by-cost across the N ranks (no
data-path collective): total work is fixed => "scaling": "strong".  One step = one evaluation
of every task (ELBO + gradient; the Hessian mode is measured beside it and reported under "hessian").
Inputs are larger than L2 (each rank reads > 126 MB of distinct pixel records per step).

Outside every timed region a seeded sample of the benched plan's tasks goes through the CPU oracle and the line
carries "parity_check" (value / gradient / Hessian relative errors, pixel-visit counters equal) for both modes, at
every N.  Each leg is timed three times (K steps each): the headline is the first, "stability" holds min / median.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# ALGORITHMIC flop per pixel-visit (the contract figure of roofline.achieved)
F_ACTIVE = {0: 570.0, 1: 2600.0, 2: 23700.0}
F_INACTIVE = 548.0
BYTES_PER_VISIT = 20.0      # f32 pixel + f32 sky + f64 constant + u8 mask + amortised per-source tables
NOMINAL_FP64_TFLOPS = 37.0   # HGX B200 datasheet, 296 TF / 8


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--fields", type=int, default=10, help="fields in the stripe (1000 sources each)")
    ap.add_argument("--sources-per-field", type=int, default=1000)
    ap.add_argument("--cpu-sample", type=int, default=1000, help="tasks per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hessian", action="store_true")
    ap.add_argument("--no-render", action="store_true", help="skip the full-image expectation render leg (row f.4)")
    ap.add_argument("--no-maximize", action="store_true", help="skip the config-5 leg (full Newton loop on field 0)")
    ap.add_argument("--maximize-timeline", action="store_true",
                    help="diagnosis: rerun the maximize leg with a sync per iteration and report [active sources, "
                         "evaluation ms, newton step ms] per lock-step iteration")
    ap.add_argument("--no-single", action="store_true", help="skip the single-call latency leg (celeste_elbo_single)")
    return ap.parse_args()


def build_stripe(n_fields, n_sources, device):
    """The synthetic stripe.  CELESTE_STRIPE_CACHE=<dir> keeps a pickle of it between processes of one
    tuning session (same seeds, same arrays; nothing about the measured path changes)."""
    import pickle
    from celeste_jl_b200 import synthetic
    cache = os.environ.get("CELESTE_STRIPE_CACHE")
    path = os.path.join(cache, f"stripe_{n_fields}x{n_sources}.pkl") if cache else None
    if path and os.path.exists(path):
        with open(path, "rb") as f:
            return pickle.load(f)
    stripe = [synthetic.FieldDataset(n_sources, H=2048, W=1489, seed=42 + f, pixel_seed=1 + f, device=device)
              for f in range(n_fields)]
    if path and int(os.environ.get("RANK", "0")) == 0:
        os.makedirs(cache, exist_ok=True)
        with open(path + ".tmp", "wb") as f:
            pickle.dump(stripe, f, protocol=4)
        os.replace(path + ".tmp", path)
    return stripe


def shard_tasks(ds, rank, world):
    """ParallelRun's cost model (sum of active pixels, ParallelRun.jl:45-56): tasks sorted by cost and dealt
    greedily to the least-loaded rank (celeste_jl_b200.parallel_run.shard_sources)."""
    from celeste_jl_b200 import parallel_run
    cost = [parallel_run.estimate_time(ds.patches[s, :]) for s in range(len(ds.catalog))]
    return parallel_run.shard_sources(cost, rank, world)


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "20"], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [r.split(",") for r in open(self.path).read().strip().splitlines() if r.strip()]
            rows = [r for r in rows if len(r) >= 9]
            # "under load" = samples drawing more than half of the peak power seen in the window
            pmax = max((float(r[3]) for r in rows), default=0.0)
            loaded = [r for r in rows if float(r[3]) >= 0.5 * pmax] or rows
            sm = [float(r[1]) for r in loaded]
            out["samples"] = len(rows)
            out["samples_under_load"] = len(loaded)
            if sm:
                out["sm_mhz"] = float(np.median(sm))
                out["sm_max_mhz"] = float(rows[0][2])
                out["power_w_max"] = max(float(r[3]) for r in rows)
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for k, nm in enumerate(names):
                    if any(r[5 + k].strip().lower().startswith("active") for r in rows):
                        out["reasons"].append(nm)
        except Exception as e:   # clocks are evidence, never fatal
            out["error"] = str(e)
        finally:
            try:
                os.unlink(self.path)
            except OSError:
                pass
        return out


def load_oracle_for_timing():
    """CPU baseline library: a -march=native build of oracle/ on this box when g++ is present (written to a temp
    dir), else the portable prebuilt one.  This is the ONLY place bench.py executes oracle/."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    try:
        out = os.path.join(tempfile.mkdtemp(prefix="celeste_oracle_"), "libceleste_oracle_native.so")
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "native", f"OUT={out}"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return oracle_lib, oracle_lib.load(out), "native"
    except Exception:
        return oracle_lib, oracle_lib.load(), "portable"


def effective_cores():
    """Host cores this process may really use: the cgroup CPU quota when there is one (the GPU boxes show 128
    logical CPUs but cap the container at 16), else the affinity mask."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, int(round(int(q) / int(p)))))
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            p = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, int(round(q / p))))
        except Exception:
            pass
    return n


def cpu_leg(ds, sample, mode, steps, warmup):
    """Time the oracle (the reference's as-written algorithm, multithreaded over sources with a shared counter
    like one_node_single_infer) on `sample` tasks of field 0, with all the host cores the box grants
    (thread count = the better of 1x and 2x the core quota, picked by one trial step each)."""
    oracle_lib, lib, build = load_oracle_for_timing()
    cores = effective_cores()
    rows, act = ds.tasks(range(min(sample, len(ds.catalog))))
    from celeste_jl_b200.flatten import csr_tasks
    tasks = [(r, a, np.stack([ds.vp[i - 1] for i in r], axis=1)) for r, a in zip(rows, act)]
    csr = csr_tasks(tasks)
    of = oracle_lib.OracleField(ds.images, ds.patches, lib=lib)
    best = None
    for nt in sorted({cores, 2 * cores}):
        of.elbo_csr(*csr, mode=mode, n_threads=nt)
        t0 = time.perf_counter()
        of.elbo_csr(*csr, mode=mode, n_threads=nt)
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, nt)
    threads = best[1]
    t0 = time.perf_counter()
    for _ in range(steps):
        out = of.elbo_csr(*csr, mode=mode, n_threads=threads)
    dt = (time.perf_counter() - t0) / steps
    visits = int(out["counters"].sum())
    return {"value": len(tasks) / dt, "unit": "sources/s", "cores": cores, "threads": threads, "kind": "port",
            "sample": f"{len(tasks)} tasks of field 0 per step ({visits} pixel-visits), mode={'grad' if mode == 1 else 'hess'}, "
                      f"oracle build={build}, {steps} steps, {threads} threads on {cores} granted cores "
                      f"({os.cpu_count()} logical CPUs visible)",
            "ms_per_step": dt * 1e3, "pixel_visits_per_s": visits / dt}


def single_call_leg(field, ds, n_threads=8, n_sources=64, seconds=1.5):
    """The drop-in call as the reference's optimiser makes it (ElboMaximize.evaluate!, ElboMaximize.jl:161-172: one
    `elbo` per Newton iterate per thread): `n_threads` host threads each loop celeste_elbo_single (value + gradient +
    Hessian, host buffers in and out) over their own sources of field 0.  Reports microseconds per call and calls/s;
    the first call of every source (plan construction) is outside the timed window."""
    import ctypes as C
    from celeste_jl_b200 import _lib
    lib = _lib.load()
    rows, act = ds.tasks(range(n_sources))
    jobs = []
    for r, a in zip(rows, act):
        src = np.asarray(r, dtype=np.int32)
        ai = np.asarray(a, dtype=np.int32)
        vp = np.concatenate([ds.vp[i - 1] for i in r])
        out = (np.zeros(1), np.zeros(44), np.zeros(44 * 44), np.zeros(2, dtype=np.int64), np.zeros(1, dtype=np.int32))
        jobs.append((src, ai, vp, out))

    def call(j):
        src, ai, vp, (v, d, h, c, f) = j
        st = lib.celeste_elbo_single(field._handle, len(src), src.ctypes.data, len(ai), ai.ctypes.data, vp.ctypes.data, 2,
                                     v.ctypes.data, d.ctypes.data, h.ctypes.data, c.ctypes.data, f.ctypes.data)
        assert st == 0, st
    for j in jobs:          # builds and caches the plans; second call captures the graph
        call(j)
        call(j)
    counts = [0] * n_threads
    lat = [[] for _ in range(n_threads)]
    stop = time.perf_counter() + seconds

    def worker(k):
        mine = jobs[k::n_threads]
        i = 0
        while time.perf_counter() < stop:
            t0 = time.perf_counter()
            call(mine[i % len(mine)])
            lat[k].append(time.perf_counter() - t0)
            i += 1
        counts[k] = i
    th = [threading.Thread(target=worker, args=(k,)) for k in range(n_threads)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    allat = np.concatenate([np.asarray(x) for x in lat])
    # one thread alone: the latency a single caller sees
    solo = []
    for _ in range(200):
        t1 = time.perf_counter()
        call(jobs[0])
        solo.append(time.perf_counter() - t1)
    return {"threads": n_threads, "calls_per_s": sum(counts) / dt, "us_per_call_median": float(np.median(allat) * 1e6),
            "us_per_call_p95": float(np.percentile(allat, 95) * 1e6), "us_per_call_single_thread_median": float(np.median(solo) * 1e6),
            "mean_sources_per_task": float(np.mean([len(j[0]) for j in jobs])),
            "what": "celeste_elbo_single, mode 2 (value + gradient + dense 44 x 44 Hessian), host buffers, cached plan + "
                    "CUDA graph per (source list, mode); ctypes from Python threads (GIL released inside the call)"}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = (f"synthetic stripe: {args.fields} fields x {args.sources_per_field} sources, 5 bands, 2048x1489 px/band, "
                f"K=2 PSF, catalog patches (radius<=25), find_neighbors tasks, Sa=1")
    config = {"workload": workload, "mode": "ELBO+grad (mode 1); Hessian mode reported under 'hessian'",
              "sharding": f"tasks dealt by active-pixel cost to {world} rank(s), no collective on the data path",
              "l2": "inputs larger than L2 (no flush)", "fields": args.fields,
              "sources": args.fields * args.sources_per_field}

    if args.impl == "reference":
        # reference arm: the reference's CPU algorithm (oracle port) on the box's host cores; rank 0 only
        if rank != 0:
            return
        from celeste_jl_b200 import synthetic
        ds = synthetic.FieldDataset(args.sources_per_field, H=2048, W=1489, seed=42, pixel_seed=1, device="cpu")
        leg = cpu_leg(ds, args.cpu_sample, 1, args.steps, args.warmup)
        line = {"impl": "reference", "metric": "sources/sec (ELBO+grad)", "value": leg["value"], "unit": "sources/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": leg["ms_per_step"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config, "cpu_baseline": leg,
                "e2e": {"value": leg["value"], "unit": "sources/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py --impl cuda needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout; keep stdout for the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    import celeste_jl_b200 as cj
    from celeste_jl_b200 import _lib

    stripe = build_stripe(args.fields, args.sources_per_field, device=str(dev))
    # one multi-field plan per rank: every task of this rank's shard, all fields, three launches per step
    fields, all_rows, all_act, task_field, vp_parts = [], [], [], [], []
    for fi, ds in enumerate(stripe):
        mine = shard_tasks(ds, rank, world)
        fields.append(cj.DeviceField(ds.images, ds.patches, device=local_rank))
        rows, act = ds.tasks(mine)
        all_rows += rows
        all_act += act
        task_field += [fi] * len(rows)
        vp_parts.append(ds.vp_flat(rows))
    plans = [cj.Plan(fields, all_rows, all_act, task_field=task_field)]
    vps = [np.concatenate(vp_parts)]
    n_tasks_total = len(all_rows)
    stream = torch.cuda.current_stream()

    # device-resident buffers (value) and pinned host buffers (e2e)
    dev_bufs, host_bufs = [], []
    for plan, vp in zip(plans, vps):
        n = plan.n_tasks
        dev_bufs.append({"vp": torch.from_numpy(vp).to(dev), "v": torch.zeros(n, dtype=torch.float64, device=dev),
                         "d": torch.zeros(n * 44, dtype=torch.float64, device=dev),
                         "h": torch.zeros(n * 44 * 44, dtype=torch.float64, device=dev),
                         "c": torch.zeros(2 * n, dtype=torch.int64, device=dev),
                         "f": torch.zeros(n, dtype=torch.int32, device=dev)})
        hb = {"vp": torch.from_numpy(vp).pin_memory(), "v": torch.zeros(n, dtype=torch.float64).pin_memory(),
              "d": torch.zeros(n * 44, dtype=torch.float64).pin_memory(),
              "h": torch.zeros(n * 44 * 44, dtype=torch.float64).pin_memory(),
              "counters": torch.zeros(2 * n, dtype=torch.int64).pin_memory(),
              "flags": torch.zeros(n, dtype=torch.int32).pin_memory()}
        host_bufs.append(hb)

    def step_device(mode):
        for plan, b in zip(plans, dev_bufs):
            plan.run_device(b["vp"].data_ptr(), mode, b["v"].data_ptr(), b["d"].data_ptr(), b["h"].data_ptr(),
                            b["c"].data_ptr(), b["f"].data_ptr(), stream=stream.cuda_stream)

    # numpy views of the pinned host buffers, made once: what a host caller holds (the views cost ~10 us per step)
    host_views = [{k: t.numpy() for k, t in hb.items()} for hb in host_bufs]
    empty = np.zeros(0)

    def step_host(mode, packed=False):
        for plan, hv in zip(plans, host_views):
            hh = empty
            if mode >= 2:
                hh = hv["h"][:406 * plan.n_tasks] if packed else hv["h"]
            out = {"v": hv["v"], "d": hv["d"] if mode >= 1 else empty, "h": hh, "counters": hv["counters"], "flags": hv["flags"]}
            plan.run_host(hv["vp"], mode, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def measure(mode, steps, warmup, sample_clocks):
        sampler = ClockSampler(local_rank)
        if sample_clocks and rank == 0:
            sampler.start()
            time.sleep(0.3)
        for _ in range(warmup):
            step_device(mode)
        barrier()
        runs = []
        for rep in range(3):                      # the headline is the FIRST K-step region; two more show stability
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                step_device(mode)
            e1.record(stream)
            barrier()
            runs.append(max_over_ranks(e0.elapsed_time(e1) / steps))
        clocks = sampler.stop() if sample_clocks and rank == 0 else None      # sampled over the three timed regions
        ms = runs[0]
        # per-kernel times, measured live with CUDA events around each launch of the evaluation
        for p in plans:
            p.enable_timing(True)
        pix_ms, set_ms, epi_ms = 0.0, 0.0, 0.0
        unit = np.zeros(3)
        reps = 5
        for _ in range(reps):
            step_device(mode)
            torch.cuda.synchronize()
            for p in plans:
                a, b, c = p.kernel_times_ms()
                set_ms += a
                pix_ms += b
                epi_ms += c
                unit += np.array(p.unit_times_ms())
        for p in plans:
            p.enable_timing(False)
        counts = np.zeros(2)
        flags = 0
        for b in dev_bufs:
            counts += b["c"].cpu().numpy().reshape(-1, 2).sum(axis=0)
            flags += int(b["f"].sum().item())
        pc = parity_check(mode)
        # e2e: public host-buffer call, H2D of vp and D2H of results inside the timed region
        def time_host(packed):
            for p in plans:
                if mode == 2:
                    p.set_hessian_layout(packed)
            for _ in range(max(1, min(warmup, 2))):
                step_host(mode, packed)
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                step_host(mode, packed)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / steps
            barrier()
            for p in plans:
                if mode == 2:
                    p.set_hessian_layout(False)
            return max_over_ranks(dt)
        e2e_s = time_host(False)
        e2e_packed_s = time_host(True) if mode == 2 else None
        return {"ms": ms, "runs": runs, "pix_ms": max_over_ranks(pix_ms / reps), "setup_ms": set_ms / reps,
                "epi_ms": epi_ms / reps, "unit_ms": [max_over_ranks(x / reps) for x in unit],
                "active": sum_over_ranks(counts[0]), "inactive": sum_over_ranks(counts[1]),
                "flags": sum_over_ranks(flags), "e2e_s": e2e_s, "e2e_packed_s": e2e_packed_s, "clocks": clocks,
                "parity_check": pc}

    def parity_check(mode, n_sample=48):
        """A seeded sample of this rank's tasks of the benched plan through the CPU oracle (the checker; outside every
        timed region): relative errors of value / gradient / Hessian and equality of the pixel-visit counters."""
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib
        plan, b = plans[0], dev_bufs[0]
        step_device(mode)
        torch.cuda.synchronize()
        n = plan.n_tasks
        pick = np.sort(np.random.default_rng(1234 + mode + 17 * rank).choice(n, min(n_sample, n), replace=False))
        v = b["v"].cpu().numpy()[pick]
        d = b["d"].cpu().numpy().reshape(n, 44)[pick] if mode >= 1 else None
        h = b["h"].cpu().numpy().reshape(n, 44 * 44)[pick] if mode >= 2 else None
        cnt = b["c"].cpu().numpy().reshape(n, 2)[pick]
        vp_all = vps[0].reshape(-1, 44)
        slot0 = np.concatenate([[0], np.cumsum([len(r) for r in all_rows])])
        oracles = {}
        rel_v = rel_d = rel_h = rel_h3 = 0.0
        counters_equal = True
        for k, t in enumerate(pick):
            fi = task_field[t]
            if fi not in oracles:
                oracles[fi] = oracle_lib.OracleField(stripe[fi].images, stripe[fi].patches)
            vpm = vp_all[slot0[t]:slot0[t + 1]].T
            ref = oracles[fi].elbo_batch([(all_rows[t], all_act[t], vpm)], mode=mode, n_threads=1)
            rel_v = max(rel_v, abs(ref["v"][0] - v[k]) / abs(ref["v"][0]))
            counters_equal = counters_equal and bool(np.array_equal(ref["counters"][0], cnt[k]))
            if mode >= 1:
                sc = np.abs(ref["d"]).max()
                rel_d = max(rel_d, float((np.abs(ref["d"] - d[k]) / np.maximum(np.abs(ref["d"]), sc * 1e-6)).max()))
            if mode >= 2:
                sc = np.abs(ref["h"]).max()
                rel_h = max(rel_h, float((np.abs(ref["h"] - h[k]) / np.maximum(np.abs(ref["h"]), sc * 1e-6)).max()))
                rel_h3 = max(rel_h3, float((np.abs(ref["h"] - h[k]) / np.maximum(np.abs(ref["h"]), sc * 1e-3)).max()))
        out = {"n": int(len(pick) * world), "max_rel_v": max_over_ranks(rel_v), "counters_equal": bool(min_over_ranks(counters_equal)),
               "tolerance": 1e-8, "what": "seeded sample of each rank's tasks vs the CPU oracle; gradient / Hessian "
               "component-wise |delta| / max(|ref_ij|, 1e-6 * max|ref|)"}
        if mode >= 1:
            out["max_rel_d"] = max_over_ranks(rel_d)
        if mode >= 2:
            out["max_rel_h"] = max_over_ranks(rel_h)
            # the same with a floor of 1e-3 of the largest entry: the 1e-6 floor is reached by entries a million times
            # smaller than the matrix scale, whose absolute error (< 1 ulp of the largest entry) is rounding noise
            out["max_rel_h_floor_1e-3"] = max_over_ranks(rel_h3)
        out["ok"] = bool(out["counters_equal"] and max(out["max_rel_v"], out.get("max_rel_d", 0), out.get("max_rel_h", 0)) <= 1e-8)
        return out

    def min_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return float(t.item())

    peak = C_peak = None
    import ctypes as C
    pk = C.c_double(0.0)
    _lib.check(_lib.load().celeste_fp64_peak(C.byref(pk), stream.cuda_stream))
    peak = pk.value

    total_sources = sum_over_ranks(n_tasks_total)
    n_slots = sum(p.n_src for p in plans)
    grad = measure(1, args.steps, args.warmup, True)
    hess = None if args.no_hessian else measure(2, max(3, args.steps // 2), args.warmup, False)

    def roofline(m, mode):
        """Roofline of the DOMINANT kernel of the step (FP64-pipe bound: ~10^2 flop per byte).
        * value / gradient: `frac` = SURVEY 8(d)'s ALGORITHMIC flop of the pixel-visits the kernel serves / its live
          time / the measured DFMA peak (the contract figure), and next to it `frac_executed` = what the kernel really
          executes (SASS 2 x DFMA + DMUL + DADD thread instructions of the ncu capture of the same workload) / live time
          / peak, and the FP64 pipe activity ncu saw.
        * Hessian: the contract count (23.7 kflop per visit) prices the reference's unfactorised 44 x 44 chain rule, which
          this library does not execute, so its ratio to the peak is NOT a fraction of peak: it is
          reported as `contract_ratio`, and `frac` is the executed fraction."""
        family = plans[0].kernel_name(mode)
        if family == "unit_kernel":
            dom, dom_ms = f"unit_walk_kernel<{mode}>", m["unit_ms"][1]
            parts = {"unit_bg_kernel": m["unit_ms"][0], dom: dom_ms}
            if mode == 2:
                parts["unit_moment_kernel"] = m["unit_ms"][2]
        else:
            dom, dom_ms = f"{family}<{mode}>", m["pix_ms"]
            parts = {dom: dom_ms}
        pk = peak * world
        contract_flop = m["active"] * F_ACTIVE[mode] + (0.0 if family == "unit_kernel" else m["inactive"] * F_INACTIVE)
        contract = contract_flop / (dom_ms * 1e-3) / 1e12
        prof, cap = {}, os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(cap):
            try:
                prof = json.load(open(cap))
            except Exception:
                prof = {}
        kernels = {}
        for name, ms_k in parts.items():
            t = prof.get(name) or next((v for k, v in prof.items() if k.startswith(name + "<")), None)
            ent = {"ms_per_step": ms_k}
            if t and t.get("sources") == int(total_sources) and ms_k > 0:
                ex = t["executed_flop_per_launch"] / (ms_k * 1e-3) / 1e12       # whole-job: the capture is the N = 1 launch
                ent.update({"executed_tflops": ex, "frac_executed": ex / pk,
                            "fp64_pipe_active_pct_ncu": t.get("fp64_pipe_active_pct"),
                            "dram_bytes_per_launch": t["dram_bytes_per_launch"] / world})
            kernels[name] = ent
        d = kernels[dom]
        r = {"bound": "fp64", "unit": "TFLOP/s", "peak": pk, "peak_per_gpu": peak, "kernel": dom,
             "kernel_ms_per_step": dom_ms, "kernel_share_of_step": dom_ms / m["ms"],
             "pixel_kernels_ms_per_step": m["pix_ms"], "pixel_kernels_share_of_step": m["pix_ms"] / m["ms"],
             "frac_executed": d.get("frac_executed"), "fp64_pipe_active_pct": d.get("fp64_pipe_active_pct_ncu"),
             "traffic": d.get("dram_bytes_per_launch"), "kernels": kernels,
             "algorithmic_flop_per_step": contract_flop, "pixel_visits_active": m["active"],
             "pixel_visits_inactive": m["inactive"],
             "peak_source": "measured live by this library (celeste_fp64_peak: register-resident DFMA chains); no FP64 "
                            "entry exists in MEASURED_PEAKS.json; nominal 148 SM x 64 DFMA/clk x 1.965 GHz = 37.2 TFLOP/s",
             "executed_source": "profiles/ncu_traffic.json (ncu --set full of tools/profile_step.py on this workload)",
             "hbm": {"achieved_gbs": (m["active"] + m["inactive"]) * BYTES_PER_VISIT / (m["pix_ms"] * 1e-3) / 1e9,
                     "peak_gbs": hbm_peak()}}
        if mode <= 1:
            r.update({"achieved": contract, "frac": contract / pk,
                      "frac_kind": "algorithmic: SURVEY 8(d) contract flop (2.6 kflop per active pixel-visit) / kernel time / peak"})
        else:
            ex = d.get("executed_tflops")
            r.update({"achieved": ex, "frac": d.get("frac_executed"),
                      "frac_kind": "executed: SASS 2 x DFMA + DMUL + DADD of the ncu capture / live kernel time / peak",
                      "contract_ratio": contract / pk,
                      "contract_ratio_note": "SURVEY 8(d)'s 23.7 kflop per visit prices the reference's unfactorised chain rule; "
                                             "not a fraction of peak"})
        return r

    # configs[4]: the full maximize! loop (ELBO + gradient + Hessian + KL, Newton trust region, 50 iterations max)
    # for every source of this rank's shard at once (ParallelRun.one_node_single_infer semantics: the target
    # starts from generic_init_source, its neighbours stay at catalog_init_source), in lock-step on the GPU
    maximize_leg = None
    if not args.no_maximize:
        from celeste_jl_b200 import deterministic_vi as dvi
        from celeste_jl_b200.elbo_maximize import BatchMaximizer
        vps0 = []
        for fi, r in zip(task_field, all_rows):
            cat = stripe[fi].catalog
            vps0.append(dvi.generic_init_source(cat[r[0] - 1].pos))
            vps0 += [dvi.catalog_init_source(cat[k - 1]) for k in r[1:]]
        vps0 = np.concatenate(vps0)
        BatchMaximizer(plans[0], vps0, include_kl=True, max_iters=2).run()          # warm-up (allocations, KL tables)
        barrier()
        t0 = time.perf_counter()
        res = BatchMaximizer(plans[0], vps0, include_kl=True).run()
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        ntot = sum_over_ranks(len(res.value))
        # catalog check (outside the timed region): what the optimiser recovered against the synthetic truth, on the
        # bright isolated sources -- the acceptance test of the reference (test/test_optimization.jl:10-32)
        from celeste_jl_b200.model import ids
        n_chk = n_type = 0
        dpos, dflux = [], []
        for k, (fi, r) in enumerate(zip(task_field, all_rows)):
            ce = stripe[fi].catalog[r[0] - 1]
            flux_r = (ce.star_fluxes if ce.is_star else ce.gal_fluxes)[2]
            if len(r) > 1 or flux_r < 30.0 or not res.converged[k]:
                continue
            vs = res.vp[k]
            a_true = 0 if ce.is_star else 1
            n_chk += 1
            n_type += int(vs[ids.is_star[a_true]] >= 0.5)
            dpos.append(float(np.abs(vs[:2] - ce.pos).max()))
            dflux.append(float(abs(math.exp(vs[ids.flux_loc[a_true]] + 0.5 * vs[ids.flux_scale[a_true]]) / flux_r - 1.0)))
        n_chk_t, n_type_t = sum_over_ranks(n_chk), sum_over_ranks(n_type)
        catalog_check = {"sources_checked": int(n_chk_t), "type_correct_fraction": n_type_t / max(n_chk_t, 1),
                         "position_error_px_max": max_over_ranks(max(dpos) if dpos else 0.0),
                         "brightness_rel_error_median": float(np.median(dflux)) if dflux else None,
                         "brightness_rel_error_max": max_over_ranks(max(dflux) if dflux else 0.0),
                         "what": "isolated sources with r-band flux >= 30 nMgy whose optimisation converged, against the "
                                 "synthetic truth (type, position, r-band brightness); rank 0's median, max over ranks"}
        maximize_leg = {"sources": int(ntot), "seconds": dt, "sources_per_s": ntot / dt,
                        "lockstep_iterations": int(max_over_ranks(res.total_steps)),
                        "mean_newton_iterations": sum_over_ranks(float(res.iterations.sum())) / ntot,
                        "converged_fraction": sum_over_ranks(float(res.converged.sum())) / ntot,
                        "catalog_check": catalog_check,
                        "what": "one_node_single_infer semantics on the whole stripe: generic init, KL included, Newton "
                                "trust region x_tol 1e-7 / f_tol 1e-6 / g_tol 1e-8 / 50 iterations, converged sources masked"}

        if args.maximize_timeline and rank == 0:
            bm = BatchMaximizer(plans[0], vps0, include_kl=True)
            tl, ev0, ev1 = [], torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            orig_eval, orig_step = bm._evaluate_plan, bm._step

            def eval_wrap():
                torch.cuda.synchronize()
                a = int(bm.mask.sum().item())
                ev0.record()
                orig_eval()
                ev1.record()
                torch.cuda.synchronize()
                tl.append([a, round(ev0.elapsed_time(ev1), 4), None])

            def step_wrap(phase):
                ev0.record()
                orig_step(phase)
                ev1.record()
                torch.cuda.synchronize()
                if tl and phase in (0, 1):
                    tl[-1][2] = round(ev0.elapsed_time(ev1), 4)

            bm._evaluate_plan, bm._step = eval_wrap, step_wrap
            bm.run()
            maximize_leg["timeline"] = tl

    # row f.4: the value-only full-image render (fill_celeste_expectation!) of field 0, through the C ABI with host
    # buffers (122 MB of float64 expectation images come back per call); rank 0 only, informational
    render_leg = None
    if rank == 0 and not args.no_render:
        ds0 = stripe[0]
        vp0 = np.stack(ds0.vp, axis=1)
        rows0 = np.arange(1, vp0.shape[1] + 1)
        fields[0].render_expectation(rows0[:8], vp0[:, :8])                          # warm-up
        t0 = time.perf_counter()
        imgs = fields[0].render_expectation(rows0, vp0)
        dt = time.perf_counter() - t0
        npix = sum(int(a.size) for a in imgs)
        render_leg = {"seconds_e2e": dt, "sources": int(vp0.shape[1]), "image_pixels": npix,
                      "megapixels_per_s_e2e": npix / dt / 1e6,
                      "what": "celeste_render_expectation: E_G - sky on every pixel of the 5 images of field 0, host buffers"}

    single_leg = None
    if rank == 0 and world == 1 and not args.no_single:
        single_leg = single_call_leg(fields[0], stripe[0])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    def e2e(m, mode, packed=False):
        h2d = n_slots * 44 * 8
        hb = (406 * 8 if packed else 44 * 44 * 8) if mode >= 2 else 0
        d2h = n_tasks_total * (8 + 16 + 4 + (44 * 8 if mode >= 1 else 0) + hb)
        sec = m["e2e_packed_s"] if packed else m["e2e_s"]
        out = {"value": total_sources / sec, "unit": "sources/s", "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": d2h * world, "ms_per_step": sec * 1e3}
        if mode >= 2:
            out["hessian_layout"] = ("packed28: 406 doubles per source, the upper triangle of the live 28 x 28 block "
                                     "(celeste_plan_set_hessian_layout)") if packed else "dense 44 x 44 per source"
        return out

    def stability(m):
        return {"ms_per_step_runs": m["runs"], "min": float(np.min(m["runs"])), "median": float(np.median(m["runs"]))}

    line = {"metric": "sources/sec (ELBO+grad)", "value": total_sources / (grad["ms"] * 1e-3), "unit": "sources/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": grad["ms"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config, "clocks": grad["clocks"], "e2e": e2e(grad, 1),
            "gpu_launches": args.steps * sum(p.launches(1) for p in plans),
            "roofline": roofline(grad, 1), "nonfinite_tasks": grad["flags"], "parity_check": grad["parity_check"],
            "stability": stability(grad),
            "pixel_visits_per_s": (grad["active"] + grad["inactive"]) / (grad["ms"] * 1e-3)}
    if hess is not None:
        line["hessian"] = {"value": total_sources / (hess["ms"] * 1e-3), "unit": "sources/s", "ms_per_step": hess["ms"],
                           "e2e": e2e(hess, 2, packed=True), "e2e_dense": e2e(hess, 2), "roofline": roofline(hess, 2),
                           "parity_check": hess["parity_check"], "stability": stability(hess),
                           "nonfinite_tasks": hess["flags"]}
    if not args.no_maximize:
        line["maximize"] = maximize_leg
    if render_leg is not None:
        line["render"] = render_leg
    if single_leg is not None:
        line["single_call"] = single_leg
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_leg(stripe[0], args.cpu_sample, 1, 3, 1)
        line["cpu_baseline"] = cb
        if hess is not None:
            line["hessian"]["cpu_baseline"] = cpu_leg(stripe[0], args.cpu_sample, 2, 2, 1)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6650.0   # fallback stated in /opt/skills/guides/B200_PROFILING.md


if __name__ == "__main__":
    main()
