#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json metric: queries/sec).

Workload (BASELINE.json configs[1]): GpuIndexFlatL2, N=10M, d=128, nq=10k, k=100, synthetic fp32.
A "step" = one search() of all nq queries over the whole database.

  python bench.py --gpus 1 --steps K --warmup W            # this framework (tcgen05 Flat path)
  python bench.py --impl reference --gpus 1 --steps K ...   # reference CPU IndexFlatL2 (oracle/_ref)
  torchrun --nproc-per-node N bench.py --gpus N ...         # database sharded over N GPUs
                                                            # (IndexShards semantics, NCCL all-gather merge)

One JSON line on stdout (rank 0).  `value` = QPS with inputs resident in HBM; `e2e` = QPS through
the public API with host (pinned) buffers, H2D/D2H inside the timed region; `roofline` = algorithmic
FLOPs of the step / device time inside the tcgen05 kernel, vs the measured bf16 GEMM peak.
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

N_TOTAL = int(os.environ.get("BENCH_N", 10_000_000))
DIM = int(os.environ.get("BENCH_D", 128))
NQ = int(os.environ.get("BENCH_NQ", 10_000))
K = int(os.environ.get("BENCH_K", 100))
CHUNK = 1_000_000  # database is generated in seeded chunks so shards do not depend on world size


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return j, "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load" = samples in the upper half of the power range
        thr = 0.5 * (max(power) + min(power))
        load = [s for s, p in zip(sm, power) if p >= thr] or sm
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": float(max(power))}


def gen_rows(torch, device, r0, r1, d):
    """rows [r0, r1) of the synthetic database: uniform [0,1) fp32, chunk c seeded with 1234 + c"""
    out = torch.empty((r1 - r0, d), dtype=torch.float32, device=device)
    c = r0 // CHUNK
    while c * CHUNK < r1:
        g = torch.Generator(device=device)
        g.manual_seed(1234 + c)
        c0, c1 = c * CHUNK, min((c + 1) * CHUNK, N_TOTAL)
        chunk = torch.rand((c1 - c0, d), dtype=torch.float32, device=device, generator=g)
        a, b = max(r0, c0), min(r1, c1)
        out[a - r0 : b - r0] = chunk[a - c0 : b - c0]
        del chunk
        c += 1
    return out


def gen_queries(torch, device, nq, d):
    g = torch.Generator(device=device)
    g.manual_seed(1235)
    return torch.rand((nq, d), dtype=torch.float32, device=device, generator=g)


def host_threads():
    """CPUs this process may actually use: affinity mask, capped by the cgroup CPU quota."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = max(1, min(n, int(float(q) / float(per) + 0.5)))
    except Exception:
        pass
    return n


def cpu_reference_qps(xb_host, xq_host, k, budget_s, steps=1, warmup=0):
    """Times the reference CPU IndexFlatL2 (oracle/_ref; else the numpy port) on a bounded sample of
    the same workload.  Returns (qps, info dict)."""
    from oracle import ref

    cores = host_threads()
    if ref.available():
        ref.set_omp_threads(cores)
        ref.set_blas_threads(cores)
        # Bound the work: the reference's brute-force search is linear in the number of database rows,
        # so when 1000 queries over all N rows would blow the time budget (it does on a 16-CPU cgroup
        # quota with this image's pthreads OpenBLAS), time a leading slice of the rows and scale.
        nfull = xb_host.shape[0]
        probe_rows = min(nfull, 250_000)
        pidx = ref.IndexFlat(xb_host.shape[1], 1)
        pidx.add(xb_host[:probe_rows])
        nsp = int(min(xq_host.shape[0], max(1000, 128000 // xb_host.shape[1] + 1)))
        t0 = time.time()
        pidx.search(xq_host[:nsp], k)
        t_probe = time.time() - t0
        del pidx
        per_step_budget = max(10.0, budget_s / max(2, steps + warmup))
        frac = 1.0
        while frac > 1.0 / 64 and t_probe * (nfull * frac / probe_rows) > per_step_budget:
            frac /= 2
        nrows = int(nfull * frac)
        xb_host = xb_host[:nrows]
        idx = ref.IndexFlat(xb_host.shape[1], 1)
        t0 = time.time()
        idx.add(xb_host)
        t_add = time.time() - t0
        # Sample = the first 1000 queries: the reference switches from its per-query SIMD loop to the
        # BLAS-blocked path at nq*d >= 128000 (faiss/utils/distances.cpp:600), i.e. nq >= 1000 at
        # d=128 -- anything smaller would time a different (much slower) code path than nq=10k uses.
        ns = int(min(xq_host.shape[0], max(1000, 128000 // xb_host.shape[1] + 1)))
        t_begin = time.time()
        ts = []
        n_runs = 0
        for i in range(max(0, warmup) + max(1, steps)):
            t0 = time.time()
            idx.search(xq_host[:ns], k)
            dt = time.time() - t0
            n_runs += 1
            if i >= warmup or dt * 2 > budget_s:
                ts.append(dt)
            # bounded: stop when the next run would exceed the budget (at least one timed run)
            if ts and (time.time() - t_begin) + dt > budget_s:
                break
        t_meas = float(np.mean(ts))
        t = t_meas * (nfull / nrows)  # exhaustive search: time linear in the rows scanned
        return ns / t, {"kind": "reference", "cores": cores, "sample": "IndexFlatL2 (oracle/_ref, %s, OpenBLAS pthreads, %d threads) first %d of %d queries, k=%d, against the first %d of N=%d rows: %.2f s measured/step over %d timed step(s) (%d run), scaled x%.0f to full N (exhaustive search is linear in N) = %.2f s/step; add %.1f s" % (
            ref.compile_options().strip(), cores, ns, xq_host.shape[0], k, nrows, nfull, t_meas, len(ts), n_runs, nfull / nrows, t, t_add), "ms_per_step": t * 1e3, "nq_sample": ns}
    from oracle import oracle_np as o

    nb = min(xb_host.shape[0], 200_000)
    ns = min(xq_host.shape[0], 64)
    t0 = time.time()
    o.knn_flat(xq_host[:ns], xb_host[:nb], k, 1)
    t = time.time() - t0
    # scale to the full database size (exhaustive search is linear in N)
    qps = ns / (t * xb_host.shape[0] / nb)
    return qps, {"kind": "port", "cores": 1, "sample": "numpy oracle port, %d queries x %d rows, scaled linearly to N=%d" % (ns, nb, xb_host.shape[0]),
                 "ms_per_step": t * 1e3, "nq_sample": ns}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    steps, warmup = max(1, args.steps), max(0, args.warmup)

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))

    import torch

    config = {"workload": "GpuIndexFlatL2 N=%d d=%d nq=%d k=%d (BASELINE configs[1])" % (N_TOTAL, DIM, NQ, K),
              "N": N_TOTAL, "d": DIM, "nq": NQ, "k": K,
              "parallelism": "IndexShards x%d (contiguous row shards, all-gather top-k merge)" % world if world > 1 else "single GPU",
              "l2_note": "inputs larger than L2 (database %.1f GB fp32 + %.1f GB fp16 copy per step vs 126 MB L2)" % (
                  N_TOTAL * DIM * 4 / 1e9 / world, N_TOTAL * DIM * 2 / 1e9 / world)}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        dev = "cuda:0" if torch.cuda.is_available() else "cpu"
        xb = gen_rows(torch, dev, 0, N_TOTAL, DIM).cpu().numpy()
        xq = gen_queries(torch, dev, NQ, DIM).cpu().numpy()
        if dev != "cpu":
            torch.cuda.empty_cache()
        qps, info = cpu_reference_qps(xb, xq, K, budget_s=150.0, steps=steps, warmup=warmup)
        out = {"impl": "reference", "metric": "queries/sec (Flat-L2 exact k-NN)", "value": qps, "unit": "queries/s",
               "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": info["ms_per_step"],
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": config,
               "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]},
               "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "gpu_launches": 0}
        print(json.dumps(out), flush=True)
        return

    # ------------------------------------------------------------------ this framework
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback exists in faiss_b200)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        # keep stdout to the one JSON line: NCCL prints its version banner to stdout when the first
        # communicator is created (any NCCL_DEBUG level >= VERSION), so that happens with fd 1 -> fd 2
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            dist.all_reduce(torch.zeros(1, device=device))
            torch.cuda.synchronize()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
    import faiss_b200 as fb
    from faiss_b200.distributed import ShardedSearcher, shard_bounds

    res = fb.StandardGpuResources()
    # order the library's work on torch's current stream so torch CUDA events bracket it
    stream = torch.cuda.current_stream(device)
    res.setDefaultStream(local_rank, stream.cuda_stream)

    r0, r1 = shard_bounds(N_TOTAL, rank, world)
    t0 = time.time()
    xb = gen_rows(torch, device, r0, r1, DIM)
    index = fb.GpuIndexFlatL2(res, DIM, device=local_rank)
    index.add(xb)
    xq = gen_queries(torch, device, NQ, DIM)
    xq_pin = torch.empty((NQ, DIM), dtype=torch.float32, pin_memory=True)
    xq_pin.copy_(xq)
    D_pin = torch.empty((NQ, K), dtype=torch.float32, pin_memory=True)
    I_pin = torch.empty((NQ, K), dtype=torch.int64, pin_memory=True)
    torch.cuda.synchronize()
    log("[rank %d] shard rows [%d,%d) built in %.1f s" % (rank, r0, r1, time.time() - t0))

    searcher = None
    if world > 1:
        searcher = ShardedSearcher(lambda q, k: index.search(q, k), r1 - r0, fb.METRIC_L2, res=res, device=local_rank)

    def step_device():
        if searcher is not None:
            return searcher.search(xq, K)
        return index.search(xq, K)

    def step_e2e():
        if searcher is not None:
            q = xq_pin.to(device, non_blocking=True)
            D, I = searcher.search(q, K)
            D_pin.copy_(D, non_blocking=True)
            I_pin.copy_(I, non_blocking=True)
            torch.cuda.synchronize()
        else:
            index.search(xq_pin.numpy(), K, D=D_pin.numpy(), I=I_pin.numpy())  # H2D + search + D2H inside

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(warmup, 3)):
        step_device()
    barrier()

    # ---- timed region: device-resident inputs
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    fb.lib.faiss_b200_kernel_timing(1)
    l0 = fb.lib.faiss_b200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(steps):
        D, I = step_device()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1) / steps
    launches = fb.lib.faiss_b200_launch_count() - l0
    import ctypes

    tc_ms = ctypes.c_double()
    tc_n = ctypes.c_int()
    fb.lib.faiss_b200_kernel_timing_collect(b"flat_tc", ctypes.byref(tc_ms), ctypes.byref(tc_n))
    fb.lib.faiss_b200_kernel_timing(0)
    clocks = sampler.stop() if rank == 0 else None
    info = index.lastSearchInfo()

    # ---- e2e: host buffers through the public API
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.time()
    for _ in range(steps):
        step_e2e()
    barrier()
    e2e_ms = (time.time() - t0) * 1e3 / steps

    if dist is not None:
        t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        pk, pk_src = peaks()
        flops_step = 2.0 * NQ * (r1 - r0) * DIM  # this rank's shard
        tc_ms_step = tc_ms.value / steps if tc_n.value else None
        peak_tf = float(pk.get("bf16_tflops_sustained", pk.get("bf16_tflops")))
        roof = {"bound": "tensor", "unit": "TFLOP/s", "peak": peak_tf, "traffic": None,
                "peak_source": "%s bf16 GEMM peak (sustained; kernel timed inside a multi-step loop), MEASURED_PEAKS.json" % pk_src,
                "kernel": "flat_tc_kernel (tcgen05 fp16 scoring + fused top-k filter), %d launches/step" % (tc_n.value // steps if tc_n.value else 0),
                "algorithmic_flops_per_step": flops_step}
        # DRAM bytes per launch of this kernel, from the committed `ncu --set full` capture of this same
        # workload (profiles/flat_tc_traffic.json, written by scripts/ncu_traffic.py); single GPU only
        tpath = os.path.join(ROOT, "profiles", "flat_tc_traffic.json")
        if world == 1 and N_TOTAL == 10_000_000 and os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                roof["traffic"] = tj["dram_bytes_per_launch"]
                roof["traffic_source"] = tj["source"]
            except Exception:
                pass
        if tc_ms_step:
            roof["achieved"] = flops_step / (tc_ms_step * 1e-3) / 1e12
            roof["frac"] = roof["achieved"] / peak_tf
            roof["kernel_ms_per_step"] = tc_ms_step
            roof["kernel_share_of_step"] = tc_ms_step / ms
        else:
            roof["achieved"] = None
            roof["frac"] = None
        out = {"metric": "queries/sec (Flat-L2 exact k-NN, recall@k = 1.0 by construction)", "value": NQ / (ms * 1e-3), "unit": "queries/s",
               "n_gpus": world, "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms, "higher_is_better": True,
               "scaling": "strong", "vs_baseline": None, "dtype": "f16 mma (fp32 accumulate) + f32 exact re-rank", "data": "synthetic",
               "config": config, "clocks": clocks,
               "e2e": {"value": NQ / (e2e_ms * 1e-3), "unit": "queries/s", "ms_per_step": e2e_ms,
                       "h2d_bytes_per_step": NQ * DIM * 4, "d2h_bytes_per_step": NQ * K * 12},
               "gpu_launches": int(launches), "roofline": roof,
               "search_info": info}
        # ---- CPU baseline (reference CPU path on this box's host cores), N=1 only
        if world == 1 and not args.no_cpu_baseline:
            try:
                xb_host = xb.cpu().numpy()
                qps, cinfo = cpu_reference_qps(xb_host, xq_pin.numpy(), K, budget_s=20.0, steps=1, warmup=0)
                out["cpu_baseline"] = {"value": qps, "unit": "queries/s", "cores": cinfo["cores"], "kind": cinfo["kind"], "sample": cinfo["sample"]}
                # parity spot check on the sample the CPU just answered
            except Exception as e:
                out["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": host_threads(), "kind": "reference", "sample": "failed: %s" % str(e)[:200]}
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
