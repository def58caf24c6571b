#!/usr/bin/env python
"""bench.py -- preprocess_cds PCA throughput (cells/s) on B200, with roofline and CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun ... bench.py --gpus N ...           (one rank per GPU, N > 1)

A STEP is one full pass of the hot path over one synthetic matrix:
  normalize (K2) -> gene select/filter + dual layout build (K0) -> gene moments (K3) ->
  IRLBA (K4-K9, num_dim=100) -> embedding.
`value`  : device-resident cells/s (raw CSC already in HBM when the clock starts).
`e2e`    : the same through the public API (new_cell_data_set + preprocess_cds) from pinned
           HOST buffers, H2D/D2H inside the timed region.
Workload : BASELINE.json configs[1] -- synthetic C. elegans L2 shape, 42k cells x 20k genes,
           ~8% dense, num_dim=100, per GPU (weak scaling: N GPUs hold 42k*N cells of ONE matrix,
           sharded by cells, one PCA over all of them).
Timing   : CUDA events on the library's own stream (m3b_timer_*), barrier + synchronize on both
           sides, max over ranks.  Inputs (0.8 GB per layout) are far larger than L2 (126 MB).
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFG2 = dict(name="cfg2: synthetic C. elegans L2 shape", n_genes=20000, cells_per_gpu=42000, density=0.08,
            num_dim=100, n_types=150, seed=1)
CPU_SAMPLE_CELLS = 3000


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells-per-gpu", type=int, default=CFG2["cells_per_gpu"])
    ap.add_argument("--genes", type=int, default=CFG2["n_genes"])
    ap.add_argument("--density", type=float, default=CFG2["density"])
    ap.add_argument("--num-dim", type=int, default=CFG2["num_dim"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-cells", type=int, default=CPU_SAMPLE_CELLS)
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end arm (used for short ncu launch lists)")
    ap.add_argument("--no-roofline-pass", action="store_true", help="skip the instrumented step (short ncu launch lists)")
    return ap.parse_args()


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region through NVML (pynvml) in a
    background thread.  (Polling the nvidia-smi binary instead perturbs the run: it re-enumerates
    the devices on every sample and stalled this launch-heavy step by ~40%.)"""

    def __init__(self, gpu_index, period=0.25):
        self.idx, self.period = gpu_index, period
        self.samples, self.stop_flag, self.thread, self.err = [], threading.Event(), None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
        except Exception as e:  # pragma: no cover
            self.err = "nvml unavailable: %s" % e
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.samples.append((sm, mx, rs, pw))
            except Exception as e:  # pragma: no cover
                self.err = str(e)
                break
            self.stop_flag.wait(self.period)

    def stop(self):
        self.stop_flag.set()
        if self.thread:
            self.thread.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err or "no samples"]}
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        reasons = sorted(k for k, bit in names.items() if any(s[2] & bit for s in self.samples))
        return {"sm_mhz": float(np.median([s[0] for s in self.samples])),
                "sm_max_mhz": float(max(s[1] for s in self.samples)), "reasons": reasons,
                "samples": len(self.samples), "power_w_max": max(s[3] for s in self.samples), "source": "nvml"}


def pinned(a):
    """Page-locked copy of a NumPy array (torch is only the allocator here)."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a))
    try:
        t = t.pin_memory()
    except Exception:
        pass
    return t.numpy(), t


def cpu_oracle_run(counts, num_dim):
    """The restated reference algorithm (oracle/) on the host cores: full preprocess_cds."""
    from oracle import monocle3_ref as ref
    t0 = time.perf_counter()
    sf = ref.estimate_size_factors(counts)
    r = ref.preprocess_cds(counts, sf, num_dim=num_dim)
    dt = time.perf_counter() - t0
    return dt, r


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  R / irlba are not
    installable here (no R in the image), so this is the oracle port (oracle/): SciPy sparse
    products (single-threaded, like CHOLMOD sdmult inside irlba) + NumPy/OpenBLAS dense ops
    on all host threads.  Each step is a full preprocess_cds on a bounded sample of the
    workload's cells."""
    if rank != 0:
        return
    from monocle3_b200.synth import SynthSpec
    import scipy.sparse as sp
    spec = SynthSpec(args.genes, args.cells_per_gpu * args.gpus, args.density, CFG2["n_types"], CFG2["seed"], device="cpu")
    n = min(args.cpu_sample_cells, spec.C)
    indptr, indices, data = spec.cells(0, n)
    counts = sp.csc_matrix((data, indices, indptr), shape=(args.genes, n))
    nd = min(args.num_dim, n - 1)
    for _ in range(args.warmup):
        cpu_oracle_run(counts, nd)
    t = []
    for _ in range(args.steps):
        dt, r = cpu_oracle_run(counts, nd)
        t.append(dt)
    ms = 1e3 * float(np.mean(t))
    val = n / (ms / 1e3)
    cores = os.cpu_count()
    sample = "full preprocess_cds(num_dim=%d) on the first %d cells x %d genes of the workload (nnz %d, mprod %d)" % (
        nd, n, args.genes, counts.nnz, r["mprod"])
    print(json.dumps({
        "impl": "reference", "metric": "cells/sec preprocess_cds PCA (num_dim=100)", "value": val, "unit": "cells/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": val, "unit": "cells/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, world):
    return {"workload": "%s, %d cells/GPU x %d genes, ~%.0f%% dense, preprocess_cds(num_dim=%d, norm_method='log', "
                        "scaling=TRUE)" % (CFG2["name"], args.cells_per_gpu, args.genes, 100 * args.density, args.num_dim),
            "cells_total": args.cells_per_gpu * world, "genes": args.genes, "density": args.density,
            "num_dim": args.num_dim, "sharding": "cells x%d" % world,
            "l2": "inputs larger than L2 (no flush needed)", "small_svd": "device"}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import torch
    import torch.distributed as td
    import scipy.sparse as sp
    import monocle3_b200 as m3
    from monocle3_b200 import engine, dist as m3dist, _lib as L
    from monocle3_b200.synth import SynthSpec
    from monocle3_b200.r_rng import rnorm_after_seed

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        td.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch with torchrun for N>1)" % (args.gpus, world))

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        td.all_reduce(t, op=td.ReduceOp.SUM)
        return float(t.item())

    ctx = engine.Context(device=local_rank, small_svd="device")
    m3.set_context(ctx)
    if world > 1:
        m3dist.init_comm(ctx)

    # ---- synthetic shard (host, pinned) ----
    C_total = args.cells_per_gpu * world
    spec = SynthSpec(args.genes, C_total, args.density, CFG2["n_types"], CFG2["seed"])
    lo, hi = rank * args.cells_per_gpu, (rank + 1) * args.cells_per_gpu
    indptr, indices, data = spec.cells(lo, hi)
    (indptr_p, k1), (indices_p, k2), (data_p, k3) = pinned(indptr.astype(np.int32)), pinned(indices), pinned(data)
    counts = sp.csc_matrix((data_p, indices_p, indptr_p), shape=(args.genes, hi - lo), copy=False)
    nnz_local = int(counts.nnz)
    shard = m3dist.make_shard(C_total) if world > 1 else None
    del spec
    torch.cuda.empty_cache()

    # ---- device-resident arm: counts + size factors already in HBM ----
    cds = m3.new_cell_data_set(counts, shard=shard)
    dev = cds.device_matrix()
    sf = m3.size_factors(cds)

    wall = {}

    def device_step():
        t0 = time.perf_counter()
        dev.normalize(sf, L.NORM_LOG2, 1.0)
        t1 = time.perf_counter()
        dev.select_genes(None)
        t2 = time.perf_counter()
        v0 = rnorm_after_seed(2016, dev.n_kept)
        nv = min(args.num_dim, min(dev.n_kept, C_total) - 1)
        t3 = time.perf_counter()
        r = dev.pca_irlba(nv, v0, center=True, scale=True, want_x=False)
        t4 = time.perf_counter()
        wall.update(normalize_ms=1e3 * (t1 - t0), select_genes_ms=1e3 * (t2 - t1), rnorm_ms=1e3 * (t3 - t2),
                    pca_irlba_call_ms=1e3 * (t4 - t3))
        return r

    for _ in range(args.warmup):
        res = device_step()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ctx.reset_stats()
    ctx.timer_start()
    step_walls = []
    for _ in range(args.steps):
        res = device_step()
        step_walls.append(dict(wall))
    ms_total = ctx.timer_stop()
    barrier()
    clocks = sampler.stop()
    if os.environ.get("M3B_BENCH_DEBUG"):
        print("DEBUG rank %d cpus=%s affinity=%d steps=%s host_ms=%s" % (
            rank, os.cpu_count(), len(os.sched_getaffinity(0)),
            [{k: round(v, 1) for k, v in w.items()} for w in step_walls],
            [round(v, 1) for v in ctx.stats()["host_ms"]]), file=sys.stderr, flush=True)
    st = ctx.stats()
    ms_step = max_over_ranks(ms_total) / args.steps
    launches = int(sum_over_ranks(st["kernel_launches"]))
    value = C_total / (ms_step / 1e3)

    roofline = None
    if not args.no_roofline_pass:
        # ---- roofline of the dominant kernel (k_flat_dot, the two SpMVs): one instrumented step ----
        ctx.set_flags(L.FLAG_TIME_SPMV)
        ctx.reset_stats()
        res_i = device_step()
        sti = ctx.stats()
        ctx.set_flags(0)
        nA, nB = sti["spmv_cells_out"], sti["spmv_genes_out"]
        bytes_total = nA * sti["spmv_bytes_cells_out"] + nB * sti["spmv_bytes_genes_out"]
        spmv_ms = sti["spmv_cells_out_ms"] + sti["spmv_genes_out_ms"]
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = bytes_total / (spmv_ms / 1e3) / 1e9 if spmv_ms > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "spmv_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        avg_us = 1e3 * spmv_ms / max(1, nA + nB)
        roofline = {"bound": "hbm", "kernel": "k_flat_dot (K4 cells-out + K5 genes-out SpMV)",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0,
                    "launches_timed": int(nA + nB), "avg_launch_us": 1e3 * spmv_ms / max(1, nA + nB),
                    "algorithmic_bytes_per_launch": bytes_total / max(1, nA + nB),
                    "cells_out": {"gbs": nA * sti["spmv_bytes_cells_out"] / max(sti["spmv_cells_out_ms"], 1e-9) / 1e6,
                                  "avg_us": 1e3 * sti["spmv_cells_out_ms"] / max(1, nA)},
                    "genes_out": {"gbs": nB * sti["spmv_bytes_genes_out"] / max(sti["spmv_genes_out_ms"], 1e-9) / 1e6,
                                  "avg_us": 1e3 * sti["spmv_genes_out_ms"] / max(1, nB)},
                    "spmv_share_of_irlba": spmv_ms / max(sti["irlba_ms"], 1e-9),
                    "stream_note": "achieved = ALGORITHMIC bytes (12 B/nnz of an int32-indexed CSR, SURVEY 8d) / time; the "
                                   "kernel streams less than that: 16-bit tile-local indices, and count-1 entries "
                                   "(index only, 2 B) carry no value, so frac can exceed 1; dram_gbs is the measured "
                                   "DRAM traffic of the ncu capture (traffic) over the same launch time",
                    "dram_gbs": (traffic / (avg_us * 1e-6) / 1e9) if (traffic and avg_us > 0) else None,
                    "dram_frac_of_peak": (traffic / (avg_us * 1e-6) / 1e9 / peak) if (traffic and avg_us > 0) else None,
                    "irlba_phase_ms": dict(zip(["spmv_cells_out", "spmv_genes_out", "orthog", "small_svd", "restart_gemm",
                                                "vector_epilogues", "allreduce"], [round(v, 3) for v in sti["phase_ms"][:7]])),
                    "irlba_ms": sti["irlba_ms"], "svd_sweeps": sti["svd_sweeps"], "svd_fallbacks": sti["svd_fallbacks"],
                    "host_wall_ms": {k: round(v, 2) for k, v in wall.items()},
                    "irlba_host_ms": dict(zip(["setup", "loop", "final", "release"], [round(v, 2) for v in sti["host_ms"]])),
                    "irlba_host_ms_untimed_step": dict(zip(["setup", "loop", "final", "release"], [round(v, 2) for v in st["host_ms"]]))}

        # ---- the other HBM-bound passes of the path (K1 totals, K2 normalize), event-timed, 5 reps ----
        def timed(fn, reps=5):
            fn()
            ctx.timer_start()
            for _ in range(reps):
                fn()
            return ctx.timer_stop() / reps
        timed(lambda: dev.cell_totals(True))
        k1_ms = ctx.stats()["k1_ms"]
        timed(lambda: dev.normalize(sf, L.NORM_LOG2, 1.0))
        k2_ms = ctx.stats()["k2_ms"]
        n_loc = hi - lo
        aux = {"K1_cell_totals": {"kernel_ms": k1_ms, "gbs": (8 * nnz_local + 16 * n_loc) / max(k1_ms, 1e-9) / 1e6,
                                  "algorithmic_bytes": 8 * nnz_local + 16 * n_loc,
                                  "frac_of_measured_peak": (8 * nnz_local + 16 * n_loc) / max(k1_ms, 1e-9) / 1e6 / peak},
               "K2_normalize_log2": {"kernel_ms": k2_ms, "gbs": (16 * nnz_local + 8 * n_loc) / max(k2_ms, 1e-9) / 1e6,
                                     "algorithmic_bytes": 16 * nnz_local + 8 * n_loc,
                                     "frac_of_measured_peak": (16 * nnz_local + 8 * n_loc) / max(k2_ms, 1e-9) / 1e6 / peak}}
        roofline["other_hbm_passes"] = aux

    e2e = None
    if not args.no_e2e:
        # ---- end-to-end arm: public API from pinned host buffers ----
        def e2e_step():
            c = m3.new_cell_data_set(counts, shard=shard)           # H2D of the CSC + size factors
            c = m3.preprocess_cds(c, num_dim=args.num_dim)          # ... embedding + loadings back on the host
            x = c.reduced_dims["PCA"]
            c._dev.free()
            return c, x

        dev.free()
        cds._dev = None
        for _ in range(max(1, min(args.warmup, 3))):
            c_e, x_e = e2e_step()
        barrier()
        ctx.reset_stats()
        ctx.timer_start()
        for _ in range(args.steps):
            c_e, x_e = e2e_step()
        ms_e2e_total = ctx.timer_stop()
        barrier()
        ste = ctx.stats()
        ms_e2e = max_over_ranks(ms_e2e_total) / args.steps
        e2e = {"value": C_total / (ms_e2e / 1e3), "unit": "cells/s",
               "h2d_bytes_per_step": int(sum_over_ranks(ste["h2d_bytes"]) / args.steps),
               "d2h_bytes_per_step": int(sum_over_ranks(ste["d2h_bytes"]) / args.steps), "ms_per_step": ms_e2e}

    # ---- CPU baseline (rank 0, N=1 only): oracle on a bounded sample ----
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n = min(args.cpu_sample_cells, hi - lo)
        sub = sp.csc_matrix((data[:indptr[n]], indices[:indptr[n]], indptr[:n + 1]), shape=(args.genes, n))
        nd = min(args.num_dim, n - 1)
        dt, r = cpu_oracle_run(sub, nd)
        cpu_baseline = {"value": n / dt, "unit": "cells/s", "cores": os.cpu_count(), "kind": "port",
                        "sample": "restated CPU path (irlba algorithm, SciPy SpMV + NumPy/OpenBLAS): full "
                                  "preprocess_cds(num_dim=%d) on the first %d cells x %d genes (nnz %d, mprod %d, %.1f s)"
                                  % (nd, n, args.genes, sub.nnz, r["mprod"], dt)}
    nnz_total = int(sum_over_ranks(nnz_local))
    if rank == 0:
        out = {"metric": "cells/sec preprocess_cds PCA (num_dim=100)", "value": value, "unit": "cells/s",
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": workload_config(args, world), "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
               "roofline": roofline, "cpu_baseline": cpu_baseline,
               "irlba": {"iter": res["iter"], "mprod": res["mprod"], "d1": float(res["d"][0]),
                         "nnz_total": nnz_total, "status": res["status"]}}
        print(json.dumps(out))
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
