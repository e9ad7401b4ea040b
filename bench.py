#!/usr/bin/env python3
"""Throughput harness for the scDD hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched with torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

Metric (BASELINE.json): gene-permutation fits per second.  One gene-permutation fit = one (gene, permutation)
unit of permMclustGene (R/permutations.R:273-282) = 2 restricted mixture fit-sets + 2 PPM scores.

Workload (config.workload): BASELINE config 3 -- simulateSet-style 20,000 genes x 2,000 cells (1,000 per
condition), permutations drawn on the device with R's sampler.  The gene matrix is sharded over the N ranks
(config 4; no collective on the data path: genes are independent).  One STEP = `perms_per_step` x N
permutations of every resident gene, i.e. 20,000 x perms_per_step gene-permutations per rank per step
(weak scaling: per-GPU work per step is constant).  A full scDD(permutations = 1000) job is 1000 / (perms per
step) such steps; the rate is what is reported.

  value  : inputs resident in HBM (scdd_plan_*), CUDA events on the launching stream, max over ranks
  e2e    : the same step through the reference-facing C-ABI call scdd_perm_bf() with HOST (pinned) buffers:
           H2D of the gene block + seeds and D2H of the Bayes-factor numerators inside the timed region
  roofline: fp64 (non-tensor) pipe -- ALGORITHMIC flops 24 T + 4 P (T, P counted by the kernel: SURVEY 8d)
            over the fit kernel's CUDA-event duration, against the DFMA peak measured in this run
  cpu_baseline: the CPU oracle (C restatement of the reference's R + mclust path, OpenMP) on a bounded
            sample of the same genes, all host cores
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

METRIC = "gene-permutation fits/sec"
UNIT = "gene-permutations/s"
PRIORS = dict(alpha=0.01, mu0=0.0, s0=0.01, a0=0.01, b0=0.01)     # R/scDD.R:203 (example / vignette priors)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--genes", type=int, default=20000)
    ap.add_argument("--cells", type=int, default=2000)
    ap.add_argument("--perms-per-step", type=int, default=16)
    ap.add_argument("--seed", type=int, default=284)                # simulateSet's default random.seed
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-balance", action="store_true", help="N > 1: keep the interleaved gene shards")
    ap.add_argument("--cpu-sample-genes", type=int, default=512)
    ap.add_argument("--cpu-sample-perms", type=int, default=8)
    return ap.parse_args()


def workload(args):
    from scdd_b200.simulate import simulate_set_style
    X, cond = simulate_set_style(args.genes, args.cells // 2, seed=args.seed)
    return X, cond


def shard_rows(ngenes, rank, world):
    """interleaved gene shard: rank r takes genes r, r+world, ... (categories are laid out in blocks, so
    interleaving gives every rank the same category mix and the same expected cost)"""
    return np.arange(rank, ngenes, world)


def balance_rows(costs, world):
    """Longest-processing-time-first assignment of genes to ranks from measured per-gene cost: returns a list of
    `world` sorted row-index arrays that partition range(len(costs)).  Deterministic (stable sort, ties by index)."""
    costs = np.asarray(costs, dtype=np.float64)
    order = np.argsort(-costs, kind="stable")
    load = np.zeros(world)
    owner = np.empty(costs.size, dtype=np.int64)
    for g in order:
        r = int(np.argmin(load))
        owner[g] = r
        load[r] += costs[g]
    return [np.nonzero(owner == r)[0] for r in range(world)]


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons while the timed region runs (ONE long-lived nvidia-smi
    process in loop mode: spawning a new one per sample perturbs the GPU it is measuring)"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, period_ms=500):
        self.idx = gpu_index
        self.period_ms = period_ms
        self.rows = []
        self._p = None
        self._t = None

    def _read(self):
        try:
            for line in self._p.stdout:
                line = line.strip()
                if line:
                    self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def start(self):
        try:
            self._p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                        "--format=csv,noheader,nounits", "-lms", str(self.period_ms)],
                                       stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self._t = threading.Thread(target=self._read, daemon=True)
            self._t.start()
        except Exception:
            self._p = None

    def stop(self):
        if self._p is not None:
            try:
                self._p.terminate()           # the exact process we started
                self._p.wait(timeout=5)
            except Exception:
                pass
            if self._t:
                self._t.join(timeout=5)
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) > 8:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        pw = [float(r[3]) for r in self.rows if len(r) > 8 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows), "power_w_max": max(pw) if pw else None}


def cpu_baseline(args, X, cond, nthreads=0):
    """the oracle on a bounded, category-stratified sample of the same workload"""
    from oracle import pyoracle as po
    from scdd_b200 import hotpath as hp
    ng = min(args.cpu_sample_genes, X.shape[0])
    rows = np.linspace(0, X.shape[0] - 1, ng).astype(int)
    logy, c01, off = hp.gene_csr(X[rows], cond)
    seeds = hp.gene_seeds(args.seed, ng)
    cores = os.cpu_count() or 1
    B = args.cpu_sample_perms
    t0 = time.time()
    po.perm_batch(logy, c01, off, B, seed6=seeds, cfg=po.make_cfg(**{**PRIORS, "mu0": PRIORS["mu0"]}), nthreads=nthreads)
    dt = time.time() - t0
    return {"value": ng * B / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{ng} genes (evenly spaced over the {X.shape[0]}-gene workload) x {B} permutations, "
                      f"{dt:.1f} s wall, OpenMP over genes"}


def run_reference(args):
    """--impl reference: the CPU restatement of the reference path (oracle) on all host threads.
    The reference itself is pure R + mclust and cannot run in this image (no R); see DESIGN.md."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as po
    from scdd_b200 import hotpath as hp
    X, cond = workload(args)
    ng = min(512, X.shape[0])        # enough genes to keep every host thread busy (OpenMP over genes)
    rows = np.linspace(0, X.shape[0] - 1, ng).astype(int)
    logy, c01, off = hp.gene_csr(X[rows], cond)
    seeds = hp.gene_seeds(args.seed, ng)
    B = 2
    cfg = po.make_cfg(**PRIORS)
    cores = os.cpu_count() or 1
    for _ in range(args.warmup):
        po.perm_batch(logy, c01, off, B, seed6=seeds, cfg=cfg, nthreads=0)
    t0 = time.time()
    for _ in range(args.steps):
        po.perm_batch(logy, c01, off, B, seed6=seeds, cfg=cfg, nthreads=0)
    dt = time.time() - t0
    val = ng * B * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"simulateSet-style {args.genes} genes x {args.cells} cells, permutations=1000 "
                                   f"(BASELINE config 3); each step = a bounded sample of it",
                       "sample_per_step": f"{ng} evenly spaced genes x {B} permutations"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{ng} genes x {B} permutations per step, OpenMP over genes, all host threads; "
                                       "C restatement of the reference's R+mclust path (R is not installable here)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from scdd_b200 import hotpath as hp
    from scdd_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    X, cond = workload(args)
    rows = shard_rows(X.shape[0], rank, world)
    Xr = X[rows]
    logy, c01, off = hp.gene_csr(Xr, cond)
    ngr = rows.size
    seeds = hp.gene_seeds(args.seed, X.shape[0])[rows]
    nperm = args.perms_per_step * world
    cfg = hp.make_cfg(**PRIORS)
    hp.set_devices([local])

    import ctypes
    clk = ctypes.c_double()
    fp64_peak = _lib.lib().scdd_measure_fp64_peak(local, ctypes.byref(clk))

    # ---------------- resident arm: `value`
    plan = hp.Plan(local, logy, c01, off, nperm, seeds, cfg=cfg)
    stream = torch.cuda.current_stream().cuda_stream
    shard_note = "interleaved" if world > 1 else "all genes on the one GPU"
    if world > 1 and not args.no_balance:
        # Untimed set-up: EM work per gene is a property of the gene, so one measured chunk on the interleaved shards
        # lets all ranks re-partition the genes longest-first by MEASURED cost (the library orders its own work the
        # same way inside a device).  Every rank holds the whole synthetic matrix, so re-sharding moves no data; the
        # only exchange is this all-gather of 8 bytes per gene, outside the timed region.
        plan.run_chunk(stream, nperm)
        mine = torch.zeros(X.shape[0], dtype=torch.float64, device="cuda")
        mine[torch.from_numpy(rows).cuda()] = torch.from_numpy(plan.gene_costs()).cuda()
        dist.all_reduce(mine, op=dist.ReduceOp.SUM)
        plan.close()
        rows = balance_rows(mine.cpu().numpy(), world)[rank]
        Xr = X[rows]
        logy, c01, off = hp.gene_csr(Xr, cond)
        ngr = rows.size
        seeds = hp.gene_seeds(args.seed, X.shape[0])[rows]
        plan = hp.Plan(local, logy, c01, off, nperm, seeds, cfg=cfg)
        shard_note = "balanced by measured per-gene EM work (one untimed chunk, all-gather of 8 B/gene)"
    for _ in range(args.warmup):
        plan.run_chunk(stream, nperm)
    barrier()
    st0 = plan.stats()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        plan.run_chunk(stream, nperm)
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    st1 = plan.stats()
    d = {k: st1[k] - st0[k] for k in st1}
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    units = torch.tensor([float(ngr) * nperm * args.steps], dtype=torch.float64, device="cuda")
    work = torch.tensor([d["triples"], d["point_iters"], d["fit_kernel_ms"], d["fit_kernel_launches"], d["rng_kernel_ms"],
                         d["fit_sets"], d["tol_near"], d["bic_near"], d["failed_sets"], d["iter_cap_hits"]],
                        dtype=torch.float64, device="cuda")
    per_rank = torch.zeros(world, 3, dtype=torch.float64, device="cuda")     # diagnostics: step ms, fit ms, SM MHz per rank
    per_rank[rank, 0] = ms / args.steps
    per_rank[rank, 1] = d["fit_kernel_ms"] / max(1.0, d["fit_kernel_launches"])
    per_rank[rank, 2] = clocks["sm_mhz"] or 0.0
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(units, op=dist.ReduceOp.SUM)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
        dist.all_reduce(per_rank, op=dist.ReduceOp.SUM)
    ms_max = float(t.item())
    total_units = float(units.item())
    value = total_units / (ms_max * 1e-3)
    plan.close()

    # ---------------- e2e arm: host buffers through scdd_perm_bf (H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        pin = lambda a: torch.from_numpy(a).pin_memory().numpy()   # noqa: E731
        h_logy, h_c01, h_off, h_seeds = pin(logy), pin(c01), pin(off), pin(np.ascontiguousarray(seeds))
        hp.perm_bf(h_logy, h_c01, h_off, nperm, seed6=h_seeds, cfg=cfg)          # warm-up
        n_e2e = max(1, min(args.steps, 3))
        barrier()
        t0 = time.perf_counter()
        h2d = d2h = 0.0
        for _ in range(n_e2e):
            out, st = hp.perm_bf(h_logy, h_c01, h_off, nperm, seed6=h_seeds, cfg=cfg)
            h2d += st["h2d_bytes"]
            d2h += st["d2h_bytes"]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        uu = torch.tensor([float(ngr) * nperm * n_e2e], dtype=torch.float64, device="cuda")
        bb = torch.tensor([h2d / n_e2e, d2h / n_e2e], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(uu, op=dist.ReduceOp.SUM)
            dist.all_reduce(bb, op=dist.ReduceOp.SUM)
        e2e = {"value": float(uu.item()) / float(tt.item()), "unit": UNIT,
               "h2d_bytes_per_step": float(bb[0].item()), "d2h_bytes_per_step": float(bb[1].item()),
               "steps": n_e2e, "timer": "host wall clock around scdd_perm_bf (includes H2D, kernels, D2H), max over ranks",
               "checksum": float(np.nansum(out))}

    if rank == 0:
        w = work.tolist()
        T, P, fit_ms, fit_launches, rng_ms = w[0], w[1], w[2], w[3], w[4]
        # per-launch averages (summed over ranks, so divide flops and time consistently)
        alg_flops = 24.0 * T + 4.0 * P
        per_gpu_achieved = alg_flops / (fit_ms * 1e-3) / 1e12 if fit_ms > 0 else None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        # algorithmic bytes of the fit kernel per step: every (gene, permutation) gathers its n_g values (8 B)
        # through the partition index (4 B) and the permutation index (2 B), and writes two 8 B scores
        nnz_total = float(off[-1]) * world
        alg_bytes = nnz_total * nperm * 14.0 + (total_units / args.steps) * 16.0
        traffic = None
        prof = os.path.join(ROOT, "profiles", "fit_kernel_latest.json")
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"simulateSet-style {args.genes} genes x {args.cells} cells ({args.cells // 2}/condition), "
                                   f"permutations=1000 (BASELINE config 3{'/4, gene-sharded' if world > 1 else ''})",
                       "step": f"{nperm} device-drawn permutations of every resident gene "
                               f"({args.genes} x {args.perms_per_step} gene-permutations per GPU per step)",
                       "genes_per_gpu": int(ngr), "parallelism": f"gene-shard x{world}, no data-path collective",
                       "shards": shard_note,
                       "l2": "per-step inputs (gene values + permutation indices, >0.6 GB) exceed the 126 MB L2",
                       "priors": PRIORS},
            "clocks": clocks,
            "per_rank": {"ms_per_step": [round(v, 2) for v in per_rank[:, 0].tolist()],
                         "fit_kernel_ms": [round(v, 2) for v in per_rank[:, 1].tolist()],
                         "sm_mhz": per_rank[:, 2].tolist()},
            "e2e": e2e,
            "gpu_launches": int(3 * args.steps * world),
            "roofline": {"bound": "fp64", "achieved": per_gpu_achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": (per_gpu_achieved / fp64_peak) if per_gpu_achieved and fp64_peak > 0 else None,
                         "traffic": traffic,
                         "kernel": "scdd::fit_kernel", "launches": fit_launches,
                         "avg_launch_ms": fit_ms / fit_launches if fit_launches else None,
                         "kernel_share_of_step": fit_ms / (ms_max * world) if ms_max > 0 else None,
                         "flops_definition": "algorithmic 24*T + 4*P per launch (T = point x component x iteration "
                                             "triples, P = point x iteration, both counted by the kernel)",
                         "peak_source": "DFMA microbenchmark in this run (MEASURED_PEAKS.json has no fp64 entry)",
                         "hbm": {"achieved_gbs": alg_bytes / (ms_max / args.steps * 1e-3) / 1e9 / world, "peak_gbs": hbm_peak,
                                 "frac": alg_bytes / (ms_max / args.steps * 1e-3) / 1e9 / world / hbm_peak,
                                 "note": "algorithmic bytes per GPU (gathered values + partition and u16 permutation indices "
                                         "+ scores) / step time; the path is fp64-bound, HBM is <1% utilised"}},
            "work": {"fit_sets": w[5], "triples": T, "point_iters": P, "tol_near": w[6], "bic_near": w[7],
                     "failed_sets": w[8], "iter_cap_hits": w[9], "rng_kernel_ms": rng_ms, "fit_kernel_ms": fit_ms},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, X, cond)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
