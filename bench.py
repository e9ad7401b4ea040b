#!/usr/bin/env python3
"""Throughput harness for the scDD hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched with torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W
    python bench.py --config 5 [--genes G] ...               BASELINE config 5 (50,000 cells, adjust.perms, CDR)
    python bench.py --full [--full-perms 1000] [--gpus N]    one whole scDD(permutations = 1000) call, wall time

Metric (BASELINE.json): gene-permutation fits per second.  One gene-permutation fit = one (gene, permutation)
unit of permMclustGene (R/permutations.R:273-282) = 2 restricted mixture fit-sets + 2 PPM scores (3 + 3 on the
adjusted path of config 5, R/permutations.R:243-267).

Workload (config.workload): BASELINE config 3 -- simulateSet-style 20,000 genes x 2,000 cells (1,000 per
condition), permutations drawn on the device with R's sampler.  The gene matrix is sharded over the N ranks
(config 4; no collective on the data path: genes are independent).  One STEP = `perms_per_step` x N
permutations of every resident gene, i.e. 20,000 x perms_per_step gene-permutations per rank per step
(weak scaling: per-GPU work per step is constant).  A full scDD(permutations = 1000) job is 1000 / (perms per
step) such steps; the rate is what is reported.

  value   : inputs resident in HBM (scdd_plan_*), CUDA events on the launching stream, max over ranks
  e2e     : the same step through the reference-facing C-ABI call scdd_perm_bf() with HOST (pinned) buffers:
            H2D of the gene block + seeds and D2H of the Bayes-factor numerators inside the timed region
  e2e_api : (N > 1, rank 0 alone after the ranks have finished) ONE process, scdd_set_devices(0..N-1), ONE
            scdd_perm_bf call on the whole gene block with pageable host arrays -- the path the R shim uses
  roofline: fp64 (non-tensor) pipe -- ALGORITHMIC flops 24 T + 4 P (T, P counted by the kernel: SURVEY 8d)
            over the fit kernel's CUDA-event duration, against the DFMA peak measured in this run
  cpu_baseline: the CPU oracle (C restatement of the reference's R + mclust path, OpenMP) on a bounded
            sample of the same genes, all host cores
"""
import argparse
import importlib.util
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
    ap.add_argument("--config", type=int, default=3, choices=[3, 5])
    ap.add_argument("--genes", type=int, default=20000)
    ap.add_argument("--cells", type=int, default=None, help="default: 2,000 (config 3) / 50,000 (config 5)")
    ap.add_argument("--perms-per-step", type=int, default=None, help="default: 16 (config 3) / 2 (config 5)")
    ap.add_argument("--seed", type=int, default=284)                # simulateSet's default random.seed
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-api-leg", action="store_true", help="N > 1: skip rank 0's single-process scdd_set_devices leg")
    ap.add_argument("--no-balance", action="store_true", help="N > 1: keep the interleaved gene shards")
    ap.add_argument("--cpu-sample-genes", type=int, default=None)
    ap.add_argument("--cpu-sample-perms", type=int, default=None)
    ap.add_argument("--full", action="store_true", help="run one whole scDD() call through the host mirror and print its wall time")
    ap.add_argument("--full-perms", type=int, default=1000)
    a = ap.parse_args()
    if a.cells is None:
        a.cells = 2000 if a.config == 3 else 50000
    if a.perms_per_step is None:
        a.perms_per_step = 16 if a.config == 3 else 2
    return a


def _simulate_module():
    """scdd_b200/simulate.py loaded by file path: the reference arm must not import the scdd_b200 package"""
    spec = importlib.util.spec_from_file_location("_scdd_simulate", os.path.join(ROOT, "scdd_b200", "simulate.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def workload_name(args, world=1):
    if args.config == 5:
        return (f"simulateSet-style {args.genes} genes x {args.cells} cells ({args.cells // 2}/condition), adjust.perms with the "
                f"CDR covariate (permMclustCov), permutations=1000 (BASELINE config 5{', gene-sharded' if world > 1 else ''})")
    return (f"simulateSet-style {args.genes} genes x {args.cells} cells ({args.cells // 2}/condition), "
            f"permutations=1000 (BASELINE config 3{'/4, gene-sharded' if world > 1 else ''})")


def workload(args):
    """config 3: the whole matrix (320 MB) on every rank"""
    X, cond = _simulate_module().simulate_set_style(args.genes, args.cells // 2, seed=args.seed)
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


def host_threads():
    """threads this process may really use (torchrun exports OMP_NUM_THREADS=1; the affinity mask is what counts)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def csr_numpy(X, cond):
    """`y <- log(y[y>0])`, `cond0 <- condition[y>0]` per gene (R/scDD.R:374-375) in plain numpy -- used by the reference arm,
    which must not touch the product library"""
    nz = X > 0
    off = np.zeros(X.shape[0] + 1, dtype=np.int64)
    np.cumsum(nz.sum(axis=1), out=off[1:])
    logy = np.log(X[nz])
    c01 = np.broadcast_to((np.asarray(cond) != cond[0]).astype(np.int32), X.shape)[nz]
    return logy, np.ascontiguousarray(c01), off


def oracle_seeds(po, seed, n):
    """L'Ecuyer-CMRG state of bplapply element i for RNGseed = seed, by the oracle's own RNG"""
    s = po.RRng(seed, kind="lecuyer").lecuyer_state()
    out = np.zeros((n, 6), dtype=np.uint32)
    for i in range(n):
        out[i] = s
        s = po.lecuyer_next_substream(s)
    return out


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


# ---------------------------------------------------------------------------------------------- CPU arms (oracle)
def _cpu_sample(args, po, ng_default, B_default):
    """a bounded, category-stratified sample of the workload in the oracle's own layout (numpy only)"""
    ng = min(args.cpu_sample_genes or ng_default, args.genes)
    B = args.cpu_sample_perms or B_default
    rows = np.linspace(0, args.genes - 1, ng).astype(int)
    sim = _simulate_module()
    if args.config == 5:
        X, cond = sim.simulate_gene_block(rows, args.genes, args.cells // 2, seed=args.seed)
    else:
        Xall, cond = workload(args)
        X = Xall[rows]
    logy, c01, off = csr_numpy(X, cond)
    cdr = None
    if args.config == 5:
        C = (X > 0).sum(axis=0) / float(X.shape[0])        # detection rate over the sampled genes (R/scDD.R:457)
        cdr = np.broadcast_to(C, X.shape)[X > 0]
    seeds = oracle_seeds(po, args.seed, ng)
    return logy, c01, off, cdr, seeds, ng, B


def cpu_baseline(args):
    """the oracle on a bounded sample of the same workload, all usable host threads"""
    from oracle import pyoracle as po
    logy, c01, off, cdr, seeds, ng, B = _cpu_sample(args, po, 512 if args.config == 3 else 16, 8 if args.config == 3 else 1)
    nt = host_threads()
    t0 = time.time()
    po.perm_batch(logy, c01, off, B, seed6=seeds, cfg=po.make_cfg(**PRIORS), nthreads=nt, cdr=cdr)
    dt = time.time() - t0
    return {"value": ng * B / dt, "unit": UNIT, "cores": nt, "kind": "port",
            "sample": f"{ng} genes (evenly spaced over the {args.genes}-gene workload) x {B} permutations, "
                      f"{dt:.1f} s wall, OpenMP over genes with {nt} threads"}


def run_reference(args):
    """--impl reference: the CPU restatement of the reference path (oracle) on all host threads.
    The reference itself is pure R + mclust and cannot run in this image (no R); see DESIGN.md.
    Nothing of the product (scdd_b200 package, libscdd_b200.so) is imported here."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as po
    logy, c01, off, cdr, seeds, ng, B = _cpu_sample(args, po, 512 if args.config == 3 else 16, 2 if args.config == 3 else 1)
    cfg = po.make_cfg(**PRIORS)
    nt = host_threads()      # explicit: torch.distributed.run exports OMP_NUM_THREADS=1, which num_threads() overrides
    for _ in range(args.warmup):
        po.perm_batch(logy, c01, off, B, seed6=seeds, cfg=cfg, nthreads=nt, cdr=cdr)
    t0 = time.time()
    for _ in range(args.steps):
        po.perm_batch(logy, c01, off, B, seed6=seeds, cfg=cfg, nthreads=nt, cdr=cdr)
    dt = time.time() - t0
    val = ng * B * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args) + "; each step = a bounded sample of it",
                       "sample_per_step": f"{ng} evenly spaced genes x {B} permutations"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": nt, "kind": "port",
                             "threads_used": nt, "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS"),
                             "sample": f"{ng} genes x {B} permutations per step, OpenMP over genes with {nt} threads "
                                       f"(num_threads clause; the process may use {nt} of {os.cpu_count()} CPUs); "
                                       "C restatement of the reference's R+mclust path (R is not installable here)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- full scDD() call
def run_full(args):
    """one whole scDD(permutations = B) call on config 3 through the host mirror of the R interface (scdd_b200.api):
    filters, observed fits, B permutations per gene on --gpus devices (scdd_set_devices), p-values, feDP, Zhat"""
    from scdd_b200 import api, hotpath as hp
    X, cond = workload(args)
    ndev = min(args.gpus, hp.device_count())
    sce = api.SingleCellExperiment({"normcounts": X}, {"condition": cond})
    pri = dict(alpha=PRIORS["alpha"], mu0=PRIORS["mu0"], s0=PRIORS["s0"], a0=PRIORS["a0"], b0=PRIORS["b0"])
    # warm-up call (context creation, cudaMalloc of the cached buffers) on a few permutations
    api.scDD(api.SingleCellExperiment({"normcounts": X[:2000]}, {"condition": cond}), prior_param=pri, permutations=16,
             testZeroes=False, param={"RNGseed": args.seed, "devices": list(range(ndev))}, verbose=False)
    t0 = time.perf_counter()
    out = api.scDD(sce, prior_param=pri, permutations=args.full_perms, testZeroes=False,
                   param={"RNGseed": args.seed, "devices": list(range(ndev))}, verbose=False)
    wall = time.perf_counter() - t0
    # testZeroes (the configs switch it off, scDD()'s own default is on): its wall time for the same matrix, separately
    t0 = time.perf_counter()
    pz = api.testZeroes(X, cond)
    wall_z = time.perf_counter() - t0
    meta = out.metadata["_scdd_b200"]
    st = meta["stats"]["permutations"]
    G = api.results(out, "Genes")
    p = np.asarray(G["nonzero.pvalue"])
    line = {"full_scdd": {"workload": workload_name(args), "permutations": args.full_perms, "n_gpus": ndev,
                          "wall_s": wall, "gene_permutations_per_s": float(meta["tofit"].size) * args.full_perms / wall,
                          "perm_call_device_ms": st["total_ms"], "fit_kernel_ms": st["fit_kernel_ms"],
                          "rng_kernel_ms": st["rng_kernel_ms"], "observed_ms": meta["stats"]["observed"]["fit_kernel_ms"],
                          "fedp_ms": meta["stats"].get("feDP", {}).get("total_ms"),
                          "host_phases_s": {k: round(v, 4) for k, v in meta["stats"].get("host_phases_s", {}).items()},
                          "genes_fitted": int(meta["tofit"].size), "const_sets": st["const_sets"], "failed_sets": st["failed_sets"],
                          "n_significant": int((np.asarray(G["nonzero.pvalue.adj"]) < 0.05).sum()),
                          "pvalue_checksum": float(np.nansum(p)),
                          "test_zeroes_wall_s": wall_z, "test_zeroes_genes": int((~np.isnan(pz)).sum()),
                          "categories": {str(k): int(v) for k, v in zip(*np.unique([str(c) for c in G["DDcategory"]], return_counts=True))}}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- GPU arm
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    if args.full:
        return run_full(args)

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

    nperm = args.perms_per_step * world
    cfg = hp.make_cfg(**PRIORS)
    hp.set_devices([local])
    sim = _simulate_module()
    X = None
    rows = shard_rows(args.genes, rank, world)
    if args.config == 5:
        # every rank draws only its shard (per-gene RNG streams); the per-cell detection rate needs the nonzero counts
        # of ALL genes: one all-reduce of ncells integers during set-up
        Xr, cond = sim.simulate_gene_block(rows, args.genes, args.cells // 2, seed=args.seed)
        colnnz = torch.from_numpy((Xr > 0).sum(axis=0).astype(np.float64)).cuda()
        if world > 1:
            dist.all_reduce(colnnz, op=dist.ReduceOp.SUM)
        C = colnnz.cpu().numpy() / float(args.genes)                     # R/scDD.R:457
        cdr = np.broadcast_to(C, Xr.shape)[Xr > 0]
    else:
        X, cond = workload(args)
        Xr = X[rows]
        cdr = None
    logy, c01, off = hp.gene_csr(Xr, cond)
    ngr = rows.size
    seeds_all = hp.gene_seeds(args.seed, args.genes)
    seeds = seeds_all[rows]

    import ctypes
    clk = ctypes.c_double()
    fp64_peak = _lib.lib().scdd_measure_fp64_peak(local, ctypes.byref(clk))

    # ---------------- resident arm: `value`
    plan = hp.Plan(local, logy, c01, off, nperm, seeds, cfg=cfg, cdr=cdr)
    stream = torch.cuda.current_stream().cuda_stream
    shard_note = "interleaved" if world > 1 else "all genes on the one GPU"
    if world > 1 and not args.no_balance and args.config == 3:
        # Untimed set-up: EM work per gene is a property of the gene, so one measured chunk on the interleaved shards
        # lets all ranks re-partition the genes longest-first by MEASURED cost (the library orders its own work the
        # same way inside a device).  Every rank holds the whole synthetic matrix, so re-sharding moves no data; the
        # only exchange is this all-gather of 8 bytes per gene, outside the timed region.
        plan.run_chunk(stream, nperm)
        mine = torch.zeros(args.genes, dtype=torch.float64, device="cuda")
        mine[torch.from_numpy(rows).cuda()] = torch.from_numpy(plan.gene_costs()).cuda()
        dist.all_reduce(mine, op=dist.ReduceOp.SUM)
        plan.close()
        rows = balance_rows(mine.cpu().numpy(), world)[rank]
        Xr = X[rows]
        logy, c01, off = hp.gene_csr(Xr, cond)
        ngr = rows.size
        seeds = seeds_all[rows]
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
                         d["fit_sets"], d["tol_near"], d["bic_near"], d["failed_sets"], d["iter_cap_hits"], d["const_sets"]],
                        dtype=torch.float64, device="cuda")
    per_rank = torch.zeros(world, 3, dtype=torch.float64, device="cuda")     # diagnostics: step ms, fit ms, SM MHz per rank
    per_rank[rank, 0] = ms / args.steps
    per_rank[rank, 1] = d["fit_kernel_ms"] / max(1.0, d["fit_kernel_launches"])
    per_rank[rank, 2] = clocks["sm_mhz"] or 0.0
    nnz_t = torch.tensor([float(off[-1])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(units, op=dist.ReduceOp.SUM)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
        dist.all_reduce(per_rank, op=dist.ReduceOp.SUM)
        dist.all_reduce(nnz_t, op=dist.ReduceOp.SUM)
    ms_max = float(t.item())
    total_units = float(units.item())
    value = total_units / (ms_max * 1e-3)
    plan.close()

    # ---------------- e2e arm: host buffers through scdd_perm_bf (H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        pin = lambda a: torch.from_numpy(a).pin_memory().numpy()   # noqa: E731
        h_logy, h_c01, h_off, h_seeds = pin(logy), pin(c01), pin(off), pin(np.ascontiguousarray(seeds))
        h_cdr = pin(np.ascontiguousarray(cdr)) if cdr is not None else None
        hp.perm_bf(h_logy, h_c01, h_off, nperm, seed6=h_seeds, cdr=h_cdr, cfg=cfg)    # warm-up: allocates the cached device buffers
        n_e2e = max(1, args.steps)
        barrier()
        t0 = time.perf_counter()
        h2d = d2h = 0.0
        for _ in range(n_e2e):
            out, st = hp.perm_bf(h_logy, h_c01, h_off, nperm, seed6=h_seeds, cdr=h_cdr, cfg=cfg)
            h2d += st["h2d_bytes"]
            d2h += st["d2h_bytes"]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        # the same call with PAGEABLE arrays (what an R vector is): reported beside the pinned figure
        t0 = time.perf_counter()
        hp.perm_bf(logy, c01, off, nperm, seed6=seeds, cdr=cdr, cfg=cfg)
        dt_pageable = time.perf_counter() - t0
        tt = torch.tensor([dt, dt_pageable], dtype=torch.float64, device="cuda")
        uu = torch.tensor([float(ngr) * nperm * n_e2e, float(ngr) * nperm], dtype=torch.float64, device="cuda")
        bb = torch.tensor([h2d / n_e2e, d2h / n_e2e], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(uu, op=dist.ReduceOp.SUM)
            dist.all_reduce(bb, op=dist.ReduceOp.SUM)
        e2e = {"value": float(uu[0].item()) / float(tt[0].item()), "unit": UNIT,
               "h2d_bytes_per_step": float(bb[0].item()), "d2h_bytes_per_step": float(bb[1].item()),
               "steps": n_e2e,
               "timer": "host wall clock around scdd_perm_bf (H2D from pinned host memory, kernels, D2H), max over ranks; one "
                        "untimed warm-up call has allocated the library's cached device buffers (as a B = 1000 call amortises them)",
               "pageable_value": float(uu[1].item()) / float(tt[1].item()),
               "checksum": float(np.nansum(out))}

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    # ---------------- N > 1: the product's own multi-device path, one process (what the R shim calls)
    e2e_api = None
    if world > 1 and not args.no_api_leg and args.config == 3:
        time.sleep(2.0)                       # let the other ranks leave their devices
        hp.release_cache()
        hp.set_devices(list(range(world)))
        la, ca, oa_ = hp.gene_csr(X, cond)
        B_api = max(nperm, 40)
        hp.perm_bf(la[:oa_[400]], ca[:oa_[400]], oa_[:401], 8, seed6=seeds_all[:400], cfg=cfg)      # contexts + streams on all devices
        t0 = time.perf_counter()
        out_api, st_api = hp.perm_bf(la, ca, oa_, B_api, seed6=seeds_all, cfg=cfg)
        dt_api = time.perf_counter() - t0
        e2e_api = {"value": args.genes * B_api / dt_api, "unit": UNIT, "n_gpus": world, "permutations": B_api,
                   "wall_s": dt_api, "device_ms": st_api["total_ms"],
                   "h2d_bytes": st_api["h2d_bytes"], "d2h_bytes": st_api["d2h_bytes"],
                   "path": "ONE process, scdd_set_devices(0..N-1), one scdd_perm_bf call on the whole 20,000-gene block with pageable "
                           "host arrays: shards gathered into pinned staging, 8 permutations on cell-count shards measure the "
                           "per-gene cost, the rest runs on shards re-dealt by measured cost",
                   "checksum": float(np.nansum(out_api))}
        hp.set_devices([local])

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
    # through the partition index (4 B) and the permutation index (2 B), and writes 8 B per fit-set
    sets = 3 if args.config == 5 else 2
    nnz_total = float(nnz_t.item())
    alg_bytes = nnz_total * nperm * 14.0 * (1.5 if sets == 3 else 1.0) + (total_units / args.steps) * 8.0 * sets
    traffic = None
    prof = os.path.join(ROOT, "profiles", "fit_kernel_latest.json")
    if os.path.exists(prof) and args.config == 3:
        try:
            traffic = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, world),
                   "step": f"{nperm} device-drawn permutations of every resident gene "
                           f"({args.genes} x {args.perms_per_step} gene-permutations per GPU per step"
                           f"{', 3 fit-sets each' if sets == 3 else ''})",
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
                     "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one fit_kernel launch, ncu --set full capture "
                                       "summarised in profiles/fit_kernel_latest.json (static: ncu cannot run inside the timed region)"
                                       if traffic else None,
                     "kernel": "scdd::fit_kernel", "launches": fit_launches,
                     "avg_launch_ms": fit_ms / fit_launches if fit_launches else None,
                     "kernel_share_of_step": fit_ms / (ms_max * world) if ms_max > 0 else None,
                     "other_kernels_ms_per_step": (ms_max * world - fit_ms) / (args.steps * world) if ms_max > 0 else None,
                     "flops_definition": "algorithmic 24*T + 4*P per launch (T = point x component x iteration "
                                         "triples, P = point x iteration, both counted by the kernel)",
                     "peak_source": "DFMA microbenchmark in this run (MEASURED_PEAKS.json has no fp64 entry)",
                     "hbm": {"achieved_gbs": alg_bytes / (ms_max / args.steps * 1e-3) / 1e9 / world, "peak_gbs": hbm_peak,
                             "frac": alg_bytes / (ms_max / args.steps * 1e-3) / 1e9 / world / hbm_peak,
                             "note": "algorithmic bytes per GPU (gathered values + partition and u16 permutation indices "
                                     "+ scores) / step time; the path is fp64-bound, HBM is <1% utilised"}},
        "work": {"fit_sets": w[5], "triples": T, "point_iters": P, "tol_near": w[6], "bic_near": w[7],
                 "failed_sets": w[8], "iter_cap_hits": w[9], "const_sets": w[10], "rng_kernel_ms": rng_ms, "fit_kernel_ms": fit_ms},
    }
    if e2e_api:
        line["e2e_api"] = e2e_api
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args)
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
