#!/usr/bin/env python
"""Headline benchmark: sweeps/sec of the O(3) spin-fermion DQMC hot path (local updates + wrap + UDT stabilization)
at L=16, beta=40 (n=1024, M=400, safe_mult=10), one independent Markov chain per GPU.

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference (oracle port)

One "step" = one sweep = M x { propagate; local_updates } incl. the K=M/safe_mult stabilizations of that
direction = half a reference "udsweep" (dqmc_framework.jl:259-262).  Prints ONE JSON line (rank 0).
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

CONFIGS = {
    "L16_beta40": dict(L=16, slices=400, safe_mult=10),   # BASELINE.json configs[3] (headline)
    "L12_beta40": dict(L=12, slices=400, safe_mult=10),   # configs[2]
    "L8_beta20": dict(L=8, slices=200, safe_mult=10),     # configs[1]
    "L4_beta5": dict(L=4, slices=50, safe_mult=10),       # configs[0]
    "L20_beta40": dict(L=20, slices=400, safe_mult=10),   # configs[4] (+ time-displaced G)
}
MODEL = dict(hoppings="1.0,0.5,-0.5,-1.0", mu=-0.5, lam=0.5, r=2.0, c=3.0, u=1.0, delta_tau=0.1, box=0.5)


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/ncu_traffic.json), or None."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(path)).get(kernel)
    except Exception:
        return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_inputs(cfg, chain, nsweeps):
    """Documented counter-based generators: field seed 1234+chain, proposal stream seed 5678+chain."""
    N, M = cfg["L"] ** 2, cfg["slices"]
    field = np.random.Generator(np.random.Philox(1234 + chain)).random((M, N, 3)).T.copy(order="F")
    u = np.random.Generator(np.random.Philox(5678 + chain)).random(4 * N * M * nsweeps)
    return field, u


# ------------------------------------------------------------------------------------------------ CPU baseline
def blas_threads(n=None):
    """Pin the BLAS/OpenMP pools of NumPy/SciPy to `n` threads (default: every host core) and return the count actually
    set.  torchrun exports OMP_NUM_THREADS=1, which would silently turn the all-cores baseline into a one-thread one."""
    n = n or os.cpu_count()
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n)
    except Exception:
        return None
    return n


class OracleChain:
    """The oracle port on the FULL chain of the workload (M slices): init once (untimed), then `block()` times one
    safe_mult block = safe_mult x {propagate; local_updates} incl. its stabilization, continuing the same Markov chain."""

    def __init__(self, cfg):
        import oracle
        from oracle.dqmc import UniformStream
        self.cfg = cfg
        L, M, sm = cfg["L"], cfg["slices"], cfg["safe_mult"]
        self.om = oracle.OracleDQMC(oracle.Params(L=L, slices=M, safe_mult=sm, lam=MODEL["lam"], all_checks=False))
        field, _ = synthetic_inputs(cfg, 0, 0)
        t0 = time.perf_counter()
        self.om.init(field)
        self.t_init = time.perf_counter() - t0
        self.gen = np.random.Generator(np.random.Philox(5678))
        self.UniformStream = UniformStream

    def block(self):
        L, sm = self.cfg["L"], self.cfg["safe_mult"]
        st = self.UniformStream(self.gen.random(4 * L * L * sm))
        t0 = time.perf_counter()
        acc = 0.0
        for _ in range(sm):
            self.om.propagate()
            acc += self.om.local_updates(st)
        return time.perf_counter() - t0, acc / sm


def cpu_block_sample(cfg, nblocks=1):
    """Time `nblocks` safe_mult blocks (safe_mult x {propagate; local_updates}, one stabilization each) of the oracle
    port on the host cores at the workload's L, on a short beta = 4 PROXY chain (4 blocks: the cost of a block does not
    depend on beta, the setup does) so the bounded sample stays within seconds; scale to a sweep."""
    import oracle
    from oracle.dqmc import UniformStream
    L, sm = cfg["L"], cfg["safe_mult"]
    Mshort = 4 * sm
    p = oracle.Params(L=L, slices=Mshort, safe_mult=sm, lam=MODEL["lam"], all_checks=False)
    om = oracle.OracleDQMC(p)
    rs = np.random.Generator(np.random.Philox(1234))
    om.init(rs.random((Mshort, L * L, 3)).T.copy())
    st = UniformStream(np.random.Generator(np.random.Philox(5678)).random(4 * L * L * sm * (nblocks + 1)))
    t0 = time.perf_counter()
    acc = 0.0
    for _ in range(nblocks * sm):
        om.propagate()
        acc += om.local_updates(st)
    dt = time.perf_counter() - t0
    blocks_per_sweep = cfg["slices"] // sm
    return dt / nblocks * blocks_per_sweep, acc / (nblocks * sm)


def cpu_one_thread(cfg):
    """The same sample at ONE BLAS thread (the reference's default: app/dqmc.jl:33-37 sets BLAS threads to 1)."""
    try:
        from threadpoolctl import threadpool_limits
    except Exception:
        return None
    with threadpool_limits(limits=1):
        return cpu_block_sample(cfg, 1)[0]


def run_reference(args, cfg, rank, world):
    """CPU arm: the oracle port on the workload's own chain (full M), all host cores, BLAS threads set explicitly.  Under
    torchrun only rank 0 works: ONE CPU chain on the whole host (it is a host figure, not a per-GPU one)."""
    if rank != 0:
        return
    cores = blas_threads() or os.cpu_count()
    chain = OracleChain(cfg)
    nblk = cfg["slices"] // cfg["safe_mult"]
    times = []
    for it in range(args.warmup + args.steps):
        t_blk, acc = chain.block()
        if it >= args.warmup:
            times.append(t_blk)
    t_step = float(np.mean(times))              # wall time of one step = one safe_mult block = 1/nblk of a sweep
    t = t_step * nblk                           # extrapolated time of a whole sweep
    val = 1.0 / t
    line = {"impl": "reference", "metric": "sweeps/sec", "value": val, "unit": "sweeps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True,
            "step_is": "1/%d of a sweep (one safe_mult block of the running chain); value = 1 / (%d x step time)" % (nblk, nblk),
            "ms_per_sweep_extrapolated": t * 1e3,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
            "config": {"workload": args.config, "note": "CPU restatement of the reference path (NumPy/SciPy, OpenBLAS %d threads)" % cores,
                       "L": cfg["L"], "slices": cfg["slices"], "safe_mult": cfg["safe_mult"]},
            "cpu_baseline": {"value": val, "unit": "sweeps/s", "cores": cores, "blas_threads": cores, "kind": "port",
                             "sample": "each step = 1 of the %d safe_mult blocks (10 x {propagate; local_updates}, 1 stabilization) "
                                       "of the workload's own chain (full M = %d, consecutive blocks of one Markov chain; stack "
                                       "init %.0f s untimed), x %d = one sweep" % (nblk, cfg["slices"], chain.t_init, nblk)},
            "e2e": {"value": val, "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------------ GPU arm
def kernel_rooflines(mc, cfg, hbm_gbs, peak_src):
    """Per-kernel roofline from CUDA-event timings inside the library (dqmc_bench_kernel)."""
    n, N, sm = mc.n, cfg["L"] ** 2, cfg["safe_mult"]
    out = {}
    t_cublas = mc.bench_kernel(2, 5)
    zg_n = 8.0 * n ** 3 / (t_cublas * 1e-3) / 1e12              # cuBLAS ZGEMM at the workload's n, TFLOP/s
    dg_big = 2.0 * 4096 ** 3 / (mc.bench_kernel(13, 5) * 1e-3) / 1e12
    zg_big = 8.0 * 4096 ** 3 / (mc.bench_kernel(14, 5) * 1e-3) / 1e12
    f64_peak = max(zg_n, dg_big, zg_big)                        # FP64 ceiling = the best cuBLAS rate measured in this run
    out["fp64_peak_tflops"] = f64_peak
    out["fp64_probes_tflops"] = {"cublas_zgemm_n%d" % n: zg_n, "cublas_dgemm_4096": dg_big, "cublas_zgemm_4096": zg_big}
    t = mc.bench_kernel(6, 20)
    out["copy_G"] = {"ms": t, "GB/s": 32.0 * n * n / (t * 1e-3) / 1e9}
    t = mc.bench_kernel(0, 20)
    out["wrap"] = {"bound": "hbm", "ms": t, "achieved": 64.0 * n * n / (t * 1e-3) / 1e9, "peak": hbm_gbs, "unit": "GB/s",
                   "algorithmic_bytes": 64.0 * n * n}
    t = mc.bench_kernel(7, 10)
    out["b_chain"] = {"bound": "hbm", "ms": t, "achieved": 32.0 * n * n / (t * 1e-3) / 1e9, "peak": hbm_gbs, "unit": "GB/s",
                      "algorithmic_bytes": 32.0 * n * n, "flops": sm * 128.0 * n * n}
    t = mc.bench_kernel(1, 10)
    out["zgemm"] = {"bound": "tensor", "ms": t, "achieved": 8.0 * n ** 3 / (t * 1e-3) / 1e12, "peak": f64_peak, "unit": "TFLOP/s"}
    t = mc.bench_kernel(3, 5)
    out["udt"] = {"bound": "tensor", "ms": t, "achieved": 10.67 * n ** 3 / (t * 1e-3) / 1e12, "peak": f64_peak, "unit": "TFLOP/s",
                  "algorithmic_flops": 10.67 * n ** 3}
    t = mc.bench_kernel(4, 5)
    # two GEMMs + QR (5.33) + apply Q^H to the rhs (8) + triangular solve (4) + final GEMM
    fl = (3 * 8 + 5.33 + 8 + 4) * n ** 3
    out["calculate_greens"] = {"bound": "tensor", "ms": t, "achieved": fl / (t * 1e-3) / 1e12, "peak": f64_peak, "unit": "TFLOP/s",
                               "algorithmic_flops": fl}
    t = mc.bench_kernel(5, 3)
    out["local_updates_slice"] = {"ms": t, "us_per_proposal": t * 1e3 / N}
    for k, v in out.items():
        if isinstance(v, dict) and "achieved" in v:
            v["frac"] = v["achieved"] / v["peak"]
            v["peak_source"] = peak_src if v["bound"] == "hbm" else "best of cuBLAS DGEMM/ZGEMM measured in this run"
    return out


def g_vs_oracle(mc, config):
    """max |G - G_oracle| / max |G| of the freshly initialised chain 0 on the sampled entries / probe vector the oracle
    left in tests/golden/bench_init_<config>.npz (tests/golden/make_bench_init_golden.py), or None if there is no fixture."""
    path = os.path.join(ROOT, "tests", "golden", f"bench_init_{config}.npz")
    if not os.path.exists(path):
        return None
    g = np.load(path)
    G = mc.greens
    n = G.shape[0]
    e_s = float(np.max(np.abs(G[np.ix_(g["rows"], g["cols"])] - g["sample"]))) / float(g["gmax"])
    e_v = float(np.max(np.abs(G @ g["v"] - g["Gv"]))) / (float(g["gmax"]) * np.sqrt(2.0 * n))
    return max(e_s, e_v)


def extra_config_series(args, local_rank):
    """The other BASELINE configs as extra series (rank 0, N=1 only): sweeps/s of configs[0..2] and configs[4] (L=20,
    beta=40) with the time of dqmc_measure_tdgfs and the memory it holds (fermion_measurements.jl:1320-1331)."""
    from dqmc_b200 import DQMC, Params
    out = {}
    for name in ("L4_beta5", "L8_beta20", "L12_beta40", "L20_beta40"):
        cfg = CONFIGS[name]
        L, M, sm = cfg["L"], cfg["slices"], cfg["safe_mult"]
        p = Params(L=L, slices=M, safe_mult=sm, delta_tau=MODEL["delta_tau"], lambda_=MODEL["lam"], r=MODEL["r"], c=MODEL["c"],
                   u=MODEL["u"], mu1=MODEL["mu"], mu2=MODEL["mu"], hoppings=MODEL["hoppings"], box=MODEL["box"],
                   Bfield=False, all_checks=True)
        mc = DQMC(p, device=local_rank, delay=args.delay)
        nsw = 3
        field, u = synthetic_inputs(cfg, 0, nsw)
        mc.init(field)
        mc.set_uniforms(u)
        mc.sweep(None)
        mc.set_timing(2); mc.timers()
        nacc = 0
        for _ in range(nsw - 1):
            nacc += mc.sweep(None)[0]
        mc.sync()
        tm = mc.timers()
        err, _ = mc.checks()
        t = tm["sweep"] * 1e-3 / (nsw - 1)
        row = {"value": 1.0 / t, "unit": "sweeps/s", "ms_per_sweep": t * 1e3, "n": mc.n, "slices": M,
               "acceptance": nacc / ((nsw - 1) * M * L * L), "max_propagation_error": err,
               "phases_ms_per_sweep": {k: tm[k] / (nsw - 1) for k in ("wrap", "local_updates", "stack_udt", "calculate_greens")}}
        if name == "L20_beta40":
            mc.set_timing(0)
            t0 = time.perf_counter()
            mc.measure_tdgfs()
            mc.sync()
            row["measure_tdgfs_s"] = time.perf_counter() - t0
            nn16 = 16.0 * mc.n ** 2
            row["tdgf_bytes"] = {"Gt0_G0t": 2 * M * nn16, "udt_chains": 4 * (M // sm) * (2 * nn16 + 8 * mc.n)}
            g1, g2 = mc.Gt0(1), mc.G0t(1)
            row["tdgf_tau0_identity_err"] = float(np.max(np.abs(g1 - g2 - np.eye(mc.n))))
            mc.deallocate_tdgfs_stacks()
        out[name] = row
        mc.close()
    return out


def run_ours(args, cfg, rank, world, local_rank):
    import torch
    from dqmc_b200 import DQMC, Params, UniformStream
    from dqmc_b200 import parallel

    torch.cuda.set_device(local_rank)
    if world > 1:
        parallel.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L, M, sm = cfg["L"], cfg["slices"], cfg["safe_mult"]
    N = L * L
    p = Params(L=L, slices=M, safe_mult=sm, delta_tau=MODEL["delta_tau"], lambda_=MODEL["lam"], r=MODEL["r"], c=MODEL["c"],
               u=MODEL["u"], mu1=MODEL["mu"], mu2=MODEL["mu"], hoppings=MODEL["hoppings"], box=MODEL["box"],
               Bfield=False, all_checks=bool(args.all_checks))
    mc = DQMC(p, device=local_rank, delay=args.delay)
    nsw = args.warmup + args.steps + 2
    field, u = synthetic_inputs(cfg, rank, nsw)
    mc.init(field)
    g_rel = g_vs_oracle(mc, args.config) if rank == 0 else None

    # ---- device-resident arm: uniforms already in HBM, timed with CUDA events on the library's stream
    mc.set_uniforms(u)
    for _ in range(args.warmup):
        mc.sweep(None)
    mc.set_timing(1)                                 # total of dqmc_sweep only: the stabilization steps run as CUDA graphs
    mc.timers()
    sampler = ClockSampler(local_rank)
    if world > 1:
        torch.distributed.barrier()
    mc.sync()
    sampler.start()
    launches0 = mc.kernel_launches()
    # DQMC_BENCH_PROFILER_RANGE=1: cudaProfilerStart/Stop around the timed region, for `ncu --profile-from-start off` (the launch
    # list of profiles/ covers exactly the timed sweeps; a number printed under a profiler is never a bench value)
    prof_range = os.environ.get("DQMC_BENCH_PROFILER_RANGE") == "1"
    if prof_range:
        torch.cuda.cudart().cudaProfilerStart()
    t0 = time.perf_counter()
    nacc = 0
    for _ in range(args.steps):
        a, _c = mc.sweep(None)
        nacc += a
    mc.sync()
    wall = time.perf_counter() - t0
    if prof_range:
        torch.cuda.cudart().cudaProfilerStop()
    launches = mc.kernel_launches() - launches0
    tm = mc.timers()
    clocks = sampler.stop()
    t_dev = tm["sweep"] * 1e-3                       # CUDA-event time of the K sweeps
    # phase breakdown from a separate pass with every phase timer on (graphs bypassed); NOT the headline
    mc.set_timing(2)
    mc.timers()
    nph = min(2, args.steps)
    for _ in range(nph):
        mc.sweep(None)
    mc.sync()
    tm_ph = mc.timers()
    mc.set_timing(0)
    if world > 1:
        tt = torch.tensor([t_dev, wall], dtype=torch.float64, device="cuda")
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        t_dev, wall = tt.tolist()
    value = world * args.steps / t_dev

    # ---- end-to-end arm: host uniforms in (pinned), configuration + counters out, every sweep, wall clock
    nsw_e = args.steps
    field_e, u_e = synthetic_inputs(cfg, 1000 + rank, nsw_e)
    upin = torch.empty(4 * N * M, dtype=torch.float64, pin_memory=True).numpy()
    if world > 1:
        torch.distributed.barrier()
    mc.sync()
    t0 = time.perf_counter()
    chi_n, chi_s1, chi_s2 = 0, 0.0, 0.0
    for k in range(nsw_e):
        upin[:] = u_e[k * 4 * N * M:(k + 1) * 4 * N * M]
        mc.sweep(UniformStream(upin))
        _conf = mc.hsfield
        chi = mc.measure_chi_dynamic()              # the step's measured observable: chi(q, i omega), read back every sweep
        chi_n += 1; chi_s1 = chi_s1 + chi; chi_s2 = chi_s2 + chi * chi
    mc.sync()
    wall_e = time.perf_counter() - t0
    chi_bytes = chi.size * 8
    if world > 1:
        tt = torch.tensor([wall_e], dtype=torch.float64, device="cuda")
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        wall_e = tt.item()
    e2e = world * nsw_e / wall_e

    # ---- extra (not the headline): two independent chains per GPU, one host thread and one stream each.  A single chain
    # leaves most SMs idle during its latency-bound stabilizations (the QR panel chain runs on one 8-CTA cluster), so a
    # second chain's local updates fill them; reported as aggregate sweeps/s next to the one-chain-per-GPU headline.
    two = None
    if args.two_chains:
        mc2 = DQMC(p, device=local_rank, delay=args.delay)
        f2, u2 = synthetic_inputs(cfg, 500 + rank, args.steps + 1)
        mc2.init(f2)
        mc2.set_uniforms(u2)
        mc.set_uniforms(u)
        mc2.sweep(None)
        mc.sweep(None)

        def _run(m):
            for _ in range(args.steps):
                m.sweep(None)
            m.sync()
        ths = [threading.Thread(target=_run, args=(m,)) for m in (mc, mc2)]
        mc.sync(); mc2.sync()
        t0 = time.perf_counter()
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        wall2 = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([wall2], dtype=torch.float64, device="cuda")
            torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
            wall2 = tt.item()
        two = {"chains_per_gpu": 2, "value": world * 2 * args.steps / wall2, "unit": "sweeps/s (aggregate)", "timing": "wall clock",
               "ms_per_sweep_per_chain": wall2 / args.steps * 1e3}
        mc2.close()

    # ---- extra: the same sweep with the magnetic flux on (complex Peierls-phase plaquette factors; all O(3) test XMLs of the
    # reference use it, the README's speed test had it off -> headline off, this series on)
    bfield = None
    if args.bfield_series:
        pb = Params(L=L, slices=M, safe_mult=sm, delta_tau=MODEL["delta_tau"], lambda_=MODEL["lam"], r=MODEL["r"], c=MODEL["c"],
                    u=MODEL["u"], mu1=MODEL["mu"], mu2=MODEL["mu"], hoppings=MODEL["hoppings"], box=MODEL["box"],
                    Bfield=True, all_checks=bool(args.all_checks))
        mcb = DQMC(pb, device=local_rank, delay=args.delay)
        mcb.init(field)
        mcb.set_uniforms(u)
        mcb.sweep(None)
        mcb.set_timing(1); mcb.timers()
        for _ in range(args.steps):
            mcb.sweep(None)
        mcb.sync()
        tb = mcb.timers()["sweep"] * 1e-3
        errb, _nr = mcb.checks()
        bfield = {"value": args.steps / tb, "unit": "sweeps/s (this rank's chain)", "ms_per_sweep": tb / args.steps * 1e3,
                  "max_propagation_error": errb}
        mcb.close()

    # ---- pooled measurement bins across chains (the path's only collective; outside the timed region)
    # the real observable bins: per (qy, qx, i omega) bin of chi the chain's [n, mean, var] -> one all-reduce of
    # [n, sum x, sum |x|^2] -> pooled mean / variance exactly as statistics.jl:22-36 (boson_measurements.jl:48-56)
    chi_mean = chi_s1 / chi_n
    chi_var = (chi_s2 - chi_n * chi_mean ** 2) / max(chi_n - 1, 1)
    pooled_mean, pooled_var = parallel.combined_mean_and_var(chi_n, chi_mean, chi_var)
    err, nonreal = mc.checks()

    extra = extra_config_series(args, local_rank) if (rank == 0 and world == 1 and args.extra_configs) else None
    if rank == 0:
        hbm_gbs, peak_src = measured_peaks()
        kr = kernel_rooflines(mc, cfg, hbm_gbs, peak_src)
        phases = {k: tm_ph[k] / nph for k in ("wrap", "local_updates", "stack_udt", "calculate_greens")}
        phases["sweep_with_phase_timers"] = tm_ph["sweep"] / nph
        # sweep-level roofline: algorithmic FP64 flops at the measured cuBLAS ZGEMM ceiling + wrap bytes at HBM peak
        n = mc.n
        acc_rate = nacc / (args.steps * M * N)
        f_sweep = acc_rate * M * N * 32.0 * n * n + (M // sm) * 107.3 * n ** 3
        # what the device really executes per stabilization.  Full matrices: UDT (QR 5.33 + Q^H 8 + T-product 8)
        # + calculate_greens (2.5 GEMMs x 8 + QR 5.33 + Q^H on the rhs 8 + triangular solve 4) = 58.7 n^3.  Half-matrix path (default
        # for this model, n % 32 == 0): UDT (paired QR 2.67 + Q^H on n/2 columns 4 + half T-product 4) + calculate_greens (three half
        # GEMMs 12 + paired QR 2.67 + Q^H on the n/2 right-hand sides 4 + triangular solve 2) = 31.3 n^3
        paired = os.environ.get("DQMC_PAIRED", "1") != "0" and n % 32 == 0
        c_own = 31.3 if paired else 58.7
        f_sweep_own = acc_rate * M * N * 32.0 * n * n + (M // sm) * c_own * n ** 3
        b_sweep = M * 64.0 * n * n
        f64_peak = kr["fp64_peak_tflops"]
        t_roof = f_sweep / (f64_peak * 1e12) + b_sweep / (hbm_gbs * 1e9)
        t_roof_own = f_sweep_own / (f64_peak * 1e12) + b_sweep / (hbm_gbs * 1e9)
        dom = max(("wrap", "local_updates", "stack_udt", "calculate_greens"), key=phases.get)
        dom_map = {"wrap": "wrap", "stack_udt": "udt", "calculate_greens": "calculate_greens", "local_updates": None}
        if dom_map[dom] is not None:
            rf = dict(kr[dom_map[dom]])
            roof = {"kernel": dom_map[dom], "bound": rf["bound"], "achieved": rf["achieved"], "peak": rf["peak"], "unit": rf["unit"],
                    "frac": rf["frac"], "traffic": ncu_traffic({"wrap": "apply_chain_kernel", "udt": "qr_panel_paired_kernel",
                                                                   "calculate_greens": "larfb_kernel"}[dom_map[dom]]),
                    "peak_source": rf["peak_source"]}
        else:
            # local updates: FP64 work = rank-4 Woodbury updates (32 n^2 flops per accepted proposal, flushed as GEMMs)
            t_lu = phases["local_updates"] * 1e-3
            ach = acc_rate * M * N * 32.0 * n * n / t_lu / 1e12
            roof = {"kernel": "lu_block_kernel", "bound": "tensor", "achieved": ach, "peak": f64_peak, "unit": "TFLOP/s",
                    "frac": ach / f64_peak, "traffic": ncu_traffic("lu_block_kernel"),
                    "traffic_note": "DRAM bytes per launch (one time slice) from ncu; algorithmic flops per launch = accepted x 32 n^2",
                    "peak_source": "best of cuBLAS DGEMM/ZGEMM measured in this run",
                    "limiter": "latency/issue of the serial Metropolis chain (the flush GEMMs are the only tensor work)",
                    "serial_floor_us_per_proposal": kr["local_updates_slice"]["us_per_proposal"]}
        ncores = blas_threads() or os.cpu_count()
        t_cpu, acc_cpu = cpu_block_sample(cfg, 1)
        t_cpu1 = cpu_one_thread(cfg)
        blas_threads()
        line = {"metric": "sweeps/sec", "value": value, "unit": "sweeps/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
                "config": {"workload": args.config, "model": "O(3) spin-fermion, square lattice, Assaad checkerboard, B-field off",
                           "L": L, "beta": M * MODEL["delta_tau"], "slices": M, "safe_mult": sm, "n": n, "delay": mc_delay(args),
                           "acceptance": acc_rate, "chains": world, "parallelism": f"{world} independent chains (replicas only)",
                           "l2": "no flush: each sweep streams the 1.3 GB UDT stack (>> 126 MB L2); G (16 MB) is L2-resident by design"},
                "e2e": {"value": e2e, "unit": "sweeps/s", "h2d_bytes_per_step": 8 * 4 * N * M,
                        "d2h_bytes_per_step": 8 * 3 * N * M + 24 + chi_bytes},
                "gpu_launches": launches, "clocks": clocks, "wall_ms_per_step": wall / args.steps * 1e3,
                "roofline": roof,
                "sweep_roofline": {"flops": f_sweep, "bytes": b_sweep, "t_roofline_ms": t_roof * 1e3, "frac": t_roof / (t_dev / args.steps),
                                   "flops_note": "reference algorithm: 107.3 n^3 per stabilization (SURVEY 8d)",
                                   "flops_executed": f_sweep_own, "frac_executed": t_roof_own / (t_dev / args.steps),
                                   "flops_executed_note": "what the device executes: %.1f n^3 per stabilization (DESIGN.md 4; %s)" % (c_own, "half-matrix path, paired Householder QR" if paired else "full matrices"),
                                   "fp64_peak_tflops": f64_peak, "hbm_gbs": hbm_gbs},
                "phases_ms_per_sweep": phases, "kernels": kr, "two_chains_per_gpu": two, "bfield_on": bfield,
                "cpu_baseline": {"value": 1.0 / t_cpu, "unit": "sweeps/s", "cores": os.cpu_count(), "kind": "port",
                                 "blas_threads": ncores,
                                 "sample": "1 of %d safe_mult blocks (10 x {propagate; local_updates}, 1 stabilization) at L=%d on a "
                                           "beta=4 proxy chain (block cost is beta-independent), scaled to a full sweep; oracle port, "
                                           "NumPy/SciPy + OpenBLAS, %d threads; --impl reference runs the full-M chain"
                                           % (M // sm, L, ncores), "acceptance": acc_cpu,
                                 "one_blas_thread_value": 1.0 / t_cpu1 if t_cpu1 else None,
                                 "one_blas_thread_note": "same sample at 1 BLAS thread, the reference driver's default (app/dqmc.jl:33-37)"},
                "checks": {"max_propagation_error": err, "nonreal_detratios": nonreal, "g_vs_oracle_rel": g_rel,
                           "g_vs_oracle_note": "init G of chain 0 against the oracle's fingerprint (tests/golden/bench_init_*.npz)",
                           "pooled_chi": {"bins": int(np.asarray(pooled_mean).size), "samples_per_chain": chi_n, "chains": world,
                                          "chi_static_mean": float(np.asarray(pooled_mean).ravel()[0]),
                                          "chi_static_var": float(np.asarray(pooled_var).ravel()[0])}},
                "extra_configs": extra}
        emit(line)
    mc.close()
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def mc_delay(args):
    return args.delay if args.delay > 0 else 16


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the process's original stdout; everything else any library prints meanwhile (NCCL's
    version banner, warnings) was redirected to stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                                   # file-descriptor level: also catches C libraries (NCCL banner)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="L16_beta40", choices=sorted(CONFIGS))
    ap.add_argument("--delay", type=int, default=0)
    ap.add_argument("--all-checks", type=int, default=1)
    ap.add_argument("--bfield-series", type=int, default=1, help="also time the sweep with the magnetic flux on (extra)")
    ap.add_argument("--two-chains", type=int, default=1, help="also time two chains per GPU (extra, not the headline)")
    ap.add_argument("--extra-configs", type=int, default=1, help="also time the other BASELINE configs (extra series, N=1 only)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
    else:
        run_ours(args, cfg, rank, world, local_rank)


if __name__ == "__main__":
    main()
