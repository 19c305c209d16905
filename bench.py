#!/usr/bin/env python
"""bench.py — reactor-solves/sec on BASELINE.json configs[1] (N1 steady-state ensemble, 65,536 reactors per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--extras]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch: K0 prolog -> K2 fused BDF solve -> K5 un-scale of the
outlet rows, for every reactor of the rank's shard (weak scaling: 65,536 reactors per GPU; N>1 adds the single
NCCL all-gather of outlet rows).  `value` times the step with inputs resident in HBM; `e2e` times the
host-buffer C-ABI call (pinned host input table -> H2D -> kernels -> D2H of outlet/status).  `cpu_baseline` /
`--impl reference` time the reference's own CPU path on a bounded sample of the same rows: the UNMODIFIED
`PyREMOT.rmtExe` (oracle/_ref, placed there by oracle/build_ref.py; kind "reference"), one call per ensemble member,
one process per host core (fork) -- BASELINE.md section 3.  If the copy is absent the oracle port is timed instead
(SciPy LSODA at the reference's default tolerances on the NumPy restatement of its RHS; kind "port").
Also emitted by default: the dynamic model (N2) at BASELINE configs[2] (4,096 x 100 nodes) and a configs[3] slice
(1,184 x 400 nodes) with their own roofline and CPU figures, and the end-to-end rate with full 101-point profiles.
"""
from __future__ import annotations

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

N_PER_GPU = 65536
RTOL, ATOL = 1e-9, 1e-13
METRIC = "reactor-solves/sec"
WORKLOAD = "N1 steady-state methanol/DME packed-bed ensemble, 65536 reactors/GPU sweeping inlet T, P, feed composition (BASELINE configs[1])"


# ----------------------------------------------------------------------------- CPU path (oracle port)
def _cpu_worker(args):
    lo, hi, tol = args
    from oracle import pbr_oracle as O
    from pyremot_b200 import cases
    rows = cases.ensemble_rows(hi - lo, offset=lo)
    ens = O.Ensemble(cases.readme_case("N1"), {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]})
    kw = {} if tol is None else {"rtol": tol[0], "atol": tol[1]}
    _, data, nfev = O.solve_n1(ens, method="LSODA", **kw)
    return hi - lo, int(nfev.sum())


def _ref_worker(args):
    """The UNMODIFIED reference: one `PyREMOT.rmtExe(modelInput)` call per ensemble member (`PyREMOT/rmt.py:21-80`);
    the worker rebuilds modelInput (incl. the rate lambdas, which cannot be pickled) from the numeric row."""
    lo, hi, _ = args
    from oracle import ref_driver as ref
    from pyremot_b200 import cases
    rows = cases.ensemble_rows(hi - lo, offset=lo)
    for k in range(hi - lo):
        mi = cases.readme_case("N1", T=float(rows["T"][k]), P=float(rows["P"][k]), conc=[float(c) for c in rows["conc"][k]],
                               Tm=float(rows["T"][k]))
        res, _ = ref.run(mi)
        assert res["resModel"][0]["successStatus"]
    return hi - lo, 0


def reference_available():
    from oracle import ref_driver as ref
    return ref.available()


def cpu_path(n_inst, cores=None, tol=None, kind="port"):
    """Solve rows [0, n_inst) on `cores` processes with the reference itself (kind "reference") or the oracle port;
    returns (solves/s, seconds, total nfev)."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    worker = _ref_worker if kind == "reference" else _cpu_worker
    per = max(1, n_inst // (cores*2))
    chunks = [(lo, min(lo + per, n_inst), tol) for lo in range(0, n_inst, per)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(worker, [(0, 1, tol)]*cores)           # warm the workers (imports) outside the timed region
        t0 = time.perf_counter()
        res = pool.map(worker, chunks, chunksize=1)
        dt = time.perf_counter() - t0
    return n_inst/dt, dt, sum(r[1] for r in res)


def _n2_port_worker(i):
    from oracle import pbr_oracle as O
    from pyremot_b200 import cases
    rows = cases.ensemble_rows(1, offset=i)
    ens = O.Ensemble(cases.readme_case("N2"), {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]})
    _, _, nfev = O.solve_n2(ens, zNo=20, tNo=5, method="LSODA")
    return int(nfev[0])


def cpu_n2(cores):
    """CPU figures for the dynamic model: RHS evaluations per second of the reference's own
    `modelEquationN2` (one core, zNo = 20 / 100 / 400; a full reference solve at zNo = 100 is ~17 min per instance), and
    whole solves per second of the oracle port at zNo = 20 (SciPy LSODA, the reference's defaults), one instance per core."""
    import multiprocessing as mp
    out = {"cores": cores}
    if reference_available():
        from oracle import ref_driver as ref
        out["rhs_kind"] = "reference"
        out["rhs_evals_per_s_per_core"] = {f"zNo{z}": ref_rhs_n2_rate(ref, z, n) for z, n in ((20, 20), (100, 8), (400, 3))}
    else:
        out["rhs_kind"] = "port"
        out["rhs_evals_per_s_per_core"] = {"zNo100": cpu_rhs_rates()["n2_z100"]}
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        t0 = time.perf_counter()
        nfev = pool.map(_n2_port_worker, range(cores), chunksize=1)
        dt = time.perf_counter() - t0
    out["solve_zNo20"] = {"kind": "port", "value": cores/dt, "unit": "solves/s", "seconds": dt, "mean_nfev": float(np.mean(nfev)),
                          "sample": f"{cores} instances (rows 0..{cores-1}), zNo=20, tNo=5, period 0.5 s, SciPy LSODA rtol 1e-3/atol 1e-6"}
    r20 = out["rhs_evals_per_s_per_core"].get("zNo20")
    if r20:
        out["solve_zNo20_reference_derived"] = {
            "value": cores*r20/float(np.mean(nfev)), "unit": "solves/s",
            "note": "derived, not timed: cores x reference RHS evals/s at zNo=20 / RHS evals per solve counted on the port"}
    return out


def ref_rhs_n2_rate(ref, zNo, n):
    """Evaluations per second of the reference's modelEquationN2 at the README N2 case's initial state (one core)."""
    from pyremot_b200 import cases
    PyREMOT, H, setting = ref.load()
    cap = {}

    class _Stop(Exception):
        pass

    orig = H.PackedBedHomoReactorClass.modelEquationN2

    def spy(t, y, p):
        cap["y"], cap["p"] = np.array(y), p
        raise _Stop

    old = setting["N2"]["zNo"]
    H.PackedBedHomoReactorClass.modelEquationN2 = spy
    setting["N2"]["zNo"] = zNo
    try:
        with ref.quiet():
            PyREMOT.rmtExe(cases.readme_case("N2"))
    except BaseException:
        pass
    finally:
        H.PackedBedHomoReactorClass.modelEquationN2 = orig
        setting["N2"]["zNo"] = old
    ref.rhs("N2", cap["y"], cap["p"])
    t0 = time.perf_counter()
    for _ in range(n):
        ref.rhs("N2", cap["y"], cap["p"])
    return n/(time.perf_counter() - t0)


def ref_rhs_m9_rate(ref, zNo, n):
    """Evaluations per second of the reference's modelEquationM5 (model M9) at the case's initial state, one core."""
    from pyremot_b200 import cases
    PyREMOT, H, setting = ref.load()
    import PyREMOT.docs.pbReactor as R
    cls = R.PackedBedReactorClass
    orig = cls.modelEquationM5
    cap = {}

    def spy(t, y, a, b, c):
        cap["y"], cap["args"] = np.array(y), (a, b, c)
        raise RuntimeError("captured")

    old = dict(setting["S2"])
    cls.modelEquationM5 = spy
    setting["S2"]["zNo"] = zNo
    try:
        with ref.quiet():
            PyREMOT.rmtExe(cases.m9_case())
    except BaseException:
        pass
    finally:
        cls.modelEquationM5 = orig
        setting["S2"].update(old)
    with ref.quiet():
        orig(0.05, cap["y"], *cap["args"])
        t0 = time.perf_counter()
        for _ in range(n):
            orig(0.05, cap["y"], *cap["args"])
    return n/(time.perf_counter() - t0)


def cpu_rhs_rates():
    """RHS evaluations per second of the CPU path, one core, called the way the reference calls its callbacks
    (one reactor per call): modelEquationN1 and modelEquationN2 at zNo = 100 (oracle port)."""
    from oracle import pbr_oracle as O
    from pyremot_b200 import cases
    e1 = O.Ensemble(cases.readme_case("N1"))
    y1 = e1.iv_n1()
    e2 = O.Ensemble(cases.readme_case("N2"))
    Y2 = e2.iv_n2(100)
    out = {}
    for name, f, n in (("n1", lambda: e1.rhs_n1(y1), 300), ("n2_z100", lambda: e2.rhs_n2(Y2), 60)):
        f()
        t0 = time.perf_counter()
        for _ in range(n):
            f()
        out[name] = n/(time.perf_counter() - t0)
    return out


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    kind = "reference" if reference_available() else "port"
    # bounded sample per step: ~0.5-0.8 s per solve per core with the reference (0.35 s with the port)
    n = max(4*cores, 32) if kind == "reference" else max(8*cores, 64)
    times = []
    for s in range(a.warmup + a.steps):
        v, dt, nfev = cpu_path(n, cores, kind=kind)
        if s >= a.warmup:
            times.append(dt)
    total = sum(times)
    value = n*len(times)/total
    what = ("the unmodified PyREMOT.rmtExe (copy of the reference package under oracle/_ref), one call per member, ivp='default' (LSODA rtol 1e-3/atol 1e-6)" if kind == "reference"
            else "oracle port: SciPy LSODA rtol 1e-3/atol 1e-6 (the reference's defaults) on the NumPy RHS restatement")
    sample = f"rows 0..{n-1} of the seeded config-2 table per step, {what}, {cores} processes"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3*total/len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_per_step": n},
        "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock + throttle reasons sampled every ~10 ms through NVML while the timed region runs
    (same fields as the recipe's nvidia-smi query: clocks.sm, clocks.max.sm, clocks_event_reasons.*)."""
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        # LOCAL_RANK indexes CUDA_VISIBLE_DEVICES; NVML indexes physical devices
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else index
        self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        self.sm, self.reasons, self.stop = [], set(), False
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        while not self.stop:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.BITS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=2)

    def summary(self):
        try:
            mx = self.nv.nvmlDeviceGetMaxClockInfo(self.h, self.nv.NVML_CLOCK_SM)
        except Exception:
            mx = None
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": mx,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


# ----------------------------------------------------------------------------- GPU arm
def algorithmic_flops(model, stats):
    """Algorithmic FP64 operation count of one K2 launch from the per-instance integrator counters.

    Per-unit figures: every +,-,*,/ and every exp/log/log10/pow/sqrt counts 1.
    rate DAG ops are counted exactly from the compiled expressions; the balance / linear-algebra terms are
    hand counts of the formulas in csrc/pbr_device.cuh."""
    nx, n, R = model.n1_vars, model.n, model.n_rxn   # nx: integrated variables (extent coordinates: R + 2)
    reduced = nx < model.nv1
    ops_r, ops_rj = model.rates.op_counts()          # exact: rate DAG, and rate DAG + its sparse partials
    base = 16*n + 14*R + 40                          # balances: scaling, x_i, M, rho, Ergun, w, Cp, dH, energy
    rhs = ops_r + base
    # Jacobian assembly on top of the rate partials (hand counts of pbr_n1x_eval / pbr_n1_eval in csrc/pbr_device.cuh):
    # sum_i x_i dr/dx_i (2Rn); one column per extent (directional chain rule 8R + 6n, balances 30) or per species
    # (3R + 3n + 25); the p and theta columns and the T-derivatives of Cp / dH (9R + 9n + 32)
    cols = R*(8*R + 6*n + 30) if reduced else n*(3*R + 3*n + 25)
    jac = ops_rj + base + 2*R*n + cols + (9*R + 9*n + 32)
    lu = (2*nx**3)//3 + nx*nx
    solve = 2*nx*nx + 6*nx                           # triangular solves + Newton update
    step = 32*nx + 14*model.nv1                      # predictor + psi, history update, three weighted norms, weights
    nsteps, nrej, nfev, nlu = [float(x) for x in stats.sum(axis=1)]
    # every step attempt evaluates the Jacobian (which yields f as well: `jac` contains `rhs`); the other nfev - nlu RHS
    # evaluations are the verification iterations and the initial-step probes; one triangular solve per Newton iteration
    flops = (nfev - nlu)*rhs + nfev*solve + nlu*(jac + lu) + nlu*step
    return flops, {"rhs": rhs, "jac": jac, "lu": lu, "solve": solve, "step": step}


def algorithmic_flops_n2(model, stats, zNo):
    """Algorithmic FP64 operation count of one N2 solver launch (block kernels, extent coordinates) from the per-instance
    counters: per node and step attempt one analytic node Jacobian (which yields f as well), one nx x nx inverse
    (Gauss-Jordan, 2 nx^3), one Newton solve (g = Ainv b and the march, 4 nx^2 + 2 nx) and the step algebra (predictor, psi,
    history update, weights, norms, the affine pressure scan); the other nfev - nlu evaluations are plain RHS evaluations
    followed by a solve.  Hand counts of pbr_n2x_node / pbr_k4b_body in csrc/pbr_device.cuh; same conventions as N1."""
    n, R = model.n, model.n_rxn
    reduced = R < n
    nx = (R if reduced else n) + (0 if model.iso else 1)
    ops_r, ops_rj = model.rates.op_counts()
    base = 16*n + 14*R + 30
    rhs = ops_r + base
    cols = R*(8*R + 6*n + 24) if reduced else n*(3*R + 3*n + 20)
    jac = ops_rj + base + 2*R*n + cols + (9*R + 9*n + 24)
    inv = 2*nx**3
    solve = 4*nx*nx + 2*nx
    step = 32*nx + 14*model.nv2 + 20
    nsteps, nrej, nfev, nlu = [float(x) for x in stats.sum(axis=1)]
    per_node = (nfev - nlu)*(rhs + solve) + nlu*(jac + inv + solve + step)
    return per_node*zNo, {"rhs": rhs, "jac": jac, "inverse": inv, "solve": solve, "step": step}


def run_gpu(a):
    import torch
    import torch.distributed as dist
    from pyremot_b200 import cases
    from pyremot_b200.api import get_model

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback for the product path)"
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"        # keep NCCL's version banner out of stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N = N_PER_GPU
    mi = cases.readme_case("N1")
    model = get_model(mi, device=local)
    rows = cases.ensemble_rows(N, offset=rank*N)          # every rank regenerates its own shard from the seed
    sweep = {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]}
    tab_h = torch.from_numpy(model.instance_table(mi, sweep)).pin_memory()
    tab_d = tab_h.cuda()
    flush = torch.empty(256*1024*1024, dtype=torch.uint8, device="cuda")      # > 126 MB L2
    gathered = torch.empty((world*N, model.nv1), dtype=torch.float64, device="cuda") if world > 1 else None

    def step_device():
        par = model.prolog(tab_d)
        r = model.solve_n1(par, rtol=RTOL, atol=ATOL, n_out=101, profile=False, stats=True)
        out = model.unscale_n1(par, r["outlet_y"][None])          # [N, nv1, 1]
        if world > 1:
            dist.all_gather_into_tensor(gathered, out.view(N, model.nv1))
        return r, out

    # host-buffer path: pinned output buffers
    out_h = {"outlet": torch.empty((N, model.nv1), dtype=torch.float64).pin_memory().numpy(),
             "status": torch.empty((N,), dtype=torch.int32).pin_memory().numpy(),
             "stats": torch.empty((4, N), dtype=torch.int32).pin_memory().numpy()}

    def step_host():
        return model.solve_n1_host(tab_h.numpy(), rtol=RTOL, atol=ATOL, n_out=101, profile=False, out=out_h)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        r, out = step_device()
    torch.cuda.synchronize()
    assert int((r["status"] != 0).sum().item()) == 0, "some reactors failed to integrate"
    stats = r["stats"].cpu().numpy()

    # ---- timed: device-resident
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    l0 = model.launch_count
    barrier()
    with ClockSampler(local) as clk:
        for k in range(a.steps):
            flush.fill_(k)                                   # L2 flush between timed iterations (untimed)
            ev[k][0].record()
            par = model.prolog(tab_d)
            kev[k][0].record()
            r = model.solve_n1(par, rtol=RTOL, atol=ATOL, n_out=101, profile=False, stats=True)
            kev[k][1].record()
            out = model.unscale_n1(par, r["outlet_y"][None])
            if world > 1:
                dist.all_gather_into_tensor(gathered, out.view(N, model.nv1))
            ev[k][1].record()
        barrier()
    launches = model.launch_count - l0
    t_dev = sum(e0.elapsed_time(e1) for e0, e1 in ev)*1e-3
    t_k2 = sum(e0.elapsed_time(e1) for e0, e1 in kev)*1e-3/a.steps
    step_ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    phase_ms = {"prolog": float(np.mean([ev[k][0].elapsed_time(kev[k][0]) for k in range(a.steps)])),
                "k2": 1e3*t_k2,
                "epilog_and_gather": float(np.mean([kev[k][1].elapsed_time(ev[k][1]) for k in range(a.steps)]))}
    # ---- timed: end-to-end through the host-buffer C-ABI call
    for _ in range(2):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for k in range(a.steps):
        rh = step_host()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    assert (rh["status"] == 0).all()
    assert np.array_equal(rh["outlet"], out.view(N, model.nv1).cpu().numpy()), "host and device paths disagree"

    northstar = bench_northstar(model, mi, world, rank)        # every rank takes part (strong-scaling record, outside the timed region)
    tt = torch.tensor([t_dev, t_e2e, t_k2], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_dev, t_e2e, t_k2 = [float(x) for x in tt.cpu()]

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        fp64_peak = model.fp64_peak()
        flops, per_unit = algorithmic_flops(model, stats)
        # second convention: without the per-attempt analytic Jacobian (RHS + LU + solves + step algebra only)
        flops_nojac = flops - float(stats[3].sum())*(per_unit["jac"] - per_unit["rhs"])
        achieved = flops/t_k2/1e12
        k2_traffic, k2_traffic_src = traffic_record("pbr_k2_solve_n1")
        # secondary: standalone N2 RHS kernel (HBM-bound side of the path), zNo=100, inputs larger than L2
        rhs = bench_rhs_n2(model, mi, peaks)
        cores = os.cpu_count() or 1
        # dynamic model under the same clock: BASELINE configs[2] whole, and one GPU's worth of 8 waves of the configs[3] grid
        n2c3 = bench_n2(model, 4096, 100, "N2 dynamic model, 100 nodes, 4096 instances, start-up transient (BASELINE configs[2])",
                        "pbr_k4bs128_solve_n2", fp64_peak)
        n2c4 = bench_n2(model, 1184, 400, "N2 dynamic model, 400 nodes, 1184-instance slice (8 per SM) of BASELINE configs[3]",
                        "pbr_k4d_solve_n2", fp64_peak)
        if world == 1:
            n2c3["cpu_baseline"] = cpu_n2(cores)
        e2e_prof = bench_e2e_profile(model, tab_h, a.steps)
        cfg5 = bench_config5(model.device)
        m9 = bench_m9(model.device, world == 1)
        kind = "reference" if reference_available() else "port"
        n_cpu = max(16*cores, 128)                      # ~0.3-0.8 s per solve per core -> 5-15 s of CPU work per leg
        # the CPU path is timed at N = 1 only (at N > 1 the other ranks would sit in the final barrier for its duration)
        cpu_v, cpu_dt = (cpu_path(n_cpu, cores, kind=kind)[:2] if world == 1 else (None, None))
        port_v = cpu_path(n_cpu, cores, kind="port")[0] if (world == 1 and kind == "reference") else None
        line = {
            "metric": METRIC, "value": world*N*a.steps/t_dev, "unit": "solves/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": 1e3*t_dev/a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "instances_per_gpu": N, "rtol": RTOL, "atol": ATOL, "integrator": "BDF(1-5)/Newton, analytic Jacobian",
                       "outputs": "outlet rows (n+2) per reactor", "l2": "flushed between timed iterations (256 MiB write)"},
            "e2e": {"value": world*N*a.steps/t_e2e, "unit": "solves/s", "h2d_bytes_per_step": int(tab_h.numel()*8),
                    "d2h_bytes_per_step": int(sum(v.nbytes for v in out_h.values()))},
            "gpu_launches": int(launches),
            "ms_each_step": step_ms, "phase_ms": phase_ms,
            "clocks": clk.summary(),
            "roofline": {"kernel": "pbr_k2_solve_n1", "bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved/fp64_peak if fp64_peak else None,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch at this size from the committed
                         # `ncu --set full` record (algorithmic: 11.5 MB of per-reactor constants in, 4.2 MB of outlet rows out)
                         "traffic": k2_traffic, "traffic_unit": f"B/launch (ncu, {k2_traffic_src})" if k2_traffic else None,
                         "frac_without_jacobian": (flops_nojac/t_k2/1e12)/fp64_peak if fp64_peak else None,
                         "peak_source": "DFMA probe pbr_k_dfma_peak measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
                         "kernel_ms": 1e3*t_k2, "algorithmic_flops_per_launch": flops, "flops_per_unit": per_unit,
                         "mean_steps": float(stats[0].mean()), "mean_nfev": float(stats[2].mean()), "mean_nlu": float(stats[3].mean()),
                         "max_over_mean_nfev": float(stats[2].max()/stats[2].mean())},
            "rhs_n2_roofline": rhs,
            "n2_config3": n2c3, "n2_config4_slice": n2c4, "e2e_profile": e2e_prof, "config5_estimation": cfg5, "m9_dynamic": m9,
            "northstar_1m": northstar,
            "cpu_baseline": None if cpu_v is None else {
                "value": cpu_v, "unit": "solves/s", "cores": cores, "kind": kind, "port_value": port_v,
                "rhs_evals_per_s_per_core_port": cpu_rhs_rates(),
                "sample": f"rows 0..{n_cpu-1} of the same table, " +
                          ("the unmodified PyREMOT.rmtExe (oracle/_ref) per member at its defaults (LSODA rtol 1e-3/atol 1e-6)" if kind == "reference"
                           else "SciPy LSODA at the reference's default rtol 1e-3/atol 1e-6 on the NumPy RHS restatement") +
                          f", {cores} processes, {cpu_dt:.1f} s"},
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def bench_n2(model, N, zNo, label, kernel, fp64_peak, reps=3):
    """Dynamic model (N2) ensemble on one GPU: N instances x zNo nodes, tNo = 5, period 0.5 s, start-up from the feed-filled
    bed, same seeded T/P/composition rows as config 2.  Device-timed (CUDA events on the launching stream, one warm-up pass,
    `reps` timed passes, L2 flushed in between); roofline of the solver kernel from its per-instance counters."""
    import torch
    from pyremot_b200 import cases
    from pyremot_b200.api import get_model
    mi2 = cases.readme_case("N2")
    m2 = get_model(mi2, device=model.device)
    tNo, period = 5, 0.5
    rows = cases.ensemble_rows(N)
    tab = torch.from_numpy(m2.instance_table(mi2, {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]})).cuda()
    par = m2.prolog(tab)
    flush = torch.empty(256*1024*1024, dtype=torch.uint8, device="cuda")
    r = m2.solve_n2(par, zNo, tNo, period, rtol=RTOL, atol=ATOL)
    torch.cuda.synchronize()
    l0 = m2.launch_count
    ms = []
    for k in range(reps):
        flush.fill_(k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = m2.solve_n2(par, zNo, tNo, period, rtol=RTOL, atol=ATOL)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    launches = m2.launch_count - l0
    t = float(np.mean(ms))*1e-3
    st = r["stats"].cpu().numpy()
    # end to end through the host-buffer C-ABI call: host table in, dataYs [N][tNo][n+1][zNo] + status + counters out
    tab_h = tab.cpu().numpy()
    m2.solve_n2_host(tab_h, zNo, tNo, period, rtol=RTOL, atol=ATOL, scaled=False)
    t0 = time.perf_counter()
    rh = m2.solve_n2_host(tab_h, zNo, tNo, period, rtol=RTOL, atol=ATOL, scaled=False)
    t_e2e = time.perf_counter() - t0
    assert (rh["status"] == 0).all()
    flops, per_unit = algorithmic_flops_n2(m2, st, zNo)
    ach = flops/t/1e12
    return {"workload": label, "kernel": kernel, "instances": N, "zNo": zNo, "tNo": tNo, "period_s": period, "rtol": RTOL, "atol": ATOL,
            "ms": 1e3*t, "ms_each": ms, "value": N/t, "unit": "solves/s", "failed": int((r["status"] != 0).sum().item()),
            "e2e": {"value": N/t_e2e, "unit": "solves/s", "ms": 1e3*t_e2e, "h2d_bytes_per_step": int(tab_h.nbytes),
                    "d2h_bytes_per_step": int(rh["dataYs"].nbytes + rh["status"].nbytes + rh["stats"].nbytes)},
            "mean_steps": float(st[0].mean()), "mean_nfev": float(st[2].mean()), "node_steps_per_s": float(st[0].sum())*zNo/t,
            "gpu_launches": int(launches),
            "roofline": {"kernel": kernel, "bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": ach/fp64_peak if fp64_peak else None, "algorithmic_flops_per_launch": flops,
                         "flops_per_node_unit": per_unit, "traffic": None,
                         "peak_source": "DFMA probe pbr_k_dfma_peak measured in this run"}}


def bench_config5(model_device):
    """BASELINE configs[4]: parameter-estimation Jacobian d(outlet)/d(theta) [8 x 6] of 16,384 base reactors -- by forward
    sensitivities (K2s: 98,304 instances at the working tolerance) and by the forward differences the config describes (114,688
    solves at rtol 1e-11 so that integration noise stays below the perturbation).  Device-resident inputs, CUDA events."""
    import torch
    from pyremot_b200 import cases
    from pyremot_b200.api import get_model
    mi = cases.readme_case("N1", theta_as_vars=True)
    theta = list(cases.MEOH_DME_THETA)
    m = get_model(mi, params=tuple(theta), device=model_device)
    Nb, p = 16384, len(theta)
    rows = cases.ensemble_rows(Nb)
    base = m.instance_table(mi, {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]})
    tab_s = torch.from_numpy(np.concatenate([base]*p)).cuda()
    kp = torch.arange(p, dtype=torch.int32, device="cuda").repeat_interleave(Nb)
    blocks = [base]
    for k in range(p):
        t = base.copy()
        t[:, 9 + m.n + k] *= 1.0 + 1e-4
        blocks.append(t)
    tab_f = torch.from_numpy(np.concatenate(blocks)).cuda()

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), r
    ms_s, rs = timed(lambda: m.solve_n1_sens(m.prolog(tab_s), kp, rtol=RTOL, atol=ATOL))
    ms_f, rf = timed(lambda: m.solve_n1(m.prolog(tab_f), rtol=1e-11, atol=1e-15, n_out=2, profile=False))
    return {"workload": "d(outlet)/d(theta), 6 kinetic constants x 16384 base reactors (BASELINE configs[4])",
            "sensitivities": {"kernel": "pbr_k2s_solve_n1_sens", "instances": Nb*p, "rtol": RTOL, "ms": ms_s, "jacobians_per_s": Nb/(ms_s*1e-3),
                              "failed": int((rs["status"] != 0).sum().item())},
            "forward_differences": {"kernel": "pbr_k2_solve_n1", "instances": Nb*(p + 1), "rtol": 1e-11, "ms": ms_f,
                                    "jacobians_per_s": Nb/(ms_f*1e-3), "failed": int((rf["status"] != 0).sum().item())}}


def bench_m9(model_device, with_cpu):
    """Widening row f3: the dynamic dimensional model M9 at the reference's own grid (solverSetting['S2']: 100 nodes, 10 intervals),
    1,184 swept reactors (8 warps per SM) through 2 s of start-up with the wall 10 K below the feed.  Device-resident inputs,
    CUDA events; the CPU figure beside it is the oracle's RHS rate (the reference's own RHS is a Python loop over nodes with a
    print per call: one rmtExe("M9") at this grid takes hours, see oracle/make_golden.py g_m9)."""
    import torch
    from pyremot_b200 import cases
    from pyremot_b200.api import get_model
    N, zNo, tNo, period = 1184, 100, 10, 2.0
    mi = cases.m9_case(period=period)
    m = get_model(mi, device=model_device)
    rows = cases.ensemble_rows(N)
    sweep = {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"] - 10.0}
    tab = torch.from_numpy(m.instance_table(mi, dict(sweep))).cuda()
    par = m.prolog(tab)
    m.solve_n2(par, zNo, tNo, period, rtol=RTOL, atol=ATOL)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = m.solve_n2(m.prolog(tab), zNo, tNo, period, rtol=RTOL, atol=ATOL)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    st = r["stats"].cpu().numpy()
    rec = {"workload": "M9 dynamic dimensional model, 100 nodes x 10 intervals (solverSetting['S2']), 1184 swept reactors, 2 s start-up",
           "kernel": "pbr_k4b_solve_n2 (block per reactor, full coordinates; pressure/velocity march by block-wide fixed-point sweeps)", "ms": ms,
           "solves_per_s": N/(ms*1e-3), "failed": int((r["status"] != 0).sum().item()), "rtol": RTOL,
           "mean_steps": float(st[0].mean()), "mean_nfev": float(st[2].mean()), "mean_nlu": float(st[3].mean())}
    if with_cpu:
        from oracle import pbr_oracle as O
        ens = O.Ensemble(mi, {k: v[:64] for k, v in sweep.items()})
        Y = ens.iv_m9(zNo)
        t0, n = time.perf_counter(), 0
        while time.perf_counter() - t0 < 2.0:
            ens.rhs_m9(Y)
            n += 64
        rec["cpu_rhs_evals_per_s_per_core_port"] = n/(time.perf_counter() - t0)
        rec["gpu_rhs_evals_per_s_inside_solver"] = float(st[2].sum())/(ms*1e-3)
        if reference_available():
            from oracle import ref_driver as ref
            r = ref_rhs_m9_rate(ref, zNo, 10)                  # the reference's own modelEquationM5 (a Python loop over the nodes)
            rec["cpu_rhs_evals_per_s_per_core_reference"] = r
            # derived, not timed: one reference solve at this grid = RHS evaluations of LSODA with a dense finite-difference Jacobian
            # (700 columns) -- hours; cores x evals/s / (the evaluations SciPy's BDF on the port needs per solve: ~2,900)
            rec["cpu_solves_per_s_reference_derived"] = {"value": (os.cpu_count() or 1)*r/2900.0, "unit": "solves/s",
                                                         "note": "derived: cores x reference RHS evals/s / 2,900 RHS evaluations per solve "
                                                                 "(oracle BDF with grouped finite-difference Jacobian, zNo 100, 10 intervals)"}
    return rec


def bench_northstar(model, mi, world, rank):
    """North-star workload: 1,048,576 N1 reactors over the `world` GPUs of the job (strong scaling: the total is fixed), rows dealt
    to the ranks by `ensemble.deal` on the inlet temperature (the stiffness proxy of DESIGN.md), ONE all-gather of outlet rows.
    Called by every rank; device-timed, max over ranks.  Returns the record on rank 0."""
    import torch
    import torch.distributed as dist
    from pyremot_b200 import cases, ensemble
    NT = 1048576
    rows = cases.ensemble_rows(NT)
    perm = ensemble.deal(rows["T"], world) if world > 1 else np.arange(NT)
    lo, hi = ensemble.shard_bounds(NT, rank, world)
    mine = perm[lo:hi]
    tab = torch.from_numpy(model.instance_table(mi, {"T": rows["T"][mine], "P": rows["P"][mine], "conc": rows["conc"][mine],
                                                       "Tm": rows["T"][mine]})).cuda()

    def run():
        par = model.prolog(tab)
        r = model.solve_n1(par, rtol=RTOL, atol=ATOL, n_out=101, profile=False, stats=False)
        out = model.unscale_n1(par, r["outlet_y"][None])[:, :, 0].contiguous()
        return ensemble.gather_rows(out, NT, world, rank), r
    for _ in range(2):                       # two warm-up passes: the second one already runs out of the caching allocator's blocks
        full, r = run()
        torch.cuda.synchronize()
        del full, r
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    full, r = run()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1), float((r["status"] != 0).sum().item())], dtype=torch.float64, device="cuda")
    if world > 1:
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t[0] = tm[0]
    ms, failed = float(t[0].item()), int(t[1].item())
    return {"workload": "1,048,576 N1 reactors over the job's GPUs (north star; strong scaling), rows dealt by ensemble.deal, one all-gather",
            "reactors": NT, "gpus": world, "ms": ms, "value": NT/(ms*1e-3), "unit": "solves/s", "failed": failed,
            "gathered_rows": int(full.shape[0])}


def traffic_record(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, read at run time from the committed ncu summary
    (profiles/traffic.json, written from an `ncu --set full` capture by tools/ncu_traffic.py); None when there is no record."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[kernel]
        return float(rec["dram_bytes_read"]) + float(rec["dram_bytes_write"]), rec.get("source")
    except Exception:
        return None, None


def bench_e2e_profile(model, tab_h, steps):
    """The rmtExe contract end to end: 101-point profiles of every reactor through the host-buffer C-ABI call with pinned
    buffers (7.9 MB H2D, 424 MB + 5.5 MB D2H per 65,536 reactors); chunked over two streams inside libpbr."""
    import torch
    N = tab_h.shape[0]
    K = 101
    out = {"dataYs": torch.empty((N, model.nv1, K), dtype=torch.float64).pin_memory().numpy(),
           "outlet": torch.empty((N, model.nv1), dtype=torch.float64).pin_memory().numpy(),
           "status": torch.empty((N,), dtype=torch.int32).pin_memory().numpy(),
           "stats": torch.empty((4, N), dtype=torch.int32).pin_memory().numpy()}
    model.solve_n1_host(tab_h.numpy(), rtol=RTOL, atol=ATOL, n_out=K, profile=True, out=out)
    reps = max(2, min(3, steps))
    t0 = time.perf_counter()
    for _ in range(reps):
        r = model.solve_n1_host(tab_h.numpy(), rtol=RTOL, atol=ATOL, n_out=K, profile=True, out=out)
    dt = (time.perf_counter() - t0)/reps
    assert (r["status"] == 0).all()
    d2h = int(sum(v.nbytes for v in out.values()))
    return {"value": N/dt, "unit": "solves/s", "ms_per_step": 1e3*dt, "points_per_profile": K, "h2d_bytes_per_step": int(tab_h.numel()*8),
            "d2h_bytes_per_step": d2h, "d2h_GBps": d2h/dt/1e9, "outputs": "dataYs [N][n+2][101] + outlet rows, pinned host buffers"}


def bench_rhs_n2(model, mi, peaks):
    """K3 standalone: RHS evals/s and achieved GB/s (16*(n+1)*zNo B per instance-eval) at zNo=100."""
    import torch
    from pyremot_b200 import cases
    from pyremot_b200.api import get_model
    mi2 = cases.readme_case("N2")
    m2 = get_model(mi2, device=model.device)
    zNo, N = 100, 65536                       # 2 x 367 MB of state: larger than L2
    rows = cases.ensemble_rows(N)
    tab = torch.from_numpy(m2.instance_table(mi2, {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]})).cuda()
    par = m2.prolog(tab)
    Y = torch.empty((N, m2.nv2, zNo), dtype=torch.float64, device="cuda")
    Y[:, :m2.n, :] = par[16:16 + m2.n].t()[:, :, None]
    Y[:, m2.n, :] = torch.linspace(0, 0.2, zNo, device="cuda", dtype=torch.float64)[None, :]
    dY = torch.empty_like(Y)
    for _ in range(3):
        m2.rhs_n2(par, Y, out=dY)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        m2.rhs_n2(par, Y, out=dY)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/reps
    bytes_ = 16.0*m2.nv2*zNo*N
    gbs = bytes_/(ms*1e-3)/1e9
    peak = peaks.get("hbm_gbs", 6650.0)
    # N1 RHS (thread per reactor): 16*(n+2) B of state per eval + 8*k_par B of per-instance constants
    N1 = 4*1024*1024
    y = torch.rand((m2.nv1, N1), dtype=torch.float64, device="cuda")*0.1 + 0.5
    y[m2.n] = 0.99
    y[m2.n + 1] = 0.1
    par1 = m2.prolog(tab[:1].repeat(N1, 1).contiguous())
    f1 = torch.empty_like(y)
    for _ in range(3):
        m2.rhs_n1(par1, y, out=f1)
    e0.record()
    for _ in range(reps):
        m2.rhs_n1(par1, y, out=f1)
    e1.record()
    torch.cuda.synchronize()
    ms1 = e0.elapsed_time(e1)/reps
    # bytes the kernel actually moves (SASS: 23 LDG.E.64 + 8 STG.E.64): state in + derivative out + the table rows the
    # plain RHS uses (the n rows of c0 and the mu row are dead in pbr_n1_eval<false> and never loaded)
    bpe = 8.0*(2*m2.nv1 + m2.k_par - m2.n - 1)
    b1 = bpe*N1
    src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    return {"kernel": "pbr_k3_rhs_n2", "bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs/peak,
            "peak_source": src, "zNo": zNo, "instances": N, "ms": ms, "rhs_evals_per_s": N/(ms*1e-3),
            "traffic": traffic_record("pbr_k3_rhs_n2")[0],
            "node_evals_per_s": N*zNo/(ms*1e-3),
            "note": "16*(n+1)*zNo B per instance-eval; at ~650 executed FP64 ops per node this kernel sits on the FP64 side of the ridge",
            "rhs_n1": {"kernel": "pbr_k_rhs_n1", "bound": "hbm", "achieved": b1/(ms1*1e-3)/1e9, "peak": peak, "unit": "GB/s",
                       "frac": b1/(ms1*1e-3)/1e9/peak, "bytes_per_eval": bpe, "instances": N1, "ms": ms1,
                       "rhs_evals_per_s": N1/(ms1*1e-3), "state_only_GBs": 16.0*m2.nv1*N1/(ms1*1e-3)/1e9}}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--extras", action="store_true", help="accepted for compatibility: the N2 lines are emitted by default")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_gpu(a)
