"""ctypes host layer over libpbr.so (include/pbr.h) — the only caller of the C ABI.

``Model`` turns the static part of a PyREMOT ``modelInput`` (components, reactions, rate expressions,
process type — what `PyREMOT/docs/rmtCore.py:63-183` prepares per call) into a compiled sm_100a module;
its methods launch the kernels on torch-owned device memory.  There is NO CPU fallback: without the
built library or without a CUDA device every compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

from . import components as DB
from . import ratec

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpbr.so")
CACHE_DIR = os.environ.get("PBR_CACHE_DIR", os.path.join(_HERE, "_cache"))

ISO_THERMAL = "iso-thermal"   # `PyREMOT/docs/modelSetting.py:21-23`


class PbrError(RuntimeError):
    pass


_lib = None


def lib():
    """Load libpbr.so (built in-tree by ``python -m pyremot_b200.build`` / ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PbrError(f"{LIB_PATH} is missing: build it with `python -m pyremot_b200.build` "
                       "(there is no CPU fallback for this path)")
    L = C.CDLL(LIB_PATH)
    vp, dp, ip, i, l, d = C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_long, C.c_double
    sigs = {
        "pbr_create": (vp, [i]),
        "pbr_destroy": (None, [vp]),
        "pbr_last_error": (C.c_char_p, [vp]),
        "pbr_version": (C.c_char_p, []),
        "pbr_set_species": (i, [vp, i, i, dp, dp, dp, dp, dp, dp, i]),
        "pbr_set_model_variant": (i, [vp, i]),
        "pbr_load_rates": (i, [vp, C.c_char_p, i, C.c_char_p, C.POINTER(C.c_char_p)]),
        "pbr_model_dims": (i, [vp] + [C.POINTER(i)]*7),
        "pbr_n1_layout": (i, [vp] + [C.POINTER(i)]*3),
        "pbr_model_source": (C.c_char_p, [vp]),
        "pbr_model_cubin": (i, [vp, C.POINTER(vp), C.POINTER(C.c_ulong)]),
        "pbr_kernel_attr": (i, [vp, C.c_char_p, C.POINTER(i), C.POINTER(i), C.POINTER(i)]),
        "pbr_prolog": (i, [vp, l, dp, dp, vp]),
        "pbr_rhs_n1": (i, [vp, l, dp, dp, dp, vp]),
        "pbr_jac_n1": (i, [vp, l, dp, dp, dp, dp, vp]),
        "pbr_rhs_n2": (i, [vp, l, i, dp, dp, dp, vp]),
        "pbr_solve_n1": (i, [vp, l, dp, d, d, i, i, dp, dp, ip, ip, vp]),
        "pbr_unscale_n1": (i, [vp, l, i, dp, dp, dp, vp]),
        "pbr_solve_n1_sens": (i, [vp, l, dp, ip, d, d, i, dp, dp, ip, ip, vp]),
        "pbr_unscale_sens_n1": (i, [vp, l, dp, dp, dp, dp, vp]),
        "pbr_solve_n1_host": (i, [vp, l, dp, d, d, i, i, dp, dp, ip, ip]),
        "pbr_solve_n2": (i, [vp, l, i, i, d, dp, d, d, i, dp, ip, ip, vp]),
        "pbr_solve_n2_resume": (i, [vp, l, i, i, i, d, dp, dp, i, d, d, i, dp, ip, ip, vp]),
        "pbr_solve_n2_host_resume": (i, [vp, l, i, i, i, d, dp, dp, i, d, d, i, dp, dp, ip, ip]),
        "pbr_unscale_n2": (i, [vp, l, i, i, dp, dp, dp, vp]),
        "pbr_solve_n2_host": (i, [vp, l, i, i, d, dp, d, d, i, dp, dp, ip, ip]),
        "pbr_math_probe": (i, [vp, l, dp, dp, vp]),
        "pbr_fp64_peak": (i, [vp, i, C.POINTER(d)]),
        "pbr_launch_count": (l, [vp]),
    }
    for name, (res, args) in sigs.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _lib = L
    return L


EXPORTS = ("pbr_create pbr_destroy pbr_last_error pbr_version pbr_set_species pbr_set_model_variant pbr_load_rates pbr_model_dims pbr_n1_layout "
           "pbr_model_source pbr_model_cubin pbr_kernel_attr pbr_prolog pbr_rhs_n1 pbr_jac_n1 pbr_rhs_n2 pbr_solve_n1 "
           "pbr_unscale_n1 pbr_solve_n1_sens pbr_unscale_sens_n1 pbr_solve_n1_host pbr_solve_n2 pbr_solve_n2_resume pbr_unscale_n2 pbr_solve_n2_host pbr_solve_n2_host_resume "
           "pbr_math_probe pbr_fp64_peak pbr_launch_count").split()


# ----------------------------------------------------------------------------- reaction parsing
_TERM = re.compile(r"([0-9.]*)([a-zA-Z0-9.]+)")


def parse_reactions(reactions, comps):
    """Reaction strings -> (S [R][n], dH25 [R] J/mol, dcp [R][4]).

    Same grammar as `PyREMOT/docs/rmtUtility.py:172-249` (regex per side, reactant coefficients
    negative); species absent from ``comps`` drop out of S (`rmtReaction.py:64-97`) but still count in
    dH25 (`rmtThermo.py:129-198`) and in the Cp sum of the dH(T) correction (`rmtThermo.py:258-312`).
    """
    S, dH25, dcp = [], [], []
    for expr in reactions.values():
        lhs, rhs = expr.replace("<", "").replace(">", "").replace(" ", "").split("=")
        terms = [(s, -(float(c) if c else 1.0)) for c, s in _TERM.findall(lhs)]
        terms += [(s, (float(c) if c else 1.0)) for c, s in _TERM.findall(rhs)]
        row = [0.0]*len(comps)
        h, poly = 0.0, [0.0]*4
        for s, nu in terms:
            if s in comps:
                row[comps.index(s)] += nu
            if s not in DB.COMPONENTS:
                raise Exception("Component database is not up to date!")
            h += nu*DB.dHf25(s)
            for k in range(4):
                poly[k] += nu*DB.cp_coef(s)[k]
        S.append(row)
        dH25.append(h*1000.0)
        dcp.append(poly)
    return np.array(S, dtype=np.float64), np.array(dH25, dtype=np.float64), np.array(dcp, dtype=np.float64)


def _arr(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.c_void_p)


class Model:
    """Compiled model = (components, reactions, rate expressions, process type) bound to one CUDA device.

    ``device=None`` -> compile-only (no GPU needed; used by the build check and CPU-side tests).
    ``params``: names of scalar VARS entries that become per-instance runtime parameters.
    """

    def __init__(self, modelInput, params=(), device=0, cache_dir=CACHE_DIR, variant=None):
        L = lib()
        comps = list(modelInput["feed"]["components"]["shell"])
        for s in comps:
            if s not in DB.componentSymbolList:          # `PyREMOT/rmt.py:55-57`
                raise Exception("Component database is not up to date!")
        self.comps = comps
        self.n = len(comps)
        self.iso = modelInput["operating-conditions"]["process-type"] == ISO_THERMAL
        self.S, self.dH25, self.dcp = parse_reactions(modelInput["reactions"], comps)
        self.rates = ratec.compile_rates(modelInput["reaction-rates"], self.n, params=params)
        if self.rates.n_rxn != self.S.shape[0]:
            raise PbrError(f"{self.S.shape[0]} reactions but {self.rates.n_rxn} rate expressions")
        self.params = list(params)
        self.device = device
        self._ctx = L.pbr_create(-1 if device is None else int(device))
        if not self._ctx:
            raise PbrError(f"pbr_create({device}) failed: no usable CUDA device (this path has no CPU fallback)")
        mw, p_mw = _arr([DB.MW(s) for s in comps])
        cp, p_cp = _arr([DB.cp_coef(s) for s in comps])
        vis, p_vis = _arr([DB.vis_coef(s) for s in comps])
        s_, p_s = _arr(self.S)
        dh, p_dh = _arr(self.dH25)
        dc, p_dc = _arr(self.dcp)
        self._check(L.pbr_set_species(self._ctx, self.n, self.S.shape[0], p_mw, p_cp, p_vis, p_s, p_dh, p_dc, int(self.iso)))
        # "M7": the dimensional steady-state model (`PyREMOT/docs/rmtCore.py:300-309`), solved through the N1 kernels
        self.variant = (modelInput.get("model") if variant is None else variant)
        self.variant = self.variant if self.variant in ("M7", "M8", "M9") else "N"
        self._check(L.pbr_set_model_variant(self._ctx, {"N": 0, "M7": 1, "M8": 2, "M9": 3}[self.variant]))
        if cache_dir:
            os.makedirs(cache_dir, exist_ok=True)
        log = C.c_char_p()
        self._check(L.pbr_load_rates(self._ctx, self.rates.cuda_source.encode(), len(self.params),
                                     (cache_dir or "").encode(), C.byref(log)))
        self.nvrtc_log = (log.value or b"").decode()
        d = [C.c_int() for _ in range(7)]
        self._check(L.pbr_model_dims(self._ctx, *[C.byref(x) for x in d]))
        _, self.n_rxn, self.n_par, self.nv1, self.nv2, self.k_in, self.k_par = [x.value for x in d]
        d = [C.c_int() for _ in range(3)]
        self._check(L.pbr_n1_layout(self._ctx, *[C.byref(x) for x in d]))
        self.n1_vars, self.n1_block, self.n1_smem = [x.value for x in d]      # K2: integrated variables, block, smem

    # -- plumbing
    def _check(self, rc):
        if rc != 0:
            raise PbrError(f"libpbr error {rc}: {lib().pbr_last_error(self._ctx).decode()}")

    def close(self):
        if getattr(self, "_ctx", None):
            lib().pbr_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def source(self):
        return lib().pbr_model_source(self._ctx).decode()

    @property
    def launch_count(self):
        return lib().pbr_launch_count(self._ctx)

    def kernel_attr(self, name):
        r, s, l = C.c_int(), C.c_int(), C.c_int()
        self._check(lib().pbr_kernel_attr(self._ctx, name.encode(), C.byref(r), C.byref(s), C.byref(l)))
        return {"regs": r.value, "static_smem": s.value, "local_bytes": l.value}

    def _stream(self):
        import torch
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self):
        import torch
        return torch.device("cuda", self.device)

    # -- per-instance input table (host)
    def instance_table(self, modelInput, sweep=None, N=None):
        """inst_in [N][k_in] (numpy, host): the numbers runN1/runN2 read from modelInput
        (`pbHomoReactor.py:2716-2782`), per instance.  ``sweep``: {"T","P","VoFlRa","conc"[N,n],"ReInDi",
        "ReLe","PaDi","BeVoFr","U","Tm","VARS:<name>"} -> per-instance arrays overriding the base value."""
        sweep = dict(sweep or {})
        if N is None:
            N = len(next(iter(sweep.values()))) if sweep else 1
        oc, fe, re_, eh = (modelInput["operating-conditions"], modelInput["feed"], modelInput["reactor"],
                           modelInput["external-heat"])
        t = np.empty((N, self.k_in), dtype=np.float64)
        base = {"T": oc["temperature"], "P": oc["pressure"], "VoFlRa": fe["volumetric-flowrate"],
                "ReInDi": re_["ReInDi"], "ReLe": re_["ReLe"], "PaDi": re_["PaDi"], "BeVoFr": re_["BeVoFr"],
                "U": eh["OvHeTrCo"], "Tm": eh["MeTe"]}
        for col, key in enumerate(("T", "P", "VoFlRa", "ReInDi", "ReLe", "PaDi", "BeVoFr", "U", "Tm")):
            t[:, col] = sweep.pop(key, base[key])
        t[:, 9:9 + self.n] = np.asarray(sweep.pop("conc", fe["concentration"]), dtype=np.float64)
        for k, name in enumerate(self.params):
            t[:, 9 + self.n + k] = sweep.pop("VARS:" + name, self.rates.param_defaults[k])
        if self.variant in ("M7", "M8", "M9"):  # `pbReactor.py:1238`, `:1545`, `:2504`: both are inputs of those models
            t[:, 9 + self.n + len(self.params)] = sweep.pop("GaMiVi", fe["mixture-viscosity"])
            t[:, 10 + self.n + len(self.params)] = sweep.pop("EfHeTrAr", eh["EfHeTrAr"])
        if self.variant in ("M8", "M9"):  # `pbReactor.py:1626`, `:2034`
            t[:, 11 + self.n + len(self.params)] = sweep.pop("SuGaVe0", fe["superficial-velocity"])
        if self.variant == "M9":          # catalyst heat capacity in the energy balance, `pbReactor.py:2048-2052`, `:2540`
            t[:, 12 + self.n + len(self.params)] = sweep.pop("CaDe", re_["CaDe"])
            t[:, 13 + self.n + len(self.params)] = sweep.pop("CaSpHeCa", re_["CaSpHeCa"])
        if sweep:
            raise KeyError(f"unknown sweep keys {sorted(sweep)} (rate parameters must be listed in params=)")
        return t

    # -- kernels on torch tensors (float64, CUDA, contiguous)
    def prolog(self, inst_in):
        import torch
        N = inst_in.shape[0]
        assert inst_in.is_cuda and inst_in.dtype == torch.float64 and inst_in.is_contiguous() and inst_in.shape[1] == self.k_in
        par = torch.empty((self.k_par, N), dtype=torch.float64, device=inst_in.device)
        self._check(lib().pbr_prolog(self._ctx, N, inst_in.data_ptr(), par.data_ptr(), self._stream()))
        return par

    def rhs_n1(self, par, y, out=None):
        import torch
        N = par.shape[1]
        assert y.shape == (self.nv1, N) and y.is_contiguous() and y.dtype == torch.float64
        f = torch.empty_like(y) if out is None else out
        self._check(lib().pbr_rhs_n1(self._ctx, N, par.data_ptr(), y.data_ptr(), f.data_ptr(), self._stream()))
        return f

    def jac_n1(self, par, y):
        import torch
        N = par.shape[1]
        f = torch.empty_like(y)
        J = torch.empty((self.nv1, self.nv1, N), dtype=torch.float64, device=y.device)
        self._check(lib().pbr_jac_n1(self._ctx, N, par.data_ptr(), y.data_ptr(), f.data_ptr(), J.data_ptr(), self._stream()))
        return f, J

    def rhs_n2(self, par, Y, out=None):
        import torch
        N, nv, zNo = Y.shape
        assert nv == self.nv2 and N == par.shape[1] and Y.is_contiguous() and Y.dtype == torch.float64
        dY = torch.empty_like(Y) if out is None else out
        self._check(lib().pbr_rhs_n2(self._ctx, N, zNo, par.data_ptr(), Y.data_ptr(), dY.data_ptr(), self._stream()))
        return dY

    def solve_n1(self, par, rtol=1e-9, atol=1e-13, n_out=101, profile=True, max_steps=0, stats=True):
        """K2 on device tensors. Returns dict(prof_y [n_out,nv1,N] or None, outlet_y [nv1,N], status [N], stats [4,N])."""
        import torch
        N = par.shape[1]
        dev = par.device
        prof = torch.empty((n_out, self.nv1, N), dtype=torch.float64, device=dev) if profile else None
        outlet = torch.empty((self.nv1, N), dtype=torch.float64, device=dev)
        status = torch.empty((N,), dtype=torch.int32, device=dev)
        st = torch.empty((4, N), dtype=torch.int32, device=dev) if stats else None
        self._check(lib().pbr_solve_n1(self._ctx, N, par.data_ptr(), rtol, atol, n_out, max_steps,
                                       prof.data_ptr() if profile else None, outlet.data_ptr(), status.data_ptr(),
                                       st.data_ptr() if stats else None, self._stream()))
        return {"prof_y": prof, "outlet_y": outlet, "status": status, "stats": st}

    def solve_n1_sens(self, par, kp, rtol=1e-9, atol=1e-13, max_steps=0):
        """K2s + K5s: outlet of every instance and its derivative with respect to runtime rate parameter ``kp[i]``.

        par [k_par, N] (CUDA), kp [N] int32 (CUDA).  Returns dict(outlet [N, nv1] = [x_i, P, T], doutlet [N, nv1], status, stats)."""
        import torch
        N = par.shape[1]
        dev = par.device
        assert kp.shape == (N,) and kp.dtype == torch.int32 and kp.device == dev
        outlet = torch.empty((self.nv1, N), dtype=torch.float64, device=dev)
        sens = torch.empty((self.nv1, N), dtype=torch.float64, device=dev)
        status = torch.empty((N,), dtype=torch.int32, device=dev)
        st = torch.empty((4, N), dtype=torch.int32, device=dev)
        self._check(lib().pbr_solve_n1_sens(self._ctx, N, par.data_ptr(), kp.data_ptr(), rtol, atol, max_steps, outlet.data_ptr(),
                                            sens.data_ptr(), status.data_ptr(), st.data_ptr(), self._stream()))
        dout = torch.empty((N, self.nv1), dtype=torch.float64, device=dev)
        self._check(lib().pbr_unscale_sens_n1(self._ctx, N, par.data_ptr(), outlet.data_ptr(), sens.data_ptr(), dout.data_ptr(),
                                              self._stream()))
        return {"outlet": self.unscale_n1(par, outlet[None])[:, :, 0], "doutlet": dout, "status": status, "stats": st}

    def unscale_n1(self, par, y):
        """K5: scaled y [K,nv1,N] -> dataYs [N,nv1,K]."""
        import torch
        K, nv, N = y.shape
        out = torch.empty((N, nv, K), dtype=torch.float64, device=y.device)
        self._check(lib().pbr_unscale_n1(self._ctx, N, K, par.data_ptr(), y.data_ptr(), out.data_ptr(), self._stream()))
        return out

    def solve_n1_host(self, inst_in, rtol=1e-9, atol=1e-13, n_out=101, profile=True, max_steps=0, out=None):
        """Whole N1 path through the host-buffer C-ABI call (H2D, K0, K2, K5, D2H inside).

        inst_in: numpy [N,k_in].  Returns dict(dataYs [N,nv1,n_out] or None, outlet [N,nv1], status [N], stats [4,N]).
        ``out`` may carry preallocated (e.g. pinned) numpy buffers under the same keys."""
        inst_in = np.ascontiguousarray(inst_in, dtype=np.float64)
        N = inst_in.shape[0]
        out = out or {}
        prof = out.get("dataYs") if profile else None
        if profile and prof is None:
            prof = np.empty((N, self.nv1, n_out), dtype=np.float64)
        outlet = out.get("outlet") if out.get("outlet") is not None else np.empty((N, self.nv1), dtype=np.float64)
        status = out.get("status") if out.get("status") is not None else np.empty((N,), dtype=np.int32)
        stats = out.get("stats") if out.get("stats") is not None else np.empty((4, N), dtype=np.int32)
        self._check(lib().pbr_solve_n1_host(
            self._ctx, N, inst_in.ctypes.data_as(C.c_void_p), rtol, atol, n_out, max_steps,
            prof.ctypes.data_as(C.c_void_p) if profile else None, outlet.ctypes.data_as(C.c_void_p),
            status.ctypes.data_as(C.c_void_p), stats.ctypes.data_as(C.c_void_p)))
        return {"dataYs": prof, "outlet": outlet, "status": status, "stats": stats}

    def solve_n2(self, par, zNo, tNo, period, rtol=1e-9, atol=1e-13, max_steps=0, stats=True, iv=None, k_first=0,
                 restart_every_interval=False):
        """K4 on device tensors. Returns dict(snaps_y [N,tNo-k_first,nv2,zNo] scaled, status [N], stats [4,N]).

        ``iv`` [N,nv2,zNo] (scaled state, e.g. ``snaps_y[:, k_first-1]`` of an earlier run) resumes at interval ``k_first``
        — the reference's own restart, ``IV = dataYs[:, -1]`` (`pbHomoReactor.py:3685`).  ``restart_every_interval``
        reproduces the reference's per-interval ``solve_ivp`` restarts (order 1, fresh initial step)."""
        import torch
        N = par.shape[1]
        dev = par.device
        assert par.device.index == self.device, "tensor lives on another GPU than this model"
        if iv is not None:
            assert iv.shape == (N, self.nv2, zNo) and iv.is_contiguous() and iv.dtype == torch.float64 and iv.device == dev
        snaps = torch.empty((N, tNo - k_first, self.nv2, zNo), dtype=torch.float64, device=dev)
        status = torch.empty((N,), dtype=torch.int32, device=dev)
        st = torch.empty((4, N), dtype=torch.int32, device=dev) if stats else None
        self._check(lib().pbr_solve_n2_resume(self._ctx, N, zNo, tNo, int(k_first), float(period), par.data_ptr(),
                                              iv.data_ptr() if iv is not None else None, 1 if restart_every_interval else 0,
                                              rtol, atol, max_steps, snaps.data_ptr(), status.data_ptr(),
                                              st.data_ptr() if stats else None, self._stream()))
        return {"snaps_y": snaps, "status": status, "stats": st}

    def unscale_n2(self, par, snaps_y):
        """K5b: scaled snapshots [N,tNo,nv2,zNo] -> dataYs [N,tNo,n+1,zNo] = [x; T]."""
        import torch
        N, tNo, nv, zNo = snaps_y.shape
        out = torch.empty((N, tNo, self.n + 1, zNo), dtype=torch.float64, device=snaps_y.device)
        self._check(lib().pbr_unscale_n2(self._ctx, N, tNo, zNo, par.data_ptr(), snaps_y.data_ptr(), out.data_ptr(), self._stream()))
        return out

    def solve_n2_host(self, inst_in, zNo, tNo, period, rtol=1e-9, atol=1e-13, max_steps=0, scaled=True, iv=None, k_first=0,
                      restart_every_interval=False):
        """Whole N2 path through the host-buffer C-ABI call. Returns dict(dataYs [N,tNo-k_first,n+1,zNo], snaps_y, status, stats).
        ``iv`` / ``k_first``: resume from a saved scaled state (see ``solve_n2``)."""
        inst_in = np.ascontiguousarray(inst_in, dtype=np.float64)
        N = inst_in.shape[0]
        nt = tNo - int(k_first)
        data = np.empty((N, nt, self.n + 1, zNo), dtype=np.float64)
        snaps = np.empty((N, nt, self.nv2, zNo), dtype=np.float64) if scaled else None
        status = np.empty((N,), dtype=np.int32)
        stats = np.empty((4, N), dtype=np.int32)
        if iv is not None:
            iv = np.ascontiguousarray(iv, dtype=np.float64)
            assert iv.shape == (N, self.nv2, zNo)
        self._check(lib().pbr_solve_n2_host_resume(
            self._ctx, N, zNo, tNo, int(k_first), float(period), inst_in.ctypes.data_as(C.c_void_p),
            iv.ctypes.data_as(C.c_void_p) if iv is not None else None, 1 if restart_every_interval else 0, rtol, atol, max_steps,
            snaps.ctypes.data_as(C.c_void_p) if scaled else None, data.ctypes.data_as(C.c_void_p),
            status.ctypes.data_as(C.c_void_p), stats.ctypes.data_as(C.c_void_p)))
        return {"dataYs": data, "snaps_y": snaps, "status": status, "stats": stats}

    def math_probe(self, x):
        """Device math primitives on a CUDA float64 vector -> [7, n] (rcp, rcp_40bit, rcp_20bit, log, sqrt, exp(-x), x^(-1/3))."""
        import torch
        assert x.is_cuda and x.dtype == torch.float64 and x.is_contiguous()
        out = torch.empty((7, x.numel()), dtype=torch.float64, device=x.device)
        self._check(lib().pbr_math_probe(self._ctx, x.numel(), x.data_ptr(), out.data_ptr(), self._stream()))
        return out

    def fp64_peak(self, iters=8192):
        t = C.c_double()
        self._check(lib().pbr_fp64_peak(self._ctx, iters, C.byref(t)))
        return t.value
