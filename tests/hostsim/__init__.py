"""Scalar host emulation of the device code — TEST HARNESS ONLY (see hostsim_driver.cpp).

Builds, with g++ -DPBR_HOST, the same program text libpbr gives NVRTC, so kernel numerics can be checked
against the oracle without a GPU.  Never imported by the product package.
"""
import ctypes as C
import hashlib
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(tempfile.gettempdir(), "pbr_hostsim")


def _build(src, driver, extra=()):
    """g++ build of `driver` around the model program text `src`; cached in the temp directory under a content hash."""
    text = "".join(open(os.path.join(_HERE, f)).read() for f in (driver,) + tuple(extra))
    key = hashlib.sha256((src + text).encode()).hexdigest()[:16]
    os.makedirs(_BUILD, exist_ok=True)
    so = os.path.join(_BUILD, f"{os.path.splitext(driver)[0]}_{key}.so")
    if not os.path.exists(so):
        cu = os.path.join(_BUILD, f"model_{key}.h")
        with open(cu, "w") as f:
            f.write(src)
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", f"-I{_HERE}",
                        f'-DPBR_MODEL_FILE="{cu}"', "-DPBR_N1_WORK_HOST=(PBR_DROWS*PBR_NV1+PBR_NV1*PBR_NV1)",
                        os.path.join(_HERE, driver), "-o", so + ".tmp"], check=True)
        os.replace(so + ".tmp", so)
    return C.CDLL(so)


class SimtSim:
    """The block-per-reactor N2 solver kernels run through the fiber-based SIMT emulator (simt.h, simt_driver.cpp)."""
    KERNELS = {"pbr_k4b_solve_n2": 0, "pbr_k4bs_solve_n2": 1, "pbr_k4c_solve_n2": 2, "pbr_k4d_solve_n2": 3}

    def __init__(self, model):
        self.L = _build(model.source, "simt_driver.cpp", extra=("simt.h",))
        self.nv2 = model.nv2

    def set_order(self, mode, seed=0):
        """Thread resume order inside a scheduling round: 0 ascending, 1 descending, 2 random per round (race check)."""
        self.L.hs_simt_set_order(C.c_int(mode), C.c_ulonglong(seed))

    def race_demo(self, racy):
        out = np.zeros(64)
        self.L.hs_simt_race_demo(C.c_int(int(racy)), HostSim._p(out))
        return out

    def rhs_n2(self, par, Y, G, staged=1):
        """K3 through the emulator; G reactors per 128-thread block; staged: state tile in shared memory."""
        par = np.ascontiguousarray(par, dtype=np.float64)
        Y = np.ascontiguousarray(Y, dtype=np.float64)
        N, nv, zNo = Y.shape
        dY = np.zeros_like(Y)
        p = HostSim._p
        self.L.hs_simt_rhs_n2(C.c_long(N), C.c_int(zNo), C.c_int(G), p(par), p(Y), p(dY), C.c_int(staged))
        return dY

    def solve_n2(self, kernel, par, zNo, tNo, period, rtol=1e-9, atol=1e-13, max_steps=50000, grid=None, iv=None, k_first=0,
                 flags=0):
        par = np.ascontiguousarray(par, dtype=np.float64)
        N = par.shape[1]
        snaps = np.zeros((N, tNo - k_first, self.nv2, zNo))
        if iv is not None:
            iv = np.ascontiguousarray(iv, dtype=np.float64)
            assert iv.shape == (N, self.nv2, zNo)
        status = np.zeros(N, dtype=np.int32)
        stats = np.zeros((4, N), dtype=np.int32)
        p = HostSim._p
        rc = self.L.hs_simt_solve_n2(self.KERNELS[kernel], C.c_long(N), C.c_int(zNo), C.c_int(tNo), C.c_double(period), p(par),
                                     C.c_double(rtol), C.c_double(atol), C.c_int(max_steps), C.c_int(grid or N), p(snaps),
                                     p(status), p(stats), p(iv) if iv is not None else None, C.c_int(k_first), C.c_int(flags))
        assert rc == 0, rc
        return snaps, status, stats


class HostSim:
    def __init__(self, model):
        src = model.source
        key = hashlib.sha256((src + open(os.path.join(_HERE, "hostsim_driver.cpp")).read()).encode()).hexdigest()[:16]
        os.makedirs(_BUILD, exist_ok=True)
        so = os.path.join(_BUILD, f"hs_{key}.so")
        if not os.path.exists(so):
            cu = os.path.join(_BUILD, f"hs_{key}.h")
            with open(cu, "w") as f:
                f.write(src)
            subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off",
                            f'-DPBR_MODEL_FILE="{cu}"', "-DPBR_N1_WORK_HOST=(PBR_DROWS*PBR_NV1+PBR_NV1*PBR_NV1)",
                            os.path.join(_HERE, "hostsim_driver.cpp"), "-o", so + ".tmp"], check=True)
            os.replace(so + ".tmp", so)
        self.L = C.CDLL(so)
        d = [C.c_int() for _ in range(4)]
        self.L.hs_dims(*[C.byref(x) for x in d])
        self.nv1, self.nv2, self.kin, self.kpar = [x.value for x in d]

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(C.c_void_p)

    def prolog(self, inst_in):
        inst_in = np.ascontiguousarray(inst_in, dtype=np.float64)
        N = inst_in.shape[0]
        par = np.zeros((self.kpar, N))
        self.L.hs_prolog(C.c_long(N), self._p(inst_in), self._p(par))
        return par

    def rhs_n1(self, par, y):
        y = np.ascontiguousarray(y, dtype=np.float64)
        f = np.zeros_like(y)
        self.L.hs_rhs_n1(C.c_long(par.shape[1]), self._p(par), self._p(y), self._p(f))
        return f

    def jac_n1(self, par, y):
        y = np.ascontiguousarray(y, dtype=np.float64)
        N = par.shape[1]
        f = np.zeros_like(y)
        J = np.zeros((self.nv1, self.nv1, N))
        self.L.hs_jac_n1(C.c_long(N), self._p(par), self._p(y), self._p(f), self._p(J))
        return f, J

    def n1x(self, par, x):
        """N1 in the integrated coordinates: x [nx, N] -> (y_full [nv1, N], f_x [nx, N], J_x [nx, nx, N])."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        nx, N = x.shape
        assert nx == self.L.hs_nx()
        y = np.zeros((self.nv1, N)); f = np.zeros_like(x); J = np.zeros((nx, nx, N))
        self.L.hs_n1x(C.c_long(N), self._p(par), self._p(x), self._p(y), self._p(f), self._p(J))
        return y, f, J

    def rhs_n2(self, par, Y, jac=False):
        Y = np.ascontiguousarray(Y, dtype=np.float64)
        N, nv, zNo = Y.shape
        dY = np.zeros_like(Y)
        if jac:
            Jd = np.zeros((N, zNo, nv, nv))
            Js = np.zeros((N, zNo, nv))
            self.L.hs_rhs_n2(C.c_long(N), C.c_int(zNo), self._p(par), self._p(Y), self._p(dY), self._p(Jd), self._p(Js))
            return dY, Jd, Js
        self.L.hs_rhs_n2(C.c_long(N), C.c_int(zNo), self._p(par), self._p(Y), self._p(dY), None, None)
        return dY

    def solve_n1(self, par, rtol=1e-8, atol=1e-12, n_out=101, max_steps=1000000):
        N = par.shape[1]
        prof = np.zeros((n_out, self.nv1, N))
        status = np.zeros(N, dtype=np.int32)
        stats = np.zeros((5, N), dtype=np.int32)
        self.L.hs_solve_n1(C.c_long(N), self._p(par), C.c_double(rtol), C.c_double(atol), C.c_int(n_out),
                           C.c_int(max_steps), self._p(prof), self._p(status), self._p(stats))
        return prof, status, stats

    def solve_n1_sens(self, par, kp, rtol=1e-9, atol=1e-13, max_steps=1000000):
        """Forward sensitivities of the outlet w.r.t. runtime rate parameter kp[i]: (outlet [nv1, N], sens [nv1, N]) scaled."""
        N = par.shape[1]
        kp = np.ascontiguousarray(kp, dtype=np.int32)
        outlet = np.zeros((self.nv1, N)); sens = np.zeros((self.nv1, N))
        status = np.zeros(N, dtype=np.int32)
        stats = np.zeros((5, N), dtype=np.int32)
        self.L.hs_solve_n1_sens(C.c_long(N), self._p(par), self._p(kp), C.c_double(rtol), C.c_double(atol), C.c_int(max_steps),
                                self._p(outlet), self._p(sens), self._p(status), self._p(stats))
        return outlet, sens, status, stats

    def solve_n2(self, par, zNo, tNo, period, rtol=1e-8, atol=1e-12, max_steps=1000000):
        N = par.shape[1]
        snaps = np.zeros((N, tNo, self.nv2, zNo))
        status = np.zeros(N, dtype=np.int32)
        stats = np.zeros((5, N), dtype=np.int32)
        self.L.hs_solve_n2(C.c_long(N), C.c_int(zNo), C.c_int(tNo), C.c_double(period), self._p(par), C.c_double(rtol),
                           C.c_double(atol), C.c_int(max_steps), self._p(snaps), self._p(status), self._p(stats))
        return snaps, status, stats
