// hostsim_driver.cpp — TEST HARNESS ONLY (not part of the product, never loaded by pyremot_b200).
// Compiles the exact device-code text that libpbr hands to NVRTC (pbr_model_source()) as plain C++
// (-DPBR_HOST) and drives the per-thread functions in scalar loops, so the physics / Jacobians / BDF logic
// can be checked against the oracle on the CPU-only build container before spending GPU time.
// The model program is #included from the file named by -DPBR_MODEL_FILE.
#define PBR_HOST 1
#include PBR_MODEL_FILE
#include <vector>

extern "C" {

void hs_dims(int* nv1, int* nv2, int* kin, int* kpar) { *nv1 = PBR_NV1; *nv2 = PBR_NV2; *kin = PBR_KIN; *kpar = PBR_KPAR; }

void hs_prolog(long N, const double* in, double* par) { for (long i = 0; i < N; ++i) pbr_prolog_one(in + i*PBR_KIN, par, N, i); }

void hs_rhs_n1(long N, const double* par, const double* y, double* f) {
    NoStore ns;
    for (long i = 0; i < N; ++i) {
        InstPar p; pbr_load_par(par, N, i, p);
        double yy[PBR_NV1], ff[PBR_NV1];
        for (int v = 0; v < PBR_NV1; ++v) yy[v] = y[v*N + i];
        pbr_n1_eval<false>(yy, p, ff, ns);
        for (int v = 0; v < PBR_NV1; ++v) f[v*N + i] = ff[v];
    }
}

struct JS { double* J; long N, i; void operator()(int r, int c, double v) const { J[(long)(r*PBR_NV1 + c)*N + i] = v; } };

void hs_jac_n1(long N, const double* par, const double* y, double* f, double* J) {
    for (long i = 0; i < N; ++i) {
        InstPar p; pbr_load_par(par, N, i, p);
        double yy[PBR_NV1], ff[PBR_NV1];
        for (int v = 0; v < PBR_NV1; ++v) yy[v] = y[v*N + i];
        JS st = {J, N, i};
        pbr_n1_eval<true>(yy, p, ff, st);
        for (int v = 0; v < PBR_NV1; ++v) f[v*N + i] = ff[v];
    }
}

// Y, dY: [N][NV2][zNo]; sequential pressure march exactly as the reference loop (:3892-3979)
struct JS2 { double* J; void operator()(int r, int c, double v) const { J[r*PBR_NV2 + c] = v; } };

void hs_rhs_n2(long N, int zNo, const double* par, const double* Y, double* dY, double* Jd, double* Js) {
#if PBR_MODEL_M9
    // M9: the warp solver's own sweep (one scalar "lane"); with Jd also the FD diagonal blocks at the marched (P, v)
    std::vector<double> work((size_t)PBR_N2_WS_ROWS*zNo), tile(PBR_N2_TILE_ROWS);
    for (long i = 0; i < N; ++i) {
        InstPar p; pbr_load_par(par, N, i, p);
        WarpCtx w; w.lane = 0; w.tile = tile.data();
        N2Ws ws = {work.data(), zNo};
        for (int v = 0; v < PBR_NV2; ++v) for (int z = 0; z < zNo; ++z) ws.vec(0, v)[z] = Y[(i*PBR_NV2 + v)*zNo + z];
        pbr_n2_sweep_eval<false>(w, p, ws, 0.0, false);
        for (int v = 0; v < PBR_NV2; ++v) for (int z = 0; z < zNo; ++z) dY[(i*PBR_NV2 + v)*zNo + z] = ws.vec(4, v)[z];
        if (Jd) {
            const double dzL = p.zf/(zNo - 1), inv_dzL = 1.0/dzL;
            double Pz = p.Pf, vz = p.vs0;
            for (int z = 0; z < zNo; ++z) {
                double yc[PBR_NV2], yb[PBR_NV2], f[PBR_NV2], sub[PBR_NV2], dP, dv;
                for (int v = 0; v < PBR_NV2; ++v) { yc[v] = Y[(i*PBR_NV2 + v)*zNo + z]; yb[v] = z ? Y[(i*PBR_NV2 + v)*zNo + z - 1] : 0.0; }
                pbr_m9_node(yc, yb, z == 0, Pz, vz, inv_dzL, p, f, dP, dv, sub);
                JS2 st = {Jd + ((i*zNo + z)*PBR_NV2*PBR_NV2)};
                pbr_m9_node_jac(yc, yb, z == 0, Pz, vz, inv_dzL, p, f, st);
                for (int v = 0; v < PBR_NV2; ++v) Js[(i*zNo + z)*PBR_NV2 + v] = sub[v];
                Pz += dzL*dP; vz += dzL*dv;
            }
        }
    }
    return;
#endif
    const double dz = 1.0/(zNo - 1), inv_dz = (double)(zNo - 1);
    for (long i = 0; i < N; ++i) {
        InstPar p; pbr_load_par(par, N, i, p);
        double Pz = p.Pf;
        for (int z = 0; z < zNo; ++z) {
            double yc[PBR_NV2], yb[PBR_NV2], f[PBR_NV2], sub[PBR_NV2];
            for (int v = 0; v < PBR_NV2; ++v) { yc[v] = Y[(i*PBR_NV2 + v)*zNo + z]; yb[v] = z ? Y[(i*PBR_NV2 + v)*zNo + z - 1] : 0.0; }
            double slope;
            if (Jd) {
                JS2 st = {Jd + ((i*zNo + z)*PBR_NV2*PBR_NV2)};
                slope = pbr_n2_node<true>(yc, yb, z == 0, Pz, inv_dz, p, f, st, sub);
                for (int v = 0; v < PBR_NV2; ++v) Js[(i*zNo + z)*PBR_NV2 + v] = sub[v];
            } else {
                NoStore ns;
                slope = pbr_n2_node<false>(yc, yb, z == 0, Pz, inv_dz, p, f, ns, (double*)0);
            }
            for (int v = 0; v < PBR_NV2; ++v) dY[(i*PBR_NV2 + v)*zNo + z] = f[v];
            Pz = slope*dz + Pz;
        }
    }
}

// one N2 node in the extent coordinates of K4b/K4c: x, xb [NX2]; outputs y_full, yb_full [NV2], f [NX2], J [NX2*NX2], sub [NX2]
struct JSX2 { double* J; void operator()(int r, int c, double v) const { J[r*PBR_NX2 + c] = v; } };
int hs_nx2() { return PBR_NX2; }
double hs_n2x_node(const double* par, const double* x, const double* xb, int first, double Pz, double inv_dz,
                   double* yfull, double* ybfull, double* f, double* J, double* sub) {
    InstPar p; pbr_load_par(par, 1, 0, p);
    pbr_n2x_expand(x, p, yfull);
    pbr_n2x_expand(xb, p, ybfull);
    JSX2 st = {J};
    return pbr_n2x_node<true>(x, xb, first != 0, Pz, inv_dz, p, f, st, sub);
}
double hs_n2_node(const double* par, const double* y, const double* yb, int first, double Pz, double inv_dz,
                  double* f, double* J, double* sub) {
    InstPar p; pbr_load_par(par, 1, 0, p);
    JS2 st = {J};
    return pbr_n2_node<true>(y, yb, first != 0, Pz, inv_dz, p, f, st, sub);
}

struct Sink { double* prof; long N, i; int n_out; const InstPar* par; void operator()(int k, const double* x) const {
    double y[PBR_NV1];
    pbr_n1x_expand(x, *par, y);
    for (int v = 0; v < PBR_NV1; ++v) prof[((long)k*PBR_NV1 + v)*N + i] = y[v]; } };

// N1 in the integrated (extent) coordinates: x [NX][N] -> y_full [NV1][N], f_x [NX][N], J_x [NX*NX][N]
struct JSX { double* J; long N, i; void operator()(int r, int c, double v) const { J[(long)(r*PBR_NX + c)*N + i] = v; } };
int hs_nx() { return PBR_NX; }
void hs_n1x(long N, const double* par, const double* x, double* yfull, double* f, double* J) {
    for (long i = 0; i < N; ++i) {
        InstPar p; pbr_load_par(par, N, i, p);
        double xx[PBR_NX], ff[PBR_NX], yy[PBR_NV1];
        for (int v = 0; v < PBR_NX; ++v) xx[v] = x[v*N + i];
        pbr_n1x_expand(xx, p, yy);
        for (int v = 0; v < PBR_NV1; ++v) yfull[v*N + i] = yy[v];
        JSX st = {J, N, i};
        pbr_n1x_eval<true>(xx, p, ff, st);
        for (int v = 0; v < PBR_NX; ++v) f[v*N + i] = ff[v];
    }
}

// prof: [n_out][NV1][N] scaled; stats [5][N]
void hs_solve_n1(long N, const double* par, double rtol, double atol, int n_out, int max_steps, double* prof, int* status, int* stats) {
    std::vector<double> work(PBR_N1_WORK_HOST);
    for (long i = 0; i < N; ++i) {
        InstPar p; pbr_load_par(par, N, i, p);
        double y0[PBR_NX];
        for (int k = 0; k < PBR_NX; ++k) y0[k] = 0.0;
#if !PBR_N1_REDUCED
        for (int k = 0; k < PBR_NC; ++k) y0[k] = p.c0[k];
#endif
        y0[PBR_NX - (PBR_ISO ? 1 : 2) - (PBR_MODEL_M8 ? 1 : 0)] = 1.0;
        if (PBR_MODEL_M8) y0[PBR_NX - 1] = 1.0;
        Strided Dw = {work.data(), 1};
        Strided LUw = {work.data() + PBR_DROWS*PBR_NV1, 1};
        Sink sk = {prof, N, i, n_out, &p};
        N1Stats st;
        status[i] = pbr_n1_solve<0>(p, y0, rtol, atol, n_out, max_steps, Dw, LUw, sk, st);
        stats[0*N + i] = st.nsteps; stats[1*N + i] = st.nrej; stats[2*N + i] = st.nfev; stats[3*N + i] = st.nlu; stats[4*N + i] = st.nfail;
    }
}

#if PBR_NP > 0
// forward sensitivities: kp[N] parameter index per instance; outlet, sens: [NV1][N] scaled (d [c, p, theta](z = 1) / d par_kp)
void hs_solve_n1_sens(long N, const double* par, const int* kp, double rtol, double atol, int max_steps, double* outlet, double* sens,
                      int* status, int* stats) {
    std::vector<double> work(2*PBR_DROWS*PBR_NV1 + PBR_NV1*PBR_NV1);
    std::vector<double> prof((size_t)2*PBR_NV1);
    for (long i = 0; i < N; ++i) {
        InstPar p; pbr_load_par(par, N, i, p);
        double y0[PBR_NX];
        for (int k = 0; k < PBR_NX; ++k) y0[k] = 0.0;
#if !PBR_N1_REDUCED
        for (int k = 0; k < PBR_NC; ++k) y0[k] = p.c0[k];
#endif
        y0[PBR_NX - (PBR_ISO ? 1 : 2)] = 1.0;
        Strided Dw = {work.data(), 1};
        Strided LUw = {work.data() + 2*PBR_DROWS*PBR_NV1, 1};
        Sink sk = {prof.data(), 1, 0, 2, &p};
        N1Stats st;
        double so[PBR_NV1];
        for (int v = 0; v < PBR_NV1; ++v) so[v] = NAN;
        status[i] = pbr_n1_solve<PBR_NX>(p, y0, rtol, atol, 2, max_steps, Dw, LUw, sk, st, true, kp[i], so);
        for (int v = 0; v < PBR_NV1; ++v) { outlet[v*N + i] = prof[(size_t)(1*PBR_NV1 + v)]; sens[v*N + i] = so[v]; }
        stats[0*N + i] = st.nsteps; stats[1*N + i] = st.nrej; stats[2*N + i] = st.nfev; stats[3*N + i] = st.nlu; stats[4*N + i] = st.nfail;
    }
}
#endif

// snaps: [N][tNo][NV2][zNo] scaled; stats [5][N]
void hs_solve_n2(long N, int zNo, int tNo, double period, const double* par, double rtol, double atol, int max_steps,
                 double* snaps, int* status, int* stats) {
    std::vector<double> work((size_t)PBR_N2_WS_ROWS*zNo);
    for (long i = 0; i < N; ++i) {
        InstPar p; pbr_load_par(par, N, i, p);
        std::vector<double> tile(PBR_N2_TILE_ROWS);
        WarpCtx w; w.lane = 0; w.tile = tile.data();
        N2Ws ws = {work.data(), zNo};
        N2Stats st;
        status[i] = pbr_n2_solve(w, p, ws, tNo, period, rtol, atol, max_steps, snaps + i*(long)tNo*PBR_NV2*zNo, st);
        stats[0*N + i] = st.nsteps; stats[1*N + i] = st.nrej; stats[2*N + i] = st.nfev; stats[3*N + i] = st.nlu; stats[4*N + i] = st.nfail;
    }
}

void hs_lu_test(int n_unused, double* A /*NV1 x NV1 row-major, overwritten*/, double* b) {
    Strided a = {A, 1};
    unsigned long long piv;
    pbr_lu_factor<PBR_NV1>(a, piv);
    pbr_lu_solve<PBR_NV1>(a, piv, b);
}

void hs_invert_test(double* A /*NV1 x NV1 row-major, overwritten by its inverse*/) {
    Strided a = {A, 1};
    pbr_invert<PBR_NV1>(a);
}

void hs_change_D(double* D, int order, double factor) { Strided d = {D, 1}; pbr_change_D<PBR_NV1>(d, order, factor); }

}
