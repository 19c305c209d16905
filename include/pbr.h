/* pbr.h — C ABI of libpbr.so, the B200-native replacement of PyREMOT's N1 / N2 hot path.
 *
 * The reference has no FFI layer: its boundary is the Python call rmtExe(modelInput)
 * (/root/reference/PyREMOT/rmt.py:21-80).  Everything below rmtExe that the path executes
 * (PyREMOT/docs/pbHomoReactor.py runN1 :2694-3015 / runN2 :3319-3704 and the SciPy solve_ivp
 * calls at :2931, :3609) is replaced by the entry points declared here; the thin Python host
 * (pyremot_b200/engine.py, ctypes) is their only caller.  INTEGRATION.md shows the binding a
 * PyREMOT maintainer would add.
 *
 * Conventions: plain C types only; every function returns 0 on success and a negative pbr_status on
 * failure, with a message available from pbr_last_error().  A pbr_ctx is bound to one CUDA device and
 * is not thread-safe (one ctx per host thread).  `stream` is a cudaStream_t passed as void* (NULL =
 * legacy default stream).  "dev" pointers are device memory owned by the caller; "host" pointers are
 * host memory (pinned or pageable).  All floating point data is IEEE binary64.
 *
 * Per-instance tables
 *   inst_in  [N][k_in]   AoS rows: T0, P0, VoFlRa, ReInDi, ReLe, PaDi, BeVoFr, U, Tm, C0[n_comp], rate_par[n_par]
 *                        (the fields runN1/runN2 read from modelInput: pbHomoReactor.py:2716-2782, :3341-3411)
 *   inst_par [k_par][N]  SoA derived constants written by pbr_prolog (K0)
 * State layouts
 *   N1  y  [nv1][N]              instance-minor; nv1 = n_comp+2 (iso-thermal: n_comp+1): c_i/Cmax, P/Pf, (T-Tf)/Tf
 *   N2  Y  [N][nv2][zNo]         the reference's var-major flatten (pbHomoReactor.py:3873); nv2 = n_comp+1 (iso: n_comp)
 */
#ifndef PBR_H
#define PBR_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pbr_ctx pbr_ctx;

typedef enum {
    PBR_SUCCESS = 0,
    PBR_E_INVALID = -1,      /* bad argument / call order */
    PBR_E_CUDA = -2,         /* CUDA runtime or driver error */
    PBR_E_NVRTC = -3,        /* rate source failed to compile (see log) */
    PBR_E_NODEVICE = -4,     /* compute entry point called on a compile-only context */
    PBR_E_NOMEM = -5
} pbr_status;

/* per-instance integrator status written to `status[]` (0 = solved) */
#define PBR_INST_OK 0
#define PBR_INST_STEP_TOO_SMALL 1
#define PBR_INST_MAX_STEPS 2
#define PBR_INST_NONFINITE 3
#define PBR_INST_SINGULAR 4
#define PBR_INST_BAD_IV 5         /* N2 resume: the saved state is not of the form c0 + S^T x (cannot come from this model) */
/* Rows of instances whose status is not 0 are filled with NaN (outlet, profile, and the N2 snapshots from the failing
 * interval on), so a failed instance can never be mistaken for a result. */

/* device >= 0: bind to that CUDA device (fails loudly if there is none). device < 0: compile-only context
 * (NVRTC needs no GPU) — used by the build check and the CPU-side tests. */
pbr_ctx* pbr_create(int device);
void pbr_destroy(pbr_ctx* ctx);
const char* pbr_last_error(pbr_ctx* ctx);
const char* pbr_version(void);

/* Species / reaction tables of the model (replaces rmtCore.initComponentData / initReaction,
 * PyREMOT/docs/rmtCore.py:129-183, and the property look-ups of rmtThermo.py:16-49, 129-198):
 *   MW[n_comp] g/mol; cp_coef[n_comp][4] (a + bT + cT^2 + dT^3, J/mol/K); vis_coef[n_comp][4] (eq1 A,B,C,D);
 *   S[n_rxn][n_comp] stoichiometric matrix; dH25[n_rxn] J/mol; dcp_coef[n_rxn][4] = sum_k nu_jk cp_coef[k]
 *   (over ALL species named in reaction j, also those absent from the component list — rmtThermo.py:258-312). */
int pbr_set_species(pbr_ctx* ctx, int n_comp, int n_rxn, const double* MW, const double* cp_coef,
                    const double* vis_coef, const double* S, const double* dH25, const double* dcp_coef,
                    int iso_thermal);

/* Model variant, to be chosen before pbr_load_rates: 0 (default) = the non-dimensional models N1 / N2; 1 = the dimensional
 * steady-state model "M7" (rmtCore.py:300-309 -> pbReactor.py runM3 :1170, modelEquationM3 :1371-1577); 2 = "M8"
 * (rmtCore.py:311-320 -> runM4 :1580, modelEquationM4 :1778-1995: the superficial velocity is a fourth kind of unknown) -- both
 * run through the N1 entry points.  Variant 1 appends two columns to inst_in: mixture viscosity [Pa.s] (feed['mixture-viscosity'])
 * and the heat-exchange area per volume (external-heat['EfHeTrAr']), both inputs of those models; variant 2 a third one, the inlet
 * superficial velocity [m/s] (feed['superficial-velocity']), and its state / profiles have one more row (nv1 = n_comp + 3: v).
 * 3 = the dynamic dimensional model "M9" (rmtCore.py:322-331 -> runM5 :1997, modelEquationM5 :2296-2666: pressure AND velocity
 * marched along the bed inside every evaluation, catalyst heat capacity in the energy balance); it runs through the N2 entry
 * points (pbr_rhs_n2, pbr_solve_n2*; state y = [C_i/Cmax ; (T - T0)/T0] per node) and appends five
 * columns to inst_in: the three of variant 2, then reactor['CaDe'] [kg/m3] and reactor['CaSpHeCa']. */
int pbr_set_model_variant(pbr_ctx* ctx, int variant);

/* Compile the model: `cuda_src` defines PBR_NP, pbr_rates() and pbr_rates_jac() (output of the rate-expression
 * compiler; replaces reactionRateExe, PyREMOT/docs/rmtReaction.py:11-61).  NVRTC -> sm_100a cubin, cached in
 * `cache_dir` (may be NULL) under a hash of the full program text.  `*log` (may be NULL) receives the NVRTC log. */
int pbr_load_rates(pbr_ctx* ctx, const char* cuda_src, int n_par, const char* cache_dir, const char** log);

int pbr_model_dims(pbr_ctx* ctx, int* n_comp, int* n_rxn, int* n_par, int* nv1, int* nv2, int* k_in, int* k_par);
/* How K2 integrates N1: n_vars = number of integrated variables (n_rxn+2 in reaction-extent coordinates when
 * n_rxn < n_comp, else nv1), threads per block and dynamic shared memory per block.  Informational (roofline counts). */
int pbr_n1_layout(pbr_ctx* ctx, int* n_vars, int* block_threads, int* smem_bytes);
const char* pbr_model_source(pbr_ctx* ctx);          /* full program text handed to NVRTC */
int pbr_model_cubin(pbr_ctx* ctx, const void** data, unsigned long* size);
int pbr_kernel_attr(pbr_ctx* ctx, const char* kernel, int* regs, int* static_smem, int* local_bytes);

/* K0 — prolog of runN1 / runN2 (pbHomoReactor.py:2744-2862, :3370-3497) */
int pbr_prolog(pbr_ctx* ctx, long N, const double* inst_in_dev, double* inst_par_dev, void* stream);

/* RHS callbacks modelEquationN1 (:3017-3314) / modelEquationN2 (:3706-4134), batched */
int pbr_rhs_n1(pbr_ctx* ctx, long N, const double* inst_par_dev, const double* y_dev, double* f_dev, void* stream);
int pbr_jac_n1(pbr_ctx* ctx, long N, const double* inst_par_dev, const double* y_dev, double* f_dev,
               double* J_dev /* [nv1*nv1][N], row-major (i,k) = d f_i / d y_k */, void* stream);
int pbr_rhs_n2(pbr_ctx* ctx, long N, int zNo, const double* inst_par_dev, const double* Y_dev, double* dYdt_dev,
               void* stream);

/* K2 — the solve_ivp call of runN1 (:2931) with t_eval = linspace(0,1,n_out) (:2862): variable-order BDF.
 *   prof_y_dev   [n_out][nv1][N] scaled state at the output points, or NULL
 *   outlet_y_dev [nv1][N]        scaled state at z = 1
 *   status_dev   [N] int         PBR_INST_*
 *   stats_dev    [4][N] int      steps, rejected+failed, RHS evals, LU factorisations; or NULL */
int pbr_solve_n1(pbr_ctx* ctx, long N, const double* inst_par_dev, double rtol, double atol, int n_out,
                 int max_steps, double* prof_y_dev, double* outlet_y_dev, int* status_dev, int* stats_dev,
                 void* stream);

/* K2s — K2 with forward sensitivities (BASELINE config 5: parameter-estimation Jacobians without finite differences).
 * Instance i integrates its reactor together with s = d x / d par_k, k = kp_dev[i] (index into the model's runtime rate
 * parameters, the `params=` of the rate compiler); the sensitivity equations use the same factorised Newton matrix as the
 * state (staggered direct method) and take no part in the error test.
 *   outlet_y_dev, sens_y_dev [nv1][N]: scaled state at z = 1 and d(scaled state)/d par_k.  Models without runtime
 *   parameters return PBR_E_INVALID. */
int pbr_solve_n1_sens(pbr_ctx* ctx, long N, const double* inst_par_dev, const int* kp_dev, double rtol, double atol,
                      int max_steps, double* outlet_y_dev, double* sens_y_dev, int* status_dev, int* stats_dev, void* stream);
/* K5s — d [x_1..x_n, P, T](z = 1) / d par_k from the scaled outlet state and its sensitivity: out_dev [N][nv1] */
int pbr_unscale_sens_n1(pbr_ctx* ctx, long N, const double* inst_par_dev, const double* outlet_y_dev,
                        const double* sens_y_dev, double* out_dev, void* stream);

/* K5 — sortResult4 + mole fractions (solResultAnalysis.py:191-249, pbHomoReactor.py:2973-2983):
 * scaled y [K][nv1][N] -> dataYs [N][nv1][K] = rows x_1..x_n, P [Pa], T [K] */
int pbr_unscale_n1(pbr_ctx* ctx, long N, int K, const double* inst_par_dev, const double* y_dev, double* out_dev,
                   void* stream);

/* Whole N1 path with HOST buffers (what rmtExe does per reactor, batched): H2D of inst_in, K0, K2, K5, D2H.
 *   prof_host   [N][nv1][n_out] dataYs per instance, or NULL;  outlet_host [N][nv1] = dataYs[:, -1]
 *   status_host [N];  stats_host [4][N] or NULL */
int pbr_solve_n1_host(pbr_ctx* ctx, long N, const double* inst_in_host, double rtol, double atol, int n_out,
                      int max_steps, double* prof_host, double* outlet_host, int* status_host, int* stats_host);

/* K4 — the interval loop of runN2 (:3589-3685): tNo intervals of length period/tNo, a snapshot at every interval end
 * (the reference restarts solve_ivp (:3609) from the previous end state at each of them; here the BDF state is carried
 * across the snapshot times, which land exactly on step ends -- same converged snapshots, ~10 % fewer steps);
 * variable-order BDF, block lower-bidiagonal Newton matrix (per-node blocks + forward substitution / affine-map scan
 * along the bed).  One block per reactor / one thread per node up to 512 nodes (state in shared memory), one warp per
 * reactor above.  IV = feed-filled bed (:3483-3497).
 *   snaps_y_dev [N][tNo][nv2][zNo] scaled state at the interval ends;  status_dev [N];  stats_dev [4][N] or NULL.
 * Scratch ((13*nv2 + 2*nv2^2 + 1)*zNo doubles per resident warp) is owned by the context. */
int pbr_solve_n2(pbr_ctx* ctx, long N, int zNo, int tNo, double period, const double* inst_par_dev, double rtol,
                 double atol, int max_steps, double* snaps_y_dev, int* status_dev, int* stats_dev, void* stream);

/* The same with a saved state: integrate intervals k_first .. tNo-1 of linspace(0, period, tNo+1) starting from
 * iv_y_dev [N][nv2][zNo] (scaled, e.g. snapshot k_first-1 of an earlier run) -- what the reference does at every interval,
 * IV = dataYs[:, -1] (pbHomoReactor.py:3685).  iv_y_dev == NULL with k_first == 0 is pbr_solve_n2.  snaps_y_dev holds the
 * tNo - k_first remaining interval ends: [N][tNo-k_first][nv2][zNo].  flags bit 0: restart the integrator (order 1, fresh
 * initial step) at every interval exactly as the reference's per-interval solve_ivp calls do. */
int pbr_solve_n2_resume(pbr_ctx* ctx, long N, int zNo, int tNo, int k_first, double period, const double* inst_par_dev,
                        const double* iv_y_dev, int flags, double rtol, double atol, int max_steps, double* snaps_y_dev,
                        int* status_dev, int* stats_dev, void* stream);

/* K5b — sortResult5 + mole fractions (solResultAnalysis.py:252-301, pbHomoReactor.py:3655-3661):
 * scaled snapshots [N][tNo][nv2][zNo] -> dataYs [N][tNo][n_comp+1][zNo] = rows x_1..x_n, T [K] */
int pbr_unscale_n2(pbr_ctx* ctx, long N, int tNo, int zNo, const double* inst_par_dev, const double* snaps_y_dev,
                   double* out_dev, void* stream);

/* Whole N2 path with HOST buffers: H2D of inst_in, K0, K4, K5b, D2H.
 *   snaps_y_host [N][tNo][nv2][zNo] or NULL;  data_host [N][tNo][n_comp+1][zNo];  status_host [N];  stats_host [4][N] or NULL */
int pbr_solve_n2_host(pbr_ctx* ctx, long N, int zNo, int tNo, double period, const double* inst_in_host, double rtol,
                      double atol, int max_steps, double* snaps_y_host, double* data_host, int* status_host,
                      int* stats_host);

int pbr_solve_n2_host_resume(pbr_ctx* ctx, long N, int zNo, int tNo, int k_first, double period, const double* inst_in_host,
                             const double* iv_y_host, int flags, double rtol, double atol, int max_steps, double* snaps_y_host,
                             double* data_host, int* status_host, int* stats_host);

/* Device math primitives on x_dev[n] -> out_dev[7][n]: 1/x at full, ~40-bit and ~20-bit precision, log, sqrt, exp(-x),
 * x^(-1/3) through single-precision exp2/log2 (the reduced ones only steer the integrators; test hook). */
int pbr_math_probe(pbr_ctx* ctx, long n, const double* x_dev, double* out_dev, void* stream);

/* FP64-pipe peak probe (DFMA streams) — roofline denominator the driver's MEASURED_PEAKS.json lacks. */
int pbr_fp64_peak(pbr_ctx* ctx, int iters, double* tflops);

/* number of kernel launches issued through this context since creation */
long pbr_launch_count(pbr_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* PBR_H */
