// simt_driver.cpp — TEST HARNESS ONLY (never loaded by pyremot_b200).
// Runs the block-per-reactor N2 solver kernels (pbr_k4b_solve_n2 serial march, pbr_k4bs_solve_n2 scan march,
// pbr_k4c_solve_n2 history in a global slab) on the CPU through the fiber-based SIMT emulator in simt.h: the exact device-code
// text libpbr hands to NVRTC (-DPBR_MODEL_FILE), compiled with -DPBR_HOST -DPBR_HOST_SIMT.
#define PBR_HOST 1
#define PBR_HOST_SIMT 1
#include "simt.h"
double pbr_smem[1 << 18];                       // 2 MB: the block's dynamic shared memory
#include PBR_MODEL_FILE
#include <vector>

extern "C" {

int hs_simt_nx2() { return PBR_NX2; }

// f(x) of the extent-form N2 system at a given iterate through pbr_k4b_eval (one block): x, f [NX2][zNo]
void hs_simt_eval(int zNo, const double* par, const double* x, double* f) {
    const unsigned block = (unsigned)((zNo + 31)/32*32);
    simt::launch(1, block, [&]() {
        double* Asm = pbr_smem; double* Ysm = Asm + PBR_NX2*PBR_NX2*PBR_K4B_ASTRIDE(zNo); double* red = Ysm + PBR_NX2*zNo;
        BlkCtx cx; cx.red = red; cx.tid = threadIdx.x; cx.lane = threadIdx.x & 31; cx.wid = threadIdx.x >> 5; cx.nw = blockDim.x >> 5;
        InstPar p; pbr_load_par(par, 1, 0, p);
        const int z = threadIdx.x; const bool act = z < zNo;
        double y[PBR_NX2], ff[PBR_NX2], sub[PBR_NX2];
        for (int v = 0; v < PBR_NX2; ++v) { y[v] = act ? x[v*zNo + z] : 1.0; ff[v] = 0; sub[v] = 0; }
        pbr_k4b_eval<false>(cx, p, zNo, act, y, 0.0, Ysm, Asm, ff, sub);
        if (act) for (int v = 0; v < PBR_NX2; ++v) f[v*zNo + z] = ff[v];
    });
}

void hs_simt_set_order(int mode, unsigned long long seed) { simt::order_mode = mode; if (seed) simt::order_state = seed; }

// Negative control of the race check: thread t publishes a value and reads its neighbour's WITHOUT a barrier in between
// (racy == 1) or with one (racy == 0).  Under different resume orders the racy version must give different results.
void hs_simt_race_demo(int racy, double* out) {
    simt::launch(1, 64, [&]() {
        const int t = threadIdx.x;
        pbr_smem[t] = 0.0;
        __syncthreads();
        pbr_smem[t] = 1.0 + t;
        if (!racy) __syncthreads();
        out[t] = pbr_smem[(t + 1) % 64];
    });
}

// K3 (the fused N2 RHS kernel) on the emulator: 128-thread blocks owning G reactors each
void hs_simt_rhs_n2(long N, int zNo, int G, const double* par, const double* Y, double* dY, int staged) {
    simt::launch((unsigned)((N + G - 1)/G), 128, [&]() { pbr_k3_rhs_n2(N, zNo, G, par, Y, dY, staged); });
}

// kernel: 0 = pbr_k4b_solve_n2, 1 = pbr_k4bs_solve_n2, 2 = pbr_k4c_solve_n2, 3 = pbr_k4d_solve_n2.  par [KPAR][N]; snaps [N][tNo][NV2][zNo];
// stats [4][N].  grid < N exercises the kernels' loop over instances.  iv (or NULL) [N][NV2][zNo], k_first, flags: resume arguments
// (snaps then holds tNo - k_first interval ends per instance).
int hs_simt_solve_n2(int kernel, long N, int zNo, int tNo, double period, const double* par, double rtol, double atol,
                     int max_steps, int grid, double* snaps, int* status, int* stats, const double* iv, int k_first, int flags) {
    const unsigned block = (unsigned)((zNo + 31)/32*32);
    if ((size_t)(7*PBR_NX2 + PBR_NX2*PBR_NX2 + 2*PBR_NX2 + PBR_NV2)*zNo + PBR_NX2*PBR_NX2 + 800 > sizeof(pbr_smem)/sizeof(double)) return -1;
    if (grid < 1) grid = 1;
    std::vector<double> gwork((size_t)grid*PBR_DROWS*PBR_NX2*zNo);
    double* gw = gwork.data();
    simt::launch((unsigned)grid, block, [&]() {
        if (kernel == 0) pbr_k4b_solve_n2(N, zNo, tNo, period, par, rtol, atol, max_steps, snaps, status, stats, iv, k_first, flags);
        else if (kernel == 1) pbr_k4bs_solve_n2(N, zNo, tNo, period, par, rtol, atol, max_steps, snaps, status, stats, iv, k_first, flags);
        else if (kernel == 3) pbr_k4d_solve_n2(N, zNo, tNo, period, par, rtol, atol, max_steps, snaps, status, stats, iv, k_first, flags);
        else pbr_k4c_solve_n2(N, zNo, tNo, period, par, rtol, atol, max_steps, gw, snaps, status, stats, iv, k_first, flags);
    });
    return 0;
}

}
