// Dependent-issue latency of the instructions the solver kernels chain (one warp, one block, clock64 around 512 dependent ops).
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/lat_probe tools/lat_probe.cu && ./tools/lat_probe
#include <cstdio>
#include <cuda_runtime.h>

#define N 512
template <int OP> __global__ void probe(double* out, long long* cyc, double a, double b, int idx) {
    __shared__ double sm[64];
    sm[threadIdx.x] = a + threadIdx.x*1e-9;
    sm[threadIdx.x + 32] = b;
    __syncthreads();
    double x = a + threadIdx.x*1e-12, y = b;
    int j = idx;
    const long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (OP == 0) x = fma(x, y, a);                                 // DFMA
        if (OP == 1) x = x + y;                                        // DADD
        if (OP == 2) x = x*y;                                          // DMUL
        if (OP == 3) x = __shfl_sync(0xffffffffu, x, j);               // 2 x SHFL.IDX (a double)
        if (OP == 4) { j = (int)sm[j & 31] & 31; }                      // LDS.64 + F2I (pointer chase through shared memory)
        if (OP == 5) { float f = (float)x; f = fmaf(f, 1.0001f, 0.5f); x = (double)f; }   // F2F + FFMA + F2F
        if (OP == 6) x = fma(x, y, a) + __shfl_sync(0xffffffffu, x, j); // DFMA || SHFL then DADD
        if (OP == 7) { float f = __int_as_float(j); f = fmaf(f, 1.0001f, 0.5f); j = __float_as_int(f); }   // FFMA
    }
    const long long t1 = clock64();
    out[threadIdx.x] = x + j;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 32*8); cudaMalloc(&cyc, 8);
    const char* names[] = {"DFMA", "DADD", "DMUL", "SHFL.IDX x2 (double)", "LDS.64 + F2I chase", "F2F.F32.F64 + FFMA + F2F.F64.F32", "DFMA || SHFL -> DADD", "FFMA"};
    for (int op = 0; op < 8; ++op) {
        for (int rep = 0; rep < 2; ++rep) {
            switch (op) {
                case 0: probe<0><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, 3); break;
                case 1: probe<1><<<1, 32>>>(out, cyc, 1.0000001, 1e-9, 3); break;
                case 2: probe<2><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, 3); break;
                case 3: probe<3><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, 3); break;
                case 4: probe<4><<<1, 32>>>(out, cyc, 5.0, 0.9999999, 3); break;
                case 5: probe<5><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, 3); break;
                case 6: probe<6><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, 3); break;
                case 7: probe<7><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, 3); break;
            }
            cudaDeviceSynchronize();
        }
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-36s %7.1f cycles per dependent op\n", names[op], (double)h/N);
    }
    return 0;
}
