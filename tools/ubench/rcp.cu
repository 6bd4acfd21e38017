// Dependent-chain latency of reciprocal variants for the LDL^T pivot (one warp).
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
__device__ __forceinline__ double rcp_a(double x) { double y = (double)__frcp_rn((float)x); y = y * (2.0 - x * y); return y * (2.0 - x * y); }
__device__ __forceinline__ double rcp_b(double x) { float yf; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(yf) : "f"((float)x)); double y = (double)yf; double e = fma(-x, y, 1.0); y = fma(y, e, y); e = fma(-x, y, 1.0); return fma(y, e, y); }
__device__ __forceinline__ double rcp_c(double x) { double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); double e = fma(-x, y, 1.0); y = fma(y, e, y); e = fma(-x, y, 1.0); y = fma(y, e, y); e = fma(-x, y, 1.0); return fma(y, e, y); }
__device__ __forceinline__ double rcp_c2(double x) { double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); double e = fma(-x, y, 1.0); y = fma(y, e, y); e = fma(-x, y, 1.0); return fma(y, e, y); }
__device__ __forceinline__ float rcpf_a(float x) { return 1.0f / x; }
__device__ __forceinline__ float rcpf_b(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); float e = fmaf(-x, y, 1.0f); return fmaf(y, e, y); }
__device__ __forceinline__ float rcpf_c(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int OP>
__global__ void lat(double* out, long long* cyc, double seed, double* err) {
    double x = seed + threadIdx.x * 1e-3; float xf = (float)x;
    double maxerr = 0;
    for (int i = 0; i < 64; ++i) { double v = 0.37 + i * 1.913 + threadIdx.x * 0.01; double r = OP == 0 ? rcp_a(v) : OP == 1 ? rcp_b(v) : OP == 2 ? rcp_c(v) : OP == 3 ? rcp_c2(v) : OP == 4 ? (double)rcpf_a((float)v) : OP == 5 ? (double)rcpf_b((float)v) : (double)rcpf_c((float)v); double e = fabs(r * v - 1.0); if (e > maxerr) maxerr = e; }
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) {
        if (OP == 0) x = rcp_a(x) + 1.5;
        if (OP == 1) x = rcp_b(x) + 1.5;
        if (OP == 2) x = rcp_c(x) + 1.5;
        if (OP == 3) x = rcp_c2(x) + 1.5;
        if (OP == 4) xf = rcpf_a(xf) + 1.5f;
        if (OP == 5) xf = rcpf_b(xf) + 1.5f;
        if (OP == 6) xf = rcpf_c(xf) + 1.5f;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + xf;
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; err[0] = maxerr; }
}
template <int OP> void run(const char* name) {
    double* out; long long* cyc; double* err; cudaMalloc(&out, 256 * 8); cudaMalloc(&cyc, 8); cudaMalloc(&err, 8);
    lat<OP><<<1, 32>>>(out, cyc, 1.0000001, err); cudaDeviceSynchronize();
    lat<OP><<<1, 32>>>(out, cyc, 1.0000001, err); cudaDeviceSynchronize();
    long long h; double e; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&e, err, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %.1f cycles (incl. one add)   max |r*x-1| = %.2e\n", name, (double)h / N, e);
}
int main() {
    run<0>("f64: frcp_rn + 2 NR (current)"); run<1>("f64: rcp.approx.f32 + 2 NR (fma form)"); run<2>("f64: rcp.approx.f64 + 3 NR");
    run<3>("f64: rcp.approx.f64 + 2 NR"); run<4>("f32: 1.0f/x (current)"); run<5>("f32: rcp.approx + 1 NR"); run<6>("f32: rcp.approx");
    return 0;
}
