// Dependent-chain latencies of the operations on the solver's serial paths (one warp, clock64 around a chain).
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <int OP>
__global__ void lat(double* out, long long* cyc, double seed, float fseed) {
    __shared__ double sm[64];
    sm[threadIdx.x & 63] = seed + threadIdx.x;
    __syncthreads();
    double x = seed + threadIdx.x * 1e-3, y = seed * 0.5;
    float xf = fseed + threadIdx.x;
    int idx = threadIdx.x & 63;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (OP == 0) x = fma(x, y, 1.0);                                    // DFMA
        if (OP == 1) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);   // 64-bit shuffle (2 SHFL)
        if (OP == 2) xf = fmaf(xf, 0.999f, 1.0f);                           // FFMA
        if (OP == 3) { x = sm[idx]; idx = ((int)x) & 63; }                  // LDS f64 + F2I dependent
        if (OP == 4) xf = __frcp_rn(xf) + 1.0f;                             // MUFU.RCP (+FADD)
        if (OP == 5) x = (double)(float)x + 1.0;                            // F2F down, F2F up, DADD
        if (OP == 6) x = 1.0 / (x + 2.0);                                   // full double division
        if (OP == 7) { asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"((unsigned)__cvta_generic_to_shared(sm + (idx & 63))) : "memory"); idx = __double2int_rn(x) & 63; }
        if (OP == 8) xf = __shfl_sync(0xffffffffu, xf, (threadIdx.x + 1) & 31);  // 32-bit shuffle
        if (OP == 9) { __syncwarp(); x = fma(x, y, 1.0); }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + xf + idx;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int OP> void run(const char* name) {
    double* out; long long* cyc; cudaMalloc(&out, 256 * 8); cudaMalloc(&cyc, 8);
    lat<OP><<<1, 32>>>(out, cyc, 1.0000001, 1.5f); cudaDeviceSynchronize();
    lat<OP><<<1, 32>>>(out, cyc, 1.0000001, 1.5f); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s %.1f cycles/op\n", name, (double)h / N);
}
// barrier cost: all threads of the CTA loop over __syncthreads
__global__ void bar(long long* cyc) {
    long long t0 = clock64();
    for (int i = 0; i < 1024; ++i) __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    run<0>("DFMA dependent"); run<1>("SHFL f64 dependent"); run<2>("FFMA dependent"); run<3>("LDS f64 + F2I dependent");
    run<4>("MUFU.RCP + FADD"); run<5>("F2F down/up + DADD"); run<6>("double division + DADD"); run<7>("ld.shared asm + D2I");
    run<8>("SHFL f32 dependent"); run<9>("syncwarp + DFMA");
    long long* cyc; cudaMalloc(&cyc, 8);
    for (int th : {64, 128, 160, 256}) { bar<<<1, th>>>(cyc); cudaDeviceSynchronize(); long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("__syncthreads %3d threads   %.1f cycles\n", th, (double)h / 1024); }
    return 0;
}
