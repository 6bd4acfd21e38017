// Kernel 2 — solve + fused epilogue: persistent CTAs pull instances from a work counter; each
// instance is solved entirely by one CTA with its working set in shared memory (spilling to the
// CTA's global scratch slot only when it does not fit), then the target / cosine loss /
// d loss / d pred epilogue runs in the same kernel (solver_core.cuh).
// Kernel 3 — finalize: fixed-order reduction of the per-instance losses (mean / sum) and dtype
// conversion of the per-instance outputs.
#include <cuda_runtime.h>
#include <stdint.h>
#include "layout.cuh"
#include "solve_kernel.cuh"
#include "solver_core.cuh"

namespace cave {

template <class T, class TIO>
__global__ void __launch_bounds__(256, 2) solve_kernel(SolveParams p) {
    extern __shared__ __align__(16) char smem[];
    __shared__ double red[64];
    __shared__ int s_b;
    if (p.cfg_id >= 0 && choose_solve_config(p.plan, sizeof(T), sizeof(TIO) == 8 ? (size_t)p.d * 4 : 0) != p.cfg_id) return;
    Ctx cx(red);
    char* slot = p.slots + (size_t)blockIdx.x * p.slot_bytes;
    EpiParams ep; ep.mode = p.mode; ep.inner_ratio = p.inner_ratio; ep.sign = p.sign; ep.gscale = p.gscale;
    SolveOpts opt; opt.max_iter = p.max_iter; opt.max_ls = p.max_ls; opt.tol = p.tol;
    const TIO* pred = (const TIO*)p.pred;
    TIO* grad = (TIO*)p.grad;
    TIO* proj = (TIO*)p.proj;
    for (;;) {
        if (cx.tid == 0) s_b = atomicAdd(p.counter, 1);
        __syncthreads();
        const int b = s_b;
        __syncthreads();
        if (b >= p.B) break;
        const size_t q = p.inst_index ? (size_t)p.inst_index[b] : (size_t)b;     // instance of the pack / of A
        Instance in;
        in.A = p.A ? p.A + q * p.m_max * p.d : nullptr;
        in.gen = p.gen + q * p.m_max;
        in.ctype = p.ctype + q * p.dpad;
        in.avg = p.avg + q * p.dpad;
        in.d = p.d; in.ngen = p.ngen[q]; in.gen_nnz = p.gennnz[q]; in.nvalid = p.nvalid[q]; in.nsingc = p.nsingc[q];
        in.csr_ok = p.csr_ok[q]; in.maxl1 = p.maxl1[q]; in.maxl2 = p.maxl2[q];
        in.ghash = p.ghash + q * p.m_max;
        in.pcol = p.csr_col + q * p.cap_nnz;
        in.pval = p.csr_val + q * p.cap_nnz;
        solve_instance<T, TIO>(cx, in, smem, (size_t)p.smem_bytes, slot, p.slot_bytes, pred + (size_t)b * p.d, ep, opt, grad + (size_t)b * p.d,
                               proj ? proj + (size_t)b * p.d : nullptr, p.loss64 + b, p.rnorm64 + b,
                               p.status + b, p.iters + b);
        __syncthreads();
    }
}

template <class TIO>
__global__ void __launch_bounds__(1024) finalize_kernel(FinalizeParams p) {
    __shared__ double part[1024];
    const int tid = threadIdx.x;
    TIO* loss_i = (TIO*)p.loss_i;
    TIO* rnorm = (TIO*)p.rnorm;
    double acc = 0.0;
    for (int i = tid; i < p.B; i += 1024) {
        const double l = p.loss64[i];
        acc += l;
        loss_i[i] = (TIO)l;
        if (rnorm) rnorm[i] = (TIO)p.rnorm64[i];
        if (p.status_out) p.status_out[i] = p.status[i];
        if (p.iters_out) p.iters_out[i] = p.iters[i];
    }
    part[tid] = acc;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if (tid < s) part[tid] += part[tid + s];
        __syncthreads();
    }
    if (tid == 0 && p.loss && p.reduction != 2)
        *(TIO*)p.loss = (TIO)(p.reduction == 0 ? part[0] / (double)p.B : part[0]);
}

template <class T, class TIO>
static cudaError_t launch_solve_t(const SolveParams& p, int grid, int threads, cudaStream_t stream) {
    static int configured = 0;
    if (configured < p.smem_bytes) {
        cudaError_t e = cudaFuncSetAttribute(solve_kernel<T, TIO>, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem_bytes);
        if (e != cudaSuccess) return e;
        configured = p.smem_bytes;
    }
    solve_kernel<T, TIO><<<grid, threads, p.smem_bytes, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_solve(const SolveParams& p, int compute_f32, int io_f32, int grid, int threads, cudaStream_t stream) {
    if (compute_f32) return io_f32 ? launch_solve_t<float, float>(p, grid, threads, stream)
                                   : launch_solve_t<float, double>(p, grid, threads, stream);
    return io_f32 ? launch_solve_t<double, float>(p, grid, threads, stream)
                  : launch_solve_t<double, double>(p, grid, threads, stream);
}

cudaError_t launch_finalize(const FinalizeParams& p, int io_f32, cudaStream_t stream) {
    if (io_f32) finalize_kernel<float><<<1, 1024, 0, stream>>>(p);
    else finalize_kernel<double><<<1, 1024, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace cave
