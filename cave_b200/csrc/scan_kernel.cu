// Kernel 1 — scan/pack: ONE streaming pass over A (the HBM-bound part of the path).
//
// One CTA per instance.  Row tiles are brought into a shared-memory ring with 1-D TMA bulk
// copies (cp.async.bulk ... mbarrier::complete_tx) issued by one thread; eight warps consume:
//   phase A  warp per row : sum|a|, sum a^2, non-zero count, position/sign of a lone non-zero
//   phase B  one thread   : ordered bookkeeping (row classes, general-row list, singleton
//                           cone types / +-1 contributions to the average)
//            all threads  : column-parallel accumulation of a/||a|| over the general rows
// Everything `_average_ctrs` (src/cave.py:222-228) and the row mask of `_project_nnls`
// (src/cave.py:303) recompute on the host for every call is produced here in one read of A.
// All accumulation orders are fixed, so the pack is bit-reproducible run to run.
#include <cuda_runtime.h>
#include <stdint.h>
#include "layout.cuh"
#include "scan_kernel.cuh"

namespace cave {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <int NT>
__global__ void __launch_bounds__(NT) scan_kernel(ScanParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    const int d = p.d, R = p.R, S = p.stages;
    const int b = blockIdx.x;

    unsigned char* ring = smem;
    float* avg_gen = (float*)(smem + (size_t)S * p.stage_stride);
    int* sing = (int*)(avg_gen + d);
    float* r_l1 = (float*)(sing + d);
    float* r_l2 = r_l1 + R;
    float* r_inv = r_l2 + R;
    float* r_val = r_inv + R;
    int* r_cnt = (int*)(r_val + R);
    int* r_k = r_cnt + R;
    int* s_cnt = r_k + R;                                   // [8] nvalid, navg, ngen, gennnz, nsingc
    uint64_t* full = (uint64_t*)align_up((size_t)(s_cnt + 8), 8);
    unsigned char* ctype = (unsigned char*)(full + S);      // [dpad]

    const int m_b = p.m_rows ? min(max(p.m_rows[b], 0), p.m_max) : p.m_max;
    const int ntiles = (m_b + R - 1) / R;
    const float* A_b = p.A + (size_t)b * p.m_max * d;

    for (int k = tid; k < d; k += NT) { avg_gen[k] = 0.f; sing[k] = 0; }
    for (int k = tid; k < (int)p.dpad; k += NT) ctype[k] = 0;
    if (tid < 8) s_cnt[tid] = 0;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int t) {
        const int s = t % S;
        const int rows = min(R, m_b - t * R);
        const uintptr_t a = (uintptr_t)(A_b + (size_t)t * R * d);
        const uintptr_t a0 = a & ~(uintptr_t)15;
        const uintptr_t a1 = (a + (size_t)rows * d * 4 + 15) & ~(uintptr_t)15;
        const uint32_t bytes = (uint32_t)(a1 - a0);
        mbar_expect_tx(full + s, bytes);
        tma_bulk_g2s(ring + (size_t)s * p.stage_stride, (const void*)a0, bytes, full + s);
    };
    if (tid == 0)
        for (int t = 0; t < S && t < ntiles; ++t) issue(t);

    // bookkeeper's private counters (thread NT-1)
    int c_nvalid = 0, c_navg = 0, c_ngen = 0, c_gennnz = 0;
    int2* gen_out = p.gen + (size_t)b * p.m_max;

    for (int t = 0; t < ntiles; ++t) {
        const int s = t % S;
        const int rows = min(R, m_b - t * R);
        const uintptr_t a = (uintptr_t)(A_b + (size_t)t * R * d);
        const int shift = (int)((a & 15) >> 2);
        const float* tile = (const float*)(ring + (size_t)s * p.stage_stride) + shift;
        mbar_wait(full + s, (uint32_t)((t / S) & 1));

        // ---- phase A: per-row statistics, one warp per row
        for (int rr = warp; rr < rows; rr += NW) {
            const float* row = tile + (size_t)rr * d;
            float l1 = 0.f, l2 = 0.f, lv = 0.f;
            int cnt = 0, lk = -1;
            for (int k = lane; k < d; k += 32) {
                float v = row[k];
                l1 += fabsf(v);
                l2 = fmaf(v, v, l2);
                if (v != 0.f) { ++cnt; lk = k; lv = v; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                l1 += __shfl_xor_sync(0xffffffffu, l1, o);
                l2 += __shfl_xor_sync(0xffffffffu, l2, o);
                cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            }
            int lkmax = lk;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) lkmax = max(lkmax, __shfl_xor_sync(0xffffffffu, lkmax, o));
            unsigned owner = __ballot_sync(0xffffffffu, lk == lkmax && lk >= 0);
            float v1 = __shfl_sync(0xffffffffu, lv, owner ? (__ffs(owner) - 1) : 0);
            if (lane == 0) {
                float nrm = sqrtf(l2);
                r_l1[rr] = l1; r_l2[rr] = nrm;
                r_inv[rr] = nrm > 1e-7f ? 1.f / fmaxf(nrm, 1e-8f) : 0.f;
                r_cnt[rr] = cnt; r_k[rr] = lkmax; r_val[rr] = v1;
            }
        }
        __syncthreads();

        // ---- phase B: ordered bookkeeping by one thread ...
        if (tid == NT - 1) {
            for (int rr = 0; rr < rows; ++rr) {
                const bool nv = r_l1[rr] > 1e-7f;       // src/cave.py:303
                const bool av = r_inv[rr] != 0.f;       // src/cave.py:224-225
                const int cnt = r_cnt[rr];
                c_nvalid += nv; c_navg += av;
                if (cnt == 1) {
                    const int k = r_k[rr];
                    const bool pos = r_val[rr] > 0.f;
                    if (nv) ctype[k] |= pos ? 1 : 2;
                    if (av) sing[k] += pos ? 1 : -1;
                } else if (cnt >= 2 && nv) {
                    gen_out[c_ngen++] = make_int2(t * R + rr, cnt);
                    c_gennnz += cnt;
                }
            }
        }
        // ... while every thread accumulates its columns over the general rows of the tile
        for (int k = tid; k < d; k += NT) {
            float acc = 0.f;
            for (int rr = 0; rr < rows; ++rr)
                if (r_cnt[rr] >= 2 && r_inv[rr] != 0.f) acc = fmaf(tile[(size_t)rr * d + k], r_inv[rr], acc);
            avg_gen[k] += acc;
        }
        __syncthreads();
        if (tid == 0 && t + S < ntiles) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(t + S);
        }
    }

    if (tid == NT - 1) { s_cnt[0] = c_nvalid; s_cnt[1] = c_navg; s_cnt[2] = c_ngen; s_cnt[3] = c_gennnz; }
    __syncthreads();
    int nsc = 0;
    for (int k = tid; k < d; k += NT) nsc += ctype[k] != 0;
    for (int o = 16; o > 0; o >>= 1) nsc += __shfl_xor_sync(0xffffffffu, nsc, o);
    if (lane == 0 && nsc) atomicAdd(&s_cnt[4], nsc);
    __syncthreads();
    const float ninv = 1.f / (float)max(s_cnt[1], 1);
    float* avg_out = p.avg + (size_t)b * p.dpad;
    for (int k = tid; k < (int)p.dpad; k += NT) avg_out[k] = k < d ? (avg_gen[k] + (float)sing[k]) * ninv : 0.f;
    uint32_t* ct_out = (uint32_t*)(p.ctype + (size_t)b * p.dpad);
    for (int k = tid; k < (int)(p.dpad / 4); k += NT) ct_out[k] = ((const uint32_t*)ctype)[k];
    if (tid == 0) {
        p.nvalid[b] = s_cnt[0]; p.navg[b] = s_cnt[1]; p.ngen[b] = s_cnt[2]; p.gennnz[b] = s_cnt[3]; p.nsingc[b] = s_cnt[4];
    }
}

size_t scan_smem_bytes(int d, int R, int stages, size_t* stage_stride_out) {
    const size_t stride = align_up((size_t)R * d * 4 + 32, 128);
    const int64_t dpad = (int64_t)align_up((size_t)d, 16);
    size_t o = (size_t)stages * stride;
    o += (size_t)d * 8;                 // avg_gen, sing
    o += (size_t)R * 24;                // row arrays
    o += 8 * 4 + 8;                     // counters (+ alignment)
    o += (size_t)stages * 8;            // mbarriers
    o += (size_t)dpad;                  // ctype
    if (stage_stride_out) *stage_stride_out = stride;
    return align_up(o, 16);
}

cudaError_t launch_scan(const ScanParams& p0, cudaStream_t stream) {
    ScanParams p = p0;
    constexpr int NT = 256;
    // tile = R rows, about 20 KB; ring depth 4 unless shared memory runs out
    int R = (int)(20480 / ((size_t)p.d * 4));
    R = R < 1 ? 1 : (R > 32 ? 32 : R);
    int stages = 4;
    size_t stride = 0, smem = scan_smem_bytes(p.d, R, stages, &stride);
    while (smem > 200 * 1024 && stages > 2) { --stages; smem = scan_smem_bytes(p.d, R, stages, &stride); }
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    p.R = R; p.stages = stages; p.stage_stride = stride;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(scan_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    scan_kernel<NT><<<dim3((unsigned)p.B), dim3(NT), smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace cave
