// Kernel 1 — scan/pack: ONE streaming pass over A (the HBM-bound part of the path).
//
// One CTA per instance.  Row tiles are brought into a shared-memory ring with 1-D TMA bulk
// copies (cp.async.bulk ... mbarrier::complete_tx) issued by one thread; eight warps consume:
//   phase A  warp per row : aligned 128-bit shared loads; a warp-uniform "all four lanes' words
//                           are zero" test skips the arithmetic for the (dominant) zero words,
//                           so sparse rows cost ~2 instructions per element.  Per row: non-zero
//                           count, and for rows that need them sum|a|, sum a^2, two order-free
//                           64-bit row hashes (for +-row matching in the solver).
//   phase B  one thread   : ordered bookkeeping of singleton rows (cone types, +-1 average terms)
//            all threads  : column-parallel accumulation of a/||a|| over the general rows,
//            warp per row : compaction of each general row into the packed CSR of the instance
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

__device__ __forceinline__ uint64_t mix64d(uint64_t x) {      // same mixer as solver_core.cuh
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

struct RowAcc {
    float l1, l2, lv;
    int lk, n;
    __device__ __forceinline__ void add(float v, int k) {
        l1 += fabsf(v);
        l2 = fmaf(v, v, l2);
        lk = k; lv = v; ++n;
    }
};

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
    uint64_t* full = (uint64_t*)align_up((size_t)(sing + d), 8);
    float* r_l1 = (float*)(full + S);
    float* r_nrm = r_l1 + R;
    float* r_inv = r_nrm + R;
    float* r_val = r_inv + R;
    int* r_cnt = (int*)(r_val + R);
    int* r_k = r_cnt + R;
    int* s_cnt = r_k + R;                                   // [8]
    unsigned char* ctype = (unsigned char*)align_up((size_t)(s_cnt + 8), 16);   // [dpad]

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

    // running totals, replicated in every thread (identical arithmetic everywhere)
    int t_ngen = 0, t_gennnz = 0;
    float t_l1max = 0.f, t_l2max = 0.f;
    // bookkeeper-only counters (thread NT-1)
    int c_nvalid = 0, c_navg = 0;
    int overflow = 0;
    int2* gen_out = p.gen + (size_t)b * p.m_max;
    ulonglong2* hash_out = p.ghash + (size_t)b * p.m_max;
    uint16_t* col_out = p.csr_col + (size_t)b * p.cap_nnz;
    float* val_out = p.csr_val + (size_t)b * p.cap_nnz;

    for (int t = 0; t < ntiles; ++t) {
        const int s = t % S;
        const int rows = min(R, m_b - t * R);
        const uintptr_t a = (uintptr_t)(A_b + (size_t)t * R * d);
        const int shift = (int)((a & 15) >> 2);
        const float* base = (const float*)(ring + (size_t)s * p.stage_stride);   // 16-byte aligned
        const float* tile = base + shift;
        mbar_wait(full + s, (uint32_t)((t / S) & 1));

        // ---- phase A: per-row statistics, one warp per row
        for (int rr = warp; rr < rows; rr += NW) {
            const int e0 = shift + rr * d, e1 = e0 + d;         // element range in `base` coordinates
            const int q0 = (e0 + 3) >> 2, q1 = e1 >> 2;         // whole float4 words inside the row
            RowAcc acc; acc.l1 = 0.f; acc.l2 = 0.f; acc.lv = 0.f; acc.lk = -1; acc.n = 0;
            int cnt = 0;
            if (q0 <= q1) {
                if (lane < 4) {                                  // ragged head and tail (< 4 elements each)
                    const int kh = e0 + lane;
                    if (kh < (q0 << 2) && kh < e1) { float v = base[kh]; if (v != 0.f) acc.add(v, kh - e0); }
                    const int kt = (q1 << 2) + lane;
                    if (kt >= (q0 << 2) && kt < e1) { float v = base[kt]; if (v != 0.f) acc.add(v, kt - e0); }
                }
                const float4* b4 = (const float4*)base;
                for (int q = q0 + lane; q < q1 + lane; q += 32) {       // uniform trip count
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (q < q1) v = b4[q];
                    const uint32_t any = (__float_as_uint(v.x) | __float_as_uint(v.y) | __float_as_uint(v.z) | __float_as_uint(v.w)) << 1;
                    if (__ballot_sync(0xffffffffu, any != 0u) == 0u) continue;      // all 128 words are zero
                    if (any != 0u) {
                        const int k = (q << 2) - e0;
                        if (v.x != 0.f) acc.add(v.x, k);
                        if (v.y != 0.f) acc.add(v.y, k + 1);
                        if (v.z != 0.f) acc.add(v.z, k + 2);
                        if (v.w != 0.f) acc.add(v.w, k + 3);
                    }
                }
            } else {                                             // row shorter than one aligned word
                for (int k = e0 + lane; k < e1; k += 32) { float v = base[k]; if (v != 0.f) acc.add(v, k - e0); }
            }
            // exact non-zero count of the row (lanes that saw nothing are skipped with one ballot)
            const unsigned sawm = __ballot_sync(0xffffffffu, acc.n > 0);
            if (sawm) {
                if ((sawm & (sawm - 1)) == 0u) {
                    cnt = __shfl_sync(0xffffffffu, acc.n, __ffs(sawm) - 1);
                } else {
                    cnt = acc.n;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
                }
            }
            float l1 = 0.f, nrm = 0.f, v1 = 0.f;
            int k1 = -1;
            if (cnt == 1) {
                const int src = __ffs(sawm) - 1;
                k1 = __shfl_sync(0xffffffffu, acc.lk, src);
                v1 = __shfl_sync(0xffffffffu, acc.lv, src);
                l1 = fabsf(v1); nrm = sqrtf(v1 * v1);
            } else if (cnt >= 2) {
                float a1 = acc.l1, a2 = acc.l2;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
                    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
                }
                l1 = a1; nrm = sqrtf(a2);
            }
            if (lane == 0) {
                r_l1[rr] = l1; r_nrm[rr] = nrm;
                r_inv[rr] = nrm > 1e-7f ? 1.f / fmaxf(nrm, 1e-8f) : 0.f;
                r_cnt[rr] = cnt; r_k[rr] = k1; r_val[rr] = v1;
            }
        }
        __syncthreads();

        // ---- phase B
        // every thread: masks of the tile's general rows (valid for the solver / valid for the average)
        unsigned gmask = 0, amask = 0;
        for (int rr = 0; rr < rows; ++rr) {
            const bool gen = r_cnt[rr] >= 2;
            if (gen && r_l1[rr] > 1e-7f) gmask |= 1u << rr;          // src/cave.py:303
            if (gen && r_inv[rr] != 0.f) amask |= 1u << rr;          // src/cave.py:224-225
        }
        if (tid == NT - 1) {                                         // ordered bookkeeping of the other rows
            for (int rr = 0; rr < rows; ++rr) {
                const bool nv = r_l1[rr] > 1e-7f, av = r_inv[rr] != 0.f;
                c_nvalid += nv; c_navg += av;
                if (r_cnt[rr] == 1) {
                    const int k = r_k[rr];
                    const bool pos = r_val[rr] > 0.f;
                    if (nv) ctype[k] |= pos ? 1 : 2;
                    if (av) sing[k] += pos ? 1 : -1;
                }
            }
        }
        if (amask) {                                                 // a / ||a|| over the general rows
            for (int k = tid; k < d; k += NT) {
                float acc = 0.f;
                for (unsigned m = amask; m; m &= m - 1) {
                    const int rr = __ffs(m) - 1;
                    acc = fmaf(tile[(size_t)rr * d + k], r_inv[rr], acc);
                }
                avg_gen[k] += acc;
            }
        }
        if (gmask) {                                                 // pack the general rows (CSR)
            int ord = 0, off = t_gennnz;
            for (unsigned m = gmask; m; m &= m - 1, ++ord) {
                const int rr = __ffs(m) - 1;
                const int cnt = r_cnt[rr];
                if (ord % NW == warp) {
                    uint64_t hp = 0, hn = 0;
                    if (off + cnt <= p.cap_nnz) {
                        const float* row = tile + (size_t)rr * d;
                        int w = off;
                        for (int k0 = 0; k0 < d; k0 += 32) {
                            const int k = k0 + lane;
                            const float v = k < d ? row[k] : 0.f;
                            const unsigned nzm = __ballot_sync(0xffffffffu, v != 0.f);
                            if (v != 0.f) {
                                const int pos = w + __popc(nzm & ((1u << lane) - 1u));
                                col_out[pos] = (uint16_t)k; val_out[pos] = v;
                                const uint32_t bits = __float_as_uint(v);
                                hp += mix64d(((uint64_t)k << 32) | bits);
                                hn += mix64d(((uint64_t)k << 32) | (bits ^ 0x80000000u));
                            }
                            w += __popc(nzm);
                        }
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            hp += __shfl_xor_sync(0xffffffffu, hp, o);
                            hn += __shfl_xor_sync(0xffffffffu, hn, o);
                        }
                    }
                    if (lane == 0) {
                        gen_out[t_ngen + ord] = make_int2(t * R + rr, cnt);
                        hash_out[t_ngen + ord] = make_ulonglong2(hp, hn);
                    }
                }
                if (off + cnt > p.cap_nnz) overflow = 1;
                off += cnt;
                t_l1max = fmaxf(t_l1max, r_l1[rr]);
                t_l2max = fmaxf(t_l2max, r_nrm[rr] * r_nrm[rr]);
            }
            t_ngen += ord; t_gennnz = off;
        }
        __syncthreads();
        if (tid == 0 && t + S < ntiles) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(t + S);
        }
    }

    if (tid == NT - 1) { s_cnt[0] = c_nvalid; s_cnt[1] = c_navg; }
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
        p.nvalid[b] = s_cnt[0]; p.navg[b] = s_cnt[1]; p.ngen[b] = t_ngen; p.gennnz[b] = t_gennnz; p.nsingc[b] = s_cnt[4];
        p.csr_ok[b] = overflow ? 0 : 1;
        p.maxl1[b] = t_l1max; p.maxl2[b] = t_l2max;
    }
}

size_t scan_smem_bytes(int d, int R, int stages, size_t* stage_stride_out) {
    const size_t stride = align_up((size_t)R * d * 4 + 32, 128);
    const int64_t dpad = (int64_t)align_up((size_t)d, 16);
    size_t o = (size_t)stages * stride;
    o += (size_t)d * 8 + 8;             // avg_gen, sing (+ alignment)
    o += (size_t)stages * 8;            // mbarriers
    o += (size_t)R * 24;                // row arrays
    o += 8 * 4 + 16;                    // counters (+ alignment)
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
