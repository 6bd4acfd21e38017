// CTA execution context.  The solver core (solver_core.cuh) is written against this small
// interface so that the same source compiles (a) as sm_100a device code, where a Ctx is one
// CTA, and (b) with CAVE_HOST_SIM as plain single-threaded C++ (tid 0 of 1, warp width 1) for
// CPU-side logic tests of the kernel (tests/hostsim).  The host build is test infrastructure
// only; the product never loads it.
#pragma once
#include <stdint.h>
#include <math.h>

#ifdef CAVE_HOST_SIM
#include <algorithm>
#include <cstring>
#define CAVE_DEV inline
#define CAVE_RESTRICT
namespace cave {
struct Ctx {
    static constexpr int WS = 1;
    int tid = 0, nthr = 1, lane = 0, warp = 0, nwarp = 1;
    void sync() {}
    void syncwarp() {}
    void bar_named(int, int) {}
    template <class T> T warp_sum(T v) { return v; }
    template <class T> T warp_max(T v) { return v; }
    template <class T> T block_sum(T v) { return v; }
    template <class T> T block_max(T v) { return v; }
    template <class T> T block_min(T v) { return v; }
    // arg-max with smallest index on ties; idx < 0 means "no candidate"
    template <class T> void block_argmax(T& v, int& idx) {}
    template <class T> void block_argmin(T& v, int& idx) {}
    template <class T> void block_sum2(T&, T&) {}
    template <class T> void block_max2(T&, T&) {}
    uint64_t warp_sum_u64(uint64_t v) { return v; }
    unsigned ballot(bool p) { return p ? 1u : 0u; }
    int lanes_below(unsigned) { return 0; }
    int popc(unsigned m) { return (int)(m & 1u); }
    int atomic_add(int* p, int v) { int o = *p; *p += v; return o; }
    template <class T> void atomic_addf(T* p, T v) { *p += v; }
    template <class T> T shfl(T v, int) { return v; }
};
CAVE_DEV float ld_stream(const float* p) { return *p; }

// "Hot pointer": on the device with HOT = true a 32-bit shared-memory address whose element accesses are
// explicit ld.shared / st.shared / atom.shared (the compiler does not reliably infer the address space
// through the arena); with HOT = false, and always in the host simulation, a plain generic pointer.
template <class U, bool HOT>
struct HPtr {
    U* p;
    HPtr() : p(nullptr) {}
    explicit HPtr(U* q) : p(q) {}
    U& operator[](size_t i) const { return p[i]; }
    HPtr operator+(size_t o) const { return HPtr(p + o); }
    U* raw() const { return p; }
    void atomic_add(size_t i, U v) const { p[i] += v; }
    U fetch_add(size_t i, U v) const { U o = p[i]; p[i] += v; return o; }
};
}  // namespace cave
#else
#define CAVE_DEV __device__ __forceinline__
#define CAVE_RESTRICT __restrict__
namespace cave {
struct Ctx {
    static constexpr int WS = 32;
    int tid, nthr, lane, warp, nwarp;
    double* red;   // shared scratch: >= 2 * 32 doubles
    __device__ Ctx(double* red_) : red(red_) {
        tid = threadIdx.x; nthr = blockDim.x; lane = tid & 31; warp = tid >> 5; nwarp = nthr >> 5;
    }
    __device__ __forceinline__ void sync() { __syncthreads(); }
    __device__ __forceinline__ void syncwarp() { __syncwarp(); }
    __device__ __forceinline__ void bar_named(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
    template <class T> __device__ __forceinline__ T shfl(T v, int src) { return __shfl_sync(0xffffffffu, v, src); }
    template <class T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    template <class T> __device__ __forceinline__ T warp_max(T v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { T u = __shfl_xor_sync(0xffffffffu, v, o); v = u > v ? u : v; }
        return v;
    }
    template <class T> __device__ __forceinline__ T warp_min(T v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { T u = __shfl_xor_sync(0xffffffffu, v, o); v = u < v ? u : v; }
        return v;
    }
    // Block reductions: fixed order (lane tree, then warp 0 over the per-warp partials), result
    // broadcast to every thread.  Two barriers each.
    template <class T> __device__ T block_sum(T v) {
        v = warp_sum(v);
        sync();
        if (lane == 0) red[warp] = (double)v;
        sync();
        double s = 0.0;
        for (int w = 0; w < nwarp; ++w) s += red[w];
        return (T)s;
    }
    template <class T> __device__ T block_max(T v) {
        v = warp_max(v);
        sync();
        if (lane == 0) red[warp] = (double)v;
        sync();
        double s = red[0];
        for (int w = 1; w < nwarp; ++w) s = red[w] > s ? red[w] : s;
        return (T)s;
    }
    template <class T> __device__ T block_min(T v) {
        v = warp_min(v);
        sync();
        if (lane == 0) red[warp] = (double)v;
        sync();
        double s = red[0];
        for (int w = 1; w < nwarp; ++w) s = red[w] < s ? red[w] : s;
        return (T)s;
    }
    template <class T> __device__ void block_argmax(T& v, int& idx) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            T u = __shfl_xor_sync(0xffffffffu, v, o);
            int j = __shfl_xor_sync(0xffffffffu, idx, o);
            bool take = (j >= 0) && (idx < 0 || u > v || (u == v && j < idx));
            if (take) { v = u; idx = j; }
        }
        sync();
        if (lane == 0) { red[warp] = (double)v; ((int*)(red + 32))[warp] = idx; }
        sync();
        double bv = 0.0; int bi = -1;
        for (int w = 0; w < nwarp; ++w) {
            double u = red[w]; int j = ((int*)(red + 32))[w];
            if (j >= 0 && (bi < 0 || u > bv || (u == bv && j < bi))) { bv = u; bi = j; }
        }
        v = (T)bv; idx = bi;
    }
    template <class T> __device__ void block_argmin(T& v, int& idx) {
        T nv = -v; block_argmax(nv, idx); v = -nv;
    }
    template <class T> __device__ void block_sum2(T& a, T& b) {
        a = warp_sum(a); b = warp_sum(b);
        sync();
        if (lane == 0) { red[warp] = (double)a; red[32 + warp] = (double)b; }
        sync();
        double sa = 0.0, sb = 0.0;
        for (int w = 0; w < nwarp; ++w) { sa += red[w]; sb += red[32 + w]; }
        a = (T)sa; b = (T)sb;
    }
    template <class T> __device__ void block_max2(T& a, T& b) {
        a = warp_max(a); b = warp_max(b);
        sync();
        if (lane == 0) { red[warp] = (double)a; red[32 + warp] = (double)b; }
        sync();
        double sa = red[0], sb = red[32];
        for (int w = 1; w < nwarp; ++w) { sa = red[w] > sa ? red[w] : sa; sb = red[32 + w] > sb ? red[32 + w] : sb; }
        a = (T)sa; b = (T)sb;
    }
    __device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    __device__ __forceinline__ int popc(unsigned m) { return __popc(m); }
    template <class T> __device__ __forceinline__ void atomic_addf(T* p, T v) { atomicAdd(p, v); }
    __device__ __forceinline__ unsigned ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
    __device__ __forceinline__ int lanes_below(unsigned mask) { return __popc(mask & ((1u << lane) - 1u)); }
    __device__ __forceinline__ int atomic_add(int* p, int v) { return atomicAdd(p, v); }
};

// ---- explicit shared-memory element access (see HPtr below)
template <class U> struct SmemOps;
template <> struct SmemOps<float> {
    static __device__ __forceinline__ float ld(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v; }
    static __device__ __forceinline__ void st(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
    static __device__ __forceinline__ void add(uint32_t a, float v) { asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
};
template <> struct SmemOps<double> {
    static __device__ __forceinline__ double ld(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); return v; }
    static __device__ __forceinline__ void st(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
    static __device__ __forceinline__ void add(uint32_t a, double v) { asm volatile("red.shared.add.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
};
template <> struct SmemOps<int> {
    static __device__ __forceinline__ int ld(uint32_t a) { int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
    static __device__ __forceinline__ void st(uint32_t a, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
    static __device__ __forceinline__ void add(uint32_t a, int v) { asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
    static __device__ __forceinline__ int fetch_add(uint32_t a, int v) { int o; asm volatile("atom.shared.add.s32 %0, [%1], %2;" : "=r"(o) : "r"(a), "r"(v) : "memory"); return o; }
};
template <> struct SmemOps<uint16_t> {
    static __device__ __forceinline__ uint16_t ld(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return (uint16_t)v; }
    static __device__ __forceinline__ void st(uint32_t a, uint16_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"((uint32_t)v) : "memory"); }
    static __device__ __forceinline__ void add(uint32_t, uint16_t) {}
    static __device__ __forceinline__ uint16_t fetch_add(uint32_t, uint16_t) { return 0; }
};
template <> struct SmemOps<uint8_t> {
    static __device__ __forceinline__ uint8_t ld(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return (uint8_t)v; }
    static __device__ __forceinline__ void st(uint32_t a, uint8_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"((uint32_t)v) : "memory"); }
    static __device__ __forceinline__ void add(uint32_t, uint8_t) {}
};
template <class U> struct SRef {
    uint32_t a;
    __device__ __forceinline__ operator U() const { return SmemOps<U>::ld(a); }
    __device__ __forceinline__ const SRef& operator=(U v) const { SmemOps<U>::st(a, v); return *this; }
    __device__ __forceinline__ const SRef& operator=(const SRef& o) const { SmemOps<U>::st(a, SmemOps<U>::ld(o.a)); return *this; }
    __device__ __forceinline__ const SRef& operator-=(U v) const { SmemOps<U>::st(a, SmemOps<U>::ld(a) - v); return *this; }
    __device__ __forceinline__ const SRef& operator+=(U v) const { SmemOps<U>::st(a, SmemOps<U>::ld(a) + v); return *this; }
    __device__ __forceinline__ const SRef& operator*=(U v) const { SmemOps<U>::st(a, SmemOps<U>::ld(a) * v); return *this; }
};
template <class U, bool HOT> struct HPtr;
template <class U> struct HPtr<U, true> {
    uint32_t a;
    __device__ __forceinline__ HPtr() : a(0) {}
    __device__ __forceinline__ explicit HPtr(U* q) : a((uint32_t)__cvta_generic_to_shared(q)) {}
    __device__ __forceinline__ SRef<U> operator[](uint32_t i) const { SRef<U> r; r.a = a + i * (uint32_t)sizeof(U); return r; }
    __device__ __forceinline__ HPtr operator+(uint32_t o) const { HPtr h; h.a = a + o * (uint32_t)sizeof(U); return h; }
    __device__ __forceinline__ U* raw() const { return (U*)__cvta_shared_to_generic((size_t)a); }
    __device__ __forceinline__ void atomic_add(uint32_t i, U v) const { SmemOps<U>::add(a + i * (uint32_t)sizeof(U), v); }
    __device__ __forceinline__ U fetch_add(uint32_t i, U v) const { return SmemOps<U>::fetch_add(a + i * (uint32_t)sizeof(U), v); }
};
template <class U> struct HPtr<U, false> {
    U* p;
    __device__ __forceinline__ HPtr() : p(nullptr) {}
    __device__ __forceinline__ explicit HPtr(U* q) : p(q) {}
    __device__ __forceinline__ U& operator[](size_t i) const { return p[i]; }
    __device__ __forceinline__ HPtr operator+(size_t o) const { return HPtr(p + o); }
    __device__ __forceinline__ U* raw() const { return p; }
    __device__ __forceinline__ void atomic_add(size_t i, U v) const { atomicAdd(p + i, v); }
    __device__ __forceinline__ U fetch_add(size_t i, U v) const { return atomicAdd(p + i, v); }
};

// streaming global load that does not pollute L1 (rows of A are touched once per phase)
CAVE_DEV float ld_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
}  // namespace cave
#endif
