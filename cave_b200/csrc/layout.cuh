// Device-resident layouts shared by the scan kernel, the solve kernel and the C ABI.
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define CAVE_HD __host__ __device__ inline
#else
#define CAVE_HD inline
#endif

namespace cave {

CAVE_HD size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// The "pack": everything the solver needs to know about A besides the general rows themselves.
//   nvalid[B]   rows with sum|a| > 1e-7            (the NNLS row mask, src/cave.py:303)
//   navg[B]     rows with ||a||_2 > 1e-7           (the average's row mask, src/cave.py:224-225)
//   ngen[B]     valid rows with >= 2 non-zeros     ("general" rows)
//   gennnz[B]   non-zeros in the general rows
//   nsingc[B]   coordinates that own at least one singleton row
//   gen[B, m_max]   int4 (row index, nnz, offset into the packed CSR, 0) of every general row, ascending row
//                   (offsets need not be monotone: rows are placed as they are scanned)
//   ctype[B, dpad]  per coordinate: bit0 = a row +a*e_k exists, bit1 = a row -a*e_k exists
//   avg[B, dpad]    float32 average unit normal     (src/cave.py:222-228)
//   csrok[B], maxl1[B], maxl2[B]   packed-CSR complete flag; max ||a||_1, max ||a||_2^2 over general rows
//   ghash[B, m_max]   indexed by ROW: order-free 64-bit hashes of the row and of its negation (general rows only)
//   csr_col/val[B, cap_nnz]   the general rows' non-zeros, row after row, columns ascending
struct PackLayout {
    size_t nvalid, navg, ngen, gennnz, nsingc, gen, ctype, avg, csrok, maxl1, maxl2, ghash, csr_col, csr_val, total;
    int64_t dpad, cap_nnz;
};

// per-instance capacity of the packed CSR: every shipped model (TSP/VRP/SP) fits; instances that do
// not (dense general rows) are flagged and the solver reads their rows from A instead
CAVE_HD int64_t pack_cap_nnz(int64_t m_max, int64_t d) {
    int64_t full = m_max * d, c = 16 * d + 4096;
    return (c < full ? c : full + 1) / 8 * 8 + 8;
}

CAVE_HD PackLayout make_pack_layout(int64_t B, int64_t m_max, int64_t d) {
    PackLayout L;
    L.dpad = (int64_t)align_up((size_t)d, 16);
    size_t o = 0;
    L.nvalid = o; o = align_up(o + (size_t)B * 4, 256);
    L.navg = o;   o = align_up(o + (size_t)B * 4, 256);
    L.ngen = o;   o = align_up(o + (size_t)B * 4, 256);
    L.gennnz = o; o = align_up(o + (size_t)B * 4, 256);
    L.nsingc = o; o = align_up(o + (size_t)B * 4, 256);
    L.gen = o;    o = align_up(o + (size_t)B * (size_t)m_max * 16, 256);
    L.ctype = o;  o = align_up(o + (size_t)B * (size_t)L.dpad, 256);
    L.avg = o;    o = align_up(o + (size_t)B * (size_t)L.dpad * 4, 256);
    L.cap_nnz = pack_cap_nnz(m_max, d);
    L.csrok = o;  o = align_up(o + (size_t)B * 4, 256);
    L.maxl1 = o;  o = align_up(o + (size_t)B * 4, 256);
    L.maxl2 = o;  o = align_up(o + (size_t)B * 4, 256);
    L.ghash = o;  o = align_up(o + (size_t)B * (size_t)m_max * 16, 256);
    L.csr_col = o; o = align_up(o + (size_t)B * (size_t)L.cap_nnz * 2, 256);
    L.csr_val = o; o = align_up(o + (size_t)B * (size_t)L.cap_nnz * 4, 256);
    L.total = o;
    return L;
}

// Solver scratch: [work counter][loss64 B][rnorm64 B][status B][iters B][one slot per CTA]
struct ScratchLayout {
    size_t counter, loss64, rnorm64, status, iters, slots, slot_bytes, total;
    int64_t n_slots;
};

CAVE_HD size_t solver_slot_bytes(int64_t d, int64_t cap_rows, int64_t cap_nnz, size_t T) {
    const size_t r = (size_t)cap_rows + 2, dd = (size_t)d + 2;
    // Newton path (upper bound over every Arena::get in nw_setup / newton_solve)
    size_t nw = 3 * dd * T + 4 * dd + 4 * dd + r * (6 * T + 4 + 4 + 1 + 4 + 1 + 4 + 4 + 8 + 4)
              + (size_t)(cap_nnz + 2) * 12 + r * r * T + (r + 1) * (r + 2) * T;
    // Lawson-Hanson path
    const size_t k = (cap_rows < d ? (size_t)cap_rows : (size_t)d) + 2;
    size_t lh = 3 * dd * T + dd * T + k * (5 * T + 8) + r * 5 + 2 * k * k * T;
    size_t m = nw > lh ? nw : lh;
    return align_up(m + 64 * 16 + 1024, 256);
}

CAVE_HD ScratchLayout make_scratch_layout(int64_t B, int64_t d, int64_t cap_rows, int64_t cap_nnz, size_t T,
                                          int64_t n_slots) {
    ScratchLayout L;
    size_t o = 0;
    L.counter = o; o = align_up(o + 256, 256);
    L.loss64 = o;  o = align_up(o + (size_t)B * 8, 256);
    L.rnorm64 = o; o = align_up(o + (size_t)B * 8, 256);
    L.status = o;  o = align_up(o + (size_t)B * 4, 256);
    L.iters = o;   o = align_up(o + (size_t)B * 4, 256);
    L.slot_bytes = solver_slot_bytes(d, cap_rows, cap_nnz, T);
    L.n_slots = n_slots;
    L.slots = o;   o += L.slot_bytes * (size_t)n_slots;
    L.total = o;
    return L;
}

}  // namespace cave
