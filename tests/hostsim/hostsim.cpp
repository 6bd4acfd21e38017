// CPU build of the solver core (CAVE_HOST_SIM: one thread, warp width 1) for logic tests of the
// kernel source.  TEST INFRASTRUCTURE ONLY: built by tests/hostsim/build.py into
// tests/hostsim/_build/libcave_hostsim.so, loaded by tests/test_hostsim.py, never by cave_b200/.
#define CAVE_HOST_SIM 1
#include "../../cave_b200/csrc/solver_core.cuh"
#include <vector>
#include <cstdlib>

using namespace cave;

// plain restatement of what the scan kernel (scan_kernel.cu) writes into the pack
struct HostPack {
    std::vector<gen_t> gen; std::vector<uint8_t> ctype; std::vector<float> avg;
    std::vector<hash_t> ghash; std::vector<uint16_t> col; std::vector<float> val;
    int nvalid = 0, navg = 0, gen_nnz = 0, nsingc = 0;
    float maxl1 = 0.f, maxl2 = 0.f;
};
static void host_pack(const float* A, int m, int d, HostPack& pk) {
    pk.ctype.assign(d, 0); pk.avg.assign(d, 0.f); pk.ghash.assign(m, hash_t{0, 0});
    std::vector<float> gen_acc(d, 0.f); std::vector<int> sing(d, 0);
    for (int i = 0; i < m; ++i) {
        const float* row = A + (size_t)i * d;
        float l1 = 0.f, l2 = 0.f; int cnt = 0, lk = -1; float lv = 0.f;
        for (int k = 0; k < d; ++k) { float v = row[k]; l1 += fabsf(v); l2 += v * v; if (v != 0.f) { ++cnt; lk = k; lv = v; } }
        bool nv = l1 > 1e-7f; float nrm = sqrtf(l2); bool av = nrm > 1e-7f;
        float inv = av ? 1.f / fmaxf(nrm, 1e-8f) : 0.f;
        if (av) pk.navg++;
        if (nv) pk.nvalid++;
        if (cnt == 1) {
            if (nv) pk.ctype[lk] |= lv > 0.f ? 1 : 2;
            if (av) sing[lk] += lv > 0.f ? 1 : -1;
        } else if (cnt >= 2) {
            if (nv) {
                gen_t g; g.x = i; g.y = cnt; g.z = pk.gen_nnz; g.w = 0; pk.gen.push_back(g); pk.gen_nnz += cnt;
                hash_t h; h.x = 0; h.y = 0;
                for (int k = 0; k < d; ++k) if (row[k] != 0.f) {
                    union { float f; uint32_t u; } cv; cv.f = row[k];
                    h.x += mix64(((uint64_t)k << 32) | cv.u); h.y += mix64(((uint64_t)k << 32) | (cv.u ^ 0x80000000u));
                    pk.col.push_back((uint16_t)k); pk.val.push_back(row[k]);
                }
                pk.ghash[i] = h;
                pk.maxl1 = fmaxf(pk.maxl1, l1); pk.maxl2 = fmaxf(pk.maxl2, nrm * nrm);
            }
            if (av) for (int k = 0; k < d; ++k) gen_acc[k] += row[k] * inv;
        }
    }
    float n = (float)(pk.navg > 1 ? pk.navg : 1);
    for (int k = 0; k < d; ++k) { pk.avg[k] = (gen_acc[k] + (float)sing[k]) / n; if (pk.ctype[k]) pk.nsingc++; }
}

template <class T>
static void run(const float* A, int B, int m, int d, const double* pred, double sign, int mode, double inner_ratio,
                double gscale, int force_path, double* loss_i, double* grad, double* proj, double* rnorm,
                int* status, int* iters) {
    size_t cap = (size_t)64 << 20;
    char* buf = (char*)malloc(cap);
    for (int b = 0; b < B; ++b) {
        HostPack pk; host_pack(A + (size_t)b * m * d, m, d, pk);
        Instance in; in.A = A + (size_t)b * m * d; in.gen = pk.gen.data(); in.ctype = pk.ctype.data(); in.avg = pk.avg.data();
        in.d = d; in.ngen = (int)pk.gen.size(); in.gen_nnz = pk.gen_nnz; in.nvalid = pk.nvalid; in.nsingc = pk.nsingc;
        bool all8 = true; for (float v : pk.val) all8 = all8 && v == (float)(int)v && fabsf(v) <= 127.f;
        in.csr_ok = (force_path & 2) ? 0 : (all8 ? 3 : 1);                                  // bit 1: rebuild the CSR from A
        in.ghash = pk.ghash.data(); in.pcol = pk.col.data(); in.pval = pk.val.data(); in.maxl1 = pk.maxl1; in.maxl2 = pk.maxl2;
        in.setup = nullptr;
        if (force_path & 1) in.nsingc = 0 == in.nsingc ? 1 : in.nsingc;       // bit 0: force the Newton path
        Ctx cx; EpiParams ep; ep.mode = mode; ep.inner_ratio = inner_ratio; ep.sign = sign; ep.gscale = gscale;
        SolveOpts opt; opt.max_iter = 0; opt.max_ls = 0; opt.tol = 0;
        // a small "shared memory" so that both placements (shared-only hot arrays / generic) get exercised
        static char fake_smem[96 * 1024];
        solve_instance<T, double>(cx, in, fake_smem, sizeof(fake_smem), buf, cap, pred + (size_t)b * d, ep, opt, grad + (size_t)b * d, proj + (size_t)b * d,
                                  loss_i + b, rnorm + b, status + b, iters + b);
    }
    free(buf);
}

extern "C" void hostsim_forward_backward(const float* A, int B, int m, int d, const double* pred, double sign, int mode,
                                         double inner_ratio, double gscale, int compute_f32, int force_path,
                                         double* loss_i, double* grad, double* proj, double* rnorm, int* status, int* iters) {
    if (compute_f32) run<float>(A, B, m, d, pred, sign, mode, inner_ratio, gscale, force_path, loss_i, grad, proj, rnorm, status, iters);
    else run<double>(A, B, m, d, pred, sign, mode, inner_ratio, gscale, force_path, loss_i, grad, proj, rnorm, status, iters);
}
