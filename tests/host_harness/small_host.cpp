// Test-only host build of csrc/cfmm_small.cuh: the same solve_one() the CUDA kernel runs per thread, looped over the
// problems of a batch, so the control flow can be checked against oracle/cfmm_oracle.py without a GPU.  Not part of the
// product: nothing under cfmm_routing_code_b200/ loads this.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../cfmm_routing_code_b200/csrc/cfmm_small.cuh"

extern "C" int small_host_solve(int n_tokens, long long n_pools, const long long* pool_ptr, const int* tok,
                                const double* R, const double* w, const double* logrw, const double* gamma,
                                const unsigned char* kind, int n_problems, const long long* pool_range, const double* c,
                                const double* a, const unsigned char* flags, double* nu, double* psi, double* stats,
                                double* delta, double* lam, long long trade_stride, double tol, int interleave) {
    using namespace cfmm_small;
    Pools P{(const int64_t*)pool_ptr, tok, R, w, logrw, gamma, kind};
    Params prm{tol, 0.1, 1e-4, 0.5, 1e-12, 60, 100};
    if (const char* e = getenv("SMALL_HOST_EPS")) sscanf(e, "%lf,%lf,%lf", &prm.eps0, &prm.eps_min, &prm.eps_shrink);   // experiments
    const int64_t nnz = pool_ptr[n_pools];
    const int64_t stride = interleave ? ((n_problems + 31) / 32) * 32 : 1;    // exercise the strided workspace too
    std::vector<double> work((size_t)work_doubles(n_tokens, nnz) * (interleave ? stride : n_problems));
    for (int p = 0; p < n_problems; ++p) {
        Problem Q;
        Q.n = n_tokens;
        Q.p0 = pool_range ? pool_range[2 * p] : 0;
        Q.p1 = pool_range ? pool_range[2 * p + 1] : n_pools;
        Q.off0 = pool_ptr[Q.p0];
        Q.c = c + (size_t)p * n_tokens; Q.a = a + (size_t)p * n_tokens; Q.flags = flags + (size_t)p * n_tokens;
        Q.delta = delta ? delta + p * trade_stride : nullptr;
        Q.lam = lam ? lam + p * trade_stride : nullptr;
        double* wk = interleave ? work.data() + p : work.data() + (size_t)p * work_doubles(n_tokens, nnz);
        Stats r = solve_one(P, Q, prm, nu + (size_t)p * n_tokens, psi + (size_t)p * n_tokens, wk, stride);
        double* st = stats + 8 * p;
        st[0] = r.value; st[1] = r.dual; st[2] = r.gap; st[3] = r.infeas; st[4] = r.err;
        st[5] = r.iters; st[6] = r.evals; st[7] = r.status;
    }
    return 0;
}
