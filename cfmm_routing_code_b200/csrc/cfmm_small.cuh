// cfmm_small.cuh -- the WHOLE dual solve of one small routing problem in one thread.
//
// The reference's three scripts are small problems (5 pools, 3-5 tokens), and two-asset.py solves 50 of them in a
// python loop, rebuilding the cvxpy problem each time (two-asset.py:40-100).  A batch of such problems is data
// parallel over PROBLEMS, not pools: one thread owns one problem and runs the complete method -- per-pool closed
// forms (arbitrage.py:60-74), psi (arbitrage.py:54), projected Newton in log-price coordinates with a dense Hessian,
// method of multipliers on the constant-sum fills -- without ever leaving the kernel.  Same algorithm, constants and
// control flow as oracle/cfmm_oracle.py::solve (which the tests check it against).
//
// All per-problem state lives in a caller-provided workspace, element-interleaved across problems (element e of
// problem p at work[e * stride + p]) so the 32 problems of a warp touch consecutive addresses.
//
// The file is plain C++ when compiled without nvcc: tests/ builds it for the host to check the control flow without a
// GPU.  The product only ever runs it through k_batch_solve (cfmm_small.cu).
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define CFMM_HD __host__ __device__
#else
#define CFMM_HD
#endif

namespace cfmm_small {

constexpr int KMAX = 8;              // largest weighted-pool arity (the reference uses 3..5; cfg3/4 use 2..8)
constexpr int NTOK_MAX = 64;         // dense n x n Newton systems per thread: keep n small
constexpr double TINY = 1e-300;
constexpr double DT_MAX = 3.0;       // largest log-price change of one Newton step

struct Pools {                       // CSR problem data (device pointers in the kernel)
    const int64_t* pool_ptr;         // [m+1]
    const int32_t* tok;              // [nnz]   token of every slot            arbitrage.py:6-12
    const double* R;                 // [nnz]   reserves                       arbitrage.py:14-20
    const double* w;                 // [nnz]   normalised weights | 0 on constant-sum pools | virtual offsets (kind 3)
    const double* logrw;             // [nnz]   log(R/w)
    const double* gamma;             // [m]     fees                           arbitrage.py:22-28
    const uint8_t* kind;             // [m]     1 = constant sum; 3 = bounded-liquidity product; 0 or 2 = weighted geometric mean
                                     //         (constant product = equal weights)
};

struct Params {
    double tol, eps0, eps_min, eps_shrink, floor_rel;
    int max_outer, max_inner;
};

struct Vec {                         // strided view into the interleaved workspace
    double* p;
    int64_t s;
    CFMM_HD double& operator[](int64_t j) const { return p[j * s]; }
};

struct Problem {
    int n;                           // tokens
    int64_t p0, p1;                  // pool range
    int64_t off0;                    // pool_ptr[p0]
    const double* c;                 // [n] objective                         arbitrage.py:31-36 / liquidation.py:57
    const double* a;                 // [n] endowment                         liquidation.py:30-36 / two-asset.py:45
    const uint8_t* flags;            // [n] bit0: psi_j + a_j == 0, bit1: psi_j free (nu_j = c_j)
    Vec theta_bar, theta_new;        // [nnz of the problem] multipliers of the constant-sum limit orders
    double* delta;                   // nullable; indexed by CSR offset
    double* lam;
};

struct Stats { double value, dual, gap, infeas, err; int iters, evals, status; };

CFMM_HD inline bool is_eq(const Problem& Q, int j) { return Q.flags[j] & 1; }
CFMM_HD inline bool is_pinned(const Problem& Q, int j) { return Q.flags[j] & 2; }

// Constant product on virtual reserves V = R + o with the real reserves R >= 0 (one Uniswap-v3 tick range; not a
// reference atom).  The optimal trade is the constant-product one on V; when it would pay out more than R_b the payout
// is capped there and the tender follows from the curve, (V_a + gamma D_a)(V_b - R_b) = V_a V_b.  hc = coefficient of
// [[1,-1],[-1,1]] in the scaled Hessian (0 at the cap: the trade no longer depends on the prices).
// Shared by the per-thread solver below and the pool-parallel kernel k_eval_pair (cfmm_kernels.cu).
CFMM_HD inline void bounded_pair(double R0, double R1, double o0, double o1, double gam, double n0, double n1,
                                 double* D, double* L, double& hc) {
    const double R[2] = {R0, R1}, o[2] = {o0, o1};
    const double V[2] = {R0 + o0, R1 + o1};
    const double pv[2] = {n0 * V[0], n1 * V[1]};
    D[0] = D[1] = L[0] = L[1] = 0.0;
    hc = 0.0;
    for (int dir = 0; dir < 2; ++dir) {
        const int ta = dir, tb = 1 - dir;
        if (gam * pv[tb] > pv[ta]) {
            const double t = sqrt(gam * pv[tb] / pv[ta]);
            const double lout = V[tb] * (1.0 - 1.0 / t);
            if (lout > R[tb]) {                                        // payout capped by the real reserve
                L[tb] = R[tb];
                D[ta] = V[ta] * R[tb] / (o[tb] * gam);
            } else {
                L[tb] = lout;
                D[ta] = V[ta] * (t - 1.0) / gam;
                hc += 0.5 * sqrt(pv[0] * pv[1] / gam);
            }
        }
    }
}

#ifdef __CUDA_ARCH__
// sum over the LANES consecutive lanes that share a problem; x + y == y + x exactly, so every lane ends with the same bits
template <int LANES>
__device__ __forceinline__ double lanes_sum(double v) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif

// One dual evaluation of the problem: psi = sum_i A_i (L_i - D_i), returns arb = sum_i nu_i'(L_i - D_i).
// Hs (n x n, nullable) receives the scaled Hessian (true Hessian = diag(1/nu) Hs diag(1/nu)).
//
// LANES > 1 (device only): LANES threads of one warp own the SAME problem.  Every lane keeps a full private copy of the
// state and runs the identical instruction stream; only this pool loop is split (lane l takes pools l, l+LANES, ...)
// and the partial psi / arb / Hessian / fills are then summed over the lanes with an xor butterfly, which leaves
// bit-identical totals in every lane -- so the lanes never diverge and need no other communication.
template <int LANES>
CFMM_HD inline double evaluate(const Pools& P, const Problem& Q, const Vec& nu, const Vec& lognu, double eps,
                               const Vec& psi, const Vec* Hs, bool trades, bool store_fill, int lane) {
    const int n = Q.n;
    for (int j = 0; j < n; ++j) { lognu[j] = log(nu[j]); psi[j] = 0.0; }
    if (Hs) for (int e = 0; e < n * n; ++e) (*Hs)[e] = 0.0;
    if (LANES > 1 && store_fill) {
        const int64_t nnz = P.pool_ptr[Q.p1] - Q.off0;
        for (int64_t x = 0; x < nnz; ++x) Q.theta_new[x] = 0.0;
    }
    double arb = 0.0;
    for (int64_t i = Q.p0 + (LANES > 1 ? lane : 0); i < Q.p1; i += LANES) {
        const int64_t off = P.pool_ptr[i];
        const int k = (int)(P.pool_ptr[i + 1] - off);
        const double gam = P.gamma[i];
        double D[KMAX], L[KMAX];
        if (P.kind[i] == 3) {
            double hc = 0.0;
            bounded_pair(P.R[off], P.R[off + 1], P.w[off], P.w[off + 1], gam, nu[P.tok[off]], nu[P.tok[off + 1]], D, L, hc);
            if (Hs && hc != 0.0) {
                const int t0 = P.tok[off], t1 = P.tok[off + 1];
                (*Hs)[t0 * n + t0] += hc; (*Hs)[t1 * n + t1] += hc;
                (*Hs)[t0 * n + t1] -= hc; (*Hs)[t1 * n + t0] -= hc;
            }
        } else if (P.kind[i] != 1) {
            double wa[KMAX];
            double M = 0.0;
            if (k == 2 && P.w[off] == P.w[off + 1]) {                      // sqrt(x0 x1) >= sqrt(R0 R1), arbitrage.py:68-70
                const double R0 = P.R[off], R1 = P.R[off + 1];
                const double p0 = nu[P.tok[off]] * R0, p1 = nu[P.tok[off + 1]] * R1;
                const bool f = gam * p1 > p0, b = gam * p0 > p1;
                const double q = f ? gam * p1 / p0 : (b ? gam * p0 / p1 : 1.0);
                const double t = sqrt(q);
                D[0] = f ? R0 * (t - 1.0) / gam : 0.0;
                L[1] = f ? R1 * (1.0 - 1.0 / t) : 0.0;
                D[1] = b ? R1 * (t - 1.0) / gam : 0.0;
                L[0] = b ? R0 * (1.0 - 1.0 / t) : 0.0;
                M = (f || b) ? 2.0 * sqrt(p0 * p1 / gam) : 0.0;
                wa[0] = wa[1] = (f || b) ? 0.5 : 0.0;
            } else {                                                       // weighted geometric mean, arbitrage.py:65
                double tA[KMAX], tB[KMAX];
                const double lg = log(gam);
                double maxB = -INFINITY, minA = INFINITY;
                for (int j = 0; j < k; ++j) {
                    tB[j] = P.logrw[off + j] + lognu[P.tok[off + j]];
                    tA[j] = tB[j] - lg;
                    maxB = fmax(maxB, tB[j]);
                    minA = fmin(minA, tA[j]);
                }
                const bool trade = maxB > minA;
                double sL = -INFINITY, hL = 0.0;                            // largest breakpoint with h <= 0
                for (int cnd = 0; cnd < 2 * k; ++cnd) {
                    const double T = cnd < k ? tA[cnd] : tB[cnd - k];
                    double hT = 0.0;
                    for (int j = 0; j < k; ++j) hT += P.w[off + j] * (fmax(T - tA[j], 0.0) + fmin(T - tB[j], 0.0));
                    if (hT <= 0.0 && T > sL) { sL = T; hL = hT; }
                }
                double W = 0.0;
                for (int j = 0; j < k; ++j) if (sL >= tA[j] || sL < tB[j]) W += P.w[off + j];
                const double s = hL < 0.0 ? sL - hL / fmax(W, TINY) : sL;
                for (int j = 0; j < k; ++j) {
                    const double zA = trade ? fmax(s - tA[j], 0.0) : 0.0;
                    const double zB = trade ? fmin(s - tB[j], 0.0) : 0.0;
                    D[j] = P.R[off + j] * expm1(zA) / gam;
                    L[j] = -P.R[off + j] * expm1(zB);
                    wa[j] = (zA > 0.0 || zB < 0.0) ? P.w[off + j] : 0.0;
                }
                M = trade ? exp(s) : 0.0;
            }
            if (Hs && M > 0.0) {
                double Wa = 0.0;
                for (int j = 0; j < k; ++j) Wa += wa[j];
                Wa = fmax(Wa, TINY);
                for (int x = 0; x < k; ++x) {
                    const int tx = P.tok[off + x];
                    (*Hs)[tx * n + tx] += M * wa[x];
                    for (int y = 0; y < k; ++y) (*Hs)[tx * n + P.tok[off + y]] -= (M / Wa) * wa[x] * wa[y];
                }
            }
        } else {                                                           // constant sum with x >= 0, arbitrage.py:73-74
            double hc = 0.0;
            D[0] = D[1] = L[0] = L[1] = 0.0;
            for (int dir = 0; dir < 2; ++dir) {                            // limit order: tender ta, receive up to R_tb of tb
                const int ta = dir, tb = 1 - dir;
                const double na = nu[P.tok[off + ta]], nb = nu[P.tok[off + tb]];
                const double r = gam * nb / na, z = r - 1.0, Rb = P.R[off + tb];
                double th, ps, curv = 0.0;
                if (eps <= 0.0) {
                    th = z > 0.0 ? Rb : 0.0;
                    ps = th * z;
                } else {
                    const double sigma = Rb / eps, bar = Q.theta_bar[off - Q.off0 + tb];
                    th = fmin(fmax(bar + sigma * z, 0.0), Rb);
                    ps = th * z - (th - bar) * (th - bar) / (2.0 * sigma);
                    curv = (th > 0.0 && th < Rb) ? sigma : 0.0;
                }
                L[tb] = th;
                D[ta] = (r * th - ps) / gam;
                hc += curv * nb * r;
            }
            if (store_fill) { Q.theta_new[off - Q.off0] = L[0]; Q.theta_new[off - Q.off0 + 1] = L[1]; }
            if (Hs && hc != 0.0) {
                const int t0 = P.tok[off], t1 = P.tok[off + 1];
                (*Hs)[t0 * n + t0] += hc; (*Hs)[t1 * n + t1] += hc;
                (*Hs)[t0 * n + t1] -= hc; (*Hs)[t1 * n + t0] -= hc;
            }
        }
        for (int j = 0; j < k; ++j) {
            const double y = L[j] - D[j];
            const int t = P.tok[off + j];
            psi[t] += y;
            arb += nu[t] * y;
            if (trades && Q.delta) { Q.delta[off + j] = D[j]; Q.lam[off + j] = L[j]; }
        }
    }
#ifdef __CUDA_ARCH__
    if (LANES > 1) {
        arb = lanes_sum<LANES>(arb);
        for (int j = 0; j < n; ++j) psi[j] = lanes_sum<LANES>(psi[j]);
        if (Hs) for (int e = 0; e < n * n; ++e) (*Hs)[e] = lanes_sum<LANES>((*Hs)[e]);
        if (store_fill) {
            const int64_t nnz = P.pool_ptr[Q.p1] - Q.off0;
            for (int64_t x = 0; x < nnz; ++x) Q.theta_new[x] = lanes_sum<LANES>(Q.theta_new[x]);
        }
    }
#endif
    return arb;
}

CFMM_HD inline double dual_value(const Problem& Q, const Vec& nu, double arb) {
    double g = arb;
    for (int j = 0; j < Q.n; ++j) g += (nu[j] - Q.c[j]) * Q.a[j];
    return g;
}

// KKT residual of the box-constrained dual: err = sum_free |nu_j (a_j + psi_j)| / |g|; fills grad / pg / free mask.
CFMM_HD inline double kkt(const Problem& Q, const Vec& nu, const Vec& psi, const Vec& lb, double g, double err_prev,
                          const Vec& grad, const Vec& pg, uint64_t* free_mask) {
    const double ep = isfinite(err_prev) ? err_prev : 1e-2;
    const double thr = fmin(1e-2, fmax(1e-3 * ep, 1e-14));       // active-set width: 1e-3 x the KKT residual (see solver.py)
    double num = 0.0, wsum = 0.0, gmax = 0.0, scl = 0.0;
    uint64_t fm = 0;
    for (int j = 0; j < Q.n; ++j) {
        const double gr = Q.a[j] + psi[j];
        const bool near = nu[j] <= lb[j] * (1.0 + thr) && !is_eq(Q, j);
        const bool fr = !(is_pinned(Q, j) || (near && gr > 0.0));
        const double v = fr ? nu[j] * gr : 0.0;
        grad[j] = gr; pg[j] = v;
        if (fr) fm |= (uint64_t)1 << j;
        num += fabs(v);
        wsum += nu[j] * fabs(gr);
        if (fr) gmax = fmax(gmax, fabs(gr));
        scl = fmax(scl, fmax(fabs(Q.a[j]), is_pinned(Q, j) ? 0.0 : fabs(psi[j])));   // scale: constrained flows only
    }
    *free_mask = fm;
    // max of the value-weighted residual and the per-token one (the reference constrains psi token by token,
    // liquidation.py:77-80 / arbitrage.py:77: a cheap token must not hide a large residual behind its price)
    return fmax(num / fmax(fmax(fabs(g), 1e-3 * wsum), TINY), gmax / fmax(scl, TINY));
}

// Solve (Hs[free,free] + mu dbar I) x = -pg[free] (dbar = mean diagonal) by Gaussian elimination with partial pivoting;
// dt = 0 off the free set.  Returns 0 = descent direction, 1 = solved but not a descent direction, 2 = singular / not finite;
// *big = largest |x_j| (the caller climbs the damping ladder while it is not a sane log-price change).
CFMM_HD inline int newton_direction(int n, uint64_t free_mask, const Vec& Hs, const Vec& pg, const Vec& A,
                                     const Vec& dt, double mu, double* big) {
    int fidx[NTOK_MAX];
    int nf = 0;
    for (int j = 0; j < n; ++j) { dt[j] = 0.0; if (free_mask >> j & 1) fidx[nf++] = j; }
    *big = 0.0;
    if (nf == 0) return 1;
    double tr = 0.0;
    for (int x = 0; x < nf; ++x) tr += Hs[fidx[x] * n + fidx[x]];
    const double reg = mu * fmax(tr / nf, TINY);
    const int ld = nf + 1;                                                  // augmented [A | rhs], row-major in A
    for (int x = 0; x < nf; ++x) {
        for (int y = 0; y < nf; ++y) A[x * ld + y] = Hs[fidx[x] * n + fidx[y]] + (x == y ? reg : 0.0);
        A[x * ld + nf] = -pg[fidx[x]];
    }
    for (int col = 0; col < nf; ++col) {
        int piv = col;
        double best = fabs(A[col * ld + col]);
        for (int r = col + 1; r < nf; ++r) { const double v = fabs(A[r * ld + col]); if (v > best) { best = v; piv = r; } }
        if (!(best > 0.0) || !isfinite(best)) return 2;
        if (piv != col)
            for (int y = col; y <= nf; ++y) { const double t = A[col * ld + y]; A[col * ld + y] = A[piv * ld + y]; A[piv * ld + y] = t; }
        const double inv = 1.0 / A[col * ld + col];
        for (int r = col + 1; r < nf; ++r) {
            const double f = A[r * ld + col] * inv;
            if (f != 0.0) for (int y = col + 1; y <= nf; ++y) A[r * ld + y] -= f * A[col * ld + y];
        }
    }
    double slope = 0.0;
    bool finite = true;
    for (int x = nf - 1; x >= 0; --x) {
        double v = A[x * ld + nf];
        for (int y = x + 1; y < nf; ++y) v -= A[x * ld + y] * dt[fidx[y]];
        v /= A[x * ld + x];
        dt[fidx[x]] = v;
        finite = finite && isfinite(v);
        *big = fmax(*big, fabs(v));
    }
    for (int x = 0; x < nf; ++x) slope += pg[fidx[x]] * dt[fidx[x]];
    if (!finite) return 2;
    return slope < 0.0 ? 0 : 1;
}

// Workspace elements one problem needs (doubles): 12 n-vectors, 2 Hessians, the augmented system, 2 multiplier sets.
CFMM_HD inline int64_t work_doubles(int n, int64_t nnz) { return 12LL * n + 2LL * n * n + (int64_t)n * (n + 1) + 2 * nnz; }

// The solve.  nu_io [n]: start prices in, optimal prices out.  psi_out [n].  `work`/`stride`: interleaved workspace.
template <int LANES = 1>
CFMM_HD inline Stats solve_one(const Pools& P, Problem Q, const Params& prm, double* nu_io, double* psi_out,
                               double* work, int64_t stride, int lane = 0) {
    const int n = Q.n;
    const int64_t nnz = P.pool_ptr[Q.p1] - Q.off0;
    int64_t e = 0;
    auto vec = [&](int64_t len) { Vec v{work + e * stride, stride}; e += len; return v; };
    Vec nuv[2] = {vec(n), vec(n)}, psiv[2] = {vec(n), vec(n)};
    Vec grad = vec(n), pg = vec(n), dt = vec(n), lb = vec(n), lognu = vec(n), grad_t = vec(n), pg_t = vec(n), spare = vec(n);
    Vec Hsv[2] = {vec(n * n), vec(n * n)};
    Vec A = vec((int64_t)n * (n + 1));
    Q.theta_bar = vec(nnz);
    Q.theta_new = vec(nnz);
    (void)spare;

    bool has_sum = false, bad = n < 1 || n > NTOK_MAX;
    for (int64_t i = Q.p0; i < Q.p1 && !bad; ++i) {                         // refuse what the closed forms do not cover
        const int64_t o = P.pool_ptr[i];
        const int k = (int)(P.pool_ptr[i + 1] - o);
        has_sum = has_sum || P.kind[i] == 1;
        bad = P.kind[i] > 3 || k < 2 || k > KMAX || ((P.kind[i] == 1 || P.kind[i] == 3) && k != 2);
        for (int j = 0; j < k && !bad; ++j) bad = P.tok[o + j] < 0 || P.tok[o + j] >= n;
    }
    if (bad) {
        Stats st;
        st.value = st.dual = st.gap = st.infeas = st.err = NAN;
        st.iters = st.evals = 0; st.status = 3;                             // 3 = rejected input
        return st;
    }
    for (int64_t x = 0; x < nnz; ++x) { Q.theta_bar[x] = 0.0; Q.theta_new[x] = 0.0; }

    double scale = 1.0;
    for (int j = 0; j < n; ++j) scale = fmax(scale, fabs(Q.c[j]));
    const double floor_ = prm.floor_rel * scale;
    int cur = 0;
    for (int j = 0; j < n; ++j) {
        lb[j] = is_eq(Q, j) ? floor_ : fmax(Q.c[j], floor_);
        nuv[0][j] = is_pinned(Q, j) ? Q.c[j] : fmax(nu_io[j], lb[j]);
    }

    int evals = 0, iters = 0, status = 1;                                   // 0 optimal, 1 max_iter, 2 stalled
    double eps_t = has_sum ? prm.eps0 : 0.0;
    double err = INFINITY, move = 1.0, g = 0.0;
    bool failed_before = false;
    uint64_t free_mask = 0, fm_t = 0;

    for (int outer = 0; outer < prm.max_outer; ++outer) {
        g = dual_value(Q, nuv[cur], evaluate<LANES>(P, Q, nuv[cur], lognu, eps_t, psiv[cur], &Hsv[cur], false, false, lane));
        ++evals;
        int inner_status = 1;
        const double inner_tol = has_sum ? fmax(prm.tol, fmin(1e-3, 1e-2 * move)) : prm.tol;
        err = kkt(Q, nuv[cur], psiv[cur], lb, g, err, grad, pg, &free_mask);
        for (int it = 0; it < prm.max_inner; ++it) {
            ++iters;
#ifdef CFMM_SMALL_TRACE
            printf("outer=%d it=%d g=%.15g err=%.3e free=%d\n", outer, iters, g, err, __builtin_popcountll(free_mask));
#endif
            if (err <= inner_tol) { inner_status = 0; break; }
            // (near-)singular free-set systems (every pool tying some free prices to the rest saturated) give an enormous
            // step along the null directions: climb the damping ladder (Levenberg-Marquardt shift mu * mean diagonal)
            // until the step is a sane price change; the null directions then get a scaled gradient step
            const double mus[6] = {1e-14, 1e-8, 1e-6, 1e-4, 1e-2, 1.0};
            const int nxt = cur ^ 1;
            double alpha = 1.0, g_t = g;
            bool ok = false;
            for (int rung = 0;;) {
                int code = 2;
                for (;; ++rung) {                                        // climb until the step is a sane price change
                    double big = 0.0;
                    code = newton_direction(n, free_mask, Hsv[cur], pg, A, dt, mus[rung], &big);
                    if ((code != 2 && big <= DT_MAX) || rung == 5) break;
                }
                if (code != 0) {                                         // fall back to scaled steepest descent
                    double mx = 0.0;
                    for (int j = 0; j < n; ++j) mx = fmax(mx, fabs(pg[j]));
                    mx = fmax(mx, TINY);
                    for (int j = 0; j < n; ++j) dt[j] = -pg[j] / mx;
                }
                {                                                        // still too long after the largest shift:
                    double big = 0.0;                                    // keep the direction, bound the step
                    for (int j = 0; j < n; ++j) big = fmax(big, fabs(dt[j]));
                    if (big > DT_MAX) for (int j = 0; j < n; ++j) dt[j] *= DT_MAX / big;
                }
                alpha = 1.0;
                double lin1 = 0.0;                                       // predicted decrease of the FULL step
                for (int ls = 0; ls < 50; ++ls) {
                    double lin = 0.0;
                    for (int j = 0; j < n; ++j) {
                        const double st = fmin(fmax(alpha * dt[j], -20.0), 20.0);
                        const double v = is_pinned(Q, j) ? Q.c[j] : fmax(nuv[cur][j] * exp(st), lb[j]);
                        nuv[nxt][j] = v;
                        lin += grad[j] * (v - nuv[cur][j]);
                    }
                    g_t = dual_value(Q, nuv[nxt], evaluate<LANES>(P, Q, nuv[nxt], lognu, eps_t, psiv[nxt], &Hsv[nxt], false, false, lane));
                    ++evals;
                    if (ls == 0) lin1 = lin;
                    if (g_t <= g + 1e-4 * lin) { ok = true; break; }
                    if (fabs(g_t - g) <= 1e-13 * fabs(g) || fabs(lin1) <= 1e-9 * fabs(g)) {  // g cannot resolve this step
                        if (kkt(Q, nuv[nxt], psiv[nxt], lb, g_t, err, grad_t, pg_t, &fm_t) < 0.99 * err) { ok = true; break; }
                        if (alpha < 1e-3) break;
                    }
                    alpha *= 0.5;
                }
                // a failed search along a barely damped direction (null-space dominated: long step, no predicted gain):
                // damp harder and try again before giving up
                if (ok || rung == 5) break;
                ++rung;
            }
#ifdef CFMM_SMALL_TRACE
            { double mx = 0; int jm = 0; for (int j = 0; j < n; ++j) if (fabs(pg[j]) > mx) { mx = fabs(pg[j]); jm = j; }
              printf("   ok=%d alpha=%.3e g_t-g=%.3e maxpg=%.3e at %d dt=%.3e nu=%.17g lb=%.17g grad=%.3e mask=%llx\n", ok, alpha, g_t - g, mx, jm, dt[jm], nuv[cur][jm], lb[jm], grad[jm], (unsigned long long)free_mask); }
#endif
            if (!ok) { inner_status = 2; break; }
            cur = nxt; g = g_t;
            err = kkt(Q, nuv[cur], psiv[cur], lb, g, err, grad, pg, &free_mask);
        }
        if (!has_sum) { status = inner_status; break; }
        // exact duality gap at the current prices (trades from the smoothed problem, dual with eps = 0)
        evaluate<LANES>(P, Q, nuv[cur], lognu, eps_t, psiv[cur ^ 1], nullptr, false, true, lane);
        const double g_exact = dual_value(Q, nuv[cur], evaluate<LANES>(P, Q, nuv[cur], lognu, 0.0, grad_t, nullptr, false, false, lane));
        evals += 2;
        double primal = 0.0;
        for (int j = 0; j < n; ++j) primal += Q.c[j] * psiv[cur ^ 1][j];
        const double gap_now = (g_exact - primal) / fmax(fabs(g_exact), TINY);
        if (inner_status == 0 && err <= prm.tol && fabs(gap_now) <= prm.tol) { status = 0; break; }   // the only certified exit
        status = inner_status != 0 ? inner_status : 1;
        // the ramp cannot get narrower and the inner solve failed twice in a row: fp64 resolution of the price
        // ratio / eps bounds the reachable residual, more passes would not help
        if (inner_status != 0 && failed_before && eps_t <= prm.eps_min) break;
        failed_before = inner_status != 0;
        move = 0.0;
        for (int64_t i = Q.p0; i < Q.p1; ++i) {
            if (P.kind[i] != 1) continue;
            const int64_t o = P.pool_ptr[i] - Q.off0;
            for (int b = 0; b < 2; ++b) {
                move = fmax(move, fabs(Q.theta_new[o + b] - Q.theta_bar[o + b]) / P.R[P.pool_ptr[i] + b]);
                Q.theta_bar[o + b] = Q.theta_new[o + b];
            }
        }
        eps_t = fmax(prm.eps_min, eps_t * prm.eps_shrink);
    }

    // final read-out: trades and psi from the (smoothed) problem, dual value from the exact one
    const Vec& psi_f = psiv[cur ^ 1];
    evaluate<LANES>(P, Q, nuv[cur], lognu, eps_t, psi_f, nullptr, true, false, lane);
    const double dval = dual_value(Q, nuv[cur], evaluate<LANES>(P, Q, nuv[cur], lognu, 0.0, grad_t, nullptr, false, false, lane));
    evals += 2;
    double primal = 0.0, viol = 0.0;
    for (int j = 0; j < n; ++j) {
        const double s = psi_f[j] + Q.a[j];
        const double v = is_pinned(Q, j) ? 0.0 : (is_eq(Q, j) ? fabs(s) : fmax(-s, 0.0));
        viol += nuv[cur][j] * v;
        primal += Q.c[j] * psi_f[j];
        if (LANES == 1 || lane == 0) { nu_io[j] = nuv[cur][j]; psi_out[j] = psi_f[j]; }
    }
    Stats st;
    st.value = primal; st.dual = dval;
    st.gap = (dval - primal) / fmax(fabs(dval), TINY);
    st.infeas = viol / fmax(fabs(dval), TINY);
    st.err = err; st.iters = iters; st.evals = evals; st.status = status;
    return st;
}

}  // namespace cfmm_small
