/* cfmm_oracle_c.c -- CPU restatement (plain C + pthreads) of the constant-product hot path, TEST INFRASTRUCTURE ONLY.
 *
 * Same math as oracle/cfmm_oracle.py::_geomean_group (constant-product branch), which restates the reference's
 * Uniswap-v2 constraint cp.geo_mean(new_reserves) >= cp.geo_mean(reserves) (arbitrage.py:68-70) with new_reserves =
 * R + gamma*D - L (arbitrage.py:60) and psi = sum_i A_i (L_i - D_i) (arbitrage.py:54); oracle_solve_pairs() restates
 * oracle/cfmm_oracle.py::solve (what replaces prob.solve(), arbitrage.py:81-82) with the Newton system solved by
 * Jacobi-preconditioned CG on Hessian-vector products instead of a dense factorisation, so that it runs at BASELINE
 * configs[4] size (1M pools, 4096 tokens).  Used only as the multi-threaded CPU baseline of bench.py (cpu_baseline /
 * --impl reference) and cross-checked against the numpy oracle in tests/test_oracle.py.  The product never links or
 * loads this file.
 *
 * Threads: a persistent pool (created once per thread count, parked on a condition variable between passes), one
 * private psi array per worker allocated once per token count -- a pass costs no thread creation and no allocation.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

static int g_threads = 0;

int oracle_num_threads(void) {
    if (g_threads <= 0) {
        const char* e = getenv("ORACLE_THREADS");
        long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
        g_threads = (n < 1) ? 1 : (n > 256 ? 256 : (int)n);
    }
    return g_threads;
}

typedef struct {
    int64_t m;
    const int32_t* idx; const double* R; const double* gamma; const double* nu;
    int32_t n; double* delta; double* lam; double* hcoef;
    const double* vt; int mode;      /* mode 0: evaluation, 1: y += Hs vt, 2: diag += diag(Hs) */
} pass_t;

/* ---- persistent pool ------------------------------------------------------------------------------------------ */
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_go = PTHREAD_COND_INITIALIZER, g_done = PTHREAD_COND_INITIALIZER;
static pthread_t g_th[256];
static int g_pool = 0;               /* workers alive (ids 1..g_pool-1; the caller is worker 0) */
static uint64_t g_gen = 0;           /* pass generation */
static uint64_t g_gen0 = 0;          /* g_gen when the current pool was created */
static int g_left = 0, g_quit = 0;
static pass_t g_pass;
static double* g_priv = 0;           /* [g_pool][g_priv_n + 1] */
static int32_t g_priv_n = -1;
static int g_priv_nt = 0;

static void do_slice(const pass_t* J, int t, int nt) {
    double* my = g_priv + (size_t)t * (size_t)(J->n + 1);
    const int32_t n = J->n;
    const int64_t lo = J->m * t / nt, hi = J->m * (t + 1) / nt;
    memset(my, 0, sizeof(double) * (size_t)(n + 1));
    if (J->mode != 0) {               /* Hessian in log-price coordinates: Hs_i = h_i [[1,-1],[-1,1]] */
        for (int64_t i = lo; i < hi; ++i) {
            const double h = J->hcoef[i];
            if (h == 0.0) continue;
            const int32_t i0 = J->idx[2 * i], i1 = J->idx[2 * i + 1];
            if (J->mode == 1) { const double c = h * (J->vt[i0] - J->vt[i1]); my[i0] += c; my[i1] -= c; }
            else { my[i0] += h; my[i1] += h; }
        }
        return;
    }
    for (int64_t i = lo; i < hi; ++i) {
        const int32_t i0 = J->idx[2 * i], i1 = J->idx[2 * i + 1];
        const double R0 = J->R[2 * i], R1 = J->R[2 * i + 1], g = J->gamma[i];
        const double n0 = J->nu[i0], n1 = J->nu[i1];
        const double p0 = n0 * R0, p1 = n1 * R1;
        double D0 = 0, D1 = 0, L0 = 0, L1 = 0;
        if (g * p1 > p0) {              /* tender token 0, receive token 1 */
            const double t_ = sqrt(g * p1 / p0);
            D0 = R0 * (t_ - 1.0) / g; L1 = R1 * (1.0 - 1.0 / t_);
        } else if (g * p0 > p1) {
            const double t_ = sqrt(g * p0 / p1);
            D1 = R1 * (t_ - 1.0) / g; L0 = R0 * (1.0 - 1.0 / t_);
        }
        const double y0 = L0 - D0, y1 = L1 - D1;
        if (J->hcoef) J->hcoef[i] = (y0 != 0.0 || y1 != 0.0) ? 0.5 * sqrt(p0 * p1 / g) : 0.0;
        my[i0] += y0; my[i1] += y1; my[n] += n0 * y0 + n1 * y1;
        if (J->delta) {
            J->delta[2 * i] = D0; J->delta[2 * i + 1] = D1; J->lam[2 * i] = L0; J->lam[2 * i + 1] = L1;
        }
    }
}

static void* pool_worker(void* arg) {
    const int t = (int)(intptr_t)arg;
    uint64_t seen = g_gen0;          /* generation at pool creation: passes posted before that are not ours */
    pthread_mutex_lock(&g_mu);
    for (;;) {
        while (g_gen == seen && !g_quit) pthread_cond_wait(&g_go, &g_mu);
        if (g_quit) break;
        seen = g_gen;
        const int nt = g_pool;
        pthread_mutex_unlock(&g_mu);
        do_slice(&g_pass, t, nt);
        pthread_mutex_lock(&g_mu);
        if (--g_left == 0) pthread_cond_signal(&g_done);
    }
    pthread_mutex_unlock(&g_mu);
    return 0;
}

static void pool_stop(void) {
    if (g_pool <= 1) { g_pool = 0; return; }
    pthread_mutex_lock(&g_mu);
    g_quit = 1;
    pthread_cond_broadcast(&g_go);
    pthread_mutex_unlock(&g_mu);
    for (int t = 1; t < g_pool; ++t) pthread_join(g_th[t], 0);
    g_quit = 0; g_pool = 0;
}

void oracle_set_threads(int n) {
    n = (n < 1) ? 1 : (n > 256 ? 256 : n);
    if (n != g_threads) { pool_stop(); g_threads = n; }
}

static int pool_ready(int32_t n) {
    const int nt = oracle_num_threads();
    if (g_pool != nt) {
        pool_stop();
        g_pool = nt;
        g_gen0 = g_gen;
        for (int t = 1; t < nt; ++t)
            if (pthread_create(&g_th[t], 0, pool_worker, (void*)(intptr_t)t)) { g_pool = t; pool_stop(); return -1; }
    }
    if (g_priv_n != n || g_priv_nt != nt) {
        free(g_priv);
        g_priv = (double*)malloc(sizeof(double) * (size_t)nt * (size_t)(n + 1));
        if (!g_priv) { g_priv_n = -1; return -1; }
        g_priv_n = n; g_priv_nt = nt;
    }
    return 0;
}

static int run_pass(const pass_t* P, double* out, double* arb) {
    if (pool_ready(P->n)) return -1;
    const int nt = g_pool;
    pthread_mutex_lock(&g_mu);
    g_pass = *P;
    g_left = nt - 1;
    ++g_gen;
    pthread_cond_broadcast(&g_go);
    pthread_mutex_unlock(&g_mu);
    do_slice(P, 0, nt);
    pthread_mutex_lock(&g_mu);
    while (g_left > 0) pthread_cond_wait(&g_done, &g_mu);
    pthread_mutex_unlock(&g_mu);
    const int32_t n = P->n;
    memset(out, 0, sizeof(double) * (size_t)n);
    if (arb) *arb = 0.0;
    for (int t = 0; t < nt; ++t) {              /* fixed order: reproducible for a given thread count */
        const double* my = g_priv + (size_t)t * (size_t)(n + 1);
        for (int32_t j = 0; j < n; ++j) out[j] += my[j];
        if (arb) *arb += my[n];
    }
    return 0;
}

/* idx: [m][2] int32, R: [m][2] f64, gamma: [m]; nu: [n]; psi: [n] out; arb: [1] out; delta/lam: [m][2] out or NULL;
 * hcoef: [m] out or NULL (curvature coefficient of each pool in log-price coordinates) */
int oracle_eval_pairs(int64_t m, const int32_t* idx, const double* R, const double* gamma, int32_t n, const double* nu,
                      double* psi, double* arb, double* delta, double* lam, double* hcoef) {
    pass_t P = {m, idx, R, gamma, nu, n, delta, lam, hcoef, 0, 0};
    return run_pass(&P, psi, arb);
}

/* y = Hs vt  (mode 1)  /  diag = diag(Hs)  (mode 2), Hs = sum_i A_i hcoef_i [[1,-1],[-1,1]] A_i' */
int oracle_hess_pairs(int64_t m, const int32_t* idx, const double* hcoef, int32_t n, const double* vt, double* out,
                      int mode) {
    pass_t P = {m, idx, 0, 0, 0, n, 0, 0, (double*)hcoef, vt, mode == 2 ? 2 : 1};
    return run_pass(&P, out, 0);
}

/* ---- the whole solve: projected Newton in log-price coordinates, Jacobi-PCG on Hessian-vector products ------------
 * minimise g(nu) = (nu - c)'a + sum_i arb_i(A_i' nu) over the box nu_j >= c_j (inequality tokens), nu_j > 0 free
 * (eq[j]), nu_j = c_j (pinned[j]); same algorithm, constants and stopping rule as oracle/cfmm_oracle.py::solve
 * restricted to constant-product pools (one outer pass: no constant-sum kinks), CG instead of a dense solve. */
typedef struct {
    double dual_value, primal_value, gap, primal_infeas, err, wall_s;
    int32_t iters, evals, hvps, status;      /* status 0 optimal, 1 max_iter, 2 stalled */
} oracle_solve_result;

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

typedef struct {
    int32_t n; const double *c, *a, *lb; const uint8_t *eq, *pinned;
    double a_inf;
} spec_t;

/* KKT residual (max of the value-weighted and the per-token one) + gradient / free set / projected gradient */
static double kkt_c(const spec_t* S, const double* nu, const double* psi, double g, double err_prev, double* grad,
                    double* fr, double* pg) {
    const double ep = isfinite(err_prev) ? err_prev : 1e-2;
    const double thr = fmin(1e-2, fmax(1e-3 * ep, 1e-14));       // active-set width: 1e-3 x the KKT residual (see solver.py)
    double num = 0, wsum = 0, gmax = 0, scl = S->a_inf;
    for (int32_t j = 0; j < S->n; ++j) {
        const double gr = S->a[j] + psi[j];
        const int near = nu[j] <= S->lb[j] * (1.0 + thr) && !S->eq[j];
        const int fre = !(S->pinned[j] || (near && gr > 0.0));
        grad[j] = gr; fr[j] = fre ? 1.0 : 0.0; pg[j] = fre ? nu[j] * gr : 0.0;
        num += fabs(pg[j]); wsum += nu[j] * fabs(gr);
        if (fre) gmax = fmax(gmax, fabs(gr));
        if (!S->pinned[j]) scl = fmax(scl, fabs(psi[j]));
    }
    return fmax(num / fmax(fmax(fabs(g), 1e-3 * wsum), 1e-300), gmax / fmax(scl, 1e-300));
}

int oracle_solve_pairs(int64_t m, const int32_t* idx, const double* R, const double* gamma, int32_t n, const double* c,
                       const double* a, const uint8_t* eq, const uint8_t* pinned, double* nu, double* psi_out,
                       double tol, int32_t max_iter, int32_t cg_max, oracle_solve_result* res) {
    const double t0 = now_s();
    double* buf = (double*)calloc((size_t)14 * (size_t)n + 2 + (size_t)m, sizeof(double));
    if (!buf) return -1;
    double *lb = buf, *psi = lb + n, *psit = psi + n, *nut = psit + n, *grad = nut + n, *fr = grad + n, *pg = fr + n,
           *dt = pg + n, *x = dt + n, *r = x + n, *z = r + n, *p = z + n, *diag = p + n, *y = diag + n, *hcoef = y + n + 2;
    double scale = 1.0, a_inf = 0.0;
    for (int32_t j = 0; j < n; ++j) { scale = fmax(scale, fabs(c[j])); a_inf = fmax(a_inf, fabs(a[j])); }
    const double floor_ = 1e-12 * scale;
    for (int32_t j = 0; j < n; ++j) {
        lb[j] = eq[j] ? floor_ : fmax(c[j], floor_);
        nu[j] = pinned[j] ? c[j] : fmax(nu[j], lb[j]);
    }
    spec_t S = {n, c, a, lb, eq, pinned, a_inf};
    int evals = 0, hvps = 0, iters = 0, status = 1, rc = 0;
    double arb = 0, g = 0, err = INFINITY;
#define DUAL(nu_, arb_) ({ double s_ = (arb_); for (int32_t j_ = 0; j_ < n; ++j_) s_ += ((nu_)[j_] - c[j_]) * a[j_]; s_; })
    rc |= oracle_eval_pairs(m, idx, R, gamma, n, nu, psi, &arb, 0, 0, hcoef); ++evals;
    g = DUAL(nu, arb);
    err = kkt_c(&S, nu, psi, g, err, grad, fr, pg);
    for (; iters < max_iter && !rc;) {
        ++iters;
        if (err <= tol) { status = 0; break; }
        /* Newton direction: Hs dt = -pg on the free set, truncated Jacobi-PCG */
        rc |= oracle_hess_pairs(m, idx, hcoef, n, 0, diag, 2);
        double rz = 0;
        for (int32_t j = 0; j < n; ++j) {
            x[j] = 0.0; r[j] = -pg[j]; z[j] = fr[j] / fmax(diag[j], 1e-300) * r[j]; p[j] = z[j]; rz += r[j] * z[j];
        }
        const double r0 = sqrt(fmax(rz, 0.0)), eta = fmin(0.1, sqrt(err));
        for (int k = 0; k < cg_max && rz > 0.0 && !rc; ++k) {
            rc |= oracle_hess_pairs(m, idx, hcoef, n, p, y, 1); ++hvps;
            double pHp = 0, pdp = 0;
            for (int32_t j = 0; j < n; ++j) { y[j] *= fr[j]; pHp += p[j] * y[j]; pdp += p[j] * p[j] * fmax(diag[j], 1e-300); }
            if (pHp <= 1e-14 * pdp) { if (k == 0) memcpy(x, p, sizeof(double) * (size_t)n); break; }   /* homogeneity direction */
            const double al = rz / pHp;
            double rzn = 0;
            for (int32_t j = 0; j < n; ++j) {
                x[j] += al * p[j]; r[j] -= al * y[j]; z[j] = fr[j] / fmax(diag[j], 1e-300) * r[j]; rzn += r[j] * z[j];
            }
            if (rzn <= 0.0 || sqrt(rzn) <= eta * r0) break;
            const double be = rzn / rz;
            for (int32_t j = 0; j < n; ++j) p[j] = z[j] + be * p[j];
            rz = rzn;
        }
        double slope = 0, mx = 0;
        for (int32_t j = 0; j < n; ++j) { slope += pg[j] * x[j]; mx = fmax(mx, fabs(pg[j])); }
        const int desc = isfinite(slope) && slope < 0.0;
        for (int32_t j = 0; j < n; ++j) dt[j] = desc ? x[j] : -pg[j] / fmax(mx, 1e-300);
        /* projected Armijo backtracking along nu * exp(alpha dt) */
        double alpha = 1.0, lin1 = 0.0, gt = g, arbt = 0;
        int ok = 0;
        for (int ls = 0; ls < 50 && !rc; ++ls) {
            double lin = 0;
            for (int32_t j = 0; j < n; ++j) {
                const double e = fmin(fmax(alpha * dt[j], -20.0), 20.0);
                nut[j] = pinned[j] ? c[j] : fmax(nu[j] * exp(e), lb[j]);
                lin += grad[j] * (nut[j] - nu[j]);
            }
            rc |= oracle_eval_pairs(m, idx, R, gamma, n, nut, psit, &arbt, 0, 0, hcoef); ++evals;
            gt = DUAL(nut, arbt);
            if (ls == 0) lin1 = lin;
            if (gt <= g + 1e-4 * lin) { ok = 1; break; }
            if (fabs(gt - g) <= 1e-13 * fabs(g) || fabs(lin1) <= 1e-9 * fabs(g)) {       /* g cannot resolve this step */
                if (kkt_c(&S, nut, psit, gt, err, r, z, p) < 0.99 * err) { ok = 1; break; }
                if (alpha < 1e-3) break;
            }
            alpha *= 0.5;
        }
        if (!ok) {                                    /* hcoef belongs to the last trial point: restore it */
            rc |= oracle_eval_pairs(m, idx, R, gamma, n, nu, psi, &arb, 0, 0, hcoef); ++evals;
            status = 2;
            break;
        }
        memcpy(nu, nut, sizeof(double) * (size_t)n); memcpy(psi, psit, sizeof(double) * (size_t)n);
        g = gt; arb = arbt;
        err = kkt_c(&S, nu, psi, g, err, grad, fr, pg);
    }
    if (status == 1 && err <= tol) status = 0;
    double primal = 0, viol = 0;
    for (int32_t j = 0; j < n; ++j) {
        const double s = psi[j] + a[j];
        viol += nu[j] * (pinned[j] ? 0.0 : (eq[j] ? fabs(s) : fmax(-s, 0.0)));
        primal += c[j] * psi[j];
        psi_out[j] = psi[j];
    }
    if (res) {
        res->dual_value = g; res->primal_value = primal; res->gap = (g - primal) / fmax(fabs(g), 1e-300);
        res->primal_infeas = viol / fmax(fabs(g), 1e-300); res->err = err; res->iters = iters; res->evals = evals;
        res->hvps = hvps; res->status = status; res->wall_s = now_s() - t0;
    }
#undef DUAL
    free(buf);
    return rc;
}
