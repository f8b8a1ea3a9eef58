/* cfmm_oracle_c.c -- CPU restatement (plain C + pthreads) of one dual evaluation, TEST INFRASTRUCTURE ONLY.
 *
 * Same math as oracle/cfmm_oracle.py::_geomean_group (constant-product branch), which restates the reference's
 * Uniswap-v2 constraint cp.geo_mean(new_reserves) >= cp.geo_mean(reserves) (arbitrage.py:68-70) with new_reserves =
 * R + gamma*D - L (arbitrage.py:60) and psi = sum_i A_i (L_i - D_i) (arbitrage.py:54).  Used only as the multi-threaded
 * CPU baseline of bench.py (cpu_baseline / --impl reference) and cross-checked against the numpy oracle in
 * tests/test_oracle.py.  The product never links or loads this file.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
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

void oracle_set_threads(int n) { g_threads = (n < 1) ? 1 : (n > 256 ? 256 : n); }

typedef struct {
    int64_t lo, hi;
    const int32_t* idx; const double* R; const double* gamma; const double* nu;
    int32_t n; double* priv; double* delta; double* lam; double* hcoef;
    const double* vt; int mode;      /* mode 0: evaluation, 1: y += Hs vt, 2: diag += diag(Hs) */
} job_t;

static void* worker(void* arg) {
    job_t* J = (job_t*)arg;
    double* my = J->priv;
    const int32_t n = J->n;
    if (J->mode != 0) {               /* Hessian in log-price coordinates: Hs_i = h_i [[1,-1],[-1,1]] */
        for (int64_t i = J->lo; i < J->hi; ++i) {
            const double h = J->hcoef[i];
            if (h == 0.0) continue;
            const int32_t i0 = J->idx[2 * i], i1 = J->idx[2 * i + 1];
            if (J->mode == 1) { const double c = h * (J->vt[i0] - J->vt[i1]); my[i0] += c; my[i1] -= c; }
            else { my[i0] += h; my[i1] += h; }
        }
        return 0;
    }
    for (int64_t i = J->lo; i < J->hi; ++i) {
        const int32_t i0 = J->idx[2 * i], i1 = J->idx[2 * i + 1];
        const double R0 = J->R[2 * i], R1 = J->R[2 * i + 1], g = J->gamma[i];
        const double n0 = J->nu[i0], n1 = J->nu[i1];
        const double p0 = n0 * R0, p1 = n1 * R1;
        double D0 = 0, D1 = 0, L0 = 0, L1 = 0;
        if (g * p1 > p0) {              /* tender token 0, receive token 1 */
            const double t = sqrt(g * p1 / p0);
            D0 = R0 * (t - 1.0) / g; L1 = R1 * (1.0 - 1.0 / t);
        } else if (g * p0 > p1) {
            const double t = sqrt(g * p0 / p1);
            D1 = R1 * (t - 1.0) / g; L0 = R0 * (1.0 - 1.0 / t);
        }
        const double y0 = L0 - D0, y1 = L1 - D1;
        if (J->hcoef) J->hcoef[i] = (y0 != 0.0 || y1 != 0.0) ? 0.5 * sqrt(p0 * p1 / g) : 0.0;
        my[i0] += y0; my[i1] += y1; my[n] += n0 * y0 + n1 * y1;
        if (J->delta) {
            J->delta[2 * i] = D0; J->delta[2 * i + 1] = D1; J->lam[2 * i] = L0; J->lam[2 * i + 1] = L1;
        }
    }
    return 0;
}

static int run_jobs(int64_t m, const int32_t* idx, const double* R, const double* gamma, int32_t n, const double* nu,
                    double* out, double* arb, double* delta, double* lam, double* hcoef, const double* vt, int mode) {
    const int nt = oracle_num_threads();
    double* priv = (double*)calloc((size_t)nt * (size_t)(n + 1), sizeof(double));
    job_t* jobs = (job_t*)calloc((size_t)nt, sizeof(job_t));
    pthread_t* th = (pthread_t*)calloc((size_t)nt, sizeof(pthread_t));
    if (!priv || !jobs || !th) { free(priv); free(jobs); free(th); return -1; }
    for (int t = 0; t < nt; ++t) {
        job_t j = {m * t / nt, m * (t + 1) / nt, idx, R, gamma, nu, n, priv + (size_t)t * (size_t)(n + 1), delta, lam,
                   hcoef, vt, mode};
        jobs[t] = j;
        if (t > 0) pthread_create(&th[t], 0, worker, &jobs[t]);
    }
    worker(&jobs[0]);
    for (int t = 1; t < nt; ++t) pthread_join(th[t], 0);
    memset(out, 0, sizeof(double) * (size_t)n);
    if (arb) *arb = 0.0;
    for (int t = 0; t < nt; ++t) {
        const double* my = priv + (size_t)t * (size_t)(n + 1);
        for (int32_t j = 0; j < n; ++j) out[j] += my[j];
        if (arb) *arb += my[n];
    }
    free(priv); free(jobs); free(th);
    return 0;
}

/* idx: [m][2] int32, R: [m][2] f64, gamma: [m]; nu: [n]; psi: [n] out; arb: [1] out; delta/lam: [m][2] out or NULL;
 * hcoef: [m] out or NULL (curvature coefficient of each pool in log-price coordinates) */
int oracle_eval_pairs(int64_t m, const int32_t* idx, const double* R, const double* gamma, int32_t n, const double* nu,
                      double* psi, double* arb, double* delta, double* lam, double* hcoef) {
    return run_jobs(m, idx, R, gamma, n, nu, psi, arb, delta, lam, hcoef, 0, 0);
}

/* y = Hs vt  (mode 1)  /  diag = diag(Hs)  (mode 2), Hs = sum_i A_i hcoef_i [[1,-1],[-1,1]] A_i' */
int oracle_hess_pairs(int64_t m, const int32_t* idx, const double* hcoef, int32_t n, const double* vt, double* out,
                      int mode) {
    return run_jobs(m, idx, 0, 0, n, 0, out, 0, 0, 0, (double*)hcoef, vt, mode == 2 ? 2 : 1);
}
