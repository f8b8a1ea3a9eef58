// cfmm_persist.cu -- the WHOLE outer loop of a constant-product routing problem in ONE persistent kernel.
//
// What replaces `prob.solve()` (arbitrage.py:81-82) for problems made of one token-blocked constant-product bucket
// (arbitrage.py:68-70): the same projected Newton / Jacobi-PCG / Armijo method as cfmm_solver.cu and solver.py, but the
// host never sees an intermediate scalar.  One cooperative launch of kCtasPerSm CTAs per SM:
//   * every CTA owns a fixed contiguous chunk of tiles and runs the pool passes (evaluation, Hessian-vector product,
//     Hessian diagonal: blocked_pass of cfmm_blocked.cuh) on it, for every pass of the solve;
//   * CTA 0 additionally owns the n_token-sized vector algebra (KKT residual, free set, PCG recurrences, line search)
//     and decides what the grid does next; the decision travels as a command word guarded by an epoch flag;
//   * between a pass and CTA 0's work the CTAs arrive on a counter (release) that CTA 0 polls (acquire); between CTA 0's
//     work and the next pass they poll the epoch flag.  Two one-way signals per pass instead of two full grid barriers.
//   * pool-sharded over several GPUs: after each pass CTA 0 all-reduces the pass's vector over NVLink peer
//     memory (the LL protocol of cfmm_allreduce.cu: 16-byte {value, seq} pushes, sum in rank order) before it does its
//     vector algebra -- every rank computes bit-identical scalars, so all ranks take the same decisions.
// The host launches once and reads one result struct.
#include <math.h>
#include <string.h>

#include "cfmm_blocked.cuh"

using namespace cfmm;

namespace {

constexpr int PT = kTileT;                       // threads per CTA
constexpr unsigned kPollSleepNs = 200;
constexpr unsigned long long kSpinLimit = 6000000000ull;     // ~3 s of SM clocks: a lost peer / launch must not hang the GPU

enum { OP_DONE = 0, OP_EVAL = 1, OP_HVP = 2, OP_DIAG = 3 };
enum { ST_EVAL0 = 0, ST_DIAG, ST_HVP, ST_TRIAL };

struct DevResult {                               // written by CTA 0, copied to the host after the kernel
    double dual_value, primal_value, gap, primal_infeas, err;
    int iters, evals, hvps, status;              // status 0 optimal, 1 max_iter, 2 stalled, 3 aborted (spin limit)
    unsigned long long seq_acc, seq_vec;
    long long prof[16];                          // CTA 0's clock64 totals: pass eval/hvp/diag, wait for the grid, vector algebra eval/hvp/diag, (unused)
};

struct PersistArgs {
    BlockedArgs B;                               // layout + slabs; vec / out / hcoef are set per pass on the device
    int n;
    const double *c, *a;
    const unsigned char *eq, *fixed;
    double* nu[2];                               // current / trial prices (nu[0] = the caller's buffer)
    double* acc[2];                              // [psi | arb] of nu[0] / nu[1]
    double *y, *diag, *hcoef;
    double *lb, *grad[2], *fr[2], *pg[2], *dt, *x, *r, *z, *p, *minv;
    unsigned* ctl;                               // 512 B: [0] arrive counter | [32] abort | [64] epoch flag, [65] command, [66..67] beta (own 128-B lines)
    double tol, nu_floor;
    int max_iter, cg_max;
    LLCell* const* recv_acc;                     // pool-sharded: receive areas (device array of `world` pointers) or null
    LLCell* const* recv_vec;
    int rank, world;
    unsigned long long seq_acc, seq_vec;
    double* nu_out;                              // == nu[0]
    double* psi_out;
    DevResult* res;
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- CTA-wide reductions (PT threads) of NS sums and NM maxima at once (two barriers in total), broadcast to every
// thread; every thread adds the warp partials in the same order, so all threads hold the same bits
template <int NS, int NM>
__device__ __forceinline__ void cta_reduce(double (&s)[NS], double (&m)[NM > 0 ? NM : 1], double* sh /* [(NS + NM) * 16] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NS; ++k) s[k] = warp_sum(s[k]);
#pragma unroll
    for (int k = 0; k < NM; ++k) m[k] = warp_max(m[k]);
    __syncthreads();                                 // previous users of sh are done
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NS; ++k) sh[k * 16 + warp] = s[k];
#pragma unroll
        for (int k = 0; k < NM; ++k) sh[(NS + k) * 16 + warp] = m[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < PT / 32; ++w) t += sh[k * 16 + w];
        s[k] = t;
    }
#pragma unroll
    for (int k = 0; k < NM; ++k) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < PT / 32; ++w) t = fmax(t, sh[(NS + k) * 16 + w]);
        m[k] = t;
    }
}

// The vector loops of CTA 0 run over n ~ 4096 entries that live in L2 (other CTAs read / wrote them): every thread
// first issues the loads of U entries of every operand (independent, all in flight), then computes -- one L2 round trip
// per U * PT entries instead of one per entry.
constexpr int U = 4;
#define BOSS_CHUNKS(n) for (int base_ = 0; base_ < (n); base_ += PT * U)
#define BOSS_IDX(u) (base_ + (u) * PT + (int)threadIdx.x)

// KKT residual (max of the value-weighted and the per-token one, as solver.py::kkt) at (nu, acc); fills grad / fr / pg.
// nu_prev / grad_prev (nullable): also returns lin = grad_prev . (nu - nu_prev), the predicted change of the line search.
struct Kkt { double err, g, primal, infeas, lin; };
__device__ __noinline__ Kkt kkt_eval(const PersistArgs& S, const double* nu, const double* acc, double thr, double* grad, double* fr,
                        double* pg, const double* nu_prev, const double* grad_prev, double* sh) {
    const int n = S.n;
    // (pointer members are copied into registers once: S lives in the kernel-parameter window and, with stores in the
    // loops, the compiler would otherwise reload every pointer from there in every chunk)
    const double *Sa = S.a, *Sc = S.c, *Slb = S.lb;
    const unsigned char *Seq = S.eq, *Sfx = S.fixed;
    double s[6] = {0, 0, 0, 0, 0, 0}, m[2] = {0, 0};
    BOSS_CHUNKS(n) {
        double nj[U], pj[U], aj[U], cj[U], lbj[U], np[U], gp[U];
        unsigned char ej[U], fj[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = min(BOSS_IDX(u), n - 1);
            nj[u] = __ldcg(nu + j); pj[u] = __ldcg(acc + j); aj[u] = Sa[j]; cj[u] = Sc[j]; lbj[u] = __ldcg(Slb + j);
            ej[u] = Seq[j]; fj[u] = Sfx[j];
            np[u] = nu_prev ? __ldcg(nu_prev + j) : 0.0; gp[u] = grad_prev ? __ldcg(grad_prev + j) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = BOSS_IDX(u);
            if (j < n) {
                const double g = aj[u] + pj[u];
                const bool near = (nj[u] <= lbj[u] * (1.0 + thr)) && !ej[u];
                const bool act = fj[u] || (near && g > 0.0);
                const double f = act ? 0.0 : 1.0;
                const double v = nj[u] * g * f;
                grad[j] = g; fr[j] = f; pg[j] = v;
                s[0] += fabs(v);
                s[1] += (nj[u] - cj[u]) * aj[u];
                s[2] += nj[u] * fabs(g);
                s[3] += cj[u] * pj[u];
                s[4] += nj[u] * (fj[u] ? 0.0 : (ej[u] ? fabs(g) : fmax(-g, 0.0)));
                s[5] += gp[u] * (nj[u] - np[u]);
                m[0] = fmax(m[0], fabs(g) * f);
                m[1] = fmax(m[1], fmax(fabs(aj[u]), fj[u] ? 0.0 : fabs(pj[u])));
            }
        }
    }
    cta_reduce<6, 2>(s, m, sh);
    Kkt k;
    k.g = s[1] + __ldcg(acc + n);
    k.err = fmax(s[0] / fmax(fmax(fabs(k.g), 1e-3 * s[2]), 1e-300), m[0] / fmax(m[1], 1e-300));
    k.primal = s[3];
    k.infeas = s[4] / fmax(fabs(k.g), 1e-300);
    k.lin = s[5];
    return k;
}

// LL all-reduce of buf[0..len) by ONE CTA, in place (see cfmm_allreduce.cu for the protocol).  Returns false on time-out.
__device__ bool ll_allreduce_cta(double* buf, int len, LLCell* const* recv, int rank, int world, unsigned long long seq) {
    const long long slot = (long long)(seq % 3) * world * len;
    for (int j = threadIdx.x; j < len; j += PT) {
        const double mine = __ldcg(buf + j);
        for (int r = 0; r < world; ++r)
            if (r != rank) st_ll(recv[r] + slot + (long long)rank * len + j, mine, seq);
    }
    bool ok = true;
    const long long t0 = clock64();
    for (int j = threadIdx.x; j < len; j += PT) {
        double s = 0.0;
        for (int r = 0; r < world; ++r) {               // rank order: same bits on every rank
            double v;
            if (r == rank) {
                v = __ldcg(buf + j);
            } else {
                const LLCell* c = recv[rank] + slot + (long long)r * len + j;
                unsigned long long f;
                do {
                    ld_ll(c, v, f);
                    if (f != seq && (unsigned long long)(clock64() - t0) > kSpinLimit) { ok = false; v = 0.0; break; }
                } while (f != seq);
            }
            s += v;
        }
        buf[j] = s;
    }
    return __syncthreads_and(ok ? 1 : 0) != 0;
}

template <int MODE>
__device__ __forceinline__ void run_pass(const PersistArgs& S, const double* vec, double* out, unsigned char* smem_raw,
                                         uint64_t* full, unsigned& phase, long long t_beg, long long t_end, double* part,
                                         double beta) {
    BlockedArgs A = S.B;
    A.vec = vec; A.out = out; A.zero_next = nullptr; A.n_zero = 0;
    asm volatile("fence.proxy.async;" ::: "memory");         // hcoef: written through the generic proxy, read by TMA (async proxy)
    if (MODE == 0) { A.arb = out + S.n; A.hcoef = S.hcoef; }
    else A.slab[0] = S.hcoef;
    if (MODE == 1) { A.vec2 = S.z; A.beta = beta; }           // direction = z + beta p, formed by the gather
    double acc = 0.0, acc2 = 0.0;
    blocked_pass<kTileP, PT, kTileStages, MODE, false, MODE == 0, true, false>(A, smem_raw, full, phase, t_beg, t_end, acc, acc2);
    if (MODE == 0) {
        cta_accumulate<PT>(acc, part, A.arb);
        asm volatile("fence.proxy.async;" ::: "memory");     // hcoef written here (generic proxy) is read by TMA in later passes
    }
    if (MODE == 1) {                                         // p'Hp and p'diag(H)p ride behind the vector: y[n], y[n+1]
        cta_accumulate<PT>(acc, part, out + S.n);
        __syncthreads();
        cta_accumulate<PT>(acc2, part, out + S.n + 1);
    }
}

// CTA 0's solver state between passes.  Lives in shared memory, not registers: every CTA runs the same kernel and the
// pass loop must not carry (and spill) a dozen doubles that only CTA 0 ever uses.
struct BossState {
    int state, cur, iters, evals, hvps, status, cg_k, ls, aborted;
    double err, g0, rz, r0n, eta, alpha, lin1, beta;
    Kkt kc;
    unsigned long long seq_acc, seq_vec;
    long long prof[16];
};

// ---- the vector loops of CTA 0, one small function each (own register allocation: the loads of a chunk stay in flight)
__device__ __noinline__ double boss_cg_init(const PersistArgs& S, int cur, double* sh) {
    const int n = S.n;
    double s[1] = {0.0}, dummy[1] = {0.0};
    const double *frc = S.fr[cur], *pgc = S.pg[cur], *Sdiag = S.diag;
    double *Sminv = S.minv, *Sx = S.x, *Sr = S.r, *Sz = S.z, *Sp = S.p, *Sy = S.y;
    BOSS_CHUNKS(n) {                                 // PCG start: x = 0, r = -pg, z = M^-1 r, p = z
        double f[U], d[U], g[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = min(BOSS_IDX(u), n - 1);
            f[u] = __ldcg(frc + j); d[u] = __ldcg(Sdiag + j); g[u] = __ldcg(pgc + j);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = BOSS_IDX(u);
            if (j < n) {
                const double mi = f[u] / fmax(d[u], 1e-300);
                const double r = -g[u];
                const double z = mi * r;
                Sminv[j] = mi; Sx[j] = 0.0; Sr[j] = r; Sz[j] = z; Sp[j] = z;
                s[0] += r * z;
            }
        }
    }
    for (int j = threadIdx.x; j < n + 2; j += PT) Sy[j] = 0.0;
    cta_reduce<1, 0>(s, dummy, sh);
    return s[0];
}

// one PCG iteration after y = Hs p arrived (p = z + beta p formed on the fly); returns the new r'z
__device__ __noinline__ double boss_cg_update(const PersistArgs& S, double beta, double al, bool flat, bool first, double* sh,
                                              long long* prof) {
    const int n = S.n;
    const long long c0 = clock64();
    double s[1] = {0.0}, dummy[1] = {0.0};
    double *Sy = S.y, *Sp = S.p, *Sz = S.z, *Sx = S.x, *Sr = S.r;
    const double* Sminv = S.minv;
    BOSS_CHUNKS(n) {                                 // (minv = 0 off the free set, so z = p = 0 there whatever r is: no mask needed)
        double y[U], pp[U], zz[U], xx[U], rr[U], mm[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = min(BOSS_IDX(u), n - 1);
            y[u] = __ldcg(Sy + j); pp[u] = __ldcg(Sp + j); zz[u] = __ldcg(Sz + j);
            xx[u] = __ldcg(Sx + j); rr[u] = __ldcg(Sr + j); mm[u] = __ldcg(Sminv + j);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = BOSS_IDX(u);
            if (j < n) {
                const double pj = fma(beta, pp[u], zz[u]);      // the direction the pass used
                Sp[j] = pj;
                if (flat) {
                    if (first) Sx[j] = pj;
                } else {
                    const double r = rr[u] - al * y[u];
                    const double z = mm[u] * r;
                    Sx[j] = fma(al, pj, xx[u]); Sr[j] = r; Sz[j] = z;
                    s[0] += r * z;
                }
            }
        }
    }
    const long long c1 = clock64();
    __syncthreads();                                 // every thread has read its y entries
    const long long c2 = clock64();
    for (int j = threadIdx.x; j < n + 2; j += PT) Sy[j] = 0.0;
    const long long c3 = clock64();
    if (flat) return 0.0;
    cta_reduce<1, 0>(s, dummy, sh);
    if (threadIdx.x == 0) { prof[12] += c1 - c0; prof[13] += c2 - c1; prof[14] += c3 - c2; prof[15] += clock64() - c3; }
    return s[0];
}

// dt <- x if it is a descent direction in value units (pg . x < 0), else scaled steepest descent
__device__ __noinline__ void boss_direction(const PersistArgs& S, int cur, double* sh) {
    const int n = S.n;
    double s[1] = {0.0}, m[1] = {0.0};
    const double *pgc = S.pg[cur], *Sx = S.x;
    double* Sdt = S.dt;
    __syncthreads();                                 // x of this iteration is complete
    BOSS_CHUNKS(n) {
        double g[U], xx[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { const int j = min(BOSS_IDX(u), n - 1); g[u] = __ldcg(pgc + j); xx[u] = __ldcg(Sx + j); }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (BOSS_IDX(u) < n) { s[0] += g[u] * xx[u]; m[0] = fmax(m[0], fabs(g[u])); }
    }
    cta_reduce<1, 1>(s, m, sh);
    const bool ok = isfinite(s[0]) && s[0] < 0.0;
    const double imx = 1.0 / fmax(m[0], 1e-300);
    BOSS_CHUNKS(n) {
        double g[U], xx[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { const int j = min(BOSS_IDX(u), n - 1); g[u] = __ldcg(pgc + j); xx[u] = __ldcg(Sx + j); }
#pragma unroll
        for (int u = 0; u < U; ++u) { const int j = BOSS_IDX(u); if (j < n) Sdt[j] = ok ? xx[u] : -g[u] * imx; }
    }
    __syncthreads();
}

// trial point nu[tr] = max(nu[cur] * exp(alpha dt), lb) (pinned tokens stay at c); clears the trial accumulator
__device__ __noinline__ void boss_take_step(const PersistArgs& S, int cur, double alpha) {
    const int n = S.n, tr = cur ^ 1;
    const double *nuc = S.nu[cur], *Sdt = S.dt, *Slb = S.lb, *Sc = S.c;
    const unsigned char* Sfx = S.fixed;
    double *nut = S.nu[tr], *acct = S.acc[tr];
    BOSS_CHUNKS(n) {
        double d[U], v[U], l[U], c[U];
        unsigned char fx[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = min(BOSS_IDX(u), n - 1);
            d[u] = __ldcg(Sdt + j); v[u] = __ldcg(nuc + j); l[u] = __ldcg(Slb + j); c[u] = Sc[j]; fx[u] = Sfx[j];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = BOSS_IDX(u);
            if (j < n) {
                const double e = fmin(fmax(alpha * d[u], -20.0), 20.0);
                nut[j] = fx[u] ? c[u] : fmax(v[u] * exp(e), l[u]);
            }
        }
    }
    for (int j = threadIdx.x; j <= n; j += PT) acct[j] = 0.0;
}

// What CTA 0 does after the grid finished a pass (`op` on set `set`): all-reduce over the peers, vector algebra,
// decision.  Returns the next command.  Every thread of the CTA computes the same scalars (broadcast reductions).
//
// PCG with ONE vector loop per iteration: the pass delivers y = Hs p together with p'Hp and p'diag(H)p (y[n], y[n+1]), so
// alpha is known on arrival; the loop forms p (= z + beta p, the same expression the pass gathered), updates x, r, z and
// accumulates r'z; beta then travels to the next pass as a scalar -- p is never materialised before it is needed.
__device__ __noinline__ unsigned boss_step(const PersistArgs& S, BossState* bs, unsigned op, unsigned set, double* sh) {
    const int tid = threadIdx.x, n = S.n;
    int state = bs->state, cur = bs->cur, iters = bs->iters, evals = bs->evals, hvps = bs->hvps, status = bs->status;
    int cg_k = bs->cg_k, ls = bs->ls;
    bool aborted = bs->aborted != 0;
    double err = bs->err, g0 = bs->g0, rz = bs->rz, r0n = bs->r0n, eta = bs->eta, alpha = bs->alpha, lin1 = bs->lin1;
    double beta = bs->beta;
    Kkt kc = bs->kc;
    unsigned long long seq_acc = bs->seq_acc, seq_vec = bs->seq_vec;
    const long long e0 = clock64();
    __syncthreads();                                 // everybody has read the state before thread 0 rewrites it below
    if (tid == 0) bs->prof[10] += clock64() - e0;
    bool finish = aborted;
    unsigned next = OP_DONE;
    if (!finish && op == OP_EVAL) {
        double* acc = S.acc[set];
        if (S.world > 1) { ++seq_acc; if (!ll_allreduce_cta(acc, n + 1, S.recv_acc, S.rank, S.world, seq_acc)) finish = aborted = true; }
        ++evals;
        if (!finish && state == ST_EVAL0) {
            kc = kkt_eval(S, S.nu[cur], acc, 1e-2, S.grad[cur], S.fr[cur], S.pg[cur], nullptr, nullptr, sh);
            err = kc.err; g0 = kc.g;
            goto newton_start;
        } else if (!finish) {                       // ST_TRIAL: Armijo test along nu * exp(alpha dt)
            const int tr = cur ^ 1;
            const double thr = fmin(1e-2, fmax(isfinite(err) ? err : 1e-2, 1e-14));
            const Kkt kt = kkt_eval(S, S.nu[tr], acc, thr, S.grad[tr], S.fr[tr], S.pg[tr], S.nu[cur], S.grad[cur], sh);
            const double gt = kt.g, lin = kt.lin;
            if (ls == 0) lin1 = lin;
            bool accept = gt <= g0 + 1e-4 * lin, stalled = false;
            if (!accept && (fabs(gt - g0) <= 1e-13 * fabs(g0) || fabs(lin1) <= 1e-9 * fabs(g0))) {
                // the (full) step is below what g resolves in fp64: judge it by the KKT residual (as solver.py)
                if (kt.err < 0.99 * err) accept = true;
                else if (alpha < 1e-3) stalled = true;
            }
            if (accept) {
                cur = tr; kc = kt; err = kt.err; g0 = kt.g;
                goto newton_start;
            }
            if (stalled || ++ls >= 50) { status = 2; finish = true; }
            else { alpha *= 0.5; goto take_step; }
        }
    } else if (!finish && op == OP_DIAG) {
        if (S.world > 1) { ++seq_vec; if (!ll_allreduce_cta(S.diag, n, S.recv_vec, S.rank, S.world, seq_vec)) finish = aborted = true; }
        if (!finish) {
            rz = boss_cg_init(S, cur, sh);
            r0n = sqrt(fmax(rz, 0.0));
            eta = fmin(0.1, sqrt(err));
            cg_k = 0; beta = 0.0;                       // first direction: p = z + 0 * p
            if (rz <= 0.0) goto direction;
            next = OP_HVP; state = ST_HVP;
        }
    } else if (!finish) {                           // OP_HVP: one PCG iteration with y = Hs p, p = z + beta p
        if (S.world > 1) { ++seq_vec; if (!ll_allreduce_cta(S.y, n + 2, S.recv_vec, S.rank, S.world, seq_vec)) finish = aborted = true; }
        ++hvps;
        if (!finish) {
            const long long q0 = clock64();
            const double pHp = __ldcg(S.y + n), pdp = __ldcg(S.y + n + 1);
            const bool flat = pHp <= 1e-14 * pdp;       // homogeneity direction: g is linear along nu
            const double al = flat ? 0.0 : rz / pHp;
            __syncthreads();                            // y[n], y[n+1] are read before the update clears y
            const long long q1 = clock64();
            const double rzn = boss_cg_update(S, beta, al, flat, cg_k == 0, sh, bs->prof);
            if (tid == 0) { bs->prof[8] += q1 - q0; bs->prof[9] += clock64() - q1; }
            bool stop = flat;
            if (!flat) {
                stop = (rzn <= 0.0) || (sqrt(fmax(rzn, 0.0)) <= eta * r0n);
                beta = rzn / rz;
                rz = rzn;
            }
            ++cg_k;
            if (!stop && cg_k < S.cg_max) next = OP_HVP;
            else goto direction;
        }
    }
    goto decided;

newton_start:                                   // same counting as cfmm_solver.cu: the final check is an iteration too
    if (iters >= S.max_iter) { status = 1; finish = true; goto decided; }
    ++iters;
    if (err <= S.tol) { status = 0; finish = true; goto decided; }
    { double* Sdiag = S.diag; for (int j = tid; j < n; j += PT) Sdiag[j] = 0.0; }
    next = OP_DIAG; state = ST_DIAG;
    goto decided;

direction:
    boss_direction(S, cur, sh);
    alpha = 1.0; ls = 0; lin1 = 0.0;
take_step:
    boss_take_step(S, cur, alpha);
    next = OP_EVAL | ((unsigned)(cur ^ 1) << 4); state = ST_TRIAL;

decided:
    if (finish) {
        // results at the accepted point: nu[cur], acc[cur] (+ the KKT data of that point in kc)
        if (status == 1 && err <= S.tol) status = 0;
        const double *nuc = S.nu[cur], *accc = S.acc[cur];
        double *po = S.psi_out, *no = S.nu_out;
        for (int j = tid; j < n; j += PT) {
            const double v = __ldcg(nuc + j);
            po[j] = __ldcg(accc + j);
            if (cur != 0) no[j] = v;
        }
        if (tid == 0) {
            DevResult R;
            R.dual_value = kc.g; R.primal_value = kc.primal;
            R.gap = (kc.g - kc.primal) / fmax(fabs(kc.g), 1e-300);
            R.primal_infeas = kc.infeas; R.err = err;
            R.iters = iters; R.evals = evals; R.hvps = hvps; R.status = aborted ? 3 : status;
            R.seq_acc = seq_acc; R.seq_vec = seq_vec;
            for (int k = 0; k < 16; ++k) R.prof[k] = bs->prof[k];
            *S.res = R;
        }
        next = OP_DONE;
    }
    if (tid == 0) {
        bs->state = state; bs->cur = cur; bs->iters = iters; bs->evals = evals; bs->hvps = hvps; bs->status = status;
        bs->cg_k = cg_k; bs->ls = ls; bs->aborted = aborted ? 1 : 0;
        bs->err = err; bs->g0 = g0; bs->rz = rz; bs->r0n = r0n; bs->eta = eta; bs->alpha = alpha; bs->lin1 = lin1;
        bs->beta = beta;
        bs->kc = kc; bs->seq_acc = seq_acc; bs->seq_vec = seq_vec;
        *reinterpret_cast<double*>(S.ctl + 66) = beta;        // travels with the HVP command
    }
    __syncthreads();
    return next;
}

__global__ void __launch_bounds__(PT, kCtasPerSm)
k_solve_persist(const __grid_constant__ PersistArgs S) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full[kTileStages];
    __shared__ double part[PT / 32];
    __shared__ double sh[8 * 16];
    __shared__ unsigned s_cmd;
    __shared__ BossState bs;
    const int tid = threadIdx.x;
    const int n = S.n;
    if (tid == 0) {
        for (int s = 0; s < kTileStages; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    const long long t_beg = (S.B.n_tiles * (long long)blockIdx.x) / gridDim.x;
    const long long t_end = (S.B.n_tiles * (long long)(blockIdx.x + 1)) / gridDim.x;
    unsigned phase = 0;
    unsigned epoch = 0;                              // passes completed so far
    const bool boss = blockIdx.x == 0;
    // ---- prologue (CTA 0): solver state, bounds, start point, clear the first accumulator; command = evaluate at nu[0]
    if (boss) {
        if (tid == 0) {
            bs.state = ST_EVAL0; bs.cur = 0; bs.iters = 0; bs.evals = 0; bs.hvps = 0; bs.status = 1; bs.cg_k = 0; bs.ls = 0;
            bs.aborted = 0;
            bs.err = INFINITY; bs.g0 = 0.0; bs.rz = 0.0; bs.r0n = 0.0; bs.eta = 0.1; bs.alpha = 1.0; bs.lin1 = 0.0; bs.beta = 0.0;
            bs.kc.err = INFINITY; bs.kc.g = 0.0; bs.kc.primal = 0.0; bs.kc.infeas = 0.0; bs.kc.lin = 0.0;
            bs.seq_acc = S.seq_acc; bs.seq_vec = S.seq_vec;
            for (int k = 0; k < 16; ++k) bs.prof[k] = 0;
        }
        for (int j = tid; j < n; j += PT) {
            const double l = S.eq[j] ? S.nu_floor : fmax(S.c[j], S.nu_floor);
            S.lb[j] = l;
            S.nu[0][j] = S.fixed[j] ? S.c[j] : fmax(S.nu[0][j], l);
        }
        for (int j = tid; j <= n; j += PT) S.acc[0][j] = 0.0;
        __threadfence();
        __syncthreads();
        if (tid == 0) { S.ctl[65] = OP_EVAL; st_release_gpu(S.ctl + 64, 1u); }
    }
    __syncthreads();
    unsigned want = 1u;                              // epoch flag value that carries this round's command
    for (;;) {
        // ---- wait for CTA 0's command of this round
        if (tid == 0) {
            const long long t0 = clock64();
            unsigned c = OP_DONE;
            bool got = true;
            while (ld_acquire_gpu(S.ctl + 64) < want) {
                // back off between polls: ~300 CTAs spinning on one L2 line slow every other access of CTA 0 down
                __nanosleep(kPollSleepNs);
                if ((unsigned long long)(clock64() - t0) > kSpinLimit || ld_acquire_gpu(S.ctl + 32) != 0u) { got = false; break; }
            }
            if (got) c = ld_acquire_gpu(S.ctl + 65);
            s_cmd = c;
        }
        __syncthreads();
        const unsigned cmd = s_cmd;
        const unsigned op = cmd & 15u, set = (cmd >> 4) & 1u;
        if (op == OP_DONE) break;
        // ---- the pass, on this CTA's tiles
        const long long tp0 = clock64();
        if (op == OP_EVAL) run_pass<0>(S, S.nu[set], S.acc[set], smem_raw, full, phase, t_beg, t_end, part, 0.0);
        else if (op == OP_HVP) run_pass<1>(S, S.p, S.y, smem_raw, full, phase, t_beg, t_end, part,
                                           __ldcg(reinterpret_cast<const double*>(S.ctl + 66)));
        else run_pass<2>(S, nullptr, S.diag, smem_raw, full, phase, t_beg, t_end, part, 0.0);
        ++epoch;
        // ---- arrive: this CTA's red.adds are ordered before the counter bump
        __threadfence();
        __syncthreads();
        if (tid == 0) atomicAdd(S.ctl, 1u);
        ++want;
        if (!boss) continue;
        // ---- CTA 0: wait for the whole grid, then vector algebra + decision
        const long long tp1 = clock64();
        if (tid == 0) {
            const unsigned target = epoch * gridDim.x;
            while (ld_acquire_gpu(S.ctl) < target) {
                if ((unsigned long long)(clock64() - tp1) > kSpinLimit) { S.ctl[32] = 1u; bs.aborted = 1; break; }
            }
        }
        __syncthreads();
        const long long tp2 = clock64();
        const unsigned next = boss_step(S, &bs, op, set, sh);
        const long long tp3 = clock64();
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            bs.prof[op - 1] += tp1 - tp0; bs.prof[3] += tp2 - tp1; bs.prof[3 + op] += clock64() - tp2;
            bs.prof[11] += clock64() - tp3;
            if ((next & 15u) == OP_DONE) for (int k = 0; k < 16; ++k) S.res->prof[k] = bs.prof[k];
            S.ctl[65] = next; st_release_gpu(S.ctl + 64, want);
        }
    }
}

// =====================================================================================================================
// Distributed variant: NO boss.  The n tokens are cut into `nsl` slices (a function of n only, so the same on every rank
// of a sharded solve); slice s belongs to CTA s mod G and is worked by one warp, one token per lane -- every load of a
// slice phase is issued at once (one L2 round trip), nothing loops over n.  After a pass:
//     grid barrier A  ->  slice phase: [NVLink LL exchange of the slice] + element-wise update + slice partial sums
//     grid barrier B  ->  decide phase: every CTA adds the slice partials in slice order (same bits everywhere, on every
//                         rank) and runs the same scalar state machine, so all CTAs (and all ranks) agree on the next
//                         pass without a command broadcast.
// A sharded solve all-reduces inside the slice phase: each lane pushes ITS token to the peers and sums what they pushed
// (rank order) -- the exchange is spread over all CTAs instead of serialised in one.
// =====================================================================================================================
// Loads of data another CTA wrote before the last grid barrier: PLAIN loads.  The barrier is fence + bar.sync + atomic on
// the writer's side and ld.acquire.gpu + bar.sync on the reader's, which orders weak accesses across it (and the acquire
// refreshes this SM's L1); ld.global.cg would also be correct but B200 issues those ~100 ns apart per thread (measured:
// 72 of them = 8 us), while plain loads pipeline.
__device__ __forceinline__ double ldw(const double* p) { return *p; }

constexpr int kSliceMax = 256;                   // slices (>= 1): min(256, ceil(n / 16))
constexpr int kQ = 10;                           // partial quantities per phase

enum { PH_KKT = 1, PH_DIAG = 2, PH_HVP = 3, PH_STEP = 4, PH_DONE = 5 };

struct DistArgs {
    PersistArgs P;
    double* accr[2];                             // all-reduced psi of nu[0] / nu[1] (what psi_out returns)
    double* y2[2];                               // Hs p ping-pong, [n + 2] each (+ p'Hp, p'diag(H)p)
    double* partial;                             // [2 parities][kQ][kSliceMax]
    int nsl;
};

struct DState {                                  // replicated in every CTA's shared memory; thread 0 updates it
    int phase, set, cur, iters, evals, hvps, status, cg_k, ls, yb, first_step, dir_ok, aborted, parity;
    double err, g0, rz, r0n, eta, alpha, lin1, beta, al, imx, thr;
    int flat;
    Kkt kc;
    unsigned long long seq_acc, seq_vec;
    unsigned bar;                                // grid barriers passed so far
    long long prof[16];
};

__device__ __forceinline__ void grid_barrier(unsigned* ctl, DState* ds) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned target = (++ds->bar) * gridDim.x;
        atomicAdd(ctl, 1u);
        const long long t0 = clock64();
        while (ld_acquire_gpu(ctl) < target) {
            if (ld_acquire_gpu(ctl + 32) != 0u) { ds->aborted = 1; break; }
            if ((unsigned long long)(clock64() - t0) > kSpinLimit) { ctl[32] = 1u; ds->aborted = 1; break; }
        }
    }
    __syncthreads();
}

// value of element j summed over the ranks (rank order), via the LL receive areas; `mine` = this rank's partial
__device__ __forceinline__ double ll_exchange(double mine, LLCell* const* recv, long long slot, long long stride, int j,
                                              int rank, int world, unsigned long long seq, bool push, unsigned* ctl, int* aborted) {
    if (push)
        for (int r = 0; r < world; ++r)
            if (r != rank) st_ll(recv[r] + slot + (long long)rank * stride + j, mine, seq);
    double s = 0.0;
    const long long t0 = clock64();
    for (int r = 0; r < world; ++r) {
        double v = mine;
        if (r != rank) {
            const LLCell* c = recv[rank] + slot + (long long)r * stride + j;
            unsigned long long f;
            do {
                ld_ll(c, v, f);
                if (f != seq && (unsigned long long)(clock64() - t0) > kSpinLimit) { ctl[32] = 1u; *aborted = 1; v = 0.0; break; }
            } while (f != seq);
        }
        s += v;
    }
    return s;
}

__global__ void __launch_bounds__(PT, kCtasPerSm)
k_solve_dist(const __grid_constant__ DistArgs D) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full[kTileStages];
    __shared__ double part[PT / 32];
    __shared__ double tot[kQ];
    __shared__ DState ds;
    const PersistArgs& S = D.P;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = S.n, nsl = D.nsl, G = gridDim.x;
    if (tid == 0) {
        for (int s = 0; s < kTileStages; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
        ds.phase = PH_KKT; ds.set = 0; ds.cur = 0; ds.iters = 0; ds.evals = 0; ds.hvps = 0; ds.status = 1; ds.cg_k = 0; ds.ls = 0;
        ds.yb = 0; ds.first_step = 1; ds.dir_ok = 0; ds.aborted = 0; ds.parity = 0; ds.flat = 0;
        ds.err = INFINITY; ds.g0 = 0.0; ds.rz = 0.0; ds.r0n = 0.0; ds.eta = 0.1; ds.alpha = 1.0; ds.lin1 = 0.0; ds.beta = 0.0;
        ds.al = 0.0; ds.imx = 0.0; ds.thr = 1e-2;
        ds.kc.err = INFINITY; ds.kc.g = 0.0; ds.kc.primal = 0.0; ds.kc.infeas = 0.0; ds.kc.lin = 0.0;
        ds.seq_acc = S.seq_acc; ds.seq_vec = S.seq_vec; ds.bar = 0;
        for (int k = 0; k < 16; ++k) ds.prof[k] = 0;
    }
    __syncthreads();
    const long long t_beg = (S.B.n_tiles * (long long)blockIdx.x) / G;
    const long long t_end = (S.B.n_tiles * (long long)(blockIdx.x + 1)) / G;
    unsigned phase_bits = 0;
    // ---- prologue: bounds, start point, cleared accumulator -- every slice by its owner
    for (int s0 = blockIdx.x + warp * G; s0 < nsl; s0 += (PT / 32) * G) {
        const int lo = (int)((long long)n * s0 / nsl), hi = (int)((long long)n * (s0 + 1) / nsl);
        for (int j = lo + lane; j < hi; j += 32) {
            const double l = S.eq[j] ? S.nu_floor : fmax(S.c[j], S.nu_floor);
            S.lb[j] = l;
            S.nu[0][j] = S.fixed[j] ? S.c[j] : fmax(S.nu[0][j], l);
            S.acc[0][j] = 0.0;
        }
        if (s0 == nsl - 1 && lane == 0) S.acc[0][n] = 0.0;
    }
    grid_barrier(S.ctl, &ds);
    for (;;) {
        const int phase = ds.phase, set = ds.set, cur = ds.cur;
        if (phase == PH_DONE || ds.aborted) break;
        const long long tp0 = clock64();
        // ---- the pass of this phase, on this CTA's tiles, then barrier A
        if (phase == PH_KKT) run_pass<0>(S, S.nu[set], S.acc[set], smem_raw, full, phase_bits, t_beg, t_end, part, 0.0);
        else if (phase == PH_HVP) run_pass<1>(S, S.p, D.y2[ds.yb], smem_raw, full, phase_bits, t_beg, t_end, part, ds.beta);
        else if (phase == PH_DIAG) run_pass<2>(S, nullptr, S.diag, smem_raw, full, phase_bits, t_beg, t_end, part, 0.0);
        const long long tp1 = clock64();
        if (phase != PH_STEP) grid_barrier(S.ctl, &ds);
        const long long tp2 = clock64();
        // ---- slice phase
        const int par = ds.parity;
        double* P0 = D.partial + (size_t)par * kQ * kSliceMax;
        const bool multi = S.world > 1;
        unsigned long long seq = 0;
        if (phase == PH_KKT) seq = ds.seq_acc + 1; else if (phase == PH_DIAG || phase == PH_HVP) seq = ds.seq_vec + 1;
        if (phase == PH_HVP) {                       // p'Hp and p'diag(H)p first: alpha feeds the element-wise update
            if (tid == 0) {
                const double* yb = D.y2[ds.yb];
                double pHp = ldw(yb + n), pdp = ldw(yb + n + 1);
                if (multi) {                         // one thread of the grid pushes them, every CTA sums what the peers pushed
                    const long long slot = (long long)(seq % 3) * S.world * (n + 2);
                    const bool owner = blockIdx.x == (nsl - 1) % G;
                    int ab = 0;
                    pHp = ll_exchange(pHp, S.recv_vec, slot, n + 2, n, S.rank, S.world, seq, owner, S.ctl, &ab);
                    pdp = ll_exchange(pdp, S.recv_vec, slot, n + 2, n + 1, S.rank, S.world, seq, owner, S.ctl, &ab);
                }
                ds.flat = pHp <= 1e-14 * pdp;        // homogeneity direction: g is linear along nu
                ds.al = ds.flat ? 0.0 : ds.rz / pHp;
            }
            __syncthreads();
        }
        const double al = ds.al;
        const int flat = ds.flat;
        for (int s0 = blockIdx.x + warp * G; s0 < nsl; s0 += (PT / 32) * G) {
            const int lo = (int)((long long)n * s0 / nsl), hi = (int)((long long)n * (s0 + 1) / nsl);
            double q[kQ];
#pragma unroll
            for (int k = 0; k < kQ; ++k) q[k] = 0.0;
            for (int j = lo + lane; j < hi; j += 32) {
                if (phase == PH_KKT) {
                    double pj = ldw(S.acc[set] + j);
                    const double nj = ldw(S.nu[set] + j), aj = S.a[j], cj = S.c[j], lbj = ldw(S.lb + j);
                    const unsigned char ej = S.eq[j], fj = S.fixed[j];
                    const bool wl = set != cur;      // trial point: also grad_cur . (nu_trial - nu_cur)
                    const double np = wl ? ldw(S.nu[cur] + j) : 0.0, gp = wl ? ldw(S.grad[cur] + j) : 0.0;
                    if (multi) {
                        int ab = 0;
                        pj = ll_exchange(pj, S.recv_acc, (long long)(seq % 3) * S.world * (n + 1), n + 1, j, S.rank, S.world, seq, true, S.ctl, &ab);
                    }
                    const double g = aj + pj;
                    const bool near = (nj <= lbj * (1.0 + ds.thr)) && !ej;
                    const bool act = fj || (near && g > 0.0);
                    const double f = act ? 0.0 : 1.0, v = nj * g * f;
                    S.grad[set][j] = g; S.fr[set][j] = f; S.pg[set][j] = v; D.accr[set][j] = pj;
                    S.diag[j] = 0.0;                 // the diagonal pass that may follow accumulates into it
                    q[0] += fabs(v); q[1] += (nj - cj) * aj; q[2] += nj * fabs(g); q[3] += cj * pj;
                    q[4] += nj * (fj ? 0.0 : (ej ? fabs(g) : fmax(-g, 0.0)));
                    q[5] += gp * (nj - np);
                    q[6] = fmax(q[6], fabs(g) * f);
                    q[7] = fmax(q[7], fmax(fabs(aj), fj ? 0.0 : fabs(pj)));
                } else if (phase == PH_DIAG) {
                    double d = ldw(S.diag + j);
                    const double f = ldw(S.fr[cur] + j), g = ldw(S.pg[cur] + j);
                    if (multi) {
                        int ab = 0;
                        d = ll_exchange(d, S.recv_vec, (long long)(seq % 3) * S.world * (n + 2), n + 2, j, S.rank, S.world, seq, true, S.ctl, &ab);
                    }
                    const double mi = f / fmax(d, 1e-300), r = -g, z = mi * r;
                    S.minv[j] = mi; S.x[j] = 0.0; S.r[j] = r; S.z[j] = z; S.p[j] = z;
                    D.y2[0][j] = 0.0;
                    q[0] += r * z;
                    q[6] = fmax(q[6], fabs(g));
                } else if (phase == PH_HVP) {
                    double yv = ldw(D.y2[ds.yb] + j);
                    const double pp = ldw(S.p + j), zz = ldw(S.z + j), xx = ldw(S.x + j), rr = ldw(S.r + j),
                                 mm = ldw(S.minv + j), g = ldw(S.pg[cur] + j);
                    if (multi) {
                        int ab = 0;
                        yv = ll_exchange(yv, S.recv_vec, (long long)(seq % 3) * S.world * (n + 2), n + 2, j, S.rank, S.world, seq, true, S.ctl, &ab);
                    }
                    const double pj = fma(ds.beta, pp, zz);             // the direction the pass used
                    double xn = xx;
                    S.p[j] = pj;
                    if (flat) {
                        if (ds.cg_k == 0) { xn = pj; S.x[j] = xn; }
                    } else {
                        const double r = rr - al * yv, z = mm * r;
                        xn = fma(al, pj, xx);
                        S.x[j] = xn; S.r[j] = r; S.z[j] = z;
                        q[0] += r * z;
                    }
                    q[1] += g * xn;                  // pg . x: is x a descent direction (needed when PCG stops here)
                    D.y2[ds.yb ^ 1][j] = 0.0;
                } else {                             // PH_STEP: direction (first step of a search) + trial point
                    const double g = ldw(S.pg[cur] + j), xx = ldw(S.x + j), v = ldw(S.nu[cur] + j), l = ldw(S.lb + j);
                    double d;
                    if (ds.first_step) { d = ds.dir_ok ? xx : -g * ds.imx; S.dt[j] = d; }
                    else d = ldw(S.dt + j);
                    const double e = fmin(fmax(ds.alpha * d, -20.0), 20.0);
                    S.nu[cur ^ 1][j] = S.fixed[j] ? S.c[j] : fmax(v * exp(e), l);
                    S.acc[cur ^ 1][j] = 0.0;
                }
            }
            if (s0 == nsl - 1 && lane == 0) {        // the extras behind the vectors
                if (phase == PH_KKT) {
                    double arb = ldw(S.acc[set] + n);
                    if (multi) {
                        int ab = 0;
                        arb = ll_exchange(arb, S.recv_acc, (long long)(seq % 3) * S.world * (n + 1), n + 1, n, S.rank, S.world, seq, true, S.ctl, &ab);
                    }
                    q[8] = arb;
                } else if (phase == PH_DIAG) { D.y2[0][n] = 0.0; D.y2[0][n + 1] = 0.0; }
                else if (phase == PH_HVP) { D.y2[ds.yb ^ 1][n] = 0.0; D.y2[ds.yb ^ 1][n + 1] = 0.0; }
                else S.acc[cur ^ 1][n] = 0.0;
            }
            if (phase != PH_STEP) {
#pragma unroll
                for (int k = 0; k < 6; ++k) q[k] = warp_sum(q[k]);
                q[6] = warp_max(q[6]); q[7] = warp_max(q[7]);
                if (lane == 0) {
#pragma unroll
                    for (int k = 0; k < 9; ++k) P0[k * kSliceMax + s0] = q[k];
                }
            }
        }
        const long long tp3 = clock64();
        grid_barrier(S.ctl, &ds);
        const long long tp4 = clock64();
        // ---- decide phase: slice partials -> totals (slice order: same bits in every CTA and on every rank)
        if (warp < 9 && phase != PH_STEP) {          // warp k totals quantity k: all loads of the phase in flight at once
            const int k = warp;
            const bool is_max = k == 6 || k == 7;
            double t = 0.0;
            for (int s = lane; s < nsl; s += 32) {
                const double v = ldw(P0 + k * kSliceMax + s);
                t = is_max ? fmax(t, v) : t + v;
            }
            t = is_max ? warp_max(t) : warp_sum(t);
            if (lane == 0) tot[k] = t;
        }
        __syncthreads();
        if (tid == 0) {
            ds.parity ^= 1;
            int next = PH_DONE;
            bool newton = false;
            if (ds.aborted) {
                next = PH_DONE;
            } else if (phase == PH_KKT) {
                ++ds.seq_acc; ++ds.evals;
                Kkt k;
                k.g = tot[1] + tot[8];
                k.err = fmax(tot[0] / fmax(fmax(fabs(k.g), 1e-3 * tot[2]), 1e-300), tot[6] / fmax(tot[7], 1e-300));
                k.primal = tot[3]; k.infeas = tot[4] / fmax(fabs(k.g), 1e-300); k.lin = tot[5];
                if (set == cur) {                    // the start point
                    ds.kc = k; ds.err = k.err; ds.g0 = k.g; newton = true;
                } else {                             // Armijo test along nu * exp(alpha dt)
                    if (ds.ls == 0) ds.lin1 = k.lin;
                    bool accept = k.g <= ds.g0 + 1e-4 * k.lin, stalled = false;
                    if (!accept && (fabs(k.g - ds.g0) <= 1e-13 * fabs(ds.g0) || fabs(ds.lin1) <= 1e-9 * fabs(ds.g0))) {
                        if (k.err < 0.99 * ds.err) accept = true;        // below what g resolves: judged by the KKT residual
                        else if (ds.alpha < 1e-3) stalled = true;
                    }
                    if (accept) { ds.cur = set; ds.kc = k; ds.err = k.err; ds.g0 = k.g; newton = true; }
                    else if (stalled || ++ds.ls >= 50) { ds.status = 2; next = PH_DONE; }
                    else { ds.alpha *= 0.5; ds.first_step = 0; next = PH_STEP; }
                }
                if (newton) {                        // same counting as cfmm_solver.cu: the final check is an iteration too
                    if (ds.iters >= S.max_iter) { ds.status = 1; next = PH_DONE; }
                    else {
                        ++ds.iters;
                        if (ds.err <= S.tol) { ds.status = 0; next = PH_DONE; }
                        else next = PH_DIAG;
                    }
                }
            } else if (phase == PH_DIAG) {
                ++ds.seq_vec;
                ds.rz = tot[0]; ds.r0n = sqrt(fmax(ds.rz, 0.0)); ds.eta = fmin(0.1, sqrt(ds.err));
                ds.imx = 1.0 / fmax(tot[6], 1e-300);
                ds.cg_k = 0; ds.beta = 0.0; ds.yb = 0;
                if (ds.rz <= 0.0) { ds.dir_ok = 0; ds.first_step = 1; ds.alpha = 1.0; ds.ls = 0; ds.lin1 = 0.0; next = PH_STEP; }
                else next = PH_HVP;
            } else if (phase == PH_HVP) {
                ++ds.seq_vec; ++ds.hvps;
                bool stop = flat != 0;
                if (!flat) {
                    const double rzn = tot[0];
                    stop = (rzn <= 0.0) || (sqrt(fmax(rzn, 0.0)) <= ds.eta * ds.r0n);
                    ds.beta = rzn / ds.rz;
                    ds.rz = rzn;
                }
                ++ds.cg_k; ds.yb ^= 1;
                if (!stop && ds.cg_k < S.cg_max) next = PH_HVP;
                else {
                    const double sdir = tot[1];
                    ds.dir_ok = (isfinite(sdir) && sdir < 0.0) ? 1 : 0;
                    ds.first_step = 1; ds.alpha = 1.0; ds.ls = 0; ds.lin1 = 0.0;
                    next = PH_STEP;
                }
            } else {                                 // PH_STEP -> evaluate the trial point
                ds.set = cur ^ 1;
                ds.thr = fmin(1e-2, fmax(isfinite(ds.err) ? ds.err : 1e-2, 1e-14));
                next = PH_KKT;
            }
            if (next == PH_DONE && ds.status == 1 && ds.err <= S.tol) ds.status = 0;
            ds.phase = next;
            if (blockIdx.x == 0) {
                ds.prof[phase == PH_KKT ? 0 : phase == PH_HVP ? 1 : phase == PH_DIAG ? 2 : 7] += tp1 - tp0;
                ds.prof[3] += tp2 - tp1; ds.prof[4] += tp3 - tp2; ds.prof[5] += tp4 - tp3; ds.prof[6] += clock64() - tp4;
            }
        }
        __syncthreads();
    }
    // ---- results: every slice owner copies its part; CTA 0 writes the scalars
    {
        const int cur = ds.cur;
        for (int s0 = blockIdx.x + warp * G; s0 < nsl; s0 += (PT / 32) * G) {
            const int lo = (int)((long long)n * s0 / nsl), hi = (int)((long long)n * (s0 + 1) / nsl);
            for (int j = lo + lane; j < hi; j += 32) {
                S.psi_out[j] = ldw(D.accr[cur] + j);
                if (cur != 0) S.nu_out[j] = ldw(S.nu[cur] + j);
            }
        }
        if (blockIdx.x == 0 && tid == 0) {
            DevResult R;
            R.dual_value = ds.kc.g; R.primal_value = ds.kc.primal;
            R.gap = (ds.kc.g - ds.kc.primal) / fmax(fabs(ds.kc.g), 1e-300);
            R.primal_infeas = ds.kc.infeas; R.err = ds.err;
            R.iters = ds.iters; R.evals = ds.evals; R.hvps = ds.hvps; R.status = ds.aborted ? 3 : ds.status;
            R.seq_acc = ds.seq_acc; R.seq_vec = ds.seq_vec;
            for (int k = 0; k < 16; ++k) R.prof[k] = ds.prof[k];
            *S.res = R;
        }
    }
}

inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }
thread_local long long g_last_prof[16] = {0};
int g_persist_mode = 0;                         // 0: distributed vector algebra (default), 1: CTA 0 does it

}  // namespace

extern "C" {

int64_t cfmm_persist_solve_work_bytes(const cfmm_blocked_pairs* b, int32_t n_tokens) {
    if (!b || n_tokens <= 0) return CFMM_E_SIZE;
    const size_t n = (size_t)n_tokens;
    size_t bytes = align_up(8 * (size_t)b->n_tiles * (size_t)b->pools_per_tile);      // hcoef
    bytes += 2 * align_up(8 * (n + 1));       // [psi | arb] of the current / trial point
    bytes += 18 * align_up(8 * n) + align_up(8 * (n + 2));    // nu trial, diag, lb, grad x2, fr x2, pg x2, dt, x, r, z, p, minv (+3 spare); y (+ p'Hp, p'Dp)
    bytes += align_up(512) + align_up(sizeof(DevResult));
    bytes += 2 * align_up(8 * n) + 2 * align_up(8 * (n + 2)) + align_up(8 * (size_t)2 * kQ * kSliceMax);   // distributed variant
    return (int64_t)bytes;
}

int cfmm_persist_solve(const cfmm_blocked_pairs* b, int32_t n_tokens, const double* c, const double* a,
                       const uint8_t* eq, const uint8_t* pinned, double* nu, double* psi_out, void* work,
                       const cfmm_solve_params* prm, cfmm_solve_result* res, cfmm_peer_ctx* peer, void* stream) {
    if (!b || !c || !a || !eq || !pinned || !nu || !psi_out || !work || !prm || !res) return CFMM_E_NULL;
    if (n_tokens <= 0 || b->n_tiles <= 0) return CFMM_E_SIZE;
    if (!b->r0 || !b->r1 || !b->gamma_inv) return CFMM_E_NULL;
    if (peer && (!peer->recv_acc_dev || !peer->recv_vec_dev)) return CFMM_E_NULL;
    if (peer && (peer->world < 2 || peer->world > 16 || peer->rank < 0 || peer->rank >= peer->world)) return CFMM_E_SIZE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    PersistArgs S;
    memset(&S, 0, sizeof(S));
    int rc = fill_blocked_args(b, S.B);
    if (rc) return rc;
    S.B.slab[0] = b->r0; S.B.slab[1] = b->r1; S.B.slab[2] = b->gamma_inv;
    const int n = n_tokens;
    S.n = n; S.c = c; S.a = a; S.eq = eq; S.fixed = pinned;
    unsigned char* w = static_cast<unsigned char*>(work);
    auto take = [&](size_t bytes) { unsigned char* p = w; w += align_up(bytes); return p; };
    auto vec = [&]() { return reinterpret_cast<double*>(take(8 * (size_t)n)); };
    S.hcoef = reinterpret_cast<double*>(take(8 * (size_t)S.B.M));
    S.acc[0] = reinterpret_cast<double*>(take(8 * (size_t)(n + 1)));
    S.acc[1] = reinterpret_cast<double*>(take(8 * (size_t)(n + 1)));
    S.nu[0] = nu; S.nu[1] = vec();
    S.y = reinterpret_cast<double*>(take(8 * (size_t)(n + 2))); S.diag = vec(); S.lb = vec();
    S.grad[0] = vec(); S.grad[1] = vec(); S.fr[0] = vec(); S.fr[1] = vec(); S.pg[0] = vec(); S.pg[1] = vec();
    S.dt = vec(); S.x = vec(); S.r = vec(); S.z = vec(); S.p = vec(); S.minv = vec();
    vec(); vec(); vec();
    S.ctl = reinterpret_cast<unsigned*>(take(512));
    S.res = reinterpret_cast<DevResult*>(take(sizeof(DevResult)));
    S.tol = prm->tol; S.nu_floor = prm->nu_floor; S.max_iter = prm->max_iter; S.cg_max = prm->cg_max;
    S.nu_out = nu; S.psi_out = psi_out;
    if (peer) {
        S.recv_acc = static_cast<LLCell* const*>(const_cast<void*>(peer->recv_acc_dev));
        S.recv_vec = static_cast<LLCell* const*>(const_cast<void*>(peer->recv_vec_dev));
        S.rank = peer->rank; S.world = peer->world; S.seq_acc = peer->seq_acc; S.seq_vec = peer->seq_vec;
    } else {
        S.rank = 0; S.world = 1;
    }
    DistArgs D;
    D.accr[0] = vec(); D.accr[1] = vec();
    D.y2[0] = reinterpret_cast<double*>(take(8 * (size_t)(n + 2))); D.y2[1] = reinterpret_cast<double*>(take(8 * (size_t)(n + 2)));
    D.partial = reinterpret_cast<double*>(take(8 * (size_t)2 * kQ * kSliceMax));
    D.nsl = (n + 15) / 16 < kSliceMax ? ((n + 15) / 16 > 0 ? (n + 15) / 16 : 1) : kSliceMax;
    D.P = S;
    const size_t sm = pass_smem_bytes<kTileP, kTileStages>(3);
    static int occ[2] = {-1, -1};
    const int mode = g_persist_mode ? 1 : 0;
    if (occ[mode] < 0) {
        cudaError_t e = mode ? cudaFuncSetAttribute(k_solve_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)
                             : cudaFuncSetAttribute(k_solve_dist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e == cudaSuccess)
            e = mode ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[mode], k_solve_persist, PT, sm)
                     : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[mode], k_solve_dist, PT, sm);
        if (e != cudaSuccess || occ[mode] < 1) { occ[mode] = -1; g_last_err = e; return CFMM_E_CUDA; }
    }
    const long long cap = (long long)occ[mode] * num_sms();
    const int grid = (int)(S.B.n_tiles < cap ? S.B.n_tiles : cap);
    cudaMemsetAsync(S.ctl, 0, 512, st);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(PT); cfg.dynamicSmemBytes = sm; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;       // all CTAs co-resident (they signal each other) or the launch fails
    at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (mode) cudaLaunchKernelEx(&cfg, k_solve_persist, S);
    else cudaLaunchKernelEx(&cfg, k_solve_dist, D);
    rc = check_launch();
    if (rc) return rc;
    static thread_local DevResult* hres = nullptr;          // pinned mirror of the result struct
    if (!hres && cudaHostAlloc(&hres, sizeof(DevResult), cudaHostAllocDefault) != cudaSuccess) return CFMM_E_CUDA;
    cudaMemcpyAsync(hres, S.res, sizeof(DevResult), cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) { g_last_err = cudaGetLastError(); return CFMM_E_CUDA; }
    res->dual_value = hres->dual_value; res->primal_value = hres->primal_value; res->gap = hres->gap;
    res->primal_infeas = hres->primal_infeas; res->err = hres->err;
    res->iters = hres->iters; res->evals = hres->evals; res->hvps = hres->hvps; res->status = hres->status;
    if (peer) { peer->seq_acc = hres->seq_acc; peer->seq_vec = hres->seq_vec; }
    for (int k = 0; k < 16; ++k) g_last_prof[k] = hres->prof[k];
    return hres->status == 3 ? CFMM_E_STATE : CFMM_OK;
}

/* experiments: 0 = every CTA owns a token slice and all CTAs decide alike (default); 1 = CTA 0 owns the vector algebra */
int cfmm_set_persist_mode(int32_t mode) {
    if (mode != 0 && mode != 1) return CFMM_E_KIND;
    g_persist_mode = mode;
    return CFMM_OK;
}

/* CTA 0's clock64 totals of the last cfmm_persist_solve of this thread (SM cycles): [0..2] its own evaluation / Hessian-
 * product / diagonal passes, [3] waiting for the rest of the grid, [4..6] vector algebra + decision after evaluation /
 * Hessian-product / diagonal passes.  A development aid (where does the solve's time go), not a contract. */
int cfmm_persist_last_profile(int64_t* out8) {
    if (!out8) return CFMM_E_NULL;
    for (int k = 0; k < 16; ++k) out8[k] = g_last_prof[k];
    return CFMM_OK;
}

}  // extern "C"
