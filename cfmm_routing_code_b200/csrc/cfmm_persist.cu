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
constexpr unsigned long long kSpinLimit = 6000000000ull;     // ~3 s of SM clocks: a lost peer / launch must not hang the GPU

enum { OP_DONE = 0, OP_EVAL = 1, OP_HVP = 2, OP_DIAG = 3 };
enum { ST_EVAL0 = 0, ST_DIAG, ST_HVP, ST_TRIAL };

struct DevResult {                               // written by CTA 0, copied to the host after the kernel
    double dual_value, primal_value, gap, primal_infeas, err;
    int iters, evals, hvps, status;              // status 0 optimal, 1 max_iter, 2 stalled, 3 aborted (spin limit)
    unsigned long long seq_acc, seq_vec;
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
    unsigned* ctl;                               // [0] arrive counter, [1] epoch flag, [2] command, [3] abort
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

// ---- CTA-wide reductions (PT threads), result broadcast to every thread ------------------------------------------------
__device__ __forceinline__ double cta_sum(double v, double* sh) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < PT / 32; ++w) t += sh[w];          // same order in every thread: same bits
    return t;
}
__device__ __forceinline__ double cta_max(double v, double* sh) {
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < PT / 32; ++w) t = fmax(t, sh[w]);
    return t;
}

// KKT residual (max of the value-weighted and the per-token one, as solver.py::kkt) at (nu, acc) into set s
struct Kkt { double err, g, primal, infeas; };
__device__ Kkt kkt_eval(const PersistArgs& S, const double* nu, const double* acc, double thr, double* grad, double* fr,
                        double* pg, double* sh) {
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, m0 = 0, m1 = 0;
    for (int j = threadIdx.x; j < S.n; j += PT) {
        const double nj = __ldcg(nu + j), pj = __ldcg(acc + j), aj = S.a[j], cj = S.c[j];
        const double g = aj + pj;
        const bool near = (nj <= S.lb[j] * (1.0 + thr)) && !S.eq[j];
        const bool act = S.fixed[j] || (near && g > 0.0);
        const double f = act ? 0.0 : 1.0;
        const double v = nj * g * f;
        grad[j] = g; fr[j] = f; pg[j] = v;
        s0 += fabs(v);
        s1 += (nj - cj) * aj;
        s2 += nj * fabs(g);
        s3 += cj * pj;
        const double viol = S.fixed[j] ? 0.0 : (S.eq[j] ? fabs(g) : fmax(-g, 0.0));
        s4 += nj * viol;
        m0 = fmax(m0, fabs(g) * f);
        m1 = fmax(m1, fmax(fabs(aj), S.fixed[j] ? 0.0 : fabs(pj)));
    }
    s0 = cta_sum(s0, sh); s1 = cta_sum(s1, sh); s2 = cta_sum(s2, sh); s3 = cta_sum(s3, sh); s4 = cta_sum(s4, sh);
    m0 = cta_max(m0, sh); m1 = cta_max(m1, sh);
    Kkt k;
    k.g = s1 + __ldcg(acc + S.n);
    k.err = fmax(s0 / fmax(fmax(fabs(k.g), 1e-3 * s2), 1e-300), m0 / fmax(m1, 1e-300));
    k.primal = s3;
    k.infeas = s4 / fmax(fabs(k.g), 1e-300);
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
                                         uint64_t* full, unsigned& phase, long long t_beg, long long t_end, double* part) {
    BlockedArgs A = S.B;
    A.vec = vec; A.out = out; A.zero_next = nullptr; A.n_zero = 0;
    asm volatile("fence.proxy.async;" ::: "memory");         // hcoef: written through the generic proxy, read by TMA (async proxy)
    if (MODE == 0) { A.arb = out + S.n; A.hcoef = S.hcoef; }
    else A.slab[0] = S.hcoef;
    double acc = 0.0;
    blocked_pass<kTileP, PT, kTileStages, MODE, false, MODE == 0, true, false>(A, smem_raw, full, phase, t_beg, t_end, acc);
    if (MODE == 0) {
        cta_accumulate<PT>(acc, part, A.arb);
        asm volatile("fence.proxy.async;" ::: "memory");     // hcoef written here (generic proxy) is read by TMA in later passes
    }
}

// CTA 0's solver state between passes.  Lives in shared memory, not registers: every CTA runs the same kernel and the
// pass loop must not carry (and spill) a dozen doubles that only CTA 0 ever uses.
struct BossState {
    int state, cur, iters, evals, hvps, status, cg_k, ls, aborted;
    double err, g0, rz, r0n, eta, alpha, lin1;
    Kkt kc;
    unsigned long long seq_acc, seq_vec;
};

// What CTA 0 does after the grid finished a pass (`op` on set `set`): all-reduce over the peers, vector algebra,
// decision.  Returns the next command.  Every thread of the CTA computes the same scalars (broadcast reductions).
__device__ __noinline__ unsigned boss_step(const PersistArgs& S, BossState* bs, unsigned op, unsigned set, double* sh) {
    const int tid = threadIdx.x, n = S.n;
    int state = bs->state, cur = bs->cur, iters = bs->iters, evals = bs->evals, hvps = bs->hvps, status = bs->status;
    int cg_k = bs->cg_k, ls = bs->ls;
    bool aborted = bs->aborted != 0;
    double err = bs->err, g0 = bs->g0, rz = bs->rz, r0n = bs->r0n, eta = bs->eta, alpha = bs->alpha, lin1 = bs->lin1;
    Kkt kc = bs->kc;
    unsigned long long seq_acc = bs->seq_acc, seq_vec = bs->seq_vec;
    __syncthreads();                                 // everybody has read the state before thread 0 rewrites it below
    bool finish = aborted;
    unsigned next = OP_DONE;
    if (!finish && op == OP_EVAL) {
        double* acc = S.acc[set];
        if (S.world > 1) { ++seq_acc; if (!ll_allreduce_cta(acc, n + 1, S.recv_acc, S.rank, S.world, seq_acc)) finish = aborted = true; }
        ++evals;
        if (!finish && state == ST_EVAL0) {
            kc = kkt_eval(S, S.nu[cur], acc, 1e-2, S.grad[cur], S.fr[cur], S.pg[cur], sh);
            err = kc.err; g0 = kc.g;
            goto newton_start;
        } else if (!finish) {                       // ST_TRIAL: Armijo test along nu * exp(alpha dt)
            const int tr = cur ^ 1;
            double s1 = 0.0, s2 = 0.0;
            for (int j = tid; j < n; j += PT) {
                const double nt = S.nu[tr][j], nc = S.nu[cur][j];
                s1 += (nt - S.c[j]) * S.a[j];
                s2 += S.grad[cur][j] * (nt - nc);
            }
            s1 = cta_sum(s1, sh); s2 = cta_sum(s2, sh);
            const double gt = s1 + __ldcg(acc + n), lin = s2;
            const double thr = fmin(1e-2, fmax(isfinite(err) ? err : 1e-2, 1e-14));
            const Kkt kt = kkt_eval(S, S.nu[tr], acc, thr, S.grad[tr], S.fr[tr], S.pg[tr], sh);
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
            double s = 0.0;
            for (int j = tid; j < n; j += PT) {
                const double mi = S.fr[cur][j] / fmax(__ldcg(S.diag + j), 1e-300);
                const double r = -S.pg[cur][j];
                const double z = mi * r;
                S.minv[j] = mi; S.x[j] = 0.0; S.r[j] = r; S.z[j] = z; S.p[j] = z;
                s += r * z;
            }
            rz = cta_sum(s, sh);
            r0n = sqrt(fmax(rz, 0.0));
            eta = fmin(0.1, sqrt(err));
            cg_k = 0;
            if (rz <= 0.0) goto direction;
            for (int j = tid; j < n; j += PT) S.y[j] = 0.0;
            next = OP_HVP; state = ST_HVP;
        }
    } else if (!finish) {                           // OP_HVP: one PCG iteration with y = Hs p
        if (S.world > 1) { ++seq_vec; if (!ll_allreduce_cta(S.y, n, S.recv_vec, S.rank, S.world, seq_vec)) finish = aborted = true; }
        ++hvps;
        if (!finish) {
            double pHp = 0.0, pdp = 0.0;
            const double* fr = S.fr[cur];
            for (int j = tid; j < n; j += PT) {
                const double hp = __ldcg(S.y + j) * fr[j], pj = S.p[j];
                pHp += pj * hp;
                pdp += pj * pj * fmax(__ldcg(S.diag + j), 1e-300);
            }
            pHp = cta_sum(pHp, sh); pdp = cta_sum(pdp, sh);
            bool stop;
            if (pHp <= 1e-14 * pdp) {               // homogeneity direction: g is linear along nu
                if (cg_k == 0)
                    for (int j = tid; j < n; j += PT) S.x[j] = S.p[j];
                stop = true;
            } else {
                const double al = rz / pHp;
                double rzn = 0.0;
                for (int j = tid; j < n; j += PT) {
                    const double hp = __ldcg(S.y + j) * fr[j];
                    S.x[j] += al * S.p[j];
                    const double r = S.r[j] - al * hp;
                    const double z = S.minv[j] * r;
                    S.r[j] = r; S.z[j] = z;
                    rzn += r * z;
                }
                rzn = cta_sum(rzn, sh);
                stop = (rzn <= 0.0) || (sqrt(fmax(rzn, 0.0)) <= eta * r0n);
                if (!stop) {
                    const double be = rzn / rz;
                    for (int j = tid; j < n; j += PT) S.p[j] = S.z[j] + be * S.p[j];
                }
                rz = rzn;
            }
            ++cg_k;
            if (!stop && cg_k < S.cg_max) {
                for (int j = tid; j < n; j += PT) S.y[j] = 0.0;
                next = OP_HVP;
            } else {
                goto direction;
            }
        }
    }
    goto decided;

newton_start:                                   // same counting as cfmm_solver.cu: the final check is an iteration too
    if (iters >= S.max_iter) { status = 1; finish = true; goto decided; }
    ++iters;
    if (err <= S.tol) { status = 0; finish = true; goto decided; }
    for (int j = tid; j < n; j += PT) S.diag[j] = 0.0;
    next = OP_DIAG; state = ST_DIAG;
    goto decided;

direction: {
        // dt <- x if it is a descent direction in value units (pg . x < 0), else scaled steepest descent
        double s = 0.0, mx = 0.0;
        const double* pg = S.pg[cur];
        for (int j = tid; j < n; j += PT) { s += pg[j] * S.x[j]; mx = fmax(mx, fabs(pg[j])); }
        s = cta_sum(s, sh); mx = cta_max(mx, sh);
        const bool ok = isfinite(s) && s < 0.0;
        for (int j = tid; j < n; j += PT) S.dt[j] = ok ? S.x[j] : -pg[j] / fmax(mx, 1e-300);
        alpha = 1.0; ls = 0; lin1 = 0.0;
    }
take_step: {
        const int tr = cur ^ 1;
        __syncthreads();
        for (int j = tid; j < n; j += PT) {
            const double e = fmin(fmax(alpha * S.dt[j], -20.0), 20.0);
            const double v = fmax(S.nu[cur][j] * exp(e), S.lb[j]);
            S.nu[tr][j] = S.fixed[j] ? S.c[j] : v;
        }
        for (int j = tid; j <= n; j += PT) S.acc[tr][j] = 0.0;
        next = OP_EVAL | ((unsigned)tr << 4); state = ST_TRIAL;
    }

decided:
    if (finish) {
        // results at the accepted point: nu[cur], acc[cur] (+ the KKT data of that point in kc)
        if (status == 1 && err <= S.tol) status = 0;
        for (int j = tid; j < n; j += PT) {
            const double v = S.nu[cur][j];
            S.psi_out[j] = __ldcg(S.acc[cur] + j);
            if (cur != 0) S.nu_out[j] = v;
        }
        if (tid == 0) {
            DevResult R;
            R.dual_value = kc.g; R.primal_value = kc.primal;
            R.gap = (kc.g - kc.primal) / fmax(fabs(kc.g), 1e-300);
            R.primal_infeas = kc.infeas; R.err = err;
            R.iters = iters; R.evals = evals; R.hvps = hvps; R.status = aborted ? 3 : status;
            R.seq_acc = seq_acc; R.seq_vec = seq_vec;
            *S.res = R;
        }
        next = OP_DONE;
    }
    if (tid == 0) {
        bs->state = state; bs->cur = cur; bs->iters = iters; bs->evals = evals; bs->hvps = hvps; bs->status = status;
        bs->cg_k = cg_k; bs->ls = ls; bs->aborted = aborted ? 1 : 0;
        bs->err = err; bs->g0 = g0; bs->rz = rz; bs->r0n = r0n; bs->eta = eta; bs->alpha = alpha; bs->lin1 = lin1;
        bs->kc = kc; bs->seq_acc = seq_acc; bs->seq_vec = seq_vec;
    }
    __syncthreads();
    return next;
}

__global__ void __launch_bounds__(PT, kCtasPerSm)
k_solve_persist(const __grid_constant__ PersistArgs S) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full[kTileStages];
    __shared__ double part[PT / 32];
    __shared__ double sh[PT / 32];
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
            bs.err = INFINITY; bs.g0 = 0.0; bs.rz = 0.0; bs.r0n = 0.0; bs.eta = 0.1; bs.alpha = 1.0; bs.lin1 = 0.0;
            bs.kc.err = INFINITY; bs.kc.g = 0.0; bs.kc.primal = 0.0; bs.kc.infeas = 0.0;
            bs.seq_acc = S.seq_acc; bs.seq_vec = S.seq_vec;
        }
        for (int j = tid; j < n; j += PT) {
            const double l = S.eq[j] ? S.nu_floor : fmax(S.c[j], S.nu_floor);
            S.lb[j] = l;
            S.nu[0][j] = S.fixed[j] ? S.c[j] : fmax(S.nu[0][j], l);
        }
        for (int j = tid; j <= n; j += PT) S.acc[0][j] = 0.0;
        __threadfence();
        __syncthreads();
        if (tid == 0) { S.ctl[2] = OP_EVAL; st_release_gpu(S.ctl + 1, 1u); }
    }
    __syncthreads();
    unsigned want = 1u;                              // epoch flag value that carries this round's command
    for (;;) {
        // ---- wait for CTA 0's command of this round
        if (tid == 0) {
            const long long t0 = clock64();
            unsigned c = OP_DONE;
            bool got = true;
            while (ld_acquire_gpu(S.ctl + 1) < want) {
                if ((unsigned long long)(clock64() - t0) > kSpinLimit || ld_acquire_gpu(S.ctl + 3) != 0u) { got = false; break; }
            }
            if (got) c = ld_acquire_gpu(S.ctl + 2);
            s_cmd = c;
        }
        __syncthreads();
        const unsigned cmd = s_cmd;
        const unsigned op = cmd & 15u, set = (cmd >> 4) & 1u;
        if (op == OP_DONE) break;
        // ---- the pass, on this CTA's tiles
        if (op == OP_EVAL) run_pass<0>(S, S.nu[set], S.acc[set], smem_raw, full, phase, t_beg, t_end, part);
        else if (op == OP_HVP) run_pass<1>(S, S.p, S.y, smem_raw, full, phase, t_beg, t_end, part);
        else run_pass<2>(S, nullptr, S.diag, smem_raw, full, phase, t_beg, t_end, part);
        ++epoch;
        // ---- arrive: this CTA's red.adds are ordered before the counter bump
        __threadfence();
        __syncthreads();
        if (tid == 0) atomicAdd(S.ctl, 1u);
        ++want;
        if (!boss) continue;
        // ---- CTA 0: wait for the whole grid, then vector algebra + decision
        if (tid == 0) {
            const long long t0 = clock64();
            const unsigned target = epoch * gridDim.x;
            while (ld_acquire_gpu(S.ctl) < target) {
                if ((unsigned long long)(clock64() - t0) > kSpinLimit) { S.ctl[3] = 1u; bs.aborted = 1; break; }
            }
        }
        __syncthreads();
        const unsigned next = boss_step(S, &bs, op, set, sh);
        __threadfence();
        __syncthreads();
        if (tid == 0) { S.ctl[2] = next; st_release_gpu(S.ctl + 1, want); }
    }
}

inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

extern "C" {

int64_t cfmm_persist_solve_work_bytes(const cfmm_blocked_pairs* b, int32_t n_tokens) {
    if (!b || n_tokens <= 0) return CFMM_E_SIZE;
    const size_t n = (size_t)n_tokens;
    size_t bytes = align_up(8 * (size_t)b->n_tiles * (size_t)b->pools_per_tile);      // hcoef
    bytes += 2 * align_up(8 * (n + 1));       // [psi | arb] of the current / trial point
    bytes += 19 * align_up(8 * n);            // nu trial, y, diag, lb, grad x2, fr x2, pg x2, dt, x, r, z, p, minv (+3 spare)
    bytes += align_up(64) + align_up(sizeof(DevResult));
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
    S.y = vec(); S.diag = vec(); S.lb = vec();
    S.grad[0] = vec(); S.grad[1] = vec(); S.fr[0] = vec(); S.fr[1] = vec(); S.pg[0] = vec(); S.pg[1] = vec();
    S.dt = vec(); S.x = vec(); S.r = vec(); S.z = vec(); S.p = vec(); S.minv = vec();
    vec(); vec(); vec();
    S.ctl = reinterpret_cast<unsigned*>(take(64));
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
    const size_t sm = pass_smem_bytes<kTileP, kTileStages>(3);
    static int occ = -1;
    if (occ < 0) {
        if (cudaFuncSetAttribute(k_solve_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_solve_persist, PT, sm) != cudaSuccess || occ < 1) {
            occ = -1; g_last_err = cudaGetLastError();
            return CFMM_E_CUDA;
        }
    }
    const long long cap = (long long)occ * num_sms();
    const int grid = (int)(S.B.n_tiles < cap ? S.B.n_tiles : cap);
    cudaMemsetAsync(S.ctl, 0, 64, st);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(PT); cfg.dynamicSmemBytes = sm; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;       // all CTAs co-resident (they signal each other) or the launch fails
    at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k_solve_persist, S);
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
    return hres->status == 3 ? CFMM_E_STATE : CFMM_OK;
}

}  // extern "C"
