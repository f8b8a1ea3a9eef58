// cfmm_blocked.cu -- token-blocked layout for 2-token pools: evaluation, Hessian-vector product and
// Hessian diagonal WITHOUT per-pool atomics.
//
// Why: psi = sum_i A_i (L_i - D_i) (arbitrage.py:54) is a scatter of 2 values per pool into n_tokens bins.
// With red.global.add.f64 per value the L2 atomic units bound the kernel at ~8x the HBM time (measured,
// profiles/r1b_*).  The sparsity pattern (local_indices, arbitrage.py:6-12) is static across dual
// iterations, so it is preprocessed once into tiles of P pools whose tokens fall in two narrow token blocks:
//   * a tile touches few distinct tokens: nu is gathered once per tile into shared memory (nu_local) and the
//     pools address it with 16-bit local ids (4 B/pool instead of 8 B of global indices);
//   * each pool thread writes its two net flows to a shared-memory array f[2P] (no atomics);
//   * "rows" = (token, <=32 entries of f) listed by a per-tile CSR are summed by one thread each, in a fixed
//     order (bit-reproducible), and only the row totals go to global memory: ~0.25 red.add per pool instead of 2.
// HBM bytes per pool: 3 x 8 (R0, R1, 1/gamma) + 4 (local ids) + 4 (row entries) + ~1-2 (row/token tables).
// Pool slabs and the per-tile tables are staged through a shared-memory ring by 1-D bulk TMA copies
// (cp.async.bulk + mbarrier), several tiles in flight per CTA.
#include <math.h>

#include "cfmm_dev.cuh"

using namespace cfmm;

namespace {

template <int P>
struct BlockedCfg {
    // the layout builder guarantees <= P distinct tokens per tile (tiles that would exceed it go to the
    // plain bucket), so rows <= P + 2P/32 (every token one row, plus one extra row per 32 entries)
    static constexpr int kTokMax = P;
    static constexpr int kRowCapMin = 8;                     // smallest row cap the tables are sized for
    static constexpr int kRowsMax = P + 2 * P / kRowCapMin + 8;

};

// one ring stage: NF per-pool f64 slabs + local ids + row entries + row table + token list
template <int P, int NF>
struct __align__(128) Stage {
    double a[NF][P];
    uint32_t lid[P];                              // lid0 | lid1 << 16
    uint32_t pos[P];                              // where this pool's two flows go in the row-ordered array g: pos0 | pos1 << 16
    uint32_t rows[BlockedCfg<P>::kRowsMax];       // start :16 | length (1..32) :6 | local token :10, longest first
    int32_t tok[BlockedCfg<P>::kTokMax];          // local token id -> global token id
    int4 desc;                                    // (ntok, nrow, groups, 0) of the tile in this stage
};

struct BlockedArgs {
    long long n_tiles;
    long long M;                  // n_tiles * P (padded pool count = slab stride)
    const double* slab[3];        // NF slabs, each [M]
    const uint32_t* lid;          // [M]
    const uint32_t* pos;          // [M]
    const uint32_t* rows;         // [n_tiles][kRowsMax]
    const int32_t* tok;           // [n_tiles][kTokMax]
    const int4* desc;             // [n_tiles] (ntok, nrow, 0, 0)
    const double* vec;            // nu (eval) or vt (hvp); unused for diag
    double* out;                  // psi / y / diag (+= via one red.add per row)
    double* zero_next;            // optional: buffer of n_zero doubles this launch clears for the NEXT call
    int n_zero;
    double* arb;                  // eval only
    double* delta;                // eval, optional: [2][M] blocked order
    double* lambda;
    double* hcoef;                // eval, optional: [M]
};

__device__ __forceinline__ unsigned round16(unsigned bytes) { return (bytes + 15u) & ~15u; }

template <int P, int NF>
__device__ __forceinline__ void issue_tile(Stage<P, NF>* st, uint64_t* bar, const BlockedArgs& A, long long tile,
                                           const int4 d) {
    const unsigned rows_b = round16(4u * (unsigned)d.y);
    const unsigned tok_b = round16(4u * (unsigned)d.x);
    mbar_expect_tx(bar, (unsigned)(NF * P * 8 + P * 4 + P * 4 + 16) + rows_b + tok_b);
    bulk_g2s(&st->desc, A.desc + tile, 16, bar);
#pragma unroll
    for (int k = 0; k < NF; ++k) bulk_g2s(st->a[k], A.slab[k] + tile * P, P * 8, bar);
    bulk_g2s(st->lid, A.lid + tile * P, P * 4, bar);
    bulk_g2s(st->pos, A.pos + tile * P, P * 4, bar);
    bulk_g2s(st->rows, A.rows + tile * BlockedCfg<P>::kRowsMax, rows_b, bar);
    bulk_g2s(st->tok, A.tok + tile * BlockedCfg<P>::kTokMax, tok_b, bar);
}

// ---- per-pool operator: the two net flows (f0, f1) of a constant-product pool (arbitrage.py:68-70) ----------
// With gi = 1/gamma, p_j = nu_j R_j and v = rsqrt(p0 p1 gi):  a = p0 v, b = p1 v  (a b gi = 1).  The KKT solution
// x_j = clip(R_j; gamma M/(2 nu_j), M/(2 nu_j)) reads  x0/R0 = clamp(1, b, b gi),  x1/R1 = clamp(1, a, a gi):
// ratio > 1 => token tendered (Delta = R (ratio-1)/gamma), ratio < 1 => token received (Lambda = R (1-ratio)),
// ratio = 1 on both => no-trade cone.  Branch-free, so the two pools a thread owns interleave in the pipeline.
// h = sqrt(p0 p1 / gamma)/2 = w v / 2 on trading pools (Hs_i = h [[1,-1],[-1,1]] in log-price coordinates).
// 1/sqrt(w) for positive, finite, normal w without the library routine's slow-path branch (which would fence the
// two pools of a thread into separate reconvergence regions): scale the exponent into [1,4), seed with the fp32
// MUFU.RSQ, two Newton steps in fp64 (22 -> 44 -> 88 bits), scale back.  ~1 ulp.
__device__ __forceinline__ double rsqrt_pos(double w) {
    const int hi = __double2hiint(w);
    const int k = ((((hi >> 20) & 0x7ff) - 1023) >> 1);
    const double ws = __hiloint2double(hi - (k << 21), __double2loint(w));
    double y = (double)rsqrtf((float)ws);
    const double hws = 0.5 * ws;
    y = y * (1.5 - hws * y * y);
    y = y * (1.5 - hws * y * y);
    return __hiloint2double(__double2hiint(y) - (k << 20), __double2loint(y));
}

struct EvalOp {
    template <bool TRADES, bool HESS>
    __device__ __forceinline__ static void apply(const BlockedArgs& A, long long q, double R0, double R1, double gi,
                                                 double n0, double n1, double& f0, double& f1, double& acc) {
        const double p0 = n0 * R0, p1 = n1 * R1;
        const double w = p0 * p1 * gi;
        const double v = rsqrt_pos(w);
        const double a = p0 * v, b = p1 * v;
        const double d0 = 1.0 - fmin(fmax(1.0, b), b * gi);
        const double d1 = 1.0 - fmin(fmax(1.0, a), a * gi);
        f0 = R0 * d0 * (d0 < 0.0 ? gi : 1.0);
        f1 = R1 * d1 * (d1 < 0.0 ? gi : 1.0);
        acc += n0 * f0 + n1 * f1;
        if (TRADES) {
            A.delta[q] = fmax(-f0, 0.0); A.delta[A.M + q] = fmax(-f1, 0.0);
            A.lambda[q] = fmax(f0, 0.0); A.lambda[A.M + q] = fmax(f1, 0.0);
        }
        if (HESS) A.hcoef[q] = (d0 != 0.0 || d1 != 0.0) ? 0.5 * w * v : 0.0;
    }
};

// row word: start (16 bits) | length (6 bits, 1..32) | local token (10 bits).  Rows of a tile are sorted by
// decreasing length by the builder, so the 32 rows of a warp have (nearly) equal trip counts.
__device__ __forceinline__ int row_start(uint32_t r) { return (int)(r & 0xffffu); }
__device__ __forceinline__ int row_len(uint32_t r) { return (int)((r >> 16) & 0x3fu); }
__device__ __forceinline__ int row_tok(uint32_t r) { return (int)(r >> 22); }

template <int P, int THREADS, int STAGES, int MODE /*0 eval, 1 hvp, 2 diag*/, bool TRADES, bool HESS>
__global__ void __launch_bounds__(THREADS)
k_blocked(const BlockedArgs A) {
    constexpr int NF = (MODE == 0) ? 3 : 1;
    using St = Stage<P, NF>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    St* stages = reinterpret_cast<St*>(smem_raw);
    double* nul = reinterpret_cast<double*>(smem_raw + (size_t)STAGES * sizeof(St));      // [P]    nu_local
    double* g = nul + P;                                                                   // [2P] flows in ROW order
    __shared__ uint64_t full[STAGES];
    __shared__ double part[THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    // Each CTA walks a CONTIGUOUS chunk of tiles.  Tiles are sorted by (token block of slot 0, of slot 1), so at any
    // moment the resident CTAs work on different token blocks and their red.adds hit different addresses (a
    // grid-strided walk would have all CTAs hammer the same ~130 tokens at once).
    const long long t_beg = (A.n_tiles * (long long)blockIdx.x) / gridDim.x;
    const long long t_end = (A.n_tiles * (long long)(blockIdx.x + 1)) / gridDim.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            const long long t = t_beg + s;
            if (t < t_end) issue_tile<P, NF>(&stages[s], &full[s], A, t, __ldg(A.desc + t));
        }
    }
    // Programmatic dependent launch: everything above touches only this launch's own shared memory and the
    // constant pool tables, so it may run while the previous kernel on the stream is still draining.  From here
    // on we read vec / write out, zero_next -- wait for the previous grid, then let the next one start its ramp.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // clear the buffer the NEXT call accumulates into (nobody touches it during this launch)
    for (int j = blockIdx.x * THREADS + tid; j < A.n_zero; j += gridDim.x * THREADS) A.zero_next[j] = 0.0;
    double acc = 0.0;
    int stage = 0;
    unsigned parity = 0;
    constexpr int NPRE = (P + THREADS - 1) / THREADS;       // nu_local values each thread prefetches
    // prologue: nu_local of this CTA's first tile
    if (t_beg < t_end) {
        mbar_wait(&full[0], 0);
        if (MODE != 2) {
            const int ntok = stages[0].desc.x;
            for (int t = tid; t < ntok; t += THREADS) nul[t] = __ldg(A.vec + stages[0].tok[t]);
        }
    }
    __syncthreads();
    for (long long tile = t_beg; tile < t_end; ++tile) {
        St& S = stages[stage];                          // full (waited for when its nu_local was fetched)
        const int4 d = S.desc;                          // (ntok, nrow, groups, -)
        // the producer thread fetches the descriptor of the tile it will issue at the end of this iteration
        const long long far = tile + STAGES;
        int4 dfar = make_int4(0, 0, 0, 0);
        if (tid == 0 && far < t_end) dfar = __ldg(A.desc + far);
        // ---- phase 2: per-pool flows, scattered into row order.  All loads and math of the thread's NPOOL pools
        // first, the shared-memory stores afterwards, so the independent chains overlap in the pipeline.
        {
            constexpr int NPOOL = P / THREADS;
            double f0[NPOOL], f1[NPOOL];
            uint32_t ps[NPOOL];
#pragma unroll
            for (int u = 0; u < NPOOL; ++u) {
                const int l = tid + u * THREADS;
                const uint32_t li = S.lid[l];
                ps[u] = S.pos[l];
                if (MODE == 0) {
                    EvalOp::apply<TRADES, HESS>(A, tile * P + l, S.a[0][l], S.a[1][l], S.a[2][l], nul[li & 0xffffu],
                                                nul[li >> 16], f0[u], f1[u], acc);
                } else if (MODE == 1) {
                    f0[u] = S.a[0][l] * (nul[li & 0xffffu] - nul[li >> 16]);
                    f1[u] = -f0[u];
                } else {
                    f0[u] = S.a[0][l];
                    f1[u] = f0[u];
                }
            }
#pragma unroll
            for (int u = 0; u < NPOOL; ++u) {
                g[ps[u] & 0xffffu] = f0[u];
                g[ps[u] >> 16] = f1[u];
            }
        }
        __syncthreads();                 // g complete; nu_local of this tile is dead from here on
        // ---- prefetch nu_local of the NEXT tile into registers: the L2 latency hides behind the row phase
        const long long nxt = tile + 1;
        int nstage = stage + 1;
        unsigned nparity = parity;
        if (nstage == STAGES) { nstage = 0; nparity ^= 1u; }
        double pre[NPRE];
        int ntok_n = 0;
        if (nxt < t_end) {
            mbar_wait(&full[nstage], nparity);           // also makes the next iteration's stage reads safe
            if (MODE != 2) {
                ntok_n = stages[nstage].desc.x;
#pragma unroll
                for (int k = 0; k < NPRE; ++k) {
                    const int t = tid + k * THREADS;
                    pre[k] = (t < ntok_n) ? __ldg(A.vec + stages[nstage].tok[t]) : 0.0;
                }
            }
        }
        // ---- phase 3: one thread per row; a row is a CONTIGUOUS run of g (the pool phase scattered the flows
        // into row order), rows are sorted by length so a warp's 32 rows have (nearly) equal trip counts.
        // Fixed summation order; one red.add per row.
        for (int r = tid; r < d.y; r += THREADS) {
            const uint32_t rw = S.rows[r];
            const double* q = g + row_start(rw);
            const int len = row_len(rw);
            double s0 = 0.0, s1 = 0.0;
            int k = 0;
            for (; k + 4 <= len; k += 4) { s0 += q[k] + q[k + 2]; s1 += q[k + 1] + q[k + 3]; }
            for (; k < len; ++k) s0 += q[k];
            const double s = s0 + s1;
            if (s != 0.0) atomicAdd(A.out + S.tok[row_tok(rw)], s);
        }
        if (MODE != 2 && nxt < t_end) {
#pragma unroll
            for (int k = 0; k < NPRE; ++k) {
                const int t = tid + k * THREADS;
                if (t < ntok_n) nul[t] = pre[k];
            }
        }
        __syncthreads();                 // stage and f are free again; nu_local of the next tile is in place
        if (tid == 0 && far < t_end) {
            fence_proxy_async();
            issue_tile<P, NF>(&S, &full[stage], A, far, dfar);
        }
        stage = nstage; parity = nparity;
    }
    if (MODE == 0) {
        acc = warp_sum(acc);
        if (lane == 0) part[warp] = acc;
        __syncthreads();
        if (tid < 32) {
            double s = (tid < THREADS / 32) ? part[tid] : 0.0;
            s = warp_sum(s);
            if (tid == 0 && s != 0.0) atomicAdd(A.arb, s);
        }
    }
}

// ---- shipped configurations (selectable at run time for tuning; the layout must be built for the same P)
struct Cfg0 { static constexpr int P = 1024, T = 512, S = 2, CTAS = 2; };   // 2 x (85 + 24) KB smem per SM
struct Cfg1 { static constexpr int P = 512, T = 512, S = 3, CTAS = 2; };    // 2 x (64 + 12) KB, deeper ring
struct Cfg2 { static constexpr int P = 512, T = 256, S = 2, CTAS = 4; };    // 4 x (43 + 12) KB
int g_cfg = 0;
int g_pdl = 1;
int g_row_cap = 32;

template <class C, int MODE, bool TRADES, bool HESS>
int launch_cfg(const BlockedArgs& A, cudaStream_t st) {
    constexpr int NF = (MODE == 0) ? 3 : 1;
    auto kern = k_blocked<C::P, C::T, C::S, MODE, TRADES, HESS>;
    const size_t sm = (size_t)C::S * sizeof(Stage<C::P, NF>) + (size_t)(3 * C::P) * sizeof(double);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        attr = true;
    }
    const long long cap = (long long)C::CTAS * num_sms();
    const int grid = (int)(A.n_tiles < cap ? A.n_tiles : cap);
    if (g_pdl) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(C::T); cfg.dynamicSmemBytes = sm; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, kern, A);
    } else {
        kern<<<grid, C::T, sm, st>>>(A);
    }
    return check_launch();
}

template <int MODE, bool TRADES, bool HESS>
int launch_blocked(const BlockedArgs& A, cudaStream_t st) {
    switch (g_cfg) {
        case 0: return launch_cfg<Cfg0, MODE, TRADES, HESS>(A, st);
        case 2: return launch_cfg<Cfg2, MODE, TRADES, HESS>(A, st);
        default: return launch_cfg<Cfg1, MODE, TRADES, HESS>(A, st);
    }
}

int cfg_P() { return g_cfg == 0 ? Cfg0::P : Cfg1::P; }

int fill_args(const cfmm_blocked_pairs* b, BlockedArgs& A) {
    if (!b) return CFMM_E_NULL;
    if (b->pools_per_tile != cfg_P()) return CFMM_E_KIND;
    const int64_t P = b->pools_per_tile;
    if (b->n_tiles < 0 || b->n_pools < 0 || b->n_pools > b->n_tiles * P) return CFMM_E_SIZE;
    if (b->n_tiles > 0 && (!b->lid || !b->pos || !b->rows || !b->tok || !b->desc)) return CFMM_E_NULL;
    A.n_tiles = b->n_tiles;
    A.M = b->n_tiles * P;
    A.lid = b->lid; A.pos = b->pos; A.rows = b->rows; A.tok = b->tok;
    A.desc = reinterpret_cast<const int4*>(b->desc);
    A.zero_next = nullptr; A.n_zero = 0;
    A.slab[0] = A.slab[1] = A.slab[2] = nullptr;
    A.vec = nullptr; A.out = nullptr; A.arb = nullptr; A.delta = A.lambda = A.hcoef = nullptr;
    return CFMM_OK;
}

}  // namespace

extern "C" {

int cfmm_blocked_layout_info(int32_t* pools_per_tile, int32_t* rows_stride, int32_t* tok_stride, int32_t* row_cap,
                             int32_t* ent_stride) {
    const int P = cfg_P();
    const int rows = P + 2 * P / 8 + 8;
    if (pools_per_tile) *pools_per_tile = P;
    if (rows_stride) *rows_stride = rows;
    if (tok_stride) *tok_stride = P;
    if (row_cap) *row_cap = g_row_cap;
    if (ent_stride) *ent_stride = 0;          /* unused since the row-ordered scatter (kept for ABI stability) */
    return CFMM_OK;
}

int cfmm_set_blocked_config(int32_t cfg) {
    if (cfg >= 300) { const int c = cfg - 300; if (c < 8 || c > 32) return CFMM_E_KIND; g_row_cap = c; return CFMM_OK; }
    if (cfg >= 200) { g_pdl = cfg - 200; return CFMM_OK; }      // 200 / 201: programmatic dependent launch off / on
    if (cfg < 0 || cfg > 2) return CFMM_E_KIND;
    g_cfg = cfg;
    return CFMM_OK;
}

int cfmm_blocked_eval(const cfmm_blocked_pairs* b, int32_t n_tokens, const double* nu, double* psi, double* arb,
                      const cfmm_eval_out* out, double* zero_next, int64_t n_zero, void* stream) {
    BlockedArgs A;
    int rc = fill_args(b, A);
    if (rc) return rc;
    if (n_tokens <= 0) return CFMM_E_SIZE;
    if (!nu || !psi || !arb) return CFMM_E_NULL;
    if (b->n_tiles == 0) return CFMM_OK;
    if (!b->r0 || !b->r1 || !b->gamma_inv) return CFMM_E_NULL;
    A.slab[0] = b->r0; A.slab[1] = b->r1; A.slab[2] = b->gamma_inv;
    A.vec = nu; A.out = psi; A.arb = arb;
    A.zero_next = zero_next; A.n_zero = zero_next ? (int)n_zero : 0;
    const bool trades = out && out->delta && out->lambda;
    const bool hess = out && out->hcoef;
    if (trades) { A.delta = out->delta; A.lambda = out->lambda; }
    if (hess) A.hcoef = out->hcoef;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (trades && hess) return launch_blocked<0, true, true>(A, st);
    if (trades) return launch_blocked<0, true, false>(A, st);
    if (hess) return launch_blocked<0, false, true>(A, st);
    return launch_blocked<0, false, false>(A, st);
}

int cfmm_blocked_hvp(const cfmm_blocked_pairs* b, int32_t n_tokens, const double* hcoef, const double* vt, double* y,
                     double* zero_next, void* stream) {
    BlockedArgs A;
    int rc = fill_args(b, A);
    if (rc) return rc;
    if (n_tokens <= 0) return CFMM_E_SIZE;
    if (!hcoef || !vt || !y) return CFMM_E_NULL;
    if (b->n_tiles == 0) return CFMM_OK;
    A.slab[0] = hcoef; A.vec = vt; A.out = y;
    A.zero_next = zero_next; A.n_zero = zero_next ? n_tokens : 0;
    return launch_blocked<1, false, false>(A, static_cast<cudaStream_t>(stream));
}

int cfmm_blocked_diag(const cfmm_blocked_pairs* b, int32_t n_tokens, const double* hcoef, double* diag, void* stream) {
    BlockedArgs A;
    int rc = fill_args(b, A);
    if (rc) return rc;
    if (n_tokens <= 0) return CFMM_E_SIZE;
    if (!hcoef || !diag) return CFMM_E_NULL;
    if (b->n_tiles == 0) return CFMM_OK;
    A.slab[0] = hcoef; A.out = diag;
    return launch_blocked<2, false, false>(A, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
