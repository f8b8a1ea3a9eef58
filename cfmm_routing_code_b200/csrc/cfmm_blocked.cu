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
    static constexpr int kRowsMax = P + 2 * P / 32 + 8;
    // rows are padded to multiples of 4 entries (pad code kZeroSlot reads a 0.0), so <= 2P + 3 per row
    static constexpr int kEntMax = (2 * P + 3 * kRowsMax + 7) / 8 * 8;
    static constexpr int kZeroSlot = 2 * P;
};

// one ring stage: NF per-pool f64 slabs + local ids + row entries + row table + token list
template <int P, int NF>
struct __align__(128) Stage {
    double a[NF][P];
    uint32_t lid[P];                              // lid0 | lid1 << 16
    uint16_t ent[BlockedCfg<P>::kEntMax];         // row-ordered, rows padded to 4: local_pool << 1 | slot
    uint32_t rows[BlockedCfg<P>::kRowsMax];       // start/4 :16 | groups of 4 entries :6 | local token :10
    int32_t tok[BlockedCfg<P>::kTokMax];          // local token id -> global token id
};

struct BlockedArgs {
    long long n_tiles;
    long long M;                  // n_tiles * P (padded pool count = slab stride)
    const double* slab[3];        // NF slabs, each [M]
    const uint32_t* lid;          // [M]
    const uint16_t* ent;          // [n_tiles][2P]
    const uint32_t* rows;         // [n_tiles][kRowsMax]
    const int32_t* tok;           // [n_tiles][kTokMax]
    const int4* desc;             // [n_tiles] (ntok, nrow, entry groups of 4, 0)
    const double* vec;            // nu (eval) or vt (hvp); unused for diag
    double* out;                  // psi / y / diag: zeroed by this kernel, filled by k_token_reduce
    int n_out;                    // entries of `out` to zero (n_tokens, +1 for arb in eval mode)
    double* partial;              // [n_tiles][rows_stride] row sums (plain stores, no atomics)
    double* arb;                  // eval only
    double* delta;                // eval, optional: [2][M] blocked order
    double* lambda;
    double* hcoef;                // eval, optional: [M]
    int n_seg;                    // token segments for k_token_reduce
    const int4* seg;              // [n_seg] (token, begin, end, multi)
    const int* pos;               // [n_rows_total] positions into partial
    int dbg;                      // MEASUREMENT ONLY: bit0 skip row phase, bit1 skip pool math, bit2 skip nu gather
};

__device__ __forceinline__ unsigned round16(unsigned bytes) { return (bytes + 15u) & ~15u; }

template <int P, int NF>
__device__ __forceinline__ void issue_tile(Stage<P, NF>* st, uint64_t* bar, const BlockedArgs& A, long long tile) {
    const int4 d = __ldg(A.desc + tile);
    const unsigned rows_b = round16(4u * (unsigned)d.y);
    const unsigned tok_b = round16(4u * (unsigned)d.x);
    const unsigned ent_b = round16(8u * (unsigned)d.z);
    mbar_expect_tx(bar, (unsigned)(NF * P * 8 + P * 4) + ent_b + rows_b + tok_b);
#pragma unroll
    for (int k = 0; k < NF; ++k) bulk_g2s(st->a[k], A.slab[k] + tile * P, P * 8, bar);
    bulk_g2s(st->lid, A.lid + tile * P, P * 4, bar);
    bulk_g2s(st->ent, A.ent + tile * BlockedCfg<P>::kEntMax, ent_b, bar);
    bulk_g2s(st->rows, A.rows + tile * BlockedCfg<P>::kRowsMax, rows_b, bar);
    bulk_g2s(st->tok, A.tok + tile * BlockedCfg<P>::kTokMax, tok_b, bar);
}

// ---- per-pool operators: produce the two values (f0, f1) that the rows sum ----------------------
// constant product with gi = 1/gamma (arbitrage.py:68-70): trade 0->1 iff nu1 R1 > nu0 R0 / gamma.
//   v = rsqrt(num den gi); t = num v = sqrt(gamma num/den); 1/t = den gi v; h = sqrt(p0 p1 / gamma)/2 = w v / 2
struct EvalOp {
    static constexpr int NF = 3;
    static constexpr bool kNeedsVec = true;
    template <bool TRADES, bool HESS>
    __device__ __forceinline__ static void apply(const BlockedArgs& A, long long q, double R0, double R1, double gi,
                                                 double n0, double n1, double& f0, double& f1, double& acc) {
        const double p0 = n0 * R0, p1 = n1 * R1;
        const bool fwd = p1 > p0 * gi;
        const bool bwd = p0 > p1 * gi;
        f0 = 0.0; f1 = 0.0;
        double h = 0.0;
        if (fwd || bwd) {
            const double num = fwd ? p1 : p0, den = fwd ? p0 : p1;
            const double w = num * den * gi;
            const double v = rsqrt(w);
            const double t = num * v;
            const double u = den * gi * v;
            const double din = (fwd ? R0 : R1) * (t - 1.0) * gi;
            const double lout = (fwd ? R1 : R0) * (1.0 - u);
            f0 = fwd ? -din : lout;
            f1 = fwd ? lout : -din;
            h = 0.5 * w * v;
            acc += n0 * f0 + n1 * f1;
        }
        if (TRADES) {
            A.delta[q] = fmax(-f0, 0.0); A.delta[A.M + q] = fmax(-f1, 0.0);
            A.lambda[q] = fmax(f0, 0.0); A.lambda[A.M + q] = fmax(f1, 0.0);
        }
        if (HESS) A.hcoef[q] = h;
    }
};

// row word: start (16 bits) | length (6 bits, 1..32) | local token (10 bits).  Rows of a tile are sorted by
// decreasing length by the builder, so the 32 rows of a warp have (nearly) equal trip counts.
__device__ __forceinline__ int row_start4(uint32_t r) { return (int)(r & 0xffffu); }
__device__ __forceinline__ int row_groups(uint32_t r) { return (int)((r >> 16) & 0x3fu); }
__device__ __forceinline__ int row_tok(uint32_t r) { return (int)(r >> 22); }

template <int P, int THREADS, int STAGES, int MODE /*0 eval, 1 hvp, 2 diag*/, bool TRADES, bool HESS>
__global__ void __launch_bounds__(THREADS)
k_blocked(const BlockedArgs A) {
    constexpr int NF = (MODE == 0) ? 3 : 1;
    using St = Stage<P, NF>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    St* stages = reinterpret_cast<St*>(smem_raw);
    double* nul0 = reinterpret_cast<double*>(smem_raw + (size_t)STAGES * sizeof(St));    // [2][P]  nu_local, double buffered
    double* f0buf = nul0 + 2 * P;                                                         // [2][2P+4] flows (+ zero slot)
    constexpr int FS = 2 * P + 4;
    __shared__ uint64_t full[STAGES];
    __shared__ double part[THREADS / 32];
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            const long long t = (long long)blockIdx.x + (long long)s * gridDim.x;
            if (t < A.n_tiles) issue_tile<P, NF>(&stages[s], &full[s], A, t);
        }
    }
    for (int j = blockIdx.x * THREADS + tid; j < A.n_out; j += gridDim.x * THREADS) A.out[j] = 0.0;
    if (tid < 8) f0buf[(tid >> 2) * FS + 2 * P + (tid & 3)] = 0.0;       // the zero slots padding entries point at
    double acc = 0.0;
    int stage = 0, pstage = 0, buf = 0;
    unsigned parity = 0;
    // prologue: nu_local of the first tile
    if ((long long)blockIdx.x < A.n_tiles) {
        mbar_wait(&full[0], 0);
        if (MODE != 2) {
            const int ntok = __ldg(A.desc + blockIdx.x).x;   
            for (int t = tid; t < ntok; t += THREADS) nul0[t] = __ldg(A.vec + stages[0].tok[t]);
        }
    }
    __syncthreads();
    // One barrier per tile.  Iteration k: pool phase of tile k -> f[k&1]; prefetch nu_local of tile k+1;
    // barrier; row phase of tile k.  The row phase of tile k overlaps the pool phase of tile k+1 in other
    // warps, and its red.adds are never followed directly by a barrier.
    bool first = true;
    for (long long tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x) {
        const int nrow = __ldg(A.desc + tile).y;
        St& S = stages[stage];
        const double* nul = nul0 + buf * P;
        double* f = f0buf + buf * FS;
        // ---- pool phase
#pragma unroll
        for (int l = tid; l < P; l += THREADS) {
            const uint32_t li = S.lid[l];
            double f0, f1;
            if (MODE == 0) {
                if (A.dbg & 2) { f0 = S.a[0][l]; f1 = S.a[1][l] + S.a[2][l] + nul[li & 0xffffu] + nul[li >> 16]; }
                else
                EvalOp::apply<TRADES, HESS>(A, tile * P + l, S.a[0][l], S.a[1][l], S.a[2][l], nul[li & 0xffffu],
                                            nul[li >> 16], f0, f1, acc);
            } else if (MODE == 1) {
                f0 = S.a[0][l] * (nul[li & 0xffffu] - nul[li >> 16]);
                f1 = -f0;
            } else {
                f0 = S.a[0][l];
                f1 = f0;
            }
            reinterpret_cast<double2*>(f)[l] = make_double2(f0, f1);
        }
        // ---- nu_local of the next tile into the other buffer
        const long long nxt = tile + gridDim.x;
        int nstage = stage + 1;
        unsigned nparity = parity;
        if (nstage == STAGES) { nstage = 0; nparity ^= 1u; }
        if (nxt < A.n_tiles) {
            mbar_wait(&full[nstage], nparity);
            if (MODE != 2 && !(A.dbg & 4)) {
                const int ntok = __ldg(A.desc + nxt).x;
                double* nn = nul0 + (buf ^ 1) * P;
                for (int t = tid; t < ntok; t += THREADS) nn[t] = __ldg(A.vec + stages[nstage].tok[t]);
            }
        }
        __syncthreads();          // f of this tile complete; everybody is also done with the PREVIOUS tile's rows
        if (tid == 0 && !first) {
            const long long far = tile + (long long)(STAGES - 1) * gridDim.x;     // (previous tile) + STAGES strides
            if (far < A.n_tiles) {
                fence_proxy_async();
                issue_tile<P, NF>(&stages[pstage], &full[pstage], A, far);
            }
        }
        // ---- row phase: one thread per row, fixed summation order, one red.add per row
        for (int r = tid; r < ((A.dbg & 1) ? 0 : nrow); r += THREADS) {
            const uint32_t rw = S.rows[r];
            const uint2* e4 = reinterpret_cast<const uint2*>(S.ent) + row_start4(rw);
            const int ng = row_groups(rw);
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int g = 0;
            for (; g + 2 <= ng; g += 2) {          // 8 independent shared-memory loads in flight
                const uint2 a = e4[g], b = e4[g + 1];
                const double v0 = f[a.x & 0xffffu], v1 = f[a.x >> 16], v2 = f[a.y & 0xffffu], v3 = f[a.y >> 16];
                const double v4 = f[b.x & 0xffffu], v5 = f[b.x >> 16], v6 = f[b.y & 0xffffu], v7 = f[b.y >> 16];
                s0 += v0; s1 += v1; s2 += v2; s3 += v3;
                s0 += v4; s1 += v5; s2 += v6; s3 += v7;
            }
            if (g < ng) {
                const uint2 a = e4[g];
                s0 += f[a.x & 0xffffu]; s1 += f[a.x >> 16]; s2 += f[a.y & 0xffffu]; s3 += f[a.y >> 16];
            }
            A.partial[tile * BlockedCfg<P>::kRowsMax + r] = (s0 + s1) + (s2 + s3);
        }
        first = false;
        pstage = stage; stage = nstage; parity = nparity; buf ^= 1;
    }
    (void)acc; (void)part;
}

// Second pass: psi[token] = sum of that token's row sums, in a fixed order (bit-reproducible), no atomics for
// tokens that fit one segment.  One warp per segment (<= kSegCap rows); arb += nu[token] * sum.
constexpr int kSegCap = 1024;
template <bool WITH_ARB>
__global__ void __launch_bounds__(256)
k_token_reduce(int n_seg, const int4* __restrict__ seg, const int* __restrict__ pos, const double* __restrict__ partial,
               const double* __restrict__ nu, double* out, double* arb) {
    __shared__ double part[8];
    const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    double contrib = 0.0;
    if (warp < n_seg) {
        const int4 sg = __ldg(seg + warp);              // (token, begin, end, multi-segment flag)
        double s = 0.0;
        for (int k = sg.y + lane; k < sg.z; k += 32) s += partial[__ldg(pos + k)];
        s = warp_sum(s);
        if (lane == 0) {
            if (sg.w) atomicAdd(out + sg.x, s); else out[sg.x] = s;
            if (WITH_ARB) contrib = __ldg(nu + sg.x) * s;
        }
    }
    if (WITH_ARB) {
        contrib = warp_sum(contrib);
        if (lane == 0) part[threadIdx.x >> 5] = contrib;
        __syncthreads();
        if (threadIdx.x < 32) {
            double t = (threadIdx.x < 8) ? part[threadIdx.x] : 0.0;
            t = warp_sum(t);
            if (threadIdx.x == 0 && t != 0.0) atomicAdd(arb, t);
        }
    }
}

// ---- shipped configurations (selectable at run time for tuning; the layout must be built for the same P)
struct Cfg0 { static constexpr int P = 1024, T = 512, S = 3, CTAS = 1; };   // 121 + 48 KB smem, one CTA per SM
struct Cfg1 { static constexpr int P = 512, T = 512, S = 4, CTAS = 2; };    // 2 x (81 + 24) KB
struct Cfg2 { static constexpr int P = 512, T = 256, S = 3, CTAS = 2; };    // 2 x (61 + 24) KB, 2 pools per thread
int g_cfg = 1;
int g_dbg = 0;

template <class C, int MODE, bool TRADES, bool HESS>
int launch_cfg(const BlockedArgs& A, cudaStream_t st) {
    constexpr int NF = (MODE == 0) ? 3 : 1;
    auto kern = k_blocked<C::P, C::T, C::S, MODE, TRADES, HESS>;
    const size_t sm = (size_t)C::S * sizeof(Stage<C::P, NF>) + (size_t)(6 * C::P + 8) * sizeof(double);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        attr = true;
    }
    const long long cap = (long long)C::CTAS * num_sms();
    const int grid = (int)(A.n_tiles < cap ? A.n_tiles : cap);
    kern<<<grid, C::T, sm, st>>>(A);
    int rc = check_launch();
    if (rc) return rc;
    const int blocks = (A.n_seg * 32 + 255) / 256;
    if (MODE == 0)
        k_token_reduce<true><<<blocks, 256, 0, st>>>(A.n_seg, A.seg, A.pos, A.partial, A.vec, A.out, A.arb);
    else
        k_token_reduce<false><<<blocks, 256, 0, st>>>(A.n_seg, A.seg, A.pos, A.partial, nullptr, A.out, nullptr);
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
    if (b->n_tiles > 0 && (!b->lid || !b->ent || !b->rows || !b->tok || !b->desc || !b->partial || !b->seg ||
                           !b->pos)) return CFMM_E_NULL;
    if (b->n_seg < 0) return CFMM_E_SIZE;
    A.n_tiles = b->n_tiles;
    A.M = b->n_tiles * P;
    A.lid = b->lid; A.ent = b->ent; A.rows = b->rows; A.tok = b->tok;
    A.desc = reinterpret_cast<const int4*>(b->desc);
    A.partial = b->partial; A.n_seg = (int)b->n_seg; A.seg = reinterpret_cast<const int4*>(b->seg); A.pos = b->pos;
    A.n_out = 0;
    A.slab[0] = A.slab[1] = A.slab[2] = nullptr;
    A.vec = nullptr; A.out = nullptr; A.arb = nullptr; A.delta = A.lambda = A.hcoef = nullptr;
    A.dbg = g_dbg;
    return CFMM_OK;
}

}  // namespace

extern "C" {

int cfmm_blocked_layout_info(int32_t* pools_per_tile, int32_t* rows_stride, int32_t* tok_stride, int32_t* row_cap,
                             int32_t* ent_stride) {
    const int P = cfg_P();
    const int rows = P + 2 * P / 32 + 8;
    if (pools_per_tile) *pools_per_tile = P;
    if (rows_stride) *rows_stride = rows;
    if (tok_stride) *tok_stride = P;
    if (row_cap) *row_cap = 32;
    if (ent_stride) *ent_stride = (2 * P + 3 * rows + 7) / 8 * 8;
    return CFMM_OK;
}

int cfmm_set_blocked_config(int32_t cfg) {
    if (cfg >= 100) { g_dbg = cfg - 100; return CFMM_OK; }      // measurement-only phase switches
    if (cfg < 0 || cfg > 2) return CFMM_E_KIND;
    g_cfg = cfg;
    return CFMM_OK;
}

int cfmm_blocked_eval(const cfmm_blocked_pairs* b, int32_t n_tokens, const double* nu, double* psi, double* arb,
                      const cfmm_eval_out* out, void* stream) {
    BlockedArgs A;
    int rc = fill_args(b, A);
    if (rc) return rc;
    if (n_tokens <= 0) return CFMM_E_SIZE;
    if (!nu || !psi || !arb) return CFMM_E_NULL;
    if (b->n_tiles == 0) return CFMM_OK;
    if (!b->r0 || !b->r1 || !b->gamma_inv) return CFMM_E_NULL;
    A.slab[0] = b->r0; A.slab[1] = b->r1; A.slab[2] = b->gamma_inv;
    A.vec = nu; A.out = psi; A.arb = arb;
    A.n_out = (arb == psi + n_tokens) ? n_tokens + 1 : n_tokens;
    if (A.n_out == n_tokens) cudaMemsetAsync(arb, 0, sizeof(double), static_cast<cudaStream_t>(stream));
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
                     void* stream) {
    BlockedArgs A;
    int rc = fill_args(b, A);
    if (rc) return rc;
    if (n_tokens <= 0) return CFMM_E_SIZE;
    if (!hcoef || !vt || !y) return CFMM_E_NULL;
    if (b->n_tiles == 0) return CFMM_OK;
    A.slab[0] = hcoef; A.vec = vt; A.out = y; A.n_out = n_tokens;
    return launch_blocked<1, false, false>(A, static_cast<cudaStream_t>(stream));
}

int cfmm_blocked_diag(const cfmm_blocked_pairs* b, int32_t n_tokens, const double* hcoef, double* diag, void* stream) {
    BlockedArgs A;
    int rc = fill_args(b, A);
    if (rc) return rc;
    if (n_tokens <= 0) return CFMM_E_SIZE;
    if (!hcoef || !diag) return CFMM_E_NULL;
    if (b->n_tiles == 0) return CFMM_OK;
    A.slab[0] = hcoef; A.out = diag; A.n_out = n_tokens;
    return launch_blocked<2, false, false>(A, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
