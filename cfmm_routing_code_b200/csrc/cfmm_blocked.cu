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

// fused LL all-reduce (see csrc/cfmm_allreduce.cu for the protocol): the CTA that finishes last pushes the finished
// vector to the peers' receive areas and sums what the peers pushed
struct PeerLL {
    LLCell* const* recv;          // device array [world] of receive areas, [3 slots][world sources][stride] cells
    unsigned int* done_ctr;       // zero-initialised; counts finished CTAs of one launch
    double* red;                  // [n] all-reduced result
    long long slot_off, stride;   // in cells
    unsigned long long seq;
    int rank, world, n;
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
    PeerLL peer;                  // optional fused all-reduce of `out` (world > 1): done by the last CTA to finish
    int tile_pools;               // pools per tile of this layout (host-side dispatch only); 0 = per-tile sizes in desc (VAR)
};

__device__ __forceinline__ unsigned round16(unsigned bytes) { return (bytes + 15u) & ~15u; }

// VAR: tiles of ANY size <= P.  The descriptor of a tile then carries (ntok, nrow, pools in the tile, offset of its first pool
// in the slab arrays); offsets are multiples of 4 pools so every bulk copy stays 16-byte aligned, and P is only the capacity
// of the shared-memory stage.  Lets the builder cut the pool list into a whole number of tiles per resident CTA, with a small
// first tile so a CTA's compute starts before the bulk of its data has landed (pools.py: plan_tiles).
template <int P, bool VAR>
__device__ __forceinline__ int tile_count(const int4 d) { return VAR ? d.z : P; }
template <int P, bool VAR>
__device__ __forceinline__ long long tile_offset(const int4 d, long long tile) { return VAR ? (long long)d.w : tile * P; }

template <int P, int NF, bool VAR = false>
__device__ __forceinline__ void issue_tile(Stage<P, NF>* st, uint64_t* bar, const BlockedArgs& A, long long tile,
                                           const int4 d) {
    const unsigned rows_b = round16(4u * (unsigned)d.y);
    const unsigned tok_b = round16(4u * (unsigned)d.x);
    const int tp = VAR ? ((d.z + 3) & ~3) : P;                 // pools copied (the slabs are padded to a multiple of 4)
    const long long off = tile_offset<P, VAR>(d, tile);
    mbar_expect_tx(bar, (unsigned)(NF * tp * 8 + tp * 4 + tp * 4 + 16) + rows_b + tok_b);
    bulk_g2s(&st->desc, A.desc + tile, 16, bar);
#pragma unroll
    for (int k = 0; k < NF; ++k) bulk_g2s(st->a[k], A.slab[k] + off, tp * 8, bar);
    bulk_g2s(st->lid, A.lid + off, tp * 4, bar);
    bulk_g2s(st->pos, A.pos + off, tp * 4, bar);
    bulk_g2s(st->rows, A.rows + tile * BlockedCfg<P>::kRowsMax, rows_b, bar);
    bulk_g2s(st->tok, A.tok + tile * BlockedCfg<P>::kTokMax, tok_b, bar);
}

// ---- per-pool operator: the two net flows (f0, f1) of a constant-product pool (arbitrage.py:68-70) ----------
// With gi = 1/gamma, p_j = nu_j R_j and v = rsqrt(p0 p1 gi):  a = p0 v, b = p1 v  (a b gi = 1).  The KKT solution is
//   b > 1 : tender token 0:  x0 = R0 b,       x1 = R1 a gi   =>  f0 = -R0 (b-1) gi,  f1 = R1 (1 - a gi)
//   a > 1 : tender token 1:  x1 = R1 a,       x0 = R0 b gi   =>  f1 = -R1 (a-1) gi,  f0 = R0 (1 - b gi)
//   else  : no-trade cone (then b gi >= 1 and a gi >= 1, so the "receive" expressions clamp to 0 by themselves).
// Written with selects only (no divergent branch on the direction).
// h = sqrt(p0 p1 / gamma)/2 = w v / 2 on trading pools (Hs_i = h [[1,-1],[-1,1]] in log-price coordinates).
// max(x, 0) on the bit pattern: a negative double has its sign bit set, so masking with ~(hi >> 31) zeroes it (3 integer
// instructions instead of the NaN-propagating fp64 max sequence)
__device__ __forceinline__ double clamp0(double x) {
    const int hi = __double2hiint(x);
    const int m = ~(hi >> 31);
    return __hiloint2double(hi & m, __double2loint(x) & m);
}

struct EvalOp {
    template <bool TRADES, bool HESS>
    __device__ __forceinline__ static void apply(const BlockedArgs& A, long long q, double R0, double R1, double gi,
                                                 double n0, double n1, double& f0, double& f1, double& acc) {
        const double p0 = n0 * R0, p1 = n1 * R1;
        const double w = p0 * p1 * gi;
        const double v = rsqrt(w);
        const double a = p0 * v, b = p1 * v;
        const double ob = 1.0 - b, oa = 1.0 - a;
        const double r0 = fma(-b, gi, 1.0), r1 = fma(-a, gi, 1.0);          // 1 - b gi, 1 - a gi  (received share)
        const double x0 = (ob < 0.0) ? ob * gi : clamp0(r0);
        const double x1 = (oa < 0.0) ? oa * gi : clamp0(r1);
        f0 = R0 * x0;
        f1 = R1 * x1;
        acc = fma(n0, f0, fma(n1, f1, acc));
        if (TRADES) {
            A.delta[q] = f0 < 0.0 ? -f0 : 0.0; A.delta[A.M + q] = f1 < 0.0 ? -f1 : 0.0;
            A.lambda[q] = f0 > 0.0 ? f0 : 0.0; A.lambda[A.M + q] = f1 > 0.0 ? f1 : 0.0;
        }
        if (HESS) A.hcoef[q] = (x0 != 0.0 || x1 != 0.0) ? 0.5 * w * v : 0.0;
    }
};

// Called by every thread at the end of a blocked kernel.  The last CTA to arrive owns the finished `out` vector of this
// rank (all red.adds of the launch are ordered before the counter bump by the fences) and runs the LL all-reduce.
template <int THREADS>
__device__ __forceinline__ void fused_allreduce_tail(const BlockedArgs& A) {
    if (A.peer.world <= 1) return;
    __shared__ int s_last;
    const int tid = threadIdx.x;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned t = atomicAdd(A.peer.done_ctr, 1u);
        s_last = (t == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    if (tid == 0) *A.peer.done_ctr = 0u;                  // next launch starts from zero (it cannot get here before we exit)
    __threadfence();
    const int rank = A.peer.rank, world = A.peer.world, n = A.peer.n;
    for (int j = tid; j < n; j += THREADS) {
        const double mine = __ldcg(A.out + j);
        for (int r = 0; r < world; ++r)
            if (r != rank) st_ll(A.peer.recv[r] + A.peer.slot_off + (long long)rank * A.peer.stride + j, mine, A.peer.seq);
    }
    for (int j = tid; j < n; j += THREADS) {
        double s = 0.0;
        for (int r = 0; r < world; ++r) {               // rank order: same bits on every rank
            double v;
            if (r == rank) {
                v = __ldcg(A.out + j);
            } else {
                const LLCell* c = A.peer.recv[rank] + A.peer.slot_off + (long long)r * A.peer.stride + j;
                unsigned long long f;
                do { ld_ll(c, v, f); } while (f != A.peer.seq);
            }
            s += v;
        }
        A.peer.red[j] = s;
    }
}

// row word: start (16 bits) | length (6 bits, 1..32) | local token (10 bits).  Rows of a tile are sorted by
// decreasing length by the builder, so the 32 rows of a warp have (nearly) equal trip counts.
__device__ __forceinline__ int row_start(uint32_t r) { return (int)(r & 0xffffu); }
__device__ __forceinline__ int row_len(uint32_t r) { return (int)((r >> 16) & 0x3fu); }
__device__ __forceinline__ int row_tok(uint32_t r) { return (int)(r >> 22); }

template <int P, int THREADS, int STAGES, int MODE /*0 eval, 1 hvp, 2 diag*/, bool TRADES, bool HESS, bool AR = false,
          bool VAR = false>
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
            if (t < t_end) issue_tile<P, NF, VAR>(&stages[s], &full[s], A, t, __ldg(A.desc + t));
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
        const int4 d = S.desc;                          // (ntok, nrow, pools, first pool) -- the last two only read by VAR
        const int tp = tile_count<P, VAR>(d);
        const long long toff = tile_offset<P, VAR>(d, tile);
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
                if (VAR && l >= tp) {                   // lane beyond the tile: zero flows into the scratch slots past 2 tp
                    f0[u] = f1[u] = 0.0;
                    ps[u] = (uint32_t)(2 * tp) | ((uint32_t)(2 * tp + 1) << 16);
                    continue;
                }
                const uint32_t li = S.lid[l];
                ps[u] = S.pos[l];
                if (MODE == 0) {
                    EvalOp::apply<TRADES, HESS>(A, toff + l, S.a[0][l], S.a[1][l], S.a[2][l], nul[li & 0xffffu],
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
#pragma unroll 1
            for (; k + 4 <= len; k += 4) { s0 += q[k] + q[k + 2]; s1 += q[k + 1] + q[k + 3]; }
            if (k + 2 <= len) { s0 += q[k]; s1 += q[k + 1]; k += 2; }
            if (k < len) s0 += q[k];
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
            issue_tile<P, NF, VAR>(&S, &full[stage], A, far, dfar);
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
    if (AR) fused_allreduce_tail<THREADS>(A);      // compiled only into the pool-sharded instantiations
}

// ---------------------------------------------------------------------------------------------------------------
// Variant "regs" (configuration 3): the per-pool slabs (R0, R1, 1/gamma, ids, positions) are read exactly once, so they
// go global -> registers directly (coalesced LDG, prefetched one tile ahead) instead of through shared memory.  Only
// the small per-tile tables (row table, token list, descriptor) ride a 4-deep TMA ring.  That frees ~80 KB of shared
// memory per CTA, which pays for double-buffered nu_local and flows, and those allow ONE barrier per tile: the row
// phase of tile k overlaps the pool phase of tile k+1 in other warps, and nu_local of tile k+1 is fetched from L2
// while tile k's pool phase computes.
// ---------------------------------------------------------------------------------------------------------------
template <int P>
struct __align__(128) TabStage {
    uint32_t rows[BlockedCfg<P>::kRowsMax];
    int32_t tok[BlockedCfg<P>::kTokMax];
    int4 desc;
};

template <int P>
__device__ __forceinline__ void issue_tables(TabStage<P>* st, uint64_t* bar, const BlockedArgs& A, long long tile,
                                             const int4 d) {
    const unsigned rows_b = round16(4u * (unsigned)d.y);
    const unsigned tok_b = round16(4u * (unsigned)d.x);
    mbar_expect_tx(bar, 16u + rows_b + tok_b);
    bulk_g2s(&st->desc, A.desc + tile, 16, bar);
    bulk_g2s(st->rows, A.rows + tile * BlockedCfg<P>::kRowsMax, rows_b, bar);
    bulk_g2s(st->tok, A.tok + tile * BlockedCfg<P>::kTokMax, tok_b, bar);
}

template <int NF>
struct PoolRegs {
    double a[NF];
    uint32_t lid, pos;
};

template <int P, int THREADS, int NF, int NPOOL, bool VAR = false>
__device__ __forceinline__ void load_pools(PoolRegs<NF> (&r)[NPOOL], const BlockedArgs& A, long long tile, int tid) {
    int4 d = make_int4(0, 0, 0, 0);
    if (VAR) d = __ldg(A.desc + tile);                  // the ring stage of this tile may not have landed yet: read it directly
    const int tp = tile_count<P, VAR>(d);
    const long long toff = tile_offset<P, VAR>(d, tile);
#pragma unroll
    for (int u = 0; u < NPOOL; ++u) {
        if (VAR && tid + u * THREADS >= tp) {           // lane beyond the tile: inert entry, flows go to the scratch slots
#pragma unroll
            for (int k = 0; k < NF; ++k) r[u].a[k] = 0.0;
            r[u].lid = 0u;
            r[u].pos = (uint32_t)(2 * tp) | ((uint32_t)(2 * tp + 1) << 16);
            continue;
        }
        const long long q = toff + tid + u * THREADS;
#pragma unroll
        for (int k = 0; k < NF; ++k) r[u].a[k] = __ldg(A.slab[k] + q);
        r[u].lid = __ldg(A.lid + q);
        r[u].pos = __ldg(A.pos + q);
    }
}

// pull the slabs of `tile` from HBM into L2 ahead of the register loads (one thread, 5 bulk prefetches)
template <int P, int NF, bool VAR = false>
__device__ __forceinline__ void prefetch_pools_l2(const BlockedArgs& A, long long tile) {
    int4 d = make_int4(0, 0, 0, 0);
    if (VAR) d = __ldg(A.desc + tile);
    const int tp = VAR ? ((d.z + 3) & ~3) : P;
    const long long toff = tile_offset<P, VAR>(d, tile);
#pragma unroll
    for (int k = 0; k < NF; ++k) bulk_prefetch_l2(A.slab[k] + toff, tp * 8);
    bulk_prefetch_l2(A.lid + toff, tp * 4);
    bulk_prefetch_l2(A.pos + toff, tp * 4);
}

template <int P, int THREADS, int STAGES, int MODE, bool TRADES, bool HESS, bool AR = false, bool VAR = false>
__global__ void __launch_bounds__(THREADS, 2)
k_blocked_regs(const BlockedArgs A) {
    constexpr int NF = (MODE == 0) ? 3 : 1;
    constexpr int NPOOL = P / THREADS;
    constexpr int NPRE = (P + THREADS - 1) / THREADS;
    using St = TabStage<P>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    St* stages = reinterpret_cast<St*>(smem_raw);
    double* nul0 = reinterpret_cast<double*>(smem_raw + (size_t)STAGES * sizeof(St));      // [2][P]  nu_local
    double* g0 = nul0 + 2 * P;                                                              // [2][2P] flows, row order
    __shared__ uint64_t full[STAGES];
    __shared__ double part[THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long t_beg = (A.n_tiles * (long long)blockIdx.x) / gridDim.x;
    const long long t_end = (A.n_tiles * (long long)(blockIdx.x + 1)) / gridDim.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            const long long t = t_beg + s;
            if (t < t_end) issue_tables<P>(&stages[s], &full[s], A, t, __ldg(A.desc + t));
        }
        if (t_beg + 1 < t_end) prefetch_pools_l2<P, NF, VAR>(A, t_beg + 1);
        if (t_beg + 2 < t_end) prefetch_pools_l2<P, NF, VAR>(A, t_beg + 2);
    }
    PoolRegs<NF> cur[NPOOL], nxt[NPOOL];
    if (t_beg < t_end) load_pools<P, THREADS, NF, NPOOL, VAR>(cur, A, t_beg, tid);      // constant tables: before the PDL wait
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    for (int j = blockIdx.x * THREADS + tid; j < A.n_zero; j += gridDim.x * THREADS) A.zero_next[j] = 0.0;
    double acc = 0.0;
    if (t_beg < t_end) {
        mbar_wait(&full[0], 0);
        if (MODE != 2) {
            const int ntok = stages[0].desc.x;
            for (int t = tid; t < ntok; t += THREADS) nul0[t] = __ldg(A.vec + stages[0].tok[t]);
        }
    }
    __syncthreads();
    int stage = 0, pstage = 0, buf = 0;
    unsigned parity = 0;
    for (long long tile = t_beg; tile < t_end; ++tile) {
        St& S = stages[stage];
        const double* nul = nul0 + buf * P;
        double* g = g0 + buf * 2 * P;
        const long long nx = tile + 1;
        int nstage = stage + 1;
        unsigned nparity = parity;
        if (nstage == STAGES) { nstage = 0; nparity ^= 1u; }
        // 1. pool slabs of the next tile -> registers; nu_local of the next tile -> registers (both land while we compute)
        double pre[NPRE];
        int ntok_n = 0;
        if (tid == 0 && tile + 3 < t_end) prefetch_pools_l2<P, NF, VAR>(A, tile + 3);      // HBM -> L2, 3 tiles ahead
        if (nx < t_end) {
            load_pools<P, THREADS, NF, NPOOL, VAR>(nxt, A, nx, tid);
            mbar_wait(&full[nstage], nparity);
            if (MODE != 2) {
                ntok_n = stages[nstage].desc.x;
#pragma unroll
                for (int k = 0; k < NPRE; ++k) {
                    const int t = tid + k * THREADS;
                    pre[k] = (t < ntok_n) ? __ldg(A.vec + stages[nstage].tok[t]) : 0.0;
                }
            }
        }
        // 2. pool phase of this tile (registers + nu_local) -> flows in row order
        {
            double f0[NPOOL], f1[NPOOL];
#pragma unroll
            for (int u = 0; u < NPOOL; ++u) {
                const uint32_t li = cur[u].lid;
                if (VAR && MODE == 0 && tid + u * THREADS >= S.desc.z) {  // inert lane: no pool behind it
                    f0[u] = f1[u] = 0.0;
                } else if (MODE == 0) {
                    EvalOp::apply<TRADES, HESS>(A, tile_offset<P, VAR>(S.desc, tile) + tid + u * THREADS, cur[u].a[0], cur[u].a[NF > 1 ? 1 : 0],
                                                cur[u].a[NF > 2 ? 2 : 0], nul[li & 0xffffu], nul[li >> 16], f0[u], f1[u],
                                                acc);
                } else if (MODE == 1) {
                    f0[u] = cur[u].a[0] * (nul[li & 0xffffu] - nul[li >> 16]);
                    f1[u] = -f0[u];
                } else {
                    f0[u] = cur[u].a[0];
                    f1[u] = f0[u];
                }
            }
#pragma unroll
            for (int u = 0; u < NPOOL; ++u) {
                g[cur[u].pos & 0xffffu] = f0[u];
                g[cur[u].pos >> 16] = f1[u];
            }
        }
        // 3. nu_local of the next tile into the other buffer (read last in the pool phase of tile-1: before barrier-1)
        if (MODE != 2 && nx < t_end) {
            double* nn = nul0 + (buf ^ 1) * P;
#pragma unroll
            for (int k = 0; k < NPRE; ++k) {
                const int t = tid + k * THREADS;
                if (t < ntok_n) nn[t] = pre[k];
            }
        }
        __syncthreads();      // flows of this tile and nu_local of the next are complete; everybody left tile-1's rows
        if (tid == 0 && tile > t_beg) {
            const long long far = tile - 1 + STAGES;
            if (far < t_end) {
                fence_proxy_async();
                issue_tables<P>(&stages[pstage], &full[pstage], A, far, __ldg(A.desc + far));
            }
        }
        // 4. row phase: overlaps the next tile's steps 1-3 in the warps that get there first
        const int nrow = S.desc.y;
        for (int r = tid; r < nrow; r += THREADS) {
            const uint32_t rw = S.rows[r];
            const double* q = g + row_start(rw);
            const int len = row_len(rw);
            double s0 = 0.0, s1 = 0.0;
            int k = 0;
#pragma unroll 1
            for (; k + 4 <= len; k += 4) { s0 += q[k] + q[k + 2]; s1 += q[k + 1] + q[k + 3]; }
            if (k + 2 <= len) { s0 += q[k]; s1 += q[k + 1]; k += 2; }
            if (k < len) s0 += q[k];
            const double s = s0 + s1;
            if (s != 0.0) atomicAdd(A.out + S.tok[row_tok(rw)], s);
        }
#pragma unroll
        for (int u = 0; u < NPOOL; ++u) cur[u] = nxt[u];
        pstage = stage; stage = nstage; parity = nparity; buf ^= 1;
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
    if (AR) fused_allreduce_tail<THREADS>(A);      // compiled only into the pool-sharded instantiations
}

// ---- configuration of the TMA-staged variant (the register-fed variant uses the same tile size): two pools per thread,
// two ring stages, two CTAs per SM.  P = 1024: 2 x (85 + 24) KB of shared memory per SM.  The smaller tile sizes exist for
// load balance: a launch walks ceil(n_tiles / (2 SMs)) tiles on its critical path, e.g. 1M pools = 977 tiles of 1024 over
// 296 CTAs -> 4 x 1024 = 4096 pools, but 1117 tiles of 896 -> 4 x 896 = 3584 (cfmm_set_blocked_config(400 + P)).
template <int P_>
struct CfgP { static constexpr int P = P_, T = P_ / 2, S = 2, CTAS = 2; };
using Cfg0 = CfgP<1024>;
int g_tile_pools = 1024;      // what cfmm_blocked_layout_info tells the layout builder
int g_cfg = -1;
int g_pdl = 1;
int g_row_cap = 32;

template <class C, int MODE, bool TRADES, bool HESS, bool VAR = false>
int launch_cfg(const BlockedArgs& A, cudaStream_t st) {
    constexpr int NF = (MODE == 0) ? 3 : 1;
    const bool ar = A.peer.world > 1;
    auto kern = ar ? k_blocked<C::P, C::T, C::S, MODE, TRADES, HESS, true, VAR>
                   : k_blocked<C::P, C::T, C::S, MODE, TRADES, HESS, false, VAR>;
    const size_t sm = (size_t)C::S * sizeof(Stage<C::P, NF>) + (size_t)(3 * C::P) * sizeof(double);
    static bool attr[2] = {false, false};
    if (!attr[ar]) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        attr[ar] = true;
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

template <int P, int MODE, bool TRADES, bool HESS, bool VAR = false>
int launch_regs(const BlockedArgs& A, cudaStream_t st) {
    constexpr int T = P / 2, S = 4;
    const bool ar = A.peer.world > 1;
    auto kern = ar ? k_blocked_regs<P, T, S, MODE, TRADES, HESS, true, VAR>
                   : k_blocked_regs<P, T, S, MODE, TRADES, HESS, false, VAR>;
    const size_t sm = (size_t)S * sizeof(TabStage<P>) + (size_t)(2 * P + 4 * P) * sizeof(double);
    static bool attr[2] = {false, false};
    if (!attr[ar]) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        attr[ar] = true;
    }
    const long long cap = 2LL * num_sms();
    const int grid = (int)(A.n_tiles < cap ? A.n_tiles : cap);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(T); cfg.dynamicSmemBytes = sm; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = g_pdl;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, A);
    return check_launch();
}

template <int P, int MODE, bool TRADES, bool HESS, bool VAR = false>
int launch_blocked_p(const BlockedArgs& A, cudaStream_t st) {
    // default (-1): evaluation through the TMA-staged slabs, Hessian products / diagonal (1 slab, less data per tile)
    // through the register-fed single-barrier variant -- each is the faster one for its mode (profiles/r1f_*)
    if (g_cfg == 3 || (g_cfg < 0 && MODE != 0)) return launch_regs<P, MODE, TRADES, HESS, VAR>(A, st);
    return launch_cfg<CfgP<P>, MODE, TRADES, HESS, VAR>(A, st);
}

// compile-time tile sizes, or 0 = every tile carries its own size and offset in its descriptor (VAR instantiations of
// P = 1024, whose stage capacity bounds the tile size)
bool tile_pools_fixed(long long P) { return P == 1024 || P == 960 || P == 896; }
bool tile_pools_ok(long long P) { return tile_pools_fixed(P) || P == 0; }

template <int MODE, bool TRADES, bool HESS>
int launch_blocked(const BlockedArgs& A, cudaStream_t st) {
    switch (A.tile_pools) {                                 // pools per tile of THIS layout (validated by fill_args)
        case 1024: return launch_blocked_p<1024, MODE, TRADES, HESS>(A, st);
        case 960: return launch_blocked_p<960, MODE, TRADES, HESS>(A, st);
        case 896: return launch_blocked_p<896, MODE, TRADES, HESS>(A, st);
        default: return launch_blocked_p<1024, MODE, TRADES, HESS, true>(A, st);
    }
}

int fill_args(const cfmm_blocked_pairs* b, BlockedArgs& A) {
    if (!b) return CFMM_E_NULL;
    if (!tile_pools_ok(b->pools_per_tile)) return CFMM_E_KIND;
    const int64_t P = b->pools_per_tile;
    if (b->n_tiles < 0 || b->n_pools < 0 || b->n_pools > b->n_tiles * (P ? P : 1024)) return CFMM_E_SIZE;
    if (b->n_tiles > 0 && (!b->lid || !b->pos || !b->rows || !b->tok || !b->desc)) return CFMM_E_NULL;
    A.n_tiles = b->n_tiles;
    A.tile_pools = (int)P;
    A.M = P ? b->n_tiles * P : ((b->n_pools + 3) & ~(int64_t)3);      // slab stride (= where slot 1 of delta / lambda starts)
    A.lid = b->lid; A.pos = b->pos; A.rows = b->rows; A.tok = b->tok;
    A.desc = reinterpret_cast<const int4*>(b->desc);
    A.zero_next = nullptr; A.n_zero = 0;
    A.slab[0] = A.slab[1] = A.slab[2] = nullptr;
    A.vec = nullptr; A.out = nullptr; A.arb = nullptr; A.delta = A.lambda = A.hcoef = nullptr;
    A.peer = PeerLL{};
    return CFMM_OK;
}

}  // namespace

extern "C" {

int cfmm_blocked_layout_info(int32_t* pools_per_tile, int32_t* rows_stride, int32_t* tok_stride, int32_t* row_cap,
                             int32_t* ent_stride) {
    // 0 = "balanced": the builder picks the tile size per bucket (any multiple of 4 in [256, 1024]; runtime-sized
    // kernels, whose row / token tables have the strides of the 1024 layout)
    const int P = g_tile_pools;
    const int cap = P ? P : 1024;
    const int rows = cap + 2 * cap / 8 + 8;
    if (pools_per_tile) *pools_per_tile = P;
    if (rows_stride) *rows_stride = rows;
    if (tok_stride) *tok_stride = cap;
    if (row_cap) *row_cap = g_row_cap;
    if (ent_stride) *ent_stride = 0;          /* unused since the row-ordered scatter (kept for ABI stability) */
    return CFMM_OK;
}

int cfmm_set_blocked_config(int32_t cfg) {
    if (cfg >= 400) {                                           // 400 + P: pools per tile of layouts built from now on
        if (cfg == 400) { g_tile_pools = 0; return CFMM_OK; }   // 400: balanced (the builder picks per bucket)
        if (!tile_pools_fixed(cfg - 400)) return CFMM_E_KIND;
        g_tile_pools = cfg - 400;
        return CFMM_OK;
    }
    if (cfg >= 300) { const int c = cfg - 300; if (c < 8 || c > 32) return CFMM_E_KIND; g_row_cap = c; return CFMM_OK; }
    if (cfg >= 200) { g_pdl = cfg - 200; return CFMM_OK; }      // 200 / 201: programmatic dependent launch off / on
    if (cfg != -1 && cfg != 0 && cfg != 3) return CFMM_E_KIND;
    g_cfg = cfg;
    return CFMM_OK;
}

static int fill_peer(const cfmm_peer_ll* p, int n, PeerLL& P) {
    if (!p) return CFMM_OK;
    if (!p->peer_recv_dev || !p->done_counter || !p->reduced) return CFMM_E_NULL;
    if (p->world < 2 || p->world > 16 || p->rank < 0 || p->rank >= p->world || p->seq == 0) return CFMM_E_SIZE;
    P.recv = static_cast<LLCell* const*>(const_cast<void*>(p->peer_recv_dev));
    P.done_ctr = p->done_counter; P.red = p->reduced;
    P.slot_off = p->slot_off_cells; P.stride = p->src_stride_cells; P.seq = p->seq;
    P.rank = p->rank; P.world = p->world; P.n = n;
    return CFMM_OK;
}

int cfmm_blocked_eval(const cfmm_blocked_pairs* b, int32_t n_tokens, const double* nu, double* psi, double* arb,
                      const cfmm_eval_out* out, double* zero_next, int64_t n_zero, void* stream) {
    return cfmm_blocked_eval_fused(b, n_tokens, nu, psi, arb, out, zero_next, n_zero, nullptr, stream);
}

int cfmm_blocked_eval_fused(const cfmm_blocked_pairs* b, int32_t n_tokens, const double* nu, double* psi, double* arb,
                            const cfmm_eval_out* out, double* zero_next, int64_t n_zero, const cfmm_peer_ll* peer,
                            void* stream) {
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
    if (peer && arb != psi + n_tokens) return CFMM_E_STATE;      // the fused reduce covers [psi | arb] as one vector
    rc = fill_peer(peer, n_tokens + 1, A.peer);
    if (rc) return rc;
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
    return cfmm_blocked_hvp_fused(b, n_tokens, hcoef, vt, y, zero_next, nullptr, stream);
}

int cfmm_blocked_hvp_fused(const cfmm_blocked_pairs* b, int32_t n_tokens, const double* hcoef, const double* vt, double* y,
                           double* zero_next, const cfmm_peer_ll* peer, void* stream) {
    BlockedArgs A;
    int rc = fill_args(b, A);
    if (rc) return rc;
    if (n_tokens <= 0) return CFMM_E_SIZE;
    if (!hcoef || !vt || !y) return CFMM_E_NULL;
    if (b->n_tiles == 0) return CFMM_OK;
    A.slab[0] = hcoef; A.vec = vt; A.out = y;
    A.zero_next = zero_next; A.n_zero = zero_next ? n_tokens : 0;
    rc = fill_peer(peer, n_tokens, A.peer);
    if (rc) return rc;
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
