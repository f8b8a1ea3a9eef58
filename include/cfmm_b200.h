/*
 * cfmm_b200.h -- C ABI of the B200-native CFMM routing hot path (libcfmm_b200.so).
 *
 * The reference (angeris/cfmm-routing-code) has no FFI: its boundary is the cvxpy call site.
 * Each entry point below names the reference lines it replaces.  All pointers are DEVICE
 * pointers unless the name ends in _host; nothing is allocated inside the *_eval/_hvp calls;
 * every call is asynchronous on `stream` (a cudaStream_t passed as void*), capturable in a CUDA
 * graph, and returns 0 or a negative CFMM_E_* code.  No host threads, no CPU fallback.
 *
 * Pool storage ("bucket"): pools of one kind and one arity k, slot-major SoA, n_pools long, slot j of
 * pool i at [j*stride + i] (per-pool outputs delta/lambda use the same stride):
 *   reserves[k][n_pools]  f64   R_i            arbitrage.py:14-20  (reserves)
 *   tok_idx [k][n_pools]  i32   local_indices  arbitrage.py:6-12   (replaces dense A_i, :42-48)
 *   gamma   [n_pools]     f64   fees[i]        arbitrage.py:22-28
 *   weights [k][n_pools]  f64   normalised p/sum(p) of cp.geo_mean(x, p=...)   arbitrage.py:65
 *   logrw   [k][n_pools]  f64   log(R/w), precomputed once (weighted pools only)
 *   theta_bar[2][n_pools] f64   constant-sum fills (multipliers of the kink), updated by the solver
 */
#ifndef CFMM_B200_H
#define CFMM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    CFMM_KIND_PRODUCT = 0, /* sqrt(x1 x2) >= sqrt(R1 R2)                 arbitrage.py:68-70 */
    CFMM_KIND_SUM = 1,     /* sum(x) >= sum(R), x >= 0                   arbitrage.py:73-74 */
    CFMM_KIND_GEOMEAN = 2, /* prod x^w >= prod R^w                       arbitrage.py:65    */
    CFMM_KIND_BOUNDED_PRODUCT = 3 /* sqrt((x1+o1)(x2+o2)) >= sqrt((R1+o1)(R2+o2)), x >= 0: constant product on virtual
                              reserves, one Uniswap-v3 tick range.  Not in the reference (a new atom for the cons list
                              of arbitrage.py:63-74); arity 2, the two offsets passed in `weights`                */
};

enum {
    CFMM_OK = 0,
    CFMM_E_NULL = -1,      /* required pointer is NULL                                  */
    CFMM_E_KIND = -2,      /* unknown kind, or arity not supported for that kind         */
    CFMM_E_SIZE = -3,      /* negative / overflowing size                                */
    CFMM_E_CUDA = -4,      /* CUDA runtime error (see cfmm_last_cuda_error)              */
    CFMM_E_NODEVICE = -5,  /* no sm_100 device                                           */
    CFMM_E_STATE = -6      /* handle used in the wrong state                             */
};

typedef struct cfmm_bucket {
    int32_t kind;          /* CFMM_KIND_*                                                */
    int32_t arity;         /* tokens per pool: 2 for PRODUCT and SUM, 2..32 for GEOMEAN  */
    int64_t n_pools;
    int64_t stride;        /* elements between consecutive slots (>= n_pools); the TMA-staged path needs
                              stride % 1024 == 0 and 16-byte aligned arrays, otherwise the LDG path runs */
    const double* reserves;
    const int32_t* tok_idx;
    const double* gamma;
    const double* weights;   /* GEOMEAN only */
    const double* logrw;     /* GEOMEAN only */
    const double* theta_bar; /* SUM only     */
} cfmm_bucket;

/* Optional per-pool outputs of one evaluation (any pointer may be NULL). */
typedef struct cfmm_eval_out {
    double* delta;   /* [arity][n_pools]  = deltas[i].value    arbitrage.py:51, two-asset.py:97  */
    double* lambda;  /* [arity][n_pools]  = lambdas[i].value   arbitrage.py:52, two-asset.py:97  */
    double* hcoef;   /* [n_pools] curvature coefficient of arb_i in log-price coordinates        */
    uint32_t* hmask; /* [n_pools] GEOMEAN: bit j set iff token j is traded                      */
} cfmm_eval_out;

/*
 * One dual evaluation over one bucket: for every pool solve the optimal-arbitrage subproblem at
 * prices nu (what `prob.solve()` does for all pools at once, arbitrage.py:81-82, restricted to
 * fixed nu), then ACCUMULATE
 *     psi[j]  += sum_i (A_i (Lambda_i - Delta_i))_j       (psi, arbitrage.py:54)
 *     arb[0]  += sum_i nu_i' (Lambda_i - Delta_i)          (the pool part of the dual value)
 * nu, log_nu: [n_tokens] (log_nu = log(nu), read by GEOMEAN buckets only).  eps: ramp width of the
 * constant-sum proximal smoothing (0 = exact bang-bang LP).  psi/arb must be zeroed by the caller
 * before the first bucket (cfmm_zero does it on the stream).
 */
int cfmm_arb_eval(const cfmm_bucket* bucket, int32_t n_tokens, const double* nu, const double* log_nu,
                  double eps, double* psi, double* arb, const cfmm_eval_out* out, void* stream);

/* y[j] += (Hs vt)_j where Hs = sum_i A_i Hs_i A_i' is the dual Hessian in log-price coordinates
 * (true Hessian = diag(1/nu) Hs diag(1/nu)), vt = v / nu.  Uses hcoef/hmask from cfmm_arb_eval. */
int cfmm_hvp(const cfmm_bucket* bucket, int32_t n_tokens, const double* hcoef, const uint32_t* hmask,
             const double* vt, double* y, void* stream);

/* diag[j] += (Hs)_jj  (Jacobi preconditioner). */
int cfmm_hess_diag(const cfmm_bucket* bucket, int32_t n_tokens, const double* hcoef, const uint32_t* hmask,
                   double* diag, void* stream);

/* H[j*n_tokens + k] += (Hs)_jk, dense row-major n_tokens x n_tokens (small n / direct solves). */
int cfmm_hess_dense(const cfmm_bucket* bucket, int32_t n_tokens, const double* hcoef, const uint32_t* hmask,
                    double* H, void* stream);

/*
 * Token-blocked storage for constant-product pools (the HBM-bound kind).  Built once per problem from
 * local_indices (arbitrage.py:6-12) by the layout builder (pools.py: build_blocked_pairs); pools are
 * reordered into tiles of `pools_per_tile` whose tokens fall into two narrow token blocks.  Per pool 28 B of
 * slabs + 4 B of local ids + 4 B of row positions; per tile a token list, a row table and (ntok, nrow).  See csrc/cfmm_blocked.cu.
 * Strides of the per-tile tables come from cfmm_blocked_layout_info().
 */
typedef struct cfmm_blocked_pairs {
    int64_t n_pools;          /* real pools (<= n_tiles * pools_per_tile; the rest is padding)          */
    int64_t n_tiles;
    int32_t pools_per_tile;   /* must equal the library's tile size                                      */
    int32_t reserved;
    const double* r0;         /* [n_tiles*P] reserves of slot 0, blocked order        arbitrage.py:14-20 */
    const double* r1;         /* [n_tiles*P] reserves of slot 1                                          */
    const double* gamma_inv;  /* [n_tiles*P] 1 / fees[i]                              arbitrage.py:22-28 */
    const uint32_t* lid;      /* [n_tiles*P] tile-local token ids: slot0 | slot1 << 16                   */
    const uint32_t* pos;      /* [n_tiles*P] where the pool's two flows go in the tile's row-ordered array:
                                 pos(slot 0) | pos(slot 1) << 16, each < 2P                              */
    const uint32_t* rows;     /* [n_tiles][rows_stride] start :16 | length 1..32 :6 | local token :10, longest first */
    const int32_t* tok;       /* [n_tiles][tok_stride] local token id -> global token id                 */
    const int32_t* desc;      /* [n_tiles][4] (ntok, nrow, 0, 0)                                          */
} cfmm_blocked_pairs;

int cfmm_blocked_layout_info(int32_t* pools_per_tile, int32_t* rows_stride, int32_t* tok_stride, int32_t* row_cap);
/* tuning knobs for experiments: 200/201 = programmatic dependent launch off/on; 300+c = row cap c (8..32) of layouts
 * built afterwards.  (The tile size is a compile-time constant of the library, 896 pools: the fastest of 1024 / 960 /
 * 896 / per-tile planned sizes measured on B200, profiles/r2a_tile_variants.txt.) */
int cfmm_set_blocked_config(int32_t cfg);

/*
 * Native layout builder (csrc/cfmm_layout.cu): the reference's literals -- local_indices as idx [m][2] int32, reserves
 * [m][2] f64, fees as gamma [m] f64 (arbitrage.py:6-28), contiguous on the device -- become the blocked layout in three
 * launches (pool keys + validation, radix sort, one CTA per tile).  `out`: a cfmm_blocked_pairs with n_pools = m,
 * n_tiles = ceil(m / P), pools_per_tile = P whose array members point at caller-allocated device buffers (strides from
 * cfmm_blocked_layout_info; slabs / lid / pos of n_tiles * P entries); all of them are filled.  order [m] uint32 (out):
 * the pool at every blocked position.  status [4] int32 (device, out): [0] tiles that touch more tokens than a tile may
 * (then the layout is unusable: use a plain bucket), [1] != 0: invalid pools (reserves <= 0 or not finite, fees outside
 * (0, 1], token ids out of range or equal), [2] rows in total.  CFMM_E_SIZE if the sort keys would not fit 32 bits
 * (token_blocks^2 * n_tokens >= 2^32).  work: cfmm_blocked_build_work_bytes(m) bytes.  Asynchronous on `stream`.
 */
int64_t cfmm_blocked_build_work_bytes(int64_t n_pools);
int cfmm_blocked_build(int64_t n_pools, int32_t n_tokens, const int32_t* idx, const double* reserves, const double* gamma,
                       const cfmm_blocked_pairs* out, uint32_t* order, int32_t* status, void* work, int64_t work_bytes,
                       void* stream);

/* Same contract as cfmm_arb_eval for a blocked constant-product bucket: psi/arb ACCUMULATE (one red.add per row
 * of <= 32 entries, ~0.35 per pool, instead of 2 per pool).  Per-pool outputs (delta/lambda [2][n_tiles*P], hcoef
 * [n_tiles*P]) are in BLOCKED order.  If zero_next != NULL the launch also clears zero_next[0..n_zero): callers that
 * ping-pong two [psi | arb] buffers never need a separate memset node between evaluations. */
int cfmm_blocked_eval(const cfmm_blocked_pairs* b, int32_t n_tokens, const double* nu, double* psi, double* arb,
                      const cfmm_eval_out* out, double* zero_next, int64_t n_zero, void* stream);
/* y += Hs vt (optionally clearing zero_next[0..n_tokens) for the next call) and diag += diag(Hs), hcoef in blocked order. */
int cfmm_blocked_hvp(const cfmm_blocked_pairs* b, int32_t n_tokens, const double* hcoef, const double* vt, double* y,
                     double* zero_next, void* stream);
int cfmm_blocked_diag(const cfmm_blocked_pairs* b, int32_t n_tokens, const double* hcoef, double* diag, void* stream);
/* H[j*n_tokens + k] += (Hs)_jk of the blocked bucket, dense row-major (the direct Newton solves of small-n mixed problems) */
int cfmm_blocked_dense(const cfmm_blocked_pairs* b, int32_t n_tokens, const double* hcoef, double* H, void* stream);

/*
 * Native outer loop (csrc/cfmm_solver.cu) for problems whose pools are ONE blocked constant-product bucket: the
 * whole of `prob.solve()` (arbitrage.py:81-82) in one call -- projected Newton-CG on the dual, all vectors on the
 * device, host loop in C++.  Utility in "linear + box" form: maximise c'psi s.t. psi_j + a_j >= 0 (eq[j]=0),
 * == 0 (eq[j]=1), unconstrained with nu_j = c_j (pinned[j]=1)  [arbitrage.py:57,77; liquidation.py:57,77-80;
 * two-asset.py:66,86].  c, a, eq, pinned, nu (in: start, out: solution), psi_out: device, n_tokens long.
 * work: device scratch of cfmm_blocked_solve_work_bytes() bytes.  res: host.  Synchronous on `stream`.
 */
typedef struct cfmm_solve_params {
    double tol;        /* stop when sum_free |nu_j (a_j + psi_j)| / |g| <= tol (bounds gap and infeasibility) */
    double nu_floor;   /* positivity floor for free prices                                                    */
    int32_t max_iter;  /* Newton iterations                                                                    */
    int32_t cg_max;    /* PCG iterations per Newton step                                                       */
} cfmm_solve_params;

typedef struct cfmm_solve_result {
    double dual_value, primal_value, gap, primal_infeas, err;
    int32_t iters, evals, hvps, status;   /* status: 0 optimal, 1 max_iter, 2 stalled */
} cfmm_solve_result;

int64_t cfmm_blocked_solve_work_bytes(const cfmm_blocked_pairs* b, int32_t n_tokens);
int cfmm_blocked_solve(const cfmm_blocked_pairs* b, int32_t n_tokens, const double* c, const double* a,
                       const uint8_t* eq, const uint8_t* pinned, double* nu, double* psi_out, void* work,
                       const cfmm_solve_params* prm, cfmm_solve_result* res, void* stream);

/*
 * The same solve with the pools SHARDED over `world` GPUs (one process per GPU): `b` holds this rank's
 * pools, every rank passes the same c / a / eq / pinned / nu and runs the same loop; each evaluation, Hessian-vector
 * product and Hessian diagonal is followed by the LL all-reduce (cfmm_allreduce_ll) of its (n_tokens+1)- or
 * n_tokens-vector over NVLink peer memory, so every rank sees bit-identical reduced vectors, takes identical decisions
 * and ends with identical nu / psi_out / res.  recv_acc_dev / recv_vec_dev: device arrays of `world` pointers to every
 * rank's receive areas ([3 slots][world][n_tokens+1] and [3 slots][world][n_tokens] cells of 16 B, zeroed once at
 * creation; torch symmetric memory: hdl.buffer_ptrs_dev).  seq_acc / seq_vec: last sequence numbers used on the two
 * channels (in), advanced by the reductions of this call (out) -- equal on all ranks.  peer == NULL: one GPU.
 */
typedef struct cfmm_peer_ctx {
    const void* recv_acc_dev;
    const void* recv_vec_dev;
    int32_t rank, world;
    uint64_t seq_acc, seq_vec;
} cfmm_peer_ctx;

int cfmm_blocked_solve_peer(const cfmm_blocked_pairs* b, int32_t n_tokens, const double* c, const double* a,
                            const uint8_t* eq, const uint8_t* pinned, double* nu, double* psi_out, void* work,
                            const cfmm_solve_params* prm, cfmm_solve_result* res, cfmm_peer_ctx* peer, void* stream);

/*
 * The same solve (one GPU or sharded, same arguments and results) as ONE persistent cooperative kernel
 * (csrc/cfmm_persist.cu): every CTA keeps its chunk of pool tiles for the whole solve and runs all evaluation /
 * Hessian-product / diagonal passes on it; every CTA also owns slices of the n_tokens-long vectors and updates them
 * between passes (two grid barriers per pass), and all CTAs run the same scalar state machine on the same reduced
 * sums, so they agree on the next pass without a broadcast; sharded runs all-reduce inside the slice update (LL pushes
 * over NVLink peer memory, every lane its own token).  The host launches once and reads one result struct -- no host
 * round trip per Newton iteration.  status 3 / CFMM_E_STATE: a peer or CTA never showed up within the spin limit (~3 s)
 * and the kernel gave up.  work: cfmm_persist_solve_work_bytes().
 */
int64_t cfmm_persist_solve_work_bytes(const cfmm_blocked_pairs* b, int32_t n_tokens);
int cfmm_set_persist_cooperative(int32_t on);   /* 1 (default) cooperative launch; 0 plain launch (single-GPU loopback tests) */
int cfmm_persist_last_profile(int64_t* out8);   /* development aid: CTA 0's cycle totals of the last persistent solve */
int cfmm_persist_solve(const cfmm_blocked_pairs* b, int32_t n_tokens, const double* c, const double* a,
                       const uint8_t* eq, const uint8_t* pinned, double* nu, double* psi_out, void* work,
                       const cfmm_solve_params* prm, cfmm_solve_result* res, cfmm_peer_ctx* peer, void* stream);

/*
 * Batches of SMALL problems (the reference's own sizes: 5 pools, 3-5 tokens), one problem per thread, the whole
 * prob.solve() (arbitrage.py:81-82, liquidation.py:84-85, two-asset.py:90-91) of every problem in ONE launch.  Replaces
 * the python loop of two-asset.py:40-100 that builds and solves 50 cvxpy problems in turn.  All problems index the same
 * CSR pool arrays (the literals of arbitrage.py:5-28 flattened); problem p uses the pools [pool_range[2p],
 * pool_range[2p+1]) or, with pool_range == NULL, all of them (a sweep over utilities).  Limits: n_tokens <= 64, weighted
 * arity <= 8, constant-sum arity 2; a problem outside them gets status 3 and NaN results.
 */
typedef struct cfmm_csr_pools {
    int32_t n_tokens;
    int64_t n_pools, nnz;
    const int64_t* pool_ptr;   /* [n_pools+1]                                                        */
    const int32_t* tok_idx;    /* [nnz]  local_indices, arbitrage.py:6-12                             */
    const double* reserves;    /* [nnz]  arbitrage.py:14-20                                          */
    const double* weights;     /* [nnz]  normalised like cp.geo_mean(p=...), arbitrage.py:65; 0 on constant-sum pools */
    const double* logrw;       /* [nnz]  log(reserves / weights) (unused on constant-sum pools)      */
    const double* gamma;       /* [n_pools] fees, arbitrage.py:22-28                                 */
    const uint8_t* kind;       /* [n_pools] CFMM_KIND_SUM | CFMM_KIND_BOUNDED_PRODUCT, else weighted geometric mean */
} cfmm_csr_pools;

typedef struct cfmm_batch {
    int32_t n_problems;
    const int64_t* pool_range; /* nullable [n_problems][2]                                           */
    const double* c;           /* [n_problems][n_tokens] objective on psi (arbitrage.py:31-36,57)    */
    const double* a;           /* [n_problems][n_tokens] endowment (liquidation.py:30-36, two-asset.py:45) */
    const uint8_t* flags;      /* [n_problems][n_tokens] bit0: psi_j + a_j == 0; bit1: psi_j free    */
    double* nu;                /* [n_problems][n_tokens] in: start prices, out: optimal prices       */
    double* psi;               /* [n_problems][n_tokens] out: psi.value (liquidation.py:87)          */
    double* stats;             /* [n_problems][8] out: value (prob.value), dual, gap, infeasibility, kkt err, iters, evals, status */
    double* delta;             /* nullable; out: problem p's Delta at delta[p * trade_stride + csr offset] */
    double* lambda;            /* nullable together with delta                                       */
    int64_t trade_stride;      /* nnz for a shared-pool sweep, 0 for disjoint pool ranges            */
} cfmm_batch;

typedef struct cfmm_batch_params {
    double tol;                /* KKT residual and relative duality gap                              */
    double eps0, eps_min, eps_shrink; /* constant-sum ramp continuation (0.1, 1e-4, 0.5)               */
    double floor_rel;          /* lower bound of free prices relative to max |c| (1e-12)             */
    int32_t max_outer, max_inner;
} cfmm_batch_params;

/* nnz_max: CSR slots of the largest single problem (<= 0: every problem may use all pools->nnz slots) */
int64_t cfmm_batch_solve_work_bytes(const cfmm_csr_pools* pools, int32_t n_problems, int64_t nnz_max);
/* threads per problem: 1 (default; throughput of large batches) or 32 (one warp per problem, the pool loop of every
 * evaluation split over the lanes: latency of small batches / problems with many pools).  Set before sizing the work
 * buffer: it scales with the lane count. */
int cfmm_set_batch_lanes(int32_t lanes);
int cfmm_batch_solve(const cfmm_csr_pools* pools, const cfmm_batch* batch, const cfmm_batch_params* prm, void* work,
                     void* stream);

/*
 * All-reduce (sum) of n doubles over NVLink peer memory, the ONE collective of a pool-sharded dual evaluation (SURVEY
 * 8e), chained into the launch sequence (programmatic dependent launch) right behind the evaluation kernels.
 * Low-latency ("LL") protocol: every rank PUSHES its n doubles into a receive area of every
 * peer as 16-byte {value, seq} cells (flag travels with the data: one NVLink one-way trip, no hand-shake) and sums what
 * it received, in rank order.  peer_recv_dev: device array of `world` pointers to every rank's receive area
 * [3 slots][world sources][src_stride cells of 16 B]; slot_off_cells = (seq % 3) * world * src_stride.  seq >= 1,
 * strictly increasing, equal on all ranks.  The receive areas must start zeroed.
 */
int cfmm_allreduce_ll(const double* local, const void* peer_recv_dev, int32_t rank, int32_t world, int32_t n,
                      int64_t slot_off_cells, int64_t src_stride_cells, double* out, uint64_t seq, void* stream);

/* SUM buckets: theta_bar <- current fills (= lambda), returns max_i |change|/R in move[0] (device). */
int cfmm_sum_update_multipliers(const cfmm_bucket* bucket, const double* lambda, double* theta_bar_out,
                                double* move, void* stream);

/* cudaMemsetAsync(ptr, 0, bytes) on the stream -- lets a host language zero psi/arb without torch. */
int cfmm_zero(void* ptr, int64_t bytes, void* stream);

/* Tuning knob for experiments (0 = auto: TMA-staged kernel when the layout allows; 1 = LDG kernel +
 * global red.add; 2 = LDG kernel + shared-memory privatised histogram; 3 = TMA-staged kernel).  Not part of the reference-facing surface. */
int cfmm_set_scatter_mode(int32_t mode);

/* Introspection: number of kernel launches issued by this library since load / last reset. */
int64_t cfmm_launch_count(void);
void cfmm_reset_launch_count(void);
const char* cfmm_last_cuda_error(void);
const char* cfmm_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CFMM_B200_H */
