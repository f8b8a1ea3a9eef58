// cfmm_layout.cu -- native builder of the token-blocked layout (cfmm_blocked_pairs) for constant-product pools.
//
// The reference describes a problem by local_indices / reserves / fees (arbitrage.py:6-28) and turns the indices into
// dense 0/1 matrices A_i (arbitrage.py:42-48).  Here the same literals, uploaded as they are, become the tiled layout the
// evaluation kernels stream (cfmm_blocked.cuh) in three launches:
//   1. k_pool_keys      one key per pool: (token block of slot 0, token block of slot 1, slot-0 token) -- pools whose two
//                       tokens fall into the same pair of narrow token blocks become neighbours; also validates the data
//   2. cub radix sort   of (key, pool id) pairs: the blocked order
//   3. k_build_tiles    ONE CTA PER TILE, everything in shared memory: sort the tile's 2P half-edges by token, number
//                       the distinct tokens (16-bit local ids, token list), cut every token's run of flows into rows of
//                       <= row_cap entries, order the rows longest first, give every half-edge its slot in the
//                       row-ordered flow array, and gather the reserves / 1/gamma slabs into blocked order.
// It replaces ~240 torch launches (sort / unique / bincount / repeat_interleave / index ...) and their host round trips.
#include <cub/cub.cuh>

#include "cfmm_blocked.cuh"

using namespace cfmm;

namespace {

constexpr int LT = kTileT;                 // threads per tile CTA
constexpr int LP = kTileP;                 // pools per tile
constexpr int LH = 2 * kTileP;             // half-edges per tile
constexpr int LI = 4;                      // items per thread of the block sorts (LT * LI == LH)
static_assert(LT * LI == LH, "block sort shape");
constexpr int LR = BlockedCfg<kTileP>::kRowsMax;

struct BuildArgs {
    long long m;
    int n_tokens, nb, row_cap, key_bits, tok_bits;
    const int32_t* idx;        // [m][2]
    const double* R;           // [m][2]
    const double* gamma;       // [m]
    const uint32_t* order;     // [m] sorted pool ids (blocked position -> pool)
    double *r0, *r1, *gi;      // [T * P]
    uint32_t *lid, *pos;       // [T * P]
    uint32_t* rows;            // [T][LR]
    int32_t* tok;              // [T][P]
    int4* desc;                // [T]
    int32_t* status;           // [0] tiles touching more than P tokens, [1] invalid pools, [2] total rows
};

__global__ void __launch_bounds__(256)
k_pool_keys(long long m, int n_tokens, int nb, const int32_t* __restrict__ idx, const double* __restrict__ R,
            const double* __restrict__ gamma, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, int32_t* status) {
    int bad = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
        const int a = idx[2 * i], b = idx[2 * i + 1];
        const double R0 = R[2 * i], R1 = R[2 * i + 1], g = gamma[i];
        const bool ok = a >= 0 && a < n_tokens && b >= 0 && b < n_tokens && a != b && R0 > 0.0 && R1 > 0.0 && isfinite(R0) &&
                        isfinite(R1) && g > 0.0 && g <= 1.0;
        bad |= ok ? 0 : 1;
        const long long ba = ok ? ((long long)a * nb) / n_tokens : 0, bb = ok ? ((long long)b * nb) / n_tokens : 0;
        keys[i] = (uint32_t)((ba * nb + bb) * n_tokens + (ok ? a : 0));
        vals[i] = (uint32_t)i;
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicAdd(status + 1, 1);
}

using Sort = cub::BlockRadixSort<uint32_t, LT, LI, uint32_t>;
using Scan = cub::BlockScan<int, LT>;

struct TileSmem {                          // dynamic shared memory of one tile CTA (> 48 KB with the cub scratch)
    union { typename Sort::TempStorage sort; typename Scan::TempStorage scan; } tmp;
    uint32_t sk[LH];                       // half-edges sorted by token: token id; later: row sort keys / sorted row lengths
    uint32_t sv[LH];                       //                              slot << 15 | pool-in-tile
    uint16_t gid[LH];                      // local token id of every sorted half-edge
    int gs[LP + 2];                        // first sorted half-edge of every local token (+ end)
    int rfirst[LP + 1];                    // first row of every local token
    uint16_t rrank[LR + 8];                // old row id -> position after the longest-first sort
    int rstart[LR + 8];                    // sorted position -> first slot of the row in the flow array
    uint16_t lidh[LP][2], posh[LP][2];
    int ntok, nrow;
};

__global__ void __launch_bounds__(LT)
k_build_tiles(const BuildArgs B) {
    extern __shared__ __align__(16) unsigned char tile_smem_raw[];
    TileSmem& M = *reinterpret_cast<TileSmem*>(tile_smem_raw);
    auto& tmp = M.tmp;
    uint32_t* sk = M.sk; uint32_t* sv = M.sv; uint16_t* gid = M.gid; int* gs = M.gs; int* rfirst = M.rfirst;
    uint32_t* rk = M.sk;                   // the token ids are dead once the token list is written
    uint16_t* rrank = M.rrank; int* rstart = M.rstart;
    auto& lidh = M.lidh; auto& posh = M.posh;
    int& s_ntok = M.ntok; int& s_nrow = M.nrow;
    const int tid = threadIdx.x;
    const long long tile = blockIdx.x;
    const long long p0 = tile * LP;
    const int np = (int)(B.m - p0 < LP ? B.m - p0 : LP);               // real pools in this tile
    const int nh = 2 * np;
    // ---- 1. the tile's half-edges, slot-major (slot 0 of every pool, then slot 1), sorted by token (stable)
    uint32_t key[LI], val[LI];
#pragma unroll
    for (int u = 0; u < LI; ++u) {
        const int h = tid * LI + u;                                     // blocked arrangement
        const int slot = h >= LP ? 1 : 0, l = h - slot * LP;
        if (l < np) {
            const uint32_t pool = B.order[p0 + l];
            key[u] = (uint32_t)B.idx[2 * (long long)pool + slot];
            val[u] = (uint32_t)(slot << 15 | l);
        } else {
            key[u] = 0xffffffffu; val[u] = 0u;                          // padding: sorts behind every token
        }
    }
    Sort(tmp.sort).Sort(key, val, 0, 32);
#pragma unroll
    for (int u = 0; u < LI; ++u) { sk[tid * LI + u] = key[u]; sv[tid * LI + u] = val[u]; }
    __syncthreads();
    // ---- 2. distinct tokens: heads -> local ids (exclusive scan), token list, group starts
    int head[LI], gpre[LI];
#pragma unroll
    for (int u = 0; u < LI; ++u) {
        const int i = tid * LI + u;
        head[u] = (i < nh && (i == 0 || sk[i] != sk[i - 1])) ? 1 : 0;
    }
    int ntok;
    Scan(tmp.scan).ExclusiveSum(head, gpre, ntok);
    __syncthreads();
    const bool bad = ntok > LP;                                          // more tokens than a tile may touch: not blockable
#pragma unroll
    for (int u = 0; u < LI; ++u) {
        const int i = tid * LI + u;
        if (i < nh) {
            const int g = gpre[u] + head[u] - 1;                         // local id of this half-edge's token
            gid[i] = (uint16_t)g;
            if (head[u] && !bad) { gs[g] = i; B.tok[tile * LP + g] = (int32_t)sk[i]; }
        }
    }
    if (tid == 0) { s_ntok = ntok; if (!bad) gs[ntok] = nh; }
    __syncthreads();
    if (bad) {
        if (tid == 0) { atomicAdd(B.status, 1); B.desc[tile] = make_int4(0, 0, 0, 0); }
        // keep the slabs and tables inert: unit reserves, zero ids / positions (never launched: the host falls back)
        for (int l = tid; l < LP; l += LT) {
            B.r0[p0 + l] = 1.0; B.r1[p0 + l] = 1.0; B.gi[p0 + l] = 1.0; B.lid[p0 + l] = 0u; B.pos[p0 + l] = 0u;
        }
        return;
    }
    // ---- 3. rows: every token's run of flows is cut into rows of <= row_cap entries
    const int cap = B.row_cap;
    int nsub[2], rpre[2];                                                 // LT * 2 = LP >= ntok tokens
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int g = tid * 2 + u;
        nsub[u] = g < ntok ? (gs[g + 1] - gs[g] + cap - 1) / cap : 0;
    }
    int nrow;
    Scan(tmp.scan).ExclusiveSum(nsub, rpre, nrow);
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int g = tid * 2 + u;
        if (g < ntok) {
            rfirst[g] = rpre[u];
            const int cnt = gs[g + 1] - gs[g];
            for (int s = 0; s < nsub[u]; ++s) {                          // old row id rpre + s: (63 - len) << 16 | id, ltok kept aside
                const int len = min(cap, cnt - cap * s);
                rk[rpre[u] + s] = (uint32_t)(63 - len) << 16 | (uint32_t)(rpre[u] + s);
            }
        }
    }
    if (tid == 0) s_nrow = nrow;
    __syncthreads();
    // ---- 4. longest rows first (the 32 rows a warp sums then have nearly equal trip counts): sort the row keys
    uint32_t rkey[LI], rdum[LI];
#pragma unroll
    for (int u = 0; u < LI; ++u) {
        const int r = tid * LI + u;
        rkey[u] = r < nrow ? rk[r] : 0xffffffffu;
        rdum[u] = 0u;
    }
    __syncthreads();
    Sort(tmp.sort).Sort(rkey, rdum, 0, 22);
    int rlen[LI], spre[LI];
#pragma unroll
    for (int u = 0; u < LI; ++u) {
        const int r = tid * LI + u;
        rlen[u] = r < nrow ? 63 - (int)(rkey[u] >> 16) : 0;
        if (r < nrow) rrank[rkey[u] & 0xffffu] = (uint16_t)r;
    }
    __syncthreads();
    int total;
    Scan(tmp.scan).ExclusiveSum(rlen, spre, total);
#pragma unroll
    for (int u = 0; u < LI; ++u) {
        const int r = tid * LI + u;
        if (r < nrow) { rstart[r] = spre[u]; rk[r] = (uint32_t)rlen[u]; }          // rk now: length of sorted row r
    }
    __syncthreads();
    // ---- 5. every half-edge: its row, its slot in the row-ordered flow array; the row table
    for (int i = tid; i < nh; i += LT) {
        const int g = gid[i], o = i - gs[g];
        const int r = rrank[rfirst[g] + o / cap];
        const int p = rstart[r] + o % cap;
        const uint32_t v = sv[i];
        const int slot = v >> 15, l = v & 0x7fffu;
        lidh[l][slot] = (uint16_t)g;
        posh[l][slot] = (uint16_t)p;
        if (o % cap == 0) B.rows[tile * LR + r] = (uint32_t)rstart[r] | rk[r] << 16 | (uint32_t)g << 22;
    }
    __syncthreads();
    // ---- 6. per-pool words and slabs, blocked order; padding pools write zero flows past the real entries
    for (int l = tid; l < LP; l += LT) {
        if (l < np) {
            const uint32_t pool = B.order[p0 + l];
            B.lid[p0 + l] = (uint32_t)lidh[l][0] | (uint32_t)lidh[l][1] << 16;
            B.pos[p0 + l] = (uint32_t)posh[l][0] | (uint32_t)posh[l][1] << 16;
            B.r0[p0 + l] = B.R[2 * (long long)pool]; B.r1[p0 + l] = B.R[2 * (long long)pool + 1];
            B.gi[p0 + l] = 1.0 / B.gamma[pool];
        } else {
            const int pl = l - np;
            B.lid[p0 + l] = 0u;
            B.pos[p0 + l] = (uint32_t)(nh + 2 * pl) | (uint32_t)(nh + 2 * pl + 1) << 16;
            B.r0[p0 + l] = 1.0; B.r1[p0 + l] = 1.0; B.gi[p0 + l] = 1.0;
        }
    }
    if (tid == 0) { B.desc[tile] = make_int4(s_ntok, s_nrow, 0, 0); atomicAdd(B.status + 2, s_nrow); }
}

inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

size_t sort_temp_bytes(long long m) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)m, 0, 32);
    return bytes;
}

}  // namespace

extern "C" {

int64_t cfmm_blocked_build_work_bytes(int64_t n_pools) {
    if (n_pools <= 0 || n_pools > 0x7fffffffLL) return CFMM_E_SIZE;
    return (int64_t)(4 * align_up(4 * (size_t)n_pools) + align_up(sort_temp_bytes(n_pools)));
}

/* Build the token-blocked layout of m constant-product pools on the device.  idx [m][2] int32, reserves [m][2] f64,
 * gamma [m] f64: the reference's local_indices / reserves / fees (arbitrage.py:6-28) as contiguous device arrays.
 * `out`: a cfmm_blocked_pairs whose array members point at caller-allocated device buffers of n_tiles = ceil(m / P)
 * tiles (strides from cfmm_blocked_layout_info); r0 / r1 / gamma_inv / lid / pos / rows / tok / desc are filled.
 * order [m] uint32 (device, out): pool at every blocked position.  status [4] int32 (device, zeroed by this call, out):
 * [0] tiles that touch more tokens than a tile may (the caller must fall back to a plain bucket for such problems),
 * [1] CTAs that saw invalid pools (reserves <= 0 or not finite, fees outside (0, 1], token ids out of range or equal),
 * [2] total rows.  Asynchronous on `stream`. */
int cfmm_blocked_build(int64_t n_pools, int32_t n_tokens, const int32_t* idx, const double* reserves, const double* gamma,
                       const cfmm_blocked_pairs* out, uint32_t* order, int32_t* status, void* work, int64_t work_bytes,
                       void* stream) {
    if (!idx || !reserves || !gamma || !out || !order || !status || !work) return CFMM_E_NULL;
    if (n_pools <= 0 || n_pools > 0x7fffffffLL || n_tokens <= 0) return CFMM_E_SIZE;
    if (work_bytes < cfmm_blocked_build_work_bytes(n_pools)) return CFMM_E_SIZE;
    const long long T = (n_pools + LP - 1) / LP;
    if (out->pools_per_tile != LP || out->n_tiles != T || out->n_pools != n_pools) return CFMM_E_SIZE;
    if (!out->r0 || !out->r1 || !out->gamma_inv || !out->lid || !out->pos || !out->rows || !out->tok || !out->desc) return CFMM_E_NULL;
    int row_cap = 32;
    cfmm_blocked_layout_info(nullptr, nullptr, nullptr, &row_cap);
    long long nb = (long long)llround(sqrt((double)n_pools / LP));
    if (nb < 1) nb = 1;
    if ((double)nb * nb * n_tokens >= 4294967296.0) return CFMM_E_SIZE;          // keys are 32-bit: the caller uses the general builder
    int key_bits = 1;
    while (key_bits < 32 && (1ull << key_bits) < (unsigned long long)(nb * nb) * (unsigned long long)n_tokens) ++key_bits;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned char* w = static_cast<unsigned char*>(work);
    auto take = [&](size_t bytes) { unsigned char* p = w; w += align_up(bytes); return p; };
    uint32_t* keys = reinterpret_cast<uint32_t*>(take(4 * (size_t)n_pools));
    uint32_t* vals = reinterpret_cast<uint32_t*>(take(4 * (size_t)n_pools));
    uint32_t* keys2 = reinterpret_cast<uint32_t*>(take(4 * (size_t)n_pools));
    take(4 * (size_t)n_pools);
    size_t temp_bytes = sort_temp_bytes(n_pools);
    void* temp = take(temp_bytes);
    cudaMemsetAsync(status, 0, 16, st);
    const int grid = (int)((n_pools + 255) / 256 < 4LL * num_sms() ? (n_pools + 255) / 256 : 4LL * num_sms());
    k_pool_keys<<<grid, 256, 0, st>>>(n_pools, n_tokens, (int)nb, idx, reserves, gamma, keys, vals, status);
    int rc = check_launch();
    if (rc) return rc;
    if (cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys, keys2, vals, order, (int)n_pools, 0, key_bits, st) != cudaSuccess) {
        g_last_err = cudaGetLastError();
        return CFMM_E_CUDA;
    }
    BuildArgs B;
    B.m = n_pools; B.n_tokens = n_tokens; B.nb = (int)nb; B.row_cap = row_cap; B.key_bits = key_bits; B.tok_bits = 32;
    B.idx = idx; B.R = reserves; B.gamma = gamma; B.order = order;
    B.r0 = const_cast<double*>(out->r0); B.r1 = const_cast<double*>(out->r1); B.gi = const_cast<double*>(out->gamma_inv);
    B.lid = const_cast<uint32_t*>(out->lid); B.pos = const_cast<uint32_t*>(out->pos);
    B.rows = const_cast<uint32_t*>(out->rows); B.tok = const_cast<int32_t*>(out->tok);
    B.desc = reinterpret_cast<int4*>(const_cast<int32_t*>(out->desc));
    B.status = status;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_build_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSmem));
        attr = true;
    }
    k_build_tiles<<<(int)T, LT, sizeof(TileSmem), st>>>(B);
    return check_launch();
}

}  // extern "C"
