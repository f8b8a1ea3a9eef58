// cfmm_allreduce.cu -- one-shot all-reduce (sum) of the small [psi | arb] vector over NVLink peer memory.
//
// pools shard across GPUs and every dual evaluation ends with ONE all-reduce of n_tokens+1 doubles
// (32 KB at 4096 tokens).  At that size NCCL is pure latency (~13 us inside a CUDA graph, ~18 us eager, measured);
// here every rank simply reads all peers' partial vectors through NVLink-mapped pointers (torch symmetric memory:
// cudaMalloc'd buffers exchanged over the process group) and sums them in rank order, so the result is
// bit-identical on all ranks (the dual iterate nu must not drift between ranks).  Hand-shake: one system-scope
// release store per peer into its signal pad, one acquire spin per peer on our own pad; per-CTA slots so the CTAs of
// the grid need no sync among themselves.  Buffers rotate over 3 slots (see pools.py), which makes reuse safe without
// a second hand-shake: a rank that has seen everybody's "ready k" knows everybody finished reading slot k-1.
#include "cfmm_dev.cuh"

using namespace cfmm;

namespace {

constexpr int kArThreads = 256;     // one output element per thread
constexpr int kArMaxCtas = 64;      // signal-pad slots reserved per channel: kArMaxCtas * world words
constexpr int kPadBase = 512;       // first signal-pad word we use (torch's own barriers live below)
constexpr int kMaxWorld = 16;

// release: orders this rank's partial vector (the pool kernels' red.adds, made visible to this grid by
// griddepcontrol.wait) before the "ready" flag; pairs with the peer's ld.acquire.sys on its pad
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// Every CTA hand-shakes on its own pad slots (no intra-grid sync), then each thread fetches its element from all
// peers with independent loads in flight and adds them in rank order.
__global__ void __launch_bounds__(kArThreads)
k_allreduce_oneshot(const double* const* __restrict__ bufs, uint32_t* const* __restrict__ pads, int rank, int world,
                    long long offset, int n, double* __restrict__ out, uint32_t seq, int channel) {
    // the partial vector of this rank was produced by earlier kernels on this stream: wait for them (PDL), after
    // that every write of theirs (red.add resolved in this GPU's L2) is visible to peers reading over NVLink
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");     // the next kernel may start its ramp now
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int tid = threadIdx.x;
    const int slot0 = kPadBase + (channel * kArMaxCtas + blockIdx.x) * world;
    if (tid < world && tid != rank) {
        st_release_sys(pads[tid] + slot0 + rank, seq);                       // "my partial `seq` is ready"
        const uint32_t* mine = pads[rank] + slot0 + tid;
        while ((int)(ld_acquire_sys(mine) - seq) < 0) { }                    // peer `tid` is ready too
    }
    __syncthreads();
    const int j = blockIdx.x * kArThreads + tid;
    if (j < n) {
        double v[kMaxWorld];
#pragma unroll
        for (int r = 0; r < kMaxWorld; ++r)
            if (r < world) v[r] = ld_relaxed_sys_f64(bufs[r] + offset + j);
        double s = 0.0;
#pragma unroll
        for (int r = 0; r < kMaxWorld; ++r)
            if (r < world) s += v[r];                                        // rank order: same bits on every rank
        out[j] = s;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Low-latency variant ("LL", the protocol NCCL uses for small messages): PUSH instead of pull.  Every rank writes its
// element j straight into a receive area of every peer as ONE 16-byte store {value, seq}; 16-byte aligned vector
// stores arrive atomically, so the flag travels with the data and no separate hand-shake or fence is needed.  The
// receiver polls its own (local) memory until the flag equals seq, then adds the values in rank order.  One NVLink
// one-way trip instead of the three of signal + remote load.  Receive areas rotate over 3 slots: a rank that has
// received everybody's step-k data knows everybody finished reading step k-1.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kArThreads)
k_allreduce_ll(const double* __restrict__ local, LLCell* const* __restrict__ recv, int rank, int world, int n,
               long long slot_off, long long src_stride, double* __restrict__ out, unsigned long long seq) {
    // Let the NEXT kernel on the stream start its ramp right away (it only touches constant tables before its own
    // griddepcontrol.wait), then wait for the kernels that produced `local`.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int j = blockIdx.x * kArThreads + threadIdx.x;
    if (j >= n) return;
    const double mine = local[j];
#pragma unroll
    for (int r = 0; r < kMaxWorld; ++r)
        if (r < world && r != rank) st_ll(recv[r] + slot_off + (long long)rank * src_stride + j, mine, seq);
    double v[kMaxWorld];
#pragma unroll
    for (int r = 0; r < kMaxWorld; ++r) {
        if (r < world) {
            if (r == rank) { v[r] = mine; continue; }
            const LLCell* c = recv[rank] + slot_off + (long long)r * src_stride + j;
            unsigned long long f;
            do { ld_ll(c, v[r], f); } while (f != seq);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < kMaxWorld; ++r)
        if (r < world) s += v[r];                                  // rank order: same bits on every rank
    out[j] = s;
}

}  // namespace

extern "C" {

int cfmm_allreduce_oneshot(const void* peer_bufs_dev, const void* peer_pads_dev, int32_t rank, int32_t world,
                           int64_t offset_elems, int32_t n, double* out, uint32_t seq, int32_t channel, void* stream) {
    if (!peer_bufs_dev || !peer_pads_dev || !out) return CFMM_E_NULL;
    if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world || n <= 0 || channel < 0 || channel > 3)
        return CFMM_E_SIZE;
    const int ctas = (n + kArThreads - 1) / kArThreads;
    if (ctas > kArMaxCtas) return CFMM_E_SIZE;        /* n_tokens <= 16383 for the fused path; larger: use NCCL */
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(kArThreads); cfg.dynamicSmemBytes = 0;
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k_allreduce_oneshot, static_cast<const double* const*>(peer_bufs_dev),
                       static_cast<uint32_t* const*>(const_cast<void*>(peer_pads_dev)), (int)rank, (int)world,
                       (long long)offset_elems, (int)n, out, seq, (int)channel);
    return check_launch();
}

int cfmm_allreduce_ll(const double* local, const void* peer_recv_dev, int32_t rank, int32_t world, int32_t n,
                      int64_t slot_off_cells, int64_t src_stride_cells, double* out, uint64_t seq, void* stream) {
    if (!local || !peer_recv_dev || !out) return CFMM_E_NULL;
    if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world || n <= 0 || seq == 0) return CFMM_E_SIZE;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((n + kArThreads - 1) / kArThreads); cfg.blockDim = dim3(kArThreads); cfg.dynamicSmemBytes = 0;
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k_allreduce_ll, local, static_cast<LLCell* const*>(const_cast<void*>(peer_recv_dev)), (int)rank,
                       (int)world, (int)n, (long long)slot_off_cells, (long long)src_stride_cells, out,
                       (unsigned long long)seq);
    return check_launch();
}

}  // extern "C"
