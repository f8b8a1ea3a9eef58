// cfmm_allreduce.cu -- all-reduce (sum) of the small [psi | arb] vector over NVLink peer memory.
//
// pools shard across GPUs and every dual evaluation ends with ONE all-reduce of n_tokens+1 doubles
// (32 KB at 4096 tokens).  At that size NCCL is pure latency (~13 us inside a CUDA graph, ~18 us eager, measured in
// round 1); here every rank pushes its partial vector straight into the peers' memory (torch symmetric memory:
// cudaMalloc'd buffers exchanged over the process group) and sums what it received in rank order, so the result is
// bit-identical on all ranks (the dual iterate nu must not drift between ranks).  A pull variant (signal pads + remote
// loads) was measured slower at N = 8 (21.8 vs 19.0 us per step) and removed.
#include "cfmm_dev.cuh"

using namespace cfmm;

namespace {

constexpr int kArThreads = 256;     // one output element per thread
constexpr int kMaxWorld = 16;

// ---------------------------------------------------------------------------------------------------------------
// Low-latency protocol ("LL", what NCCL uses for small messages): PUSH.  Every rank writes its
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
    // Poll all peers' cells TOGETHER: the loads of a round are independent (one L2 round trip for the lot instead of one
    // per peer); cells that carry this step's sequence number are taken, the rest are asked again.
    double v[kMaxWorld];
    unsigned pending = 0u;
#pragma unroll
    for (int r = 0; r < kMaxWorld; ++r) {
        v[r] = mine;
        if (r < world && r != rank) pending |= 1u << r;
    }
    const LLCell* base = recv[rank] + slot_off + j;
    const long long t0 = clock64();
    unsigned spins = 0;
    while (pending) {
        double t[kMaxWorld];
        unsigned long long f[kMaxWorld];
#pragma unroll
        for (int r = 0; r < kMaxWorld; ++r)
            if (pending >> r & 1u) ld_ll(base + (long long)r * src_stride, t[r], f[r]);
#pragma unroll
        for (int r = 0; r < kMaxWorld; ++r)
            if ((pending >> r & 1u) && f[r] == seq) { v[r] = t[r]; pending &= ~(1u << r); }
        // a peer that never shows up (crashed rank, mismatched call sequence) must not hang the GPU: ~3 s of SM clocks
        // (looked at every 64 rounds), then the element becomes NaN and the caller's certificate fails loudly
        if (pending && (++spins & 63u) == 0u && (unsigned long long)(clock64() - t0) > 6000000000ull) {
#pragma unroll
            for (int r = 0; r < kMaxWorld; ++r)
                if (pending >> r & 1u) v[r] = __longlong_as_double(0x7ff8000000000000ll);
            break;
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
