// One-shot SUM all-reduce of a small fp32 buffer over NVLink peer memory: the data-parallel training step reduces
// the network's gradients (MLPs + pyramid convolutions: 449,381 floats = 1.8 MB) every step.  As an NCCL collective that message is
// pure latency -- and on the main communicator it queues behind the 312 MB point-table all-reduce that train_step starts first.  Here every rank keeps its local
// gradients in a SYMMETRIC buffer (the same allocation mapped into every peer's address space); after a device-side barrier one
// kernel per rank reads all W peer copies straight over NVLink / NVSwitch and writes the sum to local memory:
//   * peer_sum_kernel:      W unicast loads per element, summed in rank order -> bit-identical result on every rank;
//   * multimem_sum_kernel:  ONE multimem.ld_reduce per 16 bytes on the multicast address -- the NVSwitch pulls the W copies and
//                           adds them in the switch (NVLS), 1/W of the NVLink ingress of the unicast version.
// 1.8 MB x 8 peers = 14 MB per rank over 900 GB/s links: the kernel itself takes ~10-20 us.  Measured honestly (DESIGN.md 6,
// profiles/r2_peer_check_n*.txt, r2_dp_timeline_n8*.txt): with its two multi-tensor copies and two barrier launches the call costs
// ~150 us in isolation (NCCL reduces this message in 19-32 us back to back), and inside the step the interval up to the barrier is
// 0.55-0.66 ms with EITHER implementation -- that is where the ranks' load imbalance surfaces.  What the own path buys is independence
// from the NCCL stream's issue order (the message does not queue behind the big all-reduce) and a defined summation order.
// The reference has no multi-GPU path; the semantic is NCCL's ncclAllReduce(ncclSum) on the same buffer.
#include "common.cuh"
#include "hnr.h"

namespace {

constexpr int MAX_PEERS = 16;

struct PeerArgs {
    const float* src[MAX_PEERS];
    int world;
    int64_t n4;          // number of float4 elements
    float* out;
};

__global__ void __launch_bounds__(256) peer_sum_kernel(const __grid_constant__ PeerArgs A) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < A.n4; i += stride) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
        for (int r0 = 0; r0 < A.world; r0 += 4) {
            // four peers' loads in flight together; the additions keep rank order
            float4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                v[j] = (r0 + j < A.world) ? __ldcv(reinterpret_cast<const float4*>(A.src[r0 + j]) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 4; ++j) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
        }
        reinterpret_cast<float4*>(A.out)[i] = acc;
    }
}

__global__ void __launch_bounds__(256) multimem_sum_kernel(const float* __restrict__ mc, int64_t n4, float* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 v;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                     : "l"(reinterpret_cast<const float4*>(mc) + i)
                     : "memory");
        reinterpret_cast<float4*>(out)[i] = v;
    }
}

}  // namespace

// out[i] = sum over r < world of peers[r][i], i < n (n % 4 == 0, 16-byte aligned buffers).  `peers` are the device addresses of the W
// copies of a symmetric buffer as mapped into THIS process (own copy included, in rank order).  The caller brackets the launch
// with barriers across the ranks: before (every copy is complete) and after (nobody overwrites a copy a peer still reads).
extern "C" int hnr_peer_sum_f32(const void* const* peers, int world, int64_t n, float* out, void* stream) {
    HNR_CHECK_ARG(world >= 1 && world <= MAX_PEERS, "peer_sum: 1..16 peers");
    HNR_CHECK_ARG(n % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "peer_sum: n must be a multiple of 4, out 16-byte aligned");
    if (n == 0) return HNR_OK;
    PeerArgs A{};
    for (int r = 0; r < world; ++r) {
        HNR_CHECK_ARG(peers[r] != nullptr && (reinterpret_cast<uintptr_t>(peers[r]) & 15) == 0, "peer_sum: null / unaligned peer buffer");
        A.src[r] = static_cast<const float*>(peers[r]);
    }
    A.world = world; A.n4 = n / 4; A.out = out;
    const int64_t blocks = hnr_cdiv(A.n4, 256);
    peer_sum_kernel<<<(unsigned)(blocks < 4 * HNR_NUM_SMS ? blocks : 4 * HNR_NUM_SMS), 256, 0, (cudaStream_t)stream>>>(A);
    HNR_CHECK_LAUNCH("peer_sum");
    return HNR_OK;
}

// the same through the multicast mapping of the symmetric buffer (in-switch reduction); `mc` = multicast device address
extern "C" int hnr_multimem_sum_f32(const void* mc, int64_t n, float* out, void* stream) {
    HNR_CHECK_ARG(mc != nullptr && n % 4 == 0 && (reinterpret_cast<uintptr_t>(mc) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                  "multimem_sum: multicast address required, n % 4 == 0, 16-byte aligned");
    if (n == 0) return HNR_OK;
    const int64_t n4 = n / 4, blocks = hnr_cdiv(n4, 256);
    multimem_sum_kernel<<<(unsigned)(blocks < 4 * HNR_NUM_SMS ? blocks : 4 * HNR_NUM_SMS), 256, 0, (cudaStream_t)stream>>>(
        static_cast<const float*>(mc), n4, out);
    HNR_CHECK_LAUNCH("multimem_sum");
    return HNR_OK;
}
