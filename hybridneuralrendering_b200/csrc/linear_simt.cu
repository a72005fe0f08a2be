// Exact-fp32 SIMT linear layers (forward, data-gradient, weight-gradient) used by the training
// backward pass and as the bit-faithful fp32 path next to the tcgen05 3xTF32 kernels (mlp_tc.cu).
// One layer = y = act(concat(A1,A2,A3) . W^T + b [+ residual]) with W (N,K) row-major exactly as
// torch.nn.Linear stores it, so the reference's checkpoints are consumed unchanged
// (reference layers: models/aggregators/point_aggregators.py:484-683).
//
// Classic register-tiled SGEMM: 64x64 CTA tile, BK=16, 256 threads, 4x4 outputs per thread,
// operands staged through shared memory as [BK][64+pad].  fp32 FFMA only (tensor cores would need
// the 3xTF32 split; see mlp_tc.cu).
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, PADT = 4;

struct Cat3 {
    const float* p[3];
    int ld[3];
    int k[3];        // widths; K = k0+k1+k2
    int64_t mod[3];  // if > 0 the source has `mod` rows and row m reads row m % mod (a block shared by V views)
};

__device__ __forceinline__ float cat_load(const Cat3& a, int64_t m, int k) {
    int s = 0;
    if (k >= a.k[0]) { k -= a.k[0]; s = 1; if (k >= a.k[1]) { k -= a.k[1]; s = 2; } }
    if (a.mod[s] > 0) m %= a.mod[s];
    return a.p[s][m * a.ld[s] + k];
}

struct Cat3Out {
    float* p[3];
    int ld[3];
    int k[3];
};

__device__ __forceinline__ void mma_tile(const float (&As)[BK][BM + PADT], const float (&Bs)[BK][BN + PADT], float (&acc)[4][4],
                                         int ty, int tx) {
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
}

// ---------------------------------------------------------------- forward
// Y[m,n] = act( sum_k A[m,k] W[n,k] + b[n] (+ res[m,n]) )
__global__ void __launch_bounds__(256)
linear_fwd_kernel(Cat3 A, const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ res, int ldres,
                  float* __restrict__ Y, int ldy, int64_t M, int N, int K, int act) {
    __shared__ float As[BK][BM + PADT];
    __shared__ float Bs[BK][BN + PADT];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    float acc[4][4] = {};
    const int lk = tid & 15, lr = tid >> 4;   // loader: 16 consecutive threads walk k (contiguous in A and W)
    for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int r = lr + 16 * j;
            int k = k0 + lk;
            int64_t m = m0 + r;
            As[lk][r] = (m < M && k < K) ? cat_load(A, m, k) : 0.f;
            int n = n0 + r;
            Bs[lk][r] = (n < N && k < K) ? W[(int64_t)n * K + k] : 0.f;
        }
        __syncthreads();
        mma_tile(As, Bs, acc, ty, tx);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j] + (bias ? bias[n] : 0.f);
            v = apply_act(v, act);
            if (res) v += res[m * ldres + n];
            Y[m * ldy + n] = v;
        }
    }
}

// ---------------------------------------------------------------- data gradient
// dA[m,k] = sum_n dPre[m,n] W[n,k],  dPre = dY * act'(Y)
__global__ void __launch_bounds__(256)
linear_bwd_data_kernel(const float* __restrict__ dY, int lddy, const float* __restrict__ Y, int ldy, const float* __restrict__ W,
                       Cat3Out dA, int64_t M, int N, int K, int act) {
    __shared__ float As[BK][BM + PADT];   // [n][m]
    __shared__ float Bs[BK][BN + PADT];   // [n][k]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int kout0 = blockIdx.y * BN;
    float acc[4][4] = {};
    for (int n0 = 0; n0 < N; n0 += BK) {
        {   // dPre tile: contiguous along n
            const int ln = tid & 15, lr = tid >> 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int r = lr + 16 * j;
                int64_t m = m0 + r;
                int n = n0 + ln;
                float v = 0.f;
                if (m < M && n < N) {
                    v = dY[m * lddy + n];
                    if (act != HNR_ACT_NONE) v *= act_grad_from_out(Y[m * ldy + n], act);
                }
                As[ln][r] = v;
            }
        }
        {   // W tile: rows n, contiguous along k
            const int lc = tid & 63, lr = tid >> 6;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int nn = lr + 4 * j;
                int n = n0 + nn, k = kout0 + lc;
                Bs[nn][lc] = (n < N && k < K) ? W[(int64_t)n * K + k] : 0.f;
            }
        }
        __syncthreads();
        mma_tile(As, Bs, acc, ty, tx);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int k = kout0 + tx * 4 + j;
            if (k >= K) continue;
            int kk = k, s = 0;
            if (kk >= dA.k[0]) { kk -= dA.k[0]; s = 1; if (kk >= dA.k[1]) { kk -= dA.k[1]; s = 2; } }
            if (dA.p[s]) dA.p[s][m * dA.ld[s] + kk] = acc[i][j];
        }
    }
}

// ---------------------------------------------------------------- weight gradient
// dW[n,k] += sum_m dPre[m,n] A[m,k];  db[n] += sum_m dPre[m,n].  grid.z splits M; fp32 atomics.
__global__ void __launch_bounds__(256)
linear_bwd_weight_kernel(const float* __restrict__ dY, int lddy, const float* __restrict__ Y, int ldy, Cat3 A,
                         float* __restrict__ dW, float* __restrict__ db, int64_t M, int N, int K, int act, int64_t m_per_split) {
    __shared__ float As[BK][BM + PADT];   // [m][n]
    __shared__ float Bs[BK][BN + PADT];   // [m][k]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int n0 = blockIdx.x * BM;
    const int k0 = blockIdx.y * BN;
    const int64_t mbeg = (int64_t)blockIdx.z * m_per_split;
    const int64_t mend = (mbeg + m_per_split < M) ? (mbeg + m_per_split) : M;
    float acc[4][4] = {};
    float bsum = 0.f;                      // threads with tid<64 own db[n0+tid]
    const int lc = tid & 63, lr = tid >> 6;
    for (int64_t mm0 = mbeg; mm0 < mend; mm0 += BK) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int r = lr + 4 * j;
            int64_t m = mm0 + r;
            int n = n0 + lc, k = k0 + lc;
            float v = 0.f;
            if (m < mend && n < N) {
                v = dY[m * lddy + n];
                if (act != HNR_ACT_NONE) v *= act_grad_from_out(Y[m * ldy + n], act);
            }
            As[r][lc] = v;
            Bs[r][lc] = (m < mend && k < K) ? cat_load(A, m, k) : 0.f;
        }
        __syncthreads();
        mma_tile(As, Bs, acc, ty, tx);
        if (db && blockIdx.y == 0 && tid < 64) {
#pragma unroll
            for (int r = 0; r < BK; ++r) bsum += As[r][tid];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int n = n0 + ty * 4 + i;
        if (n >= N) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int k = k0 + tx * 4 + j;
            if (k >= K) continue;
            atomicAdd(&dW[(int64_t)n * K + k], acc[i][j]);
        }
    }
    if (db && blockIdx.y == 0 && tid < 64 && n0 + tid < N) atomicAdd(&db[n0 + tid], bsum);
}

template <typename PP> bool cat_ok(PP p, const int64_t* ld, const int64_t* k, int64_t K) {
    int64_t s = 0;
    for (int i = 0; i < 3; ++i) {
        if (k[i] < 0) return false;
        if (k[i] > 0 && p && p[i] && ld[i] < k[i]) return false;
        s += k[i];
    }
    return s == K;
}

}  // namespace

// N <= 4 outputs (density / colour heads): a GEMM tile would waste 60 of its 64 columns.  One warp per row: the lanes
// stride over the concatenated K (coalesced 128-byte reads), W lives in shared memory, warp-shuffle reduction.
constexpr int SMALLN_MAXK = 512;
constexpr int FWD_ROWS = 4;
__global__ void __launch_bounds__(256) linear_fwd_smalln_kernel(Cat3 A, const float* __restrict__ W, const float* __restrict__ bias,
                                                                 const float* __restrict__ res, int ldres, float* __restrict__ Y, int ldy,
                                                                 int64_t M, int N, int K, int act) {
    __shared__ float Ws[4 * SMALLN_MAXK];
    for (int i = threadIdx.x; i < N * K; i += blockDim.x) Ws[i] = W[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int b1 = A.k[0], b2 = A.k[0] + A.k[1];
    // FWD_ROWS rows per warp iteration: their loads are independent and issued together (a row is only K floats, i.e. a handful of
    // 128-byte lines per lane pass -- one row at a time leaves the warp waiting on 4 loads); per-row arithmetic order is unchanged
    for (int64_t mb = warp0 * FWD_ROWS; mb < M; mb += nwarps * FWD_ROWS) {
        const float *r0[FWD_ROWS], *r1[FWD_ROWS], *r2[FWD_ROWS];
#pragma unroll
        for (int r = 0; r < FWD_ROWS; ++r) {
            const int64_t m = mb + r < M ? mb + r : M - 1;      // clamped: loads stay in bounds, the row is not stored
            r0[r] = A.p[0] + (A.mod[0] > 0 ? m % A.mod[0] : m) * A.ld[0];
            r1[r] = A.k[1] > 0 ? A.p[1] + (A.mod[1] > 0 ? m % A.mod[1] : m) * A.ld[1] : nullptr;
            r2[r] = A.k[2] > 0 ? A.p[2] + (A.mod[2] > 0 ? m % A.mod[2] : m) * A.ld[2] : nullptr;
        }
        float acc[FWD_ROWS][4] = {};
        for (int k = lane; k < K; k += 32) {
            float x[FWD_ROWS];
#pragma unroll
            for (int r = 0; r < FWD_ROWS; ++r) x[r] = k < b1 ? r0[r][k] : (k < b2 ? r1[r][k - b1] : r2[r][k - b2]);
#pragma unroll
            for (int n = 0; n < 4; ++n)
                if (n < N) {
                    const float w = Ws[n * K + k];
#pragma unroll
                    for (int r = 0; r < FWD_ROWS; ++r) acc[r][n] = fmaf(x[r], w, acc[r][n]);
                }
        }
#pragma unroll
        for (int r = 0; r < FWD_ROWS; ++r) {
#pragma unroll
            for (int n = 0; n < 4; ++n) acc[r][n] = warp_sum(acc[r][n]);
            const int64_t m = mb + r;
            if (m < M && lane < N) {
                float y = apply_act((lane == 0 ? acc[r][0] : lane == 1 ? acc[r][1] : lane == 2 ? acc[r][2] : acc[r][3]) + (bias ? bias[lane] : 0.f), act);
                if (res) y += res[m * ldres + lane];
                Y[m * ldy + lane] = y;
            }
        }
    }
}

// Weight gradient of the same heads (N <= 4 outputs, K <= 128 inputs): dW[n,k] += sum_m dPre[m,n] A[m,k], db[n] += sum_m dPre[m,n].
// One warp per row, lanes stride over K (coalesced), per-warp register accumulators over a grid-stride row loop, one shared-memory
// reduction per CTA, then one global atomic per (n,k) per CTA.  HBM-bound: reads M x K once (the 64x64 GEMM tile of
// linear_bwd_weight_kernel wastes 60 of 64 output rows: 514 us for the 64->1 blend-weight head at M = 602k, trace of round 1).
constexpr int SMALLN_BWD_MAXK = 128;
constexpr int SMALLN_ROWS = 4;      // rows per warp iteration: their loads are issued together (a row is only K <= 128 floats)
__global__ void __launch_bounds__(256) linear_bwd_weight_smalln_kernel(const float* __restrict__ dY, int lddy, const float* __restrict__ Y,
                                                                        int ldy, Cat3 A, float* __restrict__ dW, float* __restrict__ db,
                                                                        int64_t M, int N, int K, int act) {
    __shared__ float sW[4 * SMALLN_BWD_MAXK + 4];
    for (int i = threadIdx.x; i < 4 * SMALLN_BWD_MAXK + 4; i += blockDim.x) sW[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int b1 = A.k[0], b2 = A.k[0] + A.k[1];
    float acc[4][4] = {};
    float bs[4] = {0.f, 0.f, 0.f, 0.f};
    for (int64_t mb = warp0 * SMALLN_ROWS; mb < M; mb += nwarps * SMALLN_ROWS) {
        float dp[SMALLN_ROWS], x[SMALLN_ROWS][4];
#pragma unroll
        for (int r = 0; r < SMALLN_ROWS; ++r) {
            const int64_t m = mb + r;
            dp[r] = 0.f;
            if (m < M && lane < N) {
                dp[r] = dY[m * lddy + lane];
                if (act != HNR_ACT_NONE) dp[r] *= act_grad_from_out(Y[m * ldy + lane], act);
            }
            const int64_t mc = m < M ? m : M - 1;       // clamped row: loads stay in bounds, dp = 0 removes its contribution
            const float* r0 = A.p[0] + (A.mod[0] > 0 ? mc % A.mod[0] : mc) * A.ld[0];
            const float* r1 = A.k[1] > 0 ? A.p[1] + (A.mod[1] > 0 ? mc % A.mod[1] : mc) * A.ld[1] : nullptr;
            const float* r2 = A.k[2] > 0 ? A.p[2] + (A.mod[2] > 0 ? mc % A.mod[2] : mc) * A.ld[2] : nullptr;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = lane + 32 * j;
                x[r][j] = k < K ? (k < b1 ? r0[k] : (k < b2 ? r1[k - b1] : r2[k - b2])) : 0.f;
            }
        }
#pragma unroll
        for (int r = 0; r < SMALLN_ROWS; ++r) {
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                const float d = __shfl_sync(0xffffffffu, dp[r], n);
                bs[n] += d;
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[n][j] = fmaf(d, x[r][j], acc[n][j]);
            }
        }
    }
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        if (n >= N) break;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = lane + 32 * j;
            if (k < K) atomicAdd(&sW[n * SMALLN_BWD_MAXK + k], acc[n][j]);
        }
        if (lane == 0) atomicAdd(&sW[4 * SMALLN_BWD_MAXK + n], bs[n]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N * K; i += blockDim.x) atomicAdd(&dW[i], sW[(i / K) * SMALLN_BWD_MAXK + (i % K)]);
    if (db && threadIdx.x < N) atomicAdd(&db[threadIdx.x], sW[4 * SMALLN_BWD_MAXK + threadIdx.x]);
}

// Data gradient of a NARROW column slice of a wide layer: dA[m, 0..KN) = sum_n dPre[m,n] W[n, k0 + 0..KN), KN <= 8 output columns,
// N = 128*NV reduction columns (NV <= 2).  Used for the 7 extra inputs of block3's first layer (colour, dir - view, <dir,view>;
// point_aggregators.py:1002-1010): the tensor-core slice kernel spends a full pass (1.3 ms at M = 562k) on them because its cost is
// reading and converting dY, not the MMAs.  Warp per row, the W slice lives in registers, float4 row loads, butterfly reduction:
// HBM-bound, one read of dY and Y.
template <int NV>
__global__ void __launch_bounds__(256, 4) linear_bwd_data_narrow_kernel(const float* __restrict__ dY, int lddy, const float* __restrict__ Y,
                                                                      int ldy, int act, const float* __restrict__ W, int ldw, int k0, int KN,
                                                                      float* __restrict__ dA, int ldda, int64_t M) {
    // lane owns reduction columns n = 128*v + 4*lane + c  (v < NV, c < 4): conflict-free float4 loads of the rows and of the W slice
    __shared__ float4 sw[8][NV][32];
    for (int i = threadIdx.x; i < 8 * NV * 32; i += blockDim.x) {
        const int k = i / (NV * 32), v = (i / 32) % NV, l = i % 32;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < KN) {
            const float* wp = W + (int64_t)(128 * v + 4 * l) * ldw + k0 + k;
            t = make_float4(wp[0], wp[ldw], wp[2 * (int64_t)ldw], wp[3 * (int64_t)ldw]);
        }
        sw[k][v][l] = t;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t m = warp0; m < M; m += nwarps) {
        float4 g[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            g[v] = *reinterpret_cast<const float4*>(dY + m * lddy + 128 * v + 4 * lane);
            if (act != HNR_ACT_NONE) {
                const float4 y4 = *reinterpret_cast<const float4*>(Y + m * ldy + 128 * v + 4 * lane);
                g[v].x *= act_grad_from_out(y4.x, act); g[v].y *= act_grad_from_out(y4.y, act);
                g[v].z *= act_grad_from_out(y4.z, act); g[v].w *= act_grad_from_out(y4.w, act);
            }
        }
        float out = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float a = 0.f;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const float4 w4 = sw[k][v][lane];
                a = fmaf(g[v].x, w4.x, a); a = fmaf(g[v].y, w4.y, a); a = fmaf(g[v].z, w4.z, a); a = fmaf(g[v].w, w4.w, a);
            }
            a = warp_sum(a);
            if (lane == k) out = a;
        }
        if (lane < KN) dA[m * ldda + lane] = out;
    }
}

extern "C" int hnr_linear_bwd_data_narrow(const float* dY, int64_t lddy, const float* Y, int64_t ldy, int act, const float* W, int64_t ldw,
                                          int64_t k0, int64_t kn, float* dA, int64_t ldda, int64_t M, int64_t N, void* stream) {
    if (M == 0) return HNR_OK;
    HNR_CHECK_ARG(kn > 0 && kn <= 8 && (N == 128 || N == 256), "linear_bwd_data_narrow: 1..8 output columns, 128 or 256 reduction columns");
    HNR_CHECK_ARG(lddy % 4 == 0 && ldy % 4 == 0 && ((uintptr_t)dY & 15) == 0 && ((uintptr_t)Y & 15) == 0,
                  "linear_bwd_data_narrow: dY / Y rows must be 16-byte aligned");
    const int64_t blocks = hnr_cdiv(M, 8);
    const int g = (int)(blocks < 8 * HNR_NUM_SMS ? blocks : 8 * HNR_NUM_SMS);
    if (N == 256)
        linear_bwd_data_narrow_kernel<2><<<g, 256, 0, (cudaStream_t)stream>>>(dY, (int)lddy, Y, (int)ldy, act, W, (int)ldw, (int)k0, (int)kn, dA,
                                                                             (int)ldda, M);
    else
        linear_bwd_data_narrow_kernel<1><<<g, 256, 0, (cudaStream_t)stream>>>(dY, (int)lddy, Y, (int)ldy, act, W, (int)ldw, (int)k0, (int)kn, dA,
                                                                             (int)ldda, M);
    HNR_CHECK_LAUNCH("linear_bwd_data_narrow");
    return HNR_OK;
}

extern "C" int hnr_linear_fwd(const float* const* a_ptr, const int64_t* a_ld, const int64_t* a_k, const int64_t* a_mod, const float* W, const float* bias,
                              const float* res, int64_t ldres, float* Y, int64_t ldy, int64_t M, int64_t N, int64_t K, int act,
                              void* stream) {
    if (M == 0) return HNR_OK;
    HNR_CHECK_ARG(M > 0 && N > 0 && K > 0 && ldy >= N, "linear_fwd: bad shape");
    HNR_CHECK_ARG(cat_ok(a_ptr, a_ld, a_k, K), "linear_fwd: concat widths must sum to K");
    Cat3 A;
    for (int i = 0; i < 3; ++i) { A.p[i] = a_ptr[i]; A.ld[i] = (int)a_ld[i]; A.k[i] = (int)a_k[i]; A.mod[i] = a_mod ? a_mod[i] : 0; }
    HNR_CHECK_ARG(!(res && act != HNR_ACT_NONE), "linear_fwd: residual only with act=none");
    if (N <= 4 && K <= SMALLN_MAXK) {
        const int64_t blocks = hnr_cdiv(M, 8 * FWD_ROWS);
        const int g = (int)(blocks < 16 * HNR_NUM_SMS ? blocks : 16 * HNR_NUM_SMS);
        linear_fwd_smalln_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(A, W, bias, res, (int)ldres, Y, (int)ldy, M, (int)N, (int)K, act);
        HNR_CHECK_LAUNCH("linear_fwd(small N)");
        return HNR_OK;
    }
    dim3 grid((unsigned)hnr_cdiv(M, BM), (unsigned)hnr_cdiv(N, BN));
    linear_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, W, bias, res, (int)ldres, Y, (int)ldy, M, (int)N, (int)K, act);
    HNR_CHECK_LAUNCH("linear_fwd");
    return HNR_OK;
}

extern "C" int hnr_linear_bwd_data(const float* dY, int64_t lddy, const float* Y, int64_t ldy, const float* W, float* const* da_ptr,
                                   const int64_t* da_ld, const int64_t* a_k, int64_t M, int64_t N, int64_t K, int act, void* stream) {
    if (M == 0) return HNR_OK;
    HNR_CHECK_ARG(M > 0 && N > 0 && K > 0, "linear_bwd_data: bad shape");
    HNR_CHECK_ARG(cat_ok(da_ptr, da_ld, a_k, K), "linear_bwd_data: concat widths must sum to K");
    Cat3Out dA;
    for (int i = 0; i < 3; ++i) { dA.p[i] = da_ptr[i]; dA.ld[i] = (int)da_ld[i]; dA.k[i] = (int)a_k[i]; }
    dim3 grid((unsigned)hnr_cdiv(M, BM), (unsigned)hnr_cdiv(K, BN));
    linear_bwd_data_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dY, (int)lddy, Y, (int)ldy, W, dA, M, (int)N, (int)K, act);
    HNR_CHECK_LAUNCH("linear_bwd_data");
    return HNR_OK;
}

extern "C" int hnr_linear_bwd_weight(const float* dY, int64_t lddy, const float* Y, int64_t ldy, const float* const* a_ptr,
                                     const int64_t* a_ld, const int64_t* a_k, const int64_t* a_mod, float* dW, float* db, int64_t M, int64_t N, int64_t K,
                                     int act, void* stream) {
    if (M == 0) return HNR_OK;
    HNR_CHECK_ARG(M > 0 && N > 0 && K > 0, "linear_bwd_weight: bad shape");
    HNR_CHECK_ARG(cat_ok(a_ptr, a_ld, a_k, K), "linear_bwd_weight: concat widths must sum to K");
    Cat3 A;
    for (int i = 0; i < 3; ++i) { A.p[i] = a_ptr[i]; A.ld[i] = (int)a_ld[i]; A.k[i] = (int)a_k[i]; A.mod[i] = a_mod ? a_mod[i] : 0; }
    if (N <= 4 && K <= SMALLN_BWD_MAXK) {
        const int64_t blocks = hnr_cdiv(M, 8 * 2 * SMALLN_ROWS);    // >= 2 iterations per warp before the CTA-level reduction
        const int g = (int)(blocks < 6 * HNR_NUM_SMS ? (blocks < 1 ? 1 : blocks) : 6 * HNR_NUM_SMS);
        linear_bwd_weight_smalln_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(dY, (int)lddy, Y, (int)ldy, A, dW, db, M, (int)N, (int)K, act);
        HNR_CHECK_LAUNCH("linear_bwd_weight(small N)");
        return HNR_OK;
    }
    const int64_t tiles = hnr_cdiv(N, BM) * hnr_cdiv(K, BN);
    int64_t splits = (4 * HNR_NUM_SMS + tiles - 1) / tiles;          // aim at >= 4 CTAs per SM
    int64_t max_splits = hnr_cdiv(M, 256);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    int64_t m_per = hnr_cdiv(hnr_cdiv(M, splits), BK) * BK;
    splits = hnr_cdiv(M, m_per);
    dim3 grid((unsigned)hnr_cdiv(N, BM), (unsigned)hnr_cdiv(K, BN), (unsigned)splits);
    linear_bwd_weight_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dY, (int)lddy, Y, (int)ldy, A, dW, db, M, (int)N, (int)K, act,
                                                                    m_per);
    HNR_CHECK_LAUNCH("linear_bwd_weight");
    return HNR_OK;
}

// ------------------------------------------------------------------------------------------------------------------------------
// Backward of a ONE-output layer h = act(y . w + b) in a single pass over its input rows (the sigmoid head of the blend-weight net:
// 64 -> 1 over V*Nv = 602 k rows per training step, models/aggregators/point_aggregators.py:1199-1217).  Per row m:
//     s = dH[m] * act'(h[m])        dY[m, :] = s * w        dW += s * y[m, :]        db += s
// The generic kernels made two passes (data gradient 143 us + weight gradient 145 us, each reading or writing the 154 MB of y / dY)
// with one row per warp slot; here K/4 lanes own one row (float4 per lane), 8 row groups are in flight per warp, y is read once and
// dY written once.  K must be 64 or 128.
template <int LPR>      // lanes per row: K / 4 (16 or 32)
__global__ void __launch_bounds__(256) linear_head_bwd_kernel(const float* __restrict__ dH, const float* __restrict__ Hout, int act,
                                                              const float* __restrict__ w, const float* __restrict__ Y, int64_t ldy,
                                                              int64_t M, float* __restrict__ dY, int64_t lddy, float* __restrict__ dW,
                                                              float* __restrict__ db) {
    constexpr int RPW = 32 / LPR;           // rows per warp instruction
    constexpr int UNROLL = 8;
    __shared__ float sW[4 * LPR + 1];
    for (int i = threadIdx.x; i < 4 * LPR + 1; i += blockDim.x) sW[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, l = lane % LPR, rsub = lane / LPR;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float4 w4 = __ldg(reinterpret_cast<const float4*>(w) + l);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float bs = 0.f;
    for (int64_t m0 = warp0 * (RPW * UNROLL); m0 < M; m0 += nwarps * (RPW * UNROLL)) {
        float4 y[UNROLL];
        float sv[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t m = m0 + u * RPW + rsub;
            y[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            sv[u] = 0.f;
            if (m < M) {
                y[u] = __ldg(reinterpret_cast<const float4*>(Y + m * ldy) + l);
                sv[u] = __ldg(dH + m) * act_grad_from_out(__ldg(Hout + m), act);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t m = m0 + u * RPW + rsub;
            const float sc = sv[u];
            if (m < M) reinterpret_cast<float4*>(dY + m * lddy)[l] = make_float4(sc * w4.x, sc * w4.y, sc * w4.z, sc * w4.w);
            acc.x = fmaf(sc, y[u].x, acc.x); acc.y = fmaf(sc, y[u].y, acc.y); acc.z = fmaf(sc, y[u].z, acc.z); acc.w = fmaf(sc, y[u].w, acc.w);
            if (l == 0) bs += sc;
        }
    }
    atomicAdd(&sW[4 * l + 0], acc.x); atomicAdd(&sW[4 * l + 1], acc.y); atomicAdd(&sW[4 * l + 2], acc.z); atomicAdd(&sW[4 * l + 3], acc.w);
    if (l == 0) atomicAdd(&sW[4 * LPR], bs);
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * LPR; i += blockDim.x) atomicAdd(&dW[i], sW[i]);
    if (threadIdx.x == 0 && db) atomicAdd(db, sW[4 * LPR]);
}

extern "C" int hnr_linear_head_bwd(const float* dH, const float* Hout, int act, const float* w, const float* Y, int64_t ldy, int64_t M,
                                   int64_t K, float* dY, int64_t lddy, float* dW, float* db, void* stream) {
    if (M == 0) return HNR_OK;
    HNR_CHECK_ARG(K == 64 || K == 128, "linear_head_bwd: K must be 64 or 128");
    HNR_CHECK_ARG(ldy % 4 == 0 && lddy % 4 == 0 && ((reinterpret_cast<uintptr_t>(Y) | reinterpret_cast<uintptr_t>(dY) | reinterpret_cast<uintptr_t>(w)) & 15) == 0,
                  "linear_head_bwd: rows must be 16-byte aligned");
    const int rpw = K == 64 ? 2 : 1;
    const int64_t blocks = hnr_cdiv(M, 8 * rpw * 8 * 2);           // >= 2 iterations per warp before the CTA-level reduction
    const int g = (int)(blocks < 8 * HNR_NUM_SMS ? (blocks < 1 ? 1 : blocks) : 8 * HNR_NUM_SMS);
    if (K == 64) linear_head_bwd_kernel<16><<<g, 256, 0, (cudaStream_t)stream>>>(dH, Hout, act, w, Y, ldy, M, dY, lddy, dW, db);
    else linear_head_bwd_kernel<32><<<g, 256, 0, (cudaStream_t)stream>>>(dH, Hout, act, w, Y, ldy, M, dY, lddy, dW, db);
    HNR_CHECK_LAUNCH("linear_head_bwd");
    return HNR_OK;
}
