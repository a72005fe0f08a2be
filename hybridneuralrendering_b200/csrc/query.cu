// Voxel-grid neural-point query.  Replaces the six pycuda kernels and
// the torch glue of models/neural_points/query_point_indices_worldcoords.py (claim_occ :237-297,
// map_coor2occ :299-334, fill_occ2pnts :336-381, mask_raypos :384-408, get_shadingloc :411-433,
// query_neigh_along_ray_layered :436-522, host side :540-711).
//
// B200-first data layout (not the reference's arrival-order tables):
//   * cell_start[n_cells+1]   dense CSR over voxels (int32) built by count -> scan -> fill;
//   * pts_sorted[n_in]        float4 {x,y,z,bitcast(point id)} grouped by voxel and, inside a
//                             voxel, in ascending point id -- deterministic, and a voxel's
//                             candidates are one contiguous, coalesced 16 B-per-lane load;
//   * occ_bits[n_cells/32]    dilated occupancy as a BIT field (1 bit/voxel instead of 4 B: the
//                             whole field of a 500^2x190 room is 6 MB and lives in L2).
// The grid is built once per point-set change and cached by the caller; the reference rebuilds
// it on every forward (:616).
//
// Per forward: ray_select (one warp per ray, ballot/popc first-SR selection, positions generated
// on the fly from the per-candidate parameters t -- the (R,D,3) position tensor never exists),
// knn (one warp per shading sample, warp-distributed sorted top-K, layered shell walk with the
// reference's early exit), then a ray-level scan + compaction that also emits the valid-sample
// list, so the host needs exactly one readback (R'', Nv).
//
// Arithmetic that decides bits is kept identical to the reference's compiled kernels:
// cell = (int)floorf((p - origin) / cell) with IEEE division; d2 = fma(z,z, fma(x,x, y*y)) as in the
// sm_100a SASS of the reference source (oracle/build_ref_query_cubin.py); position = campos +
// (dir * t) with separate rounding (two torch ops in the reference).
#include <stdlib.h>
#include "common.cuh"

#include "hnr.h"

namespace {

__device__ __forceinline__ bool cell_coord(const hnr_grid_t& g, float x, float y, float z, int& cx, int& cy, int& cz) {
    cx = (int)floorf(__fdiv_rn(__fsub_rn(x, g.origin[0]), g.cell[0]));
    cy = (int)floorf(__fdiv_rn(__fsub_rn(y, g.origin[1]), g.cell[1]));
    cz = (int)floorf(__fdiv_rn(__fsub_rn(z, g.origin[2]), g.cell[2]));
    return cx >= 0 && cx < g.dims[0] && cy >= 0 && cy < g.dims[1] && cz >= 0 && cz < g.dims[2];
}
__device__ __forceinline__ int64_t cell_lin(const hnr_grid_t& g, int cx, int cy, int cz) {
    return ((int64_t)cx * g.dims[1] + cy) * g.dims[2] + cz;
}

// ------------------------------------------------------------------------------------------------
// generic int32 exclusive scan (3 phases), out has n+1 entries
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_T = 256, SCAN_ITEMS = 16, SCAN_TILE = SCAN_T * SCAN_ITEMS;

__device__ __forceinline__ int block_excl_scan(int v, int* total, int* sh) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) sh[w] = inc;
    __syncthreads();
    if (w == 0) {
        int s = lane < (SCAN_T / 32) ? sh[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane < SCAN_T / 32) sh[lane] = s;
    }
    __syncthreads();
    int base = w > 0 ? sh[w - 1] : 0;
    *total = sh[SCAN_T / 32 - 1];
    __syncthreads();
    return base + inc - v;
}

__global__ void __launch_bounds__(SCAN_T) scan_reduce_kernel(const int32_t* __restrict__ in, int64_t n, int32_t* __restrict__ bsum) {
    __shared__ int sh[32];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    int s = 0;
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t idx = base + (int64_t)i * SCAN_T + threadIdx.x;
        if (idx < n) s += in[idx];
    }
    s = __reduce_add_sync(0xffffffffu, s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < SCAN_T / 32; ++i) t += sh[i];
        bsum[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(SCAN_T) scan_bsum_kernel(int32_t* __restrict__ bsum, int64_t nb, int32_t* __restrict__ out_total) {
    __shared__ int sh[32];
    int carry = 0;
    for (int64_t b0 = 0; b0 < nb; b0 += SCAN_T) {
        int64_t i = b0 + threadIdx.x;
        int v = i < nb ? bsum[i] : 0;
        int tot;
        int ex = block_excl_scan(v, &tot, sh);
        if (i < nb) bsum[i] = ex + carry;
        carry += tot;
    }
    if (threadIdx.x == 0) *out_total = carry;
}
__global__ void __launch_bounds__(SCAN_T) scan_apply_kernel(const int32_t* __restrict__ in, int64_t n, const int32_t* __restrict__ bsum,
                                                            int32_t* __restrict__ out) {
    __shared__ int sh[32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        s += v[i];
    }
    int tot;
    int ex = block_excl_scan(s, &tot, sh) + bsum[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = ex;
        ex += v[i];
    }
}

int launch_scan(const int32_t* in, int32_t* out, int64_t n, int32_t* scratch, cudaStream_t st) {
    // out[0..n-1] exclusive prefix, out[n] = total.  scratch: ceil(n/SCAN_TILE) ints.
    int64_t nb = hnr_cdiv(n, SCAN_TILE);
    if (nb == 0) { cudaMemsetAsync(out, 0, sizeof(int32_t), st); return 0; }
    scan_reduce_kernel<<<(unsigned)nb, SCAN_T, 0, st>>>(in, n, scratch);
    scan_bsum_kernel<<<1, SCAN_T, 0, st>>>(scratch, nb, out + n);
    scan_apply_kernel<<<(unsigned)nb, SCAN_T, 0, st>>>(in, n, scratch, out);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// grid build
// ------------------------------------------------------------------------------------------------
// pass 1: voxel id of every point (-1 outside) and the smallest in-grid point id
__global__ void grid_assign_kernel(const float* __restrict__ xyz, int64_t N, hnr_grid_t g, int32_t* __restrict__ cell_of_pt,
                                   int32_t* __restrict__ first_in) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int cx, cy, cz;
    bool in = cell_coord(g, xyz[i * 3], xyz[i * 3 + 1], xyz[i * 3 + 2], cx, cy, cz);
    cell_of_pt[i] = in ? (int32_t)cell_lin(g, cx, cy, cz) : -1;
    if (in && i < *first_in) atomicMin(first_in, (int32_t)i);
}

__device__ __forceinline__ void dilate(const hnr_grid_t& g, int c, uint32_t* occ_bits) {
    int cz = c % g.dims[2];
    int cy = (c / g.dims[2]) % g.dims[1];
    int cx = c / (g.dims[2] * g.dims[1]);
    for (int x = max(0, cx - g.qhalf_lo[0]); x <= min(g.dims[0] - 1, cx + g.qhalf_hi[0]); ++x)
        for (int y = max(0, cy - g.qhalf_lo[1]); y <= min(g.dims[1] - 1, cy + g.qhalf_hi[1]); ++y)
            for (int z = max(0, cz - g.qhalf_lo[2]); z <= min(g.dims[2] - 1, cz + g.qhalf_hi[2]); ++z) {
                int64_t l = cell_lin(g, x, y, z);
                uint32_t bit = 1u << (l & 31);
                if (!(occ_bits[l >> 5] & bit)) atomicOr(&occ_bits[l >> 5], bit);
            }
}

// pass 2: per-voxel counts; the first point of each voxel dilates the occupancy.  The skip voxel
// (the reference's occupied-slot 0, :366) is marked occupied but stores no points.
__global__ void grid_count_kernel(const int32_t* __restrict__ cell_of_pt, int64_t N, hnr_grid_t g, const int32_t* __restrict__ first_in,
                                  int32_t skip_cell_arg, int32_t* __restrict__ counts, uint32_t* __restrict__ occ_bits,
                                  int32_t* __restrict__ info) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int c = cell_of_pt[i];
    if (c < 0) return;
    int skip = skip_cell_arg;
    if (skip == -2) { int f = *first_in; skip = (f >= 0 && f < N) ? cell_of_pt[f] : -1; }
    int old;
    if (c == skip) {
        old = atomicAdd(&info[3], 1);             // points dropped with the skip voxel
    } else {
        old = atomicAdd(&counts[c], 1);
    }
    if (old == 0) {
        dilate(g, c, occ_bits);
        atomicAdd(&info[0], 1);                   // occupied voxels
        info[2] = skip;                           // same value from every writer
    }
    atomicMax(&info[1], old + 1);                 // max points in a voxel
}

// pass 3: unordered fill
__global__ void grid_fill_kernel(const int32_t* __restrict__ cell_of_pt, int64_t N, const int32_t* __restrict__ cell_start,
                                 int32_t* __restrict__ cursor, const int32_t* __restrict__ info, int32_t* __restrict__ tmp_idx) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int c = cell_of_pt[i];
    if (c < 0 || c == info[2]) return;
    int pos = cell_start[c] + atomicAdd(&cursor[c], 1);
    tmp_idx[pos] = (int32_t)i;
}

// pass 4: rank inside the voxel = number of smaller ids -> deterministic ascending order
__global__ void grid_rank_kernel(const float* __restrict__ xyz, const int32_t* __restrict__ cell_of_pt, int64_t N,
                                 const int32_t* __restrict__ cell_start, const int32_t* __restrict__ info,
                                 const int32_t* __restrict__ tmp_idx, float4* __restrict__ pts_sorted) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int c = cell_of_pt[i];
    if (c < 0 || c == info[2]) return;
    int beg = cell_start[c], end = cell_start[c + 1];
    int rank = 0;
    for (int j = beg; j < end; ++j) rank += (tmp_idx[j] < (int32_t)i) ? 1 : 0;
    pts_sorted[beg + rank] = make_float4(xyz[i * 3], xyz[i * 3 + 1], xyz[i * 3 + 2], __int_as_float((int32_t)i));
}

// ------------------------------------------------------------------------------------------------
// Q3 + Q5: candidate generation, occupancy test, first-SR selection.  One warp per ray.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ray_select_kernel(const float* __restrict__ campos, const float* __restrict__ raydir, const float* __restrict__ ts, int64_t ts_stride,
                  int64_t R, int D, int SR, hnr_grid_t g, const uint32_t* __restrict__ occ_bits, float* __restrict__ sample_loc,
                  int32_t* __restrict__ nsamp) {
    const int lane = threadIdx.x & 31;
    const int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ray >= R) return;
    const float ox = campos[0], oy = campos[1], oz = campos[2];
    const float dx = raydir[ray * 3], dy = raydir[ray * 3 + 1], dz = raydir[ray * 3 + 2];
    const float* t = ts + ray * ts_stride;
    float* out = sample_loc + ray * SR * 3;
    int count = 0;
    for (int c0 = 0; c0 < D && count < SR; c0 += 32) {
        int i = c0 + lane;
        bool hit = false;
        float x = 0.f, y = 0.f, z = 0.f;
        if (i < D) {
            float tt = t[i];
            x = __fadd_rn(ox, __fmul_rn(dx, tt));
            y = __fadd_rn(oy, __fmul_rn(dy, tt));
            z = __fadd_rn(oz, __fmul_rn(dz, tt));
            int cx, cy, cz;
            if (cell_coord(g, x, y, z, cx, cy, cz)) {
                int64_t l = cell_lin(g, cx, cy, cz);
                hit = (occ_bits[l >> 5] >> (l & 31)) & 1u;
            }
        }
        unsigned b = __ballot_sync(0xffffffffu, hit);
        if (hit) {
            int slot = count + __popc(b & ((1u << lane) - 1u));
            if (slot < SR) { out[slot * 3] = x; out[slot * 3 + 1] = y; out[slot * 3 + 2] = z; }
        }
        count += __popc(b);
    }
    if (lane == 0) nsamp[ray] = min(count, SR);
}

// ------------------------------------------------------------------------------------------------
// Q6: layered radius-bounded K-nearest.  CTA per ray, warp per shading sample.
// ------------------------------------------------------------------------------------------------
constexpr unsigned long long KEY_EMPTY = 0xffffffffffffffffull;

__device__ __forceinline__ unsigned long long shfl_u64(unsigned long long v, int src) {
    return __shfl_sync(0xffffffffu, v, src);
}
__device__ __forceinline__ unsigned long long shfl_up_u64(unsigned long long v, int d) {
    return __shfl_up_sync(0xffffffffu, v, d);
}

// merge one batch of up to 32 candidate keys (one per lane, KEY_EMPTY = none) into the sorted top-K held by lanes 0..K-1
__device__ __forceinline__ void knn_merge(unsigned long long key, int K, int lane, unsigned long long& topk, int& kid) {
    unsigned cand = __ballot_sync(0xffffffffu, key != KEY_EMPTY);
    kid += __popc(cand);
    unsigned long long kth = shfl_u64(topk, K - 1);
    unsigned m = __ballot_sync(0xffffffffu, key < kth);
    while (m) {
        int src = __ffs(m) - 1;
        m &= m - 1;
        unsigned long long x = shfl_u64(key, src);
        kth = shfl_u64(topk, K - 1);
        if (x < kth) {                         // warp-uniform
            unsigned long long prev = shfl_up_u64(topk, 1);
            if (lane == 0) prev = 0ull;
            topk = (x < prev) ? prev : ((x < topk) ? x : topk);
        }
    }
}

__device__ __forceinline__ unsigned long long knn_key(const float4 p, float sx, float sy, float sz, float r2) {
    float vx = __fsub_rn(p.x, sx), vy = __fsub_rn(p.y, sy), vz = __fsub_rn(p.z, sz);
    float d2 = __fmaf_rn(vz, vz, __fmaf_rn(vx, vx, __fmul_rn(vy, vy)));
    if (r2 == 0.f || d2 <= r2) return ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)__float_as_int(p.w);
    return KEY_EMPTY;
}

// process up to 32 candidates of one voxel (lane j < n owns candidate j)
__device__ __forceinline__ void knn_visit(const float4* __restrict__ pts, int start, int n, float sx, float sy, float sz, float r2,
                                          int K, int lane, unsigned long long& topk, int& kid) {
    for (int b0 = 0; b0 < n; b0 += 32) {
        unsigned long long key = KEY_EMPTY;
        if (b0 + lane < n) key = knn_key(pts[start + b0 + lane], sx, sy, sz, r2);
        knn_merge(key, K, lane, topk, kid);
    }
}

// The points of up to 32 voxels (lane c holds voxel c's start / count) walked in FULL 32-lane batches: candidate j of the
// concatenated list belongs to the first voxel whose inclusive count prefix exceeds j (binary search over the lanes with
// shuffles).  A voxel holds <= P (12..26) points, so the per-voxel walk of knn_visit leaves 20..60 % of the lanes idle and
// pays one dependent L2 round trip per non-empty voxel; here a sample costs ceil(candidates / 32) round trips.  The result is
// the same: the top-K is the set of the K smallest (d2, id) keys, ids are unique, and the candidate count `kid` is only
// looked at after a whole shell.
__device__ __forceinline__ void knn_visit_cells(const float4* __restrict__ pts, int start, int n, float sx, float sy, float sz, float r2,
                                                int K, int lane, unsigned long long& topk, int& kid) {
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    const int excl = incl - n;
    for (int b0 = 0; b0 < total; b0 += 32) {
        const int j = b0 + lane;
        int lo = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            int v = __shfl_sync(0xffffffffu, incl, lo + step - 1);
            if (v <= j) lo += step;
        }
        const int st = __shfl_sync(0xffffffffu, start, lo), ex = __shfl_sync(0xffffffffu, excl, lo);
        unsigned long long key = KEY_EMPTY;
        if (j < total) key = knn_key(pts[st + (j - ex)], sx, sy, sz, r2);
        knn_merge(key, K, lane, topk, kid);
    }
}

template <bool BATCHED>
__global__ void __launch_bounds__(128)
knn_kernel(const float* __restrict__ sample_loc, const int32_t* __restrict__ nsamp, int64_t R, int SR, int K, hnr_grid_t g,
           const int32_t* __restrict__ cell_start, const float4* __restrict__ pts_sorted, int32_t* __restrict__ pidx,
           int32_t* __restrict__ nvalid) {
    const int64_t ray = blockIdx.x;
    const int ns = nsamp[ray];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __shared__ int s_valid;
    if (threadIdx.x == 0) s_valid = 0;
    __syncthreads();
    int my_valid = 0;
    for (int s = w; s < ns; s += nw) {
        const float* lp = sample_loc + (ray * SR + s) * 3;
        const float sx = lp[0], sy = lp[1], sz = lp[2];
        int fx, fy, fz;
        cell_coord(g, sx, sy, sz, fx, fy, fz);
        unsigned long long topk = KEY_EMPTY;      // lanes 0..K-1: ascending (d2, id) keys
        int kid = 0;
        for (int layer = 0; layer < g.layers; ++layer) {
            const int side = 2 * layer + 1, ncell = side * side * side;
            for (int c0 = 0; c0 < ncell; c0 += 32) {
                int ci = c0 + lane;
                int start = 0, n = 0;
                if (ci < ncell) {
                    int ox = ci / (side * side) - layer, oy = (ci / side) % side - layer, oz = ci % side - layer;
                    int cx = fx + ox, cy = fy + oy, cz = fz + oz;
                    bool shell = max(abs(ox), max(abs(oy), abs(oz))) == layer;
                    if (shell && cx >= 0 && cx < g.dims[0] && cy >= 0 && cy < g.dims[1] && cz >= 0 && cz < g.dims[2]) {
                        int64_t l = cell_lin(g, cx, cy, cz);
                        start = cell_start[l];
                        n = min(cell_start[l + 1] - start, g.P);
                    }
                }
                if (BATCHED) {
                    knn_visit_cells(pts_sorted, start, n, sx, sy, sz, g.radius2, K, lane, topk, kid);
                } else {
                    unsigned nonempty = __ballot_sync(0xffffffffu, n > 0);
                    while (nonempty) {
                        int src = __ffs(nonempty) - 1;
                        nonempty &= nonempty - 1;
                        int st = __shfl_sync(0xffffffffu, start, src), nn = __shfl_sync(0xffffffffu, n, src);
                        knn_visit(pts_sorted, st, nn, sx, sy, sz, g.radius2, K, lane, topk, kid);
                    }
                }
            }
            if (kid >= K) break;
        }
        if (lane < K) pidx[(ray * SR + s) * K + lane] = (topk == KEY_EMPTY) ? -1 : (int32_t)(topk & 0xffffffffull);
        if (lane == 0 && topk != KEY_EMPTY) my_valid++;
    }
    if (lane == 0 && my_valid) atomicAdd(&s_valid, my_valid);
    __syncthreads();
    if (threadIdx.x == 0) nvalid[ray] = s_valid;
}

// ---- third generation: the default since round 2 (A/B on B200: bit-exact in the query suites incl. the comparison with the
//      reference's own kernels, query stage of the 800x800 frame 6.46 -> 4.96 ms; HNR_KNN_V2=1 selects the batched kernel above) ----
// The batched kernel above is instruction-issue bound (profiles/r1_knn_batched_ncu.md: IPC 3.2 of 4, ~800 warp instructions
// per sample).  Two of its biggest items are removed here:
//   * shell sizes are compile-time for the shipped kernel_size = 3 (shell 0 = the sample's own voxel, shell 1 = the 26
//     surrounding voxels): the voxel offsets `ci / side^2`, `(ci / side) % side`, `ci % side` become constant divisions;
//   * a batch with many candidates below the current K-th key (always the first batches of a sample, whose top-K is still empty)
//     is merged by sorting: bitonic sort of the 32 keys, C[lane] = min(top[lane], batch[31 - lane]) -- the 32 smallest of the
//     union as a bitonic sequence -- and one bitonic merge; ~150 instructions instead of ~16 per serial insertion.
// Same result by construction: the top-K is the set of the K smallest unique (d2, id) keys in ascending order.
__device__ __forceinline__ unsigned long long u64_min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
__device__ __forceinline__ unsigned long long u64_max(unsigned long long a, unsigned long long b) { return a < b ? b : a; }

__device__ __forceinline__ void knn_merge_v3(unsigned long long key, int K, int lane, unsigned long long& topk, int& kid) {
    const unsigned cand = __ballot_sync(0xffffffffu, key != KEY_EMPTY);
    kid += __popc(cand);
    unsigned long long kth = shfl_u64(topk, K - 1);
    unsigned m = __ballot_sync(0xffffffffu, key < kth);
    if (__popc(m) > 8) {
        // sort the batch ascending over the lanes
        unsigned long long v = key;
#pragma unroll
        for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
                const unsigned long long p = __shfl_xor_sync(0xffffffffu, v, j);
                const bool up = (lane & k) == 0, lower = (lane & j) == 0;
                v = (lower == up) ? u64_min(v, p) : u64_max(v, p);
            }
        }
        // the 32 smallest of (top, batch) as a bitonic sequence, then sorted; lanes >= K may keep further keys: harmless, they are
        // never written out and never smaller than lane K-1
        const unsigned long long rev = shfl_u64(v, 31 - lane);
        v = u64_min(topk, rev);
#pragma unroll
        for (int j = 16; j > 0; j >>= 1) {
            const unsigned long long p = __shfl_xor_sync(0xffffffffu, v, j);
            v = (lane & j) == 0 ? u64_min(v, p) : u64_max(v, p);
        }
        topk = v;
        return;
    }
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const unsigned long long x = shfl_u64(key, src);
        kth = shfl_u64(topk, K - 1);
        if (x < kth) {                         // warp-uniform
            unsigned long long prev = shfl_up_u64(topk, 1);
            if (lane == 0) prev = 0ull;
            topk = (x < prev) ? prev : ((x < topk) ? x : topk);
        }
    }
}

// walk the concatenated candidate list of up to 32 voxels (see knn_visit_cells) with the sorting merge
__device__ __forceinline__ void knn_visit_cells_v3(const float4* __restrict__ pts, int start, int n, float sx, float sy, float sz, float r2,
                                                   int K, int lane, unsigned long long& topk, int& kid) {
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    const int excl = incl - n;
    for (int b0 = 0; b0 < total; b0 += 32) {
        const int j = b0 + lane;
        int lo = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const int v = __shfl_sync(0xffffffffu, incl, lo + step - 1);
            if (v <= j) lo += step;
        }
        const int st = __shfl_sync(0xffffffffu, start, lo), ex = __shfl_sync(0xffffffffu, excl, lo);
        unsigned long long key = KEY_EMPTY;
        if (j < total) key = knn_key(pts[st + (j - ex)], sx, sy, sz, r2);
        knn_merge_v3(key, K, lane, topk, kid);
    }
}

// one shell with a compile-time edge (SIDE = 2 * layer + 1 <= 3: at most 27 voxels = one lane group)
template <int SIDE>
__device__ __forceinline__ void knn_shell_v3(const hnr_grid_t& g, const int32_t* __restrict__ cell_start, const float4* __restrict__ pts,
                                             int fx, int fy, int fz, float sx, float sy, float sz, int K, int lane,
                                             unsigned long long& topk, int& kid) {
    constexpr int LAYER = SIDE / 2, NCELL = SIDE * SIDE * SIDE;
    static_assert(NCELL <= 32, "one voxel per lane");
    int start = 0, n = 0;
    if (lane < NCELL) {
        const int ox = lane / (SIDE * SIDE) - LAYER, oy = (lane / SIDE) % SIDE - LAYER, oz = lane % SIDE - LAYER;
        const int cx = fx + ox, cy = fy + oy, cz = fz + oz;
        const bool shell = max(abs(ox), max(abs(oy), abs(oz))) == LAYER;
        if (shell && cx >= 0 && cx < g.dims[0] && cy >= 0 && cy < g.dims[1] && cz >= 0 && cz < g.dims[2]) {
            const int64_t l = cell_lin(g, cx, cy, cz);
            start = cell_start[l];
            n = min(cell_start[l + 1] - start, g.P);
        }
    }
    knn_visit_cells_v3(pts, start, n, sx, sy, sz, g.radius2, K, lane, topk, kid);
}

__global__ void __launch_bounds__(128)
knn_kernel_v3(const float* __restrict__ sample_loc, const int32_t* __restrict__ nsamp, int64_t R, int SR, int K, hnr_grid_t g,
              const int32_t* __restrict__ cell_start, const float4* __restrict__ pts_sorted, int32_t* __restrict__ pidx,
              int32_t* __restrict__ nvalid) {
    const int64_t ray = blockIdx.x;
    const int ns = nsamp[ray];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __shared__ int s_valid;
    if (threadIdx.x == 0) s_valid = 0;
    __syncthreads();
    int my_valid = 0;
    for (int s = w; s < ns; s += nw) {
        const float* lp = sample_loc + (ray * SR + s) * 3;
        const float sx = lp[0], sy = lp[1], sz = lp[2];
        int fx, fy, fz;
        cell_coord(g, sx, sy, sz, fx, fy, fz);
        unsigned long long topk = KEY_EMPTY;      // lanes 0..K-1: ascending (d2, id) keys
        int kid = 0;
        knn_shell_v3<1>(g, cell_start, pts_sorted, fx, fy, fz, sx, sy, sz, K, lane, topk, kid);
        if (kid < K && g.layers > 1) knn_shell_v3<3>(g, cell_start, pts_sorted, fx, fy, fz, sx, sy, sz, K, lane, topk, kid);
        if (lane < K) pidx[(ray * SR + s) * K + lane] = (topk == KEY_EMPTY) ? -1 : (int32_t)(topk & 0xffffffffull);
        if (lane == 0 && topk != KEY_EMPTY) my_valid++;
    }
    if (lane == 0 && my_valid) atomicAdd(&s_valid, my_valid);
    __syncthreads();
    if (threadIdx.x == 0) nvalid[ray] = s_valid;
}

// ------------------------------------------------------------------------------------------------
// Q7: ray compaction + perspective coordinates + valid-sample list
// ------------------------------------------------------------------------------------------------
__global__ void ray_flags_kernel(const int32_t* __restrict__ nvalid, int64_t R, int32_t* __restrict__ keep) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < R) keep[i] = nvalid[i] > 0 ? 1 : 0;
}

// CTA per input ray; kept rays copy their rows to the compacted position
__global__ void __launch_bounds__(128)
ray_compact_kernel(const float* __restrict__ sample_loc_full, const int32_t* __restrict__ pidx_full, const int32_t* __restrict__ nsamp,
                   const int32_t* __restrict__ nvalid, const int32_t* __restrict__ ray_off, const int32_t* __restrict__ val_off,
                   const float* __restrict__ raydir, const float* __restrict__ campos, const float* __restrict__ camrot, int64_t R,
                   int SR, int K, int32_t* __restrict__ out_pidx, float* __restrict__ out_loc_pers, float* __restrict__ out_loc_w,
                   float* __restrict__ out_dirs, int8_t* __restrict__ ray_mask, int32_t* __restrict__ ray_ids,
                   int32_t* __restrict__ vlist, int32_t* __restrict__ counts_out) {
    const int64_t ray = blockIdx.x;
    const bool keep = nvalid[ray] > 0;
    if (threadIdx.x == 0) {
        ray_mask[ray] = keep ? 1 : 0;
        if (ray == 0) { counts_out[0] = ray_off[R]; counts_out[1] = val_off[R]; }
    }
    if (!keep) return;
    const int64_t dst = ray_off[ray];
    const int ns = nsamp[ray];
    if (threadIdx.x == 0) ray_ids[dst] = (int32_t)ray;
    const float dx = raydir[ray * 3], dy = raydir[ray * 3 + 1], dz = raydir[ray * 3 + 2];
    for (int i = threadIdx.x; i < SR; i += blockDim.x) {
        const float* src = sample_loc_full + (ray * SR + i) * 3;
        float x = src[0], y = src[1], z = src[2];
        float* ow = out_loc_w + (dst * SR + i) * 3;
        ow[0] = x; ow[1] = y; ow[2] = z;
        // w2pers (:96-103): camera-frame coordinates, x/z, y/z, z
        float sxx = x - campos[0], syy = y - campos[1], szz = z - campos[2];
        float cx = sxx * camrot[0] + syy * camrot[3] + szz * camrot[6];
        float cy = sxx * camrot[1] + syy * camrot[4] + szz * camrot[7];
        float cz = sxx * camrot[2] + syy * camrot[5] + szz * camrot[8];
        float* op = out_loc_pers + (dst * SR + i) * 3;
        op[0] = cx / cz; op[1] = cy / cz; op[2] = cz;
        float* od = out_dirs + (dst * SR + i) * 3;
        od[0] = dx; od[1] = dy; od[2] = dz;
    }
    for (int i = threadIdx.x; i < SR * K; i += blockDim.x) {
        int s = i / K;
        out_pidx[dst * SR * K + i] = (s < ns) ? pidx_full[ray * SR * K + i] : -1;
    }
    // valid-sample list in (ray, slot) order: a 128-thread ordered compaction
    __shared__ int s_cnt[5];
    int base = val_off[ray];
    for (int i0 = 0; i0 < SR; i0 += blockDim.x) {
        int i = i0 + threadIdx.x;
        bool v = i < ns && pidx_full[(ray * SR + i) * K] >= 0;
        unsigned b = __ballot_sync(0xffffffffu, v);
        int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        if (lane == 0) s_cnt[w] = __popc(b);
        __syncthreads();
        int before = 0, total = 0;
        for (int j = 0; j < (int)(blockDim.x >> 5); ++j) { if (j < w) before += s_cnt[j]; total += s_cnt[j]; }
        if (v) vlist[base + before + __popc(b & ((1u << lane) - 1u))] = (int32_t)(dst * SR + i);
        base += total;
        __syncthreads();
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" int hnr_exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, int32_t* scratch, void* stream) {
    launch_scan(in, out, n, scratch, (cudaStream_t)stream);
    HNR_CHECK_LAUNCH("exclusive_scan");
    return HNR_OK;
}

extern "C" int64_t hnr_scan_scratch_elems(int64_t n) { return hnr_cdiv(n, SCAN_TILE) + 1; }

// Build the grid.  Buffers (caller-owned, device):
//   cell_of_pt int32[N]; counts int32[n_cells] (zeroed by this call, reused as cursor);
//   cell_start int32[n_cells+1]; tmp_idx int32[N]; pts_sorted float4[N]; occ_bits u32[ceil(n_cells/32)];
//   scan_scratch int32[hnr_scan_scratch_elems(n_cells)];
//   info int32[8]: [0] occupied voxels, [1] max points in a voxel, [2] skip voxel, [3] points dropped
//   with it, [4] first in-grid point id, [5] stored points (written as cell_start[n_cells]).
// skip_cell: -1 none, -2 reference-like default (voxel of the first in-grid point), >=0 explicit.
extern "C" int hnr_grid_build(const float* xyz, int64_t N, const hnr_grid_t* g, int32_t skip_cell, int32_t* cell_of_pt,
                              int32_t* counts, int32_t* cell_start, int32_t* tmp_idx, void* pts_sorted, uint32_t* occ_bits,
                              int32_t* scan_scratch, int32_t* info, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    HNR_CHECK_ARG(g->n_cells > 0 && g->n_cells < (1ll << 31) - 64, "grid_build: voxel grid must have < 2^31 cells");
    HNR_CHECK_ARG(N >= 0 && N < (1ll << 31), "grid_build: too many points");
    HNR_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * g->n_cells, st));
    HNR_CUDA(cudaMemsetAsync(occ_bits, 0, sizeof(uint32_t) * hnr_cdiv(g->n_cells, 32), st));
    HNR_CUDA(cudaMemsetAsync(info, 0, sizeof(int32_t) * 8, st));
    int32_t big = 0x7fffffff;
    HNR_CUDA(cudaMemcpyAsync(info + 4, &big, sizeof(int32_t), cudaMemcpyHostToDevice, st));
    if (N > 0) {
        unsigned nb = (unsigned)hnr_cdiv(N, 256);
        grid_assign_kernel<<<nb, 256, 0, st>>>(xyz, N, *g, cell_of_pt, info + 4);
        grid_count_kernel<<<nb, 256, 0, st>>>(cell_of_pt, N, *g, info + 4, skip_cell, counts, occ_bits, info);
    }
    launch_scan(counts, cell_start, g->n_cells, scan_scratch, st);
    HNR_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * g->n_cells, st));
    if (N > 0) {
        unsigned nb = (unsigned)hnr_cdiv(N, 256);
        grid_fill_kernel<<<nb, 256, 0, st>>>(cell_of_pt, N, cell_start, counts, info, tmp_idx);
        grid_rank_kernel<<<nb, 256, 0, st>>>(xyz, cell_of_pt, N, cell_start, info, tmp_idx, (float4*)pts_sorted);
    }
    HNR_CUDA(cudaMemcpyAsync(info + 5, cell_start + g->n_cells, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    HNR_CHECK_LAUNCH("grid_build");
    return HNR_OK;
}

// Per-forward query over R rays.  sample_loc_full (R,SR,3) must be zero-filled by the caller.
//   ts: candidate parameters, (D) shared by all rays when ts_stride == 0, else (R,D) with ts_stride = D.
//   scratch int32: nsamp[R], nvalid[R], keep[R], ray_off[R+1], val_off[R+1], scan scratch.
// Outputs are sized for the upper bound R; counts_out[0] = R'' (rays kept), counts_out[1] = Nv.
extern "C" int hnr_query(const float* campos, const float* camrot, const float* raydir, const float* ts, int64_t ts_stride, int64_t R,
                         int64_t D, int64_t SR, int64_t K, const hnr_grid_t* g, const int32_t* cell_start, const void* pts_sorted,
                         const uint32_t* occ_bits, float* sample_loc_full, int32_t* pidx_full, int32_t* nsamp, int32_t* nvalid,
                         int32_t* keep, int32_t* ray_off, int32_t* val_off, int32_t* scan_scratch, int32_t* out_pidx,
                         float* out_loc_pers, float* out_loc_w, float* out_dirs, int8_t* ray_mask, int32_t* ray_ids, int32_t* vlist,
                         int32_t* counts_out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    HNR_CHECK_ARG(K > 0 && K <= 32, "query: K must be in 1..32");
    HNR_CHECK_ARG(SR > 0 && D > 0 && R >= 0, "query: bad shape");
    HNR_CHECK_ARG(R * SR * K < (1ll << 31), "query: R*SR*K must be < 2^31 (chunk the rays)");
    if (R == 0) { HNR_CUDA(cudaMemsetAsync(counts_out, 0, 2 * sizeof(int32_t), st)); return HNR_OK; }
    ray_select_kernel<<<(unsigned)hnr_cdiv(R * 32, 256), 256, 0, st>>>(campos, raydir, ts, ts_stride, R, (int)D, (int)SR, *g, occ_bits,
                                                                      sample_loc_full, nsamp);
    static const bool knn_per_voxel = getenv("HNR_KNN_PER_VOXEL") != nullptr;      // A/B switch: the first-generation voxel-by-voxel walk
    static const bool knn_v2 = getenv("HNR_KNN_V2") != nullptr;                    // A/B switch: second-generation batched walk
    if (!knn_v2 && !knn_per_voxel && g->layers <= 2)                                // third generation (compile-time shells, <= 2 of them)
        knn_kernel_v3<<<(unsigned)R, 128, 0, st>>>(sample_loc_full, nsamp, R, (int)SR, (int)K, *g, cell_start, (const float4*)pts_sorted,
                                                   pidx_full, nvalid);
    else if (knn_per_voxel)
        knn_kernel<false><<<(unsigned)R, 128, 0, st>>>(sample_loc_full, nsamp, R, (int)SR, (int)K, *g, cell_start, (const float4*)pts_sorted,
                                                       pidx_full, nvalid);
    else
        knn_kernel<true><<<(unsigned)R, 128, 0, st>>>(sample_loc_full, nsamp, R, (int)SR, (int)K, *g, cell_start, (const float4*)pts_sorted,
                                                      pidx_full, nvalid);
    ray_flags_kernel<<<(unsigned)hnr_cdiv(R, 256), 256, 0, st>>>(nvalid, R, keep);
    launch_scan(keep, ray_off, R, scan_scratch, st);
    launch_scan(nvalid, val_off, R, scan_scratch, st);
    ray_compact_kernel<<<(unsigned)R, 128, 0, st>>>(sample_loc_full, pidx_full, nsamp, nvalid, ray_off, val_off, raydir, campos, camrot, R,
                                                   (int)SR, (int)K, out_pidx, out_loc_pers, out_loc_w, out_dirs, ray_mask, ray_ids,
                                                   vlist, counts_out);
    HNR_CHECK_LAUNCH("query");
    return HNR_OK;
}
