// Fused Greedy Box Seeker stages for sm_100a (B200).  See include/fnp.h for the C ABI and
// DESIGN.md for the data layout and the roofline of each kernel.
//
// What these kernels replace (reference: pcdet/models/dense_heads/frustum_proposals_v1.py):
//   stage 1   :590-613,:812-815   project_to_camera x6 + boolean-mask compaction per 2D box
//   stage 1b  :616-662,:817-845   torch.quantile x3, get_cam_frustum, unprojection, clamp,
//                                 centre line
//   stage 2a  :851-911,:1392-1411 hypothesis grid, softmin front shift, distance/IoU filters
//   stage 2b  :930-932            one points_in_boxes_gpu launch + sum + D2H per hypothesis
//   stage 3   :994-1053           density+IoU score, sort, nms_normal(thresh 1), top-1
#include "fnp_common.cuh"

namespace fnp {

constexpr int kCullThreads = 256;                       // one point per thread per sub-tile
constexpr int kCullSub = FNP_CULL_TILE / kCullThreads;  // sub-tiles of one CTA tile
constexpr int kCullWarps = kCullThreads / 32;
constexpr int kCullVW = kCullSub * kCullWarps;          // "virtual warps" of a tile, in row order
constexpr int kStatsFloats = 40;
static_assert(FNP_CULL_TILE % kCullThreads == 0, "tile must be a whole number of sub-tiles");

// ======================================================================================
// Stage 1: projection + frustum cull + ordered compaction
//   cell_table_kernel:  per (frame, camera rank) a grid of 64-px image cells, each holding the
//       bitmask of the rank's candidates whose 2D box touches the cell (conservative);
//   cull_stage_kernel:  reads every point ONCE.  Per camera a division-free "certainly off
//       this image" test on packed point pairs, the reference's exact IEEE u, v only for the
//       survivors, one cell lookup, exact box tests for the few bits set there.  Membership
//       stays in registers; ballot/popc give per-(warp, candidate) populations, one atomicAdd
//       per tile reserves the tile's slice of a staging buffer, and member points are written
//       there as (x, y, z, depth) of the *unprojected* point, candidate-major, in input order;
//   scans (scan_tiles_kernel, scan_cands_kernel): exclusive prefixes of the per-tile
//       populations -> where every tile's slice lands inside every frustum;
//   cull_gather_kernel: copies the slices to their final, input-ordered position in the
//       pair-interleaved frustum buffer (deterministic, whatever order the tiles ran in).
// ======================================================================================
// Frustum points are stored pair-interleaved: points 2p and 2p+1 of the buffer share one 32-byte
// record {x0,x1, y0,y1, z0,z1, d0,d1}, so that the scoring kernel reads (x0,x1) / (y0,y1) /
// (z0,z1) as the 64-bit operands of Blackwell's packed fp32x2 instructions.  Every frustum
// starts at an even point index.  pair_slot(i) = float offset of x of point i.
__device__ __forceinline__ size_t pair_slot(int64_t i) { return (size_t)(i >> 1) * 8 + (size_t)(i & 1); }

__device__ __constant__ int kImageOrder[6] = {2, 0, 1, 5, 3, 4};   // frustum_proposals_v1.py:201

constexpr int kCellPx = 64;          // cell edge of the candidate lookup grid, pixels
constexpr float kCellInv = 1.0f / kCellPx;

__host__ __device__ inline int cell_cols(float img_w) { return (int)((img_w + kCellPx - 1) / kCellPx); }
__host__ __device__ inline int cell_rows(float img_h) { return (int)((img_h + kCellPx - 1) / kCellPx); }

struct alignas(16) CullSmem {
    float cam[6][24];       // by camera index
    int cs[8];              // candidate range per camera RANK, local to the frame: [cs[r], cs[r+1])
    int tile_base;          // first staging slot of this tile
    int tile_total;
};
static_assert(sizeof(CullSmem) % 16 == 0, "the float4 box table follows this struct in shared memory");

// bits [lo, hi) of word w (bit j of word w = candidate 32 w + j)
__device__ __forceinline__ unsigned range_bits(int lo, int hi, int w)
{
    const int a = min(max(lo - 32 * w, 0), 32), b = min(max(hi - 32 * w, 0), 32);
    const unsigned below_b = (b >= 32) ? 0xffffffffu : ((1u << b) - 1u);
    const unsigned below_a = (a >= 32) ? 0xffffffffu : ((1u << a) - 1u);
    return below_b & ~below_a;
}

// One CTA per (frame, camera rank): cell -> candidates of that rank whose box may contain a
// pixel of the cell.  Conservative (a superset); the exact test follows in cull_stage_kernel.
template <int W>
__global__ void __launch_bounds__(128) cell_table_kernel(const fnp_seeker_batch b, const int n_cu, const int n_cv)
{
    extern __shared__ unsigned s_cells[];                 // [n_cu * n_cv][W]
    const int frame = blockIdx.x / 6, r = blockIdx.x % 6;
    const int n_cells = n_cu * n_cv;
    const int c0 = b.frame_cand_start[frame];
    const int lo = b.cam_cand_start[frame * 6 + r] - c0, hi = b.cam_cand_start[frame * 6 + r + 1] - c0;
    unsigned *out = b.cell_masks + ((size_t)frame * 6 + r) * n_cells * W;
    for (int i = threadIdx.x; i < n_cells * W; i += blockDim.x) s_cells[i] = 0u;
    __syncthreads();
    for (int j = lo + threadIdx.x; j < hi; j += blockDim.x) {
        const float4 bx = reinterpret_cast<const float4 *>(b.cand_box2d)[c0 + j];
        if (!(bx.z > bx.x) || !(bx.w > bx.y)) continue;   // empty (or NaN) box: no point can match
        // cell cu covers u in [64 cu, 64 cu + 64); a member has x1 <= u < x2
        const int cu0 = max(0, (int)floorf(fmaxf(bx.x, 0.f) * kCellInv));
        const int cu1 = min(n_cu - 1, (int)floorf(fminf(bx.z, 65536.f) * kCellInv));
        const int cv0 = max(0, (int)floorf(fmaxf(bx.y, 0.f) * kCellInv));
        const int cv1 = min(n_cv - 1, (int)floorf(fminf(bx.w, 65536.f) * kCellInv));
        for (int cv = cv0; cv <= cv1; cv++)
            for (int cu = cu0; cu <= cu1; cu++) atomicOr(&s_cells[(cv * n_cu + cu) * W + (j >> 5)], 1u << (j & 31));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_cells * W; i += blockDim.x) out[i] = s_cells[i];
}

// wx, wy, wz of two points against one camera: the three rows of lidar2image, each
// fma(a2, z, fma(a1, y, a0 * x)) + a3 as in project(), evaluated on packed pairs (per-lane IEEE).
__device__ __forceinline__ unsigned long long pack2(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long row2(const float *__restrict__ a, unsigned long long x2,
                                                   unsigned long long y2, unsigned long long z2)
{
    unsigned long long t;
    const unsigned long long a0 = pack2(a[0], a[0]), a1 = pack2(a[1], a[1]), a2 = pack2(a[2], a[2]), a3 = pack2(a[3], a[3]);
    asm("{\n .reg .b64 t;\n mul.rn.f32x2 t, %1, %4;\n fma.rn.f32x2 t, %2, %5, t;\n fma.rn.f32x2 t, %3, %6, t;\n"
        " add.rn.f32x2 %0, t, %7;\n}\n"
        : "=l"(t)
        : "l"(a0), "l"(a1), "l"(a2), "l"(x2), "l"(y2), "l"(z2), "l"(a3));
    return t;
}

constexpr int kPtsPerThread = kCullSub;   // a thread owns row (sub * kCullThreads + tid) of every sub-tile

template <int W>
__global__ void __launch_bounds__(kCullThreads, 5) cull_stage_kernel(const fnp_seeker_batch b, const float img_w,
                                                                  const float img_h, const int n_cu, const int n_cv)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CullSmem &S = *reinterpret_cast<CullSmem *>(smem_raw);
    const int Cmax = b.max_cands_per_frame;
    float4 *s_box = reinterpret_cast<float4 *>(smem_raw + sizeof(CullSmem));            // [Cmax]
    int *s_cnt = reinterpret_cast<int *>(s_box + Cmax);                                  // [kCullVW][Cmax]
    int *s_off = s_cnt + kCullVW * Cmax;                                                 // [Cmax] slice offset of a candidate
    unsigned *s_rm = reinterpret_cast<unsigned *>(s_off + Cmax);                         // [6][W] rank bit ranges

    const int tile = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frame = b.tile_frame[tile];
    const int row0 = b.tile_row0[tile];
    const int64_t frow = b.frame_row_start[frame];
    const int frame_rows = (int)(b.frame_row_start[frame + 1] - frow);
    const int c0 = b.frame_cand_start[frame];
    const int nc = b.frame_cand_start[frame + 1] - c0;
    if (nc == 0) return;

    // ---- my points (issued first: the loads overlap the per-CTA setup)
    const int stride = b.point_stride;
    float x[kPtsPerThread], y[kPtsPerThread], z[kPtsPerThread];
    bool live[kPtsPerThread];
#pragma unroll
    for (int s = 0; s < kPtsPerThread; s++) {
        const int row = row0 + s * kCullThreads + tid;
        live[s] = row < frame_rows;
        x[s] = y[s] = z[s] = 0.f;
        if (live[s]) {
            const float *p = b.points + (size_t)(frow + row) * stride + b.xyz_offset;
            x[s] = __ldg(p); y[s] = __ldg(p + 1); z[s] = __ldg(p + 2);
        }
    }

    // ---- per-CTA setup
    for (int i = tid; i < 6 * 24; i += kCullThreads) S.cam[i / 24][i % 24] = b.cam_mats[(size_t)frame * 144 + i];
    for (int j = tid; j < nc; j += kCullThreads) s_box[j] = reinterpret_cast<const float4 *>(b.cand_box2d)[c0 + j];
    for (int i = tid; i < kCullVW * Cmax; i += kCullThreads) s_cnt[i] = 0;
    if (tid < 7) S.cs[tid] = b.cam_cand_start[frame * 6 + tid] - c0;
    __syncthreads();
    if (tid < 6 * W) s_rm[tid] = range_bits(S.cs[tid / W], S.cs[tid / W + 1], tid % W);

    unsigned mask[kPtsPerThread][W];
#pragma unroll
    for (int s = 0; s < kPtsPerThread; s++)
#pragma unroll
        for (int w = 0; w < W; w++) mask[s][w] = 0u;

    // conservative off-image bounds (see the exactness note in DESIGN.md, stage 1)
    const float w_hi = __fmul_rn(img_w, 1.0001f), h_hi = __fmul_rn(img_h, 1.0001f);
    const int n_cells = n_cu * n_cv;

    // ---- membership
#pragma unroll 1
    for (int r = 0; r < 6; r++) {
        if (S.cs[r] == S.cs[r + 1]) continue;                       // camera without candidates
        const float *L = S.cam[kImageOrder[r]];
        float wx[kPtsPerThread], wy[kPtsPerThread], d[kPtsPerThread];
        bool maybe[kPtsPerThread];
        bool any_maybe = false;
#pragma unroll
        for (int s = 0; s < kPtsPerThread; s += 2) {
            const unsigned long long x2 = pack2(x[s], x[s + 1]), y2 = pack2(y[s], y[s + 1]), z2 = pack2(z[s], z[s + 1]);
            float wz0, wz1;
            unpack2(row2(L + 0, x2, y2, z2), wx[s], wx[s + 1]);
            unpack2(row2(L + 4, x2, y2, z2), wy[s], wy[s + 1]);
            unpack2(row2(L + 8, x2, y2, z2), wz0, wz1);
            d[s] = fminf(fmaxf(wz0, 1e-5f), 1e5f);
            d[s + 1] = fminf(fmaxf(wz1, 1e-5f), 1e5f);
        }
#pragma unroll
        for (int s = 0; s < kPtsPerThread; s++) {
            // cheap, division-free "certainly off this image" test
            const bool off = (wx[s] < -1e-30f) | (wy[s] < -1e-30f) | (wx[s] > __fmul_rn(w_hi, d[s])) |
                             (wy[s] > __fmul_rn(h_hi, d[s]));
            maybe[s] = live[s] & !off;
            any_maybe |= maybe[s];
        }
        if (!__any_sync(0xffffffffu, any_maybe)) continue;
        const unsigned *cells = b.cell_masks + ((size_t)frame * 6 + r) * n_cells * W;
#pragma unroll
        for (int s = 0; s < kPtsPerThread; s++) {
            if (!maybe[s]) continue;
            // exact path: the reference's u, v (IEEE division) and on-image / in-box tests
            const float u = __fdiv_rn(wx[s], d[s]), v = __fdiv_rn(wy[s], d[s]);
            if (!((v < img_h) & (v >= 0.f) & (u < img_w) & (u >= 0.f))) continue;
            const int cell = min((int)(v * kCellInv), n_cv - 1) * n_cu + min((int)(u * kCellInv), n_cu - 1);
            const unsigned *cm = cells + (size_t)cell * W;
#pragma unroll
            for (int w = 0; w < W; w++) {
                unsigned m = __ldg(cm + w);
                while (m) {
                    const int jb = __ffs(m) - 1;
                    m &= m - 1;
                    const float4 bx = s_box[32 * w + jb];
                    const bool in = (v < bx.w) & (v >= bx.y) & (u < bx.z) & (u >= bx.x);
                    mask[s][w] |= (in ? 1u : 0u) << jb;
                }
            }
        }
    }

    // ---- populations per (virtual warp, candidate)
#pragma unroll
    for (int s = 0; s < kPtsPerThread; s++) {
#pragma unroll
        for (int w = 0; w < W; w++) {
            unsigned any = __reduce_or_sync(0xffffffffu, mask[s][w]);
            while (any) {
                const int j = __ffs(any) - 1;
                any &= any - 1;
                const unsigned m = __ballot_sync(0xffffffffu, (mask[s][w] >> j) & 1u);
                if (lane == 0) s_cnt[(s * kCullWarps + warp) * Cmax + 32 * w + j] = __popc(m);
            }
        }
    }
    __syncthreads();
    // exclusive prefix over the virtual warps of every candidate; the tile's population of it
    for (int j = tid; j < nc; j += kCullThreads) {
        int run = 0;
#pragma unroll
        for (int vw = 0; vw < kCullVW; vw++) {
            const int c = s_cnt[vw * Cmax + j];
            s_cnt[vw * Cmax + j] = run;
            run += c;
        }
        b.tile_counts[(size_t)tile * Cmax + j] = run;
        s_off[j] = run;
    }
    __syncthreads();
    // exclusive prefix over candidates (one warp), then reserve the tile's staging slice
    if (warp == 0) {
        int carry = 0;
        for (int j0 = 0; j0 < nc; j0 += 32) {
            const int j = j0 + lane;
            const int val = (j < nc) ? s_off[j] : 0;
            int inc = val;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += n;
            }
            if (j < nc) s_off[j] = carry + inc - val;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) {
            S.tile_total = carry;
            S.tile_base = carry ? atomicAdd(&b.status[5], carry) : 0;
            b.tile_base[tile] = S.tile_base;
        }
    }
    __syncthreads();
    const int total = S.tile_total;
    if (total == 0) return;
    const int64_t base = S.tile_base;
    if (base + total > b.pts_capacity) return;      // overflow: scan_cands_kernel raises the flag

    // ---- re-project members into their camera, unproject, ordered write into the slice
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int s = 0; s < kPtsPerThread; s++) {
        unsigned sub_any = 0u;
#pragma unroll
        for (int w = 0; w < W; w++) sub_any |= mask[s][w];
        if (!__any_sync(0xffffffffu, sub_any != 0u)) continue;
        const int row = row0 + s * kCullThreads + tid;
        const int *cnt_vw = s_cnt + (s * kCullWarps + warp) * Cmax;
#pragma unroll 1
        for (int r = 0; r < 6; r++) {
            unsigned rm[W];
            unsigned has = 0u;
#pragma unroll
            for (int w = 0; w < W; w++) { rm[w] = mask[s][w] & s_rm[r * W + w]; has |= rm[w]; }
            if (!__any_sync(0xffffffffu, has != 0u)) continue;
            const float *cm = S.cam[kImageOrder[r]];
            float u, v, dd, X = 0.f, Y = 0.f, Z = 0.f;
            project(cm, x[s], y[s], z[s], img_w, img_h, u, v, dd);
            unproject(cm + 12, cm + 21, u, v, dd, X, Y, Z);
#pragma unroll
            for (int w = 0; w < W; w++) {
                unsigned any = __reduce_or_sync(0xffffffffu, rm[w]);
                while (any) {
                    const int jb = __ffs(any) - 1;
                    any &= any - 1;
                    const bool in = (rm[w] >> jb) & 1u;
                    const unsigned m = __ballot_sync(0xffffffffu, in);
                    if (in) {
                        const int j = 32 * w + jb;
                        const int64_t pos = base + s_off[j] + cnt_vw[j] + __popc(m & lt);
                        reinterpret_cast<float4 *>(b.stage_pts)[pos] = make_float4(X, Y, Z, dd);
                        if (b.stage_idx) b.stage_idx[pos] = row;
                    }
                }
            }
        }
    }
}

// One CTA per tile: move the tile's staging slice to its final place in every frustum.
__global__ void __launch_bounds__(256) cull_gather_kernel(const fnp_seeker_batch b)
{
    extern __shared__ int s_o[];                           // [Cmax + 1] slice offsets, then [Cmax] destinations
    const int Cmax = b.max_cands_per_frame;
    int *s_dst = s_o + Cmax + 1;
    __shared__ int s_total;
    const int tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int frame = b.tile_frame[tile];
    const int c0 = b.frame_cand_start[frame];
    const int nc = b.frame_cand_start[frame + 1] - c0;
    if (nc == 0 || b.status[0] != 0) return;
    if (tid < 32) {
        int carry = 0;
        for (int j0 = 0; j0 < nc; j0 += 32) {
            const int j = j0 + lane;
            const int val = (j < nc) ? b.tile_counts[(size_t)tile * Cmax + j] : 0;
            int inc = val;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += n;
            }
            if (j < nc) {
                s_o[j] = carry + inc - val;
                s_dst[j] = b.cand_pt_start[c0 + j] + b.tile_dst[(size_t)tile * Cmax + j];
            }
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) { s_o[nc] = carry; s_total = carry; }
    }
    __syncthreads();
    const int total = s_total;
    if (total == 0) return;
    const int64_t base = b.tile_base[tile];
    for (int k = tid; k < total; k += blockDim.x) {
        // candidate of slot k: the last j with s_o[j] <= k
        int lo = 0, hi = nc;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (s_o[mid] <= k) lo = mid; else hi = mid;
        }
        const int64_t pos = (int64_t)s_dst[lo] + (k - s_o[lo]);
        const float4 rec = reinterpret_cast<const float4 *>(b.stage_pts)[base + k];
        float *dst = b.frustum_pts + pair_slot(pos);
        dst[0] = rec.x; dst[2] = rec.y; dst[4] = rec.z; dst[6] = rec.w;
        if (b.frustum_idx) b.frustum_idx[pos] = b.stage_idx[base + k];
    }
}

// exclusive prefix of one candidate's tile counts (one warp per candidate)
__global__ void __launch_bounds__(128) scan_tiles_kernel(const fnp_seeker_batch b)
{
    const int f = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (f >= b.n_cands) return;
    const int frame = b.cand_frame[f];
    const int j = f - b.frame_cand_start[frame];
    const int t0 = b.frame_tile_start[frame], t1 = b.frame_tile_start[frame + 1];
    int carry = 0;
    for (int t = t0; t < t1; t += 32) {
        const int i = t + lane;
        const size_t cell = (size_t)i * b.max_cands_per_frame + j;
        const int val = (i < t1) ? b.tile_counts[cell] : 0;
        int inc = val;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += n;
        }
        if (i < t1) b.tile_dst[cell] = carry + inc - val;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) b.cand_npts[f] = carry;
}

// exclusive prefix over candidates -> cand_pt_start, capacity check
__global__ void __launch_bounds__(1024) scan_cands_kernel(const fnp_seeker_batch b)
{
    __shared__ int s_warp[32];
    __shared__ long long s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < b.n_cands; base += 1024) {
        const int i = base + tid;
        const int val = (i < b.n_cands) ? ((b.cand_npts[i] + 1) & ~1) : 0;   // frustums start on a pair boundary
        int inc = val;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += n;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += n;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const long long carry = s_carry;
        const long long excl = carry + (warp ? s_warp[warp - 1] : 0) + inc - val;
        if (i < b.n_cands) b.cand_pt_start[i] = (int)min(excl, (long long)0x7fffffff);
        __syncthreads();
        if (tid == 1023) s_carry = carry + s_warp[31];
        __syncthreads();
    }
    if (tid == 0) {
        const long long total = s_carry;
        b.cand_pt_start[b.n_cands] = (int)min(total, (long long)0x7fffffff);
        b.status[0] = (total > b.pts_capacity) ? 1 : 0;
        b.status[1] = (int)min(total, (long long)0x7fffffff);
    }
}

// ======================================================================================
// Stage 1b: per-frustum statistics and centre line
// ======================================================================================
__device__ __forceinline__ float warp_min(float v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// k-th (0-based) and (k+1)-th smallest depth of a frustum, exactly.  Depths are >= 1e-5 > 0, so
// their bit patterns order like the values and everything below is integer arithmetic on
// rel = key - kmin.  Range-normalised MSD radix select: the first 11-bit digit already spreads the
// keys over [kmin, kmax] (2048 bins), a bin with few enough keys is finished by rank counting in
// shared memory, a crowded one is refined by the next 11 bits.  Typical frustum: one histogram
// pass + one collect pass, both from the shared-memory key cache when the frustum fits.
// All threads of the block call this; results are block-uniform.
constexpr int kSelBits = 11, kSelBins = 1 << kSelBits;
constexpr int kSelList = 512;      // keys finished by rank counting
constexpr int kStatsCache = 4096;  // depth keys cached in shared memory
constexpr int kStatsThreads = 256;

__device__ __forceinline__ float pt_depth(const float *__restrict__ pts, int i)
{
    return pts[(size_t)(i >> 1) * 8 + 6 + (i & 1)];
}

struct SelSmem {
    unsigned hist[kSelBins];
    unsigned list[kSelList];
    unsigned key[kStatsCache];
    unsigned warp_sum[kStatsThreads / 32];
    unsigned misc[8];   // 0 bin, 1 keys below it, 2 keys in it, 3 next non-empty bin, 4 list fill, 5 min key above, 6/7 results
};

// Calls f(key) for every depth key of the frustum, block-strided.  Uncached frustums (the few
// large ones, which set the duration of the whole kernel) are read as pair records with four
// independent 16-byte loads in flight per thread.
template <typename F>
__device__ __forceinline__ void for_each_key(const float *__restrict__ pts, const SelSmem &S, bool cached, int n, F f)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    if (cached) {
        for (int i = tid; i < n; i += nt) f(S.key[i]);
        return;
    }
    const float4 *rec = reinterpret_cast<const float4 *>(pts);
    const int np = (n + 1) >> 1;                         // pair records
    for (int p0 = tid; p0 < np; p0 += 4 * nt) {
        float4 c[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int p = p0 + u * nt;
            if (p < np) c[u] = __ldg(rec + 2 * p + 1);   // z0 z1 d0 d1
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int p = p0 + u * nt;
            if (p < np) {
                f(__float_as_uint(c[u].z));
                if (2 * p + 1 < n) f(__float_as_uint(c[u].w));
            }
        }
    }
}

__device__ void select_pair(const float *__restrict__ pts, SelSmem &S, bool cached, int n, int k, unsigned kmin,
                            unsigned kmax, float &v_lo, float &v_hi)
{
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    if (kmin == kmax) { v_lo = v_hi = __uint_as_float(kmin); return; }
    int shift = max(0, (32 - __clz(kmax - kmin)) - kSelBits);   // digit = (rel >> shift) & 2047; level 1: rel >> shift < 2048
    int pshift = 32;          // current set: (rel >> pshift) == (lo_rel >> pshift); 32 = every key
    unsigned lo_rel = 0;
    int rank = k;             // rank of the wanted key inside the current set
    // keys above the current set: the nearest non-empty bin seen so far (nb_*), whose minimum is only
    // needed if the (k+1)-th key is not in the final bin; it is evaluated lazily in a later pass
    bool nb_valid = false;
    unsigned nb_lo = 0;
    int nb_shift = 0;
    unsigned above = 0xffffffffu;
    for (;;) {
        for (int i = tid; i < kSelBins; i += nt) S.hist[i] = 0;
        if (tid == 0) { S.misc[3] = kSelBins; S.misc[4] = 0; S.misc[5] = 0xffffffffu; }
        __syncthreads();
        unsigned my_above = 0xffffffffu;
        for_each_key(pts, S, cached, n, [&](unsigned key) {
            const unsigned rel = key - kmin;
            if (nb_valid && (rel >> nb_shift) == (nb_lo >> nb_shift)) my_above = min(my_above, key);
            if (pshift < 32 && (rel >> pshift) != (lo_rel >> pshift)) return;
            atomicAdd(&S.hist[(rel >> shift) & (kSelBins - 1)], 1u);
        });
        if (nb_valid) {
            my_above = __reduce_min_sync(0xffffffffu, my_above);
            if (lane == 0 && my_above != 0xffffffffu) atomicMin(&S.misc[5], my_above);
        }
        __syncthreads();
        if (nb_valid) { above = min(above, S.misc[5]); nb_valid = false; }
        // ---- locate the bin holding `rank`: 8 bins per thread, block scan of the 256 partial sums
        unsigned c[kSelBins / kStatsThreads], local = 0;
#pragma unroll
        for (int j = 0; j < kSelBins / kStatsThreads; j++) { c[j] = S.hist[tid * (kSelBins / kStatsThreads) + j]; local += c[j]; }
        unsigned inc = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane == 31) S.warp_sum[warp] = inc;
        __syncthreads();
        unsigned excl = inc - local;
        for (int w = 0; w < warp; w++) excl += S.warp_sum[w];
        if ((unsigned)rank >= excl && (unsigned)rank < excl + local) {
            unsigned acc = excl;
#pragma unroll
            for (int j = 0; j < kSelBins / kStatsThreads; j++) {
                if ((unsigned)rank >= acc && (unsigned)rank < acc + c[j]) {
                    S.misc[0] = tid * (kSelBins / kStatsThreads) + j;
                    S.misc[1] = acc;
                    S.misc[2] = c[j];
                }
                acc += c[j];
            }
        }
        __syncthreads();
        const unsigned B = S.misc[0], below = S.misc[1], cB = S.misc[2];
#pragma unroll
        for (int j = 0; j < kSelBins / kStatsThreads; j++) {
            const unsigned bin = tid * (kSelBins / kStatsThreads) + j;
            if (bin > B && c[j]) atomicMin(&S.misc[3], bin);
        }
        __syncthreads();
        const unsigned Bn = S.misc[3];
        rank -= (int)below;
        const unsigned hi_part = (pshift < 32) ? ((lo_rel >> pshift) << pshift) : 0u;
        const unsigned bin_lo = hi_part | (B << shift);
        const bool has_next = Bn < (unsigned)kSelBins;
        const unsigned next_lo = hi_part | (Bn << shift);
        const bool second_in_bin = (unsigned)(rank + 1) < cB;

        if (shift == 0) {                         // a bin is one exact key
            v_lo = __uint_as_float(kmin + bin_lo);
            if (second_in_bin) v_hi = v_lo;
            else if (has_next) v_hi = __uint_as_float(kmin + next_lo);
            else v_hi = (above != 0xffffffffu) ? __uint_as_float(above) : v_lo;
            __syncthreads();
            return;
        }
        if (cB <= (unsigned)kSelList) {           // finish: collect the bin, rank by counting
            unsigned nmin = 0xffffffffu;
            for_each_key(pts, S, cached, n, [&](unsigned key) {
                const unsigned rel = key - kmin;
                if (pshift < 32 && (rel >> pshift) != (lo_rel >> pshift)) return;
                const unsigned digit = (rel >> shift) & (kSelBins - 1);
                if (digit == B) S.list[atomicAdd(&S.misc[4], 1u)] = key;
                else if (!second_in_bin && has_next && digit == Bn) nmin = min(nmin, key);
            });
            if (!second_in_bin && has_next) {
                nmin = __reduce_min_sync(0xffffffffu, nmin);
                if (lane == 0 && nmin != 0xffffffffu) atomicMin(&S.misc[5], nmin);
            }
            __syncthreads();
            const int m = (int)cB;
            for (int e = tid; e < m; e += nt) {
                const unsigned key = S.list[e];
                int lt = 0, le = 0;
                for (int j = 0; j < m; j++) { const unsigned o = S.list[j]; lt += o < key; le += o <= key; }
                if (lt <= rank && rank < le) S.misc[6] = key;               // same value from every writer
                if (second_in_bin && lt <= rank + 1 && rank + 1 < le) S.misc[7] = key;
            }
            __syncthreads();
            v_lo = __uint_as_float(S.misc[6]);
            if (second_in_bin) v_hi = __uint_as_float(S.misc[7]);
            else if (has_next) v_hi = __uint_as_float(S.misc[5]);
            else v_hi = (above != 0xffffffffu) ? __uint_as_float(above) : v_lo;
            __syncthreads();
            return;
        }
        // ---- refine the crowded bin with the next digit
        if (has_next) { nb_valid = true; nb_lo = next_lo; nb_shift = shift; }
        lo_rel = bin_lo;
        pshift = shift;
        shift = max(0, shift - kSelBits);
        __syncthreads();
    }
}

// torch.quantile(depth, q), linear interpolation (ATen Sorting.cpp quantile_compute + lerp)
__device__ float block_quantile(const float *__restrict__ pts, SelSmem &S, bool cached, int n, float q, float dmin,
                                float dmax)
{
    const float pos = __fmul_rn(q, (float)(n - 1));
    const float lo = floorf(pos), hi = ceilf(pos);
    const float w = __fsub_rn(pos, lo);
    const int klo = (int)lo, khi = (int)hi;
    float a, bv;
    if (khi == 0) { a = dmin; bv = dmin; }
    else if (klo == n - 1) { a = dmax; bv = dmax; }
    else {
        select_pair(pts, S, cached, n, klo, __float_as_uint(dmin), __float_as_uint(dmax), a, bv);
        if (khi == klo) bv = a;
    }
    const float diff = __fsub_rn(bv, a);
    return (w < 0.5f) ? __fmaf_rn(w, diff, a) : __fmaf_rn(-diff, __fsub_rn(1.0f, w), bv);
}

__global__ void __launch_bounds__(kStatsThreads) stats_kernel(const fnp_seeker_batch b, const fnp_seeker_cfg cfg)
{
    __shared__ SelSmem S;
    __shared__ float s_red[8][8];
    __shared__ float s_geo[16];  // close[3], vec[3]
    const int f = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = b.cand_npts[f];
    float *st = b.cand_stats + (size_t)f * kStatsFloats;
    if (n <= 0 || b.status[0] != 0) {
        if (tid == 0) { st[9] = 0.f; }
        return;
    }
    const float *pts = b.frustum_pts + (size_t)(b.cand_pt_start[f] >> 1) * 8;   // starts are even

    // ---- min / max of depth and of x, y, z: one 32-byte pair record per thread and iteration
    const float INF = __int_as_float(0x7f800000);
    float mn[4] = {INF, INF, INF, INF}, mx[4] = {-INF, -INF, -INF, -INF};
    const float4 *rec = reinterpret_cast<const float4 *>(pts);
    const bool cached = n <= kStatsCache;                         // depth keys stay in shared memory
    const int np = (n + 1) >> 1;
    for (int p0 = tid; p0 < np; p0 += 4 * kStatsThreads) {        // four records (8 loads) in flight per thread
        float4 ra[4], rc[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int p = p0 + u * kStatsThreads;
            if (p < np) { ra[u] = rec[2 * p]; rc[u] = rec[2 * p + 1]; }   // x0 x1 y0 y1 | z0 z1 d0 d1
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int p = p0 + u * kStatsThreads;
            if (p >= np) continue;
            const float4 a = ra[u], c = rc[u];
            if (cached) {
                S.key[2 * p] = __float_as_uint(c.z);
                if (2 * p + 1 < n) S.key[2 * p + 1] = __float_as_uint(c.w);
            }
            mn[0] = fminf(mn[0], a.x); mx[0] = fmaxf(mx[0], a.x);
            mn[1] = fminf(mn[1], a.z); mx[1] = fmaxf(mx[1], a.z);
            mn[2] = fminf(mn[2], c.x); mx[2] = fmaxf(mx[2], c.x);
            mn[3] = fminf(mn[3], c.z); mx[3] = fmaxf(mx[3], c.z);
            if (2 * p + 1 < n) {
                mn[0] = fminf(mn[0], a.y); mx[0] = fmaxf(mx[0], a.y);
                mn[1] = fminf(mn[1], a.w); mx[1] = fmaxf(mx[1], a.w);
                mn[2] = fminf(mn[2], c.y); mx[2] = fmaxf(mx[2], c.y);
                mn[3] = fminf(mn[3], c.w); mx[3] = fmaxf(mx[3], c.w);
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 4; a++) {
        mn[a] = warp_min(mn[a]);
        mx[a] = warp_max(mx[a]);
        if (lane == 0) { s_red[warp][a] = mn[a]; s_red[warp][4 + a] = mx[a]; }
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 4; a++) {
        float m0 = INF, m1 = -INF;
        for (int w = 0; w < 8; w++) { m0 = fminf(m0, s_red[w][a]); m1 = fmaxf(m1, s_red[w][4 + a]); }
        mn[a] = m0; mx[a] = m1;
    }
    __syncthreads();

    // ---- depth quantiles (frustum_proposals_v1.py:616-648)
    const float qmin = block_quantile(pts, S, cached, n, cfg.lq, mn[3], mx[3]);
    // search_depth (:619-623): the far end of the frustum is the near quantile + depth
    const float qmax = (cfg.search_depth > 0.f) ? __fadd_rn(qmin, cfg.search_depth)
                                                : block_quantile(pts, S, cached, n, cfg.uq, mn[3], mx[3]);
    const float qc = block_quantile(pts, S, cached, n, cfg.cq, mn[3], mx[3]);
    const float dmax = fminf(qmax, cfg.max_dist);
    const float dmin = fmaxf(qmin, cfg.frustum_min);

    if (tid == 0) {
        const int frame = b.cand_frame[f];
        const float *cm = b.cam_mats + ((size_t)frame * 6 + b.cand_cam[f]) * 24;
        const float *bx = b.cand_box2d + (size_t)f * 4;
        const float lo[3] = {bx[0], bx[1], dmin}, hi[3] = {bx[2], bx[3], dmax};
        const float tpl[8][3] = {{1, 1, -1}, {1, -1, -1}, {-1, -1, -1}, {-1, 1, -1},
                                 {1, 1, 1}, {1, -1, 1}, {-1, -1, 1}, {-1, 1, 1}};
        float c[8][3];
        for (int k = 0; k < 8; k++) {
            float uvd[3];
            for (int a = 0; a < 3; a++) {
                const float whl = __fsub_rn(hi[a], lo[a]);
                const float cen = __fmul_rn(__fadd_rn(hi[a], lo[a]), 0.5f);
                uvd[a] = __fadd_rn(__fmul_rn(whl, tpl[k][a] * 0.5f), cen);
            }
            unproject(cm + 12, cm + 21, uvd[0], uvd[1], uvd[2], c[k][0], c[k][1], c[k][2]);
        }
        if (cfg.clamp_bottom > 0) {
            for (int a = 0; a < 3; a++) {
                float cmin = c[0][a], cmax = c[0][a];
                for (int k = 1; k < 8; k++) { cmin = fminf(cmin, c[k][a]); cmax = fmaxf(cmax, c[k][a]); }
                const float f1 = fmaxf(mn[a], cmin), f2 = fminf(mx[a], cmax);
                for (int k = 0; k < 8; k++) c[k][a] = fminf(fmaxf(c[k][a], f1), f2);
            }
        }
        for (int a = 0; a < 3; a++) {
            float bev[4];
            for (int i = 0; i < 4; i++) bev[i] = __fmul_rn(__fadd_rn(c[2 * i][a], c[2 * i + 1][a]), 0.5f);
            const float close = __fmul_rn(__fadd_rn(bev[0], bev[1]), 0.5f);
            const float far = __fmul_rn(__fadd_rn(bev[2], bev[3]), 0.5f);
            s_geo[a] = close;
            s_geo[3 + a] = __fsub_rn(far, close);
        }
        if (cfg.search_depth > 0.f) {   // :841-842  center_vec / center_vec.norm() * search_depth
            const float nv = norm3(s_geo[3], s_geo[4], s_geo[5]);
            for (int a = 0; a < 3; a++) s_geo[3 + a] = __fmul_rn(__fdiv_rn(s_geo[3 + a], nv), cfg.search_depth);
        }
        // weighted_centre_xyz (:631-636): the 2D box centre at the cq depth quantile, unprojected
        unproject(cm + 12, cm + 21, __fmul_rn(__fadd_rn(bx[0], bx[2]), 0.5f), __fmul_rn(__fadd_rn(bx[1], bx[3]), 0.5f), qc,
                  st[10], st[11], st[12]);
        st[0] = dmin; st[1] = dmax; st[2] = qc;
        for (int a = 0; a < 3; a++) { st[3 + a] = mn[a]; st[6 + a] = mx[a]; }
        st[9] = (float)n;
        for (int k = 0; k < 8; k++)
            for (int a = 0; a < 3; a++) st[16 + k * 3 + a] = c[k][a];
    }
    __syncthreads();
    const int M = cfg.num_mags;
    for (int i = tid; i < M * 3; i += blockDim.x) {
        const int m = i / 3, a = i % 3;
        b.centres[((size_t)f * M + m) * 3 + a] = __fadd_rn(s_geo[a], __fmul_rn(s_geo[3 + a], b.mags[m]));
    }
}

// ======================================================================================
// Stage 2a: hypotheses
// ======================================================================================
// 2D IoU of the image-plane bounding box of the 8 shifted corners with a 2D box (calc_iou, :1392-1411)
__device__ __forceinline__ float view_iou(const float *__restrict__ L, const float4 box2d, const float (&cor)[8][3],
                                          const float (&shift)[3], const float img_w, const float img_h)
{
    const float area2 = __fmul_rn(__fsub_rn(box2d.z, box2d.x), __fsub_rn(box2d.w, box2d.y));
    const float INF = __int_as_float(0x7f800000);
    float x1 = INF, y1 = INF, x2 = -INF, y2 = -INF;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        float u, v, d;
        project(L, __fadd_rn(cor[k][0], shift[0]), __fadd_rn(cor[k][1], shift[1]),
                __fadd_rn(cor[k][2], shift[2]), img_w, img_h, u, v, d);
        u = fminf(fmaxf(u, 0.f), img_w);
        v = fminf(fmaxf(v, 0.f), img_h);
        x1 = fminf(x1, u); x2 = fmaxf(x2, u); y1 = fminf(y1, v); y2 = fmaxf(y2, v);
    }
    const float area1 = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
    const float lx = fmaxf(x1, box2d.x), ly = fmaxf(y1, box2d.y);
    const float rx = fminf(x2, box2d.z), ry = fminf(y2, box2d.w);
    const float iw = fmaxf(__fsub_rn(rx, lx), 0.f), ih = fmaxf(__fsub_rn(ry, ly), 0.f);
    const float inter = __fmul_rn(iw, ih);
    const float uni = __fsub_rn(__fadd_rn(area1, area2), inter);
    return __fdiv_rn(inter, uni);
}

// EXTRAS = false is the shipped configuration (single-view IoU, no hyp_dist): the optional terms are
// compiled out so that they cost the hot path no registers.
template <bool EXTRAS>
__global__ void __launch_bounds__(128) hypotheses_kernel(const fnp_seeker_batch b, const fnp_seeker_cfg cfg)
{
    __shared__ int s_wcnt[4];
    __shared__ int s_base;
    __shared__ float s_dmm[4][2];
    const int f = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int J = cfg.num_yaw_size, M = cfg.num_mags, H = J * M;
    const bool empty = (b.cand_npts[f] <= 0) || (b.status[0] != 0);
    if (empty) {
        if (tid == 0) b.hyp_nvalid[f] = 0;
        if (b.hyp_valid_dbg)
            for (int h = tid; h < H; h += blockDim.x) b.hyp_valid_dbg[(size_t)f * H + h] = 0;
        return;
    }
    const int frame = b.cand_frame[f];
    const int label = b.cand_label[f];
    const float *L = b.cam_mats + ((size_t)frame * 6 + b.cand_cam[f]) * 24;
    const float4 box2d = reinterpret_cast<const float4 *>(b.cand_box2d)[f];
    const bool multicam = EXTRAS && (cfg.flags & FNP_SEEKER_MULTICAM_IOU) != 0;
    const int fc0 = b.frame_cand_start[frame], fc1 = b.frame_cand_start[frame + 1];
    const bool want_dist = EXTRAS && b.hyp_dist != nullptr;
    float wc[3] = {0.f, 0.f, 0.f};
    if (want_dist) {
        const float *st = b.cand_stats + (size_t)f * kStatsFloats;
        wc[0] = st[10]; wc[1] = st[11]; wc[2] = st[12];
    }
    float dist_mn = __int_as_float(0x7f800000), dist_mx = -__int_as_float(0x7f800000);
    const float *bb_tab = b.base_boxes + (size_t)(label - 1) * J * 7;
    const float *bc_tab = b.base_corners + (size_t)(label - 1) * J * 24;
    if (tid == 0) s_base = 0;
    __syncthreads();

    for (int h0 = 0; h0 < H; h0 += blockDim.x) {
        const int h = h0 + tid;
        bool valid = false;
        float box[7] = {0, 0, 0, 0, 0, 0, 0};
        float iou = 0.f, dist = 0.f;
        if (h < H) {
            const int m = h / J, j = h - m * J;
            const float *ct = b.centres + ((size_t)f * M + m) * 3;
            const float *bc = bc_tab + (size_t)j * 24;
            const float *bb = bb_tab + (size_t)j * 7;
            const float ctr[3] = {ct[0], ct[1], ct[2]};
            float cor[8][3], e[8];
            float mxn = -__int_as_float(0x7f800000);
#pragma unroll
            for (int k = 0; k < 8; k++) {
#pragma unroll
                for (int a = 0; a < 3; a++) cor[k][a] = __fadd_rn(__ldg(bc + k * 3 + a), ctr[a]);
                e[k] = -norm3(cor[k][0], cor[k][1], cor[k][2]);
                mxn = fmaxf(mxn, e[k]);
            }
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < 8; k++) { e[k] = fnp_exp(__fsub_rn(e[k], mxn)); sum = __fadd_rn(sum, e[k]); }
            float front[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const float w = __fdiv_rn(e[k], sum);
#pragma unroll
                for (int a = 0; a < 3; a++) front[a] = __fadd_rn(front[a], __fmul_rn(w, cor[k][a]));
            }
            float shift[3];
#pragma unroll
            for (int a = 0; a < 3; a++) {
                const float cc = __fadd_rn(__ldg(bb + a), ctr[a]);
                shift[a] = __fsub_rn(cc, front[a]);
                box[a] = __fadd_rn(cc, shift[a]);
            }
#pragma unroll
            for (int a = 3; a < 7; a++) box[a] = __ldg(bb + a);
            const bool near_enough = norm3(front[0], front[1], front[2]) < cfg.max_dist;
            if (!multicam) iou = view_iou(L, box2d, cor, shift, cfg.img_w, cfg.img_h);
            else {
                // multicam_ious (:1413-1429): every candidate of the frame with points and the same label
                // (this one included), summed in candidate order, over (number of non-zero IoUs + 1e-6)
                float sum_iou = 0.f;
                int nz = 0;
                for (int i = fc0; i < fc1; i++) {
                    if (b.cand_label[i] != label || b.cand_npts[i] <= 0) continue;
                    const float vi = view_iou(b.cam_mats + ((size_t)frame * 6 + b.cand_cam[i]) * 24,
                                              reinterpret_cast<const float4 *>(b.cand_box2d)[i], cor, shift,
                                              cfg.img_w, cfg.img_h);
                    sum_iou = __fadd_rn(sum_iou, vi);
                    nz += vi > 0.f;
                }
                iou = __fdiv_rn(sum_iou, __fadd_rn((float)nz, 1e-6f));
            }
            if (want_dist) {   // torch.cdist(front, weighted_centre_xyz) (:889), evaluated directly
                dist = norm3(__fsub_rn(front[0], wc[0]), __fsub_rn(front[1], wc[1]), __fsub_rn(front[2], wc[2]));
                if (near_enough) { dist_mn = fminf(dist_mn, dist); dist_mx = fmaxf(dist_mx, dist); }
            }
            valid = near_enough && (iou > cfg.min_cam_iou);
            if (b.hyp_boxes_dbg) {
                float *o = b.hyp_boxes_dbg + ((size_t)f * H + h) * 7;
#pragma unroll
                for (int a = 0; a < 7; a++) o[a] = box[a];
            }
            if (b.hyp_iou_dbg) b.hyp_iou_dbg[(size_t)f * H + h] = iou;
            if (b.hyp_valid_dbg) b.hyp_valid_dbg[(size_t)f * H + h] = valid ? 1 : 0;
        }
        // ordered compaction of the valid hypotheses
        const unsigned mk = __ballot_sync(0xffffffffu, valid);
        if (lane == 0) s_wcnt[warp] = __popc(mk);
        __syncthreads();
        int base = s_base;
        for (int w = 0; w < warp; w++) base += s_wcnt[w];
        if (valid) {
            const int r = base + __popc(mk & ((1u << lane) - 1u));
            const BoxPrep p = prep_box(box);
            float4 *dst = reinterpret_cast<float4 *>(b.hyp_prep + ((size_t)f * H + r) * 8);
            dst[0] = make_float4(p.cx, p.cy, p.cz, p.hz);
            dst[1] = make_float4(p.cosa, p.sina, p.tx, p.ty);
            b.hyp_index[(size_t)f * H + r] = h;
            b.hyp_iou[(size_t)f * H + r] = iou;
            if (want_dist) b.hyp_dist[(size_t)f * H + r] = dist;
        }
        __syncthreads();
        if (tid == 0) s_base = base + s_wcnt[0] + s_wcnt[1] + s_wcnt[2] + s_wcnt[3];
        __syncthreads();
    }
    if (tid == 0) b.hyp_nvalid[f] = s_base;
    if (want_dist) {   // dists_ranked is normalised over the hypotheses within max_dist (:891)
        dist_mn = warp_min(dist_mn);
        dist_mx = warp_max(dist_mx);
        if (lane == 0) { s_dmm[warp][0] = dist_mn; s_dmm[warp][1] = dist_mx; }
        __syncthreads();
        if (tid == 0) {
            float *st = b.cand_stats + (size_t)f * kStatsFloats;
            st[13] = fminf(fminf(s_dmm[0][0], s_dmm[1][0]), fminf(s_dmm[2][0], s_dmm[3][0]));
            st[14] = fmaxf(fmaxf(s_dmm[0][1], s_dmm[1][1]), fmaxf(s_dmm[2][1], s_dmm[3][1]));
        }
    }
}

// ======================================================================================
// Stage 2b: scoring -- per-hypothesis point counts
// ======================================================================================
constexpr int kScoreThreads = 128;
constexpr int kScoreTile = 512;  // points per TMA stage (8 KB)

constexpr int kScoreKMax = 4;                                   // hypotheses per thread in a full chunk
constexpr int kScoreChunk = kScoreThreads * kScoreKMax;         // hypotheses of a full chunk

// The nv valid hypotheses of a frustum are cut into nv / 512 full chunks (4 per thread) and one
// remainder chunk of ceil(rem / 128) per thread, so that the padding stays below 128 hypotheses
// per frustum (with 512-wide chunks only, the padded work was 1.42x the useful work on cfg2).
__host__ __device__ inline int score_chunks(int nv) { return (nv + kScoreChunk - 1) / kScoreChunk; }

// Work items of the scoring stage.  Frustum f with P_f points and nv_f valid hypotheses is cut
// into S_f = ceil(P_f / split_points) point splits x score_chunks(nv_f) hypothesis
// chunks; every (split, chunk) pair is one CTA-sized item, so the largest frustums no longer
// set the kernel's duration.  A thread keeps its counts in registers for the whole item and adds
// them to row f of `counts` with one integer RED per hypothesis at the end (integer addition is
// associative: the totals do not depend on the order in which the splits finish).
__global__ void __launch_bounds__(1024) plan_items_kernel(const fnp_seeker_batch b, const int H, const int sweep)
{
    __shared__ int s_warp_i[32];
    __shared__ int s_carry_i;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry_i = 0;
    __syncthreads();
    for (int base = 0; base < b.n_cands; base += 1024) {
        const int f = base + tid;
        int items = 0;
        if (f < b.n_cands) {
            const int np = b.cand_npts[f], nv = b.hyp_nvalid[f];
            if (np > 0 && nv > 0) items = ((np + b.split_points - 1) / b.split_points) * (sweep ? 1 : score_chunks(nv));
        }
        int inc_i = items;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int a = __shfl_up_sync(0xffffffffu, inc_i, o);
            if (lane >= o) inc_i += a;
        }
        if (lane == 31) s_warp_i[warp] = inc_i;
        __syncthreads();
        if (warp == 0) {
            int wi = s_warp_i[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int a = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += a;
            }
            s_warp_i[lane] = wi;
        }
        __syncthreads();
        const int ci = s_carry_i;
        if (f < b.n_cands) b.cand_item_start[f] = ci + (warp ? s_warp_i[warp - 1] : 0) + inc_i - items;
        __syncthreads();
        if (tid == 1023) s_carry_i = ci + s_warp_i[31];
        __syncthreads();
    }
    if (tid == 0) {
        b.cand_item_start[b.n_cands] = s_carry_i;
        b.status[2] = s_carry_i;
        b.status[3] = 0;
        b.status[4] = 0;            // work-item counter of the persistent scoring CTAs
        if (s_carry_i > b.max_items) b.status[0] |= 2;
    }
}

__global__ void __launch_bounds__(128) write_items_kernel(const fnp_seeker_batch b, const int H, const int sweep)
{
    const int f = blockIdx.x;
    const int i0 = b.cand_item_start[f], n = b.cand_item_start[f + 1] - i0;
    if (n <= 0 || (b.status[0] & 2)) return;
    const int nv = b.hyp_nvalid[f];
    const int nchunks = sweep ? 1 : score_chunks(nv);   // the sweep kernel takes all hypotheses of a split at once
    for (int i = threadIdx.x; i < n; i += blockDim.x) {   // split-major: neighbours share a point tile
        const int c = i % nchunks;
        const int left = nv - c * kScoreChunk;             // hypotheses from this chunk's base on
        const int K = left >= kScoreChunk ? kScoreKMax : (left + kScoreThreads - 1) / kScoreThreads;
        reinterpret_cast<int4 *>(b.items)[i0 + i] = make_int4(f, c * kScoreChunk, i / nchunks, K);
    }
}

// Packed-fp32x2 form of the in-box predicate for TWO points against one hypothesis.  Same
// arithmetic per lane as in_box() (sub.rn, mul.rn, fma.rn are IEEE per lane):
//   sx = x - cx, sy = y - cy, sz = z - cz, lx = fma(sx, cosa, rn(sy * -sina)),
//   ly = fma(sy, cosa, rn(sx * sina)), inside = !(|sz| > hz) & |lx| <= tx & |ly| <= ty.
// 7 packed FP instructions (FADD2 x3, FMUL2 x2, FFMA2 x2) + 6 FSETP + 2 predicated IADD per
// two tests, instead of 14 + 6 + 2.
struct HypPacked {
    unsigned long long cx2, cy2, cz2, cosa2, nsina2, sina2;   // each value duplicated in both halves
    float hz, tx, ty;
};

__device__ __forceinline__ unsigned long long dup2(float v)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(v));
    return r;
}

__device__ __forceinline__ float lo_half(unsigned long long v) { return __uint_as_float((unsigned)v); }

__device__ __forceinline__ void count_pair(int &cnt, unsigned long long xx, unsigned long long yy,
                                           unsigned long long zz, const HypPacked &h)
{
    asm("{\n"
        " .reg .b64 sx, sy, sz, m1, m2, lx, ly;\n"
        " .reg .f32 a0, a1, b0, b1, c0, c1;\n"
        " .reg .pred p, q;\n"
        " sub.rn.f32x2 sx, %1, %4;\n"
        " sub.rn.f32x2 sy, %2, %5;\n"
        " sub.rn.f32x2 sz, %3, %6;\n"
        " mul.rn.f32x2 m1, sy, %8;\n"
        " mul.rn.f32x2 m2, sx, %9;\n"
        " fma.rn.f32x2 lx, sx, %7, m1;\n"
        " fma.rn.f32x2 ly, sy, %7, m2;\n"
        " mov.b64 {a0, a1}, lx;\n"
        " mov.b64 {b0, b1}, ly;\n"
        " mov.b64 {c0, c1}, sz;\n"
        " abs.f32 a0, a0;\n abs.f32 a1, a1;\n abs.f32 b0, b0;\n abs.f32 b1, b1;\n abs.f32 c0, c0;\n abs.f32 c1, c1;\n"
        " setp.leu.f32 p, c0, %10;\n"
        " setp.le.and.f32 p, a0, %11, p;\n"
        " setp.le.and.f32 p, b0, %12, p;\n"
        " setp.leu.f32 q, c1, %10;\n"
        " setp.le.and.f32 q, a1, %11, q;\n"
        " setp.le.and.f32 q, b1, %12, q;\n"
        " @p add.s32 %0, %0, 1;\n"
        " @q add.s32 %0, %0, 1;\n"
        "}\n"
        : "+r"(cnt)
        : "l"(xx), "l"(yy), "l"(zz), "l"(h.cx2), "l"(h.cy2), "l"(h.cz2), "l"(h.cosa2), "l"(h.nsina2), "l"(h.sina2),
          "f"(h.hz), "f"(h.tx), "f"(h.ty));
}

// Persistent CTAs pull (frustum, hypothesis chunk, point split) work items off a device
// counter.  A CTA keeps K hypotheses per thread in registers for the whole item and streams
// the item's points through a two-stage shared-memory ring filled by TMA bulk copies
// (cp.async.bulk + mbarrier complete_tx); every thread reads every staged pair record with
// broadcast LDS.128.
struct ScoreSmem {
    float4 tile[2][kScoreTile / 2][2];   // [stage][pair][x0x1y0y1 | z0z1d0d1]
    uint64_t bar[2];
    int item;
};

template <int K>
__device__ __forceinline__ void score_item(const fnp_seeker_batch &b, const int H, ScoreSmem &S, unsigned &it,
                                           const int f, const int h_base, const int split)
{
    const int tid = threadIdx.x;
    const int nv = b.hyp_nvalid[f];
    const int npts = b.cand_npts[f];
    const int p0 = split * b.split_points;                    // even: split_points is even
    const int n = min(npts, p0 + b.split_points) - p0;
    const int n_rec = (n + 1) >> 1;                           // pair records of this item
    const float4 *grec = reinterpret_cast<const float4 *>(b.frustum_pts) + (size_t)(b.cand_pt_start[f] + p0);
    constexpr int kRecTile = kScoreTile / 2;
    const int n_tiles = (n_rec + kRecTile - 1) / kRecTile;

    if (tid == 0) {
        for (int t = 0; t < 2 && t < n_tiles; t++) {
            const uint32_t bytes = (uint32_t)min(kRecTile, n_rec - t * kRecTile) * 32u;
            const unsigned st = (it + t) & 1u;
            mbar_expect_tx(&S.bar[st], bytes);
            tma_load_1d(S.tile[st], grec + (size_t)t * kRecTile * 2, bytes, &S.bar[st]);
        }
    }

    HypPacked hp[K];
    int cnt[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        const int r = h_base + k * kScoreThreads + tid;
        cnt[k] = 0;
        float4 a = make_float4(0.f, 0.f, 0.f, -1.f), c = make_float4(1.f, 0.f, -1.f, -1.f);   // never inside
        if (r < nv) {
            const float4 *src = reinterpret_cast<const float4 *>(b.hyp_prep + ((size_t)f * H + r) * 8);
            a = __ldg(src); c = __ldg(src + 1);
        }
        hp[k].cx2 = dup2(a.x); hp[k].cy2 = dup2(a.y); hp[k].cz2 = dup2(a.z); hp[k].hz = a.w;
        hp[k].cosa2 = dup2(c.x); hp[k].nsina2 = dup2(-c.y); hp[k].sina2 = dup2(c.y);
        hp[k].tx = c.z; hp[k].ty = c.w;
    }

    for (int t = 0; t < n_tiles; t++, it++) {
        const unsigned st = it & 1u;
        mbar_wait(&S.bar[st], (it >> 1) & 1u);
        const int m_pts = min(kScoreTile, n - t * kScoreTile);   // points in this tile
        const int m_full = m_pts >> 1;                            // complete pairs
        const ulonglong2 *tp = reinterpret_cast<const ulonglong2 *>(S.tile[st]);
        int i = 0;
        for (; i + 2 <= m_full; i += 2) {
            const ulonglong2 xy0 = tp[2 * i], zd0 = tp[2 * i + 1];
            const ulonglong2 xy1 = tp[2 * i + 2], zd1 = tp[2 * i + 3];
#pragma unroll
            for (int k = 0; k < K; k++) {
                count_pair(cnt[k], xy0.x, xy0.y, zd0.x, hp[k]);
                count_pair(cnt[k], xy1.x, xy1.y, zd1.x, hp[k]);
            }
        }
        for (; i < m_full; i++) {
            const ulonglong2 xy = tp[2 * i], zd = tp[2 * i + 1];
#pragma unroll
            for (int k = 0; k < K; k++) count_pair(cnt[k], xy.x, xy.y, zd.x, hp[k]);
        }
        if (m_pts & 1) {   // last point of the frustum: lane 0 of a half-filled record
            const float4 xy = S.tile[st][m_full][0], zd = S.tile[st][m_full][1];
#pragma unroll
            for (int k = 0; k < K; k++) {
                BoxPrep bp;
                bp.cx = lo_half(hp[k].cx2); bp.cy = lo_half(hp[k].cy2); bp.cz = lo_half(hp[k].cz2);
                bp.cosa = lo_half(hp[k].cosa2); bp.sina = lo_half(hp[k].sina2);
                bp.hz = hp[k].hz; bp.tx = hp[k].tx; bp.ty = hp[k].ty;
                count_if(cnt[k], in_box(xy.x, xy.z, zd.x, bp));
            }
        }
        __syncthreads();  // everyone is done with stage st
        if (tid == 0 && t + 2 < n_tiles) {
            const uint32_t bytes = (uint32_t)min(kRecTile, n_rec - (t + 2) * kRecTile) * 32u;
            mbar_expect_tx(&S.bar[st], bytes);
            tma_load_1d(S.tile[st], grec + (size_t)(t + 2) * kRecTile * 2, bytes, &S.bar[st]);
        }
    }

    int *out = b.counts + (size_t)f * H;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const int r = h_base + k * kScoreThreads + tid;
        if (r < nv && cnt[k]) atomicAdd(out + r, cnt[k]);      // RED.ADD, result unused
    }
}

__global__ void __launch_bounds__(kScoreThreads) score_kernel(const fnp_seeker_batch b, const int H)
{
    __shared__ __align__(128) ScoreSmem S;

    const int tid = threadIdx.x;
    if (b.status[0] & 2) return;
    const int n_items = b.status[2];
    if (tid == 0) {
        mbar_init(&S.bar[0], 1);
        mbar_init(&S.bar[1], 1);
        mbar_fence_init();
    }
    unsigned it = 0;   // tiles consumed so far by this CTA: stage = it & 1, parity = (it >> 1) & 1

    for (;;) {
        __syncthreads();                       // everyone is done with the previous item (and S.item)
        if (tid == 0) S.item = atomicAdd(&b.status[4], 1);
        __syncthreads();
        const int item_id = S.item;
        if (item_id >= n_items) break;
        const int4 item = reinterpret_cast<const int4 *>(b.items)[item_id];   // frustum, first hypothesis, split, K
        switch (item.w) {
            case 1: score_item<1>(b, H, S, it, item.x, item.y, item.z); break;
            case 2: score_item<2>(b, H, S, it, item.x, item.y, item.z); break;
            case 3: score_item<3>(b, H, S, it, item.x, item.y, item.z); break;
            default: score_item<4>(b, H, S, it, item.x, item.y, item.z); break;
        }
    }
}

// ======================================================================================
// Stage 2b, sweep mode: per-hypothesis point counts without testing every pair
// ======================================================================================
// Arithmetic and derivation: fnp_sweep.cuh (shared with the host model tools/sweep_model.cu).
// One CTA per frustum: line fit + deviation of every column.  Dynamic shared memory: 15 J words.
__global__ void __launch_bounds__(128) sweep_prep_kernel(const fnp_seeker_batch b, const int J, const int M)
{
    extern __shared__ int s_raw[];
    int *s_first = s_raw;                                  // [J]
    int *s_last = s_raw + J;                               // [J]
    int *s_r0 = s_raw + 2 * J;                             // [J] compacted slot of the column's first valid step
    float *s_c0 = reinterpret_cast<float *>(s_raw + 3 * J);   // [J][3]
    float *s_c1 = s_c0 + 3 * J;                            // [J][3], later the slopes
    unsigned *s_dev = reinterpret_cast<unsigned *>(s_c1 + 3 * J);   // [J][3] max of C - line (float bits, >= 0)
    unsigned *s_den = s_dev + 3 * J;                                // [J][3] max of line - C
    __shared__ unsigned s_maxabs, s_maxabs_z;
    const int f = blockIdx.x, tid = threadIdx.x;
    const int H = J * M;
    const int nv = b.hyp_nvalid[f];
    SweepCol *out = reinterpret_cast<SweepCol *>(b.sweep_cols) + (size_t)f * J;
    for (int j = tid; j < J; j += blockDim.x) {
        s_first[j] = 0x7fffffff;
        s_last[j] = -1;
        s_dev[3 * j] = s_dev[3 * j + 1] = s_dev[3 * j + 2] = 0u;
        s_den[3 * j] = s_den[3 * j + 1] = s_den[3 * j + 2] = 0u;
    }
    if (tid == 0) { s_maxabs = 0u; s_maxabs_z = 0u; }
    __syncthreads();
    const int *hidx = b.hyp_index + (size_t)f * H;
    float mabs = 0.f, mabs_z = 0.f;
    for (int r = tid; r < nv; r += blockDim.x) {
        const int h = hidx[r], m = h / J, j = h - m * J;
        atomicMin(&s_first[j], m);
        atomicMax(&s_last[j], m);
        const float *pp = b.hyp_prep + ((size_t)f * H + r) * 8;
        mabs = fmaxf(mabs, fmaxf(fabsf(pp[0]), fmaxf(fabsf(pp[1]), fabsf(pp[2]))));
        mabs_z = fmaxf(mabs_z, fabsf(pp[2]));
    }
    if (tid < 6 && nv > 0) {   // point AABB: pmin xyz, pmax xyz
        const float v = fabsf(b.cand_stats[(size_t)f * kStatsFloats + 3 + tid]);
        mabs = fmaxf(mabs, v);
        if (tid == 2 || tid == 5) mabs_z = v;
    }
    atomicMax(&s_maxabs, __float_as_uint(mabs));
    atomicMax(&s_maxabs_z, __float_as_uint(mabs_z));
    __syncthreads();
    for (int r = tid; r < nv; r += blockDim.x) {
        const int h = hidx[r], m = h / J, j = h - m * J;
        const bool first = (m == s_first[j]), last = (m == s_last[j]);
        if (first || last) {
            float Cv[3];
            sweep_axes(load_prep(b.hyp_prep, (size_t)f * H + r), Cv);
            for (int k = 0; k < 3; k++) {
                if (first) s_c0[3 * j + k] = Cv[k];
                if (last) s_c1[3 * j + k] = Cv[k];
            }
            if (first) s_r0[j] = r;
        }
    }
    __syncthreads();
    for (int i = tid; i < 3 * J; i += blockDim.x) {
        const int j = i / 3;
        const int span = s_last[j] - s_first[j];
        s_c1[i] = span > 0 ? __fdiv_rn(__fsub_rn(s_c1[i], s_c0[i]), (float)span) : 0.f;   // slope per depth step
    }
    __syncthreads();
    for (int r = tid; r < nv; r += blockDim.x) {
        const int h = hidx[r], m = h / J, j = h - m * J;
        float Cv[3];
        sweep_axes(load_prep(b.hyp_prep, (size_t)f * H + r), Cv);
        const float dm = (float)(m - s_first[j]);
        for (int k = 0; k < 3; k++) {
            const float line = __fmaf_rn(s_c1[3 * j + k], dm, s_c0[3 * j + k]);
            const float d = __fsub_rn(Cv[k], line);
            if (d > 0.f) atomicMax(&s_dev[3 * j + k], __float_as_uint(d));
            if (d < 0.f) atomicMax(&s_den[3 * j + k], __float_as_uint(-d));
        }
    }
    __syncthreads();
    const float eps = sweep_eps(__uint_as_float(s_maxabs)), eps_z = sweep_eps_z(__uint_as_float(s_maxabs_z));
    for (int j = tid; j < J; j += blockDim.x) {
        const int m0 = s_first[j], m1 = s_last[j];
        SweepCol c;
        if (m1 >= m0) {
            // every hypothesis of the column carries the same cosa, sina, tx, ty, hz: read the first one
            const float dev[3] = {__uint_as_float(s_dev[3 * j]), __uint_as_float(s_dev[3 * j + 1]),
                                  __uint_as_float(s_dev[3 * j + 2])};
            const float den[3] = {__uint_as_float(s_den[3 * j]), __uint_as_float(s_den[3 * j + 1]),
                                  __uint_as_float(s_den[3 * j + 2])};
            c = sweep_col_build(m0, m1, s_c0 + 3 * j, s_c1 + 3 * j, dev, den,
                                load_prep(b.hyp_prep, (size_t)f * H + s_r0[j]), eps, eps_z);
        } else {
            c = SweepCol{};
            c.m0 = 0; c.m1 = -1;
        }
        out[j] = c;
    }
}

#ifndef FNP_SWEEP_THREADS
#define FNP_SWEEP_THREADS 256
#endif
constexpr int kSweepThreads = FNP_SWEEP_THREADS;
constexpr int kSweepWarps = kSweepThreads / 32;
constexpr int kSweepChunk = 256;   // points per (column, chunk) warp item

// Entries of the uncertain-step queue of a CTA: (point | column << 16, packed steps).
__host__ __device__ inline int sweep_queue_cap(int SP, int J)
{
    // a quarter of the (point, column) pairs (~13 % have an entry on cfg2).  Measured on 128 cfg2 frames,
    // SP = 1024: a queue twice this size costs one resident CTA per SM and 19 % of the kernel's time,
    // more than the exact predicates taken in place when the queue is full.
    const int want = (SP * J / 4 + 31) & ~31;
    return want < 512 ? 512 : want > 4096 ? 4096 : want;
}

// Dynamic shared memory of sweep_score_kernel for split_points SP, H = M*J hypotheses, J columns.
__host__ __device__ inline size_t sweep_smem_bytes(int SP, int H, int J)
{
    return (size_t)J * sizeof(SweepCol) + (size_t)SP * 12 + (size_t)H * 4 + (size_t)sweep_queue_cap(SP, J) * 8 +
           (size_t)((H + 3) & ~3) * 2 + 16;
}

// Persistent CTAs pull (frustum, point split) items.  Per item:
//   stage   column parameters, the split's points as SoA, cleared difference arrays, slot table;
//   sweep   warps pull (column, 256-point chunk) pieces off a shared counter; per point one range
//           solve (sweep_solve), the definite range into the column's difference array
//           (shared-memory RED), the uncertain steps -- if any -- as one entry into the CTA's queue;
//   drain   the queue is expanded to single depth steps and spread evenly over the lanes: every
//           lane takes one exact predicate at a time, whichever point and column it belongs to;
//   scan    prefix sum over the depth steps of every column, one integer RED per valid hypothesis
//           into row f of `counts`.
#ifndef FNP_SWEEP_MIN_CTAS
#define FNP_SWEEP_MIN_CTAS 4
#endif
__global__ void __launch_bounds__(kSweepThreads, FNP_SWEEP_MIN_CTAS) sweep_score_kernel(const fnp_seeker_batch b, const int J, const int M)
{
    extern __shared__ __align__(16) unsigned char s_dyn[];
    const int H = J * M, SP = b.split_points;
    const int QCAP = sweep_queue_cap(SP, J);
    SweepCol *s_col = reinterpret_cast<SweepCol *>(s_dyn);                      // [J]   (80 B each: 16 B aligned)
    uint2 *s_q = reinterpret_cast<uint2 *>(s_col + J);                          // [QCAP] uncertain-step queue
    float *s_x = reinterpret_cast<float *>(s_q + QCAP);                         // [SP]  (SP is even: 8 B aligned rows)
    float *s_y = s_x + SP;
    float *s_z = s_y + SP;
    int *s_diff = reinterpret_cast<int *>(s_z + SP);                            // [J][M] difference array, then counts
    short *s_slot = reinterpret_cast<short *>(s_diff + H);                      // [H] compacted slot of hypothesis h, -1
    int *s_ctl = reinterpret_cast<int *>(s_slot + ((H + 3) & ~3));              // [0] item [1] next piece [2] queue size [3] queue head

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    if (b.status[0] & 2) return;
    const int n_items = b.status[2];
    auto red = [](int *p, int v) { atomicAdd(p, v); };

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            s_ctl[0] = atomicAdd(&b.status[4], 1);
            s_ctl[1] = 0; s_ctl[2] = 0; s_ctl[3] = 0;
        }
        __syncthreads();
        const int item_id = s_ctl[0];
        if (item_id >= n_items) break;
        const int4 item = reinterpret_cast<const int4 *>(b.items)[item_id];   // frustum, -, split, -
        const int f = item.x, split = item.z;
        const int nv = b.hyp_nvalid[f], npts = b.cand_npts[f];
        const int p0 = split * SP;
        const int n = min(npts, p0 + SP) - p0;
        const float *prep_f = b.hyp_prep + (size_t)f * H * 8;

        // ---- stage
        {
            const float4 *src = reinterpret_cast<const float4 *>(b.sweep_cols + (size_t)f * J * FNP_SWEEP_COL_FLOATS);
            float4 *dst = reinterpret_cast<float4 *>(s_col);
            for (int i = tid; i < J * (FNP_SWEEP_COL_FLOATS / 4); i += kSweepThreads) dst[i] = __ldg(src + i);
            const float4 *grec = reinterpret_cast<const float4 *>(b.frustum_pts) + (size_t)(b.cand_pt_start[f] + p0);
            const int n_rec = (n + 1) >> 1;
            for (int r = tid; r < n_rec; r += kSweepThreads) {
                const float4 xy = __ldg(grec + 2 * r), zd = __ldg(grec + 2 * r + 1);
                reinterpret_cast<float2 *>(s_x)[r] = make_float2(xy.x, xy.y);
                reinterpret_cast<float2 *>(s_y)[r] = make_float2(xy.z, xy.w);
                reinterpret_cast<float2 *>(s_z)[r] = make_float2(zd.x, zd.y);
            }
            for (int h = tid; h < H; h += kSweepThreads) { s_diff[h] = 0; s_slot[h] = -1; }
        }
        __syncthreads();
        {
            const int *hidx = b.hyp_index + (size_t)f * H;
            for (int r = tid; r < nv; r += kSweepThreads) s_slot[hidx[r]] = (short)r;
        }
        __syncthreads();

        // ---- sweep
        const int n_chunks = (n + kSweepChunk - 1) / kSweepChunk;
        const int n_pieces = J * n_chunks;
        for (;;) {
            int piece = 0;
            if (lane == 0) piece = atomicAdd(&s_ctl[1], 1);
            piece = __shfl_sync(0xffffffffu, piece, 0);
            if (piece >= n_pieces) break;
            const int j = piece % J, ch = piece / J;
            const SweepCol c = s_col[j];            // warp-uniform: lives in registers for the whole piece
            if (c.m1 < c.m0) continue;
            const int D = c.m1 - c.m0;
            int *diff = s_diff + j * M + c.m0;       // indexed by dm = m - m0
            const short *slot_col = s_slot + c.m0 * J + j;
            int base_cnt = 0;
            const int i_end = min(n, (ch + 1) * kSweepChunk);
            // two points per lane and pass: the two range solves are independent instruction chains
            for (int i0 = ch * kSweepChunk; i0 < i_end; i0 += 64) {
                int ia = i0 + lane, ib = i0 + 32 + lane;
                const bool live_a = ia < i_end, live_b = ib < i_end;
                ia = min(ia, i_end - 1); ib = min(ib, i_end - 1);
                const float xa = s_x[ia], ya = s_y[ia], za = s_z[ia];
                const float xb = s_x[ib], yb = s_y[ib], zb = s_z[ib];
                const SweepRanges ra = sweep_solve(c, xa, ya, za);
                const SweepRanges rb = sweep_solve(c, xb, yb, zb);
                unsigned wa = 0, wb = 0;
                if (live_a) {
                    base_cnt += sweep_add_definite(ra, D, diff, red);
                    wa = sweep_pack_uncertain(ra);
                }
                if (live_b) {
                    base_cnt += sweep_add_definite(rb, D, diff, red);
                    wb = sweep_pack_uncertain(rb);
                }
                const unsigned mka = __ballot_sync(0xffffffffu, wa != 0u), mkb = __ballot_sync(0xffffffffu, wb != 0u);
                if (mka | mkb) {
                    const int na = __popc(mka);
                    int qb = 0;
                    if (lane == 0) qb = atomicAdd(&s_ctl[2], na + __popc(mkb));
                    qb = __shfl_sync(0xffffffffu, qb, 0);
                    const int qa = qb + __popc(mka & lt_mask), qbb = qb + na + __popc(mkb & lt_mask);
                    if (wa) {
                        if (qa < QCAP) {
                            s_q[qa] = make_uint2((unsigned)ia | ((unsigned)j << 16), wa);
                        } else {   // queue full: take the exact predicates here
                            const int cnt = sweep_uncertain_count(wa);
                            for (int k = 0; k < cnt; k++)
                                sweep_exact_step_col(xa, ya, za, sweep_uncertain_step(wa, k), D, diff, slot_col, J, prep_f, c.cosa, c.sina, c.tx, c.ty, red);
                        }
                    }
                    if (wb) {
                        if (qbb < QCAP) {
                            s_q[qbb] = make_uint2((unsigned)ib | ((unsigned)j << 16), wb);
                        } else {
                            const int cnt = sweep_uncertain_count(wb);
                            for (int k = 0; k < cnt; k++)
                                sweep_exact_step_col(xb, yb, zb, sweep_uncertain_step(wb, k), D, diff, slot_col, J, prep_f, c.cosa, c.sina, c.tx, c.ty, red);
                        }
                    }
                }
            }
            base_cnt = __reduce_add_sync(0xffffffffu, base_cnt);
            if (lane == 0 && base_cnt) atomicAdd(diff, base_cnt);   // ranges that start at the column's first step
        }
        __syncthreads();

        // ---- drain: 32 queue entries per warp at a time, expanded to steps, one step per lane and pass
        const int qn = min(s_ctl[2], QCAP);
        for (;;) {
            int qb = 0;
            if (lane == 0) qb = atomicAdd(&s_ctl[3], 32);
            qb = __shfl_sync(0xffffffffu, qb, 0);
            if (qb >= qn) break;
            uint2 ent = make_uint2(0u, 0u);
            if (qb + lane < qn) ent = s_q[qb + lane];
            const int cnt = sweep_uncertain_count(ent.y);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            for (int t0 = 0; t0 < total; t0 += 32) {
                const int t = t0 + lane;
                // owner = first lane whose inclusive prefix exceeds t
                int own = 0;
#pragma unroll
                for (int step = 16; step; step >>= 1) {
                    const int v = __shfl_sync(0xffffffffu, incl, own + step - 1);
                    if (v <= t) own += step;
                }
                own = min(own, 31);
                const unsigned e0 = __shfl_sync(0xffffffffu, ent.x, own);
                const unsigned e1 = __shfl_sync(0xffffffffu, ent.y, own);
                const int first = __shfl_sync(0xffffffffu, incl - cnt, own);
                if (t < total) {
                    const int i = (int)(e0 & 0xffffu), j = (int)(e0 >> 16);
                    const float4 rot = *reinterpret_cast<const float4 *>(&s_col[j].cosa);     // cosa, sina, tx, ty
                    const int m0 = s_col[j].m0, D = s_col[j].m1 - m0;
                    sweep_exact_step_col(s_x[i], s_y[i], s_z[i], sweep_uncertain_step(e1, t - first), D, s_diff + j * M + m0,
                                         s_slot + m0 * J + j, J, prep_f, rot.x, rot.y, rot.z, rot.w, red);
                }
            }
        }
        __syncthreads();

        // ---- prefix sum over the depth steps of every column (warp per column, in place)
        for (int j = warp; j < J; j += kSweepWarps) {
            int carry = 0;
            for (int mb = 0; mb < M; mb += 32) {
                const int m = mb + lane;
                int v = m < M ? s_diff[j * M + m] : 0;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, v, o);
                    if (lane >= o) v += t;
                }
                v += carry;
                if (m < M) s_diff[j * M + m] = v;
                carry = __shfl_sync(0xffffffffu, v, 31);
            }
        }
        __syncthreads();
        int *out = b.counts + (size_t)f * H;
        for (int h = tid; h < H; h += kSweepThreads) {
            const int r = s_slot[h];
            if (r >= 0) {
                const int m = h / J, j = h - m * J;
                const int cnt = s_diff[j * M + m];
                if (cnt) atomicAdd(out + r, cnt);      // RED.ADD: the splits of a frustum add up in any order
            }
        }
    }
}

// ======================================================================================
// Stage 3: score + greedy argmax
// ======================================================================================
// Optional stage 2c: n_far of every compacted hypothesis = frustum points whose norm exceeds the norm of
// the hypothesis' nearest corner (calc_occl_scores, frustum_proposals_v1.py:408-477; the corners are rebuilt
// from the box the way boxes_to_corners_3d does, box_utils.py:28-52).  One thread per hypothesis, the
// norms of a tile of points in shared memory, read as broadcast LDS.128.
constexpr int kOcclThreads = 256;
constexpr int kOcclTile = 2048;

__global__ void __launch_bounds__(kOcclThreads) occl_kernel(const fnp_seeker_batch b, const int J, const int H)
{
    __shared__ __align__(16) float s_mag[kOcclTile];
    const int f = blockIdx.x, tid = threadIdx.x;
    const int nv = b.hyp_nvalid[f];
    const int r = blockIdx.y * kOcclThreads + tid;
    if ((int)(blockIdx.y * kOcclThreads) >= nv || b.status[0] != 0) return;
    const float INF = __int_as_float(0x7f800000);
    float m1 = INF;
    if (r < nv) {
        const float *pp = b.hyp_prep + ((size_t)f * H + r) * 8;
        const int j = b.hyp_index[(size_t)f * H + r] % J;
        const float *bb = b.base_boxes + ((size_t)(b.cand_label[f] - 1) * J + j) * 7;
        const float ca = cosf(bb[6]), sa = sinf(bb[6]);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const float sx = (k == 0 || k == 1 || k == 4 || k == 5) ? 0.5f : -0.5f;   // box_utils.py:42-45
            const float sy = (k == 0 || k == 3 || k == 4 || k == 7) ? 0.5f : -0.5f;
            const float sz = (k >= 4) ? 0.5f : -0.5f;
            const float x = __fmul_rn(bb[3], sx), y = __fmul_rn(bb[4], sy), z = __fmul_rn(bb[5], sz);
            const float cx = __fadd_rn(__fmaf_rn(y, -sa, __fmul_rn(x, ca)), pp[0]);
            const float cy = __fadd_rn(__fmaf_rn(y, ca, __fmul_rn(x, sa)), pp[1]);
            const float cz = __fadd_rn(z, pp[2]);
            m1 = fminf(m1, norm3(cx, cy, cz));
        }
    }
    const int n = b.cand_npts[f];
    const float *pts = b.frustum_pts + (size_t)(b.cand_pt_start[f] >> 1) * 8;
    int cnt = 0;
    for (int t0 = 0; t0 < n; t0 += kOcclTile) {
        const int tn = min(kOcclTile, n - t0);
        for (int i = tid; i < kOcclTile; i += kOcclThreads) {
            float m = -INF;   // padding never counts
            if (i < tn) {
                const int p = t0 + i;
                const float *rec = pts + (size_t)(p >> 1) * 8 + (p & 1);
                m = norm3(rec[0], rec[2], rec[4]);
            }
            s_mag[i] = m;
        }
        __syncthreads();
        const float4 *sm4 = reinterpret_cast<const float4 *>(s_mag);
        const int n4 = (tn + 3) >> 2;
#pragma unroll 4
        for (int i = 0; i < n4; i++) {
            const float4 m = sm4[i];
            cnt += (m.x > m1) + (m.y > m1) + (m.z > m1) + (m.w > m1);
        }
        __syncthreads();
    }
    if (r < nv) b.hyp_nfar[(size_t)f * H + r] = cnt;
}

// ======================================================================================
// Stage 3: second-stage score + greedy argmax
// ======================================================================================
__device__ __forceinline__ float block_max128(float v, float *s4, const int lane, const int warp)
{
    v = warp_max(v);
    if (lane == 0) s4[warp] = v;
    __syncthreads();
    v = fmaxf(fmaxf(s4[0], s4[1]), fmaxf(s4[2], s4[3]));
    __syncthreads();
    return v;
}

// Block-wide arg-max of (score, index) over 128 threads, lowest index on ties; every thread gets the result.
// besti == 0x7fffffff means "no entry".
__device__ __forceinline__ void block_argmax128(float &best, int &besti, float *s_f, int *s_i, const int lane,
                                                const int warp)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (oi != 0x7fffffff && (besti == 0x7fffffff || ob > best || (ob == best && oi < besti))) { best = ob; besti = oi; }
    }
    if (lane == 0) { s_f[warp] = best; s_i[warp] = besti; }
    __syncthreads();
    best = s_f[0]; besti = s_i[0];
#pragma unroll
    for (int w = 1; w < 4; w++) {
        const float ob = s_f[w];
        const int oi = s_i[w];
        if (oi != 0x7fffffff && (besti == 0x7fffffff || ob > best || (ob == best && oi < besti))) { best = ob; besti = oi; }
    }
    __syncthreads();
}

constexpr int kSelectMaxTopkH = 32768;   // hypotheses per frustum the suppression bitmask of topk > 1 covers

// EXTRAS = false is the shipped configuration (density + IoU, top-1): the optional terms, the score table and
// the suppression bitmask of topk > 1 are compiled out.
template <bool EXTRAS>
__global__ void __launch_bounds__(128) select_kernel(const fnp_seeker_batch b, const fnp_seeker_cfg cfg)
{
    __shared__ float s_f[4];
    __shared__ int s_i[4];
    __shared__ unsigned s_dead[EXTRAS ? kSelectMaxTopkH / 32 : 1];
    const int f = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int J = cfg.num_yaw_size, H = J * cfg.num_mags;
    const int T = EXTRAS ? max(cfg.topk, 1) : 1;
    const int nv = b.hyp_nvalid[f];
    if (nv <= 0 || (b.status[0] & 2)) {
        for (int k = tid; k < T; k += blockDim.x) {
            b.out_best[(size_t)f * T + k] = -1; b.out_score[(size_t)f * T + k] = 0.f; b.out_count[(size_t)f * T + k] = 0;
        }
        return;
    }
    const int *cbase = b.counts + (size_t)f * H;
    const float *pbase = b.hyp_prep + (size_t)f * H * 8;
    const bool mult = EXTRAS && (cfg.flags & FNP_SEEKER_MULT) != 0, occl_mult = EXTRAS && (cfg.flags & FNP_SEEKER_OCCL_MULT) != 0;
    const bool use_dist = EXTRAS && ((cfg.dst_w != 0.f) || mult);
    const bool use_fail = EXTRAS && ((cfg.occl_w > 0.f) || occl_mult);
    const bool use_ego = EXTRAS && cfg.ego_w > 0.f;
    const bool use_occl_w = EXTRAS && cfg.occl_w > 0.f;
    const long long npts = b.cand_npts[f];
    const int *nfar = use_fail ? b.hyp_nfar + (size_t)f * H : nullptr;
    // the reference's occlusion score: n_far * n_out (a (P,1) & (P,) broadcast, :453), stored as float
#define FNP_FAIL(r) __ll2float_rn((long long)nfar[r] * (npts - (long long)cbase[r]))
    int mxi = 0;
    float fmx = 0.f, emx = 0.f;
    for (int r = tid; r < nv; r += blockDim.x) {
        mxi = max(mxi, cbase[r]);
        if (use_fail) fmx = fmaxf(fmx, FNP_FAIL(r));
        if (use_ego) emx = fmaxf(emx, norm3(pbase[r * 8], pbase[r * 8 + 1], pbase[r * 8 + 2]));
    }
    const float den = __fadd_rn(block_max128((float)mxi, s_f, lane, warp), 1e-8f);
    float fden = 1.f, dmin = 0.f, dden = 1.f;
    if (use_fail) fden = __fadd_rn(block_max128(fmx, s_f, lane, warp), 1e-6f);
    if (use_ego) emx = block_max128(emx, s_f, lane, warp);
    if (use_dist) {
        const float *st = b.cand_stats + (size_t)f * kStatsFloats;
        dmin = st[13];
        dden = __fadd_rn(__fsub_rn(st[14], dmin), 1e-8f);
    }
    // argmax of score, lowest index wins ties (stable descending sort + top-1)
    float best = -__int_as_float(0x7f800000);
    int besti = 0x7fffffff;
    for (int r = tid; r < nv; r += blockDim.x) {
        const float dens = __fdiv_rn((float)cbase[r], den);
        const float iou = b.hyp_iou[(size_t)f * H + r];
        const float dr = use_dist ? __fsub_rn(1.0f, __fdiv_rn(__fsub_rn(b.hyp_dist[(size_t)f * H + r], dmin), dden)) : 1.0f;
        float sc;
        if (mult)
            sc = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(dens, cfg.dns_w), iou), cfg.iou_w), dr), cfg.dst_w);
        else {
            sc = __fadd_rn(__fmul_rn(dens, cfg.dns_w), __fmul_rn(iou, cfg.iou_w));
            if (use_dist) sc = __fadd_rn(sc, __fmul_rn(dr, cfg.dst_w));
        }
        if (use_occl_w) sc = __fadd_rn(sc, __fmul_rn(cfg.occl_w, __fsub_rn(1.0f, __fdiv_rn(FNP_FAIL(r), fden))));
        if (use_ego)
            sc = __fadd_rn(sc, __fmul_rn(cfg.ego_w, __fdiv_rn(norm3(pbase[r * 8], pbase[r * 8 + 1], pbase[r * 8 + 2]), emx)));
        if (occl_mult) sc = __fmul_rn(__fmul_rn(dens, iou), FNP_FAIL(r));
        if (T > 1) b.hyp_score[(size_t)f * H + r] = sc;
        if (sc > best || besti == 0x7fffffff) { best = sc; besti = r; }
    }
#undef FNP_FAIL
    if (T > 1) {
        for (int i = tid; i < (nv + 31) >> 5; i += blockDim.x) s_dead[i] = 0u;
    }
    block_argmax128(best, besti, s_f, s_i, lane, warp);     // also orders the hyp_score / s_dead writes
    const float *bb_tab = b.base_boxes + (size_t)(b.cand_label[f] - 1) * J * 7;
    for (int k = 0;;) {
        const float *pp = pbase + (size_t)besti * 8;
        const float *bb = bb_tab + (size_t)(b.hyp_index[(size_t)f * H + besti] % J) * 7;
        if (tid == 0) {
            float *o = b.out_boxes + ((size_t)f * T + k) * 7;
            o[0] = pp[0]; o[1] = pp[1]; o[2] = pp[2];
            o[3] = bb[3]; o[4] = bb[4]; o[5] = bb[5]; o[6] = bb[6];
            b.out_best[(size_t)f * T + k] = besti;
            b.out_score[(size_t)f * T + k] = best;
            b.out_count[(size_t)f * T + k] = cbase[besti];
        }
        if (++k >= T) break;
        // nms_normal_gpu (:1030, iou3d_nms.cpp:162-209): the kept box suppresses every later one with an
        // axis-aligned BEV IoU above the threshold; the next kept box is the best survivor
        const float abox[5] = {pp[0], pp[1], 0.f, bb[3], bb[4]};
        if (tid == 0) atomicOr(&s_dead[besti >> 5], 1u << (besti & 31));
        for (int r = tid; r < nv; r += blockDim.x) {
            if ((s_dead[r >> 5] >> (r & 31)) & 1u) continue;
            const float *bq = bb_tab + (size_t)(b.hyp_index[(size_t)f * H + r] % J) * 7;
            const float qbox[5] = {pbase[r * 8], pbase[r * 8 + 1], 0.f, bq[3], bq[4]};
            if (r != besti && iou_normal(abox, qbox) > cfg.nms_normal) atomicOr(&s_dead[r >> 5], 1u << (r & 31));
        }
        __syncthreads();
        best = -__int_as_float(0x7f800000);
        besti = 0x7fffffff;
        for (int r = tid; r < nv; r += blockDim.x) {
            if ((s_dead[r >> 5] >> (r & 31)) & 1u) continue;
            const float sc = b.hyp_score[(size_t)f * H + r];
            if (sc > best || besti == 0x7fffffff) { best = sc; besti = r; }
        }
        block_argmax128(best, besti, s_f, s_i, lane, warp);
        if (besti == 0x7fffffff) {       // fewer survivors than topk: the remaining slots stay empty
            for (int kk = k + tid; kk < T; kk += blockDim.x) {
                b.out_best[(size_t)f * T + kk] = -1; b.out_score[(size_t)f * T + kk] = 0.f; b.out_count[(size_t)f * T + kk] = 0;
            }
            break;
        }
    }
}

}  // namespace fnp

// ======================================================================================
// C ABI
// ======================================================================================
using namespace fnp;

static int check_batch(const fnp_seeker_cfg *cfg, const fnp_seeker_batch *b)
{
    if (!cfg || !b) return FNP_EINVAL;
    if (b->n_frames < 0 || b->n_cands < 0 || b->n_tiles < 0) return FNP_EINVAL;
    if (cfg->num_mags < 1 || cfg->num_yaw_size < 1) return FNP_EINVAL;
    if (b->max_items < 0 || b->split_points < 2 || (b->split_points & 1)) return FNP_EINVAL;
    return FNP_OK;
}

static size_t cull_stage_smem(const fnp_seeker_batch *b, int W)
{
    return sizeof(CullSmem) + (size_t)b->max_cands_per_frame * (16 + 4 * kCullVW + 4) + (size_t)6 * W * 4 + 16;
}

template <int W>
static int launch_cull(const fnp_seeker_cfg *cfg, const fnp_seeker_batch *b, cudaStream_t st)
{
    const int n_cu = cell_cols(cfg->img_w), n_cv = cell_rows(cfg->img_h);
    const size_t sa = cull_stage_smem(b, W), sc = (size_t)n_cu * n_cv * W * 4;
    const size_t sg = ((size_t)2 * b->max_cands_per_frame + 1) * 4;
    if (sa > 200 * 1024 || sc > 200 * 1024) return FNP_EINVAL;
    cudaFuncSetAttribute(cull_stage_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sa);
    cudaFuncSetAttribute(cell_table_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc);
    cell_table_kernel<W><<<b->n_frames * 6, 128, sc, st>>>(*b, n_cu, n_cv);
    cull_stage_kernel<W><<<b->n_tiles, kCullThreads, sa, st>>>(*b, cfg->img_w, cfg->img_h, n_cu, n_cv);
    scan_tiles_kernel<<<divup(b->n_cands, 4), 128, 0, st>>>(*b);
    scan_cands_kernel<<<1, 1024, 0, st>>>(*b);
    cull_gather_kernel<<<b->n_tiles, 256, sg, st>>>(*b);
    return FNP_OK;
}

extern "C" size_t fnp_seeker_cell_mask_bytes(const fnp_seeker_cfg *cfg, int n_frames, int max_cands_per_frame)
{
    const int W = fnp_seeker_mask_words(max_cands_per_frame);
    if (!cfg || n_frames < 0 || W < 0) return 0;
    return (size_t)n_frames * 6 * cell_cols(cfg->img_w) * cell_rows(cfg->img_h) * W * 4;
}

extern "C" int fnp_seeker_mask_words(int max_cands_per_frame)
{
    const int w = divup(max_cands_per_frame > 0 ? max_cands_per_frame : 1, 32);
    return w <= 1 ? 1 : w <= 2 ? 2 : w <= 4 ? 4 : w <= 8 ? 8 : -1;
}

extern "C" int fnp_seeker_cull(const fnp_seeker_cfg *cfg, const fnp_seeker_batch *b, void *stream)
{
    int rc = check_batch(cfg, b);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (b->n_cands == 0 || b->n_tiles == 0) {
        cudaMemsetAsync(b->status, 0, 8 * sizeof(int32_t), st);
        cudaMemsetAsync(b->cand_pt_start, 0, sizeof(int32_t) * (size_t)(b->n_cands + 1), st);
        if (b->n_cands) cudaMemsetAsync(b->cand_npts, 0, sizeof(int32_t) * (size_t)b->n_cands, st);
        FNP_LAUNCH_CHECK();
        return FNP_OK;
    }
    if (!b->points || !b->tile_counts || !b->tile_dst || !b->tile_base || !b->frustum_pts || !b->stage_pts ||
        !b->cell_masks || !b->cam_cand_start || b->point_stride < 3 || b->xyz_offset < 0 ||
        b->xyz_offset + 3 > b->point_stride || (b->frustum_idx && !b->stage_idx))
        return FNP_EINVAL;
    cudaMemsetAsync(b->status, 0, 8 * sizeof(int32_t), st);      // [5] = staging cursor
    const int W = fnp_seeker_mask_words(b->max_cands_per_frame);
    if (W < 0 || W != b->mask_words) return FNP_EINVAL;   // more than 256 candidates in one frame
    switch (W) {
        case 1: rc = launch_cull<1>(cfg, b, st); break;
        case 2: rc = launch_cull<2>(cfg, b, st); break;
        case 4: rc = launch_cull<4>(cfg, b, st); break;
        default: rc = launch_cull<8>(cfg, b, st); break;
    }
    if (rc) return rc;
    FNP_LAUNCH_CHECK();
    return FNP_OK;
}

extern "C" int fnp_seeker_frustum_stats(const fnp_seeker_cfg *cfg, const fnp_seeker_batch *b, void *stream)
{
    int rc = check_batch(cfg, b);
    if (rc) return rc;
    if (b->n_cands == 0) return FNP_OK;
    stats_kernel<<<b->n_cands, 256, 0, (cudaStream_t)stream>>>(*b, *cfg);
    FNP_LAUNCH_CHECK();
    return FNP_OK;
}

extern "C" int fnp_seeker_hypotheses(const fnp_seeker_cfg *cfg, const fnp_seeker_batch *b, void *stream)
{
    int rc = check_batch(cfg, b);
    if (rc) return rc;
    if (b->n_cands == 0) return FNP_OK;
    if ((cfg->flags & FNP_SEEKER_MULTICAM_IOU) || b->hyp_dist)
        hypotheses_kernel<true><<<b->n_cands, 128, 0, (cudaStream_t)stream>>>(*b, *cfg);
    else
        hypotheses_kernel<false><<<b->n_cands, 128, 0, (cudaStream_t)stream>>>(*b, *cfg);
    FNP_LAUNCH_CHECK();
    return FNP_OK;
}

// Scoring mode of a batch (see FNP_SCORE_* in fnp.h); < 0: invalid request.
static int resolve_score_mode(const fnp_seeker_cfg *cfg, const fnp_seeker_batch *b)
{
    const int J = cfg->num_yaw_size, M = cfg->num_mags;
    const long long H = (long long)J * M;
    const bool fits = b->sweep_cols && H <= 32767 && M <= 255 && J <= 65535 && b->split_points <= 65536 && sweep_smem_bytes(b->split_points, (int)H, J) <= 200 * 1024 &&
                      (size_t)15 * J * 4 <= 48 * 1024;
    switch (b->score_mode) {
        case FNP_SCORE_DIRECT: return FNP_SCORE_DIRECT;
        case FNP_SCORE_SWEEP: return fits ? FNP_SCORE_SWEEP : -1;
        case FNP_SCORE_AUTO: return (fits && M >= FNP_SWEEP_MIN_MAGS) ? FNP_SCORE_SWEEP : FNP_SCORE_DIRECT;
        default: return -1;
    }
}

extern "C" int fnp_seeker_score(const fnp_seeker_cfg *cfg, const fnp_seeker_batch *b, void *stream)
{
    int rc = check_batch(cfg, b);
    if (rc) return rc;
    if (b->n_cands == 0) return FNP_OK;
    const int J = cfg->num_yaw_size, M = cfg->num_mags, H = M * J;
    cudaStream_t st = (cudaStream_t)stream;
    if (!b->items || !b->cand_item_start || !b->counts) return FNP_EINVAL;
    const int mode = resolve_score_mode(cfg, b);
    if (mode < 0) return FNP_EINVAL;
    const int sweep = mode == FNP_SCORE_SWEEP;
    // per device: a process may run engines on several GPUs (function attributes and the SM count are per device)
    static int sms_of[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    int n_sms = 0;
    if (dev >= 0 && dev < 64 && sms_of[dev]) n_sms = sms_of[dev];
    else {
        cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev);
        if (dev >= 0 && dev < 64) sms_of[dev] = n_sms;
    }
    cudaMemsetAsync(b->counts, 0, sizeof(int32_t) * (size_t)b->n_cands * H, st);
    if (sweep) sweep_prep_kernel<<<b->n_cands, 128, (size_t)15 * J * 4, st>>>(*b, J, M);
    plan_items_kernel<<<1, 1024, 0, st>>>(*b, H, sweep);
    write_items_kernel<<<b->n_cands, 128, 0, st>>>(*b, H, sweep);
    if (b->max_items > 0) {
        // persistent CTAs: a whole number of CTAs per SM (148 SMs on B200), capped by the item capacity
        if (sweep) {
            const size_t smem = sweep_smem_bytes(b->split_points, H, J);
            static size_t smem_set[64] = {0};
            if (dev < 0 || dev >= 64 || smem > smem_set[dev]) {
                cudaFuncSetAttribute(sweep_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (dev >= 0 && dev < 64) smem_set[dev] = smem;
            }
            int per_sm = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sweep_score_kernel, kSweepThreads, smem);
            if (per_sm < 1) return FNP_EINVAL;
            const int grid = b->max_items < n_sms * per_sm ? b->max_items : n_sms * per_sm;
            sweep_score_kernel<<<grid, kSweepThreads, smem, st>>>(*b, J, M);
        } else {
            const int per_sm = 6;
            const int grid = b->max_items < n_sms * per_sm ? b->max_items : n_sms * per_sm;
            score_kernel<<<grid, kScoreThreads, 0, st>>>(*b, H);
        }
    }
    FNP_LAUNCH_CHECK();
    return FNP_OK;
}

/* Which scoring kernel fnp_seeker_score would run for this batch: FNP_SCORE_DIRECT or FNP_SCORE_SWEEP
 * (FNP_EINVAL: the requested mode cannot run, e.g. SWEEP without sweep_cols). */
extern "C" int fnp_seeker_score_mode(const fnp_seeker_cfg *cfg, const fnp_seeker_batch *b)
{
    if (!cfg || !b) return FNP_EINVAL;
    const int m = resolve_score_mode(cfg, b);
    return m < 0 ? FNP_EINVAL : m;
}

extern "C" int fnp_seeker_select(const fnp_seeker_cfg *cfg, const fnp_seeker_batch *b, void *stream)
{
    int rc = check_batch(cfg, b);
    if (rc) return rc;
    if (((cfg->dst_w != 0.f) || (cfg->flags & FNP_SEEKER_MULT)) && !b->hyp_dist) return FNP_EINVAL;
    if (((cfg->occl_w > 0.f) || (cfg->flags & FNP_SEEKER_OCCL_MULT)) && !b->hyp_nfar) return FNP_EINVAL;
    if (cfg->topk > 1 && (!b->hyp_score || cfg->num_yaw_size * cfg->num_mags > kSelectMaxTopkH)) return FNP_EINVAL;
    if (b->n_cands == 0) return FNP_OK;
    const bool extras = cfg->dst_w != 0.f || cfg->ego_w > 0.f || cfg->occl_w > 0.f || cfg->flags != 0 || cfg->topk > 1;
    if (extras)
        select_kernel<true><<<b->n_cands, 128, 0, (cudaStream_t)stream>>>(*b, *cfg);
    else
        select_kernel<false><<<b->n_cands, 128, 0, (cudaStream_t)stream>>>(*b, *cfg);
    FNP_LAUNCH_CHECK();
    return FNP_OK;
}

extern "C" int fnp_seeker_occlusion(const fnp_seeker_cfg *cfg, const fnp_seeker_batch *b, void *stream)
{
    int rc = check_batch(cfg, b);
    if (rc) return rc;
    if (!(cfg->occl_w > 0.f) && !(cfg->flags & FNP_SEEKER_OCCL_MULT)) return FNP_OK;
    if (!b->hyp_nfar) return FNP_EINVAL;
    if (b->n_cands == 0) return FNP_OK;
    const int J = cfg->num_yaw_size, H = J * cfg->num_mags;
    occl_kernel<<<dim3(b->n_cands, divup(H, kOcclThreads)), kOcclThreads, 0, (cudaStream_t)stream>>>(*b, J, H);
    FNP_LAUNCH_CHECK();
    return FNP_OK;
}

extern "C" int fnp_seeker_run(const fnp_seeker_cfg *cfg, const fnp_seeker_batch *b, void *stream)
{
    int rc;
    if ((rc = fnp_seeker_cull(cfg, b, stream))) return rc;
    if ((rc = fnp_seeker_frustum_stats(cfg, b, stream))) return rc;
    if ((rc = fnp_seeker_hypotheses(cfg, b, stream))) return rc;
    if ((rc = fnp_seeker_score(cfg, b, stream))) return rc;
    if ((rc = fnp_seeker_occlusion(cfg, b, stream))) return rc;
    return fnp_seeker_select(cfg, b, stream);
}
