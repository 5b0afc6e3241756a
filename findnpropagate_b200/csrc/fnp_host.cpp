// Host-side planning of a seeker batch (C ABI, no CUDA calls): which GLIP 2D boxes become
// candidate frustums, and in which order.
//
// Reference: pcdet/models/dense_heads/frustum_proposals_v1.py:582-595 -- per frame, cameras are
// visited in image_order [2,0,1,5,3,4]; each camera's boxes go through
// torchvision.ops.batched_nms(boxes, scores, labels, nms_2d) on the CPU (indices come back in
// descending score order) and boxes with score < score_thr are skipped.  torchvision's CPU
// arithmetic is restated here (ops/boxes.py batched_nms + csrc/ops/cpu/nms_kernel.cpp):
//   * <= 1000 boxes in the group: coordinate trick -- every box is shifted by
//     label * (max coordinate of the group + 1) in fp32, then one class-agnostic greedy NMS;
//   * more: one greedy NMS per label on the unshifted boxes;
//   * greedy NMS: stable descending score order, areas and IoU in fp32,
//     suppressed when (double)iou > iou_threshold.
// The reference makes 6 Python-level calls per frame; this does a whole batch in one call.
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <vector>

#include "../../include/fnp.h"

namespace {

const int kCamRank[6] = {1, 2, 0, 4, 5, 3};  // rank of camera c in image_order [2,0,1,5,3,4]

struct GreedyNms {
    std::vector<float> x1, y1, x2, y2, area;
    std::vector<uint8_t> dead;
    // boxes given in priority order; keep[i] = survives
    void run(int n, double thr, std::vector<uint8_t> &keep)
    {
        area.resize(n);
        dead.assign(n, 0);
        keep.assign(n, 0);
        for (int i = 0; i < n; i++) area[i] = (x2[i] - x1[i]) * (y2[i] - y1[i]);
        for (int i = 0; i < n; i++) {
            if (dead[i]) continue;
            keep[i] = 1;
            const float ix1 = x1[i], iy1 = y1[i], ix2 = x2[i], iy2 = y2[i], ia = area[i];
            for (int j = i + 1; j < n; j++) {
                if (dead[j]) continue;
                const float xx1 = std::max(ix1, x1[j]), yy1 = std::max(iy1, y1[j]);
                const float xx2 = std::min(ix2, x2[j]), yy2 = std::min(iy2, y2[j]);
                const float w = std::max(0.0f, xx2 - xx1), h = std::max(0.0f, yy2 - yy1);
                const float inter = w * h;
                if (!(inter > 0.0f)) continue;      // disjoint: ovr is 0 (or 0/0 = NaN), never > thr; skips the division
                const float ovr = inter / (ia + area[j] - inter);
                if ((double)ovr > thr) dead[j] = 1;
            }
        }
    }
};

}  // namespace

extern "C" int fnp_host_select_candidates(const float *det_boxes, const int64_t *det_labels,
                                          const float *det_scores, const int64_t *det_frame,
                                          const int64_t *det_cam, int n_dets, int n_frames, float nms_2d,
                                          float score_thr, int32_t *cand_det, int32_t *frame_cand_start)
{
    if (n_dets < 0 || n_frames < 0 || !frame_cand_start) return FNP_EINVAL;
    if (n_dets > 0 && (!det_boxes || !det_labels || !det_scores || !det_frame || !det_cam || !cand_det))
        return FNP_EINVAL;
    for (int b = 0; b <= n_frames; b++) frame_cand_start[b] = 0;
    if (n_dets == 0) return 0;
    std::vector<int64_t> group(n_dets);
    for (int i = 0; i < n_dets; i++) {
        if (det_cam[i] < 0 || det_cam[i] > 5 || det_frame[i] < 0 || det_frame[i] >= n_frames) return FNP_EINVAL;
        group[i] = det_frame[i] * 6 + kCamRank[det_cam[i]];
    }
    // group ascending, score descending, original index ascending: a counting sort over the
    // (frame, camera rank) groups (stable), then a stable sort by score inside each small group
    // (one global comparison sort of all detections was 3/4 of this function's time)
    const int n_groups = n_frames * 6;
    std::vector<int32_t> gstart(n_groups + 1, 0);
    for (int i = 0; i < n_dets; i++) gstart[group[i] + 1]++;
    for (int g = 0; g < n_groups; g++) gstart[g + 1] += gstart[g];
    std::vector<int32_t> order(n_dets), fill(gstart.begin(), gstart.end() - 1);
    for (int i = 0; i < n_dets; i++) order[fill[group[i]]++] = i;
    for (int g = 0; g < n_groups; g++) {
        const int n = gstart[g + 1] - gstart[g];
        int32_t *o = order.data() + gstart[g];
        if (n > 1 && n <= 48) {          // the usual case: a stable insertion sort, no temporary buffer
            for (int i = 1; i < n; i++) {
                const int32_t v = o[i];
                const float sv = det_scores[v];
                int j = i - 1;
                while (j >= 0 && det_scores[o[j]] < sv) { o[j + 1] = o[j]; j--; }
                o[j + 1] = v;
            }
        } else if (n > 1) {
            std::stable_sort(o, o + n, [&](int32_t a, int32_t b) { return det_scores[a] > det_scores[b]; });
        }
    }
    GreedyNms nms;
    std::vector<uint8_t> keep, keep_l;
    std::vector<int> idx_l;
    const double thr = (double)nms_2d;
    int n_out = 0;
    for (int s = 0; s < n_dets;) {
        int e = s;
        while (e < n_dets && group[order[e]] == group[order[s]]) e++;
        const int n = e - s;
        keep.assign(n, 0);
        if (n * 4 <= 4000) {
            float mx = det_boxes[(size_t)order[s] * 4];
            for (int i = 0; i < n; i++)
                for (int k = 0; k < 4; k++) mx = std::max(mx, det_boxes[(size_t)order[s + i] * 4 + k]);
            const float step = mx + 1.0f;
            nms.x1.resize(n); nms.y1.resize(n); nms.x2.resize(n); nms.y2.resize(n);
            for (int i = 0; i < n; i++) {
                const float *bx = det_boxes + (size_t)order[s + i] * 4;
                const float off = (float)det_labels[order[s + i]] * step;
                nms.x1[i] = bx[0] + off; nms.y1[i] = bx[1] + off; nms.x2[i] = bx[2] + off; nms.y2[i] = bx[3] + off;
            }
            nms.run(n, thr, keep);
        } else {
            // per-label NMS on the raw boxes (torchvision's "vanilla" path for large groups)
            std::vector<int64_t> labels;
            for (int i = 0; i < n; i++) labels.push_back(det_labels[order[s + i]]);
            std::sort(labels.begin(), labels.end());
            labels.erase(std::unique(labels.begin(), labels.end()), labels.end());
            for (int64_t lab : labels) {
                idx_l.clear();
                for (int i = 0; i < n; i++)
                    if (det_labels[order[s + i]] == lab) idx_l.push_back(i);
                const int m = (int)idx_l.size();
                nms.x1.resize(m); nms.y1.resize(m); nms.x2.resize(m); nms.y2.resize(m);
                for (int i = 0; i < m; i++) {
                    const float *bx = det_boxes + (size_t)order[s + idx_l[i]] * 4;
                    nms.x1[i] = bx[0]; nms.y1[i] = bx[1]; nms.x2[i] = bx[2]; nms.y2[i] = bx[3];
                }
                nms.run(m, thr, keep_l);
                for (int i = 0; i < m; i++) keep[idx_l[i]] = keep_l[i];
            }
        }
        const int frame = (int)det_frame[order[s]];
        for (int i = 0; i < n; i++) {
            const int d = order[s + i];
            if (keep[i] && !(det_scores[d] < score_thr)) {
                cand_det[n_out++] = d;
                frame_cand_start[frame + 1]++;
            }
        }
        s = e;
    }
    for (int b = 0; b < n_frames; b++) frame_cand_start[b + 1] += frame_cand_start[b];
    return n_out;
}

// Priority order of stage 4 (rotated-BEV NMS of a frame's proposals): inside every frame the
// candidates by descending 2D score, ties by candidate index (what np.lexsort((index, -score,
// frame)) gives; pseudo_loader.py:29-55 sorts the same way before its CPU NMS).
extern "C" int fnp_host_nms_order(const float *cand_score, const int32_t *frame_cand_start, int n_frames,
                                  int32_t *order)
{
    if (n_frames < 0 || !frame_cand_start || (n_frames > 0 && frame_cand_start[n_frames] > 0 && (!cand_score || !order)))
        return FNP_EINVAL;
    for (int b = 0; b < n_frames; b++) {
        const int s = frame_cand_start[b], e = frame_cand_start[b + 1];
        if (e < s) return FNP_EINVAL;
        for (int i = s; i < e; i++) order[i] = i;
        std::stable_sort(order + s, order + e, [&](int32_t a, int32_t c) { return cand_score[a] > cand_score[c]; });
    }
    return FNP_OK;
}

// ---------------------------------------------------------------------------------------
// Per-box constants of the reference's CPU point-in-box predicate (roiaware_pool3d.cpp:121-140), computed
// where the reference computes them -- on the host, with the host's libm:
//     cosa = cos(-rz), sina = sin(-rz)                        (float overloads: glibc cosf / sinf)
//     in_z  = !(fabsf(z - cz) > dz / 2.0)                     (fp64 compare)
//     in_xy = fabs(lx) < dx / 2.0 + MARGIN, MARGIN = 1e-2f    (fp64 compare, MARGIN a float constant)
// The fp64 compares are hoisted into fp32 thresholds exactly:  |l| < t  <=>  |l| <= pred(t), pred(t) the
// largest float strictly below t;  |sz| > h  <=>  |sz| > rd(h), rd = round down to float.
// ---------------------------------------------------------------------------------------
#include <cmath>
#include <cstring>

static float strict_below(double t)
{
    if (!(t > 0.0)) return (t != t) ? std::nanf("") : -1.0f;      // nothing is inside
    float f = (float)t;                                           // round to nearest
    if ((double)f >= t) f = std::nextafterf(f, -INFINITY);        // largest float < t
    return f;
}

extern "C" int fnp_host_prep_boxes_cpu(const float *boxes_host, float *prep_host, int N)
{
    if (N < 0 || (N > 0 && (!boxes_host || !prep_host))) return FNP_EINVAL;
    for (int i = 0; i < N; i++) {
        const float *b = boxes_host + (size_t)i * 7;
        float *q = prep_host + (size_t)i * 8;
        const float rz = b[6];
        q[0] = b[0]; q[1] = b[1]; q[2] = b[2];
        const double h = (double)b[5] / 2.0;
        float hz = (float)h;
        if ((double)hz > h) hz = std::nextafterf(hz, -INFINITY);  // rd(h): |sz| > h  <=>  |sz| > rd(h) for float |sz|
        q[3] = hz;
        q[4] = cosf(-rz);
        q[5] = sinf(-rz);
        q[6] = strict_below((double)b[3] / 2.0 + (double)1e-2f);
        q[7] = strict_below((double)b[4] / 2.0 + (double)1e-2f);
    }
    return FNP_OK;
}

// ---------------------------------------------------------------------------------------
// Whole-batch planning in one call (what SeekerEngine.plan used to do in numpy, 3 ms per 256 frames
// on the main thread): candidate selection as above over the frames' own detection arrays, the
// per-(frame, camera rank) candidate ranges, the stage-4 priority order, the point-tile table and
// the camera-matrix block, written straight into caller-provided arrays.
// ---------------------------------------------------------------------------------------
extern "C" int fnp_host_plan_sizes(const fnp_host_frame *frames, int n_frames, int64_t *sizes3)
{
    if (n_frames < 0 || !sizes3 || (n_frames > 0 && !frames)) return FNP_EINVAL;
    int64_t n_dets = 0, n_tiles = 0, rows = 0;
    for (int b = 0; b < n_frames; b++) {
        if (frames[b].n_rows < 0 || frames[b].n_dets < 0) return FNP_EINVAL;
        n_dets += frames[b].n_dets;
        n_tiles += (frames[b].n_rows + FNP_CULL_TILE - 1) / FNP_CULL_TILE;
        rows += frames[b].n_rows;
    }
    if (n_dets > 0x7fffffff || n_tiles > 0x7fffffff) return FNP_EINVAL;
    sizes3[0] = n_dets; sizes3[1] = n_tiles; sizes3[2] = rows;
    return FNP_OK;
}

extern "C" int fnp_host_plan(const fnp_host_frame *frames, int n_frames, float nms_2d, float score_thr, int box_xywh,
                             int topk, fnp_host_plan_out *out)
{
    if (n_frames < 0 || !out || (n_frames > 0 && !frames)) return FNP_EINVAL;
    int64_t sz[3];
    int rc = fnp_host_plan_sizes(frames, n_frames, sz);
    if (rc) return rc;
    const int n_dets = (int)sz[0];
    if (!out->frame_row_start || !out->frame_tile_start || !out->frame_cand_start || !out->cam_cand_start) return FNP_EINVAL;
    if (sz[1] > 0 && (!out->tile_frame || !out->tile_row0)) return FNP_EINVAL;
    if (n_frames > 0 && !out->cam_mats) return FNP_EINVAL;
    // ---- rows, tiles (tiles never straddle a frame), camera matrices
    int64_t row = 0;
    int tile = 0;
    for (int b = 0; b < n_frames; b++) {
        out->frame_row_start[b] = row;
        out->frame_tile_start[b] = tile;
        const int nt = (int)((frames[b].n_rows + FNP_CULL_TILE - 1) / FNP_CULL_TILE);
        for (int t = 0; t < nt; t++) { out->tile_frame[tile + t] = b; out->tile_row0[tile + t] = t * FNP_CULL_TILE; }
        tile += nt;
        row += frames[b].n_rows;
        if (!frames[b].cam_mats) return FNP_EINVAL;
        std::copy(frames[b].cam_mats, frames[b].cam_mats + 144, out->cam_mats + (size_t)b * 144);
    }
    out->frame_row_start[n_frames] = row;
    out->frame_tile_start[n_frames] = tile;
    out->n_tiles = tile;
    out->total_rows = row;
    // ---- detections of the batch, flat (a few thousand rows)
    std::vector<float> boxes((size_t)n_dets * 4), scores(n_dets);
    std::vector<int64_t> labels(n_dets), dframe(n_dets), dcam(n_dets);
    int d0 = 0;
    for (int b = 0; b < n_frames; b++) {
        const fnp_host_frame &f = frames[b];
        if (f.n_dets > 0 && (!f.det_boxes || !f.det_labels || !f.det_scores || !f.det_cam)) return FNP_EINVAL;
        for (int i = 0; i < f.n_dets; i++) {
            for (int k = 0; k < 4; k++) boxes[(size_t)(d0 + i) * 4 + k] = f.det_boxes[(size_t)i * 4 + k];
            scores[d0 + i] = f.det_scores[i];
            labels[d0 + i] = f.det_labels[i];
            dcam[d0 + i] = f.det_cam[i];
            dframe[d0 + i] = b;
        }
        d0 += f.n_dets;
    }
    if (n_dets > 0 && (!out->cand_det || !out->cand_frame || !out->cand_cam || !out->cand_label || !out->cand_box2d ||
                       !out->cand_score || !out->nms_order))
        return FNP_EINVAL;
    const int F = fnp_host_select_candidates(boxes.data(), labels.data(), scores.data(), dframe.data(), dcam.data(), n_dets,
                                             n_frames, nms_2d, score_thr, out->cand_det, out->frame_cand_start);
    if (F < 0) return F;
    out->n_cands = F;
    for (int g = 0; g <= n_frames * 6; g++) out->cam_cand_start[g] = 0;
    int max_c = 0;
    for (int b = 0; b < n_frames; b++) max_c = std::max(max_c, out->frame_cand_start[b + 1] - out->frame_cand_start[b]);
    out->max_cands_per_frame = max_c;
    for (int i = 0; i < F; i++) {
        const int d = out->cand_det[i];
        const int fr = (int)dframe[d], cam = (int)dcam[d];
        out->cand_frame[i] = fr;
        out->cand_cam[i] = cam;
        out->cand_label[i] = (int32_t)labels[d];
        out->cand_score[i] = scores[d];
        float *bx = out->cand_box2d + (size_t)i * 4;
        for (int k = 0; k < 4; k++) bx[k] = boxes[(size_t)d * 4 + k];
        if (box_xywh) { bx[2] = bx[2] + bx[0]; bx[3] = bx[3] + bx[1]; }   // frustum_proposals_v1.py:597-601, after the NMS
        out->cam_cand_start[fr * 6 + kCamRank[cam] + 1]++;                 // candidates are already in this order
    }
    for (int g = 0; g < n_frames * 6; g++) out->cam_cand_start[g + 1] += out->cam_cand_start[g];
    rc = fnp_host_nms_order(out->cand_score, out->frame_cand_start, n_frames, out->nms_order);
    if (rc) return rc;
    if (topk > 1) {   // every candidate owns topk consecutive proposal slots; a candidate's slots stay together, best first
        if (!out->frame_prop_start || (F > 0 && !out->prop_order)) return FNP_EINVAL;
        for (int b = 0; b <= n_frames; b++) out->frame_prop_start[b] = out->frame_cand_start[b] * topk;
        for (int i = 0; i < F; i++)
            for (int k = 0; k < topk; k++) out->prop_order[(size_t)i * topk + k] = out->nms_order[i] * topk + k;
    }
    return FNP_OK;
}

// ---------------------------------------------------------------------------------------
// Host-side column gather of the point table.
//
// The reference uploads every column of `points` ([batch_idx,] x, y, z, intensity, time:
// pcdet/models/__init__.py:23-36 load_data_to_gpu) although the seeker reads xyz only
// (frustum_proposals_v1.py:571-575).  With host buffers on one side of a PCIe link the point
// table IS the end-to-end cost (20 B/point at ~55 GB/s against ~2.5 ms of kernels per 128
// frames), so the host side of the path gathers x, y, z into a pinned staging buffer and only
// those 12 B/point cross the link.  Worker threads split the rows.
// ---------------------------------------------------------------------------------------
#include <atomic>
#include <cstdint>
#include <mutex>
#include <thread>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace {

// Rows [r0, r1).  Four rows make three 16-byte vectors, written with non-temporal stores: the
// staging buffer is not read into the cache before it is overwritten (measured on the 16-core
// GPU box, 41 M rows of 5 floats: 84 GB/s in with plain stores, 100 GB/s with these; 4-byte
// non-temporal stores were slower than plain ones, tools/probes/pack_probe.cpp).
void pack_rows(const float *src, int64_t r0, int64_t r1, int stride, int off, float *dst)
{
    const float *s = src + r0 * stride + off;
    float *d = dst + r0 * 3;
    int64_t r = r0;
#if defined(__SSE2__)
    for (; r < r1 && (reinterpret_cast<uintptr_t>(d) & 15); r++, s += stride, d += 3) { d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; }
    if (stride == 5) {
        // 4 rows = 20 floats a[0..3] b[4..7] c[8..11] e[12..15] f[16..19]; wanted: 0 1 2 5 | 6 7 10 11 | 12 15 16 17
        for (; r + 4 <= r1; r += 4, s += 20, d += 12) {
            const __m128 a = _mm_loadu_ps(s), b = _mm_loadu_ps(s + 4), c = _mm_loadu_ps(s + 8), e = _mm_loadu_ps(s + 12),
                         f = (r + 5 <= r1 || off == 0) ? _mm_loadu_ps(s + 16) : _mm_set_ps(0.f, 0.f, s[17], s[16]);
            const __m128 o0 = _mm_shuffle_ps(a, _mm_shuffle_ps(a, b, _MM_SHUFFLE(1, 1, 2, 2)), _MM_SHUFFLE(2, 0, 1, 0));
            const __m128 o1 = _mm_shuffle_ps(_mm_shuffle_ps(b, b, _MM_SHUFFLE(3, 3, 3, 2)), c, _MM_SHUFFLE(3, 2, 1, 0));
            const __m128 o2 = _mm_shuffle_ps(_mm_shuffle_ps(e, e, _MM_SHUFFLE(3, 3, 3, 0)), f, _MM_SHUFFLE(1, 0, 1, 0));
            _mm_stream_ps(d, o0); _mm_stream_ps(d + 4, o1); _mm_stream_ps(d + 8, o2);
        }
    } else {
        for (; r + 4 <= r1; r += 4, s += 4 * (int64_t)stride, d += 12) {
            const float *s1 = s + stride, *s2 = s1 + stride, *s3 = s2 + stride;
            _mm_stream_ps(d, _mm_set_ps(s1[0], s[2], s[1], s[0]));
            _mm_stream_ps(d + 4, _mm_set_ps(s2[1], s2[0], s1[2], s1[1]));
            _mm_stream_ps(d + 8, _mm_set_ps(s3[2], s3[1], s3[0], s2[2]));
        }
    }
#endif
    for (; r < r1; r++, s += stride, d += 3) { d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; }
#if defined(__SSE2__)
    _mm_sfence();
#endif
}

struct PackJob {
    std::vector<std::thread> threads;
    bool used = false;
};
constexpr int kMaxPackJobs = 16;
PackJob g_jobs[kMaxPackJobs];
std::mutex g_jobs_mutex;

}  // namespace

// Worker: the rows [g0, g1) of the concatenation of the segments (segment s holds seg_rows[s] rows
// at seg_src[s]; its first row is row seg_first[s] of dst).
static void pack_segments(std::vector<const float *> seg_src, std::vector<int64_t> seg_first, int64_t g0, int64_t g1,
                          int stride, int off, float *dst)
{
    const int n_seg = (int)seg_src.size();
    int s = (int)(std::upper_bound(seg_first.begin(), seg_first.end(), g0) - seg_first.begin()) - 1;
    while (g0 < g1 && s < n_seg) {
        const int64_t seg_end = (s + 1 < n_seg) ? seg_first[s + 1] : g1;
        const int64_t e = std::min(g1, seg_end);
        if (e > g0) pack_rows(seg_src[s], g0 - seg_first[s], e - seg_first[s], stride, off, dst + seg_first[s] * 3);
        g0 = std::max(g0, e);
        s++;
    }
}

extern "C" int fnp_host_pack_xyz_multi_begin(const float *const *seg_src_host, const int64_t *seg_rows, int n_segments,
                                             int stride, int xyz_offset, float *dst_host, int n_threads)
{
    if (n_segments < 0 || stride < 3 || xyz_offset < 0 || xyz_offset + 3 > stride) return FNP_EINVAL;
    if (n_segments > 0 && (!seg_src_host || !seg_rows)) return FNP_EINVAL;
    std::vector<const float *> src;
    std::vector<int64_t> first;
    int64_t rows = 0;
    for (int s = 0; s < n_segments; s++) {
        if (seg_rows[s] < 0 || (seg_rows[s] > 0 && !seg_src_host[s])) return FNP_EINVAL;
        if (seg_rows[s] == 0) continue;
        src.push_back(seg_src_host[s]);
        first.push_back(rows);
        rows += seg_rows[s];
    }
    if (rows > 0 && !dst_host) return FNP_EINVAL;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    int ticket = -1;
    {
        std::lock_guard<std::mutex> lk(g_jobs_mutex);
        for (int i = 0; i < kMaxPackJobs; i++)
            if (!g_jobs[i].used) { ticket = i; g_jobs[i].used = true; break; }
    }
    if (ticket < 0) return FNP_EWORKSPACE;      // too many gathers in flight
    PackJob &job = g_jobs[ticket];
    // thread shares are multiples of 4 rows so that every share but the last starts 16-byte aligned
    const int64_t per = ((rows + n_threads - 1) / n_threads + 3) & ~(int64_t)3;
    for (int t = 0; t < n_threads; t++) {
        const int64_t g0 = (int64_t)t * per, g1 = std::min(rows, g0 + per);
        if (g0 >= g1) break;
        job.threads.emplace_back(pack_segments, src, first, g0, g1, stride, xyz_offset, dst_host);
    }
    return ticket;
}

extern "C" int fnp_host_pack_xyz_begin(const float *src_host, int64_t rows, int stride, int xyz_offset, float *dst_host,
                                       int n_threads)
{
    if (rows < 0 || (rows > 0 && !src_host)) return FNP_EINVAL;
    const float *srcs[1] = {src_host};
    return fnp_host_pack_xyz_multi_begin(srcs, &rows, 1, stride, xyz_offset, dst_host, n_threads);
}

extern "C" int fnp_host_pack_wait(int ticket)
{
    if (ticket < 0 || ticket >= kMaxPackJobs) return FNP_EINVAL;
    PackJob &job = g_jobs[ticket];
    {
        std::lock_guard<std::mutex> lk(g_jobs_mutex);
        if (!job.used) return FNP_EINVAL;
    }
    for (auto &t : job.threads) t.join();
    job.threads.clear();
    std::lock_guard<std::mutex> lk(g_jobs_mutex);
    job.used = false;
    return FNP_OK;
}

extern "C" int fnp_host_pack_xyz(const float *src_host, int64_t rows, int stride, int xyz_offset, float *dst_host,
                                 int n_threads)
{
    const int t = fnp_host_pack_xyz_begin(src_host, rows, stride, xyz_offset, dst_host, n_threads);
    if (t < 0) return t;
    return fnp_host_pack_wait(t);
}
