// target_loss.cuh -- YOLOLoss.forward(input, targets): get_target
// (models/yolo_loss.py:77-178) fused with weighted_mse_loss (:53-60, :219-236).
//
// The reference materialises targets / targets_weight (2 x (N,A,H,W,C+1)) and
// walks GT boxes in a Python loop with ~5 device syncs per assignment.  Only
// cells with weight > 0 contribute to the loss, so nothing dense is needed:
//
//   one CTA per image
//   P1  thread per GT box: xyxy (:112-113), IoU against ALL anchor shapes (:132),
//       best anchor = first argmax (:133), cell (gj, gi) (:128,136-137), assignment
//       flags k == index(best) or iou[mask[k]] > iou_thresh (:138-145); assigned
//       (GT, k) pairs are appended to a shared-memory list and their cells flagged.
//   P2  thread per cell, coalesced head reads: conf = sigmoid(tc); flagged cells
//       contribute (conf-1)^2 (:149-150); otherwise the decoded box (:84-92) is
//       tested against every GT box staged in shared memory: weight 1 / target 0
//       iff max_g IoU < ignore_threshold (:115-125), else the cell is ignored.
//   P3  thread per assignment: CIoU term (box_ciou :257-293), recall / iou / obj /
//       class-score stats (:151-169); the first assignment of each distinct cell
//       adds the class-channel loss with the union of assigned classes at 0.95,
//       the rest at 0.05 (class_loss :425-434 -- order independent, duplicates
//       counted exactly like the sequential reference).
//   P4  deterministic reduction: per-image partial sums -> workspace, then a
//       single-CTA kernel adds them in image order into sums[16].
//
// Loss normalisers are batch-global (:55, :224), so the kernel returns SUMS; the
// division happens after the (optional) cross-rank all-reduce, in
// b200yolo_loss_finalize.
#pragma once
#include "common.cuh"

namespace b200yolo {

constexpr int kTLThreads = 256;
constexpr int kTLWarps = kTLThreads / 32;
constexpr int kTLMaxGT = 1024;         // GT boxes per image staged in shared memory
constexpr int kTLMaxAllAnchors = 16;
constexpr int kTLMaxAnchors = 8;
constexpr int kTLSums = 16;

struct TLParams {
    const float *head;
    int N, A, C, attrs, H, W, HW, cells, NA;
    float invHW, invW, fW, fH;
    float aw_all[kTLMaxAllAnchors], ah_all[kTLMaxAllAnchors];
    int mask[kTLMaxAnchors];
    const float *gt;
    const int *gt_off;
    int G;
    float ignore_thr, iou_thr;
    double *sums;
    double *partial;  // [N][kTLSums]
    int *assign;      // [G][A][4] or null
    float *terms;     // [G][A][2] or null
    int *status;
};

struct TLAssign {
    uint32_t cell;  // (k*H + gj)*W + gi
    uint32_t t;     // GT index inside the image
};

__host__ __device__ inline uint32_t tl_smem_bytes(int cells) {
    uint32_t o = 0;
    o += 16 * kTLMaxGT;                       // gt xyxy
    o += 4 * kTLMaxGT;                        // gt area
    o += 4 * kTLMaxGT;                        // gt class (0-based)
    o += 8 * kTLMaxGT * kTLMaxAnchors / 2;    // assignment list (capacity 4 * kTLMaxGT entries)
    o += ((uint32_t)cells + 15u) / 16u * 16u; // assigned-cell flags
    o += 8 * kTLSums * kTLWarps;              // reduction scratch
    o += 64;
    return o;
}
constexpr int kTLMaxAssign = kTLMaxGT * kTLMaxAnchors / 2;

// decode one cell's box exactly like get_target (:84-92) + wh_to_x2y2 (:243-247)
__device__ __forceinline__ float4 tl_decode_box(const float *q, int HW, int i, int j, float fW, float fH, float aw,
                                                float ah) {
    const float tx = __ldg(q), ty = __ldg(q + HW), tw = __ldg(q + 2 * HW), th = __ldg(q + 3 * HW);
    const float sx = sigmoid_f(tx), sy = sigmoid_f(ty);
    const float ew = expf(tw), eh = expf(th);
    const float cx = __fdiv_rn(__fadd_rn(sx, (float)i), fW);
    const float cy = __fdiv_rn(__fadd_rn(sy, (float)j), fH);
    const float bw = __fmul_rn(ew, aw), bh = __fmul_rn(eh, ah);
    float4 b;
    b.x = __fsub_rn(cx, __fmul_rn(bw, 0.5f));
    b.y = __fsub_rn(cy, __fmul_rn(bh, 0.5f));
    b.z = __fadd_rn(bw, b.x);
    b.w = __fadd_rn(bh, b.y);
    return b;
}

// utils/iou.py:32-49 for one pair
__device__ __forceinline__ float tl_iou(const float4 &a, float area_a, const float4 &b, float area_b) {
    const float inter = pair_inter(a, b);
    return __fdiv_rn(inter, pair_union(area_a, area_b, inter));
}

// box_ciou (yolo_loss.py:257-293) with box1 = gt, box2 = pred; returns v = iou - term
__device__ __forceinline__ float tl_box_ciou(const float4 &b1, const float4 &b2, float *iou_out) {
    const float l = fminf(b1.x, b2.x), t = fminf(b1.y, b2.y);   // box_c :250-253
    const float r = fmaxf(b1.z, b2.z), bt = fmaxf(b1.w, b2.w);
    const float c = __fmul_rn(__fsub_rn(r, l), __fsub_rn(bt, t));  // :264
    const float iou = tl_iou(b1, box_area(b1), b2, box_area(b2));  // :265
    const float w1 = __fsub_rn(b1.z, b1.x), h1 = __fsub_rn(b1.w, b1.y);  // :267
    const float w2 = __fsub_rn(b2.z, b2.x), h2 = __fsub_rn(b2.w, b2.y);  // :268
    const float x1 = __fmul_rn(__fadd_rn(b1.z, b1.x), 0.5f), y1 = __fmul_rn(__fadd_rn(b1.y, b1.w), 0.5f);  // :269
    const float x2 = __fmul_rn(__fadd_rn(b2.z, b2.x), 0.5f), y2 = __fmul_rn(__fadd_rn(b2.y, b2.w), 0.5f);  // :270
    const float dx = __fsub_rn(x1, x2), dy = __fsub_rn(y1, y2);
    const float u = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));  // :272
    const float d = __fdiv_rn(u, c);                                  // :277
    const float ar_gt = __fdiv_rn(w2, h2), ar_pred = __fdiv_rn(w1, h1);  // :279-280
    const float k = 0.40528473456935109f;                             // 4/(pi*pi) -> fp32
    const float dl = __fsub_rn(atanf(ar_gt), atanf(ar_pred));
    const float ar_loss = __fmul_rn(__fmul_rn(k, dl), dl);            // :282
    const float alpha = __fdiv_rn(ar_loss, __fadd_rn(__fadd_rn(__fsub_rn(1.0f, iou), ar_loss), 0.000001f));  // :283
    float term = __fadd_rn(d, __fmul_rn(alpha, ar_loss));             // :284
    const float m = (c == 0.0f) ? 1.0f : 0.0f;                        // :286
    term = __fadd_rn(__fmul_rn(term, 1.0f - m), __fmul_rn(iou, m));   // :287
    *iou_out = iou;
    return __fsub_rn(iou, term);                                      // :293
}

// box_giou (yolo_loss.py:295-317; dead code upstream, kept as a device routine)
__device__ __forceinline__ float tl_box_giou(const float4 &b1, const float4 &b2, float *iou_out) {
    const float l = fminf(b1.x, b2.x), t = fminf(b1.y, b2.y);
    const float r = fmaxf(b1.z, b2.z), bt = fmaxf(b1.w, b2.w);
    const float c = __fmul_rn(__fsub_rn(r, l), __fsub_rn(bt, t));
    const float inter = pair_inter(b1, b2);
    const float u = pair_union(box_area(b1), box_area(b2), inter);
    const float iou = __fdiv_rn(inter, u);
    float term = __fdiv_rn(__fsub_rn(c, u), c);
    const float m = (c == 0.0f) ? 1.0f : 0.0f;
    term = __fadd_rn(__fmul_rn(term, 1.0f - m), __fmul_rn(iou, m));
    *iou_out = iou;
    return __fsub_rn(iou, term);
}

__global__ void __launch_bounds__(kTLThreads) target_loss_kernel(const TLParams p) {
    extern __shared__ __align__(16) unsigned char tl_smem[];
    float4 *s_gbox = reinterpret_cast<float4 *>(tl_smem);
    float *s_garea = reinterpret_cast<float *>(tl_smem + 16 * kTLMaxGT);
    int *s_gcls = reinterpret_cast<int *>(tl_smem + 20 * kTLMaxGT);
    TLAssign *s_list = reinterpret_cast<TLAssign *>(tl_smem + 24 * kTLMaxGT);
    uint8_t *s_flag = reinterpret_cast<uint8_t *>(tl_smem + 24 * kTLMaxGT + 8 * kTLMaxAssign);
    const uint32_t flag_bytes = ((uint32_t)p.cells + 15u) / 16u * 16u;
    double *s_red = reinterpret_cast<double *>(tl_smem + 24 * kTLMaxGT + 8 * kTLMaxAssign + flag_bytes);
    int *s_misc = reinterpret_cast<int *>(s_red + kTLSums * kTLWarps);

    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g0 = p.gt_off[b];
    int nG = p.gt_off[b + 1] - g0;
    const int HW = p.HW, W = p.W, H = p.H, A = p.A, C = p.C;

    if (tid == 0) s_misc[0] = 0;  // assignment list length
    for (int c = tid; c < (int)flag_bytes; c += kTLThreads) s_flag[c] = 0;
    if (nG > kTLMaxGT) {
        if (tid == 0) atomicMax(p.status, 2);  // too many GT boxes for one image
        nG = 0;                                // (the shim raises; keep the kernel well defined)
    }
    __syncthreads();

    double acc[12];
#pragma unroll
    for (int q = 0; q < 12; ++q) acc[q] = 0.0;

    // ---------------- P1: per-GT anchor matching ----------------
    for (int t = tid; t < nG; t += kTLThreads) {
        const float *g = p.gt + 5 * (size_t)(g0 + t);
        const float gc = __ldg(g), gx = __ldg(g + 1), gy = __ldg(g + 2), gw = __ldg(g + 3), gh = __ldg(g + 4);
        float4 bx;                                                  // :112-113 wh_to_x2y2
        bx.x = __fsub_rn(gx, __fmul_rn(gw, 0.5f));
        bx.y = __fsub_rn(gy, __fmul_rn(gh, 0.5f));
        bx.z = __fadd_rn(gw, bx.x);
        bx.w = __fadd_rn(gh, bx.y);
        s_gbox[t] = bx;
        s_garea[t] = box_area(bx);
        const int cls = (int)__fsub_rn(gc, 1.0f);                   // :131,147
        s_gcls[t] = cls;
        const int gi = (int)__fmul_rn(gx, p.fW), gj = (int)__fmul_rn(gy, p.fH);  // :128,136-137
        const bool ok = gi >= 0 && gi < W && gj >= 0 && gj < H && cls >= 0 && cls < C;
        if (!ok) atomicMax(p.status, 1);
        // anchor-vs-GT IoU on (0,0,w,h) shapes, ALL anchors (:129-133)
        const float4 gb = make_float4(0.f, 0.f, gw, gh);
        const float ga = box_area(gb);
        float best = 0.f;
        int best_n = 0;
        unsigned over = 0u;  // bit n: iou[n] > iou_thresh
        for (int n = 0; n < p.NA; ++n) {
            const float4 ab = make_float4(0.f, 0.f, p.aw_all[n], p.ah_all[n]);
            const float v = tl_iou(gb, ga, ab, box_area(ab));
            if (n == 0 || v > best) { best = v; best_n = n; }       // argmax = first maximum
            if (v > p.iou_thr) over |= 1u << n;                      // :139
        }
        for (int k = 0; k < A; ++k) {
            const bool asg = ok && (p.mask[k] == best_n || ((over >> p.mask[k]) & 1u));  // :141-145
            if (p.assign) {
                int *r = p.assign + ((size_t)(g0 + t) * A + k) * 4;
                r[0] = asg ? 1 : 0; r[1] = gj; r[2] = gi; r[3] = best_n;
            }
            if (p.terms && !asg) {
                float *r = p.terms + ((size_t)(g0 + t) * A + k) * 2;
                r[0] = 0.f; r[1] = 0.f;
            }
            if (asg) {
                const uint32_t cell = (uint32_t)((k * H + gj) * W + gi);
                const int e = atomicAdd(&s_misc[0], 1);
                if (e < kTLMaxAssign) {
                    s_list[e].cell = cell;
                    s_list[e].t = (uint32_t)t;
                } else {
                    atomicMax(p.status, 2);
                }
                s_flag[cell] = 1;
            }
        }
    }
    __syncthreads();
    const int nE = min(s_misc[0], kTLMaxAssign);

    // ---------------- P2: per-cell objectness / ignore mask ----------------
    const int cells_pad = (p.cells + 31) & ~31;
    for (int cell = tid; cell < cells_pad; cell += kTLThreads) {
        const bool valid = cell < p.cells;
        bool undecided = false;  // still "max IoU < ignore_threshold so far"
        float conf = 0.f;
        float4 pb = make_float4(0.f, 0.f, 0.f, 0.f);
        float pa = 0.f;
        if (valid) {
            const int a = (int)(((float)cell + 0.5f) * p.invHW);
            const int pos = cell - a * HW;
            const float *q = p.head + ((size_t)(b * A + a) * p.attrs) * HW + pos;
            conf = sigmoid_f(__ldg(q + 4 * (size_t)HW));            // output[...,0] :87
            acc[B200YOLO_S_CONF_ALL] += (double)conf;               // :98
            if (s_flag[cell]) {                                     // :149-150 target 1, weight 1
                const float df = __fsub_rn(conf, 1.0f);
                acc[B200YOLO_S_SQW] += (double)__fmul_rn(df, df);
                acc[B200YOLO_S_W] += 1.0;
            } else if (nG == 0) {                                   // :108-111
                acc[B200YOLO_S_SQW] += (double)__fmul_rn(conf, conf);
                acc[B200YOLO_S_W] += 1.0;
            } else {
                const int j = (int)(((float)pos + 0.5f) * p.invW);
                const int i = pos - j * W;
                pb = tl_decode_box(q, HW, i, j, p.fW, p.fH, p.aw_all[p.mask[a]], p.ah_all[p.mask[a]]);
                pa = box_area(pb);
                undecided = true;
            }
        }
        if (nG > 0) {
            bool below = undecided;
            for (int t = 0; t < nG; ++t) {
                if (!__any_sync(kFullMask, below)) break;
                if (below) {
                    const float4 gb = s_gbox[t];
                    const float inter = pair_inter(gb, pb);
                    const float u = pair_union(s_garea[t], pa, inter);
                    // iou < thr, decided without the divide unless within 4e-6 of the threshold
                    bool lt;
                    const float d = __fmaf_rn(-p.ignore_thr, u, inter);
                    const float e = __fmul_rn(4e-6f, u);
                    if (u > 0.0f && d < -e) lt = true;
                    else if (u > 0.0f && d > e) lt = false;
                    else lt = __fdiv_rn(inter, u) < p.ignore_thr;   // NaN -> false (torch.max propagates NaN)
                    below = lt;
                }
            }
            if (undecided && below) {                               // :123-125 weight 1, target 0
                acc[B200YOLO_S_SQW] += (double)__fmul_rn(conf, conf);
                acc[B200YOLO_S_W] += 1.0;
            }
        }
    }

    // ---------------- P3: per-assignment terms ----------------
    for (int e = tid; e < nE; e += kTLThreads) {
        const uint32_t cell = s_list[e].cell;
        const int t = (int)s_list[e].t;
        const int a = (int)(((float)cell + 0.5f) * p.invHW);
        const int pos = (int)cell - a * HW;
        const int j = (int)(((float)pos + 0.5f) * p.invW);
        const int i = pos - j * W;
        const float *q = p.head + ((size_t)(b * A + a) * p.attrs) * HW + pos;
        const float4 pb = tl_decode_box(q, HW, i, j, p.fW, p.fH, p.aw_all[p.mask[a]], p.ah_all[p.mask[a]]);
        const float conf = sigmoid_f(__ldg(q + 4 * (size_t)HW));
        const float4 gb = s_gbox[t];
        float iou;
        const float v = tl_box_ciou(gb, pb, &iou);                  // :157
        const float wt = __fsub_rn(2.0f, s_garea[t]);               // :160
        const float dv = __fsub_rn(v, 1.0f);
        acc[B200YOLO_S_IOU_SQ] += (double)__fmul_rn(dv, dv);
        acc[B200YOLO_S_IOU_W] += (double)wt;
        acc[B200YOLO_S_NASSIGN] += 1.0;                             // :146
        acc[B200YOLO_S_OBJ] += (double)conf;                        // :152
        acc[B200YOLO_S_IOU] += (double)iou;                         // :165
        if (iou > p.ignore_thr) acc[B200YOLO_S_RECALL] += 1.0;      // :163
        const int cls = s_gcls[t];
        acc[B200YOLO_S_CLS] += (double)sigmoid_f(__ldg(q + (size_t)(5 + cls) * HW));  // :169
        if (p.terms) {
            // locate k: cell = (k*H+gj)*W+gi -> k == a
            float *r = p.terms + ((size_t)(g0 + t) * A + a) * 2;
            r[0] = v; r[1] = iou;
        }
        // class channels: once per distinct cell, by its first list entry
        bool first = true;
        for (int f = 0; f < e; ++f)
            if (s_list[f].cell == cell) { first = false; break; }
        if (first) {
            double sq = 0.0;
            for (int c = 0; c < C; ++c) {
                const float o = sigmoid_f(__ldg(q + (size_t)(5 + c) * HW));
                bool hit = (c == cls);
                if (!hit)
                    for (int f = e + 1; f < nE; ++f)
                        if (s_list[f].cell == cell && s_gcls[s_list[f].t] == c) { hit = true; break; }
                const float tv = hit ? 0.95f : 0.05f;               // :426-433
                const float df = __fsub_rn(o, tv);
                sq += (double)__fmul_rn(df, df);
            }
            acc[B200YOLO_S_SQW] += sq;
            acc[B200YOLO_S_W] += (double)C;
        }
    }

    // ---------------- P4: block reduction -> per-image partials ----------------
#pragma unroll
    for (int q = 0; q < 10; ++q) {
        double v = acc[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
        if (lane == 0) s_red[q * kTLWarps + warp] = v;
    }
    __syncthreads();
    if (tid < kTLSums) {
        double v = 0.0;
        if (tid < 10)
            for (int w = 0; w < kTLWarps; ++w) v += s_red[tid * kTLWarps + w];
        else if (tid == B200YOLO_S_NCELLS) v = (double)p.cells;
        else if (tid == B200YOLO_S_NIMG) v = 1.0;
        p.partial[(size_t)b * kTLSums + tid] = v;
    }
}

// fixed-order sum over images -> sums[16] (bitwise reproducible run to run)
__global__ void __launch_bounds__(kTLSums * 32) target_loss_reduce_kernel(const double *partial, int N, double *sums) {
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double v = 0.0;
    for (int b = lane; b < N; b += 32) v += partial[(size_t)b * kTLSums + q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    if (lane == 0) sums[q] = v;
}

}  // namespace b200yolo
