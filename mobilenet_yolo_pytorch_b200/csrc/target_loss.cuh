// target_loss.cuh -- YOLOLoss.forward(input, targets): get_target
// (models/yolo_loss.py:77-178) fused with weighted_mse_loss (:53-60, :219-236).
//
// The reference materialises targets / targets_weight (2 x (N,A,H,W,C+1)) and
// walks GT boxes in a Python loop with ~5 device syncs per assignment.  Only
// cells with weight > 0 contribute to the loss, so nothing dense is needed:
//
//   grid N*S: S CTAs per image, each owning a slice of the image's cells
//   P1  thread per GT box (replicated in the S CTAs of an image): xyxy (:112-113), IoU
//       against ALL anchor shapes (:132), best anchor = first argmax (:133), cell (gj, gi)
//       (:128,136-137), assignment flags k == index(best) or iou[mask[k]] > iou_thresh
//       (:138-145); assigned (GT, k) pairs go to a shared-memory list in (GT, k) order (block
//       scan, so every sum is the same run to run) and their cells are flagged; duplicate
//       assignments of a cell are chained and the distinct cells listed.
//   P2  thread per cell, coalesced head reads: conf = sigmoid(tc); flagged cells
//       contribute (conf-1)^2 (:149-150); otherwise the decoded box (:84-92) is tested
//       against every GT box staged in shared memory: weight 1 / target 0 iff
//       max_g IoU < ignore_threshold (:115-125), else the cell is ignored.  This is the
//       hot loop (config 4: 92.9 M box pairs) and it is divide-free and branch-free:
//       iou < thr <=> inter < t*(area_g+area_p), t = thr/(1+thr); d = t*(areas) - w*h is
//       one FFMA, the cell keeps min d and min(|d| - 1e-5*t*(areas)); only a cell with a pair
//       closer than 1e-5 to the threshold (or a degenerate box) re-runs its GT loop with
//       the reference's IEEE arithmetic (inter/union < thr).
//   P3  (first CTA of the image) thread per assignment: CIoU term (box_ciou :257-293),
//       recall / iou / obj / class-score stats (:151-169); then thread per (distinct assigned
//       cell, class): the class-channel loss with the union of
//       assigned classes at 0.95, the rest at 0.05 (class_loss :425-434 -- order
//       independent, duplicates counted exactly like the sequential reference).
//   P4  deterministic reduction: per-CTA partial sums -> workspace, then a
//       single-CTA kernel adds them in a fixed order into sums[16].
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
constexpr int kTLMaxSplit = 8;         // CTAs per image
constexpr float kTLEps = 1e-5f;
constexpr int kTLStatusSlot = 12;      // slot of the per-CTA partial sums that carries the status (max-reduced)

struct TLParams {
    const float *head;
    int N, A, C, attrs, H, W, HW, cells, NA;
    int S, chunk;  // CTAs per image; cells per CTA (multiple of 32)
    int gcap;      // GT boxes per image the shared-memory staging holds (<= kTLMaxGT); the list holds A*gcap assignments
    float invHW, invW, fW, fH;
    float aw_all[kTLMaxAllAnchors], ah_all[kTLMaxAllAnchors];
    int mask[kTLMaxAnchors];
    const float *gt;
    const int *gt_off;
    int G;
    float ignore_thr, iou_thr;
    float ts;        // ignore_thr/(1+ignore_thr) * 2^-13 (0: the divide-free test is off)
    float th16;      // ignore_thr/(1+ignore_thr) * 64 * (1 - 2^-8): the fp16 prefilter's area factor (0: prefilter off)
    double *sums;
    double *partial;  // [N][S][kTLSums]
    int *assign;      // [G][A][4] or null
    float *terms;     // [G][A][2] or null
    int *status;
    unsigned char *cell_state;  // [N][cells] or null: 0 ignored, 1 no-object (weight 1, target 0), 2 assigned
    // backward only
    const unsigned char *cell_state_in;
    const float *grad_out;      // device scalar d(total)/d(loss), or null = 1
    float *grad_input;          // (N, A*(5+C), H, W)
    float iou_weighting;
    int wait_inputs;            // 1: griddepcontrol.wait before the first global read (the predecessor in the stream may
                                // have produced head / gt); 0: the caller vouches for the inputs (b200yolo_set_inputs_ready)
};

struct TLAssign {
    uint32_t cell;  // (k*H + gj)*W + gi
    uint32_t t;     // GT index inside the image
};

// Shared-memory layout of both kernels
struct TLSmem {
    float4 *gbox;      // [gcap] GT xyxy
    float *garea;      // [gcap]
    float *gta;        // [gcap] t * area * 2^-13
    int *gcls;         // [gcap] class (0-based)
    int2 *gtmp;        // [gcap] pass-1 result of the matching: {cell base gj*W+gi, assignment mask over k}
    TLAssign *list;    // [A*gcap] assignments in (GT, k) order
    uint16_t *ucell;   // [A*gcap] list index of the first assignment of every distinct cell
    uint8_t *flag;     // [cells] cell is assigned
    double *red;       // [kTLSums][kTLWarps]
    int *misc;         // [16]: 0 nE, 1 any degenerate GT, 2 status, 3 nU, 4..11 warp totals of the block scans
    H16Tile *gh16;     // [ceil(gcap/32)] conservative fp16 images of the GT boxes (decode_nms.cuh, h16_prefilter)
    double4 *contrib;  // backward: [A*gcap] box-gradient contribution of every assignment
};

__host__ __device__ inline uint32_t tl_up16(uint32_t v) { return (v + 15u) / 16u * 16u; }

__host__ __device__ inline uint32_t tl_smem_bytes(int cells, int gcap, int A, bool backward = false) {
    const uint32_t g = (uint32_t)gcap, l = (uint32_t)(gcap * A);
    uint32_t o = 36 * g + 8 * l + tl_up16(2 * l) + tl_up16((uint32_t)cells) + 8 * kTLSums * kTLWarps + 64;
    o = tl_up16(o) + (uint32_t)sizeof(H16Tile) * ((g + 31) / 32);
    if (backward) o += 32 * l;
    return o;
}

__device__ __forceinline__ TLSmem tl_carve(unsigned char *base, int cells, int gcap, int A) {
    const uint32_t g = (uint32_t)gcap, l = (uint32_t)(gcap * A);
    TLSmem s;
    uint32_t o = 0;
    s.gbox = reinterpret_cast<float4 *>(base + o); o += 16 * g;
    s.garea = reinterpret_cast<float *>(base + o); o += 4 * g;
    s.gta = reinterpret_cast<float *>(base + o); o += 4 * g;
    s.gcls = reinterpret_cast<int *>(base + o); o += 4 * g;
    s.gtmp = reinterpret_cast<int2 *>(base + o); o += 8 * g;
    s.list = reinterpret_cast<TLAssign *>(base + o); o += 8 * l;
    s.ucell = reinterpret_cast<uint16_t *>(base + o); o += tl_up16(2 * l);
    s.flag = reinterpret_cast<uint8_t *>(base + o); o += tl_up16((uint32_t)cells);
    s.red = reinterpret_cast<double *>(base + o); o += 8 * kTLSums * kTLWarps;
    s.misc = reinterpret_cast<int *>(base + o); o += 64;
    o = tl_up16(o);
    s.gh16 = reinterpret_cast<H16Tile *>(base + o); o += (uint32_t)sizeof(H16Tile) * ((g + 31) / 32);
    s.contrib = reinterpret_cast<double4 *>(base + o);
    return s;
}

// s_list[e].t packs the GT index (bits 0-10) and 1 + the list index of the next assignment of the same cell (bits 11-24)
constexpr uint32_t kTLGtMask = 0x7ffu;
__device__ __forceinline__ int tl_entry_gt(uint32_t tn) { return (int)(tn & kTLGtMask); }
__device__ __forceinline__ int tl_entry_next(uint32_t tn) { return (int)(tn >> 11) - 1; }

// block-wide exclusive scan of one int per thread (warp totals in misc[4..11]); returns the exclusive prefix and the
// block total.  Two barriers; every thread of the CTA must call it.
__device__ __forceinline__ int tl_block_scan(int v, int *misc, int *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int inc = warp_inclusive_scan(v, lane);
    __syncthreads();  // (misc[4..11] free again)
    if (lane == 31) misc[4 + warp] = inc;
    __syncthreads();
    int before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < kTLWarps; ++w) {
        const int t = misc[4 + w];
        before += (w < warp) ? t : 0;
        all += t;
    }
    *total = all;
    return before + inc - v;
}

// The matching runs either on the whole CTA (NT = kTLThreads) or on warp 0 alone (NT = 32: the backward kernel, whose
// other warps stream the dense gradient planes meanwhile): scan and barrier of the cooperating group
template <int NT>
__device__ __forceinline__ int tl_group_scan(int v, int *misc, int *total) {
    if (NT == 32) {
        const int lane = threadIdx.x & 31;
        const int inc = warp_inclusive_scan(v, lane);
        *total = __shfl_sync(kFullMask, inc, 31);
        return inc - v;
    }
    return tl_block_scan(v, misc, total);
}
template <int NT>
__device__ __forceinline__ void tl_group_sync() {
    if (NT == 32) __syncwarp();
    else __syncthreads();
}

// decode one cell's box exactly like get_target (:84-92) + wh_to_x2y2 (:243-247)
__device__ __forceinline__ float4 tl_decode_box(float tx, float ty, float tw, float th, int i, int j, float fW, float fH,
                                                float aw, float ah) {
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

// conservative fp16 image of a box (decode_nms.cuh: h16_store / h16_prefilter): outward-rounded coordinates and a
// lower bound of t*area*64; boxes the fp16 format cannot hold get TA = -inf ("always maybe")
__device__ __forceinline__ void tl_h16_put(H16Tile *tiles, int idx, const float4 &b, float area, float th16) {
    H16Tile &tile = tiles[idx >> 5];
    const int k = idx & 15, hi = (idx >> 4) & 1;
    const float big = fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w)));
    const bool ok = th16 > 0.0f && area >= 3.814697265625e-06f && area <= 256.0f && big <= 8.0f;   // (false for NaN)
    __half *xy = reinterpret_cast<__half *>(&tile.xy[k]);
    xy[0 + hi] = ok ? __float2half_rd(__fmul_rn(b.x, kH16SX)) : __ushort_as_half((unsigned short)0);
    xy[2 + hi] = ok ? __float2half_rd(__fmul_rn(b.y, kH16SY)) : __ushort_as_half((unsigned short)0);
    xy[4 + hi] = ok ? __float2half_ru(__fmul_rn(b.z, kH16SX)) : __ushort_as_half((unsigned short)0);
    xy[6 + hi] = ok ? __float2half_ru(__fmul_rn(b.w, kH16SY)) : __ushort_as_half((unsigned short)0);
    reinterpret_cast<__half *>(&tile.ta[k])[hi] = ok ? __float2half_rd(__fmul_rn(area, th16)) : __ushort_as_half((unsigned short)0xfc00);
}

__device__ __forceinline__ H16Row tl_h16_row(const float4 &b, float area, float th16) {
    const float big = fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w)));
    const bool ok = th16 > 0.0f && area >= 3.814697265625e-06f && area <= 256.0f && big <= 8.0f;
    auto dup = [](__half h) { const uint32_t u = (uint32_t)__half_as_ushort(h); return u | (u << 16); };
    H16Row r;
    r.x1 = ok ? dup(__float2half_rd(__fmul_rn(b.x, kH16SX))) : 0u;
    r.y1 = ok ? dup(__float2half_rd(__fmul_rn(b.y, kH16SY))) : 0u;
    r.x2 = ok ? dup(__float2half_ru(__fmul_rn(b.z, kH16SX))) : 0u;
    r.y2 = ok ? dup(__float2half_ru(__fmul_rn(b.w, kH16SY))) : 0u;
    r.nta = ok ? (dup(__float2half_rd(__fmul_rn(area, th16))) ^ 0x80008000u) : 0x7c007c00u;   // -TA, or +inf
    return r;
}

// t * area * 2^-13 for the divide-free test; NaN when the box must take the exact path
__device__ __forceinline__ float tl_ta(const float4 &b, float area, float ts) {
    const float big = fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w)));
    const bool ok = ts > 0.0f && area >= 1e-20f && area <= 1e20f && big < 4096.0f;
    return ok ? __fmul_rn(area, ts) : __int_as_float(0x7fc00000);
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

// exact max_g IoU < ignore_threshold (:115-125): torch.max propagates NaN, NaN < thr is false
__device__ __noinline__ bool tl_below_exact(const float4 *gbox, const float *garea, int nG, const float4 pb, float pa, float thr) {
    bool below = true;
    for (int t = 0; t < nG; ++t) {
        const float4 gb = gbox[t];
        const float inter = pair_inter(gb, pb);
        const float u = pair_union(garea[t], pa, inter);
        below = below && (__fdiv_rn(inter, u) < thr);
    }
    return below;
}

// P1: per-GT anchor matching (yolo_loss.py:112-113, 127-145): stages the image's GT boxes, then appends the assigned
// (GT, k) pairs to the list in (GT, k) order (block scan: the list, and with it every sum, is the same run to run) and
// flags their cells.  `lead` CTAs also write the per-GT outputs and the status (misc[2]; it travels to the host through
// slot 12 of the CTA's partial sums).  Ends with the list complete and visible (trailing barrier); misc[0] = length.
template <int NT>
__device__ __forceinline__ void tl_match_gt(const TLParams &p, int g0, int nG, bool lead, const TLSmem &s) {
    const int tid = threadIdx.x;
    const int W = p.W, H = p.H, A = p.A, C = p.C;
    int carry = 0;
    for (int t0 = 0; t0 < nG; t0 += NT) {  // (uniform trip count)
        const int t = t0 + tid;
        int cellbase = 0;
        unsigned amask = 0u;
        if (t < nG) {
            const float *g = p.gt + 5 * (size_t)(g0 + t);
            const float gc = __ldg(g), gx = __ldg(g + 1), gy = __ldg(g + 2), gw = __ldg(g + 3), gh = __ldg(g + 4);
            float4 bx;                                                  // :112-113 wh_to_x2y2
            bx.x = __fsub_rn(gx, __fmul_rn(gw, 0.5f));
            bx.y = __fsub_rn(gy, __fmul_rn(gh, 0.5f));
            bx.z = __fadd_rn(gw, bx.x);
            bx.w = __fadd_rn(gh, bx.y);
            const float barea = box_area(bx);
            const float bta = tl_ta(bx, barea, p.ts);
            s.gbox[t] = bx;
            s.garea[t] = barea;
            s.gta[t] = bta;
            tl_h16_put(s.gh16, t, bx, barea, p.th16);
            if (bta != bta) s.misc[1] = 1;
            const int cls = (int)__fsub_rn(gc, 1.0f);                   // :131,147
            s.gcls[t] = cls;
            const int gi = (int)__fmul_rn(gx, p.fW), gj = (int)__fmul_rn(gy, p.fH);  // :128,136-137
            const bool ok = gi >= 0 && gi < W && gj >= 0 && gj < H && cls >= 0 && cls < C;
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
                const bool want = (p.mask[k] == best_n || ((over >> p.mask[k]) & 1u));      // :141-145
                // The reference indexes [gj, gi] and the class only for a GT it assigns on THIS head (:146-169): a box
                // outside the grid that another head owns trains fine upstream, so only an assigned one is an error
                // (status 1 -> IndexError in the shim).  Divergence kept on purpose: negative cells / class 0 wrap
                // around silently in the reference (python negative indexing); here they are errors too.
                if (want && !ok && lead) atomicMax(&s.misc[2], 1);
                const bool asg = ok && want;
                if (asg) amask |= 1u << k;
                if (p.assign && lead) {
                    int *r = p.assign + ((size_t)(g0 + t) * A + k) * 4;
                    r[0] = asg ? 1 : 0; r[1] = gj; r[2] = gi; r[3] = best_n;
                }
                if (p.terms && lead && !asg) {
                    float *r = p.terms + ((size_t)(g0 + t) * A + k) * 2;
                    r[0] = 0.f; r[1] = 0.f;
                }
            }
            cellbase = gj * W + gi;
        }
        int total;
        int e = carry + tl_group_scan<NT>(__popc(amask), s.misc, &total);
        for (int k = 0; k < A; ++k) {
            if ((amask >> k) & 1u) {
                const uint32_t cell = (uint32_t)(k * H * W + cellbase);
                s.list[e].cell = cell;
                s.list[e].t = (uint32_t)t;
                s.flag[cell] = 1;
                ++e;
            }
        }
        carry += total;
    }
    if (tid == 0) s.misc[0] = carry;
    tl_group_sync<NT>();
}

// Duplicate chains and the list of distinct assigned cells (in list order): list[e].t gets the index of the next
// assignment of the same cell; ucell[u] = index of the first assignment of the u-th distinct cell; misc[3] = count.
// Every thread of the CTA must call it; ends with a barrier.
template <int NT>
__device__ __forceinline__ void tl_unique_cells(int nE, const TLSmem &s) {
    const int tid = threadIdx.x;
    int carry = 0;
    for (int e0 = 0; e0 < nE; e0 += NT) {
        const int e = e0 + tid;
        int first = 0;
        if (e < nE) {
            const uint32_t cell = s.list[e].cell;
            int nx = -1;
            for (int f = e + 1; f < nE; ++f)
                if (s.list[f].cell == cell) { nx = f; break; }
            first = 1;
            for (int f = 0; f < e; ++f)
                if (s.list[f].cell == cell) { first = 0; break; }
            s.list[e].t |= (uint32_t)(nx + 1) << 11;
        }
        int total;
        const int u = carry + tl_group_scan<NT>(first, s.misc, &total);
        if (first) s.ucell[u] = (uint16_t)e;
        carry += total;
    }
    if (tid == 0) s.misc[3] = carry;
    tl_group_sync<NT>();
}

__global__ void __launch_bounds__(kTLThreads) target_loss_kernel(const TLParams p) {
    extern __shared__ __align__(16) unsigned char tl_smem[];
    const TLSmem sm = tl_carve(tl_smem, p.cells, p.gcap, p.A);
    float4 *s_gbox = sm.gbox;
    float *s_garea = sm.garea, *s_gta = sm.gta;
    int *s_gcls = sm.gcls;
    TLAssign *s_list = sm.list;
    uint8_t *s_flag = sm.flag;
    double *s_red = sm.red;
    int *s_misc = sm.misc;
    const uint32_t flag_bytes = tl_up16((uint32_t)p.cells);

    const int b = blockIdx.x / p.S, split = blockIdx.x - b * p.S;
    const bool lead = (split == 0);  // the CTA of the image that owns the per-GT outputs and the assignments
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int HW = p.HW, W = p.W, A = p.A, C = p.C;

    // Programmatic dependent launch (see decode_nms.cuh): the next kernel of the stream may start now.  This one waits
    // for its predecessor before its first global READ unless the caller has declared the inputs ready (the kernel
    // before it in the stream may be the producer of head / gt, and its writes are only guaranteed visible after
    // griddepcontrol.wait), and in any case before its first global store (the workspace of partial sums is shared
    // by consecutive calls; the optional per-GT / per-cell outputs are written early, so they wait here).
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (p.wait_inputs || p.assign || p.terms || p.cell_state) asm volatile("griddepcontrol.wait;" ::: "memory");
    const int g0 = p.gt_off[b];
    int nG = p.gt_off[b + 1] - g0;
    if (tid < 4) s_misc[tid] = 0;  // assignment list length; any degenerate GT box; status; distinct cells
    for (int c = tid; c < (int)flag_bytes; c += kTLThreads) s_flag[c] = 0;
    __syncthreads();
    if (nG > p.gcap) {
        if (tid == 0 && lead) s_misc[2] = 2;  // more GT boxes in one image than the staging holds
        nG = 0;                               // (the shim raises; keep the kernel well defined)
    }

    double acc[10];
#pragma unroll
    for (int q = 0; q < 10; ++q) acc[q] = 0.0;

    tl_match_gt<kTLThreads>(p, g0, nG, lead, sm);
    const int nE = s_misc[0];
    const bool gt_degenerate = s_misc[1] != 0;
    if (lead) tl_unique_cells<kTLThreads>(nE, sm);  // (lead is uniform in the CTA)
    const int nU = s_misc[3];

    // ---------------- P2: per-cell objectness / ignore mask ----------------
    const int cell_lo = split * p.chunk, cell_hi = min(cell_lo + p.chunk, p.cells);
    for (int cell = cell_lo + tid; cell < cell_hi; cell += kTLThreads) {
        const int a = (int)(((float)cell + 0.5f) * p.invHW);
        const int pos = cell - a * HW;
        const float *q = p.head + ((size_t)(b * A + a) * p.attrs) * HW + pos;
        const bool flagged = s_flag[cell] != 0;
        const bool need_box = !flagged && nG > 0;
        float tx = 0.f, ty = 0.f, tw = 0.f, th = 0.f;
        if (need_box) {
            tx = __ldcs(q);
            ty = __ldcs(q + HW);
            tw = __ldcs(q + 2 * (size_t)HW);
            th = __ldcs(q + 3 * (size_t)HW);
        }
        const float conf = sigmoid_f(__ldg(q + 4 * (size_t)HW));    // output[...,0] :87
        acc[B200YOLO_S_CONF_ALL] += (double)conf;                   // :98
        unsigned char state = 1;
        if (flagged) {                                              // :149-150 target 1, weight 1
            const float df = __fsub_rn(conf, 1.0f);
            acc[B200YOLO_S_SQW] += (double)__fmul_rn(df, df);
            acc[B200YOLO_S_W] += 1.0;
            state = 2;
        } else if (nG == 0) {                                       // :108-111
            acc[B200YOLO_S_SQW] += (double)__fmul_rn(conf, conf);
            acc[B200YOLO_S_W] += 1.0;
        } else {
            const int j = (int)(((float)pos + 0.5f) * p.invW);
            const int i = pos - j * W;
            const float4 pb = tl_decode_box(tx, ty, tw, th, i, j, p.fW, p.fH, p.aw_all[p.mask[a]], p.ah_all[p.mask[a]]);
            const float pa = box_area(pb);
            const float pta = tl_ta(pb, pa, p.ts);
            // Pass over the GT boxes.  First the conservative fp16 prefilter of decode_nms.cuh, two GT boxes per
            // instruction: a clear bit PROVES iou < ignore_thr * (1 - 0.003) for that box; then, only for the boxes it
            // cannot rule out, the divide-free fp32 test (d > 0 <=> iou < thr) with its guard band.
            float dmin = INFINITY, m = INFINITY;
            const H16Row row = tl_h16_row(pb, pa, p.th16);
            for (int t0 = 0; t0 < nG; t0 += 32) {
                const int ncol = min(32, nG - t0);
                uint32_t maybe = h16_prefilter(sm.gh16[t0 >> 5], row, ncol) & ((ncol >= 32) ? 0xffffffffu : ((1u << ncol) - 1u));
                while (maybe) {
                    const int t = t0 + __ffs(maybe) - 1;
                    maybe &= maybe - 1u;
                    const float4 gb = s_gbox[t];
                    const float w = __fsub_rn(fminf(pb.z, gb.z), fmaxf(pb.x, gb.x));
                    const float h = __fsub_rn(fminf(pb.w, gb.w), fmaxf(pb.y, gb.y));
                    const float ws = __saturatef(__fmul_rn(w, 1.220703125e-4f));  // max(w,0) * 2^-13
                    const float sum = __fadd_rn(pta, s_gta[t]);
                    const float d = __fmaf_rn(-ws, h, sum);
                    dmin = fminf(dmin, d);
                    m = fminf(m, __fmaf_rn(sum, -kTLEps, fabsf(d)));
                }
            }
            bool below;
            if (gt_degenerate || pta != pta || !(m > 0.0f)) below = tl_below_exact(s_gbox, s_garea, nG, pb, pa, p.ignore_thr);
            else below = dmin > 0.0f;
            if (below) {                                            // :123-125 weight 1, target 0
                acc[B200YOLO_S_SQW] += (double)__fmul_rn(conf, conf);
                acc[B200YOLO_S_W] += 1.0;
            } else {
                state = 0;
            }
        }
        if (p.cell_state) p.cell_state[(size_t)b * p.cells + cell] = state;
    }

    // ---------------- P3a: per-assignment terms (thread per assignment) ----------------
    if (lead) {
        for (int e = tid; e < nE; e += kTLThreads) {
            const uint32_t cell = s_list[e].cell;
            const int t = tl_entry_gt(s_list[e].t);
            const int a = (int)(((float)cell + 0.5f) * p.invHW);
            const int pos = (int)cell - a * HW;
            const int j = (int)(((float)pos + 0.5f) * p.invW);
            const int i = pos - j * W;
            const float *q = p.head + ((size_t)(b * A + a) * p.attrs) * HW + pos;
            const int cls = s_gcls[t];
            const float tx = __ldg(q), ty = __ldg(q + HW), tw = __ldg(q + 2 * (size_t)HW), th = __ldg(q + 3 * (size_t)HW);
            const float tc = __ldg(q + 4 * (size_t)HW), tk = __ldg(q + (size_t)(5 + cls) * HW);
            const float4 pb = tl_decode_box(tx, ty, tw, th, i, j, p.fW, p.fH, p.aw_all[p.mask[a]], p.ah_all[p.mask[a]]);
            const float conf = sigmoid_f(tc);
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
            acc[B200YOLO_S_CLS] += (double)sigmoid_f(tk);               // :169
            if (p.terms) {
                float *r = p.terms + ((size_t)(g0 + t) * A + a) * 2;    // cell = (k*H+gj)*W+gi -> k == a
                r[0] = v; r[1] = iou;
            }
        }
        // ------------ P3b: class channels, once per distinct cell: thread per (cell, class), all logit loads in flight
        for (int it = tid; it < nU * C; it += kTLThreads) {
            const int u = it / C, c = it - u * C;
            int f = sm.ucell[u];
            const uint32_t cell = s_list[f].cell;
            const int a = (int)(((float)cell + 0.5f) * p.invHW);
            const int pos = (int)cell - a * HW;
            const float o = sigmoid_f(__ldg(p.head + ((size_t)(b * A + a) * p.attrs + 5 + c) * HW + pos));
            bool hit = false;                                           // is c one of the classes assigned to the cell
            while (f >= 0) {
                const uint32_t tn = s_list[f].t;
                hit = hit || (s_gcls[tl_entry_gt(tn)] == c);
                f = tl_entry_next(tn);
            }
            const float df = __fsub_rn(o, hit ? 0.95f : 0.05f);         // :426-433
            acc[B200YOLO_S_SQW] += (double)__fmul_rn(df, df);
            acc[B200YOLO_S_W] += 1.0;
        }
    }

    // ---------------- P4: block reduction -> per-CTA partials ----------------
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
        else if (tid == B200YOLO_S_NCELLS && lead) v = (double)p.cells;
        else if (tid == B200YOLO_S_NIMG && lead) v = 1.0;
        else if (tid == kTLStatusSlot) v = (double)s_misc[2];
        asm volatile("griddepcontrol.wait;" ::: "memory");
        p.partial[((size_t)b * p.S + split) * kTLSums + tid] = v;
    }
}

// ---------------------------------------------------------------------------
// Backward: d loss / d input as the reference's autograd graph defines it (yolo_loss.py:15-32, 84-92,
// 154-159, 219-234).  The custom sigmoid passes gradients through unchanged, exp has its true derivative;
// only entries of `targets` overwritten with constants carry a gradient 2 (o - t) w / sum(w); the CIoU loss
// sum_i (v_i - 1)^2 / n_assign (weights cancel, :224) reaches tx, ty, tw, th of the assigned cells through
// the decoded box (alpha is not detached, :283).  The normalisers sum(w) and n_assign are read from the
// (all-reduced) partial sums in device memory, so the gradient of a data-parallel shard is scaled by the
// batch-global denominators.  Every element of grad_input is written exactly once.
// ---------------------------------------------------------------------------
struct TLBox4d { double x, y, z, w; };

// v = iou - ciou_term and dv/d(pred xyxy) (box1 = gt, box2 = pred), fp64
__device__ __forceinline__ double tl_ciou_grad(const float4 &gt, const TLBox4d &pr, double dv[4]) {
    const double a1 = gt.x, b1 = gt.y, a2 = gt.z, b2 = gt.w;
    const double p1 = pr.x, q1 = pr.y, p2 = pr.z, q2 = pr.w;
    auto sel = [](double lhs, double rhs) { return lhs > rhs ? 1.0 : (lhs == rhs ? 0.5 : 0.0); };  // d max(lhs, rhs) / d lhs
    const double iw_raw = fmin(a2, p2) - fmax(a1, p1), ih_raw = fmin(b2, q2) - fmax(b1, q1);
    const double iw = fmax(iw_raw, 0.0), ih = fmax(ih_raw, 0.0);
    double d_iw[4] = {0, 0, 0, 0}, d_ih[4] = {0, 0, 0, 0};
    if (iw_raw >= 0.0) { d_iw[2] = sel(a2, p2); d_iw[0] = -sel(p1, a1); }   // d min(a2,p2)/dp2 = [p2 < a2]
    if (ih_raw >= 0.0) { d_ih[3] = sel(b2, q2); d_ih[1] = -sel(q1, b1); }
    const double inter = iw * ih;
    const double area1 = (a2 - a1) * (b2 - b1);
    const double w2 = p2 - p1, h2 = q2 - q1;
    const double d_w2[4] = {-1, 0, 1, 0}, d_h2[4] = {0, -1, 0, 1};
    const double uni = area1 + w2 * h2 - inter;
    const double iou = inter / uni;
    const double cw = fmax(a2, p2) - fmin(a1, p1), ch = fmax(b2, q2) - fmin(b1, q1);
    const double d_cw[4] = {-sel(a1, p1), 0, sel(p2, a2), 0};               // d(-min(a1,p1))/dp1 = -[p1 < a1]
    const double d_ch[4] = {0, -sel(b1, q1), 0, sel(q2, b2)};
    const double c = cw * ch;
    const double dx = (a2 + a1) * 0.5 - (p2 + p1) * 0.5, dy = (b1 + b2) * 0.5 - (q1 + q2) * 0.5;
    const double u = dx * dx + dy * dy;
    const double d_u[4] = {-dx, -dy, -dx, -dy};
    const double kk = 4.0 / (3.14159265358979323846 * 3.14159265358979323846);
    // atan(x) - atan(y) = atan((x - y) / (1 + x y)) for x y > -1: one atan and one divide for positive sizes (every
    // box the matching lets through); anything else takes the two-atan form
    const double w1 = a2 - a1, h1 = b2 - b1;
    const double delta = (w1 > 0.0 && h1 > 0.0 && w2 > 0.0 && h2 > 0.0) ? atan((w2 * h1 - w1 * h2) / (h2 * h1 + w2 * w1))
                                                                       : atan(w2 / h2) - atan(w1 / h1);
    const double Aar = kk * delta * delta;
    const double D = 1.0 - iou + Aar + 0.000001;
    // the chain below is what the CTA waits for: reciprocals once, multiplications per coordinate
    const double r_uni2 = 1.0 / (uni * uni), r_c2 = 1.0 / (c * c), r_hw = 1.0 / (h2 * h2 + w2 * w2), r_D2 = 1.0 / (D * D);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const double d_inter = d_iw[q] * ih + iw * d_ih[q];
        const double d_union = d_w2[q] * h2 + w2 * d_h2[q] - d_inter;
        const double d_iou = (d_inter * uni - inter * d_union) * r_uni2;
        const double d_c = d_cw[q] * ch + cw * d_ch[q];
        const double d_dd = (d_u[q] * c - u * d_c) * r_c2;
        const double d_A = 2.0 * kk * delta * (h2 * d_w2[q] - w2 * d_h2[q]) * r_hw;
        const double d_D = -d_iou + d_A;
        const double d_f = (2.0 * Aar * d_A * D - Aar * Aar * d_D) * r_D2;
        dv[q] = (c == 0.0) ? 0.0 : d_iou - d_dd - d_f;
    }
    return (c == 0.0) ? 0.0 : iou - (u / c + Aar * Aar / D);
}

// (min 4 CTAs per SM: the fp64 CIoU gradient, run by a few threads of the first CTA of an image, would otherwise
// raise the register count to 115 and halve the occupancy of the streaming part)
__global__ void __launch_bounds__(kTLThreads, 4) target_loss_backward_kernel(const TLParams p) {
    extern __shared__ __align__(16) unsigned char tl_smem[];
    const TLSmem sm = tl_carve(tl_smem, p.cells, p.gcap, p.A);
    float4 *s_gbox = sm.gbox;
    int *s_gcls = sm.gcls;
    TLAssign *s_list = sm.list;
    int *s_misc = sm.misc;

    const int b = blockIdx.x / p.S, split = blockIdx.x - b * p.S;
    const int tid = threadIdx.x;
    const int HW = p.HW, W = p.W, A = p.A;
    // Programmatic dependent launch: the next kernel of the stream may start; this one reads what the forward call
    // produced (sums, cell states), so it waits before its first global read.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int g0 = p.gt_off[b];
    int nG = p.gt_off[b + 1] - g0;
    if (nG > p.gcap) nG = 0;  // (the forward call reported it)
    float *gbase = p.grad_input + (size_t)b * A * p.attrs * HW;
    const float *hbase = p.head + (size_t)b * A * p.attrs * HW;
    const int cell_lo = split * p.chunk, cell_hi = min(cell_lo + p.chunk, p.cells);
    if (tid < 4) s_misc[tid] = 0;
    {
        // ---- streaming pass first: every cell as if it were not assigned (only the objectness channel can carry a
        // gradient).  Its stores need nothing from the matching below, so they drain while the CTA sits in the
        // matching's barriers; the assigned cells are overwritten after those barriers.
        constexpr int NS = kTLThreads;
        const int st = tid;
        const double inv_w0 = (p.grad_out ? (double)__ldg(p.grad_out) : 1.0) * 2.0 / p.sums[B200YOLO_S_W];
        const bool vec = (HW & 3) == 0 && (((uintptr_t)p.grad_input | (uintptr_t)p.head) & 15) == 0;
        if (vec) {
            // 4 consecutive cells of one anchor plane per thread: 16-byte loads / stores (the slice bounds are multiples of 32)
            for (int cell = cell_lo + 4 * st; cell < cell_hi; cell += 4 * NS) {
                const int a = (int)(((float)cell + 0.5f) * p.invHW);
                const int pos = cell - a * HW;   // multiple of 4; HW is a multiple of 4: the group stays inside the plane
                const size_t off = (size_t)a * p.attrs * HW + pos;
                const uchar4 cs4 = *reinterpret_cast<const uchar4 *>(p.cell_state_in + (size_t)b * p.cells + cell);
                const float4 tc = __ldg(reinterpret_cast<const float4 *>(hbase + off + 4 * (size_t)HW));
                float4 gc;
                gc.x = (cs4.x == 1) ? (float)((double)sigmoid_fast(tc.x) * inv_w0) : 0.f;   // target 0
                gc.y = (cs4.y == 1) ? (float)((double)sigmoid_fast(tc.y) * inv_w0) : 0.f;
                gc.z = (cs4.z == 1) ? (float)((double)sigmoid_fast(tc.z) * inv_w0) : 0.f;
                gc.w = (cs4.w == 1) ? (float)((double)sigmoid_fast(tc.w) * inv_w0) : 0.f;
                float *g = gbase + off;
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int t = 0; t < p.attrs; ++t) __stcs(reinterpret_cast<float4 *>(g + (size_t)t * HW), t == 4 ? gc : z);
            }
        } else {
            for (int cell = cell_lo + st; cell < cell_hi; cell += NS) {
                const int a = (int)(((float)cell + 0.5f) * p.invHW);
                const int pos = cell - a * HW;
                const size_t off = (size_t)a * p.attrs * HW + pos;
                const unsigned char cs1 = p.cell_state_in[(size_t)b * p.cells + cell];
                const float tc = __ldg(hbase + off + 4 * (size_t)HW);
                float gconf = 0.f;
                if (cs1 == 1) gconf = (float)((double)sigmoid_fast(tc) * inv_w0);  // target 0
                float *g = gbase + off;
                for (int t = 0; t < p.attrs; ++t) __stcs(g + (size_t)t * HW, t == 4 ? gconf : 0.f);
            }
        }
    }
    // ---- the image's assignments (the same deterministic list as the forward's)
    __syncthreads();
    tl_match_gt<kTLThreads>(p, g0, nG, false, sm);
    tl_unique_cells<kTLThreads>(s_misc[0], sm);
    __syncthreads();
    const int nE = s_misc[0], nU = s_misc[3];
    const double go = p.grad_out ? (double)__ldg(p.grad_out) : 1.0;
    const double inv_w = go * 2.0 / p.sums[B200YOLO_S_W];                                    // d L_dense / d o = 2 (o - t) w / sum w
    const double n_assign = p.sums[B200YOLO_S_NASSIGN];
    const double inv_n = n_assign > 0.0 ? go * (double)p.iou_weighting * 2.0 / n_assign : 0.0;  // d (w_iou * L_iou) / d v = 2 (v - 1) / n

    // ---- assigned cells of this CTA's slice (after the barrier: they overwrite what the streaming pass stored)
    {
        // thread per assignment: its CIoU gradient w.r.t. (tx, ty, tw, th) of its cell, fp64 on the fp32 decoded box
        for (int e = tid; e < nE; e += kTLThreads) {
            const uint32_t cell = s_list[e].cell;
            if ((int)cell < cell_lo || (int)cell >= cell_hi) continue;
            const int a = (int)(((float)cell + 0.5f) * p.invHW);
            const int pos = (int)cell - a * HW;
            const int j = (int)(((float)pos + 0.5f) * p.invW);
            const int i = pos - j * W;
            const float *q = hbase + (size_t)a * p.attrs * HW + pos;
            const float tx = __ldg(q), ty = __ldg(q + HW), tw = __ldg(q + 2 * (size_t)HW), th = __ldg(q + 3 * (size_t)HW);
            const float aw = p.aw_all[p.mask[a]], ah = p.ah_all[p.mask[a]];
            const float4 pb = tl_decode_box(tx, ty, tw, th, i, j, p.fW, p.fH, aw, ah);  // the forward's fp32 box (:84-92)
            TLBox4d pr;
            pr.x = pb.x; pr.y = pb.y; pr.z = pb.z; pr.w = pb.w;
            double dv[4];
            const double v = tl_ciou_grad(s_gbox[tl_entry_gt(s_list[e].t)], pr, dv);
            const double gl = inv_n * (v - 1.0);
            const double bw = (double)__fmul_rn(expf(tw), aw), bh = (double)__fmul_rn(expf(th), ah);
            double4 c4;
            c4.x = gl * (dv[0] + dv[2]) / (double)p.fW;   // d x1/d sx = d x2/d sx = 1/W; the sigmoid passes through
            c4.y = gl * (dv[1] + dv[3]) / (double)p.fH;
            c4.z = gl * (dv[2] - dv[0]) * 0.5 * bw;       // d x1/d bw = -1/2, d x2/d bw = +1/2, d bw/d tw = bw
            c4.w = gl * (dv[3] - dv[1]) * 0.5 * bh;
            sm.contrib[e] = c4;
        }
        // Objectness / class channels of the distinct assigned cells need nothing from the fp64 terms above, which
        // keep the first ceil(nE / 32) warps busy for a long dependent chain: the OTHER warps take these channels
        // meanwhile (thread per (cell, channel), all logit loads in flight together).
        const int attrs = p.attrs;
        {
            int wbusy = (nE + 31) >> 5;
            if (wbusy > kTLWarps - 1) wbusy = 0;               // (everybody is busy: share the work evenly afterwards)
            const int first = 32 * wbusy, nw = kTLThreads - first;
            const int per = attrs - 4;
            if (tid >= first || wbusy == 0) {
                for (int it = tid - first; it < nU * per; it += nw) {
                    const int u = it / per, t = 4 + (it - u * per);
                    int f = sm.ucell[u];
                    const uint32_t cell = s_list[f].cell;
                    if ((int)cell < cell_lo || (int)cell >= cell_hi) continue;
                    const int a = (int)(((float)cell + 0.5f) * p.invHW);
                    const int pos = (int)cell - a * HW;
                    const size_t off = ((size_t)a * attrs + t) * HW + pos;
                    const float o = sigmoid_f(__ldg(hbase + off));
                    float target = 1.0f;                       // objectness of an assigned cell (:149-150)
                    if (t > 4) {                               // class t-5: 0.95 if assigned to the cell, else 0.05 (:425-434)
                        bool hit = false;
                        while (f >= 0) {
                            const uint32_t tn = s_list[f].t;
                            hit = hit || (s_gcls[tl_entry_gt(tn)] == t - 5);
                            f = tl_entry_next(tn);
                        }
                        target = hit ? 0.95f : 0.05f;
                    }
                    gbase[off] = (float)(((double)o - (double)target) * inv_w);
                }
            }
        }
        __syncthreads();
        // box channels: thread per (distinct cell, channel), sum over the cell's assignments (duplicates add up)
        for (int it = tid; it < nU * 4; it += kTLThreads) {
            const int u = it >> 2, t = it & 3;
            int f = sm.ucell[u];
            const uint32_t cell = s_list[f].cell;
            if ((int)cell < cell_lo || (int)cell >= cell_hi) continue;
            const int a = (int)(((float)cell + 0.5f) * p.invHW);
            const int pos = (int)cell - a * HW;
            const size_t off = ((size_t)a * attrs + t) * HW + pos;
            double acc4 = 0.0;
            while (f >= 0) {
                const double4 c4 = sm.contrib[f];
                acc4 += (t == 0) ? c4.x : (t == 1) ? c4.y : (t == 2) ? c4.z : c4.w;
                f = tl_entry_next(s_list[f].t);
            }
            gbase[off] = (float)acc4;
        }
    }
}

// fixed-order sum over the per-CTA partials -> sums[16] (bitwise reproducible run to run); slot 12 carries the
// status and is max-reduced into status[0]
__global__ void __launch_bounds__(kTLSums * 32) target_loss_reduce_kernel(const double *partial, int rows, double *sums, int *status) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");  // the partial sums of the kernel before
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double v = 0.0;
    if (q == kTLStatusSlot) {
        for (int b = lane; b < rows; b += 32) v = fmax(v, partial[(size_t)b * kTLSums + q]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(kFullMask, v, o));
        if (lane == 0) { status[0] = (int)v; sums[q] = 0.0; }
        return;
    }
    for (int b = lane; b < rows; b += 32) v += partial[(size_t)b * kTLSums + q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    if (lane == 0) sums[q] = v;
}

}  // namespace b200yolo
