// decode_nms.cuh -- YOLO-head decode + confidence threshold + per-class NMS.
//
// One CTA per image.  Everything about one image (<= a few thousand candidate
// boxes) is staged in shared memory; the head tensors are read from HBM once,
// coalesced (threads walk the contiguous H*W plane of one attribute), and only
// the kept rows are written back.  Phases (all inside one launch):
//
//   A1  objectness plane -> sigmoid -> `conf > thr` -> ORDER-PRESERVING compaction
//       (warp ballot + popc prefix, block scan of warp totals): candidate k keeps
//       the reference's (head, a, j, i) row-major order (yolo_loss.py:201-203).
//   A2  one thread per surviving candidate: box decode (yolo_loss.py:186-196,
//       243-247), class max/argmax (:198), per-class histogram.
//   B   exclusive scan of the class histogram -> class segments (box.py:20-22).
//   C   scatter 64-bit sort keys (score desc, candidate order asc = stable sort
//       of torchvision.ops.nms) into class segments; rank-sort inside each
//       segment; permute boxes into sorted order.
//   D   per class (one warp each): 32x32 bitmask-tiled greedy NMS -- the diagonal
//       tile is resolved with a ballot/bitmask sweep, its kept rows are then
//       applied to all later column tiles (suppressed rows are never visited).
//   E   class-ascending / score-descending output order (box.py:29-30): scan of
//       per-class kept counts, then a flat coalesced store of the kept rows.
//
// The same phases are reused by the stand-alone decode (A1,A2 + store) and NMS
// (load rows, B..E) kernels, which back YOLOLoss.forward(input) and
// utils.box.nms separately.
#pragma once
#include "common.cuh"

namespace b200yolo {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxAnchors = 8;

enum { MODE_FUSED = 0, MODE_DECODE = 1, MODE_NMS = 2 };

struct HeadDesc {
    const float *ptr;
    int H, W, HW, cells;  // cells = A*H*W
    float invHW, invW, fW, fH;
    float aw[kMaxAnchors], ah[kMaxAnchors];  // anchors / img_size (yolo_loss.py:214)
};

struct DNParams {
    HeadDesc head[2];
    int nheads;
    int N, A, C, attrs;
    int Kmax;  // row stride of out / out_idx (= total cells, or stride0+stride1 in NMS mode)
    float conf_thr;
    IouThr iou;
    float *out;
    int *out_count;
    int *out_idx;
    // MODE_NMS inputs
    const float *cand[2];
    const int *cand_count[2];
    int cand_stride[2];
};

struct SmemLayout {
    uint32_t box, sbox, key, conf, cscore, sarea, cell, order, cls, alive;
    uint32_t hist, start, kcount, kstart, wcount, misc, total;
};

__host__ __device__ inline uint32_t align_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

__host__ __device__ inline SmemLayout make_layout(int Kmax, int C, int mode) {
    SmemLayout L;
    uint32_t K = align_up((uint32_t)(Kmax > 0 ? Kmax : 1), 16);
    uint32_t Cp = align_up((uint32_t)C + 2, 4);
    uint32_t o = 0;
    const bool nms = (mode != MODE_DECODE);
    L.box = o; o += 16 * K;
    L.sbox = o; o += nms ? 16 * K : 0;
    L.key = o; o += nms ? 8 * K : 0;   // (outsrc, uint16, aliases key after the rank phase)
    L.conf = o; o += 4 * K;
    L.cscore = o; o += 4 * K;
    L.sarea = o; o += nms ? 4 * K : 0;
    L.cell = o; o += 4 * K;
    L.order = o; o += nms ? 2 * K : 0;
    L.cls = o; o += 2 * K;
    L.alive = o; o += nms ? K : 0;
    o = align_up(o, 16);
    L.hist = o; o += 4 * Cp;
    L.start = o; o += 4 * Cp;
    L.kcount = o; o += 4 * Cp;
    L.kstart = o; o += 4 * Cp;
    L.wcount = o; o += 4 * 2 * kWarps;
    L.misc = o; o += 64;
    L.total = align_up(o, 16);
    return L;
}

struct Smem {
    float4 *box, *sbox;
    unsigned long long *key;
    uint16_t *outsrc;
    float *conf, *cscore, *sarea;
    uint32_t *cell;
    uint16_t *order, *cls;
    uint8_t *alive;
    int *hist, *start, *kcount, *kstart, *wcount, *misc;
};

__device__ __forceinline__ Smem carve(unsigned char *base, const SmemLayout &L) {
    Smem s;
    s.box = reinterpret_cast<float4 *>(base + L.box);
    s.sbox = reinterpret_cast<float4 *>(base + L.sbox);
    s.key = reinterpret_cast<unsigned long long *>(base + L.key);
    s.outsrc = reinterpret_cast<uint16_t *>(base + L.key);
    s.conf = reinterpret_cast<float *>(base + L.conf);
    s.cscore = reinterpret_cast<float *>(base + L.cscore);
    s.sarea = reinterpret_cast<float *>(base + L.sarea);
    s.cell = reinterpret_cast<uint32_t *>(base + L.cell);
    s.order = reinterpret_cast<uint16_t *>(base + L.order);
    s.cls = reinterpret_cast<uint16_t *>(base + L.cls);
    s.alive = reinterpret_cast<uint8_t *>(base + L.alive);
    s.hist = reinterpret_cast<int *>(base + L.hist);
    s.start = reinterpret_cast<int *>(base + L.start);
    s.kcount = reinterpret_cast<int *>(base + L.kcount);
    s.kstart = reinterpret_cast<int *>(base + L.kstart);
    s.wcount = reinterpret_cast<int *>(base + L.wcount);
    s.misc = reinterpret_cast<int *>(base + L.misc);
    return s;
}

constexpr uint16_t kNoClass = 0xffffu;

// ---------------------------------------------------------------------------
// A1: objectness threshold + order-preserving compaction.  Returns K (uniform).
// ---------------------------------------------------------------------------
__device__ __forceinline__ int phase_threshold_compact(const DNParams &p, const Smem &s, int b) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cells0 = p.head[0].cells;
    const int total = cells0 + (p.nheads > 1 ? p.head[1].cells : 0);
    int base_k = 0;
    int parity = 0;
    for (int base = 0; base < total; base += kThreads, parity ^= 1) {
        const int cid = base + tid;
        bool pass = false;
        float conf = 0.0f;
        if (cid < total) {
            const bool h1 = cid >= cells0;
            const HeadDesc &hd = h1 ? p.head[1] : p.head[0];
            const int local = h1 ? cid - cells0 : cid;
            const int a = (int)(((float)local + 0.5f) * hd.invHW);
            const int pos = local - a * hd.HW;
            const float *q = hd.ptr + ((size_t)(b * p.A + a) * p.attrs + 4) * hd.HW + pos;
            conf = sigmoid_f(ld_stream_f(q));   // yolo_loss.py:189,197
            pass = conf > p.conf_thr;           // :201 (threshold already rounded to fp32)
        }
        const unsigned bal = __ballot_sync(kFullMask, pass);
        if (lane == 0) s.wcount[parity * kWarps + warp] = __popc(bal);
        __syncthreads();
        int before = 0, all = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const int c = s.wcount[parity * kWarps + w];
            before += (w < warp) ? c : 0;
            all += c;
        }
        if (pass) {
            const int k = base_k + before + __popc(bal & lanemask_lt());
            s.cell[k] = (uint32_t)cid;
            s.conf[k] = conf;
        }
        base_k += all;
        // wcount is double-buffered by chunk parity, so one barrier per chunk suffices
    }
    return base_k;
}

// ---------------------------------------------------------------------------
// A2: decode the surviving candidates.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void phase_decode(const DNParams &p, const Smem &s, int b, int K, bool want_hist) {
    const int cells0 = p.head[0].cells;
    const int C = p.C;
    for (int k = threadIdx.x; k < K; k += kThreads) {
        const int cid = (int)s.cell[k];
        const bool h1 = cid >= cells0;
        const HeadDesc &hd = h1 ? p.head[1] : p.head[0];
        const int local = h1 ? cid - cells0 : cid;
        const int a = (int)(((float)local + 0.5f) * hd.invHW);
        const int pos = local - a * hd.HW;
        const int j = (int)(((float)pos + 0.5f) * hd.invW);
        const int i = pos - j * hd.W;
        const int HW = hd.HW;
        const float *q = hd.ptr + ((size_t)(b * p.A + a) * p.attrs) * HW + pos;
        const float tx = ld_stream_f(q), ty = ld_stream_f(q + HW);
        const float tw = ld_stream_f(q + 2 * HW), th = ld_stream_f(q + 3 * HW);
        const float *qc = q + 5 * (size_t)HW;

        // class max over the raw logits; sigmoid is only evaluated where it can
        // change the (value, first-argmax) of torch.max(sigmoid(logits)) (:198)
        float m1 = ld_stream_f(qc), m2 = -INFINITY;
        int i1 = 0;
        int c = 1;
        for (; c + 4 <= C; c += 4) {
            float x0 = ld_stream_f(qc + (size_t)(c + 0) * HW), x1 = ld_stream_f(qc + (size_t)(c + 1) * HW);
            float x2 = ld_stream_f(qc + (size_t)(c + 2) * HW), x3 = ld_stream_f(qc + (size_t)(c + 3) * HW);
            if (x0 > m1) { m2 = m1; m1 = x0; i1 = c; } else m2 = fmaxf(m2, x0);
            if (x1 > m1) { m2 = m1; m1 = x1; i1 = c + 1; } else m2 = fmaxf(m2, x1);
            if (x2 > m1) { m2 = m1; m1 = x2; i1 = c + 2; } else m2 = fmaxf(m2, x2);
            if (x3 > m1) { m2 = m1; m1 = x3; i1 = c + 3; } else m2 = fmaxf(m2, x3);
        }
        for (; c < C; ++c) {
            float x0 = ld_stream_f(qc + (size_t)c * HW);
            if (x0 > m1) { m2 = m1; m1 = x0; i1 = c; } else m2 = fmaxf(m2, x0);
        }
        const float e1 = expf(-m1);
        float best = __fdiv_rn(1.0f, __fadd_rn(1.0f, e1));  // sigmoid(m1)
        int bi = i1;
        // d/dt ln(sigmoid(t)) = 1 - sigmoid(t) >= e1*best on (-inf, m1], so any logit
        // below m1 - 2^-19/(e1*best) has a sigmoid smaller by > 2^-19 relative (>10x
        // the evaluation error) and cannot win or tie.  Everything inside the window
        // is evaluated exactly like the reference (sigmoid first, then first max).
        const float win = __fdiv_rn(1.9073486e-06f, __fmul_rn(e1, best));
        if (C > 1 && !(m2 < __fsub_rn(m1, win))) {
            const float lo = __fsub_rn(m1, win);
            best = -1.0f;
            bi = 0;
            for (int cc = 0; cc < C; ++cc) {
                const float x = __ldg(qc + (size_t)cc * HW);
                if (!(x < lo)) {
                    const float sg = sigmoid_f(x);
                    if (sg > best) { best = sg; bi = cc; }
                }
            }
            if (best < 0.0f) { best = sigmoid_f(m1); bi = i1; }  // only NaN logits in the window
        }

        const float sx = sigmoid_f(tx), sy = sigmoid_f(ty);          // :187
        const float ew = expf(tw), eh = expf(th);                    // :188
        const float cx = __fdiv_rn(__fadd_rn(sx, (float)i), hd.fW);  // :194
        const float cy = __fdiv_rn(__fadd_rn(sy, (float)j), hd.fH);
        const float bw = __fmul_rn(ew, hd.aw[a]);                    // :195
        const float bh = __fmul_rn(eh, hd.ah[a]);
        float4 bx;
        bx.x = __fsub_rn(cx, __fmul_rn(bw, 0.5f));                   // :244
        bx.y = __fsub_rn(cy, __fmul_rn(bh, 0.5f));                   // :245
        bx.z = __fadd_rn(bw, bx.x);                                  // :246
        bx.w = __fadd_rn(bh, bx.y);                                  // :247
        s.box[k] = bx;
        s.cscore[k] = best;
        s.cls[k] = (uint16_t)bi;
        if (want_hist) atomicAdd(&s.hist[bi], 1);
    }
}

// ---------------------------------------------------------------------------
// MODE_NMS: load already-decoded rows (two heads, box.py:17) into the same staging
// ---------------------------------------------------------------------------
__device__ __forceinline__ int phase_load_rows(const DNParams &p, const Smem &s, int b) {
    const int K0 = p.cand_count[0][b];
    const int K1 = p.cand[1] ? p.cand_count[1][b] : 0;
    const int K = K0 + K1;
    const float *r0 = p.cand[0] + (size_t)b * p.cand_stride[0] * 7;
    const float *r1 = p.cand[1] ? p.cand[1] + (size_t)b * p.cand_stride[1] * 7 : nullptr;
    float *boxf = reinterpret_cast<float *>(s.box);
    for (int f = threadIdx.x; f < 7 * K; f += kThreads) {
        const int row = f / 7, col = f - 7 * row;
        const float v = (row < K0) ? __ldg(r0 + f) : __ldg(r1 + (f - 7 * K0));
        if (col < 4) boxf[4 * row + col] = v;
        else if (col == 4) s.conf[row] = v;
        else if (col == 5) s.cscore[row] = v;
        else {
            const int c = (int)v;  // rows whose class column is not an integer in [0,C) match no `== i` (box.py:21)
            const bool ok = (v == (float)c) && c >= 0 && c < p.C;
            s.cls[row] = ok ? (uint16_t)c : kNoClass;
            s.cell[row] = (uint32_t)row;
            if (ok) atomicAdd(&s.hist[c], 1);
        }
    }
    return K;
}

// ---------------------------------------------------------------------------
// B: exclusive scan of hist[0..C) -> start[0..C]; hist is zeroed (reused as the
// scatter cursor).  Executed by warp 0.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void warp_scan_classes(int *cnt, int *start, int C, bool zero_cnt) {
    const int lane = threadIdx.x & 31;
    int carry = 0;
    for (int c0 = 0; c0 < C; c0 += 32) {
        const int c = c0 + lane;
        const int v = (c < C) ? cnt[c] : 0;
        const int inc = warp_inclusive_scan(v, lane);
        if (c < C) {
            start[c] = carry + inc - v;
            if (zero_cnt) cnt[c] = 0;
        }
        carry += __shfl_sync(kFullMask, inc, 31);
    }
    if (lane == 0) start[C] = carry;
}

// ---------------------------------------------------------------------------
// C: class-segmented stable sort by score (descending)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void phase_scatter_keys(const DNParams &p, const Smem &s, int K) {
    for (int k = threadIdx.x; k < K; k += kThreads) {
        const uint16_t c = s.cls[k];
        if (c == kNoClass) continue;
        const float sc = __fmul_rn(s.cscore[k], s.conf[k]);  // box.py:27 scores = col5*col4
        const unsigned long long key =
            ((unsigned long long)float_order_key(sc) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)k);
        const int slot = s.start[c] + atomicAdd(&s.hist[c], 1);
        s.key[slot] = key;
    }
}

__device__ __forceinline__ void phase_rank_sort(const Smem &s, int Kv) {
    for (int t = threadIdx.x; t < Kv; t += kThreads) {
        const unsigned long long key = s.key[t];
        const uint32_t k = 0xffffffffu - (uint32_t)(key & 0xffffffffu);
        const int c = s.cls[k];
        const int st = s.start[c], en = s.start[c + 1];
        int rank = 0;
        for (int u = st; u < en; ++u) rank += (s.key[u] > key) ? 1 : 0;
        const int pos = st + rank;
        const float4 bx = s.box[k];
        s.order[pos] = (uint16_t)k;
        s.sbox[pos] = bx;
        s.sarea[pos] = box_area(bx);
        s.alive[pos] = 1;
    }
}

// ---------------------------------------------------------------------------
// D: greedy NMS of one class segment by one warp, 32x32 bitmask tiles.
// Returns the number of kept boxes (warp-uniform).
// ---------------------------------------------------------------------------
__device__ __forceinline__ int warp_nms_class(const Smem &s, int st, int n, const IouThr &thr) {
    const int lane = threadIdx.x & 31;
    const int ntiles = (n + 31) >> 5;
    int kept_total = 0;
    for (int rt = 0; rt < ntiles; ++rt) {
        const int r = rt * 32 + lane;
        const bool valid = r < n;
        float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
        float marea = 0.f;
        bool alive_me = false;
        if (valid) {
            alive_me = s.alive[st + r] != 0;
            me = s.sbox[st + r];
            marea = s.sarea[st + r];
        }
        const unsigned alive_in = __ballot_sync(kFullMask, alive_me);
        if (alive_in == 0u) continue;
        // diagonal tile: bit i of colword = "row i (earlier, still alive) overlaps me"
        unsigned colword = 0u;
        const float4 *rowb = s.sbox + st + rt * 32;
        const float *rowa = s.sarea + st + rt * 32;
        // (the last alive row of the tile has no later column inside it)
        for (unsigned rem = alive_in; rem & (rem - 1u);) {
            const int i = __ffs(rem) - 1;
            rem &= rem - 1u;
            const float4 rb = rowb[i];
            const float ra = rowa[i];
            if (alive_me && lane > i && nms_suppress(rb, ra, me, marea, thr)) colword |= 1u << i;
        }
        // sweep: rows without any overlap bit are kept outright; the others are
        // resolved in ascending order against the kept mask built so far
        const unsigned nz = __ballot_sync(kFullMask, colword != 0u);
        unsigned kept = alive_in & ~nz;
        for (unsigned rem = nz; rem;) {
            const int i = __ffs(rem) - 1;
            rem &= rem - 1u;
            const unsigned cw = __shfl_sync(kFullMask, colword, i);
            if ((cw & kept) == 0u) kept |= 1u << i;
        }
        if (valid && alive_me && !((kept >> lane) & 1u)) s.alive[st + r] = 0;
        kept_total += __popc(kept);
        // apply this tile's kept rows to every later column tile
        for (int ct = rt + 1; ct < ntiles; ++ct) {
            const int j = ct * 32 + lane;
            bool a = (j < n) && (s.alive[st + j] != 0);
            if (!__any_sync(kFullMask, a)) continue;
            float4 cb = make_float4(0.f, 0.f, 0.f, 0.f);
            float ca = 0.f;
            if (a) { cb = s.sbox[st + j]; ca = s.sarea[st + j]; }
            const bool was = a;
            for (unsigned rem = kept; rem;) {
                const int i = __ffs(rem) - 1;
                rem &= rem - 1u;
                const float4 rb = rowb[i];
                const float ra = rowa[i];
                if (a && nms_suppress(rb, ra, cb, ca, thr)) a = false;
            }
            if (was && !a) s.alive[st + j] = 0;
        }
        __syncwarp();
    }
    return kept_total;
}

// E1: per class, map kept sorted positions to output rows
__device__ __forceinline__ void warp_emit_class(const Smem &s, int st, int n, int out_base) {
    const int lane = threadIdx.x & 31;
    int run = out_base;
    for (int r0 = 0; r0 < n; r0 += 32) {
        const int r = r0 + lane;
        const bool a = (r < n) && (s.alive[st + r] != 0);
        const unsigned bal = __ballot_sync(kFullMask, a);
        if (a) s.outsrc[run + __popc(bal & lanemask_lt())] = s.order[st + r];
        run += __popc(bal);
    }
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 2) decode_nms_kernel(const DNParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SmemLayout L = make_layout(p.Kmax, p.C, MODE);
    const Smem s = carve(smem_raw, L);
    const int b = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int C = p.C;

    if (MODE != MODE_DECODE) {
        for (int c = tid; c <= C; c += kThreads) { s.hist[c] = 0; s.kcount[c] = 0; }
        __syncthreads();
    }

    int K;
    if (MODE == MODE_NMS) {
        K = phase_load_rows(p, s, b);
    } else {
        K = phase_threshold_compact(p, s, b);
        __syncthreads();
        phase_decode(p, s, b, K, MODE == MODE_FUSED);
    }
    __syncthreads();

    if (MODE == MODE_DECODE) {
        // YOLOLoss.get_pred_boxes output: rows in candidate order (:203)
        float *o = p.out + (size_t)b * p.Kmax * 7;
        const float *boxf = reinterpret_cast<const float *>(s.box);
        for (int f = tid; f < 7 * K; f += kThreads) {
            const int row = f / 7, col = f - 7 * row;
            float v;
            if (col < 4) v = boxf[4 * row + col];
            else if (col == 4) v = s.conf[row];
            else if (col == 5) v = s.cscore[row];
            else v = (float)s.cls[row];   // cls_idx.float() :199
            o[f] = v;
        }
        if (p.out_idx) {
            const int cells0 = p.head[0].cells;
            (void)cells0;
            for (int k = tid; k < K; k += kThreads) p.out_idx[(size_t)b * p.Kmax + k] = (int)s.cell[k];
        }
        if (tid == 0) p.out_count[b] = K;
        return;
    }

    // B
    if (warp == 0) warp_scan_classes(s.hist, s.start, C, true);
    __syncthreads();
    const int Kv = s.start[C];
    // C
    phase_scatter_keys(p, s, K);
    __syncthreads();
    phase_rank_sort(s, Kv);
    __syncthreads();
    // D
    for (int c = warp; c < C; c += kWarps) {
        const int st = s.start[c], n = s.start[c + 1] - st;
        if (n == 0) continue;
        const int kept = warp_nms_class(s, st, n, p.iou);
        if ((tid & 31) == 0) s.kcount[c] = kept;
    }
    __syncthreads();
    // E
    if (warp == 0) warp_scan_classes(s.kcount, s.kstart, C, false);
    __syncthreads();
    const int T = s.kstart[C];
    for (int c = warp; c < C; c += kWarps) {
        const int st = s.start[c], n = s.start[c + 1] - st;
        if (n == 0) continue;
        warp_emit_class(s, st, n, s.kstart[c]);
    }
    __syncthreads();
    {
        float *o = p.out + (size_t)b * p.Kmax * 7;
        const float *boxf = reinterpret_cast<const float *>(s.box);
        for (int f = tid; f < 7 * T; f += kThreads) {
            const int row = f / 7, col = f - 7 * row;
            const int k = s.outsrc[row];
            float v;
            if (MODE == MODE_NMS) {
                // gather the caller's own row (pred_this_cls[index], box.py:29) bit-for-bit
                const int K0 = p.cand_count[0][b];
                v = (k < K0) ? __ldg(p.cand[0] + ((size_t)b * p.cand_stride[0] + k) * 7 + col)
                             : __ldg(p.cand[1] + ((size_t)b * p.cand_stride[1] + (k - K0)) * 7 + col);
            } else {
                if (col < 4) v = boxf[4 * k + col];
                else if (col == 4) v = s.conf[k];
                else if (col == 5) v = s.cscore[k];
                else v = (float)s.cls[k];
            }
            o[f] = v;
        }
        if (p.out_idx)
            for (int r = tid; r < T; r += kThreads) p.out_idx[(size_t)b * p.Kmax + r] = (int)s.cell[s.outsrc[r]];
        if (tid == 0) p.out_count[b] = T;
    }
}

}  // namespace b200yolo
