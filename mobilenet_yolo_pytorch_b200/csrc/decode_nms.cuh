// decode_nms.cuh -- YOLO-head decode + confidence threshold + per-class NMS (sm_100a).
//
// One CTA per image; everything about one image lives in shared memory, the head
// tensors are read from HBM exactly once and only kept rows are written back.
//
// Shared memory is kept SMALL on purpose (<= 80 KB per CTA for the 352x352 VOC heads):
// the decode phase streams the heads with ordinary coalesced loads whose in-flight
// lines live in L1, and L1 is what the CTAs' shared memory leaves of the SM's 228 KB
// (profiles/micro/load_pattern.cu: the same load pattern runs at 4.3 TB/s with 64 KB of
// L1 and at 2.9 TB/s with none).  So: structure-of-arrays records, and one region `U`
// that is reused by phase (class/arrival index + sort keys -> sorted order + pair masks
// -> output order).
//
//   P1 decode     thread per cell, ALL 5+C attribute planes of the cell loaded at
//                 once (coalesced: consecutive lanes = consecutive cells of a plane).
//                 conf = sigmoid(tc) > thr (yolo_loss.py:189,201); for passing cells:
//                 class max / argmax (:198), box (:186-196,243-247) -> box[], cs[], ta[],
//                 and the per-(class, score bucket) arrival index (one shared atomic).
//   P2 scans      exclusive scan of the (class, bucket) histogram -> class segments
//                 (box.py:20-22), kept-bitmap tiles, the round table.
//   P3 key scatter 64-bit keys (score desc, candidate order asc == the stable sort of
//                 torchvision.ops.nms) into their (class, bucket) segment.
//   P4 rank       rank inside the bucket -> `sord` (sorted position -> cell id).
//   P5 pairs+sweep  dynamic warp tasks.  A task is one 32-row strip of one class: the
//                 lane keeps its row's box in registers and walks the later columns
//                 (broadcast from shared memory), 32 columns per mask word.  The test is
//                 divide-free: iou > thr  <=>  inter > t*(area_r+area_c), t = thr/(1+thr);
//                 d = t*(area_r+area_c) - w*h is one FFMA and its SIGN BIT is the mask
//                 bit (one funnel shift).  |d| <= 1e-5*t*(area sum) (or any degenerate
//                 box in the class) sends the lane to the exact torchvision arithmetic
//                 (inter/(Sa+Sb-inter) > thr in double).  The warp that finishes the last
//                 strip of a class sweeps it at once (no block barrier): per 32-column
//                 tile, OR the mask words of the kept earlier rows, then resolve the
//                 diagonal block by walking only the rows that suppress something.
//   P6 output     class-ascending / score-descending rows (box.py:29-30): scan of the
//                 kept bitmap, then a flat coalesced store.
//
// If the masks of all classes do not fit `U`, P5 runs in rounds (groups of whole
// classes, or column-tile chunks of one huge class) with a block barrier in between.
//
// The stand-alone decode (P1 + ordered compaction + store) and NMS (load rows,
// P2..P6) kernels back YOLOLoss.forward(input) and utils.box.nms separately.
#pragma once
#include "common.cuh"

namespace b200yolo {

constexpr int kMaxAnchors = 8;

enum { MODE_FUSED = 0, MODE_DECODE = 1, MODE_NMS = 2 };

struct HeadDesc {
    const float *ptr;
    int H, W, HW, cells;         // cells = A*H*W
    uint32_t magicHW, magicW;    // ceil(2^32/d) for exact n/d, n < 65536 (0: d == 1)
    float fW, fH, rW, rH;        // grid size and its fp32 reciprocal
    float aw[kMaxAnchors], ah[kMaxAnchors];  // anchors / img_size (yolo_loss.py:214)
};

struct DNParams {
    HeadDesc head[2];
    int nheads;
    int N, A, C, attrs;
    int K;               // candidate slots per image = row stride of out / out_idx
    int B;               // score buckets per class of the counting sort (power of two)
    int flags;           // experiment switches (B200YOLO_FLAGS env): 1 = no L2 prefetch, 8 = prefetch both heads, 4 = old smem sizing
    unsigned long long *dbg;  // optional [N][16] phase time stamps (ns, globaltimer), NULL in production
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
    uint32_t box, cs, ta, cntb, cls, tasks, keptbits, tilepref, passbits, rounds, misc, U, total;
    uint32_t u_bytes;     // bytes in U
    uint32_t mask_off;    // masks / output order start here (after sord)
    uint32_t mask_words;  // 32-bit words available for pair masks
};

__host__ __device__ inline uint32_t align_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

// per-class int arrays, each (C+1) long
enum { CA_CNT = 0, CA_START, CA_KTILE, CA_MASK, CA_TBASE, CA_DONE, CA_FLAG, CA_NUM };
// misc ints
enum { M_NROUNDS = 0, M_CLO, M_CHI, M_T0, M_T1, M_NTASK, M_CTR, M_TOTAL, M_KV, M_WSUM = 16, M_NUM = 16 + 32 };

// score buckets per class: C*B counters, at most 2048 (8 KB)
__host__ __device__ inline int pick_buckets(int C) {
    int B = 64;
    while (B > 1 && C * B > 1536) B >>= 1;
    return B;
}

// U region by phase (Kp = K rounded up to 32):
//   decode .. key scatter   clsidx u32[Kp] | key u64[Kp]
//   rank .. sweep           sord u16[Kp+32] | pair masks u32[mask_words]
//   output                  sord            | outsrc u16[Kp] | outcls u16[Kp]
//   (MODE_DECODE)           clsidx u32[Kp]  | outsrc u16[Kp]
__host__ __device__ inline SmemLayout make_layout(int K, int C, int mode, uint32_t extra_mask_bytes) {
    SmemLayout L;
    const uint32_t Kp = align_up((uint32_t)(K > 0 ? K : 1), 32);
    const uint32_t Cp = align_up((uint32_t)C + 1, 4);
    const uint32_t tiles = Kp / 32 + (uint32_t)C + 1;
    const bool nms = (mode != MODE_DECODE);
    uint32_t o = 0;
    L.box = o; o += 16 * Kp;
    L.cs = o; o += 8 * Kp;
    L.ta = o; o += nms ? 4 * Kp : 0;
    L.cntb = o; o += nms ? 4 * align_up((uint32_t)(C * pick_buckets(C)) + 1, 4) : 0;
    L.cls = o; o += nms ? 4 * Cp * CA_NUM : 0;
    L.tasks = o; o += nms ? 4 * tiles : 0;
    L.keptbits = o; o += nms ? 4 * tiles : 0;
    L.tilepref = o; o += 4 * (tiles + 1);
    L.passbits = o; o += nms ? 0 : 4 * (Kp / 32);
    o = align_up(o, 8);
    L.rounds = o; o += nms ? 8 * (tiles + (uint32_t)C + 2) : 0;
    L.misc = o; o += 4 * M_NUM;
    o = align_up(o, 16);
    L.U = o;
    L.u_bytes = nms ? 12 * Kp + (extra_mask_bytes & ~15u) : 6 * Kp;
    const uint32_t sord_bytes = nms ? align_up(2 * (Kp + 32), 16) : 4 * Kp;
    L.mask_off = L.U + sord_bytes;
    L.mask_words = (L.u_bytes - sord_bytes) / 4;
    L.total = align_up(L.U + L.u_bytes, 16);
    return L;
}

struct Smem {
    float4 *box;        // [Kp] x1 y1 x2 y2 by cell id            (output columns 0-3)
    float2 *cs;         // [Kp] conf, class score                 (columns 4, 5)
    float *ta;          // [Kp] t * area * 2^-13 (NaN: degenerate box, the class takes the exact path)
    uint32_t *clsidx;   // [Kp] (class << 16) | arrival index inside the (class, bucket); ~0u: not a candidate
    unsigned long long *key;
    uint16_t *sord;     // sorted position -> cell id
    uint32_t *mask;     // pair-mask words
    uint16_t *outsrc;   // output row -> cell id
    uint16_t *outcls;   // output row -> class
    uint32_t *passbits, *keptbits, *tilepref, *tasks;
    int *cnt, *start, *ktile, *maskbase, *tbase, *done, *flag;
    int *cntb;          // [C*B+1] per (class, score bucket): arrival counter, then exclusive prefix
    uint2 *rounds;      // x = c_lo | c_hi << 16, y = t0 | t1 << 16
    int *misc;
};

__device__ __forceinline__ Smem carve(unsigned char *base, const SmemLayout &L, int K, int C, int mode) {
    Smem s;
    const uint32_t Kp = align_up((uint32_t)(K > 0 ? K : 1), 32);
    const uint32_t Cp = align_up((uint32_t)C + 1, 4);
    s.box = reinterpret_cast<float4 *>(base + L.box);
    s.cs = reinterpret_cast<float2 *>(base + L.cs);
    s.ta = reinterpret_cast<float *>(base + L.ta);
    s.clsidx = reinterpret_cast<uint32_t *>(base + L.U);
    s.key = reinterpret_cast<unsigned long long *>(base + L.U + 4 * Kp);
    s.sord = reinterpret_cast<uint16_t *>(base + L.U);
    s.mask = reinterpret_cast<uint32_t *>(base + L.mask_off);
    s.outsrc = reinterpret_cast<uint16_t *>(base + L.mask_off);
    s.outcls = reinterpret_cast<uint16_t *>(base + L.mask_off + 2 * Kp);
    s.passbits = reinterpret_cast<uint32_t *>(base + L.passbits);
    s.keptbits = reinterpret_cast<uint32_t *>(base + L.keptbits);
    s.tilepref = reinterpret_cast<uint32_t *>(base + L.tilepref);
    s.tasks = reinterpret_cast<uint32_t *>(base + L.tasks);
    int *ca = reinterpret_cast<int *>(base + L.cls);
    s.cnt = ca + CA_CNT * Cp;
    s.start = ca + CA_START * Cp;
    s.ktile = ca + CA_KTILE * Cp;
    s.maskbase = ca + CA_MASK * Cp;
    s.tbase = ca + CA_TBASE * Cp;
    s.done = ca + CA_DONE * Cp;
    s.flag = ca + CA_FLAG * Cp;
    s.cntb = reinterpret_cast<int *>(base + L.cntb);
    s.rounds = reinterpret_cast<uint2 *>(base + L.rounds);
    s.misc = reinterpret_cast<int *>(base + L.misc);
    (void)mode;
    return s;
}

__device__ __forceinline__ int fastdiv(int n, int d, uint32_t magic) {
    (void)d;
    return magic ? (int)__umulhi((uint32_t)n, magic) : n;  // exact for n < 65536
}

__device__ __forceinline__ int tri(int x) { return (x * (x + 1)) >> 1; }

__device__ __forceinline__ void stamp(const DNParams &p, int b, int k) {
    if (p.dbg && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.dbg[(size_t)b * 16 + k] = t;
    }
}

// t * area * 2^-13 of one box for the divide-free pair test, NaN when the fast test
// must not be trusted for this box (the whole class then runs the exact arithmetic)
__device__ __forceinline__ float make_ta(const float4 &b, const IouThr &t) {
    const float a = box_area(b);
    const float big = fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w)));
    // comparisons are false for NaN coordinates / areas
    const bool ok = t.fast_ok && a >= 1e-20f && a <= 1e20f && big < 4096.0f;  // a NaN coordinate makes the area NaN
    return ok ? __fmul_rn(a, t.ts) : __int_as_float(0x7fc00000);
}

// score bucket of the counting sort: monotone non-increasing in the sort key
// (NaN scores sort first, like torch's descending sort)
__device__ __forceinline__ int score_bucket(float sc, int B) {
    const float top = (float)(B - 1);
    const float f = (sc != sc) ? top : fminf(fmaxf(__fmul_rn(sc, (float)B), 0.0f), top);
    return B - 1 - (int)f;
}

__device__ __forceinline__ float order_key_to_float(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);  // 0xffffffff -> NaN
}

// Ask the L2 to fetch [ptr, ptr+bytes) from HBM (cp.async.bulk.prefetch: the copy engine
// streams it, no registers, no issue slots).  Only the 16-byte-aligned interior is
// requested; the few bytes around it arrive with the ordinary loads.
__device__ __forceinline__ void l2_prefetch_span(const void *ptr, size_t bytes) {
    const uintptr_t lo = ((uintptr_t)ptr + 15) & ~(uintptr_t)15;
    const uintptr_t hi = ((uintptr_t)ptr + bytes) & ~(uintptr_t)15;
    if (hi > lo) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(lo), "r"((uint32_t)(hi - lo)) : "memory");
}

// class score of a cell whose top logits are closer than the sigmoid's evaluation
// error: evaluate like the reference (sigmoid first, then first max, yolo_loss.py:198)
__device__ __noinline__ float class_tie_break(const float *qc, int HW, int C, float lo, float m1, int i1, int *bi_out) {
    float best = -1.0f;
    int bi = 0;
    for (int cc = 0; cc < C; ++cc) {
        const float x = __ldg(qc + (size_t)cc * HW);
        if (!(x < lo)) {
            const float sg = sigmoid_fast(x);
            if (sg > best) { best = sg; bi = cc; }
        }
    }
    if (best < 0.0f) { best = sigmoid_fast(m1); bi = i1; }  // only NaN logits in the window
    *bi_out = bi;
    return best;
}

constexpr int kClsChunk = 24;  // class planes loaded per batch (all in flight before the first use)

// ---------------------------------------------------------------------------
// P1: decode every cell of the image, single pass over the heads.  All 5+C plane
// loads of a cell are issued before the first use (C <= 24; larger C: batches of 24).
// ---------------------------------------------------------------------------
template <int THREADS, int MODE>
__device__ __forceinline__ void phase_decode(const DNParams &p, const Smem &s, int b) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int C = p.C;
    int cid0 = 0;
#pragma unroll 1
    for (int hh = 0; hh < p.nheads; ++hh) {
        const HeadDesc &hd = p.head[hh];
        const int HW = hd.HW;
        const float *hb = hd.ptr + (size_t)b * p.A * p.attrs * HW;  // uniform
#pragma unroll 1
        for (int base = 0; base < hd.cells; base += THREADS) {
            const int local = base + tid;
            bool pass = false;
            if (local < hd.cells) {
                const int cid = cid0 + local;
                const int a = fastdiv(local, HW, hd.magicHW);
                const int pos = local - a * HW;
                // addresses: one IMAD.WIDE each (64-bit base + 32-bit plane stride * constant)
                const char *q = reinterpret_cast<const char *>(hb + (uint32_t)(a * p.attrs * HW + pos));
                const uint32_t st = (uint32_t)HW * 4u;  // plane stride in bytes
#define B200_LD(u) __ldcs(reinterpret_cast<const float *>(q + (uint64_t)st * (uint32_t)(u)))
                const float tx = B200_LD(0);
                const float ty = B200_LD(1);
                const float tw = B200_LD(2);
                const float th = B200_LD(3);
                const float tc = B200_LD(4);
                q += (uint64_t)st * 5u;
                const float *qc = reinterpret_cast<const float *>(q);
                // class max over the raw logits: value, first argmax and whether any other
                // logit lies within `win` of it (then sigmoid rounding could change the result
                // of torch.max(sigmoid(logits)), :198, and the exact tie-break runs)
                float m1 = -INFINITY;
                int i1 = 0;
                bool tie = false;
                float conf = 0.f, e1 = 0.f, best = 0.f, win = 0.f;
#pragma unroll 1
                for (int c0 = 0; c0 < C; c0 += kClsChunk) {
                    // groups of 4 planes behind uniform branches; only the last, partial group clamps
                    // its plane index to the last class (duplicates are masked out of `near` below),
                    // so the loads carry no predicate and no default value
                    float x[kClsChunk];
                    const int nv = min(C - c0, kClsChunk);  // uniform
#pragma unroll
                    for (int g = 0; g < kClsChunk / 4; ++g) {
                        if (4 * g + 4 <= nv) {
#pragma unroll
                            for (int u = 4 * g; u < 4 * g + 4; ++u) x[u] = B200_LD(u);
                        } else if (4 * g < nv) {
#pragma unroll
                            for (int u = 4 * g; u < 4 * g + 4; ++u) x[u] = B200_LD(min(u, nv - 1));
                        }
                    }
                    q += (uint64_t)st * (uint32_t)nv;
                    if (c0 == 0) {
                        conf = sigmoid_fast(tc);   // yolo_loss.py:189,197
                        pass = conf > p.conf_thr;  // :201 (threshold already rounded to fp32)
                    }
                    float cm = x[0];
#pragma unroll
                    for (int g = 0; g < kClsChunk / 4; ++g) {
                        if (4 * g < nv) cm = fmaxf(fmaxf(cm, fmaxf(x[4 * g], x[4 * g + 1])), fmaxf(x[4 * g + 2], x[4 * g + 3]));
                    }
                    if (pass) {
                        // d/dt ln(sigmoid(t)) = 1 - sigmoid(t) >= e*s on (-inf, m], so a logit below
                        // m - 2^-17/(e*s) has a sigmoid smaller by > 2^-17 relative (>10x the
                        // evaluation error) and cannot win or tie
                        const float m_new = fmaxf(m1, cm);
                        e1 = exp_fast(-m_new);
                        best = __fdividef(1.0f, __fadd_rn(1.0f, e1));  // sigmoid(m_new)
                        win = __fdividef(7.6293945e-06f, __fmul_rn(e1, best));
                        const float lo = __fsub_rn(m_new, win);
                        float near = 0.f;  // bit u: x[u] >= lo   (FSET + FFMA: exact for 24 bits)
#pragma unroll
                        for (int g = 0; g < kClsChunk / 4; ++g) {
                            if (4 * g < nv) {
#pragma unroll
                                for (int u = 4 * g; u < 4 * g + 4; ++u)
                                    near = __fmaf_rn((x[u] >= lo) ? 1.0f : 0.0f, (float)(1u << u), near);
                            }
                        }
                        const uint32_t nb = __float2uint_rn(near) & (0xffffffffu >> (32 - nv));
                        // previous chunks: their max m1 must lie below the window too
                        const bool prev_near = (c0 > 0) && !(m1 < lo);
                        if (cm > m1 || c0 == 0) { i1 = c0 + __ffs(nb) - 1; tie = (nb & (nb - 1u)) != 0u || prev_near || nb == 0u; }
                        else tie = tie || nb != 0u;
                        m1 = m_new;
                    } else {
                        m1 = fmaxf(m1, cm);
                    }
                }
                if (pass) {
                    int bi = i1;
                    if (C > 1 && tie) best = class_tie_break(qc, HW, C, __fsub_rn(m1, win), m1, i1, &bi);
                    const int j = fastdiv(pos, hd.W, hd.magicW);
                    const int i = pos - j * hd.W;
                    const float sx = sigmoid_fast(tx), sy = sigmoid_fast(ty);    // :187
                    const float ew = exp_fast(tw), eh = exp_fast(th);            // :188
                    const float cx = __fmul_rn(__fadd_rn(sx, (float)i), hd.rW);  // :194 (x * 1/W)
                    const float cy = __fmul_rn(__fadd_rn(sy, (float)j), hd.rH);
                    const float bw = __fmul_rn(ew, hd.aw[a]);                    // :195
                    const float bh = __fmul_rn(eh, hd.ah[a]);
                    float4 bx;
                    bx.x = __fsub_rn(cx, __fmul_rn(bw, 0.5f));                   // :244
                    bx.y = __fsub_rn(cy, __fmul_rn(bh, 0.5f));                   // :245
                    bx.z = __fadd_rn(bw, bx.x);                                  // :246
                    bx.w = __fadd_rn(bh, bx.y);                                  // :247
                    s.box[cid] = bx;
                    s.cs[cid] = make_float2(conf, best);
                    uint32_t idx = 0;
                    if (MODE == MODE_FUSED) {
                        s.ta[cid] = make_ta(bx, p.iou);
                        idx = (uint32_t)atomicAdd(&s.cntb[bi * p.B + score_bucket(__fmul_rn(best, conf), p.B)], 1);
                    }
                    s.clsidx[cid] = ((uint32_t)bi << 16) | idx;
                } else if (MODE == MODE_FUSED) {
                    s.clsidx[cid] = 0xffffffffu;
                }
            }
#undef B200_LD
            if (MODE == MODE_DECODE) {  // single head: candidate ids are 32-aligned per warp
                const unsigned bal = __ballot_sync(kFullMask, pass);
                if (lane == 0 && local < hd.cells) s.passbits[local >> 5] = bal;
            }
            if (MODE == MODE_FUSED) stamp(p, b, 8 + min(6, hh * 4 + base / THREADS));
        }
        cid0 += hd.cells;
    }
}

// ---------------------------------------------------------------------------
// MODE_NMS: load already-decoded rows (two heads, box.py:17) into the records
// ---------------------------------------------------------------------------
template <int THREADS>
__device__ __forceinline__ void phase_load_rows(const DNParams &p, const Smem &s, int b) {
    const int tid = threadIdx.x;
    const int K0 = min(p.cand_count[0][b], p.cand_stride[0]);
    const int K1 = p.cand[1] ? min(p.cand_count[1][b], p.cand_stride[1]) : 0;
    const int Kb = K0 + K1;
    const float *r0 = p.cand[0] + (size_t)b * p.cand_stride[0] * 7;
    const float *r1 = p.cand[1] ? p.cand[1] + (size_t)b * p.cand_stride[1] * 7 : nullptr;
    for (int base = 0; base < p.K; base += THREADS) {
        const int row = base + tid;
        bool ok = false;
        if (row < Kb) {
            const float *src = (row < K0) ? r0 + (size_t)row * 7 : r1 + (size_t)(row - K0) * 7;
            const float4 bx = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), __ldg(src + 3));
            const float conf = __ldg(src + 4), score = __ldg(src + 5);
            const float v = __ldg(src + 6);
            const int c = (int)v;  // rows whose class column is not an integer in [0,C) match no `== i` (box.py:21)
            ok = (v == (float)c) && c >= 0 && c < p.C;
            if (ok) {
                s.box[row] = bx;
                s.cs[row] = make_float2(conf, score);
                s.ta[row] = make_ta(bx, p.iou);
                const uint32_t idx = (uint32_t)atomicAdd(&s.cntb[c * p.B + score_bucket(__fmul_rn(score, conf), p.B)], 1);
                s.clsidx[row] = ((uint32_t)c << 16) | idx;
            }
        }
        if (row < p.K && !ok) s.clsidx[row] = 0xffffffffu;
    }
}

// ---------------------------------------------------------------------------
// P2 (warp 0): class segments, kept-bitmap tiles, and the round table
// ---------------------------------------------------------------------------
__device__ __forceinline__ void warp_class_scan(const DNParams &p, const Smem &s, int mask_cap_words) {
    const int lane = threadIdx.x & 31;
    const int C = p.C;
    int carryT = 0, words = 0;
    for (int c0 = 0; c0 < C; c0 += 32) {
        const int c = c0 + lane;
        const int st = (c < C) ? s.cntb[c * p.B] : 0;
        const int n = (c < C) ? s.cntb[(c + 1) * p.B] - st : 0;
        const int T = (n + 31) >> 5;
        const int incT = warp_inclusive_scan(T, lane);
        if (c < C) {
            s.cnt[c] = n;
            s.start[c] = st;
            s.ktile[c] = carryT + incT - T;
        }
        carryT += __shfl_sync(kFullMask, incT, 31);
        words += 32 * tri(T);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) words += __shfl_xor_sync(kFullMask, words, o);
    if (lane == 0) {
        s.cnt[C] = 0;
        s.start[C] = s.cntb[C * p.B];
        s.ktile[C] = carryT;
        const int cap = mask_cap_words;
        if (words <= cap) {
            s.rounds[0] = make_uint2((uint32_t)C << 16, 0xffffu << 16);
            s.misc[M_NROUNDS] = 1;
        } else {
            // rare: groups of whole classes, or column-tile chunks of one huge class
            int r = 0, c = 0;
            while (c < C) {
                const int T = (s.cnt[c] + 31) >> 5;
                if (32 * tri(T) > cap) {
                    int t0 = 0;
                    while (t0 < T) {
                        int t1 = t0, acc = 0;
                        while (t1 < T && acc + 32 * (t1 + 1) <= cap) { acc += 32 * (t1 + 1); ++t1; }
                        if (t1 == t0) ++t1;  // cannot happen: cap >= Kp >= 32*T
                        s.rounds[r++] = make_uint2((uint32_t)c | ((uint32_t)(c + 1) << 16), (uint32_t)t0 | ((uint32_t)t1 << 16));
                        t0 = t1;
                    }
                    ++c;
                } else {
                    const int c_lo = c;
                    int acc = 0;
                    while (c < C) {
                        const int w = 32 * tri((s.cnt[c] + 31) >> 5);
                        if (acc + w > cap) break;
                        acc += w;
                        ++c;
                    }
                    s.rounds[r++] = make_uint2((uint32_t)c_lo | ((uint32_t)c << 16), 0xffffu << 16);
                }
            }
            s.misc[M_NROUNDS] = r;
        }
    }
    __syncwarp();
}

// per round (warp 0): mask offsets, strip-task table and completion counters of the round's classes
__device__ __forceinline__ void warp_round_prefix(const Smem &s, int r) {
    const int lane = threadIdx.x & 31;
    const uint2 rd = s.rounds[r];
    const int c_lo = rd.x & 0xffff, c_hi = rd.x >> 16, t0 = rd.y & 0xffff, t1 = rd.y >> 16;
    int carryS = 0, carryM = 0;
    for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
        const int c = c0 + lane;
        int strips = 0, words = 0;
        if (c < c_hi) {
            const int T = (s.cnt[c] + 31) >> 5;
            const int te = min(T, t1);
            if (te > t0) {
                strips = te;  // every row tile below te has blocks in column tiles [max(rt,t0), te)
                words = 32 * (tri(te) - tri(t0));
            }
        }
        const int incS = warp_inclusive_scan(strips, lane);
        const int incM = warp_inclusive_scan(words, lane);
        if (c < c_hi) {
            const int tb = carryS + incS - strips;
            s.tbase[c] = tb;
            s.maskbase[c] = carryM + incM - words;
            s.done[c] = 0;
            for (int rt = 0; rt < strips; ++rt) s.tasks[tb + rt] = ((uint32_t)c << 16) | (uint32_t)rt;
        }
        carryS += __shfl_sync(kFullMask, incS, 31);
        carryM += __shfl_sync(kFullMask, incM, 31);
    }
    if (lane == 0) {
        s.misc[M_CLO] = c_lo; s.misc[M_CHI] = c_hi; s.misc[M_T0] = t0; s.misc[M_T1] = t1;
        s.misc[M_NTASK] = carryS;
        s.misc[M_CTR] = 0;
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------
// P2a (all threads): exclusive scan of the (class, score bucket) counters in place
// -> sorted-position base of every bucket; total -> misc[M_KV]
// ---------------------------------------------------------------------------
template <int THREADS>
__device__ __forceinline__ void block_scan_buckets(const DNParams &p, const Smem &s) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NB = p.C * p.B;
    const int per = (NB + THREADS - 1) / THREADS;
    const int lo = min(tid * per, NB), hi = min(lo + per, NB);
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += s.cntb[i];
    const int inc = warp_inclusive_scan(sum, lane);
    if (lane == 31) s.misc[M_WSUM + warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const int v = (lane < THREADS / 32) ? s.misc[M_WSUM + lane] : 0;
        const int w = warp_inclusive_scan(v, lane);
        s.misc[M_WSUM + lane] = w - v;
        if (lane == 31) { s.misc[M_KV] = w; s.cntb[NB] = w; }
    }
    __syncthreads();
    int base = s.misc[M_WSUM + warp] + inc - sum;
    for (int i = lo; i < hi; ++i) {
        const int t = s.cntb[i];
        s.cntb[i] = base;
        base += t;
    }
}

// ---------------------------------------------------------------------------
// P3/P4: class-major counting sort on score buckets, then rank inside the bucket:
// the stable descending score sort of torchvision.ops.nms, per class
// key = score order key (32) | class (16) | 0xffff - cell id (16)
// ---------------------------------------------------------------------------
template <int THREADS>
__device__ __forceinline__ void phase_scatter_keys(const DNParams &p, const Smem &s) {
    for (int cid = threadIdx.x; cid < p.K; cid += THREADS) {
        const uint32_t ci = s.clsidx[cid];
        if (ci == 0xffffffffu) continue;
        const float2 cs = s.cs[cid];
        const float sc = __fmul_rn(cs.y, cs.x);  // box.py:27 scores = col5*col4
        const uint32_t c = ci >> 16;
        const unsigned long long key =
            ((unsigned long long)float_order_key(sc) << 32) | (unsigned long long)((c << 16) | (0xffffu - (uint32_t)cid));
        s.key[s.cntb[(int)c * p.B + score_bucket(sc, p.B)] + (int)(ci & 0xffffu)] = key;
        const float ta = s.ta[cid];
        if (ta != ta) s.flag[c] = 1;  // degenerate box: the class runs the exact pair arithmetic
    }
}

template <int THREADS>
__device__ __forceinline__ void phase_rank_sort(const DNParams &p, const Smem &s, int Kv) {
    for (int t = threadIdx.x; t < Kv; t += THREADS) {
        const unsigned long long key = s.key[t];
        const uint32_t lo32 = (uint32_t)key;
        const uint32_t cid = 0xffffu - (lo32 & 0xffffu);
        const int cq = (int)(lo32 >> 16) * p.B + score_bucket(order_key_to_float((uint32_t)(key >> 32)), p.B);
        const int st = s.cntb[cq], en = s.cntb[cq + 1];
        int rank = 0;
        for (int u = st; u < en; ++u) rank += (s.key[u] > key) ? 1 : 0;
        s.sord[st + rank] = (uint16_t)cid;
    }
}

// ---------------------------------------------------------------------------
// P5: pair masks (row-major words: bit k of word (rt, ct)[lane] = row 32rt+lane suppresses
// column 32ct+k) and the sweep
// ---------------------------------------------------------------------------
constexpr float kPairEps = 1e-5f;

// exact torchvision decisions of one row against the columns of one tile (slow path)
__device__ __noinline__ uint32_t block_exact(const float4 *box, const uint16_t *ordc, int ncol, const float4 R, double thr) {
    const float ra = box_area(R);
    uint32_t word = 0u;
    for (int k = 0; k < ncol; ++k) {
        const float4 Cb = box[ordc[k]];
        if (nms_suppress_exact(R, ra, Cb, box_area(Cb), thr)) word |= 1u << k;
    }
    return word;
}

__device__ __forceinline__ uint32_t block_fast(const Smem &s, const uint16_t *ordc, int ncol, const float4 R, float rta,
                                               double thr) {
    uint32_t bits = 0u;
    float m = INFINITY;
#pragma unroll 4
    for (int k = 0; k < ncol; ++k) {
        const int ccid = ordc[k];
        const float4 Cb = s.box[ccid];
        const float cta = s.ta[ccid];
        const float w = __fsub_rn(fminf(R.z, Cb.z), fmaxf(R.x, Cb.x));
        const float h = __fsub_rn(fminf(R.w, Cb.w), fmaxf(R.y, Cb.y));
        const float ws = __saturatef(__fmul_rn(w, 1.220703125e-4f));  // max(w, 0) * 2^-13 (exact; |coords| < 4096)
        const float sum = __fadd_rn(rta, cta);                         // t * (area_r + area_c) * 2^-13
        const float d = __fmaf_rn(-ws, h, sum);                        // < 0  <=>  inter > t * (area sum)
        m = fminf(m, __fmaf_rn(sum, -kPairEps, fabsf(d)));             // <= 0: too close to call
        bits = __funnelshift_l(__float_as_uint(d), bits, 1);           // sign bit -> mask bit
    }
    uint32_t word = __brev(bits) >> (32 - ncol);  // column k was shifted in k-th: bit ncol-1-k -> bit k
    if (m <= 0.0f) word = block_exact(s.box, ordc, ncol, R, thr);
    return word;
}

// sweep of one class over the column tiles [t0, te) (earlier tiles' kept words are final)
__device__ __forceinline__ void sweep_class(const Smem &s, int c, int t0, int te, int tri0) {
    const int lane = threadIdx.x & 31;
    const int n = s.cnt[c];
    uint32_t *kept_w = s.keptbits + s.ktile[c];
    const uint32_t *mb = s.mask + s.maskbase[c];
    for (int ct = t0; ct < te; ++ct) {
        const uint32_t *blk = mb + (tri(ct) - tri0) * 32 + lane;
        uint32_t sup = 0u;
        for (int rt = 0; rt < ct; ++rt) {
            const uint32_t wv = blk[rt * 32];
            if ((kept_w[rt] >> lane) & 1u) sup |= wv;
        }
        sup = __reduce_or_sync(kFullMask, sup);
        const int ncol = min(32, n - 32 * ct);
        const uint32_t valid = (ncol >= 32) ? 0xffffffffu : ((1u << ncol) - 1u);
        const uint32_t alive = valid & ~sup;
        // diagonal block: only rows that are alive and suppress an alive column need a turn
        const uint32_t D = ((alive >> lane) & 1u) ? (blk[ct * 32] & alive) : 0u;
        uint32_t nz = __ballot_sync(kFullMask, D != 0u);
        uint32_t rem = alive;
        while (nz) {
            const int i = __ffs(nz) - 1;
            nz &= nz - 1u;
            const uint32_t Di = __shfl_sync(kFullMask, D, i);
            if ((rem >> i) & 1u) rem &= ~Di;
        }
        if (lane == 0) kept_w[ct] = rem;
        __syncwarp();
    }
}

__device__ __forceinline__ void phase_pairs_sweep(const DNParams &p, const Smem &s) {
    const int lane = threadIdx.x & 31;
    const int t0 = s.misc[M_T0], t1 = s.misc[M_T1];
    const int ntask = s.misc[M_NTASK];
    const int tri0 = tri(t0);
    for (;;) {
        int q = 0;
        if (lane == 0) q = atomicAdd(&s.misc[M_CTR], 1);
        q = __shfl_sync(kFullMask, q, 0);
        if (q >= ntask) break;
        const uint32_t tk = s.tasks[q];
        const int c = (int)(tk >> 16), rt = (int)(tk & 0xffffu);
        const int n = s.cnt[c];
        const int te = min((n + 31) >> 5, t1);
        const uint16_t *ord = s.sord + s.start[c];
        const int row = 32 * rt + lane;
        const int rcid = ord[min(row, n - 1)];
        const float4 R = s.box[rcid];
        const float rta = s.ta[rcid];
        const bool slow = s.flag[c] != 0;
        uint32_t *mb = s.mask + s.maskbase[c];
        for (int ct = max(rt, t0); ct < te; ++ct) {
            const int ncol = min(32, n - 32 * ct);
            uint32_t word = slow ? block_exact(s.box, ord + 32 * ct, ncol, R, p.iou.thr)
                                 : block_fast(s, ord + 32 * ct, ncol, R, rta, p.iou.thr);
            if (ct == rt) word &= ~((2u << lane) - 1u);  // only LATER columns (2u << 31 == 0: none)
            if (row >= n) word = 0u;
            mb[(tri(ct) - tri0 + rt) * 32 + lane] = word;
        }
        __syncwarp();
        int old = 0;
        if (lane == 0) {
            __threadfence_block();
            old = atomicAdd(&s.done[c], 1);
        }
        old = __shfl_sync(kFullMask, old, 0);
        if (old == te - 1) {  // last strip of the class in this round: its masks are complete
            __threadfence_block();
            sweep_class(s, c, t0, te, tri0);
        }
    }
}

// ---------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------
template <int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS, (THREADS == 512) ? 2 : 1) decode_nms_kernel(const DNParams p, const SmemLayout L) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem s = carve(smem_raw, L, p.K, p.C, MODE);
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = p.C, K = p.K;
    constexpr int kWarps = THREADS / 32;

    stamp(p, b, 0);
    if (MODE != MODE_NMS) {
        // Start the HBM -> L2 stream of the FIRST head now, so that the first decode round (which can
        // only issue after the launch ramp) finds its lines on the way.  Prefetching the whole image
        // is slower: the demand loads then queue behind 46 MB of prefetches (profiles/r01/NOTES.md).
        if (tid == 0 && !(p.flags & 1)) {
            const size_t img0 = (size_t)p.A * p.attrs * p.head[0].HW;
            l2_prefetch_span(p.head[0].ptr + (size_t)b * img0, img0 * sizeof(float));
            if (p.nheads > 1 && (p.flags & 8)) {
                const size_t img1 = (size_t)p.A * p.attrs * p.head[1].HW;
                l2_prefetch_span(p.head[1].ptr + (size_t)b * img1, img1 * sizeof(float));
            }
        }
    }
    if (MODE != MODE_DECODE) {
        for (int i = tid; i <= C * p.B; i += THREADS) s.cntb[i] = 0;
        for (int i = tid; i <= C; i += THREADS) s.flag[i] = 0;
        __syncthreads();
    }
    stamp(p, b, 15);

    if (MODE == MODE_NMS) phase_load_rows<THREADS>(p, s, b);
    else phase_decode<THREADS, MODE>(p, s, b);
    __syncthreads();
    stamp(p, b, 1);

    const int nwords = (K + 31) >> 5;

    if (MODE == MODE_DECODE) {
        // YOLOLoss.get_pred_boxes output: rows in candidate order (:203)
        if (warp == 0) {
            int carry = 0;
            for (int w0 = 0; w0 < nwords; w0 += 32) {
                const int w = w0 + lane;
                const int v = (w < nwords) ? __popc(s.passbits[w]) : 0;
                const int inc = warp_inclusive_scan(v, lane);
                if (w < nwords) s.tilepref[w] = carry + inc - v;
                carry += __shfl_sync(kFullMask, inc, 31);
            }
            if (lane == 0) s.misc[M_TOTAL] = carry;
        }
        __syncthreads();
        for (int cid = tid; cid < K; cid += THREADS) {
            const uint32_t bits = s.passbits[cid >> 5];
            if ((bits >> (cid & 31)) & 1u) s.outsrc[s.tilepref[cid >> 5] + __popc(bits & ((1u << (cid & 31)) - 1u))] = (uint16_t)cid;
        }
        __syncthreads();
        const int T = s.misc[M_TOTAL];
        float *o = p.out + (size_t)b * K * 7;
        const float *boxf = reinterpret_cast<const float *>(s.box);
        const float *csf = reinterpret_cast<const float *>(s.cs);
        for (int f = tid; f < 7 * T; f += THREADS) {
            const int row = f / 7, col = f - 7 * row;
            const int cid = s.outsrc[row];
            o[f] = (col < 4) ? boxf[4 * cid + col] : (col < 6) ? csf[2 * cid + col - 4] : (float)(s.clsidx[cid] >> 16);  // cls_idx.float() :199
        }
        if (p.out_idx)
            for (int r = tid; r < T; r += THREADS) p.out_idx[(size_t)b * K + r] = (int)s.outsrc[r];
        if (tid == 0) p.out_count[b] = T;
        return;
    }

    // P2, P3 (warp 0's class bookkeeping overlaps the key scatter of the others)
    block_scan_buckets<THREADS>(p, s);
    __syncthreads();
    if (warp == 0) {
        warp_class_scan(p, s, (int)L.mask_words);
        warp_round_prefix(s, 0);
    }
    phase_scatter_keys<THREADS>(p, s);
    __syncthreads();
    stamp(p, b, 2);
    const int Kv = s.misc[M_KV];
    const int nrounds = s.misc[M_NROUNDS];
    // P4
    phase_rank_sort<THREADS>(p, s, Kv);
    __syncthreads();
    stamp(p, b, 3);
    // P5
    for (int r = 0;;) {
        phase_pairs_sweep(p, s);
        __syncthreads();
        if (++r >= nrounds) break;
        if (warp == 0) warp_round_prefix(s, r);
        __syncthreads();
    }
    stamp(p, b, 4);
    stamp(p, b, 5);
    // P6: output row -> cell id (the mask buffer is dead now; outsrc aliases it).  Every
    // warp sums the kept counts of the tiles before its class itself (no serial scan phase).
    for (int c = warp; c <= C; c += kWarps) {
        const int kt = s.ktile[c];
        int before = 0;
        for (int g = lane; g < kt; g += 32) before += __popc(s.keptbits[g]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(kFullMask, before, o);
        if (c == C) {
            if (lane == 0) s.misc[M_TOTAL] = before;
            break;
        }
        const int n = s.cnt[c], st = s.start[c];
        for (int ct = 0; 32 * ct < n; ++ct) {
            const uint32_t word = s.keptbits[kt + ct];
            if ((word >> lane) & 1u) {
                const int r = before + __popc(word & lanemask_lt());
                s.outsrc[r] = s.sord[st + 32 * ct + lane];
                s.outcls[r] = (uint16_t)c;
            }
            before += __popc(word);
        }
    }
    __syncthreads();
    stamp(p, b, 6);
    // flat coalesced store
    {
        const int T = s.misc[M_TOTAL];
        float *o = p.out + (size_t)b * K * 7;
        if (MODE == MODE_NMS) {
            // gather the caller's own row (pred_this_cls[index], box.py:29) bit-for-bit
            const int K0 = min(p.cand_count[0][b], p.cand_stride[0]);
            const float *r0 = p.cand[0] + (size_t)b * p.cand_stride[0] * 7;
            const float *r1 = p.cand[1] ? p.cand[1] + (size_t)b * p.cand_stride[1] * 7 : nullptr;
            for (int f = tid; f < 7 * T; f += THREADS) {
                const int row = f / 7, col = f - 7 * row;
                const int k = s.outsrc[row];
                o[f] = (k < K0) ? __ldg(r0 + (size_t)k * 7 + col) : __ldg(r1 + (size_t)(k - K0) * 7 + col);
            }
        } else {
            const float *boxf = reinterpret_cast<const float *>(s.box);
            const float *csf = reinterpret_cast<const float *>(s.cs);
            for (int f = tid; f < 7 * T; f += THREADS) {
                const int row = f / 7, col = f - 7 * row;
                const int cid = s.outsrc[row];
                float v;
                if (col < 4) v = boxf[4 * cid + col];
                else if (col < 6) v = csf[2 * cid + col - 4];
                else v = (float)s.outcls[row];
                o[f] = v;
            }
        }
        if (p.out_idx)
            for (int r = tid; r < T; r += THREADS) p.out_idx[(size_t)b * K + r] = (int)s.outsrc[r];
        if (tid == 0) p.out_count[b] = T;
    }
    stamp(p, b, 7);
}

}  // namespace b200yolo
