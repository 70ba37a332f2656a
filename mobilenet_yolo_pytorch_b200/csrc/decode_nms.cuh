// decode_nms.cuh -- YOLO-head decode + confidence threshold + per-class NMS (sm_100a).
//
// One CTA per image; everything about one image lives in shared memory, the head
// tensors are read from HBM exactly once and only kept rows are written back.
//
//   P1 decode     thread per cell, ALL 5+C attribute planes of the cell loaded at
//                 once (coalesced: consecutive lanes = consecutive cells of a plane;
//                 every load is independent, so ~13 x THREADS x 4 B are in flight
//                 per CTA).  conf = sigmoid(tc) > thr (yolo_loss.py:189,201); for
//                 passing cells: class max / argmax (:198), box (:186-196,243-247),
//                 one 32-byte record in shared memory, pass bit (ballot) and the
//                 per-class arrival index (one shared atomic).
//   P2 class scan exclusive scan of the class histogram -> class segments
//                 (box.py:20-22), tiles, pair-slot and bitmask offsets.
//   P3 key scatter 64-bit keys (score desc, candidate order asc == the stable sort
//                 of torchvision.ops.nms) into class segments.
//   P4 rank sort  rank inside the class segment -> `sord` (sorted position -> record).
//   P5 pair masks ALL threads: for every pair (row i < column j) of a class, bit i
//                 of column j's mask = "i suppresses j".  Column j and column
//                 n-1-j share a lane, so every lane does n-1 tests (balanced) while
//                 row records are broadcast from shared memory.  The test is a
//                 divide-free two-sided filter (inter > t*(area_i+area_j) with a
//                 1e-5 guard band, t = thr/(1+thr)); only pairs inside the band run
//                 the exact torchvision arithmetic (inter/(Sa+Sb-inter) > thr in
//                 double).  Result bits are accumulated with FFMA (fma pipe) so the
//                 alu pipe (FMNMX/FSET) is not the only one working.
//   P6 sweep      one warp per class walks 32-column tiles: columns suppressed by a
//                 kept row of an earlier tile drop out with one AND per tile, the
//                 diagonal tile is resolved with ballot/shfl in ascending order.
//   P7-9 output   class-ascending / score-descending rows (box.py:29-30): scan of the
//                 kept bitmap, then a flat coalesced store.
//
// If the masks of all classes do not fit the shared-memory budget, P5/P6 run in
// rounds (groups of whole classes, or column-tile chunks of one huge class).
//
// The stand-alone decode (P1 + ordered compaction + store) and NMS (load rows,
// P2..P9) kernels back YOLOLoss.forward(input) and utils.box.nms separately.
#pragma once
#include "common.cuh"

namespace b200yolo {

constexpr int kMaxAnchors = 8;
constexpr uint32_t kNoClass = 0xffffu;

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
    int mask_cap_words;  // capacity of the pair-mask buffer (32-bit words)
    int B;               // score buckets per class of the counting sort (power of two)
    int flags;           // experiment switches (B200YOLO_FLAGS env): 1 = no L2 prefetch
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

// 32-byte candidate record, indexed by cell id (candidate id)
struct __align__(16) Rec {
    float4 box;   // x1 y1 x2 y2            (output columns 0-3)
    float conf;   //                         (column 4)
    float score;  // class score             (column 5)
    float ta_hi;  // t*(1+1e-5)*area   (+inf: always take the exact path)
    float ta_lo;  // t*(1-1e-5)*area   (-inf: always take the exact path)
};

struct SmemLayout {
    uint32_t rec, clsidx, sord, key, passbits, keptbits, tilepref, cls, cntb, rounds, misc, total;
    uint32_t mask_words;  // words available at `key` (keys are dead once ranks are known)
};

__host__ __device__ inline uint32_t align_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

// per-class int arrays, each (C+1) long
enum { CA_CNT = 0, CA_START, CA_KTILE, CA_SLOT, CA_MASK, CA_NUM };
// misc ints
enum { M_NROUNDS = 0, M_CLO, M_CHI, M_T0, M_T1, M_SLOTS, M_CTR, M_TOTAL, M_KV, M_WSUM = 16, M_NUM = 16 + 32 };

// score buckets per class: C*B counters, at most 2048 (8 KB)
__host__ __device__ inline int pick_buckets(int C) {
    int B = 64;
    while (B > 1 && C * B > 2048) B >>= 1;
    return B;
}

__host__ __device__ inline SmemLayout make_layout(int K, int C, int mode, uint32_t extra_mask_bytes) {
    SmemLayout L;
    const uint32_t Kp = align_up((uint32_t)(K > 0 ? K : 1), 32);
    const uint32_t Cp = align_up((uint32_t)C + 1, 4);
    const uint32_t tiles = Kp / 32 + (uint32_t)C + 1;
    const bool nms = (mode != MODE_DECODE);
    uint32_t o = 0;
    L.rec = o; o += 32 * Kp;
    L.clsidx = o; o += 4 * Kp;
    L.sord = o; o += nms ? 4 * (2 * Kp + 64) : 2 * Kp;  // decode mode: u16 output order
    L.key = o; o += nms ? 8 * Kp + (extra_mask_bytes & ~15u) : 0;
    L.mask_words = nms ? (8 * Kp + (extra_mask_bytes & ~15u)) / 4 : 0;
    L.passbits = o; o += 4 * (Kp / 32);
    L.keptbits = o; o += nms ? 4 * tiles : 0;
    L.tilepref = o; o += 4 * (tiles + 1);
    o = align_up(o, 16);
    L.cls = o; o += nms ? 4 * Cp * CA_NUM : 0;
    L.cntb = o; o += nms ? 4 * align_up((uint32_t)(C * pick_buckets(C)) + 1, 4) : 0;
    L.rounds = o; o += nms ? 8 * (tiles + (uint32_t)C + 2) : 0;
    L.misc = o; o += 4 * M_NUM;
    L.total = align_up(o, 16);
    return L;
}

struct Smem {
    Rec *rec;
    uint32_t *clsidx;   // (class << 16) | arrival index inside the class
    uint32_t *sord;     // sorted position -> shared-space address of the record
    uint16_t *outsrc;   // output row -> cell id (aliases sord in decode mode, keys otherwise)
    unsigned long long *key;
    uint32_t *mask;     // aliases key
    uint32_t *passbits, *keptbits, *tilepref;
    int *cnt, *start, *ktile, *slot, *maskbase;
    int *cntb;          // [C*B+1] per (class, score bucket): arrival counter, then exclusive prefix
    uint2 *rounds;      // x = c_lo | c_hi << 16, y = t0 | t1 << 16
    int *misc;
    uint32_t rec_saddr;
};

__device__ __forceinline__ Smem carve(unsigned char *base, const SmemLayout &L, int C, int mode) {
    Smem s;
    const uint32_t Cp = align_up((uint32_t)C + 1, 4);
    s.rec = reinterpret_cast<Rec *>(base + L.rec);
    s.clsidx = reinterpret_cast<uint32_t *>(base + L.clsidx);
    s.sord = reinterpret_cast<uint32_t *>(base + L.sord);
    s.key = reinterpret_cast<unsigned long long *>(base + L.key);
    s.mask = reinterpret_cast<uint32_t *>(base + L.key);
    s.outsrc = (mode == MODE_DECODE) ? reinterpret_cast<uint16_t *>(base + L.sord)
                                     : reinterpret_cast<uint16_t *>(base + L.key);
    s.passbits = reinterpret_cast<uint32_t *>(base + L.passbits);
    s.keptbits = reinterpret_cast<uint32_t *>(base + L.keptbits);
    s.tilepref = reinterpret_cast<uint32_t *>(base + L.tilepref);
    int *ca = reinterpret_cast<int *>(base + L.cls);
    s.cnt = ca + CA_CNT * Cp;
    s.start = ca + CA_START * Cp;
    s.ktile = ca + CA_KTILE * Cp;
    s.slot = ca + CA_SLOT * Cp;
    s.maskbase = ca + CA_MASK * Cp;
    s.cntb = reinterpret_cast<int *>(base + L.cntb);
    s.rounds = reinterpret_cast<uint2 *>(base + L.rounds);
    s.misc = reinterpret_cast<int *>(base + L.misc);
    s.rec_saddr = (uint32_t)__cvta_generic_to_shared(s.rec);
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

__device__ __forceinline__ void lds_rec(uint32_t saddr, float4 &box, float2 &ta) {
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(box.x), "=f"(box.y), "=f"(box.z), "=f"(box.w)
                 : "r"(saddr));
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+24];" : "=f"(ta.x), "=f"(ta.y) : "r"(saddr));
}

__device__ __forceinline__ float4 lds_box(uint32_t saddr) {
    float4 b;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(saddr));
    return b;
}

// guard-banded area terms of one box for the divide-free pair filter
__device__ __forceinline__ void make_ta(const float4 &b, const IouThr &t, float &hi, float &lo) {
    const float a = box_area(b);
    const bool ok = t.fast_ok && a >= 1e-30f && a <= 1e30f;
    hi = ok ? __fmul_rn(a, t.t_hi) : INFINITY;
    lo = ok ? __fmul_rn(a, t.t_lo) : -INFINITY;
}

// score bucket of the counting sort: monotone non-increasing in the sort key
// (NaN scores sort first, like torch's descending sort)
__device__ __forceinline__ int score_bucket(float sc, int B) {
    const float top = (float)(B - 1);
    const float f = (sc != sc) ? top : fminf(fmaxf(__fmul_rn(sc, (float)B), 0.0f), top);
    return B - 1 - (int)f;
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
    if ((p.flags >> 8) && (tid >> 5) >= THREADS / 64) __nanosleep((unsigned)(p.flags >> 8) * 100u);  // experiment
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
                const float *q = hb + (uint32_t)(a * p.attrs * HW + pos);
                const float tx = __ldcs(q); q += HW;
                const float ty = __ldcs(q); q += HW;
                const float tw = __ldcs(q); q += HW;
                const float th = __ldcs(q); q += HW;
                const float tc = __ldcs(q); q += HW;
                const float *qc = q;
                // class max over the raw logits: value, first argmax and whether any other
                // logit lies within `win` of it (then sigmoid rounding could change the result
                // of torch.max(sigmoid(logits)), :198, and the exact tie-break runs)
                float m1 = -INFINITY;
                int i1 = 0;
                bool tie = false;
                float conf = 0.f, e1 = 0.f, best = 0.f, win = 0.f;
#pragma unroll 1
                for (int c0 = 0; c0 < C; c0 += kClsChunk) {
                    // groups of 4 planes behind uniform branches; inside a group the plane index is
                    // clamped to the last class (duplicates are masked out of `near` below), so the
                    // loads carry no predicate and no default value
                    float x[kClsChunk];
                    const int nv = min(C - c0, kClsChunk);  // uniform
#pragma unroll
                    for (int g = 0; g < kClsChunk / 4; ++g) {
                        if (4 * g < nv) {
#pragma unroll
                            for (int u = 4 * g; u < 4 * g + 4; ++u) x[u] = __ldcs(q + (size_t)min(u, nv - 1) * HW);
                        }
                    }
                    q += (size_t)nv * HW;
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
                    float hi, lo;
                    make_ta(bx, p.iou, hi, lo);
                    float4 *dst = reinterpret_cast<float4 *>(&s.rec[cid]);
                    dst[0] = bx;
                    dst[1] = make_float4(conf, best, hi, lo);
                    uint32_t idx = 0;
                    if (MODE == MODE_FUSED)
                        idx = (uint32_t)atomicAdd(&s.cntb[bi * p.B + score_bucket(__fmul_rn(best, conf), p.B)], 1);
                    s.clsidx[cid] = ((uint32_t)bi << 16) | idx;
                } else if (MODE == MODE_FUSED) {
                    s.clsidx[cid] = 0xffffffffu;
                }
            }
            if (MODE == MODE_DECODE) {  // single head: candidate ids are 32-aligned per warp
                const unsigned bal = __ballot_sync(kFullMask, pass);
                if (lane == 0 && local < hd.cells) s.passbits[local >> 5] = bal;
            }
            if (MODE == MODE_FUSED) stamp(p, b, 8 + min(7, hh * 4 + base / THREADS));
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
            Rec r;
            r.box = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), __ldg(src + 3));
            r.conf = __ldg(src + 4);
            r.score = __ldg(src + 5);
            const float v = __ldg(src + 6);
            const int c = (int)v;  // rows whose class column is not an integer in [0,C) match no `== i` (box.py:21)
            ok = (v == (float)c) && c >= 0 && c < p.C;
            if (ok) {
                make_ta(r.box, p.iou, r.ta_hi, r.ta_lo);
                float4 *dst = reinterpret_cast<float4 *>(&s.rec[row]);
                dst[0] = r.box;
                dst[1] = make_float4(r.conf, r.score, r.ta_hi, r.ta_lo);
                const uint32_t idx =
                    (uint32_t)atomicAdd(&s.cntb[c * p.B + score_bucket(__fmul_rn(r.score, r.conf), p.B)], 1);
                s.clsidx[row] = ((uint32_t)c << 16) | idx;
            }
        }
        if (row < p.K && !ok) s.clsidx[row] = 0xffffffffu;
    }
}

// ---------------------------------------------------------------------------
// P2 (warp 0): class segments, kept-bitmap tiles, and the round table
// ---------------------------------------------------------------------------
__device__ __forceinline__ void warp_class_scan(const DNParams &p, const Smem &s) {
    const int lane = threadIdx.x & 31;
    const int C = p.C;
    int carryS = 0, carryT = 0, words = 0;
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
    carryS = s.cntb[C * p.B];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) words += __shfl_xor_sync(kFullMask, words, o);
    if (lane == 0) {
        s.cnt[C] = 0;
        s.start[C] = carryS;
        s.ktile[C] = carryT;
        const int cap = p.mask_cap_words;
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
                        if (t1 == t0) ++t1;  // cannot happen: host guarantees cap >= Kp >= 32*T
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

// per round (warp 0): pair-slot and mask offsets of the round's classes
__device__ __forceinline__ void warp_round_prefix(const Smem &s, int r) {
    const int lane = threadIdx.x & 31;
    const uint2 rd = s.rounds[r];
    const int c_lo = rd.x & 0xffff, c_hi = rd.x >> 16, t0 = rd.y & 0xffff, t1 = rd.y >> 16;
    int carryS = 0, carryM = 0;
    for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
        const int c = c0 + lane;
        int slots = 0, words = 0;
        if (c < c_hi) {
            const int n = s.cnt[c];
            const int T = (n + 31) >> 5;
            const int te = min(T, t1);
            const int j0 = 32 * t0, j1 = min(n, 32 * te);
            if (j1 > j0) {
                slots = (j1 - j0 + 1) >> 1;
                words = 32 * (tri(te) - tri(t0));
            }
        }
        const int incS = warp_inclusive_scan(slots, lane);
        const int incM = warp_inclusive_scan(words, lane);
        if (c < c_hi) {
            s.slot[c] = carryS + incS - slots;
            s.maskbase[c] = carryM + incM - words;
        }
        carryS += __shfl_sync(kFullMask, incS, 31);
        carryM += __shfl_sync(kFullMask, incM, 31);
    }
    if (lane == 0) {
        s.slot[c_hi] = carryS;
        s.misc[M_CLO] = c_lo; s.misc[M_CHI] = c_hi; s.misc[M_T0] = t0; s.misc[M_T1] = t1;
        s.misc[M_SLOTS] = carryS;
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
// ---------------------------------------------------------------------------
template <int THREADS>
__device__ __forceinline__ void phase_scatter_keys(const DNParams &p, const Smem &s) {
    for (int cid = threadIdx.x; cid < p.K; cid += THREADS) {
        const uint32_t ci = s.clsidx[cid];
        if (ci == 0xffffffffu) continue;
        const Rec &r = s.rec[cid];
        const float sc = __fmul_rn(r.score, r.conf);  // box.py:27 scores = col5*col4
        const unsigned long long key =
            ((unsigned long long)float_order_key(sc) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)cid);
        s.key[s.cntb[(int)(ci >> 16) * p.B + score_bucket(sc, p.B)] + (int)(ci & 0xffffu)] = key;
    }
}

template <int THREADS>
__device__ __forceinline__ void phase_rank_sort(const DNParams &p, const Smem &s, int Kv, int sord_len) {
    for (int t = threadIdx.x; t < sord_len; t += THREADS) {
        if (t >= Kv) {
            s.sord[t] = s.rec_saddr;  // padding: any valid record
            continue;
        }
        const unsigned long long key = s.key[t];
        const uint32_t cid = 0xffffffffu - (uint32_t)(key & 0xffffffffu);
        const Rec &r = s.rec[cid];
        const int cq = (int)(s.clsidx[cid] >> 16) * p.B + score_bucket(__fmul_rn(r.score, r.conf), p.B);
        const int st = s.cntb[cq], en = s.cntb[cq + 1];
        int rank = 0;
        for (int u = st; u < en; ++u) rank += (s.key[u] > key) ? 1 : 0;
        s.sord[st + rank] = s.rec_saddr + 32u * cid;
    }
}

// ---------------------------------------------------------------------------
// P5: pair masks
// ---------------------------------------------------------------------------
// exact torchvision decision for one pair (row record address, column box)
__device__ __noinline__ bool pair_exact(uint32_t row_saddr, const float4 &cb, const IouThr &thr) {
    const float4 rb = lds_box(row_saddr);
    return nms_suppress_exact(rb, box_area(rb), cb, box_area(cb), thr);
}

// one 16-row group for up to two columns; returns 16 result bits per column
template <bool DO_A>
__device__ __forceinline__ void pair_group16(const uint32_t *ord, const float4 &ca, float ca_hi, float ca_lo,
                                             const float4 &cb, float cb_hi, float cb_lo, const IouThr &thr,
                                             uint32_t &bitsA, uint32_t &bitsB) {
    float aS = 0.f, aM = 0.f, bS = 0.f, bM = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        float4 R;
        float2 T;
        lds_rec(ord[k], R, T);
        const float wk = (float)(1u << k);
        {
            const float w = fmaxf(__fsub_rn(fminf(R.z, cb.z), fmaxf(R.x, cb.x)), 0.0f);
            const float h = __fsub_rn(fminf(R.w, cb.w), fmaxf(R.y, cb.y));
            const float inter = __fmul_rn(w, h);
            const float pS = (inter > __fadd_rn(T.x, cb_hi)) ? 1.0f : 0.0f;
            const float pM = (inter > __fadd_rn(T.y, cb_lo)) ? 1.0f : 0.0f;
            bS = __fmaf_rn(pS, wk, bS);
            bM = __fmaf_rn(pM, wk, bM);
        }
        if (DO_A) {
            const float w = fmaxf(__fsub_rn(fminf(R.z, ca.z), fmaxf(R.x, ca.x)), 0.0f);
            const float h = __fsub_rn(fminf(R.w, ca.w), fmaxf(R.y, ca.y));
            const float inter = __fmul_rn(w, h);
            const float pS = (inter > __fadd_rn(T.x, ca_hi)) ? 1.0f : 0.0f;
            const float pM = (inter > __fadd_rn(T.y, ca_lo)) ? 1.0f : 0.0f;
            aS = __fmaf_rn(pS, wk, aS);
            aM = __fmaf_rn(pM, wk, aM);
        }
    }
    uint32_t uS = __float2uint_rn(bS), uM = __float2uint_rn(bM);
    if (uS != uM) {  // pairs inside the guard band (or flagged boxes): exact arithmetic decides
        for (uint32_t amb = uS ^ uM; amb;) {
            const int k = __ffs(amb) - 1;
            amb &= amb - 1u;
            if (pair_exact(ord[k], cb, thr)) uS |= 1u << k; else uS &= ~(1u << k);
        }
    }
    bitsB = uS;
    if (DO_A) {
        uS = __float2uint_rn(aS);
        uM = __float2uint_rn(aM);
        if (uS != uM) {
            for (uint32_t amb = uS ^ uM; amb;) {
                const int k = __ffs(amb) - 1;
                amb &= amb - 1u;
                if (pair_exact(ord[k], ca, thr)) uS |= 1u << k; else uS &= ~(1u << k);
            }
        }
        bitsA = uS;
    }
}

__device__ __forceinline__ void phase_pairs(const DNParams &p, const Smem &s) {
    const int lane = threadIdx.x & 31;
    const int c_lo = s.misc[M_CLO], c_hi = s.misc[M_CHI], t0 = s.misc[M_T0], t1 = s.misc[M_T1];
    const int total = s.misc[M_SLOTS];
    const int tri0 = tri(t0);
    for (;;) {
        int chunk = 0;
        if (lane == 0) chunk = atomicAdd(&s.misc[M_CTR], 1);
        chunk = __shfl_sync(kFullMask, chunk, 0);
        if (chunk * 32 >= total) break;
        const int slot = chunk * 32 + lane;
        const bool act = slot < total;
        int lo = c_lo, hi = c_hi;  // largest c in [c_lo, c_hi) with slot[c] <= slot
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (s.slot[mid] <= slot) lo = mid; else hi = mid;
        }
        const int c = lo;
        const int n = s.cnt[c];
        const int rowbase = s.start[c];
        const int j0 = 32 * t0, j1 = min(n, 32 * min((n + 31) >> 5, t1));
        const int h = slot - s.slot[c];
        const int jA = j0 + h, jB = j1 - 1 - h;   // the short and the long column of this lane
        const bool hasA = act && jA < jB;
        const int tripA = hasA ? jA : 0, tripB = act ? jB : 0;  // rows 0..trip-1 precede the column
        const uint32_t *ord = s.sord + rowbase;
        float4 ca, cb;
        float2 ta, tb;
        lds_rec(ord[hasA ? jA : 0], ca, ta);
        lds_rec(ord[act ? jB : 0], cb, tb);
        const int maxA = __reduce_max_sync(kFullMask, tripA), maxB = __reduce_max_sync(kFullMask, tripB);
        const int ngA = (maxA + 15) >> 4, ngB = (maxB + 15) >> 4;  // 16-row groups
        const int ctA = jA >> 5, ctB = jB >> 5;
        uint32_t *mA = s.mask + s.maskbase[c] + (tri(ctA) - tri0) * 32 + (jA & 31);
        uint32_t *mB = s.mask + s.maskbase[c] + (tri(ctB) - tri0) * 32 + (jB & 31);
        uint32_t wA = 0u, wB = 0u;
        for (int g = 0; g < ngB; ++g) {
            uint32_t a16 = 0u, b16;
            if (g < ngA) pair_group16<true>(ord + 16 * g, ca, ta.x, ta.y, cb, tb.x, tb.y, p.iou, a16, b16);
            else pair_group16<false>(ord + 16 * g, ca, ta.x, ta.y, cb, tb.x, tb.y, p.iou, a16, b16);
            if (!(g & 1)) {
                wA = a16;
                wB = b16;
                if (g + 1 < ngB) continue;
            } else {
                wA |= a16 << 16;
                wB |= b16 << 16;
            }
            // rows >= the column index (own bit, later rows, other classes) are masked off
            const int w = g >> 1;
            if (act && w <= ctB) mB[w * 32] = (w == ctB) ? (wB & ((1u << (jB & 31)) - 1u)) : wB;
            if (hasA && w <= ctA) mA[w * 32] = (w == ctA) ? (wA & ((1u << (jA & 31)) - 1u)) : wA;
        }
        // ceil(j/32) row words were computed; the diagonal word of a column with
        // j % 32 == 0 holds no earlier row and may lie beyond the warp's loop
        const int nwB = (ngB + 1) >> 1;
        if (act && nwB <= ctB) mB[ctB * 32] = 0u;
        if (hasA && nwB <= ctA) mA[ctA * 32] = 0u;
    }
}

// ---------------------------------------------------------------------------
// P6: sweep -- one warp per class, 32-column tiles in score order
// ---------------------------------------------------------------------------
template <int THREADS>
__device__ __forceinline__ void phase_sweep(const Smem &s) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c_lo = s.misc[M_CLO], c_hi = s.misc[M_CHI], t0 = s.misc[M_T0], t1 = s.misc[M_T1];
    const int tri0 = tri(t0);
    for (int c = c_lo + warp; c < c_hi; c += THREADS / 32) {
        const int n = s.cnt[c];
        const int T = (n + 31) >> 5;
        const int te = min(T, t1);
        uint32_t *kept_w = s.keptbits + s.ktile[c];
        const uint32_t *mbase = s.mask + s.maskbase[c];
        for (int ct = t0; ct < te; ++ct) {
            const int j = 32 * ct + lane;
            const uint32_t *col = mbase + (tri(ct) - tri0) * 32 + lane;
            uint32_t sup = 0u;
#pragma unroll 2
            for (int rt = 0; rt < ct; ++rt) sup |= col[rt * 32] & kept_w[rt];
            const bool alive = (j < n) && (sup == 0u);
            const unsigned alive_mask = __ballot_sync(kFullMask, alive);
            const uint32_t diag = alive ? (col[ct * 32] & alive_mask) : 0u;
            const unsigned nz = __ballot_sync(kFullMask, diag != 0u);
            // columns without any alive earlier overlap are kept outright; the rest are
            // decided in parallel as soon as all their earlier overlaps are decided (the
            // lowest undecided column always is, so every pass makes progress)
            unsigned kept = alive_mask & ~nz;
            for (unsigned und = nz; und;) {
                const bool mine = (und >> lane) & 1u;
                const bool dead = mine && (diag & kept) != 0u;
                const bool keep = mine && !dead && (diag & und) == 0u;
                const unsigned d = __ballot_sync(kFullMask, dead), k = __ballot_sync(kFullMask, keep);
                kept |= k;
                und &= ~(d | k);
            }
            if (lane == 0) kept_w[ct] = kept;
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------
template <int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS, (THREADS == 512) ? 2 : 1) decode_nms_kernel(const DNParams p, const SmemLayout L) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem s = carve(smem_raw, L, p.C, MODE);
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = p.C, K = p.K;
    constexpr int kWarps = THREADS / 32;

    stamp(p, b, 0);
    if (MODE != MODE_NMS) {
        // one thread per (head, anchor) slab: start the HBM -> L2 stream of this image now
        if (tid == 0 && !(p.flags & 1)) {
            const size_t img0 = (size_t)p.A * p.attrs * p.head[0].HW;
            l2_prefetch_span(p.head[0].ptr + (size_t)b * img0, img0 * sizeof(float));
            if (p.nheads > 1 && !(p.flags & 2)) {
                const size_t img1 = (size_t)p.A * p.attrs * p.head[1].HW;
                l2_prefetch_span(p.head[1].ptr + (size_t)b * img1, img1 * sizeof(float));
            }
        }
    }
    if (MODE != MODE_DECODE) {
        for (int i = tid; i <= C * p.B; i += THREADS) s.cntb[i] = 0;
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
        const float *recf = reinterpret_cast<const float *>(s.rec);
        for (int f = tid; f < 7 * T; f += THREADS) {
            const int row = f / 7, col = f - 7 * row;
            const int cid = s.outsrc[row];
            o[f] = (col < 6) ? recf[8 * cid + col] : (float)(s.clsidx[cid] >> 16);  // cls_idx.float() :199
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
        warp_class_scan(p, s);
        warp_round_prefix(s, 0);
    }
    phase_scatter_keys<THREADS>(p, s);
    __syncthreads();
    stamp(p, b, 2);
    const int Kv = s.misc[M_KV];
    const int nrounds = s.misc[M_NROUNDS];
    // P4
    phase_rank_sort<THREADS>(p, s, Kv, 2 * (int)align_up((uint32_t)K, 32) + 64);
    __syncthreads();
    stamp(p, b, 3);
    // P5, P6
    for (int r = 0;;) {
        phase_pairs(p, s);
        __syncthreads();
        stamp(p, b, 4);
        phase_sweep<THREADS>(s);
        __syncthreads();
        stamp(p, b, 5);
        if (++r >= nrounds) break;
        if (warp == 0) warp_round_prefix(s, r);
        __syncthreads();
    }
    // P7/P8: output row -> cell id (the mask buffer is dead now; outsrc aliases it).  Every
    // warp sums the kept counts of the tiles before its class itself (no serial scan phase).
    const int ntiles = s.ktile[C];
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
                const uint32_t cid = (s.sord[st + 32 * ct + lane] - s.rec_saddr) >> 5;
                s.outsrc[before + __popc(word & lanemask_lt())] = (uint16_t)cid;
            }
            before += __popc(word);
        }
    }
    (void)ntiles;
    __syncthreads();
    stamp(p, b, 6);
    // P9: flat coalesced store
    {
        const int T = s.misc[M_TOTAL];
        float *o = p.out + (size_t)b * K * 7;
        const float *recf = reinterpret_cast<const float *>(s.rec);
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
            for (int f = tid; f < 7 * T; f += THREADS) {
                const int row = f / 7, col = f - 7 * row;
                const int cid = s.outsrc[row];
                o[f] = (col < 6) ? recf[8 * cid + col] : (float)(s.clsidx[cid] >> 16);
            }
        }
        if (p.out_idx)
            for (int r = tid; r < T; r += THREADS) p.out_idx[(size_t)b * K + r] = (int)s.outsrc[r];
        if (tid == 0) p.out_count[b] = T;
    }
    stamp(p, b, 7);
}

}  // namespace b200yolo
