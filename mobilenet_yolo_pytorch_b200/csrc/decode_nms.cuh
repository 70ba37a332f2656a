// decode_nms.cuh -- YOLO-head decode + confidence threshold + per-class NMS (sm_100a).
//
// One CTA per image; everything about one image lives in shared memory, the head
// tensors are read from HBM exactly once and only kept rows are written back.
//
// Shared memory is kept SMALL on purpose (<= 80 KB per CTA for the 352x352 VOC heads):
// the decode phase streams the heads with ordinary coalesced loads whose in-flight
// lines live in L1, and L1 is what the CTAs' shared memory leaves of the SM's 228 KB
// (profiles/micro/load_pattern.cu: the same load pattern runs at 4.3 TB/s with 64 KB of
// L1 and at 2.9 TB/s with none).  So: structure-of-arrays records by cell id (box 16 B,
// conf/score 8 B) and one region `U` that is reused by phase.
//
// The kernel is issue-bound once the heads are in (profiles/r01: 0.67 of the issue
// slots used while an SM is active), so every phase is written for instruction count:
//
//   P1 decode     thread per cell, ALL 5+C attribute planes of the cell loaded at
//                 once (coalesced: consecutive lanes = consecutive cells of a plane).
//                 The reference's shapes (VOC 352 / 416, BDD 640x384) are compiled with
//                 the plane stride as a constant, so the 5+C loads are one base register
//                 plus immediates; any other shape takes the runtime-stride path.
//                 conf = sigmoid(tc) > thr (yolo_loss.py:189,201); for passing cells:
//                 class max / argmax (:198), box (:186-196,243-247) -> box[], cs[],
//                 and the per-(class, score bucket) arrival index (one shared atomic).
//   P2 scans      warp per class: exclusive scan of the class's score-bucket histogram; warp 0:
//                 class segment starts (box.py:20-22).  Then warp 0 builds the kept-bitmap
//                 tiles, the round table and the strip tasks WHILE the other warps sort (named
//                 barrier between P3 and P4): the tables are off the critical path.
//   P3 key scatter 64-bit keys (score desc, candidate order asc == the stable sort of
//                 torchvision.ops.nms) into their (class, bucket) segment.
//   P4 rank       rank inside the bucket -> `sord`: sorted position -> {shared-memory
//                 address of the box, t * area}.  The pair loop then needs two loads and no
//                 address arithmetic per column.
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
//   P6 output     class-ascending / score-descending rows (box.py:29-30).  A warp owns
//                 32-column tiles of the kept bitmap: the rows of a tile are consecutive
//                 in the output, so the warp assembles them in a 896-byte scratch and
//                 stores them with coalesced 4-byte stores.
//
// If the masks of all classes do not fit `U`, P5 runs in rounds (groups of whole
// classes, or column-tile chunks of one huge class) with a block barrier in between.
//
// The stand-alone decode (P1 + ordered compaction + store) and NMS (load rows,
// P2..P6) kernels back YOLOLoss.forward(input) and utils.box.nms separately.
//
// Variants of the same kernel: channels-last heads (decode_head_nhwc: warps stage 32 cells and transpose through
// shared memory, no NCHW copy); GATHER = 1 (b200yolo_decode_nms_gather): P6 stores every tile of kept rows into the
// gather buffer of EVERY rank of the box over NVLink peer mappings -- the data-parallel all-gather fused into the
// kernel.  Images with more cells than `U` can hold take large_nms.cuh.
//
// Consecutive launches overlap (programmatic dependent launch, see pdl_trigger / pdl_wait below): a launch
// starts on the SM slots its predecessor leaves free and streams its heads under the predecessor's NMS.
#pragma once
#include "common.cuh"

namespace b200yolo {

constexpr int kMaxAnchors = 8;
constexpr int kMaxPeers = 8;        // GPUs of one NVSwitch box

enum { MODE_FUSED = 0, MODE_DECODE = 1, MODE_NMS = 2 };

// Compile-time head shapes (0 = runtime).  SHAPE ids are chosen by the host from (C, H, W) of both heads.
template <int SHAPE> struct ShapeT { static constexpr int C = 0, HW0 = 0, W0 = 0, HW1 = 0, W1 = 0; };
template <> struct ShapeT<1> { static constexpr int C = 20, HW0 = 121, W0 = 11, HW1 = 484, W1 = 22; };  // VOC 352x352 (models/voc/config.yaml)
template <> struct ShapeT<2> { static constexpr int C = 20, HW0 = 169, W0 = 13, HW1 = 676, W1 = 26; };  // 416x416 (inference.py:112)
template <> struct ShapeT<3> { static constexpr int C = 10, HW0 = 240, W0 = 20, HW1 = 960, W1 = 40; };  // BDD100k 640x384, 10 classes
// single heads of the same configurations (the stand-alone decode kernel, YOLOLoss.forward(input))
template <> struct ShapeT<11> { static constexpr int C = 20, HW0 = 121, W0 = 11, HW1 = 0, W1 = 0; };
template <> struct ShapeT<12> { static constexpr int C = 20, HW0 = 484, W0 = 22, HW1 = 0, W1 = 0; };
template <> struct ShapeT<13> { static constexpr int C = 20, HW0 = 169, W0 = 13, HW1 = 0, W1 = 0; };
template <> struct ShapeT<14> { static constexpr int C = 20, HW0 = 676, W0 = 26, HW1 = 0, W1 = 0; };
template <> struct ShapeT<15> { static constexpr int C = 10, HW0 = 240, W0 = 20, HW1 = 0, W1 = 0; };
template <> struct ShapeT<16> { static constexpr int C = 10, HW0 = 960, W0 = 40, HW1 = 0, W1 = 0; };
// channels-last heads (decode_head_nhwc): only the class count is a compile-time constant
template <> struct ShapeT<21> { static constexpr int C = 20, HW0 = 0, W0 = 0, HW1 = 0, W1 = 0; };
template <> struct ShapeT<22> { static constexpr int C = 10, HW0 = 0, W0 = 0, HW1 = 0, W1 = 0; };
template <int SHAPE> struct ShapeIsNhwc { static constexpr bool value = SHAPE >= 20 && SHAPE < 30; };

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
    int flags;           // experiment switches (B200YOLO_FLAGS env): 1 = no L2 prefetch, 8 = prefetch both heads
    int nhwc;            // > 0: heads are channels-last, (N, H, W, A*(5+C)) in memory (fused mode only); the value is the
                         // number of warps that stage + decode (32 cells each per step; what the free shared memory allows)
    unsigned long long *dbg;  // optional [N][32] phase time stamps (16 x globaltimer ns, 16 x SM clock), NULL in production
    float conf_thr;
    IouThr iou;
    float *out;
    int *out_count;
    int *out_idx;
    // fused all-gather (b200yolo_decode_nms_gather): the output phase stores every kept row into the gather buffer of
    // EVERY rank (its own and, through NVLink peer mappings, the others'), image slot gslot + b; out / out_count unused
    int gR;                      // ranks (0: ordinary single-buffer output)
    int gslot;                   // first image slot of this rank = rank * N
    float *gout[kMaxPeers];      // [gR] rank r's buffer [gR*N][K][7]
    int *gcount[kMaxPeers];      // [gR] rank r's counts [gR*N]
    // MODE_NMS inputs
    const float *cand[2];
    const int *cand_count[2];
    int cand_stride[2];
};

struct SmemLayout {
    uint32_t box, cs, cntb, cls, tasks, tilecls, keptbits, tilepref, passbits, rounds, misc, U, total;
    uint32_t u_bytes;     // bytes in U
    uint32_t mask_off;    // keys, then pair masks, then the output scratch start here (after sord)
    uint32_t mask_words;  // 32-bit words available for pair masks
};

__host__ __device__ inline uint32_t align_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

// per-class int arrays, each (C+1) long
enum { CA_CNT = 0, CA_START, CA_KTILE, CA_MASK, CA_TBASE, CA_DONE, CA_FLAG, CA_NUM };
// misc ints
enum { M_NROUNDS = 0, M_CLO, M_CHI, M_T0, M_T1, M_NTASK, M_CTR, M_TOTAL, M_KV, M_WSUM = 16, M_NUM = 16 + 32 };

// score buckets per class: C*B counters, at most 1536 (6 KB)
__host__ __device__ inline int pick_buckets(int C) {
    int B = 64;
    while (B > 1 && C * B > 1536) B >>= 1;
    return B;
}

constexpr uint32_t kScratchPerWarp = 32 * 7 * 4;  // P6: 32 output rows of 7 floats

// U region by phase (Kp = K rounded up to 32):
//   decode .. key scatter   clsidx u32[Kp] ........ | key u64[Kp]
//   rank .. sweep           sord uint2[Kp + 8] .... | (key, dead after rank) pair masks u32[mask_words]
//   output                  sord .................. | per-warp row scratch
//   (MODE_DECODE)           clsidx u32[Kp] | outsrc u16[Kp]
__host__ __device__ inline SmemLayout make_layout(int K, int C, int mode, int threads, uint32_t extra_mask_bytes) {
    SmemLayout L;
    const uint32_t Kp = align_up((uint32_t)(K > 0 ? K : 1), 32);
    const uint32_t Cp = align_up((uint32_t)C + 1, 4);
    const uint32_t tiles = Kp / 32 + (uint32_t)C + 1;
    const bool nms = (mode != MODE_DECODE);
    uint32_t o = 0;
    L.box = o; o += 16 * Kp;
    L.cs = o; o += 8 * Kp;
    L.cntb = o; o += nms ? 4 * align_up((uint32_t)(C * pick_buckets(C)) + 1, 4) : 0;
    L.cls = o; o += nms ? 4 * Cp * CA_NUM : 0;
    L.tasks = o; o += nms ? 4 * tiles : 0;
    L.keptbits = o; o += nms ? 4 * tiles : 0;
    L.tilepref = o; o += nms ? 0 : 4 * (tiles + 1);
    L.passbits = o; o += nms ? 0 : 4 * (Kp / 32);
    L.tilecls = o; o += nms ? align_up(2 * tiles, 4) : 0;
    o = align_up(o, 8);
    L.rounds = o; o += nms ? 8 * (tiles + (uint32_t)C + 2) : 0;
    L.misc = o; o += 4 * M_NUM;
    o = align_up(o, 16);
    L.U = o;
    const uint32_t scratch = (uint32_t)(threads / 32) * kScratchPerWarp;
    uint32_t tail = 8 * Kp + (extra_mask_bytes & ~15u);  // keys; pair masks; output scratch
    if (tail < scratch) tail = scratch;
    const uint32_t sord_bytes = nms ? 8 * Kp + 64 : 4 * Kp;  // + 8 padding entries (the pair loop reads whole chunks of 8)
    L.u_bytes = nms ? sord_bytes + tail : 6 * Kp;
    L.mask_off = L.U + sord_bytes;
    L.mask_words = (L.u_bytes - sord_bytes) / 4;
    L.total = align_up(L.U + L.u_bytes, 16);
    return L;
}

struct Smem {
    float4 *box;        // [Kp] x1 y1 x2 y2 by cell id            (output columns 0-3)
    float2 *cs;         // [Kp] conf, class score                 (columns 4, 5)
    uint32_t *clsidx;   // [Kp] (class << 16) | arrival index inside the (class, bucket); ~0u: not a candidate
    unsigned long long *key;
    uint2 *sord;        // sorted position -> {shared address of box[cell], t * area * 2^-13 (NaN: degenerate)}
    uint32_t *mask;     // pair-mask words
    float *scratch;     // P6: per-warp 32 x 7 floats
    uint16_t *outsrc;   // MODE_DECODE: output row -> cell id
    uint16_t *tilecls;  // kept-bitmap tile -> class
    uint32_t *passbits, *keptbits, *tilepref, *tasks;
    int *cnt, *start, *ktile, *maskbase, *tbase, *done, *flag;
    int *cntb;          // [C*B+1] per (class, score bucket): arrival counter, then exclusive prefix
    uint2 *rounds;      // x = c_lo | c_hi << 16, y = t0 | t1 << 16
    int *misc;
    uint32_t box_saddr; // shared-window address of box[0]
};

__device__ __forceinline__ Smem carve(unsigned char *base, const SmemLayout &L, int K, int C, int mode) {
    Smem s;
    const uint32_t Kp = align_up((uint32_t)(K > 0 ? K : 1), 32);
    const uint32_t Cp = align_up((uint32_t)C + 1, 4);
    s.box = reinterpret_cast<float4 *>(base + L.box);
    s.cs = reinterpret_cast<float2 *>(base + L.cs);
    s.clsidx = reinterpret_cast<uint32_t *>(base + L.U);
    s.key = reinterpret_cast<unsigned long long *>(base + L.mask_off);
    s.sord = reinterpret_cast<uint2 *>(base + L.U);
    s.mask = reinterpret_cast<uint32_t *>(base + L.mask_off);
    s.scratch = reinterpret_cast<float *>(base + L.mask_off);
    s.outsrc = reinterpret_cast<uint16_t *>(base + L.U + 4 * Kp);
    s.tilecls = reinterpret_cast<uint16_t *>(base + L.tilecls);
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
    s.box_saddr = (uint32_t)__cvta_generic_to_shared(base + L.box);
    (void)mode;
    return s;
}

__device__ __forceinline__ int fastdiv(int n, uint32_t magic) {
    return magic ? (int)__umulhi((uint32_t)n, magic) : n;  // exact for n < 65536
}

__device__ __forceinline__ int tri(int x) { return (x * (x + 1)) >> 1; }

// profiling aid: slot k of image b gets %globaltimer (ns, 256 ns resolution, comparable across SMs) and,
// 16 slots further, the SM's cycle counter (comparable inside the CTA only)
__device__ __forceinline__ void stamp(const DNParams &p, int b, int k) {
    if (p.dbg && threadIdx.x == 0) {
        if (k == 0 || k == 7) {  // (reading %globaltimer costs several hundred ns: only at the CTA's two ends)
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            p.dbg[(size_t)b * 32 + k] = t;
        }
        p.dbg[(size_t)b * 32 + 16 + k] = (unsigned long long)clock64();
    }
}
// inside the decode loop the check is hoisted: `on` = p.dbg != nullptr && threadIdx.x == 0
__device__ __forceinline__ void stamp_if(bool on, const DNParams &p, int b, int k) {
    if (on) p.dbg[(size_t)b * 32 + 16 + k] = (unsigned long long)clock64();
}

// Programmatic dependent launch (the host sets cudaLaunchAttributeProgrammaticStreamSerialization): the NEXT
// kernel in the stream may start once every CTA of this one has executed pdl_trigger(); it must execute
// pdl_wait() -- which returns when this grid has completed and its writes are visible -- before it touches
// anything this kernel writes.  Only a kernel that triggers can be overtaken, so an ordinary producer of the
// head tensors (a convolution) still completes before any of our CTAs starts; between two of our own launches
// the only shared data are the output buffers, which every mode writes after pdl_wait().  Without the launch
// attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// shared-memory loads by shared-window address (the pair loop: no address arithmetic)
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
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

constexpr int kClsChunk = 24;  // class planes loaded per batch on the runtime-shape path

// Window below the largest logit m inside which another logit could tie with it after the sigmoid is
// rounded: d/dt ln(sigmoid(t)) = 1 - sigmoid(t) >= e*s on (-inf, m] (e = exp(-m), s = sigmoid(m)), so a
// logit below m - 2^-17/(e*s) has a sigmoid smaller by > 2^-17 relative (>10x the evaluation error) and
// cannot win or tie.  Returns sigmoid(m) in *best.
__device__ __forceinline__ float tie_window(float m, float *best) {
    const float e1 = exp_fast(-m);
    const float s1 = rcp_fast(__fadd_rn(1.0f, e1));
    *best = s1;
    return __fmul_rn(7.6293945e-06f, rcp_fast(__fmul_rn(e1, s1)));
}

// ---------------------------------------------------------------------------
// P1: decode every cell of one head, single pass.  All 5+C plane loads of a cell are
// issued before the first use.
// ---------------------------------------------------------------------------
// box arithmetic + records of one passing cell (yolo_loss.py:186-199, 243-247)
template <int MODE>
__device__ __forceinline__ void emit_candidate(const DNParams &p, const Smem &s, const HeadDesc &hd, int cid, int a, int i, int j,
                                               float tx, float ty, float tw, float th, float conf, float best, int bi) {
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
    if (MODE == MODE_FUSED) idx = (uint32_t)atomicAdd(&s.cntb[bi * p.B + score_bucket(__fmul_rn(best, conf), p.B)], 1);
    s.clsidx[cid] = ((uint32_t)bi << 16) | idx;
}

// Compile-time class count and grid: the 5+C loads are one base register plus immediates.  FIRST: this
// is the first decode of the kernel -- the block barrier that orders the zeroing of the histogram before
// the first shared atomic is taken AFTER the first round's loads are in flight (hides ~0.5 us of start-up
// behind the first HBM round trip).
template <int THREADS, int MODE, int CT, int HWT, int WT, bool FIRST>
__device__ __forceinline__ void decode_head_static(const DNParams &p, const Smem &s, int b, const HeadDesc &hd, int cid0, int hh) {
    static_assert(CT >= 1 && CT <= 24, "compile-time shapes keep all class bits in one fp32 accumulator");
    const int tid = threadIdx.x, lane = tid & 31;
    const bool dbg_on = p.dbg != nullptr && tid == 0;
    constexpr int attrs = CT + 5;
    const int cells = p.A * HWT;
    const float *hb = hd.ptr + (size_t)b * p.A * attrs * HWT;  // uniform
#pragma unroll 1
    for (int base = 0; base < cells; base += THREADS) {
        const int local = base + tid;
        const bool active = local < cells;
        const int a = local / HWT;
        const int pos = local - a * HWT;
        const float *q = hb + (uint32_t)(a * attrs * HWT + pos);
        float tx = 0.f, ty = 0.f, tw = 0.f, th = 0.f, tc = 0.f;
        float x[CT];
        if (active) {
            tx = __ldcs(q);
            ty = __ldcs(q + HWT);
            tw = __ldcs(q + 2 * HWT);
            th = __ldcs(q + 3 * HWT);
            tc = __ldcs(q + 4 * HWT);
#pragma unroll
            for (int u = 0; u < CT; ++u) x[u] = __ldcs(q + (5 + u) * HWT);
        }
        if (FIRST && base == 0) __syncthreads();
        bool pass = false;
        if (active) {
            const int cid = cid0 + local;
            const float conf = sigmoid_fast(tc);  // yolo_loss.py:189,197
            pass = conf > p.conf_thr;             // :201 (threshold already rounded to fp32)
            if (pass) {
                float m1 = x[0];
#pragma unroll
                for (int u = 1; u < CT; ++u) m1 = fmaxf(m1, x[u]);
                float best;
                const float win = tie_window(m1, &best);
                const float lo = __fsub_rn(m1, win);
                float near = 0.f;  // bit u: x[u] >= lo   (FSET + FFMA: exact for 24 bits)
#pragma unroll
                for (int u = 0; u < CT; ++u) near = __fmaf_rn((x[u] >= lo) ? 1.0f : 0.0f, (float)(1u << u), near);
                const uint32_t nb = __float2uint_rn(near);
                int bi = nb ? __ffs(nb) - 1 : 0;
                const bool tie = (nb & (nb - 1u)) != 0u || nb == 0u;
                if (CT > 1 && tie) best = class_tie_break(q + 5 * HWT, HWT, CT, lo, m1, bi, &bi);
                const int j = pos / WT;
                emit_candidate<MODE>(p, s, hd, cid, a, pos - j * WT, j, tx, ty, tw, th, conf, best, bi);
            } else if (MODE == MODE_FUSED) {
                s.clsidx[cid] = 0xffffffffu;
            }
        }
        if (MODE == MODE_DECODE) {  // single head: candidate ids are 32-aligned per warp
            const unsigned bal = __ballot_sync(kFullMask, pass);
            if (lane == 0 && active) s.passbits[local >> 5] = bal;
        }
        if (MODE == MODE_FUSED) stamp_if(dbg_on, p, b, 8 + min(3, hh + base / THREADS));
    }
}

// Runtime class count and grid (any shape): plane stride in a register, classes in chunks of 24.
template <int THREADS, int MODE>
__device__ __forceinline__ void decode_head_rt(const DNParams &p, const Smem &s, int b, const HeadDesc &hd, int cid0, int hh) {
    const int tid = threadIdx.x, lane = tid & 31;
    const bool dbg_on = p.dbg != nullptr && tid == 0;
    const int C = p.C;
    const int attrs = C + 5;
    const int HW = hd.HW;
    const int cells = p.A * HW;
    const float *hb = hd.ptr + (size_t)b * p.A * attrs * HW;  // uniform
#pragma unroll 1
    for (int base = 0; base < cells; base += THREADS) {
        const int local = base + tid;
        bool pass = false;
        if (local < cells) {
            const int cid = cid0 + local;
            const int a = fastdiv(local, hd.magicHW);
            const int pos = local - a * HW;
            const float *q = hb + (uint32_t)(a * attrs * HW + pos);
            float m1 = -INFINITY, best = 0.f, win = 0.f, conf = 0.f;
            int i1 = 0;
            bool tie = false;
            const char *qb = reinterpret_cast<const char *>(q);
            const uint32_t st = (uint32_t)HW * 4u;  // plane stride in bytes
#define B200_LD(u) __ldcs(reinterpret_cast<const float *>(qb + (uint64_t)st * (uint32_t)(u)))
            const float tx = B200_LD(0);
            const float ty = B200_LD(1);
            const float tw = B200_LD(2);
            const float th = B200_LD(3);
            const float tc = B200_LD(4);
            qb += (uint64_t)st * 5u;
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
                qb += (uint64_t)st * (uint32_t)nv;
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
                    const float m_new = fmaxf(m1, cm);
                    win = tie_window(m_new, &best);
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
#undef B200_LD
            if (pass) {
                int bi = i1;
                if (C > 1 && tie) best = class_tie_break(q + 5 * HW, HW, C, __fsub_rn(m1, win), m1, i1, &bi);
                const int j = fastdiv(pos, hd.magicW);
                emit_candidate<MODE>(p, s, hd, cid, a, pos - j * hd.W, j, tx, ty, tw, th, conf, best, bi);
            } else if (MODE == MODE_FUSED) {
                s.clsidx[cid] = 0xffffffffu;
            }
        }
        if (MODE == MODE_DECODE) {  // single head: candidate ids are 32-aligned per warp
            const unsigned bal = __ballot_sync(kFullMask, pass);
            if (lane == 0 && local < cells) s.passbits[local >> 5] = bal;
        }
        if (MODE == MODE_FUSED) stamp_if(dbg_on, p, b, 8 + min(3, hh + base / THREADS));
    }
}

// Channels-last heads (what cuDNN prefers for the head's last convolution): in memory the image is
// (H, W, A*(5+C)), so the 5+C values of a cell are contiguous and the cells follow each other in (j, i, a) order.
// The kernel is issue-bound after the loads (profiles/r01/NOTES.md), so what counts is instructions per cell:
// a staging warp copies 32 consecutive cells (contiguous floats: perfectly coalesced, 5+C words per lane) into
// its shared-memory scratch, then every lane reads ONE cell at stride 5+C (conflict-free for the odd 5+C of the
// reference's configurations) and runs the same arithmetic as the planar paths; the loads of the warp's next 32
// cells are issued before that arithmetic, so they are in flight while it runs.  Only p.nhwc warps stage -- as
// many as the free part of U holds scratch for (7 for the VOC heads) -- the others skip the decode: a warp that
// fills half of its lanes costs the same issue slots as a full one.  Records keep the reference's candidate order
// (a, j, i).  CT: compile-time class count (0 = runtime).
constexpr int kNhwcMaxAttrs = 32;   // 5 + C <= 32 on this path

template <int THREADS, int CT>
__device__ __forceinline__ void decode_head_nhwc(const DNParams &p, const Smem &s, int b, const HeadDesc &hd, int cid0, float *scr_base) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = p.nhwc;
    if (warp >= nwarps) return;
    const int C = CT ? CT : p.C, attrs = 5 + C, A = p.A;
    const int cells = hd.cells;
    const float *hb = hd.ptr + (size_t)b * cells * attrs;  // uniform
    float *scr = scr_base + warp * (32 * attrs);
    constexpr int kLoads = CT ? 5 + CT : kNhwcMaxAttrs;    // words per lane per step
    float v[kLoads];
    auto issue = [&](int base) {
        const int nw = min(32, cells - base) * attrs;
        const float *src = hb + (size_t)base * attrs;
#pragma unroll
        for (int k = 0; k < kLoads; ++k) {
            const int w = 32 * k + lane;
            v[k] = (w < nw) ? __ldcs(src + w) : 0.f;
        }
    };
    int base = warp * 32;
    if (base < cells) issue(base);
#pragma unroll 1
    for (; base < cells; base += nwarps * 32) {
        const int nc = min(32, cells - base);
        const int nw = nc * attrs;
#pragma unroll
        for (int k = 0; k < kLoads; ++k) {
            const int w = 32 * k + lane;
            if (w < nw) scr[w] = v[k];
        }
        __syncwarp();
        if (base + nwarps * 32 < cells) issue(base + nwarps * 32);  // in flight during the arithmetic below
        if (lane < nc) {
            const int m = base + lane;                       // memory order: (j*W + i)*A + a
            const int pos = m / A, a = m - pos * A;
            const int j = fastdiv(pos, hd.magicW), i = pos - j * hd.W;
            const int cid = cid0 + a * hd.HW + pos;          // reference order: (a*H + j)*W + i
            const float *xs = scr + lane * attrs;
            const float conf = sigmoid_fast(xs[4]);         // yolo_loss.py:189,197
            if (conf > p.conf_thr) {                        // :201
                float best;
                int bi;
                if constexpr (CT > 0) {
                    float x[CT];
#pragma unroll
                    for (int u = 0; u < CT; ++u) x[u] = xs[5 + u];
                    float m1 = x[0];
#pragma unroll
                    for (int u = 1; u < CT; ++u) m1 = fmaxf(m1, x[u]);
                    const float win = tie_window(m1, &best);
                    const float lo = __fsub_rn(m1, win);
                    float near = 0.f;  // bit u: x[u] >= lo   (FSET + FFMA: exact for 24 bits)
#pragma unroll
                    for (int u = 0; u < CT; ++u) near = __fmaf_rn((x[u] >= lo) ? 1.0f : 0.0f, (float)(1u << u), near);
                    const uint32_t nb = __float2uint_rn(near);
                    bi = nb ? __ffs(nb) - 1 : 0;
                    const bool tie = (nb & (nb - 1u)) != 0u || nb == 0u;
                    if (CT > 1 && tie) best = class_tie_break(hb + (size_t)m * attrs + 5, 1, CT, lo, m1, bi, &bi);
                } else {
                    float m1 = xs[5];
                    for (int c = 1; c < C; ++c) m1 = fmaxf(m1, xs[5 + c]);
                    const float win = tie_window(m1, &best);
                    const float lo = __fsub_rn(m1, win);
                    int nnear = 0;
                    bi = -1;
                    for (int c = 0; c < C; ++c) {
                        const bool nr = xs[5 + c] >= lo;
                        if (nr && bi < 0) bi = c;
                        nnear += nr ? 1 : 0;
                    }
                    bi = max(bi, 0);
                    if (C > 1 && nnear != 1) best = class_tie_break(hb + (size_t)m * attrs + 5, 1, C, lo, m1, bi, &bi);
                }
                emit_candidate<MODE_FUSED>(p, s, hd, cid, a, i, j, xs[0], xs[1], xs[2], xs[3], conf, best, bi);
            } else {
                s.clsidx[cid] = 0xffffffffu;
            }
        }
        __syncwarp();
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
// (warp 0) class segment starts: exclusive scan of cnt[] -> start[]; total -> misc[M_KV]
__device__ __forceinline__ void warp_class_starts(const DNParams &p, const Smem &s) {
    const int lane = threadIdx.x & 31;
    const int C = p.C;
    int carry = 0;
    for (int c0 = 0; c0 < C; c0 += 32) {
        const int c = c0 + lane;
        const int n = (c < C) ? s.cnt[c] : 0;
        const int inc = warp_inclusive_scan(n, lane);
        if (c < C) s.start[c] = carry + inc - n;
        carry += __shfl_sync(kFullMask, inc, 31);
    }
    if (lane == 0) {
        s.cnt[C] = 0;
        s.start[C] = carry;
        s.misc[M_KV] = carry;
    }
    __syncwarp();
}

__device__ __forceinline__ void warp_class_scan(const DNParams &p, const Smem &s, int mask_cap_words) {
    const int lane = threadIdx.x & 31;
    const int C = p.C;
    int carryT = 0, words = 0;
    for (int c0 = 0; c0 < C; c0 += 32) {
        const int c = c0 + lane;
        const int n = (c < C) ? s.cnt[c] : 0;
        const int T = (n + 31) >> 5;
        const int incT = warp_inclusive_scan(T, lane);
        if (c < C) {
            const int kt = carryT + incT - T;
            s.ktile[c] = kt;
            for (int t = 0; t < T; ++t) s.tilecls[kt + t] = (uint16_t)c;
        }
        carryT += __shfl_sync(kFullMask, incT, 31);
        words += 32 * tri(T);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) words += __shfl_xor_sync(kFullMask, words, o);
    __syncwarp();  // lane 0 reads the other lanes' cnt[] below
    if (lane == 0) {
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
// P2a (warp per class): exclusive scan of the class's score-bucket counters in place (position of the
// bucket INSIDE the class segment); cnt[c] = candidates of the class
// ---------------------------------------------------------------------------
template <int THREADS>
__device__ __forceinline__ void scan_buckets_per_class(const DNParams &p, const Smem &s) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int B = p.B;                    // power of two <= 64
    const int per = (B + 31) >> 5;        // 1 or 2 consecutive buckets per lane
    for (int c = warp; c < p.C; c += THREADS / 32) {
        int *cb = s.cntb + c * B;
        const int i0 = lane * per;
        const int v0 = (i0 < B) ? cb[i0] : 0;
        const int v1 = (per == 2) ? cb[i0 + 1] : 0;
        const int sum = v0 + v1;
        const int inc = warp_inclusive_scan(sum, lane);
        if (i0 < B) cb[i0] = inc - sum;
        if (per == 2) cb[i0 + 1] = inc - sum + v0;
        if (lane == 31) s.cnt[c] = inc;
    }
}

// ---------------------------------------------------------------------------
// P3/P4: class-major counting sort on score buckets, then rank inside the bucket:
// the stable descending score sort of torchvision.ops.nms, per class
// key = score order key (32) | class (16) | 0xffff - cell id (16)
// ---------------------------------------------------------------------------
template <int THREADS>
__device__ __forceinline__ void phase_scatter_keys(const DNParams &p, const Smem &s, int tid, int nthr) {
    for (int cid = tid; cid < p.K; cid += nthr) {
        const uint32_t ci = s.clsidx[cid];
        if (ci == 0xffffffffu) continue;
        const float2 cs = s.cs[cid];
        const float sc = __fmul_rn(cs.y, cs.x);  // box.py:27 scores = col5*col4
        const uint32_t c = ci >> 16;
        const unsigned long long key =
            ((unsigned long long)float_order_key(sc) << 32) | (unsigned long long)((c << 16) | (0xffffu - (uint32_t)cid));
        s.key[s.start[c] + s.cntb[(int)c * p.B + score_bucket(sc, p.B)] + (int)(ci & 0xffffu)] = key;
    }
}

template <int THREADS>
__device__ __forceinline__ void phase_rank_sort(const DNParams &p, const Smem &s, int Kv, int tid, int nthr) {
    if (tid < 8) s.sord[Kv + tid] = make_uint2(s.box_saddr, 0x7fc00000u);  // padding: a valid address, NaN area
    for (int t = tid; t < Kv; t += nthr) {
        const unsigned long long key = s.key[t];
        const uint32_t lo32 = (uint32_t)key;
        const uint32_t cid = 0xffffu - (lo32 & 0xffffu);
        const int c = (int)(lo32 >> 16);
        const int bk = score_bucket(order_key_to_float((uint32_t)(key >> 32)), p.B);
        const int base = s.start[c];
        const int st = base + s.cntb[c * p.B + bk];
        const int en = base + ((bk == p.B - 1) ? s.cnt[c] : s.cntb[c * p.B + bk + 1]);
        int rank = 0;
#pragma unroll 4
        for (int u = st; u < en; ++u) rank += (s.key[u] > key) ? 1 : 0;
        const float ta = make_ta(s.box[cid], p.iou);
        if (ta != ta) s.flag[c] = 1;  // degenerate box: the class runs the exact pair arithmetic
        s.sord[st + rank] = make_uint2(s.box_saddr + 16u * cid, __float_as_uint(ta));
    }
}

// ---------------------------------------------------------------------------
// P5: pair masks (row-major words: bit k of word (rt, ct)[lane] = row 32rt+lane suppresses
// column 32ct+k) and the sweep
// ---------------------------------------------------------------------------
constexpr float kPairEps = 1e-5f;

// exact torchvision decisions of one row against the columns of one tile (slow path)
__device__ __noinline__ uint32_t block_exact(const uint2 *ordc, int ncol, const float4 R, double thr) {
    const float ra = box_area(R);
    uint32_t word = 0u;
    for (int k = 0; k < ncol; ++k) {
        const float4 Cb = lds_f4(ordc[k].x);
        if (nms_suppress_exact(R, ra, Cb, box_area(Cb), thr)) word |= 1u << k;
    }
    return word;
}

__device__ __forceinline__ void pair_step(const uint2 e, const float4 &R, float rta, uint32_t &bits, float &m) {
    const float4 Cb = lds_f4(e.x);
    const float cta = __uint_as_float(e.y);
    const float w = __fsub_rn(fminf(R.z, Cb.z), fmaxf(R.x, Cb.x));
    const float h = __fsub_rn(fminf(R.w, Cb.w), fmaxf(R.y, Cb.y));
    const float ws = __saturatef(__fmul_rn(w, 1.220703125e-4f));  // max(w, 0) * 2^-13 (exact; |coords| < 4096)
    const float sum = __fadd_rn(rta, cta);                         // t * (area_r + area_c) * 2^-13
    const float d = __fmaf_rn(-ws, h, sum);                        // < 0  <=>  inter > t * (area sum)
    m = fminf(m, __fmaf_rn(sum, -kPairEps, fabsf(d)));             // <= 0: too close to call
    bits = __funnelshift_l(__float_as_uint(d), bits, 1);           // sign bit -> mask bit
}

__device__ __forceinline__ uint32_t block_fast(const uint2 *ordc, int ncol, const float4 R, float rta, double thr) {
    uint32_t bits = 0u;
    float m = INFINITY;
    // chunks of 8 columns, then one chunk of 4 or 8 for the remainder; a partial last chunk reads up to 7 entries
    // past the class (the next class's entries, or the NaN-area padding after the last one): their bits are
    // dropped below, a NaN never lowers m, and a false "too close" only costs the exact path
    const int n8 = ncol & ~7, rem = ncol - n8;
    if (ncol == 32) {  // a full tile: one straight-line block
#pragma unroll
        for (int k = 0; k < 32; ++k) pair_step(ordc[k], R, rta, bits, m);
    } else {
#pragma unroll 1
        for (int k0 = 0; k0 < n8; k0 += 8) {
#pragma unroll
            for (int k = 0; k < 8; ++k) pair_step(ordc[k0 + k], R, rta, bits, m);
        }
    }
    int nr = n8;
    if (rem > 4) {
#pragma unroll
        for (int k = 0; k < 8; ++k) pair_step(ordc[n8 + k], R, rta, bits, m);
        nr += 8;
    } else if (rem > 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) pair_step(ordc[n8 + k], R, rta, bits, m);
        nr += 4;
    }
    uint32_t word = (__brev(bits) >> (32 - nr)) & (0xffffffffu >> (32 - ncol));  // column k was shifted in k-th
    if (m <= 0.0f) word = block_exact(ordc, ncol, R, thr);
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
        const uint2 *ord = s.sord + s.start[c];
        const int row = 32 * rt + lane;
        const uint2 re = ord[min(row, n - 1)];
        const float4 R = lds_f4(re.x);
        const float rta = __uint_as_float(re.y);
        const bool slow = s.flag[c] != 0;
        uint32_t *mb = s.mask + s.maskbase[c] + (rt - tri0) * 32 + lane;
        for (int ct = max(rt, t0); ct < te; ++ct) {
            const int ncol = min(32, n - 32 * ct);
            uint32_t word = slow ? block_exact(ord + 32 * ct, ncol, R, p.iou.thr) : block_fast(ord + 32 * ct, ncol, R, rta, p.iou.thr);
            if (ct == rt) word &= ~((2u << lane) - 1u);  // only LATER columns (2u << 31 == 0: none)
            if (row >= n) word = 0u;
            mb[tri(ct) * 32] = word;
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
// P6: output.  Tile g of the kept bitmap holds up to 32 rows that are consecutive in the
// output (class-ascending, score-descending); the warp assembles them in its scratch and
// writes them with coalesced stores.
// ---------------------------------------------------------------------------
template <int MODE, int THREADS, int GATHER>
__device__ __forceinline__ void phase_output(const DNParams &p, const Smem &s, int b) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = THREADS / 32;
    const int ntiles = s.ktile[p.C];
    const int nbuf = (GATHER != 0) ? p.gR : 1;   // fused all-gather: one copy per rank
    const size_t img = (GATHER != 0) ? (size_t)(p.gslot + b) : (size_t)b;
    float *o = (GATHER != 0) ? nullptr : p.out + img * p.K * 7;
    float *scr = s.scratch + warp * (32 * 7);
    // MODE_NMS: the caller's own rows are gathered bit-for-bit (pred_this_cls[index], box.py:29)
    const int K0 = (MODE == MODE_NMS) ? min(p.cand_count[0][b], p.cand_stride[0]) : 0;
    const float *r0 = (MODE == MODE_NMS) ? p.cand[0] + (size_t)b * p.cand_stride[0] * 7 : nullptr;
    const float *r1 = (MODE == MODE_NMS && p.cand[1]) ? p.cand[1] + (size_t)b * p.cand_stride[1] * 7 : nullptr;
    // each warp owns a contiguous range of tiles: one prefix over the kept bitmap per warp
    const int per = (ntiles + kWarps) / kWarps;  // ceil((ntiles + 1) / kWarps): "tile" ntiles writes the count
    const int g0 = warp * per, g1 = min(g0 + per, ntiles + 1);
    int before = 0;
    for (int t = lane; t < g0; t += 32) before += __popc(s.keptbits[t]);
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) before += __shfl_xor_sync(kFullMask, before, sh);
    for (int g = g0; g < g1; ++g) {
        if (g == ntiles) {
            if (GATHER != 0) {
                if (lane == 0) {
#pragma unroll
                    for (int r = 0; r < kMaxPeers; ++r)
                        if (r < nbuf) p.gcount[r][img] = before;
                }
            } else if (lane == 0) {
                p.out_count[b] = before;
            }
            break;
        }
        const uint32_t word = s.keptbits[g];
        const int nk = __popc(word);
        if (nk == 0) continue;
        const int c = s.tilecls[g];
        const int pos = s.start[c] + 32 * (g - s.ktile[c]) + lane;
        if ((word >> lane) & 1u) {
            const int r = __popc(word & lanemask_lt());
            const uint32_t cid = (s.sord[pos].x - s.box_saddr) >> 4;
            float *d = scr + 7 * r;
            if (MODE == MODE_NMS) {
                const float *src = ((int)cid < K0) ? r0 + (size_t)cid * 7 : r1 + (size_t)((int)cid - K0) * 7;
#pragma unroll
                for (int k = 0; k < 7; ++k) d[k] = __ldg(src + k);
            } else {
                const float4 bx = s.box[cid];
                const float2 cs = s.cs[cid];
                d[0] = bx.x; d[1] = bx.y; d[2] = bx.z; d[3] = bx.w;
                d[4] = cs.x; d[5] = cs.y;
                d[6] = (float)c;  // cls_idx.float() (yolo_loss.py:199)
            }
            if (p.out_idx) p.out_idx[(size_t)b * p.K + before + r] = (int)cid;
        }
        __syncwarp();
        const int nf = 7 * nk;
        if (GATHER != 0) {
            // NVLink likes long requests: the chunk (<= 224 floats, contiguous in every buffer, buffers 256-byte aligned)
            // goes out as <= 2 sixteen-byte stores per lane plus one scalar store for the unaligned head and tail
            const size_t e0 = (img * p.K + before) * 7;          // float index of the chunk in a buffer
            const int hcnt = min((int)((4 - (e0 & 3)) & 3), nf);
            const int nvec = (nf - hcnt) >> 2;
            const int tail0 = hcnt + 4 * nvec;
            float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
            if (lane < nvec) {
                const float *q = scr + hcnt + 4 * lane;
                a0 = make_float4(q[0], q[1], q[2], q[3]);
            }
            if (lane + 32 < nvec) {
                const float *q = scr + hcnt + 4 * (lane + 32);
                a1 = make_float4(q[0], q[1], q[2], q[3]);
            }
            int hidx = -1;                                       // lanes 0-2: head floats, lanes 4-6: tail floats
            if (lane < hcnt) hidx = lane;
            else if (lane >= 4 && lane < 8 && tail0 + lane - 4 < nf) hidx = tail0 + lane - 4;
            const float hv = (hidx >= 0) ? scr[hidx] : 0.f;
#pragma unroll
            for (int r = 0; r < kMaxPeers; ++r) {  // peer stores travel over NVLink while the next tile is assembled
                if (r >= nbuf) break;
                float *dst = p.gout[r] + e0;
                if (lane < nvec) *reinterpret_cast<float4 *>(dst + hcnt + 4 * lane) = a0;
                if (lane + 32 < nvec) *reinterpret_cast<float4 *>(dst + hcnt + 4 * (lane + 32)) = a1;
                if (hidx >= 0) dst[hidx] = hv;
            }
        } else {
            float *dst = o + (size_t)7 * before;
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                const int f = 32 * k + lane;
                if (f < nf) dst[f] = scr[f];
            }
        }
        __syncwarp();
        before += nk;
    }
}

// ---------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------
// GATHER: the fused all-gather variant of the output phase (its own instantiation, so the ordinary kernel's register
// allocation is untouched)
template <int MODE, int THREADS, int SHAPE, int GATHER = 0>
__global__ void __launch_bounds__(THREADS, (THREADS == 512) ? 2 : 1) decode_nms_kernel(const DNParams p, const SmemLayout L) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem s = carve(smem_raw, L, p.K, p.C, MODE);
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = p.C, K = p.K;
    using SH = ShapeT<SHAPE>;

    pdl_trigger();  // the next launch may start filling free SM slots right away (it waits before it writes)
    if (MODE == MODE_NMS || p.dbg) pdl_wait();  // (inputs written by our own decode kernels; debug stamps are global stores)
    stamp(p, b, 0);
    if (MODE != MODE_NMS) {
        // Start the HBM -> L2 stream of the FIRST head now, so that the first decode round (which can
        // only issue after the launch ramp) finds its lines on the way (-1 us).  Prefetching more --
        // the whole image up front, or one decode round ahead -- is slower by 1-2 us (measured,
        // profiles/r01/NOTES.md): the demand loads then queue behind the prefetches.
        if (tid == 0 && !(p.flags & 1)) {
            const size_t img0 = (size_t)p.A * p.attrs * p.head[0].HW;
            l2_prefetch_span(p.head[0].ptr + (size_t)b * img0, img0 * sizeof(float));
            if (p.nheads > 1 && (p.flags & 8)) {
                const size_t img1 = (size_t)p.A * p.attrs * p.head[1].HW;
                l2_prefetch_span(p.head[1].ptr + (size_t)b * img1, img1 * sizeof(float));
            }
        }
    }
    constexpr bool kNhwcShape = (MODE == MODE_FUSED && ShapeIsNhwc<SHAPE>::value);
    constexpr bool kStaticShape = (MODE != MODE_NMS && SH::C > 0 && !ShapeIsNhwc<SHAPE>::value);
    if (MODE != MODE_DECODE) {
        for (int i = tid; i <= C * p.B; i += THREADS) s.cntb[i] = 0;
        for (int i = tid; i <= C; i += THREADS) s.flag[i] = 0;
        if (!kStaticShape) __syncthreads();  // (compile-time shapes: the barrier sits behind the first round's loads)
    }
    stamp(p, b, 15);

    if constexpr (MODE == MODE_NMS) {
        phase_load_rows<THREADS>(p, s, b);
    } else if constexpr (kStaticShape && MODE == MODE_DECODE) {
        decode_head_static<THREADS, MODE, SH::C, SH::HW0, SH::W0, false>(p, s, b, p.head[0], 0, 0);
    } else if constexpr (kStaticShape) {
        decode_head_static<THREADS, MODE, SH::C, SH::HW0, SH::W0, true>(p, s, b, p.head[0], 0, 0);
        decode_head_static<THREADS, MODE, SH::C, SH::HW1, SH::W1, false>(p, s, b, p.head[1], p.head[0].cells, 1);
    } else if constexpr (kNhwcShape) {
        // scratch: the part of U that is free during the decode (behind clsidx); the host checked that it fits
        float *scr = reinterpret_cast<float *>(smem_raw + L.U + 4 * align_up((uint32_t)(K > 0 ? K : 1), 32));
        decode_head_nhwc<THREADS, SH::C>(p, s, b, p.head[0], 0, scr);
        decode_head_nhwc<THREADS, SH::C>(p, s, b, p.head[1], p.head[0].cells, scr);
    } else {
        if (MODE == MODE_FUSED && p.nhwc) {
            float *scr = reinterpret_cast<float *>(smem_raw + L.U + 4 * align_up((uint32_t)(K > 0 ? K : 1), 32));
            decode_head_nhwc<THREADS, 0>(p, s, b, p.head[0], 0, scr);
            decode_head_nhwc<THREADS, 0>(p, s, b, p.head[1], p.head[0].cells, scr);
        } else {
            decode_head_rt<THREADS, MODE>(p, s, b, p.head[0], 0, 0);
            if (MODE == MODE_FUSED) decode_head_rt<THREADS, MODE>(p, s, b, p.head[1], p.head[0].cells, 1);
        }
    }
    __syncthreads();
    stamp(p, b, 1);

    if (MODE == MODE_DECODE) {
        // YOLOLoss.get_pred_boxes output: rows in candidate order (:203)
        const int nwords = (K + 31) >> 5;
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
        pdl_wait();
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

    // P2: bucket positions inside each class (warp per class), then the class starts (warp 0)
    scan_buckets_per_class<THREADS>(p, s);
    __syncthreads();
    stamp(p, b, 12);
    if (warp == 0) warp_class_starts(p, s);
    __syncthreads();
    stamp(p, b, 13);
    const int Kv = s.misc[M_KV];
    if (warp == 0) {
        // kept-bitmap tiles, round table, strip tasks: only the pair phase needs them, so warp 0 builds
        // them while the other warps sort
        warp_class_scan(p, s, (int)L.mask_words);
        warp_round_prefix(s, 0);
    } else {
        // P3 / P4 on the other warps (named barrier 1 between the key scatter and the ranking)
        phase_scatter_keys<THREADS>(p, s, tid - 32, THREADS - 32);
        asm volatile("bar.sync 1, %0;" ::"n"(THREADS - 32) : "memory");
        phase_rank_sort<THREADS>(p, s, Kv, tid - 32, THREADS - 32);
    }
    __syncthreads();
    stamp(p, b, 3);
    const int nrounds = s.misc[M_NROUNDS];
    // P5
    for (int r = 0;;) {
        phase_pairs_sweep(p, s);
        __syncthreads();
        if (++r >= nrounds) break;
        if (warp == 0) warp_round_prefix(s, r);
        __syncthreads();
    }
    stamp(p, b, 4);
    // P6 (the mask buffer is dead now; the row scratch aliases it)
    pdl_wait();
    phase_output<MODE, THREADS, GATHER>(p, s, b);
    stamp(p, b, 7);
}

}  // namespace b200yolo
