// decode_nms.cuh -- YOLO-head decode + confidence threshold + per-class NMS (sm_100a).
//
// One CTA per image; everything about one image lives in shared memory, the head
// tensors are read from HBM exactly once and only kept rows are written back.
//
// Shared memory is kept SMALL on purpose (<= 80 KB per CTA for the 352x352 VOC heads):
// the decode phase streams the heads with ordinary coalesced loads whose in-flight
// lines live in L1, and L1 is what the CTAs' shared memory leaves of the SM's 228 KB
// (profiles/micro/load_pattern.cu: the same load pattern runs at 4.3 TB/s with 64 KB of
// L1 and at 2.9 TB/s with none).  So: records by cell id (box 16 B, conf/score 8 B) and
// one region `U` that is reused by phase.
//
// The kernel is issue-bound once the heads are in (profiles/r02/NOTES.md: round 1 executed 86.7 k
// warp instructions per image, 44 % of them the fp32 box-pair loop), so every phase is
// written for instruction count:
//
//   P1 decode     thread per cell, ALL 5+C attribute planes of the cell loaded at
//                 once (coalesced: consecutive lanes = consecutive cells of a plane).
//                 The reference's shapes (VOC 352 / 416, BDD 640x384) are compiled with
//                 the plane stride as a constant, so the 5+C loads are one base register
//                 plus immediates; any other shape takes the runtime-stride path.
//                 conf = sigmoid(tc) > thr (yolo_loss.py:189,201); for passing cells:
//                 class max / argmax (:198), box (:186-196,243-247) -> box[], cs[],
//                 and the per-(class, score bucket) arrival index (one shared atomic).
//                 A warp whose first-head cells mostly fail (trained heads) reads the second head's
//                 objectness plane first and only the passing cells' other planes
//                 (decode_head_static_sparse): same results, fewer round trips and bytes.
//   P2 scans      warp per class: exclusive scan of the class's score-bucket histogram; warp 0:
//                 class segment starts (box.py:20-22).  Nothing else sits in a window that one warp
//                 executes while fifteen wait.
//   P3 key scatter 64-bit keys (score desc, candidate order asc == the stable sort of
//                 torchvision.ops.nms) into their (class, bucket) segment; warp 0 builds the tile
//                 tables (a class owns ceil(n/32) TILES of 32 sorted positions; named barrier 2 orders
//                 them before the ranking) and the pair-task list meanwhile.
//   P4 rank       rank inside the bucket = sorted position.  Every candidate is then written to the
//                 tile-padded sorted tables: its cell id (scid) and a CONSERVATIVE fp16 image of
//                 its box (H16Tile: x1/y1 rounded down, x2/y2 rounded up, a lower bound of
//                 t*area; two columns per 32-bit word).
//   P5 pairs + sweep as you go.  Warp tasks claimed from a queue in dependency order; a task is
//                 one 32x32 block (rows of tile rt, columns of tile ct) of one class.  The block
//                 is first PREFILTERED in packed fp16 (HMNMX2 / HADD2 / HFMA2, two columns per
//                 instruction: 8.4 issue cycles per row x column slot against 17.0 for the fp32
//                 loop, profiles/micro/pair_loop.cu).  The rounding directions and the area
//                 margin make "inter < t*(Sa+Sb)" in fp16 a PROOF that torchvision does not
//                 suppress the pair (h16_prefilter below); the few pairs it cannot rule out
//                 ("maybe") are decided in fp32 with torchvision's exact arithmetic behind a
//                 guard band (pair_decide).  An off-diagonal block reduces at once to ONE word
//                 -- the columns suppressed by the KEPT rows of tile rt -- that is OR-ed into the
//                 column tile's word; the diagonal task of a tile runs when all earlier row
//                 tiles have contributed and resolves the tile in score order.  No n^2 mask in
//                 memory, no separate sweep, rows that are already suppressed cost nothing.
//   P6 output     class-ascending / score-descending rows (box.py:29-30).  A warp owns
//                 tiles of the kept bitmap: the rows of a tile are consecutive
//                 in the output, so the warp assembles them in a 896-byte scratch and
//                 stores them with coalesced 4-byte stores.
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
// starts on the SM slots its predecessor leaves free and streams its heads under the predecessor's NMS.  Whether it
// must wait for that predecessor before it stores depends on the output buffers of the list (DNParams::chain).
#pragma once
#include <cuda_fp16.h>

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
    int Bshift;          // log2(B)
    int wait_inputs;     // 1: griddepcontrol.wait before the first global read (see launch_dn_t)
    // Launches of a list whose consecutive batches write disjoint outputs (b200yolo_decode_nms_batches) need not wait for
    // their predecessor before they store.  What keeps the list ordered is a chain: 1 = first launch of such a list: every
    // CTA waits for whatever precedes the list and only then lets the next launch start; 2 = a later launch: its CTAs
    // never wait, and one extra CTA (blockIdx.x == N) waits for the previous launch to complete before IT lets the next
    // launch start -- so launch k + 1 starts after launch k - 1 has completed (two launches in flight, outputs in a
    // ring of two are safe) and every launch completes after its predecessor.  3 = a later launch of a list in which NO
    // two batches share an output: nothing orders the stores, the extra CTA only waits (a launch still completes after
    // its predecessor), and launch k + 1 starts as soon as every CTA of launch k has.  0 = single call: wait before storing.
    int chain;
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
    int gnbuf;                   // buffers the output phase stores into: gR peer mappings, or 1 NVSwitch multicast view
    int gslot;                   // first image slot of this rank = rank * N
    float *gout[kMaxPeers];      // [gR] rank r's buffer [gR*N][K][7]
    int *gcount[kMaxPeers];      // [gR] rank r's counts [gR*N]
    // back-pressure of the fused all-gather: before its first store the kernel waits until every rank's arrival flag
    // (this rank's own array, written by the peers over NVLink) has reached gwait_value (0: no wait)
    const int *gwait_flags;
    int gwait_value;
    int *gtimed_out;
    long long gwait_cycles;
    // arrival signal of the fused all-gather, riding on the NEXT step's launch: when gsignal > 0 the grid has one extra
    // CTA that waits for the previous launch in the stream (the previous step) to complete and then raises slot grank
    // of every rank's flag array to gsignal
    int *gflags[kMaxPeers];
    int grank, gsignal;
    // MODE_NMS inputs
    const float *cand[2];
    const int *cand_count[2];
    int cand_stride[2];
};


// fp16 column table of one tile (32 sorted positions of one class): entry k packs columns k (low half) and k + 16
// (high half), so one HMNMX2 / HADD2 / HFMA2 works on two columns.  Written in P4, read in P5.
struct H16Tile {
    uint4 xy[16];      // .x = X1 pair, .y = Y1 pair, .z = X2 pair, .w = Y2 pair (half2 each); X = x/16, Y = 1024*y
    uint32_t ta[16];   // TA pair: a lower bound of 64 * t * area, or -inf ("always maybe": degenerate box)
};
static_assert(sizeof(H16Tile) == 320, "H16Tile layout");
constexpr uint32_t kTileBytes = sizeof(H16Tile) + 64;   // + scid u16[32]

struct SmemLayout {
    uint32_t box, cs, cls, tiles, tasks, misc, U, total;
    uint32_t passbits, tilepref;   // MODE_DECODE
    uint32_t u_bytes;     // bytes in U
    uint32_t key_off;     // keys (8 B per cell), also the channels-last staging scratch during the decode
    uint32_t cntb_off;    // score-bucket counters (live from the decode to the ranking)
    uint32_t h16_off;     // fp16 tile table (after the ranking); the P6 row scratch aliases it
    uint32_t nt_max;      // tiles an image can need: Kp/32 + C
    uint32_t task_cap;    // entries of the pair-task table
};

__host__ __device__ inline uint32_t align_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

// per-class int arrays, each (C+1) long
enum { CA_CNT = 0, CA_START, CA_KTILE, CA_NUM };
// misc ints
enum { M_NTASK = 0, M_CTR, M_NLEVELS, M_TOTAL, M_KV, M_NUM = 16 };

// score buckets per class: C*B counters, at most 1536 (6 KB)
__host__ __device__ inline int pick_buckets(int C) {
    int B = 64;
    while (B > 1 && C * B > 1536) B >>= 1;
    return B;
}

constexpr uint32_t kScratchFloats = 320;              // P6: per-warp staging: one tile of 32 x 7 floats + what the previous flush left (< 96)
constexpr uint32_t kScratchPerWarp = kScratchFloats * 4;
constexpr int kRankItems = 8;                     // P4: sorted positions a thread carries over the barrier

// U region by phase (Kp = K rounded up to 32, NT = Kp/32 + C tiles):
//   decode .. rank   clsidx u32[Kp] | key u64[Kp] | cntb int[C*B+1]
//   pairs            scid u16[32*NT] | H16Tile[NT]
//   output           scid ........... | per-warp row scratch
//   (MODE_DECODE)    clsidx u32[Kp] | outsrc u16[Kp]
__host__ __device__ inline SmemLayout make_layout(int K, int C, int mode, int threads, uint32_t extra) {
    SmemLayout L;
    const uint32_t Kp = align_up((uint32_t)(K > 0 ? K : 1), 32);
    const uint32_t Cp = align_up((uint32_t)C + 1, 4);
    const uint32_t NT = Kp / 32 + (uint32_t)C;
    const bool nms = (mode != MODE_DECODE);
    uint32_t o = 0;
    L.box = o; o += 16 * Kp;
    L.cs = o; o += 8 * Kp;
    L.cls = o; o += nms ? 4 * Cp * CA_NUM : 0;
    L.tiles = o; o += nms ? 16 * align_up(NT + 1, 4) : 0;   // keptw, supw, arrived, (ready | class) per tile
    L.tilepref = o; o += nms ? 0 : 4 * (NT + 2);
    L.passbits = o; o += nms ? 0 : 4 * (Kp / 32);
    L.misc = o; o += 4 * M_NUM;
    (void)extra;
    L.task_cap = nms ? align_up(2 * NT + 1, 4) : 0;   // a head task and at most one tail task per tile
    L.tasks = o; o += 4 * L.task_cap;
    o = align_up(o, 16);
    L.U = o;
    const uint32_t sort_bytes = 4 * Kp + 8 * Kp + 4 * align_up((uint32_t)(C * pick_buckets(C)) + 1, 4);
    const uint32_t pair_bytes = 64 * NT + (uint32_t)sizeof(H16Tile) * NT;
    const uint32_t out_bytes = 64 * NT + (uint32_t)(threads / 32) * kScratchPerWarp;
    uint32_t u = sort_bytes > pair_bytes ? sort_bytes : pair_bytes;
    if (out_bytes > u) u = out_bytes;
    L.u_bytes = nms ? align_up(u, 16) : 6 * Kp;
    L.key_off = L.U + 4 * Kp;
    L.cntb_off = L.U + 12 * Kp;
    L.h16_off = L.U + 64 * NT;
    L.nt_max = NT;
    L.total = align_up(L.U + L.u_bytes, 16);
    return L;
}

struct Smem {
    float4 *box;        // [Kp] x1 y1 x2 y2 by cell id            (output columns 0-3)
    float2 *cs;         // [Kp] conf, class score                 (columns 4, 5)
    uint32_t *clsidx;   // [Kp] (global bucket c*B+bk << 16) | arrival index inside the bucket; ~0u: not a candidate
    unsigned long long *key;
    int *cntb;          // [C*B+1] per (class, score bucket): arrival counter, then exclusive prefix inside the class
    uint16_t *scid;     // [32*NT] tile-padded sorted position -> cell id
    H16Tile *h16;       // [NT]
    float *scratch;     // P6: per-warp 32 x 7 floats
    uint16_t *outsrc;   // MODE_DECODE: output row -> cell id
    uint32_t *passbits, *tilepref;
    uint32_t *keptw;    // [NT] kept bitmap of the tile (final once `ready`)
    uint32_t *supw;     // [NT] columns of the tile suppressed by kept rows of EARLIER tiles (OR-accumulated)
    int *arrived;       // [NT] earlier row tiles that have contributed to supw
    int *ready;         // [NT] class of the tile << 16 | 1 once keptw is final (set for tiles that have a tail task)
    uint32_t *tasks;    // [task_cap] class << 16 | kTaskTail? | tile (phase_pairs)
    int *cnt, *start, *ktile;
    int *misc;
};

__device__ __forceinline__ Smem carve(unsigned char *base, const SmemLayout &L, int K, int C) {
    Smem s;
    const uint32_t Kp = align_up((uint32_t)(K > 0 ? K : 1), 32);
    const uint32_t Cp = align_up((uint32_t)C + 1, 4);
    const uint32_t NTp = align_up(L.nt_max + 1, 4);
    s.box = reinterpret_cast<float4 *>(base + L.box);
    s.cs = reinterpret_cast<float2 *>(base + L.cs);
    s.clsidx = reinterpret_cast<uint32_t *>(base + L.U);
    s.key = reinterpret_cast<unsigned long long *>(base + L.key_off);
    s.cntb = reinterpret_cast<int *>(base + L.cntb_off);
    s.scid = reinterpret_cast<uint16_t *>(base + L.U);
    s.h16 = reinterpret_cast<H16Tile *>(base + L.h16_off);
    s.scratch = reinterpret_cast<float *>(base + L.h16_off);
    s.outsrc = reinterpret_cast<uint16_t *>(base + L.U + 4 * Kp);
    s.passbits = reinterpret_cast<uint32_t *>(base + L.passbits);
    s.tilepref = reinterpret_cast<uint32_t *>(base + L.tilepref);
    s.keptw = reinterpret_cast<uint32_t *>(base + L.tiles);
    s.supw = s.keptw + NTp;
    s.arrived = reinterpret_cast<int *>(s.supw + NTp);
    s.ready = s.arrived + NTp;
    s.tasks = reinterpret_cast<uint32_t *>(base + L.tasks);
    int *ca = reinterpret_cast<int *>(base + L.cls);
    s.cnt = ca + CA_CNT * Cp;
    s.start = ca + CA_START * Cp;
    s.ktile = ca + CA_KTILE * Cp;
    s.misc = reinterpret_cast<int *>(base + L.misc);
    return s;
}

__device__ __forceinline__ int fastdiv(int n, uint32_t magic) {
    return magic ? (int)__umulhi((uint32_t)n, magic) : n;  // exact for n < 65536
}

__device__ __forceinline__ int tri(int x) { return (x * (x + 1)) >> 1; }

// profiling aid (DBG instantiations only; the production kernels carry no trace of it): slot k of image b gets
// %globaltimer (ns, comparable across SMs) at the CTA's two ends and, 16 slots further, the SM's cycle counter
template <bool DBG>
__device__ __forceinline__ void stamp(const DNParams &p, int b, int k) {
    if constexpr (DBG) {
        if (p.dbg && threadIdx.x == 0) {
            if (k == 0 || k == 7) {  // (reading %globaltimer costs several hundred ns: only at the CTA's two ends)
                unsigned long long t;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                p.dbg[(size_t)b * 32 + k] = t;
            }
            p.dbg[(size_t)b * 32 + 16 + k] = (unsigned long long)clock64();
        }
    }
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
// error: evaluate like the reference (IEEE sigmoid first, then first max, yolo_loss.py:198)
__device__ __noinline__ float class_tie_break(const float *qc, int HW, int C, float lo, float m1, int i1, int *bi_out) {
    float best = -1.0f;
    int bi = 0;
    for (int cc = 0; cc < C; ++cc) {
        const float x = __ldg(qc + (size_t)cc * HW);
        if (!(x < lo)) {
            const float sg = sigmoid_f(x);   // (the IEEE form torch.sigmoid computes: the tie is decided like the reference's)
            if (sg > best) { best = sg; bi = cc; }
        }
    }
    if (best < 0.0f) { best = sigmoid_f(m1); bi = i1; }  // only NaN logits in the window
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

// conf = sigmoid(tc) against the threshold (yolo_loss.py:189,201).  The SFU sigmoid is within ~4 ulp of the IEEE one, so
// a cell whose confidence lands within 2e-6 (relative) of the threshold is re-evaluated with the IEEE form
// 1/(1+expf(-x)) -- what torch.sigmoid computes -- and that value decides and is reported: the candidate SET then
// matches the reference's own arithmetic.  EXACT (b200yolo_set_exact_decode): every cell takes the IEEE form.
constexpr float kConfBand = 2e-6f;

__device__ __noinline__ float sigmoid_ieee_call(float x) { return sigmoid_f(x); }   // (out of line: the hot path keeps a branch only)

template <bool EXACT>
__device__ __forceinline__ bool conf_pass(float tc, float thr, float *conf_out) {
    float conf = EXACT ? sigmoid_f(tc) : sigmoid_fast(tc);
    if (!EXACT && fabsf(__fsub_rn(conf, thr)) <= __fmul_rn(kConfBand, fabsf(thr))) conf = sigmoid_ieee_call(tc);
    *conf_out = conf;
    return conf > thr;
}

// ---------------------------------------------------------------------------
// P1: decode every cell of one head, single pass.  All 5+C plane loads of a cell are
// issued before the first use.
// ---------------------------------------------------------------------------
// box arithmetic + records of one passing cell (yolo_loss.py:186-199, 243-247)
template <int MODE, bool EXACT>
__device__ __forceinline__ void emit_candidate(const DNParams &p, const Smem &s, const HeadDesc &hd, int cid, int a, int i, int j,
                                               float tx, float ty, float tw, float th, float conf, float best, int bi) {
    // EXACT (b200yolo_set_exact_decode; its own instantiations, the default kernels carry none of it): the reference's
    // own operations -- IEEE sigmoid / expf and a true division by the grid size -- instead of the SFU forms and the
    // multiplication by 1/W: rows bit-identical to the reference on CUDA (profiles/exact_decode.py)
    const float sx = EXACT ? sigmoid_f(tx) : sigmoid_fast(tx), sy = EXACT ? sigmoid_f(ty) : sigmoid_fast(ty);    // :187
    const float ew = EXACT ? expf(tw) : exp_fast(tw), eh = EXACT ? expf(th) : exp_fast(th);                      // :188
    const float cx = EXACT ? __fdiv_rn(__fadd_rn(sx, (float)i), hd.fW) : __fmul_rn(__fadd_rn(sx, (float)i), hd.rW);  // :194
    const float cy = EXACT ? __fdiv_rn(__fadd_rn(sy, (float)j), hd.fH) : __fmul_rn(__fadd_rn(sy, (float)j), hd.rH);
    const float bw = __fmul_rn(ew, hd.aw[a]);                    // :195
    const float bh = __fmul_rn(eh, hd.ah[a]);
    float4 bx;
    bx.x = __fsub_rn(cx, __fmul_rn(bw, 0.5f));                   // :244
    bx.y = __fsub_rn(cy, __fmul_rn(bh, 0.5f));                   // :245
    bx.z = __fadd_rn(bw, bx.x);                                  // :246
    bx.w = __fadd_rn(bh, bx.y);                                  // :247
    s.box[cid] = bx;
    s.cs[cid] = make_float2(conf, best);
    if (MODE == MODE_FUSED) {
        const uint32_t gb = (uint32_t)(bi * p.B + score_bucket(__fmul_rn(best, conf), p.B));  // global score bucket
        s.clsidx[cid] = (gb << 16) | (uint32_t)atomicAdd(&s.cntb[gb], 1);
    } else {
        s.clsidx[cid] = (uint32_t)bi << 16;
    }
}

// Compile-time class count and grid: the 5+C loads are one base register plus immediates.  FIRST: this
// is the first decode of the kernel -- the block barrier that orders the zeroing of the histogram before
// the first shared atomic is taken AFTER the first round's loads are in flight (hides ~0.5 us of start-up
// behind the first HBM round trip).
template <int THREADS, int MODE, int CT, int HWT, int WT, bool FIRST, bool DBG, bool EXACT>
__device__ __forceinline__ void decode_head_static(const DNParams &p, const Smem &s, int b, const HeadDesc &hd, int cid0, int hh,
                                                   int *warp_active = nullptr, int *warp_pass = nullptr) {
    static_assert(CT >= 1 && CT <= 24, "compile-time shapes keep all class bits in one fp32 accumulator");
    const int tid = threadIdx.x, lane = tid & 31;
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
            float conf;
            pass = conf_pass<EXACT>(tc, p.conf_thr, &conf);   // yolo_loss.py:189,197,201 (threshold already rounded to fp32)
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
                else if (EXACT) best = sigmoid_f(m1);
                const int j = pos / WT;
                emit_candidate<MODE, EXACT>(p, s, hd, cid, a, pos - j * WT, j, tx, ty, tw, th, conf, best, bi);
            } else if (MODE == MODE_FUSED) {
                s.clsidx[cid] = 0xffffffffu;
            }
        }
        if (MODE == MODE_DECODE) {  // single head: candidate ids are 32-aligned per warp
            const unsigned bal = __ballot_sync(kFullMask, pass);
            if (lane == 0 && active) s.passbits[local >> 5] = bal;
        }
        if (FIRST && MODE == MODE_FUSED && base == 0 && warp_active != nullptr) {   // (decode_head_static_sparse)
            *warp_active = __popc(__ballot_sync(kFullMask, active));
            *warp_pass = __popc(__ballot_sync(kFullMask, pass));
        }
        if (MODE == MODE_FUSED) stamp<DBG>(p, b, 8 + min(3, hh + base / THREADS));
    }
}

// The second head of a compile-time shape when the image looks sparse to this warp (trained heads pass a few percent of
// the cells).  The ordinary decode reads all 5 + C planes of every cell -- one HBM round trip per 512 cells, 25 loads a
// cell -- although only the objectness plane decides whether the other 24 are ever looked at.  Here the warp reads the
// objectness of ALL its cells of the head first (one load per cell, one round trip for the whole head), and only the
// lanes that got a passing cell assigned read that cell's other planes (a second round trip): three round trips
// instead of four per image and a fraction of the bytes on trained-like heads.  Decided per warp, no barrier: a warp
// whose first-head cells mostly fail tries it; a warp that finds more than 32 passing cells falls back to the ordinary
// loop.  Only the loads differ -- every passing cell runs the same arithmetic, so the results are identical either way.
// Returns false: nothing done, decode the head with decode_head_static.
template <int THREADS, int MODE, int CT, int HWT, int WT, bool EXACT>
__device__ __forceinline__ bool decode_head_static_sparse(const DNParams &p, const Smem &s, int b, const HeadDesc &hd, int cid0) {
    constexpr int attrs = CT + 5;
    constexpr int kMaxRounds = 4;
    const int tid = threadIdx.x, lane = tid & 31;
    const int cells = p.A * HWT;
    const int rounds = (cells + THREADS - 1) / THREADS;
    if (rounds > kMaxRounds) return false;
    const float *hb = hd.ptr + (size_t)b * p.A * attrs * HWT;  // uniform
    float tc[kMaxRounds];
#pragma unroll
    for (int r = 0; r < kMaxRounds; ++r) {
        const int local = r * THREADS + tid;
        tc[r] = 0.f;
        if (r < rounds && local < cells) {
            const int a = local / HWT;
            tc[r] = __ldcs(hb + (uint32_t)(a * attrs * HWT + (local - a * HWT)) + 4 * HWT);
        }
    }
    uint32_t bal[kMaxRounds];
    int total = 0;
#pragma unroll
    for (int r = 0; r < kMaxRounds; ++r) {
        const int local = r * THREADS + tid;
        bool pass = false;
        if (r < rounds && local < cells) {
            float conf;
            pass = conf_pass<EXACT>(tc[r], p.conf_thr, &conf);
        }
        bal[r] = __ballot_sync(kFullMask, pass);
        total += __popc(bal[r]);
    }
    if (total > 32) return false;   // dense after all: the ordinary loop (it writes every record)
    // cells that fail (the ordinary loop marks them as it goes)
#pragma unroll
    for (int r = 0; r < kMaxRounds; ++r) {
        const int local = r * THREADS + tid;
        if (r < rounds && local < cells && !((bal[r] >> lane) & 1u)) s.clsidx[cid0 + local] = 0xffffffffu;
    }
    // lane L takes the L-th passing cell of the warp (round-major)
    int local = -1, off = 0;
#pragma unroll
    for (int r = 0; r < kMaxRounds; ++r) {
        const int n = __popc(bal[r]);
        if (lane >= off && lane < off + n) local = r * THREADS + (tid & ~31) + (int)__fns(bal[r], 0, lane - off + 1);
        off += n;
    }
    if (local >= 0) {
        const int a = local / HWT;
        const int pos = local - a * HWT;
        const float *q = hb + (uint32_t)(a * attrs * HWT + pos);
        const float tx = __ldcs(q), ty = __ldcs(q + HWT), tw = __ldcs(q + 2 * HWT), th = __ldcs(q + 3 * HWT), tcv = __ldcs(q + 4 * HWT);
        float x[CT];
#pragma unroll
        for (int u = 0; u < CT; ++u) x[u] = __ldcs(q + (5 + u) * HWT);
        float conf;
        const bool pass = conf_pass<EXACT>(tcv, p.conf_thr, &conf);   // (true: the same value decided above)
        if (pass) {
            float m1 = x[0];
#pragma unroll
            for (int u = 1; u < CT; ++u) m1 = fmaxf(m1, x[u]);
            float best;
            const float win = tie_window(m1, &best);
            const float lo = __fsub_rn(m1, win);
            float near = 0.f;
#pragma unroll
            for (int u = 0; u < CT; ++u) near = __fmaf_rn((x[u] >= lo) ? 1.0f : 0.0f, (float)(1u << u), near);
            const uint32_t nb = __float2uint_rn(near);
            int bi = nb ? __ffs(nb) - 1 : 0;
            const bool tie = (nb & (nb - 1u)) != 0u || nb == 0u;
            if (CT > 1 && tie) best = class_tie_break(q + 5 * HWT, HWT, CT, lo, m1, bi, &bi);
            else if (EXACT) best = sigmoid_f(m1);
            const int j = pos / WT;
            emit_candidate<MODE, EXACT>(p, s, hd, cid0 + local, a, pos - j * WT, j, tx, ty, tw, th, conf, best, bi);
        }
    }
    return true;
}

// Runtime class count and grid (any shape): plane stride in a register, classes in chunks of 24.
template <int THREADS, int MODE, bool DBG, bool EXACT>
__device__ __forceinline__ void decode_head_rt(const DNParams &p, const Smem &s, int b, const HeadDesc &hd, int cid0, int hh) {
    const int tid = threadIdx.x, lane = tid & 31;
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
                    pass = conf_pass<EXACT>(tc, p.conf_thr, &conf);   // yolo_loss.py:189,197,201
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
                else if (EXACT) best = sigmoid_f(m1);
                const int j = fastdiv(pos, hd.magicW);
                emit_candidate<MODE, EXACT>(p, s, hd, cid, a, pos - j * hd.W, j, tx, ty, tw, th, conf, best, bi);
            } else if (MODE == MODE_FUSED) {
                s.clsidx[cid] = 0xffffffffu;
            }
        }
        if (MODE == MODE_DECODE) {  // single head: candidate ids are 32-aligned per warp
            const unsigned bal = __ballot_sync(kFullMask, pass);
            if (lane == 0 && local < cells) s.passbits[local >> 5] = bal;
        }
        if (MODE == MODE_FUSED) stamp<DBG>(p, b, 8 + min(3, hh + base / THREADS));
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

template <int THREADS, int CT, bool EXACT>
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
            float conf;
            if (conf_pass<EXACT>(xs[4], p.conf_thr, &conf)) {   // yolo_loss.py:189,197,201
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
                    else if (EXACT) best = sigmoid_f(m1);
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
                    else if (EXACT) best = sigmoid_f(m1);
                }
                emit_candidate<MODE_FUSED, EXACT>(p, s, hd, cid, a, i, j, xs[0], xs[1], xs[2], xs[3], conf, best, bi);
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
                const uint32_t gb = (uint32_t)(c * p.B + score_bucket(__fmul_rn(score, conf), p.B));
                s.clsidx[row] = (gb << 16) | (uint32_t)atomicAdd(&s.cntb[gb], 1);
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
#pragma unroll 1
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
// P2 (warp 0): tile tables.  A class owns ceil(n/32) tiles of 32 sorted positions; tile g of the image is
// tile g - ktile[c] of class c = ready[g] >> 16.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void warp_tile_tables(const DNParams &p, const Smem &s) {
    const int lane = threadIdx.x & 31;
    const int C = p.C;
    int carryT = 0, tmax = 0;
#pragma unroll 1
    for (int c0 = 0; c0 < C; c0 += 32) {
        const int c = c0 + lane;
        const int n = (c < C) ? s.cnt[c] : 0;
        const int T = (n + 31) >> 5;
        const int incT = warp_inclusive_scan(T, lane);
        if (c < C) {
            const int kt = carryT + incT - T;
            s.ktile[c] = kt;
#pragma unroll 1
            for (int t = 0; t < T; ++t) {
                s.keptw[kt + t] = 0u;
                s.supw[kt + t] = 0u;
                s.arrived[kt + t] = 0;
                s.ready[kt + t] = c << 16;
            }
        }
        carryT += __shfl_sync(kFullMask, incT, 31);
        tmax = max(tmax, T);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tmax = max(tmax, __shfl_xor_sync(kFullMask, tmax, o));
    if (lane == 0) {
        s.ktile[C] = carryT;
        s.misc[M_NLEVELS] = tmax;   // tiles of the largest class
    }
    __syncwarp();
}

constexpr int kTailMin = 4;          // a tail task exists when tile rt has at least this many tiles beyond rt + 1
constexpr uint32_t kTaskTail = 0x8000u;

__device__ __forceinline__ bool has_tail(int T, int rt) { return T - rt - 2 >= kTailMin; }

// (warp 0) the task table, tile-major: head tasks of tile 0 of every class that has one, their tail tasks, head tasks
// of tile 1, ... (class << 16 | kTaskTail? | tile); at most 2 * NT entries
__device__ __forceinline__ void warp_build_tasks(const DNParams &p, const Smem &s) {
    const int lane = threadIdx.x & 31;
    const int C = p.C;
    const int tmax = s.misc[M_NLEVELS];
    int q = 0;
    for (int t = 0; t < tmax; ++t) {
        for (int pass = 0; pass < 2; ++pass) {
            for (int c0 = 0; c0 < C; c0 += 32) {
                const int c = c0 + lane;
                const int T = (c < C) ? (s.cnt[c] + 31) >> 5 : 0;
                const bool has = pass ? has_tail(T, t) : T > t;
                const uint32_t bal = __ballot_sync(kFullMask, has);
                if (has) s.tasks[q + __popc(bal & lanemask_lt())] = ((uint32_t)c << 16) | (pass ? kTaskTail : 0u) | (uint32_t)t;
                q += __popc(bal);
            }
        }
    }
    if (lane == 0) {
        s.misc[M_NTASK] = q;
        s.misc[M_CTR] = 0;
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------
// P3/P4: class-major counting sort on score buckets, then rank inside the bucket:
// the stable descending score sort of torchvision.ops.nms, per class
// key = score order key (32) | 0xffff - cell id (16) | global bucket c*B+bk (16)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void phase_scatter_keys(const DNParams &p, const Smem &s, int tid, int nthr) {
    for (int cid = tid; cid < p.K; cid += nthr) {
        const uint32_t ci = s.clsidx[cid];
        if (ci == 0xffffffffu) continue;
        const float2 cs = s.cs[cid];
        const float sc = __fmul_rn(cs.y, cs.x);  // box.py:27 scores = col5*col4
        const uint32_t gb = ci >> 16;
        const unsigned long long key =
            ((unsigned long long)float_order_key(sc) << 32) | (unsigned long long)(((0xffffu - (uint32_t)cid) << 16) | gb);
        s.key[s.start[gb >> p.Bshift] + s.cntb[gb] + (int)(ci & 0xffffu)] = key;
    }
}

// rank of every key inside its bucket -> tile-padded sorted position pp = 32 * tile + column, carried in registers
// (pp << 16 | cell id) over the barrier after which the keys are dead
template <int ITEMS>
__device__ __forceinline__ void phase_rank(const DNParams &p, const Smem &s, int Kv, int tid, int nthr, uint32_t (&item)[ITEMS]) {
    const int Bm = p.B - 1;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const int t = tid + i * nthr;
        item[i] = 0xffffffffu;
        if (t < Kv) {
            const unsigned long long key = s.key[t];
            const uint32_t lo32 = (uint32_t)key;
            const int gb = (int)(lo32 & 0xffffu);
            const int c = gb >> p.Bshift;
            const int base = s.start[c];
            const int st = base + s.cntb[gb];
            const int en = base + (((gb & Bm) == Bm) ? s.cnt[c] : s.cntb[gb + 1]);
            int rank = 0;
#pragma unroll 4
            for (int u = st; u < en; ++u) rank += (s.key[u] > key) ? 1 : 0;
            const int pc = st + rank - base;                      // position inside the class
            const uint32_t pp = (uint32_t)(32 * s.ktile[c] + pc);  // tile-padded position
            item[i] = (pp << 16) | (0xffffu - (lo32 >> 16));
        }
    }
}

// fp16 images of a box for the prefilter (see h16_prefilter).  Scales: X = x / 16, Y = 1024 * y, so for |coordinates|
// <= 8 the enlarged overlap width is <= 1 (HADD2.SAT clamps it at 0 from below for free), overlap heights stay below
// 16384, and areas down to 3.8e-6 (a 0.7 x 0.7 pixel box at 352 x 352) keep a NORMAL fp16 t*area*64.  Anything else --
// non-finite or far-away coordinates, tiny or huge areas, a threshold outside [0.01, 1] -- gets TA = -inf: every
// pair with the box is a "maybe" and is decided by the exact arithmetic.
constexpr float kH16SX = 0.0625f, kH16SY = 1024.0f;

__device__ __forceinline__ void h16_store(const Smem &s, uint32_t pp, const float4 &b, const IouThr &t) {
    H16Tile &tile = s.h16[pp >> 5];
    const int k = pp & 15, hi = (pp >> 4) & 1;
    const float a = box_area(b);
    const float big = fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w)));
    const bool ok = t.fast_ok && a >= 3.814697265625e-06f && a <= 256.0f && big <= 8.0f;  // (false for NaN)
    __half *xy = reinterpret_cast<__half *>(&tile.xy[k]);
    xy[0 + hi] = ok ? __float2half_rd(__fmul_rn(b.x, kH16SX)) : __ushort_as_half((unsigned short)0);
    xy[2 + hi] = ok ? __float2half_rd(__fmul_rn(b.y, kH16SY)) : __ushort_as_half((unsigned short)0);
    xy[4 + hi] = ok ? __float2half_ru(__fmul_rn(b.z, kH16SX)) : __ushort_as_half((unsigned short)0);
    xy[6 + hi] = ok ? __float2half_ru(__fmul_rn(b.w, kH16SY)) : __ushort_as_half((unsigned short)0);
    // lower bound of t * area * 64: the 2^-8 margin covers the three fp16 roundings of the test with room to
    // spare (see h16_prefilter)
    reinterpret_cast<__half *>(&tile.ta[k])[hi] = ok ? __float2half_rd(__fmul_rn(a, t.th)) : __ushort_as_half((unsigned short)0xfc00);
}

template <int ITEMS>
__device__ __forceinline__ void phase_write_sorted(const DNParams &p, const Smem &s, const uint32_t (&item)[ITEMS]) {
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        if (item[i] != 0xffffffffu) {
            const uint32_t pp = item[i] >> 16, cid = item[i] & 0xffffu;
            s.scid[pp] = (uint16_t)cid;
            h16_store(s, pp, s.box[cid], p.iou);
        }
    }
}

// ---------------------------------------------------------------------------
// P5: pair blocks.
// ---------------------------------------------------------------------------
constexpr float kPairEps = 1e-5f;

__device__ __forceinline__ uint32_t h2u(__half2 h) { return *reinterpret_cast<uint32_t *>(&h); }
__device__ __forceinline__ __half2 u2h(uint32_t u) { return *reinterpret_cast<__half2 *>(&u); }

// The row operands of the prefilter: the row's own H16Tile entry, broadcast into both halves
struct H16Row { uint32_t x1, y1, x2, y2, nta; };

__device__ __forceinline__ H16Row h16_row(const H16Tile &t, int lane) {
    const int k = lane & 15;
    const uint32_t sel = (lane & 16) ? 0x3232u : 0x1010u;
    const uint4 e = t.xy[k];
    H16Row r;
    r.x1 = __byte_perm(e.x, 0, sel);
    r.y1 = __byte_perm(e.y, 0, sel);
    r.x2 = __byte_perm(e.z, 0, sel);
    r.y2 = __byte_perm(e.w, 0, sel);
    r.nta = __byte_perm(t.ta[k], 0, sel) ^ 0x80008000u;   // -TA
    return r;
}

// Conservative fp16 prefilter of one row against the 32 columns of a tile: bit j CLEAR = torchvision provably
// does not suppress column j by this row; bit j set = maybe.
//   W = sat(min(X2r, X2c) - max(X1r, X1c)),  H = min(Y2r, Y2c) - max(Y1r, Y1c),  D = W * H - (TAr + TAc);  maybe iff D >= 0
// Why D < 0 is a proof.  The fp16 boxes contain the fp32 boxes (x1, y1 rounded down, x2, y2 rounded up; min / max
// are exact), so the exact overlap extents W*, H* of the fp16 boxes bound the true ones from above (times the
// scales).  W and H are single roundings of W*, H* (relative error <= 2^-11; differences of fp16 numbers that
// land below the normal range are exact), the sum is one more rounding, and the FMA computes W * H - SUM with
// ONE rounding, which never changes a sign (a nonzero result that underflows keeps its sign bit).  Hence
//   D < 0  =>  64 * inter_true * (1 - 2^-11)^2  <=  W * H  <  SUM  <=  64 * t * (Sa + Sb) * (1 - 2^-8) * (1 + 2^-11)
//          =>  inter_true  <  t * (Sa + Sb) * (1 - 0.0024)            (Sa, Sb: the fp32 areas torchvision uses)
// i.e. IoU < thr * (1 - 0.003) in exact arithmetic, and torchvision's fp32 evaluation of the IoU is within 4e-7
// of the exact one.  If an extent is <= 0 the true boxes do not overlap on that axis either: W clamps to 0, or
// H <= 0, and D = -SUM < 0.  Degenerate boxes carry TA = -inf, so D = +inf: maybe.  NaN cannot appear: every
// table entry of a valid column is finite except TA = -inf, and -inf only ever meets finite numbers or itself.
// S: steps = column pairs (k, k + 16) visited; S < 16 serves tiles with at most S valid columns (the class's last tile)
template <int S>
__device__ __forceinline__ uint32_t h16_prefilter_steps(const H16Tile &t, const H16Row &r) {
    uint32_t acc = 0u;
#pragma unroll
    for (int k = 0; k < S; ++k) {
        const uint4 c = t.xy[k];
        const __half2 w = __hsub2_sat(__hmin2(u2h(r.x2), u2h(c.z)), __hmax2(u2h(r.x1), u2h(c.x)));
        const __half2 h = __hsub2(__hmin2(u2h(r.y2), u2h(c.w)), __hmax2(u2h(r.y1), u2h(c.y)));
        const __half2 nsum = __hsub2(u2h(r.nta), u2h(t.ta[k]));   // -(TAr + TAc)
        const uint32_t d = h2u(__hfma2(w, h, nsum));
        acc = (acc >> 1) | (d & 0x80008000u);                      // sign bits: step k's land at 16 - S + k and 32 - S + k
    }
    return (S == 16) ? ~acc : ((~acc >> (16 - S)) & ((1u << S) - 1u));
}

__device__ __forceinline__ uint32_t h16_prefilter(const H16Tile &t, const H16Row &r, int ncol) {
    if (ncol > 12) return h16_prefilter_steps<16>(t, r);   // (13..16 columns: the high halves are masked by `valid`)
    if (ncol > 8) return h16_prefilter_steps<12>(t, r);
    if (ncol > 4) return h16_prefilter_steps<8>(t, r);
    return h16_prefilter_steps<4>(t, r);
}

// One pair, decided exactly as torchvision decides it (call site utils/box.py:28).  The divide-free comparison
// inter > t * (Sa + Sb) is used when it is not a close call: its roundings and torchvision's are each < 5e-7
// relative, so outside a 1e-5 band the two agree; inside the band, for degenerate magnitudes and for thresholds
// outside [0.01, 1] the exact routine (IEEE divide, compare in double) decides.
__device__ __forceinline__ bool pair_decide(const float4 &R, float ra, const float4 &Cb, const IouThr &t) {
    const float ca = box_area(Cb);
    const float w = fmaxf(0.0f, __fsub_rn(fminf(R.z, Cb.z), fmaxf(R.x, Cb.x)));
    const float h = fmaxf(0.0f, __fsub_rn(fminf(R.w, Cb.w), fmaxf(R.y, Cb.y)));
    const float inter = __fmul_rn(w, h);
    const float sum = __fadd_rn(ra, ca);
    const float rhs = __fmul_rn(sum, t.tf);
    const float diff = __fsub_rn(inter, rhs);
    const bool safe = t.fast_ok && sum > 1e-30f && sum < 1e30f && inter < 1e30f;   // (false for NaN)
    if (safe && fabsf(diff) > __fmul_rn(kPairEps, rhs)) return diff > 0.0f;
    return nms_suppress_exact(R, ra, Cb, ca, t.thr);
}

// the exact decisions for the "maybe" bits of one block: bit j of the result = this lane's row suppresses column j
// of tile gct.  Either every lane walks its own bits (max popcount turns), or the warp walks the union column by column
// (uniform column loads, popcount(union) turns): whichever is fewer instructions.
__device__ __forceinline__ uint32_t resolve_maybe(const DNParams &p, const Smem &s, int gct, uint32_t maybe, int row_pp) {
    const int mx = __reduce_max_sync(kFullMask, (unsigned)__popc(maybe));
    if (mx == 0) return 0u;
    // (lanes without a "maybe" bit may sit past the end of the class: their scid entry is not a cell id)
    const float4 R = maybe ? s.box[s.scid[row_pp]] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float ra = box_area(R);
    const uint16_t *cids = s.scid + 32 * gct;
    uint32_t word = 0u;
    uint32_t uni = __reduce_or_sync(kFullMask, maybe);
    if (5 * mx <= 4 * __popc(uni)) {
        for (int it = 0; it < mx; ++it) {
            if (maybe) {
                const int j = __ffs(maybe) - 1;
                maybe &= maybe - 1u;
                if (pair_decide(R, ra, s.box[cids[j]], p.iou)) word |= 1u << j;
            }
        }
    } else {
        while (uni) {
            const int j = __ffs(uni) - 1;
            uni &= uni - 1u;
            const float4 Cb = s.box[cids[j]];
            if (((maybe >> j) & 1u) && pair_decide(R, ra, Cb, p.iou)) word |= 1u << j;
        }
    }
    return word;
}

// Shared-memory hand-overs between the warps of a CTA: release / acquire at CTA scope, and EVERY access to a word that
// another warp may touch concurrently is an atomic instruction (ATOMS) -- the hardware cost is a handful of
// instructions per task, and compute-sanitizer's racecheck, which only understands barriers and atomics, stays clean
// (profiles/r02/sanitizer_racecheck.log) instead of reporting hand-overs that are ordered by flags.
__device__ __forceinline__ int ld_acquire_s(const int *a) {
    int v;
    asm volatile("atom.acquire.cta.shared.or.b32 %0, [%1], 0;" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(a)) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_s(int *a, int v) {
    int old;
    asm volatile("atom.release.cta.shared.exch.b32 %0, [%1], %2;" : "=r"(old) : "r"((uint32_t)__cvta_generic_to_shared(a)), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_add_s(int *a, int v) {
    asm volatile("red.release.cta.shared.add.s32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(a)), "r"(v) : "memory");
}
__device__ __forceinline__ void red_or_s(uint32_t *a, uint32_t v) {
    asm volatile("red.relaxed.cta.shared.or.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(a)), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_s(const uint32_t *a) {
    uint32_t v;
    asm volatile("atom.relaxed.cta.shared.or.b32 %0, [%1], 0;" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(a)) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_s(uint32_t *a, uint32_t v) {
    uint32_t old;
    asm volatile("atom.relaxed.cta.shared.exch.b32 %0, [%1], %2;" : "=r"(old) : "r"((uint32_t)__cvta_generic_to_shared(a)), "r"(v) : "memory");
}

// Tasks.  HEAD task of tile rt of class c: (1) when the kept rows of all earlier tiles have contributed to the tile's
// suppressed-column word (arrived[g] == rt), the tile is resolved in score order (its diagonal block) and its kept word
// is final; (2) its kept rows meet the columns of tile rt + 1 -- the block the next head task waits for -- and, unless
// the class has many tiles left, of every later tile too.  TAIL task of tile rt (classes with many tiles): the kept
// rows of tile rt against the tiles rt + 2, rt + 3, ... in that order, on another warp.  Every block reduces at once
// to one word (the columns its KEPT rows suppress) that is OR-ed into supw[ct]; rows that are already suppressed and
// columns that are already suppressed cost nothing.  The head tasks of a class form a chain (two blocks per link); the
// tail strips run beside it and deliver a column tile before the chain needs it.  Tasks are claimed in table order --
// heads of tile 0, tails of tile 0, heads of tile 1, ... -- so a warp only ever waits for tasks that some warp has
// already claimed: no deadlock.
// (Tried and dropped, profiles/r02/NOTES.md: evaluating a task's blocks for ALL rows before looking at its dependency --
// full parallelism, but 19 % slower on cfg2: more code in the loop body and no skipping.)

// kept rows of tile grt (word rem, this lane's row operands `row`) against the column tiles [ct0, ct1) of the class
__device__ __forceinline__ void strip_blocks(const DNParams &p, const Smem &s, int n, int grt, int rt, int ct0, int ct1, uint32_t rem,
                                             const H16Row &row) {
    const int lane = threadIdx.x & 31;
    const bool kept = (rem >> lane) & 1u;
    for (int ct = ct0; ct < ct1; ++ct) {
        const int gct = grt + (ct - rt);
        const int ncol = min(32, n - 32 * ct);
        uint32_t supd = 0u;
        if (lane == 0) supd = ld_relaxed_s(&s.supw[gct]);
        supd = __shfl_sync(kFullMask, supd, 0);
        const uint32_t open = ((ncol >= 32) ? 0xffffffffu : ((1u << ncol) - 1u)) & ~supd;
        uint32_t sup = 0u;
        if (rem != 0u && open != 0u) {
            const uint32_t maybe = kept ? (h16_prefilter(s.h16[gct], row, ncol) & open) : 0u;
            sup = __reduce_or_sync(kFullMask, resolve_maybe(p, s, gct, maybe, 32 * grt + lane));
        }
        if (lane == 0) {
            if (sup) red_or_s(&s.supw[gct], sup);
            red_release_add_s(&s.arrived[gct], 1);
        }
    }
}

__device__ __forceinline__ void phase_pairs(const DNParams &p, const Smem &s) {
    const int lane = threadIdx.x & 31;
    const int ntask = s.misc[M_NTASK];
    for (;;) {
        int q = 0;
        if (lane == 0) q = atomicAdd(&s.misc[M_CTR], 1);
        q = __shfl_sync(kFullMask, q, 0);
        if (q >= ntask) break;
        const uint32_t tk = s.tasks[q];
        const int c = (int)(tk >> 16), rt = (int)(tk & 0x7fffu);
        const int n = s.cnt[c];
        const int T = (n + 31) >> 5;
        const int grt = s.ktile[c] + rt;
        const H16Row row = h16_row(s.h16[grt], lane);
        uint32_t rem;
        int ct0, ct1;
        if (tk & kTaskTail) {
            rem = 0u;
            if (lane == 0) {
                while (!(ld_acquire_s(&s.ready[grt]) & 1)) __nanosleep(20);
                rem = ld_relaxed_s(&s.keptw[grt]);
            }
            rem = __shfl_sync(kFullMask, rem, 0);
            ct0 = rt + 2;
            ct1 = T;
        } else {
            const int nrow = min(32, n - 32 * rt);
            const uint32_t validr = (nrow >= 32) ? 0xffffffffu : ((1u << nrow) - 1u);
            // (1) the tile's own turns
            uint32_t supd = 0u;
            if (lane == 0) {
                if (rt > 0) {
                    while (ld_acquire_s(&s.arrived[grt]) < rt) __nanosleep(20);
                }
                supd = ld_relaxed_s(&s.supw[grt]);
            }
            supd = __shfl_sync(kFullMask, supd, 0);
            const uint32_t alive = validr & ~supd;
            rem = alive;
            if (alive & (alive - 1u)) {   // two or more alive candidates
                uint32_t maybe = 0u;
                if ((alive >> lane) & 1u) maybe = h16_prefilter(s.h16[grt], row, nrow) & alive & ~((2u << lane) - 1u);  // LATER columns
                const uint32_t D = resolve_maybe(p, s, grt, maybe, 32 * grt + lane);
                uint32_t nz = __ballot_sync(kFullMask, D != 0u);
                while (nz) {
                    const int i = __ffs(nz) - 1;
                    nz &= nz - 1u;
                    const uint32_t Di = __shfl_sync(kFullMask, D, i);
                    if ((rem >> i) & 1u) rem &= ~Di;
                }
            }
            const bool tail = has_tail(T, rt);
            if (lane == 0) {
                st_relaxed_s(&s.keptw[grt], rem);
                if (tail) st_release_s(&s.ready[grt], (c << 16) | 1);
            }
            // (2) the kept rows against the next tile (the tail task takes the others), or against all later tiles
            ct0 = rt + 1;
            ct1 = tail ? rt + 2 : T;
        }
        strip_blocks(p, s, n, grt, rt, ct0, ct1, rem, row);   // (one call site: the block code exists once)
        __syncwarp();
    }
}

// fp32 pair blocks over a {shared address of the box, t*area} table: used by large_nms.cuh
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

// ---------------------------------------------------------------------------
// P6: output.  Tile g of the kept bitmap holds up to 32 rows that are consecutive in the output (class-ascending,
// score-descending); a warp owns a contiguous range of tiles, hence a CONTIGUOUS range of the output.  It streams
// that range through its scratch: rows are appended as 7 floats each, and whenever >= 96 floats are pending (and at
// the end of the range) the pending floats leave as 16-byte stores aligned to 16 bytes in global memory -- only the
// first and the last < 4 floats of the warp's range go out as scalars.  Few, long, aligned requests: what HBM likes
// and, for the fused all-gather, what NVLink likes (trained-like heads keep ~4 rows per class: a store per tile
// would be a 100-byte packet per class and peer).
// ---------------------------------------------------------------------------
template <int MODE, int THREADS, int GATHER>
__device__ __forceinline__ void phase_output_stream(const DNParams &p, const Smem &s, int b) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = THREADS / 32;
    const int ntiles = s.ktile[p.C];
    const int nbuf = (GATHER != 0) ? p.gnbuf : 1;   // fused all-gather: one copy per rank, or one multicast store
    const size_t img = (GATHER != 0) ? (size_t)(p.gslot + b) : (size_t)b;
    float *scr = s.scratch + warp * kScratchFloats;
    // MODE_NMS: the caller's own rows are gathered bit-for-bit (pred_this_cls[index], box.py:29)
    const int K0 = (MODE == MODE_NMS) ? min(p.cand_count[0][b], p.cand_stride[0]) : 0;
    const float *r0 = (MODE == MODE_NMS) ? p.cand[0] + (size_t)b * p.cand_stride[0] * 7 : nullptr;
    const float *r1 = (MODE == MODE_NMS && p.cand[1]) ? p.cand[1] + (size_t)b * p.cand_stride[1] * 7 : nullptr;
    // each warp owns a contiguous range of tiles: one prefix over the kept bitmap per warp
    const int per = (ntiles + kWarps) / kWarps;  // ceil((ntiles + 1) / kWarps): "tile" ntiles writes the count
    const int g0 = warp * per, g1 = min(g0 + per, ntiles + 1);
    int before = 0;
    for (int t = lane; t < g0; t += 32) before += __popc(s.keptw[t]);
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) before += __shfl_xor_sync(kFullMask, before, sh);
    // Streaming state.  e = float index (in a gather buffer / in `out`) of the next float this warp emits; scratch[i]
    // holds the float that belongs at gbase + i, gbase a multiple of 4 in the sense of the 16-byte alignment of the
    // destination; the first `hole` entries (only before the first flush) belong to the warp before this one.
    size_t e = (img * (size_t)p.K + (size_t)before) * 7;
    const float *abase = (GATHER != 0) ? p.gout[0] : p.out;          // (every buffer has the same alignment: 256 B)
    int hole = (int)((((uintptr_t)(abase + e)) >> 2) & 3);
    size_t gbase = e - (size_t)hole;
    int pend = hole;
    // the pending floats [hole, pend) leave: scalars for a first group with holes, 16-byte stores for whole groups, and
    // (final) scalars for the last < 4 floats; what stays moves to the front of the scratch
    auto flush = [&](bool final) {
        __syncwarp();
        int v0 = 0;
        const int nvec = pend >> 2;
        if (hole > 0 && (nvec > 0 || final)) {
            const int hi = min(4, pend);
            if (lane >= hole && lane < hi) {
                const float v = scr[lane];
#pragma unroll
                for (int r = 0; r < kMaxPeers; ++r) {
                    if (r >= nbuf) break;
                    ((GATHER != 0) ? p.gout[r] : p.out)[gbase + lane] = v;
                }
            }
            hole = 0;
            v0 = 1;
        }
        for (int v = v0 + lane; v < nvec; v += 32) {
            const float4 a = *reinterpret_cast<const float4 *>(scr + 4 * v);
#pragma unroll
            for (int r = 0; r < kMaxPeers; ++r) {  // (peer stores travel over NVLink while the next tiles are assembled)
                if (r >= nbuf) break;
                *reinterpret_cast<float4 *>(((GATHER != 0) ? p.gout[r] : p.out) + gbase + 4 * (size_t)v) = a;
            }
        }
        const int done = min(pend, 4 * max(nvec, v0));   // scratch entries that are out (a first group with holes counts)
        const int rem = pend - done;                      // 0..3
        float keep = 0.f;
        if (lane < rem) keep = scr[done + lane];
        if (final && lane < rem && done + lane >= hole) {
#pragma unroll
            for (int r = 0; r < kMaxPeers; ++r) {
                if (r >= nbuf) break;
                ((GATHER != 0) ? p.gout[r] : p.out)[gbase + done + lane] = keep;
            }
        }
        __syncwarp();
        if (!final && done > 0 && lane < rem) scr[lane] = keep;
        if (!final) {
            gbase += (size_t)done;
            pend = rem;
        }
        __syncwarp();
    };
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
        const uint32_t word = s.keptw[g];
        const int nk = __popc(word);
        if (nk == 0) continue;
        if ((word >> lane) & 1u) {
            const int r = __popc(word & lanemask_lt());
            const uint32_t cid = s.scid[32 * g + lane];
            float *d = scr + pend + 7 * r;
            if (MODE == MODE_NMS) {
                const float *src = ((int)cid < K0) ? r0 + (size_t)cid * 7 : r1 + (size_t)((int)cid - K0) * 7;
#pragma unroll
                for (int k = 0; k < 7; ++k) d[k] = __ldg(src + k);
            } else {
                const float4 bx = s.box[cid];
                const float2 cs = s.cs[cid];
                d[0] = bx.x; d[1] = bx.y; d[2] = bx.z; d[3] = bx.w;
                d[4] = cs.x; d[5] = cs.y;
                d[6] = (float)(s.ready[g] >> 16);  // cls_idx.float() (yolo_loss.py:199)
            }
            if (p.out_idx) p.out_idx[(size_t)b * p.K + before + r] = (int)cid;
        }
        pend += 7 * nk;
        before += nk;
        if (pend >= 96) flush(false);   // (a tile adds at most 224 floats: 95 + 224 < kScratchFloats)
    }
    if (pend > hole) flush(true);
}

// The ordinary (single-buffer) kernel keeps the simpler per-tile writer: a tile's rows are assembled in the scratch and
// leave with coalesced 4-byte stores.  HBM does not care about the request size the way NVLink does, and the
// streaming writer's bookkeeping costs this kernel 4-7 % (measured, profiles/r02/NOTES.md).
template <int MODE, int THREADS, int GATHER>
__device__ __forceinline__ void phase_output(const DNParams &p, const Smem &s, int b) {
    if constexpr (GATHER != 0) {
        phase_output_stream<MODE, THREADS, GATHER>(p, s, b);
        return;
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = THREADS / 32;
    const int ntiles = s.ktile[p.C];
    float *o = p.out + (size_t)b * p.K * 7;
    float *scr = s.scratch + warp * kScratchFloats;
    // MODE_NMS: the caller's own rows are gathered bit-for-bit (pred_this_cls[index], box.py:29)
    const int K0 = (MODE == MODE_NMS) ? min(p.cand_count[0][b], p.cand_stride[0]) : 0;
    const float *r0 = (MODE == MODE_NMS) ? p.cand[0] + (size_t)b * p.cand_stride[0] * 7 : nullptr;
    const float *r1 = (MODE == MODE_NMS && p.cand[1]) ? p.cand[1] + (size_t)b * p.cand_stride[1] * 7 : nullptr;
    // each warp owns a contiguous range of tiles: one prefix over the kept bitmap per warp
    const int per = (ntiles + kWarps) / kWarps;  // ceil((ntiles + 1) / kWarps): "tile" ntiles writes the count
    const int g0 = warp * per, g1 = min(g0 + per, ntiles + 1);
    int before = 0;
    for (int t = lane; t < g0; t += 32) before += __popc(s.keptw[t]);
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) before += __shfl_xor_sync(kFullMask, before, sh);
    for (int g = g0; g < g1; ++g) {
        if (g == ntiles) {
            if (lane == 0) p.out_count[b] = before;
            break;
        }
        const uint32_t word = s.keptw[g];
        const int nk = __popc(word);
        if (nk == 0) continue;
        if ((word >> lane) & 1u) {
            const int r = __popc(word & lanemask_lt());
            const uint32_t cid = s.scid[32 * g + lane];
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
                d[6] = (float)(s.ready[g] >> 16);  // cls_idx.float() (yolo_loss.py:199)
            }
            if (p.out_idx) p.out_idx[(size_t)b * p.K + before + r] = (int)cid;
        }
        __syncwarp();
        const int nf = 7 * nk;
        float *dst = o + (size_t)7 * before;
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const int f = 32 * k + lane;
            if (f < nf) dst[f] = scr[f];
        }
        __syncwarp();
        before += nk;
    }
}

// ---------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------
// GATHER: the fused all-gather variant of the output phase (its own instantiation, so the ordinary kernel's register
// allocation is untouched).  DBG: phase time stamps (profiles/phase_times.py).  EXACT: the reference's own decode
// arithmetic (b200yolo_set_exact_decode), instantiated for the runtime-shape variants only.
template <int MODE, int THREADS, int SHAPE, int GATHER = 0, bool DBG = false, bool EXACT = false>
__global__ void __launch_bounds__(THREADS, (THREADS == 384) ? 3 : (THREADS == 512) ? 2 : 1) decode_nms_kernel(const DNParams p, const SmemLayout L) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem s = carve(smem_raw, L, p.K, p.C);
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = p.C, K = p.K;
    using SH = ShapeT<SHAPE>;

    if constexpr (GATHER == 0) {
        if (p.chain >= 2 && b == p.N) {   // the chain CTA
            if (p.chain == 3) pdl_trigger();
            pdl_wait();
            pdl_trigger();
            return;
        }
        if (p.chain != 1) pdl_trigger();  // the next launch may start filling free SM slots right away
    } else {
        pdl_trigger();
    }
    if constexpr (GATHER != 0) {
        // The arrival signal of the PREVIOUS step rides on this launch (a kernel of its own per step costs ~2.3 us of
        // launch processing, and a system-scope fence in every CTA of the step itself flushes the L1 the decode
        // streams through: +8 us, profiles/r02/NOTES.md).  The extra CTA idles in one of 296 slots until the
        // previous launch has completed and flushed, then publishes: the peers may read that step's rows.
        if (b == p.N) {
            pdl_wait();
            if (tid < p.gR) {
                __threadfence_system();
                *reinterpret_cast<volatile int *>(p.gflags[tid] + p.grank) = p.gsignal;
            }
            return;
        }
    }
    // Inputs another kernel of this library wrote (rows of our own decode kernels), or a producer launched with the
    // programmatic attribute, must be complete and visible before the first global read: p.wait_inputs is set by the
    // host unless the caller vouches that the inputs were produced in ordinary stream order (see launch_dn_t).
    if (MODE == MODE_NMS || p.wait_inputs || p.chain == 1 || (DBG && p.dbg && !(p.flags & 512))) pdl_wait();
    if constexpr (GATHER == 0) {
        if (p.chain == 1) pdl_trigger();
    }
    stamp<DBG>(p, b, 0);
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
        if (!kStaticShape) __syncthreads();  // (compile-time shapes: the barrier sits behind the first round's loads)
    }
    stamp<DBG>(p, b, 15);

    if constexpr (MODE == MODE_NMS) {
        phase_load_rows<THREADS>(p, s, b);
    } else if constexpr (kStaticShape && MODE == MODE_DECODE) {
        decode_head_static<THREADS, MODE, SH::C, SH::HW0, SH::W0, false, DBG, EXACT>(p, s, b, p.head[0], 0, 0);
    } else if constexpr (kStaticShape) {
        int h0_active = 0, h0_pass = 0;
        decode_head_static<THREADS, MODE, SH::C, SH::HW0, SH::W0, true, DBG, EXACT>(p, s, b, p.head[0], 0, 0, &h0_active, &h0_pass);
        // (warp-uniform) at most an eighth of this warp's first-head cells passed: objectness first.  Warps that own no
        // first-head cell run ahead with the ordinary loop (probing the first head from them, or letting them try
        // unconditionally, costs dense images 1 %; loading the second head's objectness together with the first head's
        // planes costs them 6 % and sparse ones gain nothing more: measured, profiles/r02/NOTES.md).  Flag 16384: off.
        bool done = false;
        if (MODE == MODE_FUSED && !(p.flags & 16384) && h0_active > 0 && 8 * h0_pass <= h0_active)
            done = decode_head_static_sparse<THREADS, MODE, SH::C, SH::HW1, SH::W1, EXACT>(p, s, b, p.head[1], p.head[0].cells);
        if (!done)
            decode_head_static<THREADS, MODE, SH::C, SH::HW1, SH::W1, false, DBG, EXACT>(p, s, b, p.head[1], p.head[0].cells, 1);
    } else if constexpr (kNhwcShape) {
        // scratch: the key region of U, free during the decode; the host checked that it fits
        float *scr = reinterpret_cast<float *>(smem_raw + L.key_off);
        decode_head_nhwc<THREADS, SH::C, EXACT>(p, s, b, p.head[0], 0, scr);
        decode_head_nhwc<THREADS, SH::C, EXACT>(p, s, b, p.head[1], p.head[0].cells, scr);
    } else {
        if (MODE == MODE_FUSED && p.nhwc) {
            float *scr = reinterpret_cast<float *>(smem_raw + L.key_off);
            decode_head_nhwc<THREADS, 0, EXACT>(p, s, b, p.head[0], 0, scr);
            decode_head_nhwc<THREADS, 0, EXACT>(p, s, b, p.head[1], p.head[0].cells, scr);
        } else {
            decode_head_rt<THREADS, MODE, DBG, EXACT>(p, s, b, p.head[0], 0, 0);
            if (MODE == MODE_FUSED) decode_head_rt<THREADS, MODE, DBG, EXACT>(p, s, b, p.head[1], p.head[0].cells, 1);
        }
    }
    __syncthreads();
    stamp<DBG>(p, b, 1);

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

    // P2: bucket positions inside each class (warp per class), then the class starts and tile tables (warp 0)
    scan_buckets_per_class<THREADS>(p, s);
    __syncthreads();
    stamp<DBG>(p, b, 12);
    // A phase that one warp executes alone runs at instruction-fetch latency (the two resident CTAs are in different
    // phases of a ~100 KB kernel: ~200 cycles per 8 instructions, profiles/r02/NOTES.md), and the other 15 warps wait
    // for it: only the class starts, which the key scatter needs, stay in such a window.
    if (warp == 0) warp_class_starts(p, s);
    __syncthreads();
    stamp<DBG>(p, b, 13);
    const int Kv = s.misc[M_KV];
    if (warp == 0) {
        // the tile tables (the ranking needs ktile: barrier 2) and the strip tasks (only the pair phase needs them)
        // are built while the other warps sort
        warp_tile_tables(p, s);
        asm volatile("bar.arrive 2, %0;" ::"n"(THREADS) : "memory");
        warp_build_tasks(p, s);
    } else {
        // P3 / P4 on the other warps (barrier 2 between the key scatter and the ranking: all keys in place, and warp 0's
        // tile tables; named barrier 1 between the ranking and the writes that reuse the key memory)
        constexpr int kSorters = THREADS - 32;
        uint32_t item[kRankItems];
        phase_scatter_keys(p, s, tid - 32, kSorters);
        asm volatile("bar.sync 2, %0;" ::"n"(THREADS) : "memory");
        phase_rank<kRankItems>(p, s, Kv, tid - 32, kSorters, item);
        asm volatile("bar.sync 1, %0;" ::"n"(kSorters) : "memory");
        phase_write_sorted<kRankItems>(p, s, item);
    }
    __syncthreads();
    stamp<DBG>(p, b, 3);
    // P5
    phase_pairs(p, s);
    __syncthreads();
    stamp<DBG>(p, b, 4);
    // P6 (the fp16 tables are dead now; the row scratch aliases them).  The ordinary kernel waits for its predecessor
    // before its first store (consecutive launches may share the output buffers).  The gather variant needs no such
    // wait: consecutive steps store into different buffers, and what orders a step against the previous users of ITS
    // buffer is the flag wait below -- so a step's stores, and their NVLink round trips, overlap the previous step's tail.
    if constexpr (GATHER == 0) {
        if (p.chain < 2) pdl_wait();
    }
    if constexpr (GATHER != 0) {
        // the gather buffer this step writes was last used three steps ago: wait until every rank has completed the
        // previous step, i.e. has moved past everything it ran on that buffer (dist.PeerGather, back-pressure)
        if (p.gwait_value > 0) {
            if (warp == 0 && lane < p.gR) {
                const long long t0 = clock64();
                while (*reinterpret_cast<const volatile int *>(p.gwait_flags + lane) < p.gwait_value) {
                    if (clock64() - t0 > p.gwait_cycles) { *p.gtimed_out = 1; break; }   // a dead peer must not hang the GPU
                    __nanosleep(64);
                }
                // (no fence: what follows are STORES that depend on this load's outcome; a system-scope fence here would
                // also flush the L1 the co-resident CTA streams its heads through)
            }
            __syncthreads();
        }
    }
    stamp<DBG>(p, b, 5);
    phase_output<MODE, THREADS, GATHER>(p, s, b);
    stamp<DBG>(p, b, 7);
}

}  // namespace b200yolo
