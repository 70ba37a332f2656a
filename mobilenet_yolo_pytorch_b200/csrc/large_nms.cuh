// large_nms.cuh -- decode + per-class NMS for images with more candidate cells than one CTA can keep in shared
// memory (decode_nms.cuh stages ~40 B per cell: about 5.6 k cells in 227 KB).  SURVEY.md 8(d) asks for the 832x832
// variant of config 5: heads (N,75,26,26) + (N,75,52,52) = 10 140 cells per image.
//
// One CTA of 1024 threads per image, same reference semantics as the fused kernel (yolo_loss.py:186-203,
// utils/box.py:16-30, torchvision nms) and the same arithmetic (SFU sigmoid / exp, class near-tie fallback,
// conservative fp16 prefilter of the box pairs with the exact torchvision decision for the pairs it cannot rule out):
//   P1  decode every cell (thread per cell, plane loads coalesced); a passing cell's record {box, conf, score,
//       t*area*2^-13, class} goes to the caller's workspace (32 B per cell, stays in L2) and its 64-bit sort key
//       (class asc | score desc | cell id asc) to shared memory -- 8 B per cell is all the CTA stages;
//   P2  bitonic sort of the keys (padded to a power of two with ~0);
//   P3  class starts / tile table (warp 0);
//   P4  greedy NMS over tasks = (class, tile of 32 sorted candidates), claimed from a counter in tile-major
//       order (tile 0 of every class, tile 1 of every class, ...: the classes advance side by side).  A task
//       (lane = candidate) stages each earlier tile of its class as the 32 columns of one pair block (fp32 boxes
//       + their H16Tile), runs the fp16 prefilter of decode_nms.cuh (h16_prefilter) and decides exactly
//       (pair_decide) only the "maybe" pairs whose column was KEPT -- the tile's kept word comes from a per-tile
//       ready flag; a task only depends on tasks claimed before it, so the spin cannot deadlock -- then resolves
//       its own tile in score order and publishes its kept word.  (Round 1 evaluated the fp32 pair block for every
//       column: 0.53 ms per 128-image step at 832x832; now 0.43 ms.)  No n^2 mask storage: the worst case (one class holding everything) is
//       bounded by time, not memory;
//   P5  kept rows in (class asc, score desc) order -> out / out_idx / out_count.
#pragma once
#include "decode_nms.cuh"

namespace b200yolo {

constexpr int kLargeThreads = 1024;
constexpr int kLargeMaxKeys = 16384;   // 128 KB of keys
constexpr int kLargeStageBytes = 32 * 16 + 320;   // per warp: 32 staged column boxes + their H16Tile (decode_nms.cuh)

struct LargeParams {
    HeadDesc head[2];
    int N, A, C, attrs;
    int K, P;            // cells per image; P = smallest power of two >= max(K, 32)
    int T;               // tile slots: K/32 + C + 1
    float conf_thr;
    IouThr iou;
    float4 *rec;         // workspace [N][K][2] (heads as the source)
    const float *cand[2];        // SRC = 1 (utils.box.nms on caller rows): [N][stride][7] candidate rows of the two heads,
    const int *cand_count[2];    // their per-image counts
    int cand_stride[2];
    float *out;          // [N][K][7]
    int *out_count;      // [N]
    int *out_idx;        // [N][K] or null
};

__host__ __device__ inline size_t large_smem_bytes(int P, int C, int T) {
    size_t o = (size_t)8 * P + (size_t)4 * (3 * (C + 2)) + (size_t)12 * T + 64;
    o = (o + 15) & ~(size_t)15;
    return o + (size_t)(kLargeThreads / 32) * kLargeStageBytes;
}

struct LargeCell { float4 bx; float conf, best; int bi; };

// one cell of a head (yolo_loss.py:186-199, 243-247); false: conf <= val_conf (:201)
__device__ __forceinline__ bool large_decode_cell(const HeadDesc &hd, const float *hb, int C, int local, float conf_thr,
                                                  LargeCell &o) {
    const int attrs = 5 + C, HW = hd.HW;
    const int a = fastdiv(local, hd.magicHW);
    const int pos = local - a * HW;
    const float *q = hb + (size_t)a * attrs * HW + pos;
    const float tx = __ldcs(q), ty = __ldcs(q + HW), tw = __ldcs(q + 2 * (size_t)HW), th = __ldcs(q + 3 * (size_t)HW);
    o.conf = sigmoid_fast(__ldcs(q + 4 * (size_t)HW));             // :189,197
    if (!(o.conf > conf_thr)) return false;                        // :201
    const float *qc = q + 5 * (size_t)HW;
    float m1 = __ldg(qc);
    for (int c = 1; c < C; ++c) m1 = fmaxf(m1, __ldg(qc + (size_t)c * HW));
    float best;
    const float win = tie_window(m1, &best);
    const float lo = __fsub_rn(m1, win);
    int bi = -1, nnear = 0;
    for (int c = 0; c < C; ++c) {
        const bool nr = __ldg(qc + (size_t)c * HW) >= lo;
        if (nr && bi < 0) bi = c;
        nnear += nr ? 1 : 0;
    }
    bi = max(bi, 0);
    if (C > 1 && nnear != 1) best = class_tie_break(qc, HW, C, lo, m1, bi, &bi);   // :198
    const int j = fastdiv(pos, hd.magicW), i = pos - j * hd.W;
    const float sx = sigmoid_fast(tx), sy = sigmoid_fast(ty);      // :187
    const float ew = exp_fast(tw), eh = exp_fast(th);              // :188
    const float cx = __fmul_rn(__fadd_rn(sx, (float)i), hd.rW);    // :194
    const float cy = __fmul_rn(__fadd_rn(sy, (float)j), hd.rH);
    const float bw = __fmul_rn(ew, hd.aw[a]), bh = __fmul_rn(eh, hd.ah[a]);   // :195
    o.bx.x = __fsub_rn(cx, __fmul_rn(bw, 0.5f));                   // :244-247
    o.bx.y = __fsub_rn(cy, __fmul_rn(bh, 0.5f));
    o.bx.z = __fadd_rn(bw, o.bx.x);
    o.bx.w = __fadd_rn(bh, o.bx.y);
    o.best = best;
    o.bi = bi;
    return true;
}

__device__ __forceinline__ unsigned long long large_key(int cls, float sc, uint32_t cid) {
    return ((unsigned long long)cls << 48) | ((unsigned long long)(~float_order_key(sc)) << 16) | (unsigned long long)cid;
}

__device__ __forceinline__ void large_decode_head(const LargeParams &p, int b, const HeadDesc &hd, int cid0,
                                                  unsigned long long *keys, int *cnt) {
    const int cells = p.A * hd.HW;
    const float *hb = hd.ptr + (size_t)b * p.A * p.attrs * hd.HW;
    for (int local = threadIdx.x; local < cells; local += kLargeThreads) {
        const int cid = cid0 + local;
        unsigned long long key = ~0ull;
        LargeCell o;
        if (large_decode_cell(hd, hb, p.C, local, p.conf_thr, o)) {
            float4 *r = p.rec + ((size_t)b * p.K + cid) * 2;
            r[0] = o.bx;
            r[1] = make_float4(o.conf, o.best, make_ta(o.bx, p.iou), __int_as_float(o.bi));
            key = large_key(o.bi, __fmul_rn(o.best, o.conf), (uint32_t)cid);   // box.py:27
            atomicAdd(&cnt[o.bi], 1);
        }
        keys[cid] = key;
    }
}

// SRC = 1: the caller's candidate rows (utils.box.nms, box.py:17: head 0 rows then head 1 rows)
__device__ __forceinline__ const float *large_row(const LargeParams &p, int b, int K0, uint32_t cid) {
    return ((int)cid < K0) ? p.cand[0] + ((size_t)b * p.cand_stride[0] + cid) * 7
                           : p.cand[1] + ((size_t)b * p.cand_stride[1] + (cid - K0)) * 7;
}

__device__ __forceinline__ void large_load_rows(const LargeParams &p, int b, int K0, int Kb, unsigned long long *keys, int *cnt) {
    for (int row = threadIdx.x; row < p.K; row += kLargeThreads) {
        unsigned long long key = ~0ull;
        if (row < Kb) {
            const float *src = large_row(p, b, K0, (uint32_t)row);
            const float conf = __ldg(src + 4), score = __ldg(src + 5), v = __ldg(src + 6);
            const int c = (int)v;  // rows whose class column is not an integer in [0,C) match no `== i` (box.py:21)
            if ((v == (float)c) && c >= 0 && c < p.C) {
                key = large_key(c, __fmul_rn(score, conf), (uint32_t)row);
                atomicAdd(&cnt[c], 1);
            }
        }
        keys[row] = key;
    }
}

// box and t*area*2^-13 of candidate `cid`
template <int SRC>
__device__ __forceinline__ void large_fetch(const LargeParams &p, int b, int K0, const float4 *rec, uint32_t cid, float4 &B, float &ta) {
    if (SRC == 0) {
        B = rec[2 * cid];              // (plain loads: the records were written by this kernel)
        ta = rec[2 * cid + 1].z;
    } else {
        const float *src = large_row(p, b, K0, cid);
        B = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), __ldg(src + 3));
        ta = make_ta(B, p.iou);
    }
}

__device__ __forceinline__ void large_bitonic_sort(unsigned long long *k, int P) {
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < (P >> 1); i += kLargeThreads) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long a = k[lo], b = k[hi];
                if ((a > b) == up) { k[lo] = b; k[hi] = a; }
            }
            __syncthreads();
        }
    }
}

template <int SRC>
__global__ void __launch_bounds__(kLargeThreads, 1) decode_nms_large_kernel(const LargeParams p) {
    extern __shared__ __align__(16) unsigned char lsm[];
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(lsm);
    int *cnt = reinterpret_cast<int *>(keys + p.P);       // [C+2]
    int *start = cnt + (p.C + 2);                         // [C+2]
    int *ktile = start + (p.C + 2);                       // [C+2]
    uint32_t *keptw = reinterpret_cast<uint32_t *>(ktile + (p.C + 2));   // [T]
    int *ready = reinterpret_cast<int *>(keptw + p.T);    // [T]  (P5: exclusive prefix of the kept counts)
    uint32_t *tasks = reinterpret_cast<uint32_t *>(ready + p.T);   // [T] (class << 16 | tile), tile-major
    int *misc = reinterpret_cast<int *>(tasks + p.T);     // [0] task counter, [1] total kept
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = p.C, K = p.K;
    // per-warp staging of one tile of columns for the pair blocks of decode_nms.cuh
    const size_t stage0 = (((size_t)8 * p.P + (size_t)4 * (3 * (C + 2)) + (size_t)12 * p.T + 64) + 15) & ~(size_t)15;
    float4 *sbox = reinterpret_cast<float4 *>(lsm + stage0 + (size_t)warp * kLargeStageBytes);
    H16Tile *stile = reinterpret_cast<H16Tile *>(sbox + 32);

    for (int i = tid; i < C + 2; i += kLargeThreads) cnt[i] = 0;
    for (int i = tid; i < p.T; i += kLargeThreads) { keptw[i] = 0u; ready[i] = 0; }
    for (int i = K + tid; i < p.P; i += kLargeThreads) keys[i] = ~0ull;
    if (tid < 2) misc[tid] = 0;
    __syncthreads();
    // P1
    int K0 = 0;
    if (SRC == 0) {
        large_decode_head(p, b, p.head[0], 0, keys, cnt);
        large_decode_head(p, b, p.head[1], p.head[0].cells, keys, cnt);
    } else {
        K0 = min(p.cand_count[0][b], p.cand_stride[0]);
        const int K1 = p.cand[1] ? min(p.cand_count[1][b], p.cand_stride[1]) : 0;
        large_load_rows(p, b, K0, K0 + K1, keys, cnt);
    }
    __syncthreads();
    // P2
    large_bitonic_sort(keys, p.P);
    // P3: start[c] = first sorted position of class c, ktile[c] = first tile slot of class c; then the task list,
    // tile-major (tile 0 of every class, tile 1 of every class, ...): the classes' chains advance side by side,
    // and a task still depends only on tasks that come before it
    if (warp == 0) {
        int cs = 0, ct = 0, maxt = 0;
        for (int c0 = 0; c0 <= C; c0 += 32) {
            const int c = c0 + lane;
            const int n = (c < C) ? cnt[c] : 0;
            const int nt = (n + 31) >> 5;
            const int in = warp_inclusive_scan(n, lane), it = warp_inclusive_scan(nt, lane);
            if (c <= C) { start[c] = cs + in - n; ktile[c] = ct + it - nt; }
            cs += __shfl_sync(kFullMask, in, 31);
            ct += __shfl_sync(kFullMask, it, 31);
            maxt = max(maxt, nt);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) maxt = max(maxt, __shfl_xor_sync(kFullMask, maxt, o));
        int q = 0;
        for (int t = 0; t < maxt; ++t) {
            for (int c0 = 0; c0 < C; c0 += 32) {
                const int c = c0 + lane;
                const bool has = c < C && ((cnt[c] + 31) >> 5) > t;
                const uint32_t bal = __ballot_sync(kFullMask, has);
                if (has) tasks[q + __popc(bal & lanemask_lt())] = ((uint32_t)c << 16) | (uint32_t)t;
                q += __popc(bal);
            }
        }
    }
    __syncthreads();
    const int ntask = ktile[C];
    const float4 *rec = p.rec + (size_t)b * K * 2;
    // P4
    for (;;) {
        int q = 0;
        if (lane == 0) q = atomicAdd(&misc[0], 1);
        q = __shfl_sync(kFullMask, q, 0);
        if (q >= ntask) break;
        const uint32_t tk = tasks[q];
        const int c = (int)(tk >> 16), t = (int)(tk & 0xffffu);
        const int n = cnt[c], s0 = start[c], g0 = ktile[c];
        const int idx = 32 * t + lane;
        const int ncol = min(32, n - 32 * t);
        bool alive = idx < n;
        // my candidate (lanes past the end of the class take the last one; they are never alive)
        const uint32_t cid = (uint32_t)(keys[s0 + min(idx, n - 1)] & 0xffffull);
        float4 R;
        float rta;
        large_fetch<SRC>(p, b, K0, rec, cid, R, rta);
        // The pair blocks use the conservative fp16 prefilter of decode_nms.cuh (h16_prefilter: a clear bit proves that
        // torchvision does not suppress the pair) and decide the pairs it cannot rule out exactly (pair_decide).
        const float ra = box_area(R);
        const H16Row myrow = tl_h16_row(R, ra, p.iou.th);
        // the exact decisions for this lane's "maybe" columns of the staged tile
        auto resolve = [&](uint32_t maybe) -> uint32_t {
            uint32_t word = 0u;
            while (maybe) {
                const int j = __ffs(maybe) - 1;
                maybe &= maybe - 1u;
                if (pair_decide(R, ra, sbox[j], p.iou)) word |= 1u << j;
            }
            return word;
        };
        // earlier tiles of the class: their 32 boxes are staged as the columns of one pair block (the records were
        // written in P1, so the loads do not wait for the tile's flag; the next tile's are issued before this one
        // is processed); only columns that were KEPT can suppress my candidate
        float4 nB = R;
        float nta = rta;
        if (t > 0) large_fetch<SRC>(p, b, K0, rec, (uint32_t)(keys[s0 + lane] & 0xffffull), nB, nta);
        for (int rt = 0; rt < t; ++rt) {
            __syncwarp();   // (every lane has finished reading the previous tile's columns out of the staging buffer)
            sbox[lane] = nB;
            tl_h16_put(stile, lane, nB, box_area(nB), p.iou.th);
            if (rt + 1 < t) large_fetch<SRC>(p, b, K0, rec, (uint32_t)(keys[s0 + 32 * (rt + 1) + lane] & 0xffffull), nB, nta);
            __syncwarp();
            uint32_t maybe = alive ? h16_prefilter(*stile, myrow, 32) : 0u;
            // (hand-over through atomics with release / acquire semantics: see decode_nms.cuh, ld_acquire_s)
            uint32_t kw = 0u;
            if (lane == 0) {
                while (ld_acquire_s(&ready[g0 + rt]) == 0) __nanosleep(40);
                kw = ld_relaxed_s(&keptw[g0 + rt]);
            }
            kw = __shfl_sync(kFullMask, kw, 0);
            if (resolve(maybe & kw)) alive = false;
            if (!__any_sync(kFullMask, alive)) break;
        }
        // own tile: later columns a candidate would suppress, then the turns in score order
        __syncwarp();
        sbox[lane] = R;
        tl_h16_put(stile, lane, R, ra, p.iou.th);
        __syncwarp();
        const uint32_t validc = (ncol >= 32) ? 0xffffffffu : ((1u << ncol) - 1u);
        uint32_t D = alive ? resolve(h16_prefilter(*stile, myrow, ncol) & validc & ~((2u << lane) - 1u)) : 0u;   // only LATER columns
        uint32_t rem = __ballot_sync(kFullMask, alive);
        uint32_t nz = __ballot_sync(kFullMask, (D & rem) != 0u);
        while (nz) {
            const int i = __ffs(nz) - 1;
            nz &= nz - 1u;
            const uint32_t Di = __shfl_sync(kFullMask, D, i);
            if ((rem >> i) & 1u) rem &= ~Di;
        }
        __syncwarp();
        if (lane == 0) {
            st_relaxed_s(&keptw[g0 + t], rem);
            st_release_s(&ready[g0 + t], 1);
        }
    }
    __syncthreads();
    // P5: exclusive prefix of the kept counts over the tile slots (reuses ready[]), then the rows
    if (warp == 0) {
        int carry = 0;
        for (int g = 0; g < ntask; g += 32) {
            const int v = (g + lane < ntask) ? __popc(keptw[g + lane]) : 0;
            const int inc = warp_inclusive_scan(v, lane);
            if (g + lane < ntask) ready[g + lane] = carry + inc - v;
            carry += __shfl_sync(kFullMask, inc, 31);
        }
        if (lane == 0) misc[1] = carry;
    }
    __syncthreads();
    float *o = p.out + (size_t)b * K * 7;
    for (int g = warp; g < ntask; g += kLargeThreads / 32) {
        const uint32_t word = keptw[g];
        if (!((word >> lane) & 1u)) continue;
        int lo = 0, hi = C - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (ktile[mid] <= g) lo = mid; else hi = mid - 1;
        }
        const int c = lo;
        const uint32_t cid = (uint32_t)(keys[start[c] + 32 * (g - ktile[c]) + lane] & 0xffffull);
        const int r = ready[g] + __popc(word & lanemask_lt());
        float *d = o + (size_t)r * 7;
        if (SRC == 0) {
            const float4 bx = rec[2 * cid], cs = rec[2 * cid + 1];
            d[0] = bx.x; d[1] = bx.y; d[2] = bx.z; d[3] = bx.w;
            d[4] = cs.x; d[5] = cs.y; d[6] = (float)c;              // cls_idx.float() yolo_loss.py:199
        } else {                                                    // the caller's own rows, bit for bit (box.py:29)
            const float *src = large_row(p, b, K0, cid);
#pragma unroll
            for (int k = 0; k < 7; ++k) d[k] = __ldg(src + k);
        }
        if (p.out_idx) p.out_idx[(size_t)b * K + r] = (int)cid;
    }
    if (tid == 0) p.out_count[b] = misc[1];
}

// YOLOLoss.forward(input) for a head with more cells than the stand-alone decode kernel stages (e.g. 52x52x3 = 8112):
// CTA per image, 1024 cells per round, rows written straight from registers in candidate order (a, j, i) (:203).
struct LargeDecodeParams {
    HeadDesc head;
    int N, A, C, K;
    float conf_thr;
    float *rows;   // [N][K][7]
    int *count;    // [N]
    int *ids;      // [N][K] or null
};

__global__ void __launch_bounds__(kLargeThreads, 1) decode_head_large_kernel(const LargeDecodeParams p) {
    __shared__ int s_warp[kLargeThreads / 32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cells = p.K;
    const float *hb = p.head.ptr + (size_t)b * p.A * (5 + p.C) * p.head.HW;
    float *o = p.rows + (size_t)b * p.K * 7;
    int running = 0;
    for (int base = 0; base < cells; base += kLargeThreads) {
        const int local = base + tid;
        LargeCell c;
        const bool pass = local < cells && large_decode_cell(p.head, hb, p.C, local, p.conf_thr, c);
        const uint32_t bal = __ballot_sync(kFullMask, pass);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < kLargeThreads / 32; ++w) {
            const int v = s_warp[w];
            before += (w < warp) ? v : 0;
            total += v;
        }
        if (pass) {
            const int r = running + before + __popc(bal & lanemask_lt());
            float *d = o + (size_t)r * 7;
            d[0] = c.bx.x; d[1] = c.bx.y; d[2] = c.bx.z; d[3] = c.bx.w;
            d[4] = c.conf; d[5] = c.best; d[6] = (float)c.bi;       // :199
            if (p.ids) p.ids[(size_t)b * p.K + r] = local;
        }
        running += total;
        __syncthreads();
    }
    if (tid == 0) p.count[b] = running;
}

}  // namespace b200yolo
