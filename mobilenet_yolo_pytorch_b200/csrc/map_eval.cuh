// map_eval.cuh -- utils/eval_mAP.py::calculate_mAP on the device: the consumer of
// the NMS output in train.py:test() (:367-421).  The reference walks every (class, image, detection) in Python
// and calls find_jaccard_overlap once per detection.
//
//   map_match_kernel   CTA per image, thread per class: the greedy matching of eval_single_image_recall
//                      (:8-63) in the detections' given order -- max IoU over the image's class-c objects (first
//                      maximum, NaN -> no match), > 0.5 and not difficult and not yet detected -> TP, else FP
//                      (matched-difficult: neither) -- plus the per-(class, image) detection counts and the
//                      number of non-difficult objects per class.
//   map_class_kernel   CTA per class: eval_class_ap (:65-132): gather the class's (score, tp, fp) over all images,
//                      sort by score descending (bitonic, 64-bit keys; ties by detection index), cumulative sums,
//                      precision = ctp/(ctp+cfp+1e-10), recall = ctp/n_easy, 11-point interpolated AP.
//
// IoU uses the reference's IEEE operations (utils/iou.py), so every TP/FP decision is bit-exact.
#pragma once
#include "common.cuh"

namespace b200yolo {

constexpr int kMapMatchThreads = 128;
constexpr int kMapClassThreads = 1024;
constexpr int kMapSmemKeys = 4096;   // segments up to this size are sorted in shared memory
constexpr int kMapMaxThr = 16;

struct MapParams {
    const float *det_boxes;    // [D][4]
    const int *det_labels;     // [D] 1..n_classes-1
    const float *det_scores;   // [D]
    const int *det_off;        // [N+1]
    const float *true_boxes;   // [T][4]
    const int *true_labels;    // [T]
    const unsigned char *true_difficult;  // [T]
    const int *true_off;       // [N+1]
    int N, Cf;                 // images, foreground classes (n_classes - 1)
    float iou_thr;
    int n_thr;
    float thr[kMapMaxThr];     // recall thresholds (torch.arange(0, 1.1, .1), eval_mAP.py:120)
    // workspace
    unsigned char *detected;   // [T]
    unsigned char *flags;      // [D] bit 0 tp, bit 1 fp
    int *M;                    // [Cf] detections per class
    int *n_easy;               // [Cf]
    int *cnt;                  // [Cf][N] detections per (class, image); becomes the offset inside the class segment
    unsigned long long *keys;  // [2 D + Cf]
    // outputs
    float *ap, *tp_sum, *fp_sum;  // [Cf]
};

__global__ void __launch_bounds__(kMapMatchThreads) map_match_kernel(const MapParams p) {
    const int b = blockIdx.x;
    const int d0 = p.det_off[b], d1 = p.det_off[b + 1];
    const int t0 = p.true_off[b], t1 = p.true_off[b + 1];
    for (int c = threadIdx.x; c < p.Cf; c += kMapMatchThreads) {
        const int label = c + 1;
        int n_easy = 0, n_obj = 0;
        for (int t = t0; t < t1; ++t)
            if (__ldg(p.true_labels + t) == label) { ++n_obj; n_easy += 1 - (int)__ldg(p.true_difficult + t); }  // :17
        int n_det = 0;
        for (int d = d0; d < d1; ++d) {
            if (__ldg(p.det_labels + d) != label) continue;
            ++n_det;
            unsigned char fl = 0;
            if (n_obj == 0) {
                fl = 2;                                                    // :37-39 false positive
            } else {
                const float4 db = __ldg(reinterpret_cast<const float4 *>(p.det_boxes) + d);
                const float da = box_area(db);
                float best = 0.f;
                int ind = -1;
                bool nan = false;
                for (int t = t0; t < t1; ++t) {
                    if (__ldg(p.true_labels + t) != label) continue;
                    const float4 tb = __ldg(reinterpret_cast<const float4 *>(p.true_boxes) + t);
                    const float inter = pair_inter(db, tb);                // utils/iou.py:4-13
                    const float v = __fdiv_rn(inter, pair_union(da, box_area(tb), inter));  // :32-49
                    if (v != v) nan = true;                                // torch.max propagates NaN; NaN > 0.5 is False
                    if (ind < 0 || v > best) { best = v; ind = t; }        // first maximum (:42)
                }
                if (!nan && best > p.iou_thr) {                            // :51
                    if (__ldg(p.true_difficult + ind) == 0) {              // :53
                        if (p.detected[ind] == 0) { fl = 1; p.detected[ind] = 1; }  // :55-57 true positive
                        else fl = 2;                                       // :59-60
                    }
                } else {
                    fl = 2;                                                // :61-62
                }
            }
            p.flags[d] = fl;
        }
        p.cnt[(size_t)c * p.N + b] = n_det;
        if (n_det) atomicAdd(p.M + c, n_det);
        if (n_easy) atomicAdd(p.n_easy + c, n_easy);
    }
}

__device__ __forceinline__ unsigned next_pow2(unsigned v) {
    unsigned p = 1;
    while (p < v) p <<= 1;
    return p;
}

__device__ __forceinline__ void bitonic_sort(unsigned long long *k, int P) {
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < (P >> 1); i += kMapClassThreads) {
                const int lo = 2 * i - (i & (stride - 1));  // index with bit `stride` clear
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long a = k[lo], b = k[hi];
                if ((a > b) == up) { k[lo] = b; k[hi] = a; }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(kMapClassThreads) map_class_kernel(const MapParams p) {
    __shared__ unsigned long long s_keys[kMapSmemKeys];
    __shared__ int s_warp[2][32];
    __shared__ int s_carry[2];
    __shared__ int s_prec[kMapMaxThr];
    __shared__ int s_base;
    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int label = c + 1;
    const int M = p.M[c];
    const int P = (int)next_pow2((unsigned)max(M, 1));
    if (tid == 0) {
        long long base = 0;
        for (int cc = 0; cc < c; ++cc) base += next_pow2((unsigned)max(p.M[cc], 1));
        s_base = (int)base;
        s_carry[0] = 0;
        s_carry[1] = 0;
    }
    if (tid < kMapMaxThr) s_prec[tid] = 0;
    __syncthreads();
    unsigned long long *seg = p.keys + s_base;
    unsigned long long *k = (P <= kMapSmemKeys) ? s_keys : seg;

    // (b) offset of every image inside the class segment: exclusive scan of cnt[c][*] in place
    int *cnt = p.cnt + (size_t)c * p.N;
    for (int b0 = 0; b0 < p.N; b0 += kMapClassThreads) {
        const int b = b0 + tid;
        const int v = (b < p.N) ? cnt[b] : 0;
        const int inc = warp_inclusive_scan(v, lane);
        if (lane == 31) s_warp[0][warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const int w = s_warp[0][lane];
            const int winc = warp_inclusive_scan(w, lane);
            s_warp[0][lane] = winc - w;
        }
        __syncthreads();
        const int carry = s_carry[0];
        if (b < p.N) cnt[b] = carry + s_warp[0][warp] + inc - v;
        __syncthreads();
        if (tid == kMapClassThreads - 1) s_carry[0] = carry + s_warp[0][warp] + inc;
        __syncthreads();
    }
    // (c) keys: ascending key order = descending score, then ascending detection index
    for (int i = tid; i < P; i += kMapClassThreads) k[i] = ~0ull;
    __syncthreads();
    for (int b = tid; b < p.N; b += kMapClassThreads) {
        int o = cnt[b];
        for (int d = p.det_off[b]; d < p.det_off[b + 1]; ++d) {
            if (__ldg(p.det_labels + d) != label) continue;
            const unsigned sk = ~float_order_key(__ldg(p.det_scores + d));
            k[o++] = ((unsigned long long)sk << 32) | ((unsigned long long)(unsigned)d << 2) | (unsigned long long)p.flags[d];
        }
    }
    __syncthreads();
    // (d) sort
    bitonic_sort(k, P);
    // (e) cumulative tp / fp in score order, precision / recall, running max of the precision per recall threshold
    const float n_easy = (float)p.n_easy[c];
    if (tid == 0) { s_carry[0] = 0; s_carry[1] = 0; }
    __syncthreads();
    for (int i0 = 0; i0 < M; i0 += kMapClassThreads) {
        const int i = i0 + tid;
        const unsigned fl = (i < M) ? (unsigned)(k[i] & 3ull) : 0u;
        const int vt = (int)(fl & 1u), vf = (int)((fl >> 1) & 1u);
        const int it = warp_inclusive_scan(vt, lane), jf = warp_inclusive_scan(vf, lane);
        if (lane == 31) { s_warp[0][warp] = it; s_warp[1][warp] = jf; }
        __syncthreads();
        if (warp == 0) {
            const int w0 = s_warp[0][lane], w1 = s_warp[1][lane];
            const int a0 = warp_inclusive_scan(w0, lane), a1 = warp_inclusive_scan(w1, lane);
            s_warp[0][lane] = a0 - w0;
            s_warp[1][lane] = a1 - w1;
        }
        __syncthreads();
        const int ct = s_carry[0] + s_warp[0][warp] + it;   // cumulative true positives up to and including i (:113)
        const int cf = s_carry[1] + s_warp[1][warp] + jf;   // cumulative false positives (:114)
        if (i < M) {
            const float ctp = (float)ct, cfp = (float)cf;
            const float prec = __fdiv_rn(ctp, __fadd_rn(__fadd_rn(ctp, cfp), 1e-10f));   // :115-116
            const float rec = __fdiv_rn(ctp, n_easy);                                      // :117 (x/0 -> inf / NaN like torch)
            for (int q = 0; q < p.n_thr; ++q)
                if (rec >= p.thr[q]) atomicMax(&s_prec[q], __float_as_int(prec));          // :121-126 (prec >= 0)
        }
        __syncthreads();
        if (tid == kMapClassThreads - 1) { s_carry[0] = ct; s_carry[1] = cf; }
        __syncthreads();
    }
    if (tid == 0) {
        float sum = 0.f;
        for (int q = 0; q < p.n_thr; ++q) sum = __fadd_rn(sum, __int_as_float(s_prec[q]));
        p.ap[c] = __fdiv_rn(sum, (float)p.n_thr);                                          // precisions.mean() (:128)
        p.tp_sum[c] = (float)s_carry[0];
        p.fp_sum[c] = (float)s_carry[1];
    }
}

// ---------------------------------------------------------------------------
// Compaction of fixed-stride detections (N, K, 7) + counts into packed rows + offsets: the send buffer of the
// data-parallel all-gather carries only kept rows (dist.all_gather_detections_compact).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) compact_offsets_kernel(const int *count, int N, int K, int *offsets) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < N; b0 += 1024) {
        const int b = b0 + tid;
        const int v = (b < N) ? min(max(count[b], 0), K) : 0;
        const int inc = warp_inclusive_scan(v, lane);
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const int w = s_warp[lane];
            const int winc = warp_inclusive_scan(w, lane);
            s_warp[lane] = winc - w;
        }
        __syncthreads();
        const int carry = s_carry;
        if (b < N) offsets[b] = carry + s_warp[warp] + inc - v;
        __syncthreads();
        if (tid == 1023) s_carry = carry + s_warp[warp] + inc;
        __syncthreads();
    }
    if (tid == 0) offsets[N] = s_carry;
}

__global__ void __launch_bounds__(128) compact_rows_kernel(const float *dets, const int *count, const int *offsets, int K,
                                                           float *packed) {
    const int b = blockIdx.x;
    const int n = min(max(count[b], 0), K) * 7;
    const float *src = dets + (size_t)b * K * 7;
    float *dst = packed + (size_t)offsets[b] * 7;
    for (int f = threadIdx.x; f < n; f += 128) dst[f] = __ldg(src + f);
}

}  // namespace b200yolo
