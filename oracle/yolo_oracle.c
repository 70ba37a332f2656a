/*
 * oracle/yolo_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, CPU restatement of the detection hot path of
 * eric612/Mobilenet-YOLO-Pytorch: YOLO-head decode, confidence threshold +
 * per-class greedy IoU NMS, pairwise IoU, and the YOLOLoss target assignment
 * and loss.  It exists only to check the CUDA path (tests/, smoke(), and the
 * cpu_baseline / --impl reference legs of bench.py).  Nothing under
 * mobilenet_yolo_pytorch_b200/ may import, link or call it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function in
 * this file against fixtures under tests/golden/ that were produced by running
 * the reference's own Python (models/yolo_loss.py, utils/box.py, utils/iou.py,
 * torchvision.ops.nms) in the build container (tests/golden/make_golden.py).
 *
 * All reference citations are relative to /root/reference/.
 * Compile with:  gcc -O2 -fno-fast-math -ffp-contract=off -fopenmp -shared -fPIC
 * (no FMA contraction: the reference evaluates every op as a separate fp32
 * rounding).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

static int g_threads = 0; /* 0 = OpenMP default */

ORACLE_API void oracle_set_threads(int n) { g_threads = n; }
ORACLE_API int oracle_get_max_threads(void) {
#ifdef _OPENMP
    return g_threads > 0 ? g_threads : omp_get_max_threads();
#else
    return 1;
#endif
}
static int nthreads(void) { return oracle_get_max_threads(); }

/* yolo_loss.py:19 (training flavour 1/(1+exp(-x))); yolo_loss.py:187-189 uses
 * torch.sigmoid, which differs from this by <= 1 ulp -- float parity tolerance
 * is 1e-5 relative, see tests. */
static inline float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

/* ------------------------------------------------------------------------ */
/* A.1  decode of one cell (yolo_loss.py:186-199 eval, 84-92 train)          */
/* ------------------------------------------------------------------------ */
/* element (b,a,t,j,i) of a (N, A*attrs, H, W) head: yolo_loss.py:84,186 */
static inline size_t head_off(int b, int a, int t, int j, int i, int A, int attrs, int H, int W) {
    return ((((size_t)b * A + a) * attrs + t) * H + j) * (size_t)W + i;
}

static inline void decode_box(const float *x, int b, int a, int j, int i, int A, int attrs, int H, int W,
                              float aw, float ah, float box[4]) {
    float tx = x[head_off(b, a, 0, j, i, A, attrs, H, W)];
    float ty = x[head_off(b, a, 1, j, i, A, attrs, H, W)];
    float tw = x[head_off(b, a, 2, j, i, A, attrs, H, W)];
    float th = x[head_off(b, a, 3, j, i, A, attrs, H, W)];
    float sx = sigmoid_f(tx), sy = sigmoid_f(ty);          /* :187 / :85 */
    float ew = expf(tw), eh = expf(th);                    /* :188 / :86 */
    float cx = (sx + (float)i) / (float)W;                 /* :194 / :90, grid from pre_maps :71-72 */
    float cy = (sy + (float)j) / (float)H;
    float bw = ew * aw;                                    /* :195 / :91 */
    float bh = eh * ah;
    float x1 = cx - bw / 2.0f;                             /* wh_to_x2y2 :244 */
    float y1 = cy - bh / 2.0f;                             /* :245 */
    float x2 = bw + x1;                                    /* :246 (NOT cx + bw/2) */
    float y2 = bh + y1;                                    /* :247 */
    box[0] = x1; box[1] = y1; box[2] = x2; box[3] = y2;
}

/*
 * YOLOLoss.get_pred_boxes (yolo_loss.py:180-204) for one head.
 *   x          (N, A*(5+C), H, W) fp32 contiguous
 *   anchor_wh  [A][2]  this head's anchors already divided by img_size
 *              (yolo_loss.py:214, pre_maps :66-70), rounded to fp32
 *   rows       [N][A*H*W][7]  rows that pass, in (a,j,i) row-major order (:203)
 *   count      [N]
 *   ids        [N][A*H*W] cell id (a*H+j)*W+i of each emitted row (may be NULL)
 */
ORACLE_API void oracle_decode_head(const float *x, int N, int A, int C, int H, int W, const float *anchor_wh,
                                   float conf_thr, float *rows, int *count, int *ids) {
    const int attrs = 5 + C;
    const int cells = A * H * W;
#pragma omp parallel for schedule(static) num_threads(nthreads())
    for (int b = 0; b < N; ++b) {
        int n = 0;
        float *out = rows + (size_t)b * cells * 7;
        for (int a = 0; a < A; ++a)
            for (int j = 0; j < H; ++j)
                for (int i = 0; i < W; ++i) {
                    float conf = sigmoid_f(x[head_off(b, a, 4, j, i, A, attrs, H, W)]); /* :189,197 */
                    if (!(conf > conf_thr)) continue;                                  /* :201 */
                    float box[4];
                    decode_box(x, b, a, j, i, A, attrs, H, W, anchor_wh[2 * a], anchor_wh[2 * a + 1], box);
                    float best = sigmoid_f(x[head_off(b, a, 5, j, i, A, attrs, H, W)]);
                    int bi = 0;
                    for (int c = 1; c < C; ++c) { /* torch.max(dim) :198 -> first maximum */
                        float s = sigmoid_f(x[head_off(b, a, 5 + c, j, i, A, attrs, H, W)]);
                        if (s > best) { best = s; bi = c; }
                    }
                    float *r = out + (size_t)n * 7;
                    r[0] = box[0]; r[1] = box[1]; r[2] = box[2]; r[3] = box[3];
                    r[4] = conf; r[5] = best; r[6] = (float)bi;                          /* :199 */
                    if (ids) ids[(size_t)b * cells + n] = (a * H + j) * W + i;
                    ++n;
                }
        count[b] = n;
    }
}

/* ------------------------------------------------------------------------ */
/* A.2  NMS: utils/box.py:11-31 driving torchvision.ops.nms                   */
/* ------------------------------------------------------------------------ */
typedef struct { float s; int idx; } sort_item;

/* stable descending sort by score (torchvision nms_kernel: scores.sort(stable,
 * descending)); ties keep candidate order; NaN sorts first like torch. */
static int cmp_desc(const void *pa, const void *pb) {
    const sort_item *a = (const sort_item *)pa, *b = (const sort_item *)pb;
    int an = isnan(a->s), bn = isnan(b->s);
    if (an != bn) return an ? -1 : 1;
    if (!an) {
        if (a->s > b->s) return -1;
        if (a->s < b->s) return 1;
    }
    return (a->idx > b->idx) - (a->idx < b->idx);
}

/* greedy NMS over `n` rows selected by sel[] (indices into rows, 7 floats per
 * row), torchvision/csrc/ops/cpu/nms_kernel.cpp semantics: area=(x2-x1)*(y2-y1),
 * inter=max(0,..)*max(0,..), ovr = inter/((ai+aj)-inter), suppress iff
 * (double)ovr > thr.  Returns kept count; kept[] holds row indices in score
 * order. */
static int nms_one_class(const float *rows, const int *sel, int n, double thr, sort_item *items, unsigned char *supp,
                         int *kept) {
    for (int k = 0; k < n; ++k) {
        const float *r = rows + (size_t)sel[k] * 7;
        items[k].s = r[5] * r[4]; /* box.py:27 scores = col5*col4 */
        items[k].idx = k;
        supp[k] = 0;
    }
    qsort(items, (size_t)n, sizeof(sort_item), cmp_desc);
    int nk = 0;
    for (int _i = 0; _i < n; ++_i) {
        int i = items[_i].idx;
        if (supp[i]) continue;
        kept[nk++] = sel[i];
        const float *bi = rows + (size_t)sel[i] * 7;
        float ix1 = bi[0], iy1 = bi[1], ix2 = bi[2], iy2 = bi[3];
        float iarea = (ix2 - ix1) * (iy2 - iy1);
        for (int _j = _i + 1; _j < n; ++_j) {
            int j = items[_j].idx;
            if (supp[j]) continue;
            const float *bj = rows + (size_t)sel[j] * 7;
            float xx1 = fmaxf(ix1, bj[0]), yy1 = fmaxf(iy1, bj[1]);
            float xx2 = fminf(ix2, bj[2]), yy2 = fminf(iy2, bj[3]);
            float w = fmaxf(0.0f, xx2 - xx1), h = fmaxf(0.0f, yy2 - yy1);
            float inter = w * h;
            float jarea = (bj[2] - bj[0]) * (bj[3] - bj[1]);
            float ovr = inter / (iarea + jarea - inter);
            if ((double)ovr > thr) supp[j] = 1;
        }
    }
    return nk;
}

/*
 * utils.box.nms (box.py:11-31) on already-concatenated per-image candidates.
 *   rows      [N][Kstride][7], count[N]: candidates (head0 then head1, box.py:17)
 *   out       [N][Kstride][7], out_count[N]: class-ascending blocks, each in
 *             descending score order (box.py:20-29)
 *   out_idx   [N][Kstride] index (into the image's candidate list) of each
 *             output row; may be NULL
 */
ORACLE_API void oracle_nms(const float *rows, const int *count, int N, int Kstride, int C, double iou_thr, float *out,
                           int *out_count, int *out_idx) {
#pragma omp parallel num_threads(nthreads())
    {
        int *sel = (int *)malloc(sizeof(int) * (size_t)(Kstride > 0 ? Kstride : 1));
        int *kept = (int *)malloc(sizeof(int) * (size_t)(Kstride > 0 ? Kstride : 1));
        sort_item *items = (sort_item *)malloc(sizeof(sort_item) * (size_t)(Kstride > 0 ? Kstride : 1));
        unsigned char *supp = (unsigned char *)malloc((size_t)(Kstride > 0 ? Kstride : 1));
#pragma omp for schedule(dynamic, 1)
        for (int b = 0; b < N; ++b) {
            const float *r = rows + (size_t)b * Kstride * 7;
            float *o = out + (size_t)b * Kstride * 7;
            int n = count[b], no = 0;
            for (int c = 0; c < C; ++c) { /* box.py:20 */
                int ns = 0;
                for (int k = 0; k < n; ++k)
                    if (r[(size_t)k * 7 + 6] == (float)c) sel[ns++] = k; /* box.py:21-22 */
                if (!ns) continue;                                        /* box.py:24 */
                int nk = nms_one_class(r, sel, ns, iou_thr, items, supp, kept);
                for (int k = 0; k < nk; ++k) { /* box.py:29 */
                    memcpy(o + (size_t)no * 7, r + (size_t)kept[k] * 7, 7 * sizeof(float));
                    if (out_idx) out_idx[(size_t)b * Kstride + no] = kept[k];
                    ++no;
                }
            }
            out_count[b] = no;
        }
        free(sel); free(kept); free(items); free(supp);
    }
}

/*
 * Whole inference post-process for a two-head detector, as driven by
 * mbv2_yolo.py:158-160: get_pred_boxes on each head, then utils.box.nms.
 * Kmax = A*H0*W0 + A*H1*W1.  cand/cand_count/cand_ids are optional outputs of
 * the intermediate candidate list (ids are global: head1 ids are offset by
 * A*H0*W0).
 */
ORACLE_API void oracle_decode_nms(const float *head0, const float *head1, int N, int A, int C, int H0, int W0, int H1,
                                  int W1, const float *anchor_wh /*[2][A][2]*/, float conf_thr, double iou_thr,
                                  float *out, int *out_count, int *out_idx, float *cand, int *cand_count,
                                  int *cand_ids) {
    const int c0 = A * H0 * W0, c1 = A * H1 * W1, K = c0 + c1;
    float *r0 = (float *)malloc(sizeof(float) * 7 * (size_t)N * c0);
    float *r1 = (float *)malloc(sizeof(float) * 7 * (size_t)N * c1);
    int *n0 = (int *)malloc(sizeof(int) * (size_t)N), *n1 = (int *)malloc(sizeof(int) * (size_t)N);
    int *i0 = (int *)malloc(sizeof(int) * (size_t)N * c0), *i1 = (int *)malloc(sizeof(int) * (size_t)N * c1);
    float *cat = cand ? cand : (float *)malloc(sizeof(float) * 7 * (size_t)N * K);
    int *ncat = cand_count ? cand_count : (int *)malloc(sizeof(int) * (size_t)N);
    oracle_decode_head(head0, N, A, C, H0, W0, anchor_wh, conf_thr, r0, n0, i0);
    oracle_decode_head(head1, N, A, C, H1, W1, anchor_wh + 2 * A, conf_thr, r1, n1, i1);
    for (int b = 0; b < N; ++b) { /* box.py:17 torch.cat((preds[0][b], preds[1][b])) */
        memcpy(cat + (size_t)b * K * 7, r0 + (size_t)b * c0 * 7, sizeof(float) * 7 * (size_t)n0[b]);
        memcpy(cat + ((size_t)b * K + n0[b]) * 7, r1 + (size_t)b * c1 * 7, sizeof(float) * 7 * (size_t)n1[b]);
        ncat[b] = n0[b] + n1[b];
        if (cand_ids) {
            for (int k = 0; k < n0[b]; ++k) cand_ids[(size_t)b * K + k] = i0[(size_t)b * c0 + k];
            for (int k = 0; k < n1[b]; ++k) cand_ids[(size_t)b * K + n0[b] + k] = c0 + i1[(size_t)b * c1 + k];
        }
    }
    oracle_nms(cat, ncat, N, K, C, iou_thr, out, out_count, out_idx);
    free(r0); free(r1); free(n0); free(n1); free(i0); free(i1);
    if (!cand) free(cat);
    if (!cand_count) free(ncat);
}

/* ------------------------------------------------------------------------ */
/* A8  utils/iou.py                                                          */
/* ------------------------------------------------------------------------ */
static inline float inter_f(const float *a, const float *b) { /* iou.py:4-13 */
    float lx = fmaxf(a[0], b[0]), ly = fmaxf(a[1], b[1]);
    float ux = fminf(a[2], b[2]), uy = fminf(a[3], b[3]);
    float dx = ux - lx, dy = uy - ly;
    /* torch.clamp(min=0) propagates NaN; fmaxf would drop it */
    dx = (dx < 0.0f) ? 0.0f : dx;
    dy = (dy < 0.0f) ? 0.0f : dy;
    return dx * dy;
}
static inline float area_f(const float *a) { return (a[2] - a[0]) * (a[3] - a[1]); } /* iou.py:39-40 */
static inline float union_f(const float *a, const float *b, float inter) {          /* iou.py:44 */
    return area_f(a) + area_f(b) - inter;
}
static inline float iou_f(const float *a, const float *b) { /* iou.py:32-49 */
    float in = inter_f(a, b);
    return in / union_f(a, b, in);
}

/* mode 0: find_intersection, 1: find_union, 2: find_jaccard_overlap; out (n1,n2) */
ORACLE_API void oracle_pairwise(const float *s1, int n1, const float *s2, int n2, int mode, float *out) {
#pragma omp parallel for schedule(static) num_threads(nthreads())
    for (int p = 0; p < n1; ++p)
        for (int q = 0; q < n2; ++q) {
            const float *a = s1 + 4 * (size_t)p, *b = s2 + 4 * (size_t)q;
            float in = inter_f(a, b);
            float v = in;
            if (mode == 1) v = union_f(a, b, in);
            else if (mode == 2) v = in / union_f(a, b, in);
            out[(size_t)p * n2 + q] = v;
        }
}

/* ------------------------------------------------------------------------ */
/* A14  box_ciou / box_giou  (yolo_loss.py:249-317)                           */
/* ------------------------------------------------------------------------ */
/* box1 = gt, box2 = pred, both xyxy; returns v = iou - term in out[0], iou in out[1] */
ORACLE_API void oracle_box_ciou(const float *b1, const float *b2, float *out) {
    float l = fminf(b1[0], b2[0]), t = fminf(b1[1], b2[1]);  /* box_c :250-253 */
    float r = fmaxf(b1[2], b2[2]), bt = fmaxf(b1[3], b2[3]);
    float c = (r - l) * (bt - t);                              /* :264 (an AREA) */
    float iou = iou_f(b1, b2);                                 /* :265 */
    float w1 = b1[2] - b1[0], h1 = b1[3] - b1[1];              /* :267 */
    float w2 = b2[2] - b2[0], h2 = b2[3] - b2[1];              /* :268 */
    float x1 = (b1[2] + b1[0]) / 2.0f, y1 = (b1[1] + b1[3]) / 2.0f; /* :269 */
    float x2 = (b2[2] + b2[0]) / 2.0f, y2 = (b2[1] + b2[3]) / 2.0f; /* :270 */
    float u = (x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2);   /* :272 */
    float d = u / c;                                           /* :277 */
    float ar_gt = w2 / h2, ar_pred = w1 / h1;                  /* :279-280 (names as in the reference) */
    float k = (float)(4.0 / (M_PI * M_PI));                    /* python double -> f32 scalar */
    float dl = atanf(ar_gt) - atanf(ar_pred);
    float ar_loss = (k * dl) * dl;                             /* :282 left-to-right */
    float alpha = ar_loss / (((1.0f - iou) + ar_loss) + 0.000001f); /* :283 */
    float term = d + alpha * ar_loss;                          /* :284 */
    float m = (c == 0.0f) ? 1.0f : 0.0f;                       /* :286 */
    term = term * (1.0f - m) + iou * m;                        /* :287 (bool tensors promote to 0/1) */
    out[0] = iou - term;                                       /* :293 */
    out[1] = iou;
}

ORACLE_API void oracle_box_giou(const float *b1, const float *b2, float *out) { /* :295-317 (dead code upstream) */
    float l = fminf(b1[0], b2[0]), t = fminf(b1[1], b2[1]);
    float r = fmaxf(b1[2], b2[2]), bt = fmaxf(b1[3], b2[3]);
    float c = (r - l) * (bt - t);
    float in = inter_f(b1, b2);
    float u = union_f(b1, b2, in);
    float iou = in / u;
    float term = (c - u) / c;
    float m = (c == 0.0f) ? 1.0f : 0.0f;
    term = term * (1.0f - m) + iou * m;
    out[0] = iou - term;
    out[1] = iou;
}

/* ------------------------------------------------------------------------ */
/* A10-A16  YOLOLoss.get_target + loss (yolo_loss.py:77-178, 206-236)         */
/* ------------------------------------------------------------------------ */
/*
 * x            (N, A*(5+C), H, W)
 * anchors_all  [NA][2] ALL anchors divided by img_size (yolo_loss.py:214), fp32
 * mask         [A] anchor indices of this head
 * gt           [G][5] rows [cls(1-based), cx, cy, w, h] (folder2lmdb.py:145-151),
 *              gt_off [N+1] offsets of each image's rows
 * scalars (double[16]) out:
 *   0 loss (L_dense + iou_weighting*L_iou)   1 recall   2 avg_iou   3 obj
 *   4 no_obj   5 cls   6 count/N            (the 7-tuple of :236)
 *   7 L_dense  8 L_iou  9 sum_w  10 n_assign  11 sum_(o-t)^2 w  12 sum_i (v_i-1)^2
 *   13 sum_i (2-area_i) (informational: cancels, quirk Q6)  14 sum_conf_all  15 n_recall
 * assign       [max_assign][6] = (b, t, k, gj, gi, best_n) in reference loop
 *              order; terms [max_assign][3] = (v, iou, weight); both optional
 * targets/weights  optional (N,A,H,W,C+1) dense tensors exactly as get_target
 *              returns them (:178)
 * returns number of assignments, or -1 if a GT cell index is out of range (the
 * reference raises IndexError / wraps negatives there).
 */
ORACLE_API int oracle_target_loss(const float *x, int N, int A, int C, int H, int W, const float *anchors_all, int NA,
                                  const int *mask, const float *gt, const int *gt_off, float ignore_thr,
                                  float iou_thr, float iou_weighting, double *scalars, int *assign, float *terms,
                                  int max_assign, float *targets, float *weights) {
    const int attrs = 5 + C;
    const int cells = A * H * W;
    const int CH = C + 1;
    const size_t tot = (size_t)N * cells * CH;
    float *outp = (float *)malloc(sizeof(float) * tot);       /* output = sigmoid(pred[...,4:]) :87 */
    float *tgt = targets ? targets : (float *)malloc(sizeof(float) * tot);
    float *wgt = weights ? weights : (float *)malloc(sizeof(float) * tot);
    float *pbox = (float *)malloc(sizeof(float) * 4 * (size_t)N * cells);
    memset(wgt, 0, sizeof(float) * tot);                      /* :82 */

#pragma omp parallel for schedule(static) num_threads(nthreads())
    for (int b = 0; b < N; ++b)
        for (int a = 0; a < A; ++a)
            for (int j = 0; j < H; ++j)
                for (int i = 0; i < W; ++i) {
                    size_t cell = (size_t)b * cells + ((size_t)a * H + j) * W + i;
                    decode_box(x, b, a, j, i, A, attrs, H, W, anchors_all[2 * mask[a]], anchors_all[2 * mask[a] + 1],
                               pbox + 4 * cell);
                    for (int c = 0; c < CH; ++c) {
                        float s = sigmoid_f(x[head_off(b, a, 4 + c, j, i, A, attrs, H, W)]);
                        outp[cell * CH + c] = s;
                        tgt[cell * CH + c] = s;               /* targets = output.clone() :97 */
                    }
                }

    double sum_conf = 0.0;                                    /* no_obj = sum(output[...,0]) :98 */
    for (size_t cell = 0; cell < (size_t)N * cells; ++cell) sum_conf += outp[cell * CH];
    const double no_cnt = (double)N * cells;                  /* :99 */

    /* ignore / no-object mask, :107-125 */
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads())
    for (int b = 0; b < N; ++b) {
        int g0 = gt_off[b], n = gt_off[b + 1] - gt_off[b];
        if (n == 0) {                                         /* :108-111 */
            for (int p = 0; p < cells; ++p) {
                size_t cell = (size_t)b * cells + p;
                wgt[cell * CH] = 1.0f;
                tgt[cell * CH] = 0.0f;
            }
            continue;
        }
        float *gx = (float *)malloc(sizeof(float) * 4 * (size_t)n);
        for (int t = 0; t < n; ++t) {                         /* :112-113 wh_to_x2y2 */
            const float *g = gt + 5 * (size_t)(g0 + t);
            float x1 = g[1] - g[3] / 2.0f, y1 = g[2] - g[4] / 2.0f;
            gx[4 * t] = x1; gx[4 * t + 1] = y1; gx[4 * t + 2] = g[3] + x1; gx[4 * t + 3] = g[4] + y1;
        }
        for (int p = 0; p < cells; ++p) {
            size_t cell = (size_t)b * cells + p;
            float mx = iou_f(gx, pbox + 4 * cell);            /* :116-118; torch.max propagates NaN */
            for (int t = 1; t < n; ++t) {
                float v = iou_f(gx + 4 * t, pbox + 4 * cell);
                if (isnan(v) || v > mx) { if (!isnan(mx)) mx = v; }
            }
            if (mx < ignore_thr) {                            /* :123-125 */
                wgt[cell * CH] = 1.0f;
                tgt[cell * CH] = 0.0f;
            }
        }
        free(gx);
    }

    /* anchor matching + sequential assignment loop, :127-169 */
    int count = 0, n_recall = 0, bad = 0;
    double obj = 0.0, ious = 0.0, cls_score = 0.0;
    float no_obj = (float)0; /* set below, fp32 running value like the reference tensor */
    {
        /* torch.sum in fp32; we keep the double sum and round once */
        no_obj = (float)sum_conf;
    }
    double iou_num = 0.0, iou_wsum = 0.0;
    for (int b = 0; b < N && !bad; ++b) {
        int g0 = gt_off[b], n = gt_off[b + 1] - gt_off[b];
        for (int t = 0; t < n && !bad; ++t) {
            const float *g = gt + 5 * (size_t)(g0 + t);
            float gxf = g[1] * (float)W, gyf = g[2] * (float)H; /* :128 */
            int gi = (int)gxf, gj = (int)gyf;                   /* :136-137 trunc */
            if (gi < 0 || gi >= W || gj < 0 || gj >= H) { bad = 1; break; }
            float gbox[4] = {0.0f, 0.0f, g[3], g[4]};           /* :129-130 */
            float best = -1.0f; int best_n = 0;
            float aiou[64];
            for (int a = 0; a < NA; ++a) {                      /* :132 over ALL anchors */
                float anc[4] = {0.0f, 0.0f, anchors_all[2 * a], anchors_all[2 * a + 1]};
                aiou[a] = iou_f(gbox, anc);
                if (a == 0 || aiou[a] > best) { best = aiou[a]; best_n = a; } /* argmax :133, first maximum */
            }
            int bn = NA + 1;                                    /* :140 */
            for (int k = 0; k < A; ++k) if (mask[k] == best_n) { bn = k; break; } /* :141-142 */
            float gxy[4];
            {
                float x1 = g[1] - g[3] / 2.0f, y1 = g[2] - g[4] / 2.0f;
                gxy[0] = x1; gxy[1] = y1; gxy[2] = g[3] + x1; gxy[3] = g[4] + y1;
            }
            int cls_index = (int)(g[0] - 1.0f);                 /* :131,147 */
            for (int k = 0; k < A; ++k) {
                if (!(k == bn || aiou[mask[k]] > iou_thr)) continue; /* :139,145 */
                size_t cell = (size_t)b * cells + ((size_t)k * H + gj) * W + gi;
                tgt[cell * CH] = 1.0f;                          /* :149 */
                wgt[cell * CH] = 1.0f;                          /* :150 */
                float conf = outp[cell * CH];                   /* :151 */
                obj += (double)conf;                            /* :152 */
                no_obj = (float)((double)no_obj - (double)conf);/* :153 fp32 tensor minus python float */
                float cv[2];
                oracle_box_ciou(gxy, pbox + 4 * cell, cv);      /* :157 */
                float wt = 2.0f - area_f(gxy);                  /* :160 */
                iou_num += (double)((cv[0] - 1.0f) * (cv[0] - 1.0f)); /* see Q6 below: weights cancel */
                iou_wsum += (double)wt;
                if (cv[1] > ignore_thr) ++n_recall;             /* :163 */
                ious += (double)cv[1];                          /* :165 */
                /* class_loss :425-434 */
                if (cls_index >= 0 && cls_index < C) {
                    if (wgt[cell * CH + 1 + cls_index] > 0.0f) {
                        tgt[cell * CH + 1 + cls_index] = 0.95f;
                        wgt[cell * CH + 1 + cls_index] = 1.0f;
                    } else {
                        for (int c = 0; c < C; ++c) { tgt[cell * CH + 1 + c] = 0.05f; wgt[cell * CH + 1 + c] = 1.0f; }
                        tgt[cell * CH + 1 + cls_index] = 0.95f;
                    }
                    cls_score += (double)outp[cell * CH + 1 + cls_index]; /* :169 */
                } else {
                    bad = 1;
                }
                if (assign && count < max_assign) {
                    int *r = assign + 6 * (size_t)count;
                    r[0] = b; r[1] = t; r[2] = k; r[3] = gj; r[4] = gi; r[5] = best_n;
                }
                if (terms && count < max_assign) {
                    float *r = terms + 3 * (size_t)count;
                    r[0] = cv[0]; r[1] = cv[1]; r[2] = wt;
                }
                ++count;
            }
        }
    }

    /* weighted_mse_loss :53-60 over the dense tensors, :219 */
    double num = 0.0, sw = 0.0;
#pragma omp parallel for reduction(+ : num, sw) schedule(static) num_threads(nthreads())
    for (size_t e = 0; e < tot; ++e) {
        float df = outp[e] - tgt[e];
        num += (double)(df * df) * (double)wgt[e];
        sw += (double)wgt[e];
    }
    double L_dense = num / sw;
    /* :222-224.  Quirk Q6: iou_losses is (n,1) (box_ciou returns (1,1) tensors, :289-293) but
     * iou_weights is (n,) (:160-162), so `out * weights / total` in weighted_mse_loss (:56)
     * broadcasts to (n,n) and the sum is  sum_i (v_i-1)^2 * (sum_j w_j / sum_j w_j):
     * the 2-area weights cancel and L_iou = mean((v-1)^2).  Verified against the reference
     * (tests/golden/loss_*.npz). */
    double L_iou = 0.0;
    if (count > 0) L_iou = iou_num / (double)count;

    if (scalars) {
        memset(scalars, 0, sizeof(double) * 16);
        scalars[0] = L_dense + L_iou * (double)iou_weighting; /* :234 */
        if (count > 0) {                                      /* :170-175 */
            scalars[1] = (double)n_recall / count;
            scalars[2] = ious / count;
            scalars[3] = obj / count;
            scalars[4] = (double)no_obj / (no_cnt - count);
            scalars[5] = cls_score / count;
        }
        scalars[6] = (double)count / N;                       /* :178 count/bs */
        scalars[7] = L_dense; scalars[8] = L_iou; scalars[9] = sw; scalars[10] = count;
        scalars[11] = num; scalars[12] = iou_num; scalars[13] = iou_wsum; scalars[14] = sum_conf;
        scalars[15] = n_recall;
    }
    free(outp); free(pbox);
    if (!targets) free(tgt);
    if (!weights) free(wgt);
    return bad ? -1 : count;
}
