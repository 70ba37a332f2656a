/*
 * b200yolo.h -- C ABI of libb200yolo.so: the B200 (sm_100a) detection hot path of
 * MobileNet-YOLO (decode -> confidence threshold -> per-class NMS; pairwise IoU;
 * YOLOLoss target assignment + loss).
 *
 * The library is the drop-in boundary.  Every entry point
 * replaces one Python-level function of the reference; the reference-side
 * binding is a ctypes stub.  Reference citations are relative
 * to the upstream repo root (eric612/Mobilenet-YOLO-Pytorch).
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / C++ types.
 *  - "dev" pointers are device memory owned by the caller; the library never
 *    allocates, frees or retains device memory passed to it (only the convenience
 *    form b200yolo_decode_nms_host keeps a staging buffer of its own; the _ws form takes the caller's).
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*;
 *    NULL = legacy default stream) and performs no host synchronisation, so
 *    calls are CUDA-graph capturable.  The *_host entry point is synchronous.
 *  - return value: 0 on success, a negative B200YOLO_E* code otherwise;
 *    b200yolo_last_error() returns a thread-local message.  Nothing throws.
 *  - head tensors are fp32, contiguous, NCHW: (N, A*(5+C), H, W); element
 *    (b,a,t,j,i) at (((b*A+a)*(5+C)+t)*H+j)*W+i   (models/yolo_loss.py:84,186).
 *  - candidate / detection rows are 7 floats:
 *    [x1, y1, x2, y2, conf, class_score, class_index]   (models/yolo_loss.py:199).
 *  - there is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef B200YOLO_H
#define B200YOLO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200YOLO_VERSION 100 /* 0.1.0 */

#define B200YOLO_OK 0
#define B200YOLO_EINVAL (-1)      /* bad argument / shape */
#define B200YOLO_EUNSUPPORTED (-2) /* shape exceeds what one CTA's shared memory can stage */
#define B200YOLO_ECUDA (-3)       /* a CUDA runtime call failed (see last_error) */
#define B200YOLO_ERANGE (-4)      /* a GT box maps outside the grid (reference: IndexError) */

#define B200YOLO_MAX_ANCHORS 8      /* anchors per head (reference: 3) */
#define B200YOLO_MAX_ALL_ANCHORS 16 /* anchors over all heads (reference: 6) */

int b200yolo_version(void);
const char *b200yolo_last_error(void);

/* Number of CUDA kernels this library has launched in this process (all
 * streams); bench.py reports the delta over the timed region as gpu_launches. */
unsigned long long b200yolo_launch_count(void);

/* Profiling aid (profiles/phase_times.py): when set to a device buffer of
 * [N][32] uint64, the fused / NMS kernels record %globaltimer (ns; slots 0-15) and the SM
 * cycle counter (slots 16-31) at their phase boundaries for every image.  NULL (the
 * default) disables it. */
void b200yolo_debug_phase_stamps(unsigned long long *dev_buf);

/* Experiment / test switches of the fused kernel (initialised from the B200YOLO_FLAGS
 * environment variable): 1 = no L2 prefetch of head 0, 2 = no programmatic dependent launch (consecutive
 * launches then run in plain stream order), 8 = prefetch both heads, 16 = never
 * use the compile-time head shapes (every shape then runs the runtime-stride decode), 32 = always wait for the
 * previous kernel before the first global read, 64 = three CTAs of 384 threads per SM for the VOC-352 shape
 * (experiment), 128 = "exact" decode: IEEE sigmoid / expf and a true division by the grid size, the reference's own
 * operations, instead of the SFU forms and the multiplication by the reciprocal (profiles/exact_decode.py),
 * 512 = with phase stamps: the buffer is a ring of 16 launches ([16][N][32]) and the stamped launches overlap like
 * production launches (steady-state phase times), 1024 = b200yolo_decode_nms_batches: every launch waits for its
 * predecessor before it stores even when the outputs are disjoint, 2048 = ... at most two launches in flight even when
 * no two batches share an output, 16384 = no objectness-first decode of the second head (decode_nms.cuh,
 * decode_head_static_sparse). */
void b200yolo_debug_set_flags(int flags);

/* Largest number of candidate cells per image (sum over heads of A*H*W) that
 * the fused / NMS kernels can stage in one CTA's shared memory on `device`. */
int b200yolo_max_cells(int device);

/*
 * YOLOLoss.get_pred_boxes  (models/yolo_loss.py:180-204; dispatch :238-241).
 * One head.  anchor_wh is a HOST array [A][2]: this head's anchors already
 * divided by img_size (yolo_loss.py:214) and rounded to fp32.
 *   rows  dev [N][A*H*W][7]  rows with conf > conf_thr, (a,j,i) row-major (:203)
 *   count dev [N]
 *   ids   dev [N][A*H*W] cell id (a*H+j)*W+i per emitted row, or NULL
 */
int b200yolo_decode_head(const float *head, int N, int A, int C, int H, int W, const float *anchor_wh,
                         float conf_thr, float *rows, int *count, int *ids, void *stream);

/*
 * utils.box.nms  (utils/box.py:11-31) incl. the arithmetic of
 * torchvision.ops.nms (call site utils/box.py:28): per image concatenate the
 * two heads' candidates (:17), per class (:20-22) score = col5*col4 (:27),
 * stable descending sort, greedy suppression iff (double)iou > iou_thr.
 *   cand0/cand1 dev [N][stride0|1][7], count0/count1 dev [N]
 *   out       dev [N][stride0+stride1][7]: class-ascending blocks, score-descending
 *   out_count dev [N]
 *   out_idx   dev [N][stride0+stride1]: index into the image's concatenated
 *             candidate list of every output row (the "keep indices"), or NULL
 */
int b200yolo_nms(const float *cand0, const int *count0, int stride0, const float *cand1, const int *count1,
                 int stride1, int N, int C, double iou_thr, float *out, int *out_count, int *out_idx, void *stream);

/*
 * Fused inference post-process of a two-head detector: what
 * models/mbv2_yolo.py:158-160 computes with two YOLOLoss.forward calls and
 * utils.box.nms, in ONE kernel launch with no host synchronisation.
 *   anchor_wh HOST [2][A][2] scaled anchors of head 0 then head 1
 *   out       dev [N][K][7], K = A*H0*W0 + A*H1*W1; out_count dev [N]
 *   out_idx   dev [N][K] global cell id of every kept row (head-1 ids are offset
 *             by A*H0*W0), or NULL
 */
int b200yolo_decode_nms(const float *head0, const float *head1, int N, int A, int C, int H0, int W0, int H1,
                        int W1, const float *anchor_wh, float conf_thr, double iou_thr, float *out,
                        int *out_count, int *out_idx, void *stream);

/*
 * The same for a LIST of batches of one shape: what a loop over `yolo.forward` of models/mbv2_yolo.py:158-160 does
 * batch after batch (the evaluation loop train.py:357-395, a serving queue), as one call.  Launch k runs the fused
 * kernel on batches[k]; because launch k > 0 directly follows the library's own launch k - 1 in the stream -- which
 * writes nothing that launch k reads -- it is allowed to start on the SM slots its predecessor leaves free and to
 * stream its heads under the predecessor's NMS (programmatic dependent launch).  The first launch waits for whatever
 * precedes the call in the stream before it reads.  No batch's `out` / `out_count` / `out_idx` may alias another batch's
 * heads.  Outputs and ordering (decided per call from the pointers):
 *   - some consecutive batches share an output buffer: every launch waits for its predecessor before it stores;
 *   - consecutive batches write disjoint buffers (results in a ring of two or more): no launch waits before it
 *     stores; one extra CTA per launch keeps at most two launches in flight, so a buffer is never written while the
 *     launch that used it two steps earlier still runs;
 *   - no two batches of the list share an output buffer: launches are not ordered against each other at all.
 * In every case a launch completes after its predecessor, so whatever follows the call in the stream sees all results.
 */
typedef struct b200yolo_batch {
    const float *head0, *head1; /* dev (N, A*(5+C), H0, W0), (N, A*(5+C), H1, W1) */
    float *out;                 /* dev [N][K][7] */
    int *out_count;             /* dev [N] */
    int *out_idx;               /* dev [N][K] or NULL */
} b200yolo_batch;
int b200yolo_decode_nms_batches(const b200yolo_batch *batches, int n_batches, int N, int A, int C, int H0, int W0, int H1,
                                int W1, const float *anchor_wh, float conf_thr, double iou_thr, void *stream);

/*
 * The same list as a replayable plan: the launches of b200yolo_decode_nms_batches captured once into a CUDA graph (with
 * their programmatic-launch edges) and issued by ONE cudaGraphLaunch per b200yolo_plan_launch -- for a loop that visits
 * the same device buffers again and again (an evaluation loop over a resident ring of head buffers, train.py:357-395; a
 * serving queue with fixed slots).  Creation must run with the buffers' device current and does not touch the caller's
 * streams (thread-local capture on a private stream); the pointers and thresholds are frozen into the plan.  A plan may
 * be launched any number of times on any stream of its device, one launch at a time; destroy it when no launch is in
 * flight.
 */
typedef struct b200yolo_plan b200yolo_plan;
int b200yolo_plan_create(const b200yolo_batch *batches, int n_batches, int N, int A, int C, int H0, int W0, int H1, int W1,
                         const float *anchor_wh, float conf_thr, double iou_thr, b200yolo_plan **plan_out);
int b200yolo_plan_launch(b200yolo_plan *plan, void *stream);
int b200yolo_plan_destroy(b200yolo_plan *plan);

/*
 * Single calls (b200yolo_decode_nms & co.) execute griddepcontrol.wait before their first global read, because the
 * kernel that precedes them in the stream may be the producer of their inputs.  A caller that issues them back to
 * back on buffers no kernel of this library writes (a benchmark loop over resident head tensors) may declare that
 * with b200yolo_set_inputs_ready(1): consecutive single calls then overlap like the launches of
 * b200yolo_decode_nms_batches.  Process-wide; default 0.
 */
void b200yolo_set_inputs_ready(int ready);

/*
 * Decode arithmetic of YOLOLoss.get_pred_boxes (models/yolo_loss.py:186-199).  Default (0): SFU sigmoid / exp and a
 * multiplication by the reciprocal of the grid size -- within 1e-5 of the reference (the north star's tolerance).
 * exact != 0: the reference's own operations -- 1/(1+expf(-x)), expf, a true division by the grid size -- so that
 * decoded rows AND final detections are bit-identical to the unmodified reference running on device='cuda'
 * (tests/test_reference_integration.py, profiles/exact_decode.py); costs ~3 % of the fused kernel.  Process-wide.
 */
void b200yolo_set_exact_decode(int exact);

/*
 * The same for channels-last heads: head tensors laid out (N, H, W, A*(5+C)) in memory -- what cuDNN prefers for the
 * last convolution of the head (models/mbv2_yolo.py:82,144,153) -- so no NCHW copy is needed before the
 * post-processing.  Same results, bit for bit, as b200yolo_decode_nms on the permuted
 * tensors.  Limits: 5+C <= 32; B200YOLO_EUNSUPPORTED when the image leaves no shared memory for the staging.
 */
int b200yolo_decode_nms_nhwc(const float *head0, const float *head1, int N, int A, int C, int H0, int W0, int H1,
                             int W1, const float *anchor_wh, float conf_thr, double iou_thr, float *out,
                             int *out_count, int *out_idx, void *stream);

/*
 * The same for images with more candidate cells than one CTA can stage in shared memory (b200yolo_decode_nms
 * answers B200YOLO_EUNSUPPORTED above b200yolo_max_cells(): ~5.6 k cells), e.g. the 832x832 variant of the
 * dense-candidate stress configuration: heads (N,75,26,26) + (N,75,52,52) = 10 140 cells per image.  One CTA per
 * image keeps only 8-byte sort keys in shared memory; the decoded records live in the caller's workspace
 * (b200yolo_decode_nms_large_workspace_bytes(N, cells) bytes, 16-byte aligned) and the per-class greedy NMS
 * (utils/box.py:16-30 + torchvision nms) runs tile by tile against the kept boxes, without an n^2 mask.
 * Same outputs, same order, as b200yolo_decode_nms.  Limits: cells per image <= 16384.
 * b200yolo_decode_head and b200yolo_nms switch to their large-image kernels by themselves (no workspace needed:
 * the decode writes rows straight from registers, the NMS reads boxes from the caller's rows).
 */
size_t b200yolo_decode_nms_large_workspace_bytes(int N, int cells_per_image);
int b200yolo_decode_nms_large(const float *head0, const float *head1, int N, int A, int C, int H0, int W0, int H1,
                              int W1, const float *anchor_wh, float conf_thr, double iou_thr, float *out,
                              int *out_count, int *out_idx, void *workspace, size_t workspace_bytes, void *stream);

/*
 * Decode + NMS fused with the data-parallel all-gather of the detections (the reference is single-GPU; the
 * north star shards the batch by image over the GPUs of one NVSwitch box and all-gathers the per-rank detections).
 * Rank `rank` of R post-processes its N images and its output phase stores every kept row straight into the gather
 * buffer of EVERY rank -- peer_out[r] dev [R*N][K][7] and peer_count[r] dev int32 [R*N] on rank r, image slot
 * rank*N + b -- its own through local stores, the others' through NVLink peer mappings (b200yolo_peer_open), so the
 * transfer overlaps the NMS of the images still in flight instead of following the kernel as a separate collective.
 * peer_out / peer_count are HOST arrays of R device pointers (R <= 8).  When every rank's launch has completed
 * (any stream-ordered barrier across the ranks, e.g. a 1-element all-reduce) all R buffers hold the whole batch.
 */
int b200yolo_decode_nms_gather(const float *head0, const float *head1, int N, int A, int C, int H0, int W0, int H1,
                               int W1, const float *anchor_wh, float conf_thr, double iou_thr, float *const *peer_out,
                               int *const *peer_count, int R, int rank, void *stream);

/*
 * Peer-visible device memory for the call above: b200yolo_peer_alloc = cudaMalloc (zero-filled) + a 64-byte CUDA IPC
 * handle that the other ranks of the node pass to b200yolo_peer_open (cudaIpcOpenMemHandle with lazy peer access);
 * b200yolo_peer_close / b200yolo_peer_free undo them.
 */
int b200yolo_peer_alloc(size_t bytes, void **dev_ptr, unsigned char *handle64);
int b200yolo_peer_open(const unsigned char *handle64, void **dev_ptr);
/*
 * Fence for the fused all-gather without a collective.  b200yolo_peer_signal (stream-ordered after the launch): rank
 * `rank` raises slot [rank] of the flag array of EVERY rank (peer_flags: HOST array of R device pointers to int[R]
 * arrays living in the peer-visible buffers) to `value`, a step number that only grows.  b200yolo_peer_wait: returns
 * to the stream once all R slots of this rank's own array are >= value, i.e. every rank's launch of this step has
 * completed and its rows are in this rank's buffer; *timed_out (dev int, zero it once) is set instead if a peer does
 * not arrive within timeout_s seconds (<= 60; default 5), so a dead peer cannot hang the GPU.
 */
int b200yolo_peer_signal(int *const *peer_flags, int R, int rank, int value, void *stream);
int b200yolo_peer_wait(const int *own_flags, int R, int value, double timeout_s, int *timed_out, void *stream);
int b200yolo_peer_close(void *dev_ptr);
int b200yolo_peer_free(void *dev_ptr);
/*
 * The two fence kernels above as launches that do not serialise the stream (programmatic launch): the signal kernel
 * waits for the b200yolo_decode_nms_gather launch right before it to complete, then raises this rank's flag in every
 * rank's array (`value`: the number of steps this rank has completed, it only grows); the wait kernel polls own_flags
 * until all R slots have reached `value`.  A kernel launched after them may start early; it must not rely on them for
 * ordering against the decode launch (b200yolo_decode_nms_gather_steps uses the flags themselves for that).
 */
int b200yolo_peer_fence(int *const *peer_flags, const int *own_flags, int R, int rank, int value, double timeout_s,
                        int *timed_out, void *stream);

/*
 * A sequence of data-parallel steps in one call -- ONE kernel launch per step.  Step s (= first_step + k) runs the
 * fused decode + NMS + all-gather on batches[k] (head0 / head1; the other members are ignored) into gather buffer
 * s % 3 of every rank.  This rank's arrival flag for step s (value s + 1 in every rank's flag array) is raised by an
 * extra CTA of the launch of step s + 1, which waits for the launch before it to complete (no separate signal launch:
 * an extra kernel per step costs ~2.3 us of launch processing); the call's last step is announced by the final fence,
 * or, with final_fence == 0, by the caller's b200yolo_peer_fence.
 * Back-pressure without a release signal: before its first store the kernel of step s waits (in the kernel, on this
 * rank's own flag array) until EVERY rank has completed step s - 2; the launch of step s - 2 follows -- in that rank's
 * stream -- whatever it ran on the buffer of step s - 3, the previous user of buffer s % 3.  In a pipeline that runs
 * in step these flags arrived a whole step ago: the wait never stalls.
 * final_fence != 0 appends b200yolo_peer_fence (value first_step + n_steps): when it completes, every rank's rows of
 * the call's last step are in this rank's buffers.
 * Contract for consumers: read the buffers of step s after its fence, with ordinary launches, before launching step
 * s + 1.
 */
/*
 * NVSwitch multicast memory for the gather buffers: a store to the multicast view is replicated by the switch into
 * the memory every GPU of the group has bound at the same offset, so a kept row leaves its GPU once instead of once per
 * peer (the all-gather's egress drops from (R - 1) x to 1 x).  Set-up, one process per GPU, `bytes` equal on all:
 *   rank 0:       b200yolo_mc_create -> a POSIX file descriptor, to be passed to the other processes (SCM_RIGHTS)
 *   other ranks:  b200yolo_mc_import(fd)
 *   all ranks:    b200yolo_mc_add_device; barrier; b200yolo_mc_bind -> local_ptr (this GPU's memory, for readers) and
 *                 mc_ptr (the multicast view, for the kernel's stores); barrier before the first store
 * b200yolo_mc_supported: 1 when the device and driver offer it.  The driver API is resolved at run time (dlopen), so
 * the library has no link dependency on libcuda.
 */
typedef struct b200yolo_mc b200yolo_mc;
int b200yolo_mc_supported(int device);
int b200yolo_mc_create(size_t bytes, int n_devices, int *fd, b200yolo_mc **mc);
int b200yolo_mc_import(size_t bytes, int n_devices, int fd, b200yolo_mc **mc);
int b200yolo_mc_add_device(b200yolo_mc *mc);
int b200yolo_mc_bind(b200yolo_mc *mc, void **local_ptr, void **mc_ptr);
int b200yolo_mc_free(b200yolo_mc *mc);

#define B200YOLO_GATHER_BUFFERS 3
typedef struct b200yolo_gather {
    int R, rank;
    float *peer_out[B200YOLO_GATHER_BUFFERS][8];  /* [buffer][r]: rank r's buffer dev [R*N][K][7] (own: local, others: b200yolo_peer_open) */
    int *peer_count[B200YOLO_GATHER_BUFFERS][8];  /* [buffer][r]: rank r's counts dev int32 [R*N] */
    int *peer_flags[8];     /* [r]: rank r's arrival flags dev int32 [8] */
    int *timed_out;         /* own dev int32[1], zero it once */
    double timeout_s;       /* per wait, <= 60; <= 0: 5 s */
    int multicast;          /* != 0: peer_out[buffer][0] / peer_count[buffer][0] are NVSwitch multicast addresses
                               (b200yolo_mc_bind) that reach the buffer of EVERY rank; the other entries are ignored */
} b200yolo_gather;
int b200yolo_decode_nms_gather_steps(const b200yolo_gather *g, const b200yolo_batch *batches, int n_steps, int first_step,
                                     int final_fence, int N, int A, int C, int H0, int W0, int H1, int W1,
                                     const float *anchor_wh, float conf_thr, double iou_thr, void *stream);

/*
 * Same computation from HOST buffers (the reference-facing call bench.py times as "e2e"; inference.py:121 and
 * train.py:366 hand host data to the model): heads are copied host->device in image chunks on three streams,
 * post-processed, and counts + detections copied back, overlapped.  Only rows that can be kept travel back: per chunk
 * ONE strided copy whose width is the largest count of the chunk (rows past an image's count are left untouched).
 * Pinned host memory is recommended (pageable works, slower).  Synchronous; thread-safe without a lock (streams
 * are per thread).
 *   b200yolo_decode_nms_host_ws   device staging provided by the CALLER (SURVEY 8b ownership): dev_workspace of
 *                                 b200yolo_decode_nms_host_workspace_bytes(...) bytes on `device`; concurrent calls
 *                                 need distinct workspaces
 *   b200yolo_decode_nms_host      convenience form: the library keeps one grow-only workspace per (thread, device)
 *   b200yolo_host_last_d2h_bytes  device->host bytes of this thread's last call (counts + rows)
 */
size_t b200yolo_decode_nms_host_workspace_bytes(int N, int A, int C, int H0, int W0, int H1, int W1);
int b200yolo_decode_nms_host_ws(const float *head0, const float *head1, int N, int A, int C, int H0, int W0,
                                int H1, int W1, const float *anchor_wh, float conf_thr, double iou_thr,
                                float *out, int *out_count, void *dev_workspace, size_t dev_workspace_bytes, int device);
int b200yolo_decode_nms_host(const float *head0, const float *head1, int N, int A, int C, int H0, int W0,
                             int H1, int W1, const float *anchor_wh, float conf_thr, double iou_thr,
                             float *out, int *out_count, int device);
size_t b200yolo_host_last_d2h_bytes(void);

/*
 * Compaction of fixed-stride detections for the data-parallel all-gather (no counterpart in the reference, which is
 * single-GPU): packed dev [sum count][7] gets the kept rows of image 0, 1, ... back to back (capacity N*K rows is
 * always enough), offsets dev int32 [N+1] the first packed row of every image (offsets[N] = total).
 */
int b200yolo_compact_rows(const float *dets, const int *count, int N, int K, float *packed, int *offsets, void *stream);

/*
 * utils.iou.find_intersection / find_union / find_jaccard_overlap
 * (utils/iou.py:4-13, 14-31, 32-49).  mode 0 / 1 / 2.  set1 dev [n1][4],
 * set2 dev [n2][4] xyxy; out dev [n1][n2].  mode 3 / 4: YOLOLoss.box_giou / box_ciou value
 * iou - term with box1 = set1 row, box2 = set2 row (models/yolo_loss.py:295-317 / 257-293).
 */
int b200yolo_pairwise(const float *set1, int n1, const float *set2, int n2, int mode, float *out, void *stream);

/*
 * YOLOLoss.forward(input, targets)  (models/yolo_loss.py:206-236): get_target
 * (:77-178: decode, pred-vs-GT ignore mask, anchor-vs-GT IoU, best-anchor
 * assignment, CIoU terms, class targets, stats) fused with weighted_mse_loss
 * (:53-60) into partial sums.  One head.
 *   anchors_all HOST [NA][2] ALL anchors / img_size, fp32;  mask HOST [A]
 *   gt        dev [G][5] rows [cls(1-based), cx, cy, w, h]; gt_off dev [N+1]
 *   G         total number of GT rows (= gt_off[N] on the host)
 *   max_gt_per_image  upper bound on the GT rows of any one image (sizes the shared-memory
 *             staging; 0 = unknown: min(G, 1024) is assumed).  At most 1024.
 *   sums      dev double[16], OVERWRITTEN with this call's partial sums (see
 *             B200YOLO_S_* below).  Data-parallel callers all-reduce(SUM) this
 *             vector across ranks, then call b200yolo_loss_finalize.
 *   assign    dev int32 [G][A][4] = (assigned?, gj, gi, best_n) per (GT, k), or NULL
 *   terms     dev float [G][A][2] = (ciou value v, iou) per (GT, k), or NULL
 *   status    dev int32[1], overwritten: 0 ok, 1 a GT maps outside the grid or has
 *             a class outside [1,C] (reference: IndexError), 2 an image has more GT
 *             boxes than max_gt_per_image / 1024
 *   cell_state dev uint8 [N][A*H*W] or NULL: per cell 0 = ignored (weight 0), 1 = no-object (weight 1,
 *             target 0), 2 = assigned; b200yolo_target_loss_backward consumes it
 *   workspace dev, b200yolo_target_loss_workspace_bytes(N) bytes: per-CTA partial sums
 *             (up to 8 CTAs share an image), reduced in a fixed order so results are
 *             bitwise reproducible run to run
 */
size_t b200yolo_target_loss_workspace_bytes(int N);
int b200yolo_target_loss(const float *head, int N, int A, int C, int H, int W, const float *anchors_all, int NA,
                         const int *mask, const float *gt, const int *gt_off, int G, float ignore_thr,
                         float iou_thr, int max_gt_per_image, double *sums, int *assign, float *terms, int *status,
                         unsigned char *cell_state, void *workspace, size_t workspace_bytes, void *stream);

/*
 * Backward of YOLOLoss.forward(input, targets): d loss / d input exactly as the reference's autograd
 * graph defines it (loss.backward() in train.py:282): the custom sigmoid passes gradients through
 * unchanged (models/yolo_loss.py:15-32), exp has its true derivative (:86), objectness / class entries
 * with weight 1 get 2 (o - t) / sum(w) (:53-60), the CIoU loss (:154-159, 224) reaches tx, ty, tw, th of
 * the assigned cells (alpha is not detached, :283).
 *   cell_state dev: what b200yolo_target_loss wrote for the same head, GT and thresholds
 *   sums       dev double[16]: the partial sums AFTER the cross-rank all-reduce (read on the device, no
 *              host synchronisation): the gradient of a shard is scaled by the batch-global normalisers
 *   grad_out   dev float[1] = d(total)/d(loss), or NULL for 1
 *   grad_input dev (N, A*(5+C), H, W): every element is written
 */
int b200yolo_target_loss_backward(const float *head, int N, int A, int C, int H, int W, const float *anchors_all, int NA,
                                  const int *mask, const float *gt, const int *gt_off, int G, float iou_thr,
                                  int max_gt_per_image, const unsigned char *cell_state, const double *sums,
                                  float iou_weighting, const float *grad_out, float *grad_input, void *stream);

/*
 * utils.eval_mAP.calculate_mAP  (utils/eval_mAP.py:134-188, eval_class_ap :65-132,
 * eval_single_image_recall :8-63; called from train.py:421 on the NMS output): per class the greedy
 * detection <-> ground-truth matching of every image (IoU > iou_thr, 'difficult' objects ignored), then
 * the 11-point interpolated average precision over the score-sorted detections.
 *   det_boxes dev [D][4] xyxy (16-byte aligned), det_labels dev int32 [D] in 1..n_classes-1,
 *   det_scores dev [D], det_off dev int32 [N+1]: the detections of image b are rows det_off[b]..det_off[b+1]
 *             in the order the reference would see them (train.py:385-388)
 *   true_boxes dev [T][4], true_labels dev int32 [T], true_difficult dev uint8 [T], true_off dev [N+1]
 *   recall_thresholds HOST float[n_thresholds <= 16]: torch.arange(0, 1.1, .1) in the reference (:120)
 *   ap, tp_sum, fp_sum dev float [n_classes-1]: average precision, #true / #false positives per class
 *             (mAP = mean of ap, taken by the caller)
 *   workspace dev, b200yolo_map_eval_workspace_bytes(D, T, N, n_classes) bytes
 * Score ties are ordered by detection index (the reference's torch.sort leaves them unspecified).
 */
size_t b200yolo_map_eval_workspace_bytes(int D, int T, int N, int n_classes);
int b200yolo_map_eval(const float *det_boxes, const int *det_labels, const float *det_scores, const int *det_off, int D,
                      const float *true_boxes, const int *true_labels, const unsigned char *true_difficult,
                      const int *true_off, int T, int N, int n_classes, float iou_thr, const float *recall_thresholds,
                      int n_thresholds, float *ap, float *tp_sum, float *fp_sum, void *workspace, size_t workspace_bytes,
                      void *stream);

/*
 * models.seg_loss.SegLoss  (models/seg_loss.py:14-81; called at models/mbv2_yolo.py:163,170): the
 * drivable-area head of the BDD100k multi-task model.
 *   b200yolo_seg_loss           forward(input, targets) (:51-76): input dev (N, C, H, W), truth dev (N, H, W, C)
 *                               (the reference permutes it, :54).  sums dev double[8], overwritten:
 *                               [0] sum (sigmoid(x) - t)^2, [1] numel, [2] sum sigmoid(x) where t >= 0.5,
 *                               [3] count(t >= 0.5), [4] sum sigmoid(x) where t < 0.5, [5] count(t < 0.5);
 *                               loss = 0.05 * [0]/[1], obj = [2]/[3], no_obj = [4]/[5]  (:40-45, 65-66, 76)
 *   b200yolo_seg_loss_backward  d (0.05 * mse) / d input with the pass-through sigmoid (:15-31):
 *                               grad_out[0] * 0.05 * 2 (sigmoid(x) - t) / numel; grad_out dev float[1] or NULL = 1
 *   b200yolo_seg_sigmoid        forward(input) (:77-80): out[i] = 1/(1+exp(-input[i])), i < count
 */
size_t b200yolo_seg_loss_workspace_bytes(void);
int b200yolo_seg_loss(const float *input, const float *truth, int N, int C, int H, int W, double *sums, void *workspace,
                      size_t workspace_bytes, void *stream);
int b200yolo_seg_loss_backward(const float *input, const float *truth, int N, int C, int H, int W, const float *grad_out,
                               float *grad_input, void *stream);
int b200yolo_seg_sigmoid(const float *input, long long count, float *out, void *stream);

/* indices into the partial-sum vector of b200yolo_target_loss */
enum {
    B200YOLO_S_SQW = 0,      /* sum (o-t)^2 w           (yolo_loss.py:54-58) */
    B200YOLO_S_W = 1,        /* sum w                   (:55) */
    B200YOLO_S_IOU_SQ = 2,   /* sum_i (v_i-1)^2         (:224, quirk Q6: weights cancel) */
    B200YOLO_S_IOU_W = 3,    /* sum_i (2-area_i)        (:160) */
    B200YOLO_S_NASSIGN = 4,  /* count                   (:146) */
    B200YOLO_S_OBJ = 5,      /* sum conf at assigned    (:152) */
    B200YOLO_S_CONF_ALL = 6, /* sum conf over all cells (:98) */
    B200YOLO_S_CLS = 7,      /* sum class score         (:169) */
    B200YOLO_S_IOU = 8,      /* sum iou                 (:165) */
    B200YOLO_S_RECALL = 9,   /* #(iou > ignore_thr)     (:163-164) */
    B200YOLO_S_NCELLS = 10,  /* N*A*H*W                 (:99) */
    B200YOLO_S_NIMG = 11,    /* N                       (:178) */
    B200YOLO_S_COUNT = 16
};

/*
 * Host-side finalisation (yolo_loss.py:170-178, 219-236) from (all-reduced)
 * partial sums: result[7] = loss, recall, avg_iou, obj, no_obj, cls, count/N.
 */
int b200yolo_loss_finalize(const double *sums, float iou_weighting, double *result);
/* The same on the device (sums dev double[16] -> result dev float[7], one tiny launch): a training step can keep the
 * loss tensor on the device and read the statistics later, without a host round trip per head. */
int b200yolo_loss_finalize_dev(const double *sums, float iou_weighting, float *result, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B200YOLO_H */
