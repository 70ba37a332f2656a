// b200yolo.cu -- C ABI (include/b200yolo.h) over the sm_100a kernels.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false
//        -shared -Xcompiler -fPIC  (see mobilenet_yolo_pytorch_b200/build.py)
#include "../../include/b200yolo.h"

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>      // driver API types; the entry points are resolved with dlopen (no link dependency)
#include <dlfcn.h>

#include <algorithm>
#include <mutex>
#include <vector>
#include <new>

#include "decode_nms.cuh"  // (after the helpers it uses)
#include "target_loss.cuh"
#include "pairwise.cuh"
#include "map_eval.cuh"
#include "seg_loss.cuh"
#include "large_nms.cuh"

using namespace b200yolo;

namespace {

thread_local char g_err[512] = "";
std::atomic<unsigned long long *> g_dbg{nullptr};  // profiling aid, see b200yolo_debug_phase_stamps
std::atomic<unsigned> g_dbg_seq{0};
std::atomic<unsigned long long> g_launches{0};
std::atomic<int> g_flags{[] { const char *e = getenv("B200YOLO_FLAGS"); return e ? atoi(e) : 0; }()};  // experiment switches
std::atomic<int> g_inputs_ready{0};   // see b200yolo_set_inputs_ready

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char *what) {
    return fail(B200YOLO_ECUDA, "%s: %s", what, cudaGetErrorString(e));
}

#define CUDA_TRY(expr)                                         \
    do {                                                       \
        cudaError_t _e = (expr);                               \
        if (_e != cudaSuccess) return cuda_fail(_e, #expr);    \
    } while (0)

int smem_optin(int device) {
    static std::atomic<int> cache[64];  // (zero-initialised; the attribute never changes)
    if (device >= 0 && device < 64) {
        const int c = cache[device].load(std::memory_order_relaxed);
        if (c > 0) return c;
    }
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess) return 0;
    if (device >= 0 && device < 64) cache[device].store(v, std::memory_order_relaxed);
    return v;
}

int current_device(int *dev) {
    cudaError_t e = cudaGetDevice(dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    return 0;
}

uint32_t div_magic(int d) { return d <= 1 ? 0u : (uint32_t)((0x100000000ull + (unsigned long long)d - 1) / (unsigned long long)d); }

void fill_head(HeadDesc &h, const float *ptr, int A, int H, int W, const float *anchor_wh) {
    h.ptr = ptr;
    h.H = H;
    h.W = W;
    h.HW = H * W;
    h.cells = A * H * W;
    h.magicHW = div_magic(H * W);
    h.magicW = div_magic(W);
    h.fW = (float)W;
    h.fH = (float)H;
    h.rW = 1.0f / (float)W;
    h.rH = 1.0f / (float)H;
    for (int a = 0; a < kMaxAnchors; ++a) {
        h.aw[a] = (a < A) ? anchor_wh[2 * a] : 0.f;
        h.ah[a] = (a < A) ? anchor_wh[2 * a + 1] : 0.f;
    }
}

IouThr make_thr(double thr) {
    IouThr t;
    t.thr = thr;
    t.fast_ok = (thr >= 0.01 && thr <= 1.0) ? 1 : 0;
    const double tt = t.fast_ok ? thr / (1.0 + thr) : 0.0;  // iou > thr <=> inter > tt * (area_a + area_b)
    t.ts = (float)tt * 1.220703125e-4f;                     // * 2^-13, exact
    t.tf = (float)tt;
    t.th = (float)(tt * 64.0 * (1.0 - 0.00390625));         // fp16 prefilter: X = x/16, Y = 1024*y, margin 2^-8
    return t;
}

// Shared-memory tiers.  The decode phase streams the heads through L1, and L1 is what
// the resident CTAs' shared memory leaves of the SM's 228 KB (carve-out steps ... 132,
// 164, 196, 228 KB), so the layout is kept as small as the image allows:
//   <= 80 KB : two CTAs of 512 threads per SM, 164 KB carve-out, 64 KB of L1
//   <= 113 KB: two CTAs of 512 threads per SM, no L1 to speak of
//   <= 162 KB: one CTA of 1024 threads per SM, 64 KB of L1
//   else     : one CTA of 1024 threads, up to the opt-in limit (227 KB)
constexpr int kTier2x64 = 80 * 1024;
constexpr int kTier2x0 = (228 * 1024) / 2 - 1024;
constexpr int kTier1x64 = 162 * 1024;

template <int MODE, int THREADS, int SHAPE>
int launch_dn_t(DNParams &p, const SmemLayout &L, int dev, cudaStream_t st) {
    static std::mutex mu;
    static int configured[5][64] = {{0}, {0}, {0}, {0}, {0}};  // smem size opted into, per kernel variant and device
    using Kern = void (*)(const DNParams, const SmemLayout);
    Kern kern = decode_nms_kernel<MODE, THREADS, SHAPE, 0>;
    int variant = 0;
    if constexpr (SHAPE == 0 && MODE != MODE_NMS && (THREADS == 512 || THREADS == 1024)) {
        if (p.flags & 128) {   // exact decode (b200yolo_set_exact_decode): the runtime-shape variants carry it
            if constexpr (MODE == MODE_FUSED) {
                if (p.gR > 0) { kern = decode_nms_kernel<MODE, THREADS, 0, 1, false, true>; variant = 4; }
                else { kern = decode_nms_kernel<MODE, THREADS, 0, 0, false, true>; variant = 3; }
            } else {
                kern = decode_nms_kernel<MODE, THREADS, 0, 0, false, true>;
                variant = 3;
            }
        }
    }
    if (variant >= 3) {
        // (fall through to the launch)
    } else
    if constexpr (MODE == MODE_FUSED) {
        if (p.gR > 0) {  // fused all-gather: the output phase stores into every rank's buffer
            kern = decode_nms_kernel<MODE, THREADS, SHAPE, 1>;
            variant = 1;
        } else if constexpr (THREADS == 512 && (SHAPE == 0 || SHAPE == 1)) {
            if (p.dbg) {  // phase time stamps (profiles/phase_times.py): a separate instantiation
                kern = decode_nms_kernel<MODE, THREADS, SHAPE, 0, true>;
                variant = 2;
            }
        }
    }
    {
        std::lock_guard<std::mutex> g(mu);
        if (dev < 64 && configured[variant][dev] < (int)L.total) {
            CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin(dev)));
            configured[variant][dev] = smem_optin(dev);
        }
    }
    // programmatic dependent launch: consecutive launches of this kernel overlap (decode_nms.cuh, pdl_trigger / pdl_wait)
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    const bool gather_variant = (variant == 1 || variant == 4);
    if (gather_variant || MODE != MODE_FUSED) p.chain = 0;
    // (+ the signalling CTA of the gather, or the chain CTA of a list of batches: decode_nms.cuh)
    cfg.gridDim = dim3((unsigned)p.N + ((gather_variant ? p.gsignal > 0 : p.chain >= 2) ? 1u : 0u));
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = L.total;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = (p.flags & 2) ? 0 : 1;  // flag 2: plain stream order
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, p, L));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// compile-time head shapes of the fused kernel (decode_nms.cuh, ShapeT): the reference's own configurations
template <int S>
bool shape_is(const DNParams &p) {
    using SH = ShapeT<S>;
    return p.nheads == 2 && p.C == SH::C && p.head[0].HW == SH::HW0 && p.head[0].W == SH::W0 && p.head[1].HW == SH::HW1 &&
           p.head[1].W == SH::W1;
}

template <int S>
bool head_is(const DNParams &p) {
    using SH = ShapeT<S>;
    return p.nheads == 1 && p.C == SH::C && p.head[0].HW == SH::HW0 && p.head[0].W == SH::W0;
}

template <int MODE, int THREADS>
int launch_dn_shape(DNParams &p, const SmemLayout &L, int dev, cudaStream_t st) {
    if (MODE == MODE_FUSED && p.nhwc) {
        // the channels-last decode stages 32 cells per warp in the part of U behind clsidx: as many warps as fit
        const uint32_t Kp = align_up((uint32_t)(p.K > 0 ? p.K : 1), 32);
        const uint32_t free_bytes = 8 * Kp;   // the key region of U (decode_nms.cuh, make_layout)
        int nwarps = (int)(free_bytes / (32u * (uint32_t)p.attrs * sizeof(float)));
        if (nwarps > THREADS / 32) nwarps = THREADS / 32;
        if (nwarps < 2)
            return fail(B200YOLO_EUNSUPPORTED, "decode_nms_nhwc: %d cells per image leave no room for the channels-last staging "
                        "(convert the heads to NCHW)", p.K);
        p.nhwc = nwarps;
        if (THREADS == 512 && !(p.flags & (16 | 128))) {
            constexpr bool kOn = (MODE == MODE_FUSED && THREADS == 512);
            if (p.C == 20) return launch_dn_t<MODE, THREADS, kOn ? 21 : 0>(p, L, dev, st);
            if (p.C == 10) return launch_dn_t<MODE, THREADS, kOn ? 22 : 0>(p, L, dev, st);
        }
    }
    if (MODE == MODE_FUSED && !(p.flags & (16 | 128)) && !p.nhwc) {  // flag 16: force the runtime-shape path (tests); 128: exact decode
        if (THREADS == 512 && shape_is<1>(p)) return launch_dn_t<MODE, THREADS, (MODE == MODE_FUSED && THREADS == 512) ? 1 : 0>(p, L, dev, st);
        if (THREADS == 512 && shape_is<2>(p)) return launch_dn_t<MODE, THREADS, (MODE == MODE_FUSED && THREADS == 512) ? 2 : 0>(p, L, dev, st);
        if (THREADS == 1024 && shape_is<3>(p)) return launch_dn_t<MODE, THREADS, (MODE == MODE_FUSED && THREADS == 1024) ? 3 : 0>(p, L, dev, st);
    }
    if (MODE == MODE_DECODE && THREADS == 512 && !(p.flags & (16 | 128))) {
        constexpr bool kOn = (MODE == MODE_DECODE && THREADS == 512);
        if (head_is<11>(p)) return launch_dn_t<MODE, THREADS, kOn ? 11 : 0>(p, L, dev, st);
        if (head_is<12>(p)) return launch_dn_t<MODE, THREADS, kOn ? 12 : 0>(p, L, dev, st);
        if (head_is<13>(p)) return launch_dn_t<MODE, THREADS, kOn ? 13 : 0>(p, L, dev, st);
        if (head_is<14>(p)) return launch_dn_t<MODE, THREADS, kOn ? 14 : 0>(p, L, dev, st);
        if (head_is<15>(p)) return launch_dn_t<MODE, THREADS, kOn ? 15 : 0>(p, L, dev, st);
        if (head_is<16>(p)) return launch_dn_t<MODE, THREADS, kOn ? 16 : 0>(p, L, dev, st);
    }
    return launch_dn_t<MODE, THREADS, 0>(p, L, dev, st);
}

// limits of the one-CTA-per-image layout beyond shared memory (decode_nms.cuh): the sorting warps carry kRankItems
// sorted positions per thread over a barrier, tile-padded positions are 16 bits, a class has at most 255 tiles
bool dn_counts_ok(int K, int C, int mode, int threads) {
    if (mode == MODE_DECODE) return true;
    const long long Kp = ((long long)(K > 0 ? K : 1) + 31) / 32 * 32;
    return K <= kRankItems * (threads - 32) && Kp + 32LL * C <= 65535 && K <= 255 * 32;
}

template <int MODE>
int launch_dn(DNParams &p, cudaStream_t st) {
    int dev = 0;
    if (int rc = current_device(&dev)) return rc;
    const int lim = smem_optin(dev);
    const SmemLayout base1k = make_layout(p.K, p.C, MODE, 1024, 0);
    if ((int)base1k.total > lim || !dn_counts_ok(p.K, p.C, MODE, 1024))
        return fail(B200YOLO_EUNSUPPORTED,
                    "%d candidate cells per image with %d classes need %u B of shared memory (limit %d B)", p.K, p.C,
                    base1k.total, lim);
    if (p.N == 0) return 0;
    p.B = pick_buckets(p.C);
    p.Bshift = 0;
    while ((1 << p.Bshift) < p.B) ++p.Bshift;
    p.dbg = g_dbg.load();
    p.flags = g_flags.load();
    // flag 512 (profiling aid): the stamp buffer is a ring of 16 launches of N x 32 words and the stamped launches
    // overlap like production launches (phase times in steady state, profiles/phase_times.py --steady)
    if (p.dbg && (p.flags & 512)) p.dbg += (size_t)(g_dbg_seq.fetch_add(1) % 16) * (size_t)p.N * 32;
    // Programmatic dependent launch lets this kernel start while its predecessor in the stream is still running.
    // That is only harmless when the predecessor does not produce this kernel's inputs -- true for back-to-back
    // launches of this library on different batches, NOT true in general (CUDA makes a predecessor's writes
    // visible only after griddepcontrol.wait).  So the kernel waits before its first global read unless the
    // previous kernel in the stream is known not to produce them: launches 2..n of b200yolo_decode_nms_batches
    // (their predecessor is our own launch on another batch), or a caller that vouches for it with
    // b200yolo_set_inputs_ready.  Flag 32 forces the wait.
    if (g_inputs_ready.load()) p.wait_inputs = 0;
    if (p.flags & 32) p.wait_inputs = 1;
    const int need512 = (int)make_layout(p.K, p.C, MODE, 512, 0).total;
    const int need = (int)base1k.total;
    // the pair masks get whatever the tier leaves (MODE_DECODE has none)
    if constexpr (MODE == MODE_FUSED) {
        // experiment (flag 64): three CTAs of 384 threads per SM (<= 74 KB of shared memory each, <= 56 registers)
        if ((p.flags & 64) && !p.nhwc && p.gR == 0 && shape_is<1>(p) && dn_counts_ok(p.K, p.C, MODE, 384)) {
            const SmemLayout L3 = make_layout(p.K, p.C, MODE, 384, 0);
            if ((int)L3.total <= 74 * 1024) return launch_dn_t<MODE, 384, 1>(p, L3, dev, st);
        }
    }
    const bool ok512 = dn_counts_ok(p.K, p.C, MODE, 512);
    if (ok512 && need512 <= kTier2x64 && kTier2x64 <= lim)
        return launch_dn_shape<MODE, 512>(p, make_layout(p.K, p.C, MODE, 512, (MODE == MODE_DECODE) ? 0u : (uint32_t)(kTier2x64 - need512)), dev, st);
    if (ok512 && need512 <= kTier2x0 && kTier2x0 <= lim)
        return launch_dn_shape<MODE, 512>(p, make_layout(p.K, p.C, MODE, 512, (MODE == MODE_DECODE) ? 0u : (uint32_t)(kTier2x0 - need512)), dev, st);
    const int budget = (need <= kTier1x64 && kTier1x64 <= lim) ? kTier1x64 : lim;
    return launch_dn_shape<MODE, 1024>(p, make_layout(p.K, p.C, MODE, 1024, (MODE == MODE_DECODE) ? 0u : (uint32_t)(budget - need)), dev, st);
}

// Fence of the fused all-gather without a collective: after its launch every rank raises its slot of the flag array
// in EVERY rank's buffer to the step number (the stores of the launch before it in the stream are complete at that
// kernel boundary), then waits until all slots of its own array have reached it.
struct PeerFlagPtrs { int *p[kMaxPeers]; };

__global__ void peer_signal_kernel(PeerFlagPtrs flags, int R, int rank, int value) {
    const int r = threadIdx.x;
    if (r < R) {
        __threadfence_system();
        *reinterpret_cast<volatile int *>(flags.p[r] + rank) = value;
    }
}

__global__ void peer_wait_kernel(const int *own_flags, int R, int value, long long max_cycles, int *timed_out) {
    const int r = threadIdx.x;
    if (r < R) {
        const long long t0 = clock64();
        while (*reinterpret_cast<const volatile int *>(own_flags + r) < value) {
            if (clock64() - t0 > max_cycles) {   // a peer died: report instead of hanging the GPU
                *timed_out = 1;
                break;
            }
            __nanosleep(200);
        }
        __threadfence_system();
    }
}

// The two halves of the gather's fence as kernels that do not serialise the stream (programmatic launch, trigger at
// their start).  peer_signal_pdl_kernel directly follows the decode launch whose rows it announces: it waits for
// that launch to complete and flush (griddepcontrol.wait), then raises this rank's arrival flag in every rank's
// array.  peer_wait_pdl_kernel depends on nothing in its stream: it only polls this rank's own flags.
__global__ void peer_signal_pdl_kernel(PeerFlagPtrs flags, int R, int rank, int value) {
    pdl_trigger();
    pdl_wait();
    const int r = threadIdx.x;
    if (r < R) {
        __threadfence_system();
        *reinterpret_cast<volatile int *>(flags.p[r] + rank) = value;
    }
}

__global__ void peer_wait_pdl_kernel(const int *own_flags, int R, int value, long long max_cycles, int *timed_out) {
    pdl_trigger();
    const int r = threadIdx.x;
    if (r < R) {
        const long long t0 = clock64();
        while (*reinterpret_cast<const volatile int *>(own_flags + r) < value) {
            if (clock64() - t0 > max_cycles) {   // a peer died: report instead of hanging the GPU
                *timed_out = 1;
                break;
            }
            __nanosleep(100);
        }
        __threadfence_system();
    }
}

long long timeout_cycles(double timeout_s) {
    if (!(timeout_s > 0.0) || timeout_s > 60.0) timeout_s = 5.0;
    return (long long)(timeout_s * 2.0e9);   // SM clock <= 2 GHz: at least timeout_s seconds
}

template <typename... Args>
int launch_pdl_1warp(cudaStream_t st, void (*kern)(Args...), Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(32);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = (g_flags.load() & 2) ? 0 : 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, args...));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

// yolo_loss.py:170-178, 219-236 on the device (same formulas as b200yolo_loss_finalize): result[7] = loss, recall,
// avg_iou, obj, no_obj, cls, count/N as fp32 -- lets a training step keep the loss on the device (no host round trip)
__global__ void loss_finalize_kernel(const double *s, float iou_weighting, float *r) {
    if (threadIdx.x != 0) return;
    const double count = s[B200YOLO_S_NASSIGN];
    const double l_dense = s[B200YOLO_S_SQW] / s[B200YOLO_S_W];
    const double l_iou = count > 0 ? s[B200YOLO_S_IOU_SQ] / count : 0.0;
    r[0] = (float)(l_dense + l_iou * (double)iou_weighting);
    if (count > 0) {
        r[1] = (float)(s[B200YOLO_S_RECALL] / count);
        r[2] = (float)(s[B200YOLO_S_IOU] / count);
        r[3] = (float)(s[B200YOLO_S_OBJ] / count);
        r[4] = (float)((s[B200YOLO_S_CONF_ALL] - s[B200YOLO_S_OBJ]) / (s[B200YOLO_S_NCELLS] - count));
        r[5] = (float)(s[B200YOLO_S_CLS] / count);
    } else {
        r[1] = r[2] = r[3] = r[4] = r[5] = 0.f;
    }
    r[6] = (float)(s[B200YOLO_S_NIMG] > 0 ? count / s[B200YOLO_S_NIMG] : 0.0);
}

// large_nms.cuh: one CTA per image, sort keys in shared memory.  SRC 0: heads (records in p.rec), 1: caller rows
template <int SRC>
int launch_large(LargeParams &p, cudaStream_t st, const char *who) {
    if (p.K > kLargeMaxKeys)
        return fail(B200YOLO_EUNSUPPORTED, "%s: %d cells per image (limit %d)", who, p.K, kLargeMaxKeys);
    p.P = 32;
    while (p.P < p.K) p.P <<= 1;
    p.T = p.K / 32 + p.C + 1;
    int dev = 0;
    if (int rc = current_device(&dev)) return rc;
    const size_t smem = large_smem_bytes(p.P, p.C, p.T);
    const int lim = smem_optin(dev);
    if (smem > (size_t)lim)
        return fail(B200YOLO_EUNSUPPORTED, "%s: %d cells, %d classes need %zu B of shared memory (limit %d B)", who, p.K, p.C,
                    smem, lim);
    {
        static std::mutex mu;
        static bool configured[64] = {false};
        std::lock_guard<std::mutex> g(mu);
        if (dev < 64 && !configured[dev]) {
            CUDA_TRY(cudaFuncSetAttribute(decode_nms_large_kernel<SRC>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
            configured[dev] = true;
        }
    }
    decode_nms_large_kernel<SRC><<<p.N, kLargeThreads, smem, st>>>(p);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// does the shared-memory layout of decode_nms.cuh hold K cells at all (one CTA of 1024 threads, opt-in limit)?
bool fits_one_cta(int K, int C, int mode) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return true;  // (the regular path reports the error)
    return (int)make_layout(K, C, mode, 1024, 0).total <= smem_optin(dev) && dn_counts_ok(K, C, mode, 1024);
}

}  // namespace

extern "C" {

int b200yolo_version(void) { return B200YOLO_VERSION; }
const char *b200yolo_last_error(void) { return g_err; }
unsigned long long b200yolo_launch_count(void) { return g_launches.load(); }

void b200yolo_debug_phase_stamps(unsigned long long *dev_buf) { g_dbg.store(dev_buf); g_dbg_seq.store(0); }
void b200yolo_debug_set_flags(int flags) { g_flags.store(flags); }
void b200yolo_set_inputs_ready(int ready) { g_inputs_ready.store(ready ? 1 : 0); }
void b200yolo_set_exact_decode(int exact) {
    if (exact) g_flags.fetch_or(128); else g_flags.fetch_and(~128);
}

int b200yolo_max_cells(int device) {
    const int lim = smem_optin(device);
    int lo = 0, hi = 1 << 16;
    while (lo < hi) {  // largest K whose fused layout fits (C = 80 as a conservative class count)
        int mid = (lo + hi + 1) / 2;
        if ((int)make_layout(mid, 80, MODE_FUSED, 1024, 0).total <= lim && dn_counts_ok(mid, 80, MODE_FUSED, 1024)) lo = mid; else hi = mid - 1;
    }
    return lo;
}

int b200yolo_decode_head(const float *head, int N, int A, int C, int H, int W, const float *anchor_wh,
                         float conf_thr, float *rows, int *count, int *ids, void *stream) {
    if (!head || !anchor_wh || !rows || !count) return fail(B200YOLO_EINVAL, "decode_head: null pointer");
    if (N < 0 || A < 1 || A > kMaxAnchors || C < 1 || C > 4096 || H < 1 || W < 1)
        return fail(B200YOLO_EINVAL, "decode_head: bad shape N=%d A=%d C=%d H=%d W=%d", N, A, C, H, W);
    if ((long long)A * H * W > 65535) return fail(B200YOLO_EUNSUPPORTED, "decode_head: more than 65535 cells per image");
    DNParams p;
    memset(&p, 0, sizeof(p));
    fill_head(p.head[0], head, A, H, W, anchor_wh);
    p.nheads = 1;
    p.N = N; p.A = A; p.C = C; p.attrs = 5 + C;
    p.K = A * H * W;
    p.conf_thr = conf_thr;
    p.iou = make_thr(0.45);
    p.out = rows; p.out_count = count; p.out_idx = ids;
    if (!fits_one_cta(p.K, C, MODE_DECODE)) {  // e.g. a 52x52 head (8112 cells): rows straight from registers
        if (N == 0) return 0;
        LargeDecodeParams q;
        memset(&q, 0, sizeof(q));
        q.head = p.head[0];
        q.N = N; q.A = A; q.C = C; q.K = p.K;
        q.conf_thr = conf_thr;
        q.rows = rows; q.count = count; q.ids = ids;
        decode_head_large_kernel<<<N, kLargeThreads, 0, (cudaStream_t)stream>>>(q);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
    p.wait_inputs = 1;
    return launch_dn<MODE_DECODE>(p, (cudaStream_t)stream);
}

int b200yolo_nms(const float *cand0, const int *count0, int stride0, const float *cand1, const int *count1,
                 int stride1, int N, int C, double iou_thr, float *out, int *out_count, int *out_idx, void *stream) {
    if (!cand0 || !count0 || !out || !out_count) return fail(B200YOLO_EINVAL, "nms: null pointer");
    if ((cand1 == nullptr) != (count1 == nullptr)) return fail(B200YOLO_EINVAL, "nms: cand1/count1 mismatch");
    if (!cand1) stride1 = 0;
    if (N < 0 || C < 1 || C > 4096 || stride0 < 0 || stride1 < 0 || stride0 + stride1 < 1)
        return fail(B200YOLO_EINVAL, "nms: bad shape N=%d C=%d strides=%d,%d", N, C, stride0, stride1);
    if (stride0 + stride1 > 65535) return fail(B200YOLO_EUNSUPPORTED, "nms: more than 65535 candidates per image");
    if (!(iou_thr == iou_thr)) return fail(B200YOLO_EINVAL, "nms: NaN threshold");
    DNParams p;
    memset(&p, 0, sizeof(p));
    p.nheads = 0;
    p.N = N; p.A = 0; p.C = C; p.attrs = 5 + C;
    p.K = stride0 + stride1;
    p.iou = make_thr(iou_thr);
    p.out = out; p.out_count = out_count; p.out_idx = out_idx;
    p.cand[0] = cand0; p.cand_count[0] = count0; p.cand_stride[0] = stride0;
    p.cand[1] = cand1; p.cand_count[1] = count1; p.cand_stride[1] = stride1;
    if (!fits_one_cta(p.K, C, MODE_NMS)) {  // more rows per image than one CTA stages: keys only (large_nms.cuh)
        if (N == 0) return 0;
        LargeParams q;
        memset(&q, 0, sizeof(q));
        q.N = N; q.C = C; q.attrs = 5 + C;
        q.K = p.K;
        q.iou = p.iou;
        q.out = out; q.out_count = out_count; q.out_idx = out_idx;
        q.cand[0] = cand0; q.cand_count[0] = count0; q.cand_stride[0] = stride0;
        q.cand[1] = cand1; q.cand_count[1] = count1; q.cand_stride[1] = stride1;
        return launch_large<1>(q, (cudaStream_t)stream, "nms");
    }
    p.wait_inputs = 1;
    return launch_dn<MODE_NMS>(p, (cudaStream_t)stream);
}

int b200yolo_decode_nms(const float *head0, const float *head1, int N, int A, int C, int H0, int W0, int H1,
                        int W1, const float *anchor_wh, float conf_thr, double iou_thr, float *out,
                        int *out_count, int *out_idx, void *stream) {
    if (!head0 || !head1 || !anchor_wh || !out || !out_count) return fail(B200YOLO_EINVAL, "decode_nms: null pointer");
    if (N < 0 || A < 1 || A > kMaxAnchors || C < 1 || C > 4096 || H0 < 1 || W0 < 1 || H1 < 1 || W1 < 1)
        return fail(B200YOLO_EINVAL, "decode_nms: bad shape");
    if (!(iou_thr == iou_thr)) return fail(B200YOLO_EINVAL, "decode_nms: NaN threshold");
    const long long cells = (long long)A * H0 * W0 + (long long)A * H1 * W1;
    if (cells > 65535) return fail(B200YOLO_EUNSUPPORTED, "decode_nms: more than 65535 cells per image");
    DNParams p;
    memset(&p, 0, sizeof(p));
    fill_head(p.head[0], head0, A, H0, W0, anchor_wh);
    fill_head(p.head[1], head1, A, H1, W1, anchor_wh + 2 * A);
    p.nheads = 2;
    p.N = N; p.A = A; p.C = C; p.attrs = 5 + C;
    p.K = (int)cells;
    p.conf_thr = conf_thr;
    p.iou = make_thr(iou_thr);
    p.out = out; p.out_count = out_count; p.out_idx = out_idx;
    p.wait_inputs = 1;
    return launch_dn<MODE_FUSED>(p, (cudaStream_t)stream);
}

int b200yolo_decode_nms_batches(const b200yolo_batch *batches, int n_batches, int N, int A, int C, int H0, int W0, int H1,
                                int W1, const float *anchor_wh, float conf_thr, double iou_thr, void *stream) {
    if (n_batches < 0 || (n_batches > 0 && !batches)) return fail(B200YOLO_EINVAL, "decode_nms_batches: bad batch list");
    if (!anchor_wh) return fail(B200YOLO_EINVAL, "decode_nms_batches: null pointer");
    if (N < 0 || A < 1 || A > kMaxAnchors || C < 1 || C > 4096 || H0 < 1 || W0 < 1 || H1 < 1 || W1 < 1)
        return fail(B200YOLO_EINVAL, "decode_nms_batches: bad shape");
    if (!(iou_thr == iou_thr)) return fail(B200YOLO_EINVAL, "decode_nms_batches: NaN threshold");
    const long long cells = (long long)A * H0 * W0 + (long long)A * H1 * W1;
    if (cells > 65535) return fail(B200YOLO_EUNSUPPORTED, "decode_nms_batches: more than 65535 cells per image");
    for (int k = 0; k < n_batches; ++k)
        if (!batches[k].head0 || !batches[k].head1 || !batches[k].out || !batches[k].out_count)
            return fail(B200YOLO_EINVAL, "decode_nms_batches: null pointer in batch %d", k);
    DNParams p;
    memset(&p, 0, sizeof(p));
    p.nheads = 2;
    p.N = N; p.A = A; p.C = C; p.attrs = 5 + C;
    p.K = (int)cells;
    p.conf_thr = conf_thr;
    p.iou = make_thr(iou_thr);
    // consecutive batches with disjoint outputs (a ring of two or more result buffers): the launches are chained and
    // do not wait for their predecessor before they store (DNParams::chain); flag 1024 switches it off
    bool chained = n_batches >= 2 && N > 0 && !(g_flags.load() & 1024);
    for (int k = 1; k < n_batches && chained; ++k) {
        const b200yolo_batch &x = batches[k], &y = batches[k - 1];
        const size_t rows = (size_t)N * (size_t)cells;
        const void *px[3] = {x.out, x.out_count, x.out_idx}, *py[3] = {y.out, y.out_count, y.out_idx};
        const size_t len[3] = {rows * 7 * sizeof(float), (size_t)N * sizeof(int), rows * sizeof(int)};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                if (!px[i] || !py[j]) continue;
                const uintptr_t a0 = (uintptr_t)px[i], a1 = a0 + len[i], b0 = (uintptr_t)py[j], b1 = b0 + len[j];
                if (a0 < b1 && b0 < a1) chained = false;
            }
    }
    // ... and when no two batches of the list share an output at all, nothing needs to order the stores (chain 3);
    // flag 2048 keeps the two-in-flight chain
    bool all_distinct = chained && !(g_flags.load() & 2048);
    if (all_distinct) {
        struct Span { uintptr_t lo, hi; };
        std::vector<Span> spans;
        spans.reserve(3 * (size_t)n_batches);
        const size_t rows = (size_t)N * (size_t)cells;
        for (int k = 0; k < n_batches; ++k) {
            spans.push_back({(uintptr_t)batches[k].out, (uintptr_t)batches[k].out + rows * 7 * sizeof(float)});
            spans.push_back({(uintptr_t)batches[k].out_count, (uintptr_t)batches[k].out_count + (size_t)N * sizeof(int)});
            if (batches[k].out_idx) spans.push_back({(uintptr_t)batches[k].out_idx, (uintptr_t)batches[k].out_idx + rows * sizeof(int)});
        }
        std::sort(spans.begin(), spans.end(), [](const Span &a, const Span &b) { return a.lo < b.lo; });
        for (size_t i = 1; i < spans.size(); ++i)
            if (spans[i].lo < spans[i - 1].hi) all_distinct = false;
    }
    for (int k = 0; k < n_batches; ++k) {
        p.chain = chained ? (k == 0 ? 1 : all_distinct ? 3 : 2) : 0;
        fill_head(p.head[0], batches[k].head0, A, H0, W0, anchor_wh);
        fill_head(p.head[1], batches[k].head1, A, H1, W1, anchor_wh + 2 * A);
        p.out = batches[k].out; p.out_count = batches[k].out_count; p.out_idx = batches[k].out_idx;
        // launch k > 0 follows our own launch k - 1 in the stream, which writes nothing this one reads: it may start
        // on the SM slots its predecessor leaves free and stream its heads under the predecessor's NMS
        p.wait_inputs = (k == 0) ? 1 : 0;
        if (int rc = launch_dn<MODE_FUSED>(p, (cudaStream_t)stream)) return rc;
    }
    return 0;
}

// The list of batches as ONE CUDA graph: the launches of b200yolo_decode_nms_batches are captured (thread-local capture on
// a private stream, so nothing else the process does is affected) with their programmatic-launch edges, and replayed by
// one cudaGraphLaunch -- the host's per-launch cost (~2.5 us each, which the GPU outruns at the start of a short list)
// is paid once at creation.
struct b200yolo_plan {
    cudaGraphExec_t exec;
    int device, n_batches;
    unsigned long long kernel_nodes;   // what one launch of the plan adds to b200yolo_launch_count()
};

int b200yolo_plan_create(const b200yolo_batch *batches, int n_batches, int N, int A, int C, int H0, int W0, int H1, int W1,
                         const float *anchor_wh, float conf_thr, double iou_thr, b200yolo_plan **plan_out) {
    if (!plan_out) return fail(B200YOLO_EINVAL, "plan_create: null pointer");
    *plan_out = nullptr;
    if (n_batches < 1) return fail(B200YOLO_EINVAL, "plan_create: empty batch list");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    cudaStream_t s = nullptr;
    CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    cudaError_t e = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) {
        cudaStreamDestroy(s);
        return fail(B200YOLO_ECUDA, "plan_create: cudaStreamBeginCapture: %s", cudaGetErrorString(e));
    }
    const unsigned long long counted = g_launches.load();
    const int rc = b200yolo_decode_nms_batches(batches, n_batches, N, A, C, H0, W0, H1, W1, anchor_wh, conf_thr, iou_thr, s);
    g_launches.fetch_sub(g_launches.load() - counted);   // captured, not executed (a plan is created by one thread at a time)
    cudaGraph_t graph = nullptr;
    e = cudaStreamEndCapture(s, &graph);
    if (rc) {                                  // b200yolo_last_error() already names the cause
        if (graph) cudaGraphDestroy(graph);
        cudaStreamDestroy(s);
        (void)cudaGetLastError();
        return rc;
    }
    if (e != cudaSuccess || !graph) {
        cudaStreamDestroy(s);
        return fail(B200YOLO_ECUDA, "plan_create: cudaStreamEndCapture: %s", cudaGetErrorString(e));
    }
    size_t nodes = 0;
    cudaGraphGetNodes(graph, nullptr, &nodes);
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e == cudaSuccess) {                    // move the executable graph to the device now, not inside the first launch
        e = cudaGraphUpload(exec, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) cudaGraphExecDestroy(exec);
    }
    cudaStreamDestroy(s);
    if (e != cudaSuccess) return fail(B200YOLO_ECUDA, "plan_create: cudaGraphInstantiate / Upload: %s", cudaGetErrorString(e));
    b200yolo_plan *pl = new (std::nothrow) b200yolo_plan{exec, dev, n_batches, (unsigned long long)nodes};
    if (!pl) {
        cudaGraphExecDestroy(exec);
        return fail(B200YOLO_ECUDA, "plan_create: out of host memory");
    }
    *plan_out = pl;
    return 0;
}

int b200yolo_plan_launch(b200yolo_plan *plan, void *stream) {
    if (!plan || !plan->exec) return fail(B200YOLO_EINVAL, "plan_launch: null plan");
    CUDA_TRY(cudaGraphLaunch(plan->exec, (cudaStream_t)stream));
    g_launches.fetch_add(plan->kernel_nodes, std::memory_order_relaxed);
    return 0;
}

int b200yolo_plan_destroy(b200yolo_plan *plan) {
    if (!plan) return 0;
    cudaError_t e = plan->exec ? cudaGraphExecDestroy(plan->exec) : cudaSuccess;
    delete plan;
    if (e != cudaSuccess) return fail(B200YOLO_ECUDA, "plan_destroy: %s", cudaGetErrorString(e));
    return 0;
}

int b200yolo_decode_nms_gather(const float *head0, const float *head1, int N, int A, int C, int H0, int W0, int H1,
                               int W1, const float *anchor_wh, float conf_thr, double iou_thr, float *const *peer_out,
                               int *const *peer_count, int R, int rank, void *stream) {
    if (!head0 || !head1 || !anchor_wh || !peer_out || !peer_count) return fail(B200YOLO_EINVAL, "decode_nms_gather: null pointer");
    if (N < 0 || A < 1 || A > kMaxAnchors || C < 1 || C > 4096 || H0 < 1 || W0 < 1 || H1 < 1 || W1 < 1)
        return fail(B200YOLO_EINVAL, "decode_nms_gather: bad shape");
    if (R < 1 || R > kMaxPeers || rank < 0 || rank >= R)
        return fail(B200YOLO_EINVAL, "decode_nms_gather: %d ranks (1..%d), rank %d", R, kMaxPeers, rank);
    if (!(iou_thr == iou_thr)) return fail(B200YOLO_EINVAL, "decode_nms_gather: NaN threshold");
    for (int r = 0; r < R; ++r)
        if (!peer_out[r] || !peer_count[r]) return fail(B200YOLO_EINVAL, "decode_nms_gather: null buffer of rank %d", r);
    const long long cells = (long long)A * H0 * W0 + (long long)A * H1 * W1;
    if (cells > 65535) return fail(B200YOLO_EUNSUPPORTED, "decode_nms_gather: more than 65535 cells per image");
    DNParams p;
    memset(&p, 0, sizeof(p));
    fill_head(p.head[0], head0, A, H0, W0, anchor_wh);
    fill_head(p.head[1], head1, A, H1, W1, anchor_wh + 2 * A);
    p.nheads = 2;
    p.N = N; p.A = A; p.C = C; p.attrs = 5 + C;
    p.K = (int)cells;
    p.conf_thr = conf_thr;
    p.iou = make_thr(iou_thr);
    p.gR = R;
    p.gslot = rank * N;
    // every rank starts with its own buffer and walks the ring from there, so at any moment the ranks store into
    // different peers instead of all hitting rank 0 first
    p.gnbuf = R;
    for (int i = 0; i < R; ++i) { p.gout[i] = peer_out[(rank + i) % R]; p.gcount[i] = peer_count[(rank + i) % R]; }
    p.wait_inputs = 1;
    return launch_dn<MODE_FUSED>(p, (cudaStream_t)stream);
}

int b200yolo_peer_fence(int *const *peer_flags, const int *own_flags, int R, int rank, int value, double timeout_s,
                        int *timed_out, void *stream) {
    if (!peer_flags || !own_flags || !timed_out || R < 1 || R > kMaxPeers || rank < 0 || rank >= R)
        return fail(B200YOLO_EINVAL, "peer_fence: bad argument");
    PeerFlagPtrs f;
    memset(&f, 0, sizeof(f));
    for (int r = 0; r < R; ++r) {
        if (!peer_flags[r]) return fail(B200YOLO_EINVAL, "peer_fence: null flag array of rank %d", r);
        f.p[r] = peer_flags[r];
    }
    if (int rc = launch_pdl_1warp((cudaStream_t)stream, peer_signal_pdl_kernel, f, R, rank, value)) return rc;
    return launch_pdl_1warp((cudaStream_t)stream, peer_wait_pdl_kernel, own_flags, R, value, timeout_cycles(timeout_s), timed_out);
}

int b200yolo_decode_nms_gather_steps(const b200yolo_gather *g, const b200yolo_batch *batches, int n_steps, int first_step,
                                     int final_fence, int N, int A, int C, int H0, int W0, int H1, int W1,
                                     const float *anchor_wh, float conf_thr, double iou_thr, void *stream) {
    if (!g || n_steps < 0 || (n_steps > 0 && !batches) || !anchor_wh || first_step < 0)
        return fail(B200YOLO_EINVAL, "decode_nms_gather_steps: bad argument");
    if (N < 0 || A < 1 || A > kMaxAnchors || C < 1 || C > 4096 || H0 < 1 || W0 < 1 || H1 < 1 || W1 < 1)
        return fail(B200YOLO_EINVAL, "decode_nms_gather_steps: bad shape");
    const int R = g->R, rank = g->rank;
    if (R < 1 || R > kMaxPeers || rank < 0 || rank >= R || !g->timed_out)
        return fail(B200YOLO_EINVAL, "decode_nms_gather_steps: %d ranks (1..%d), rank %d", R, kMaxPeers, rank);
    if (!(iou_thr == iou_thr)) return fail(B200YOLO_EINVAL, "decode_nms_gather_steps: NaN threshold");
    for (int par = 0; par < B200YOLO_GATHER_BUFFERS; ++par)
        for (int r = 0; r < R; ++r)
            if (!g->peer_flags[r] || ((r == 0 || !g->multicast) && (!g->peer_out[par][r] || !g->peer_count[par][r])))
                return fail(B200YOLO_EINVAL, "decode_nms_gather_steps: null buffer of rank %d", r);
    const long long cells = (long long)A * H0 * W0 + (long long)A * H1 * W1;
    if (cells > 65535) return fail(B200YOLO_EUNSUPPORTED, "decode_nms_gather_steps: more than 65535 cells per image");
    if (n_steps == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const long long cyc = timeout_cycles(g->timeout_s);
    DNParams p;
    memset(&p, 0, sizeof(p));
    p.nheads = 2;
    p.N = N; p.A = A; p.C = C; p.attrs = 5 + C;
    p.K = (int)cells;
    p.conf_thr = conf_thr;
    p.iou = make_thr(iou_thr);
    p.gR = R;
    p.gslot = rank * N;
    p.gwait_flags = g->peer_flags[rank];
    p.gtimed_out = g->timed_out;
    p.gwait_cycles = cyc;
    p.grank = rank;
    for (int r = 0; r < R; ++r) p.gflags[r] = g->peer_flags[r];
    for (int k = 0; k < n_steps; ++k) {
        if (!batches[k].head0 || !batches[k].head1) return fail(B200YOLO_EINVAL, "decode_nms_gather_steps: null head in step %d", k);
        const int step = first_step + k, par = step % B200YOLO_GATHER_BUFFERS;
        fill_head(p.head[0], batches[k].head0, A, H0, W0, anchor_wh);
        fill_head(p.head[1], batches[k].head1, A, H1, W1, anchor_wh + 2 * A);
        // every rank starts with its own buffer and walks the ring from there (the ranks hit different peers)
        if (g->multicast) {   // one store through the switch reaches every rank's buffer
            p.gnbuf = 1;
            p.gout[0] = g->peer_out[par][0];
            p.gcount[0] = g->peer_count[par][0];
        } else {
            p.gnbuf = R;
            for (int i = 0; i < R; ++i) { p.gout[i] = g->peer_out[par][(rank + i) % R]; p.gcount[i] = g->peer_count[par][(rank + i) % R]; }
        }
        // Buffer step % 3 was last used by step - 3.  A rank whose flag shows step - 1 has COMPLETED step - 2, whose
        // launch follows -- in that rank's stream -- everything it ran on the buffers of step - 3.  Waiting for the
        // flags of two steps ago never stalls a pipeline that runs in step.
        p.gwait_value = step - 1;
        p.wait_inputs = (k == 0) ? 1 : 0;   // k > 0: the predecessor is our own launch
        // k > 0: this launch also announces that step - 1 -- the launch right before it in the stream -- is complete
        p.gsignal = (k > 0) ? step : 0;
        if (int rc = launch_dn<MODE_FUSED>(p, st)) return rc;
    }
    // the last step's arrival signal and the consumer-side fence: when it completes, the rows of the call's last step
    // -- and of every step before it -- are in this rank's buffers
    if (final_fence) {
        PeerFlagPtrs f;
        memset(&f, 0, sizeof(f));
        for (int r = 0; r < R; ++r) f.p[r] = g->peer_flags[r];
        if (int rc = launch_pdl_1warp(st, peer_signal_pdl_kernel, f, R, rank, first_step + n_steps)) return rc;
        return launch_pdl_1warp(st, peer_wait_pdl_kernel, (const int *)g->peer_flags[rank], R, first_step + n_steps, cyc, g->timed_out);
    }
    return 0;
}

}  // extern "C"

namespace {

// ---------------------------------------------------------------------------
// NVSwitch multicast for the gather buffers.  One multicast object spans the GPUs of the job; every rank binds its own
// physical memory to it and maps two views: the ordinary (unicast) one its consumers read, and the multicast one --
// a store to a multicast address is replicated by the switch into the bound memory of EVERY GPU at the same offset,
// so a kept row leaves its GPU once instead of once per peer.  Driver API (virtual memory management + cuMulticast*),
// resolved at run time with dlopen so that the library still loads on a machine without a driver.
// ---------------------------------------------------------------------------
struct DriverApi {
    void *lib = nullptr;
    bool ok = false;
    decltype(&cuDeviceGet) DeviceGet = nullptr;
    decltype(&cuDeviceGetAttribute) DeviceGetAttribute = nullptr;
    decltype(&cuGetErrorString) GetErrorString = nullptr;
    decltype(&cuMulticastGetGranularity) MulticastGetGranularity = nullptr;
    decltype(&cuMulticastCreate) MulticastCreate = nullptr;
    decltype(&cuMulticastAddDevice) MulticastAddDevice = nullptr;
    decltype(&cuMulticastBindMem) MulticastBindMem = nullptr;
    decltype(&cuMulticastUnbind) MulticastUnbind = nullptr;
    decltype(&cuMemGetAllocationGranularity) MemGetAllocationGranularity = nullptr;
    decltype(&cuMemCreate) MemCreate = nullptr;
    decltype(&cuMemRelease) MemRelease = nullptr;
    decltype(&cuMemAddressReserve) MemAddressReserve = nullptr;
    decltype(&cuMemAddressFree) MemAddressFree = nullptr;
    decltype(&cuMemMap) MemMap = nullptr;
    decltype(&cuMemUnmap) MemUnmap = nullptr;
    decltype(&cuMemSetAccess) MemSetAccess = nullptr;
    decltype(&cuMemExportToShareableHandle) MemExportToShareableHandle = nullptr;
    decltype(&cuMemImportFromShareableHandle) MemImportFromShareableHandle = nullptr;
};

const DriverApi &driver_api() {
    static DriverApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        api.lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
        if (!api.lib) return;
        bool all = true;
#define B200_DRV(field, name)                                              \
        api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, name)); \
        all = all && api.field != nullptr;
        B200_DRV(DeviceGet, "cuDeviceGet")
        B200_DRV(DeviceGetAttribute, "cuDeviceGetAttribute")
        B200_DRV(GetErrorString, "cuGetErrorString")
        B200_DRV(MulticastGetGranularity, "cuMulticastGetGranularity")
        B200_DRV(MulticastCreate, "cuMulticastCreate")
        B200_DRV(MulticastAddDevice, "cuMulticastAddDevice")
        B200_DRV(MulticastBindMem, "cuMulticastBindMem")
        B200_DRV(MulticastUnbind, "cuMulticastUnbind")
        B200_DRV(MemGetAllocationGranularity, "cuMemGetAllocationGranularity")
        B200_DRV(MemCreate, "cuMemCreate")
        B200_DRV(MemRelease, "cuMemRelease")
        B200_DRV(MemAddressReserve, "cuMemAddressReserve")
        B200_DRV(MemAddressFree, "cuMemAddressFree")
        B200_DRV(MemMap, "cuMemMap")
        B200_DRV(MemUnmap, "cuMemUnmap")
        B200_DRV(MemSetAccess, "cuMemSetAccess")
        B200_DRV(MemExportToShareableHandle, "cuMemExportToShareableHandle")
        B200_DRV(MemImportFromShareableHandle, "cuMemImportFromShareableHandle")
#undef B200_DRV
        api.ok = all;
    });
    return api;
}

int drv_fail(CUresult r, const char *what) {
    const char *msg = nullptr;
    const DriverApi &d = driver_api();
    if (d.GetErrorString) d.GetErrorString(r, &msg);
    return fail(B200YOLO_ECUDA, "%s: %s (CUresult %d)", what, msg ? msg : "?", (int)r);
}

#define DRV_TRY(expr)                                         \
    do {                                                      \
        CUresult _r = (expr);                                 \
        if (_r != CUDA_SUCCESS) return drv_fail(_r, #expr);   \
    } while (0)

}  // namespace

struct b200yolo_mc {
    CUmemGenericAllocationHandle mc = 0, mem = 0;
    CUdeviceptr local_va = 0, mc_va = 0;
    size_t size = 0;
    int device = -1, n_devices = 0;
    bool added = false, bound = false, mapped_local = false, mapped_mc = false;
};

namespace {

int mc_prepare(size_t bytes, int n_devices, int *device, CUmulticastObjectProp *prop) {
    const DriverApi &d = driver_api();
    if (!d.ok) return fail(B200YOLO_EUNSUPPORTED, "multicast: the CUDA driver library (libcuda.so.1) with cuMulticast* is not available");
    if (bytes == 0 || n_devices < 2 || n_devices > kMaxPeers) return fail(B200YOLO_EINVAL, "multicast: %d devices (2..%d)", n_devices, kMaxPeers);
    CUDA_TRY(cudaFree(nullptr));   // (the primary context exists and is current)
    if (int rc = current_device(device)) return rc;
    CUdevice dev;
    DRV_TRY(d.DeviceGet(&dev, *device));
    int sup = 0;
    DRV_TRY(d.DeviceGetAttribute(&sup, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev));
    if (!sup) return fail(B200YOLO_EUNSUPPORTED, "multicast: device %d does not support NVSwitch multicast", *device);
    memset(prop, 0, sizeof(*prop));
    prop->numDevices = (unsigned)n_devices;
    prop->handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    prop->flags = 0;
    prop->size = bytes;
    size_t gran = 0;
    DRV_TRY(d.MulticastGetGranularity(&gran, prop, CU_MULTICAST_GRANULARITY_RECOMMENDED));
    CUmemAllocationProp ap;
    memset(&ap, 0, sizeof(ap));
    ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    ap.location.id = *device;
    ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t agran = 0;
    DRV_TRY(d.MemGetAllocationGranularity(&agran, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    if (agran > gran) gran = agran;
    if (gran == 0) gran = (size_t)2 << 20;
    prop->size = (bytes + gran - 1) / gran * gran;
    return 0;
}

}  // namespace

extern "C" {

int b200yolo_mc_supported(int device) {
    const DriverApi &d = driver_api();
    if (!d.ok) return 0;
    CUdevice dev;
    int sup = 0;
    if (cudaFree(nullptr) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    if (d.DeviceGet(&dev, device) != CUDA_SUCCESS) return 0;
    if (d.DeviceGetAttribute(&sup, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) != CUDA_SUCCESS) return 0;
    return sup ? 1 : 0;
}

int b200yolo_mc_create(size_t bytes, int n_devices, int *fd, b200yolo_mc **out) {
    if (!fd || !out) return fail(B200YOLO_EINVAL, "mc_create: null pointer");
    *out = nullptr;
    int device = 0;
    CUmulticastObjectProp prop;
    if (int rc = mc_prepare(bytes, n_devices, &device, &prop)) return rc;
    const DriverApi &d = driver_api();
    b200yolo_mc *m = new (std::nothrow) b200yolo_mc();
    if (!m) return fail(B200YOLO_ECUDA, "mc_create: out of host memory");
    m->device = device; m->n_devices = n_devices; m->size = prop.size;
    CUresult r = d.MulticastCreate(&m->mc, &prop);
    if (r != CUDA_SUCCESS) { delete m; return drv_fail(r, "cuMulticastCreate"); }
    int h = -1;
    r = d.MemExportToShareableHandle(&h, m->mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
    if (r != CUDA_SUCCESS) { d.MemRelease(m->mc); delete m; return drv_fail(r, "cuMemExportToShareableHandle"); }
    *fd = h;
    *out = m;
    return 0;
}

int b200yolo_mc_import(size_t bytes, int n_devices, int fd, b200yolo_mc **out) {
    if (!out || fd < 0) return fail(B200YOLO_EINVAL, "mc_import: bad argument");
    *out = nullptr;
    int device = 0;
    CUmulticastObjectProp prop;
    if (int rc = mc_prepare(bytes, n_devices, &device, &prop)) return rc;
    const DriverApi &d = driver_api();
    b200yolo_mc *m = new (std::nothrow) b200yolo_mc();
    if (!m) return fail(B200YOLO_ECUDA, "mc_import: out of host memory");
    m->device = device; m->n_devices = n_devices; m->size = prop.size;
    CUresult r = d.MemImportFromShareableHandle(&m->mc, (void *)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    if (r != CUDA_SUCCESS) { delete m; return drv_fail(r, "cuMemImportFromShareableHandle"); }
    *out = m;
    return 0;
}

int b200yolo_mc_add_device(b200yolo_mc *m) {
    if (!m || !m->mc) return fail(B200YOLO_EINVAL, "mc_add_device: null object");
    const DriverApi &d = driver_api();
    CUdevice dev;
    DRV_TRY(d.DeviceGet(&dev, m->device));
    DRV_TRY(d.MulticastAddDevice(m->mc, dev));
    m->added = true;
    return 0;
}

int b200yolo_mc_bind(b200yolo_mc *m, void **local_ptr, void **mc_ptr) {
    if (!m || !m->mc || !m->added || !local_ptr || !mc_ptr) return fail(B200YOLO_EINVAL, "mc_bind: bad argument (add the device first)");
    const DriverApi &d = driver_api();
    CUmemAllocationProp ap;
    memset(&ap, 0, sizeof(ap));
    ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    ap.location.id = m->device;
    ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    DRV_TRY(d.MemCreate(&m->mem, m->size, &ap, 0));
    DRV_TRY(d.MulticastBindMem(m->mc, 0, m->mem, 0, m->size, 0));
    m->bound = true;
    CUmemAccessDesc acc;
    memset(&acc, 0, sizeof(acc));
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = m->device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    DRV_TRY(d.MemAddressReserve(&m->local_va, m->size, 0, 0, 0));
    DRV_TRY(d.MemMap(m->local_va, m->size, 0, m->mem, 0));
    m->mapped_local = true;
    DRV_TRY(d.MemSetAccess(m->local_va, m->size, &acc, 1));
    DRV_TRY(d.MemAddressReserve(&m->mc_va, m->size, 0, 0, 0));
    DRV_TRY(d.MemMap(m->mc_va, m->size, 0, m->mc, 0));
    m->mapped_mc = true;
    DRV_TRY(d.MemSetAccess(m->mc_va, m->size, &acc, 1));
    CUDA_TRY(cudaMemset((void *)m->local_va, 0, m->size));
    CUDA_TRY(cudaDeviceSynchronize());
    *local_ptr = (void *)m->local_va;
    *mc_ptr = (void *)m->mc_va;
    return 0;
}

int b200yolo_mc_free(b200yolo_mc *m) {
    if (!m) return 0;
    const DriverApi &d = driver_api();
    if (d.ok) {
        if (m->mapped_mc) d.MemUnmap(m->mc_va, m->size);
        if (m->mc_va) d.MemAddressFree(m->mc_va, m->size);
        if (m->mapped_local) d.MemUnmap(m->local_va, m->size);
        if (m->local_va) d.MemAddressFree(m->local_va, m->size);
        if (m->bound) {
            CUdevice dev;
            if (d.DeviceGet(&dev, m->device) == CUDA_SUCCESS) d.MulticastUnbind(m->mc, dev, 0, m->size);
        }
        if (m->mem) d.MemRelease(m->mem);
        if (m->mc) d.MemRelease(m->mc);
    }
    delete m;
    return 0;
}

/* Peer-visible device memory for the fused all-gather: cudaMalloc + an IPC handle another process of the node opens. */
int b200yolo_peer_alloc(size_t bytes, void **dev_ptr, unsigned char *handle64) {
    if (!dev_ptr || !handle64 || bytes == 0) return fail(B200YOLO_EINVAL, "peer_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void *ptr = nullptr;
    CUDA_TRY(cudaMalloc(&ptr, bytes));
    cudaError_t e = cudaMemset(ptr, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, ptr);
    if (e != cudaSuccess) {
        cudaFree(ptr);
        return cuda_fail(e, "peer_alloc");
    }
    memcpy(handle64, &h, 64);
    *dev_ptr = ptr;
    return 0;
}

int b200yolo_peer_open(const unsigned char *handle64, void **dev_ptr) {
    if (!dev_ptr || !handle64) return fail(B200YOLO_EINVAL, "peer_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    CUDA_TRY(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int b200yolo_peer_signal(int *const *peer_flags, int R, int rank, int value, void *stream) {
    if (!peer_flags || R < 1 || R > kMaxPeers || rank < 0 || rank >= R) return fail(B200YOLO_EINVAL, "peer_signal: bad argument");
    PeerFlagPtrs f;
    memset(&f, 0, sizeof(f));
    for (int r = 0; r < R; ++r) {
        if (!peer_flags[r]) return fail(B200YOLO_EINVAL, "peer_signal: null flag array of rank %d", r);
        f.p[r] = peer_flags[r];
    }
    peer_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(f, R, rank, value);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int b200yolo_peer_wait(const int *own_flags, int R, int value, double timeout_s, int *timed_out, void *stream) {
    if (!own_flags || !timed_out || R < 1 || R > kMaxPeers) return fail(B200YOLO_EINVAL, "peer_wait: bad argument");
    if (!(timeout_s > 0.0) || timeout_s > 60.0) timeout_s = 5.0;
    const long long max_cycles = (long long)(timeout_s * 2.0e9);   // SM clock <= 2 GHz: at least timeout_s seconds
    peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(own_flags, R, value, max_cycles, timed_out);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int b200yolo_peer_close(void *dev_ptr) {
    if (dev_ptr) CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
    return 0;
}

int b200yolo_peer_free(void *dev_ptr) {
    if (dev_ptr) CUDA_TRY(cudaFree(dev_ptr));
    return 0;
}

int b200yolo_decode_nms_nhwc(const float *head0, const float *head1, int N, int A, int C, int H0, int W0, int H1,
                             int W1, const float *anchor_wh, float conf_thr, double iou_thr, float *out,
                             int *out_count, int *out_idx, void *stream) {
    if (!head0 || !head1 || !anchor_wh || !out || !out_count) return fail(B200YOLO_EINVAL, "decode_nms_nhwc: null pointer");
    if (N < 0 || A < 1 || A > kMaxAnchors || C < 1 || H0 < 1 || W0 < 1 || H1 < 1 || W1 < 1)
        return fail(B200YOLO_EINVAL, "decode_nms_nhwc: bad shape");
    if (5 + C > kNhwcMaxAttrs)
        return fail(B200YOLO_EUNSUPPORTED, "decode_nms_nhwc: more than %d classes (convert the heads to NCHW)", kNhwcMaxAttrs - 5);
    if (!(iou_thr == iou_thr)) return fail(B200YOLO_EINVAL, "decode_nms_nhwc: NaN threshold");
    const long long cells = (long long)A * H0 * W0 + (long long)A * H1 * W1;
    if (cells > 65535) return fail(B200YOLO_EUNSUPPORTED, "decode_nms_nhwc: more than 65535 cells per image");
    DNParams p;
    memset(&p, 0, sizeof(p));
    fill_head(p.head[0], head0, A, H0, W0, anchor_wh);
    fill_head(p.head[1], head1, A, H1, W1, anchor_wh + 2 * A);
    p.nheads = 2;
    p.N = N; p.A = A; p.C = C; p.attrs = 5 + C;
    p.K = (int)cells;
    p.conf_thr = conf_thr;
    p.iou = make_thr(iou_thr);
    p.out = out; p.out_count = out_count; p.out_idx = out_idx;
    p.nhwc = 1;
    p.wait_inputs = 1;
    return launch_dn<MODE_FUSED>(p, (cudaStream_t)stream);
}

int b200yolo_pairwise(const float *set1, int n1, const float *set2, int n2, int mode, float *out, void *stream) {
    if (n1 < 0 || n2 < 0 || mode < 0 || mode > 4) return fail(B200YOLO_EINVAL, "pairwise: bad argument");
    if (n1 == 0 || n2 == 0) return 0;
    if (!set1 || !set2 || !out) return fail(B200YOLO_EINVAL, "pairwise: null pointer");
    if (((uintptr_t)set1 | (uintptr_t)set2) & 15) return fail(B200YOLO_EINVAL, "pairwise: boxes must be 16-byte aligned");
    dim3 grid((n2 + kPairTile - 1) / kPairTile, (n1 + kPairRows - 1) / kPairRows);
    if (grid.y > 65535) return fail(B200YOLO_EUNSUPPORTED, "pairwise: n1 too large");
    pairwise_kernel<<<grid, kPairTile, 0, (cudaStream_t)stream>>>((const float4 *)set1, n1, (const float4 *)set2, n2,
                                                                   mode, out);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int b200yolo_target_loss(const float *head, int N, int A, int C, int H, int W, const float *anchors_all, int NA,
                         const int *mask, const float *gt, const int *gt_off, int G, float ignore_thr,
                         float iou_thr, int max_gt_per_image, double *sums, int *assign, float *terms, int *status,
                         unsigned char *cell_state, void *workspace, size_t workspace_bytes, void *stream) {
    if (!head || !anchors_all || !mask || !gt_off || !sums || !status)
        return fail(B200YOLO_EINVAL, "target_loss: null pointer");
    if (G > 0 && !gt) return fail(B200YOLO_EINVAL, "target_loss: null gt");
    if (N < 0 || A < 1 || A > kMaxAnchors || NA < A || NA > B200YOLO_MAX_ALL_ANCHORS || C < 1 || C > 4096 || H < 1 ||
        W < 1 || G < 0)
        return fail(B200YOLO_EINVAL, "target_loss: bad shape");
    for (int k = 0; k < A; ++k)
        if (mask[k] < 0 || mask[k] >= NA) return fail(B200YOLO_EINVAL, "target_loss: mask[%d]=%d out of range", k, mask[k]);
    TLParams p;
    memset(&p, 0, sizeof(p));
    p.head = head;
    p.N = N; p.A = A; p.C = C; p.attrs = 5 + C; p.H = H; p.W = W; p.HW = H * W; p.cells = A * H * W;
    p.NA = NA;
    p.invHW = 1.0f / (float)(H * W);
    p.invW = 1.0f / (float)W;
    p.fW = (float)W; p.fH = (float)H;
    for (int a = 0; a < NA; ++a) { p.aw_all[a] = anchors_all[2 * a]; p.ah_all[a] = anchors_all[2 * a + 1]; }
    for (int k = 0; k < A; ++k) p.mask[k] = mask[k];
    p.gt = gt; p.gt_off = gt_off; p.G = G;
    p.ignore_thr = ignore_thr; p.iou_thr = iou_thr;
    // iou < thr <=> inter < thr/(1+thr) * (area_g + area_p); outside [0.01, 1] every cell takes the exact path
    p.ts = (ignore_thr >= 0.01f && ignore_thr <= 1.0f) ? (float)((double)ignore_thr / (1.0 + (double)ignore_thr)) * 1.220703125e-4f : 0.0f;
    p.th16 = (ignore_thr >= 0.01f && ignore_thr <= 1.0f) ? (float)((double)ignore_thr / (1.0 + (double)ignore_thr) * 64.0 * (1.0 - 0.00390625)) : 0.0f;
    p.sums = sums; p.assign = assign; p.terms = terms; p.status = status; p.cell_state = cell_state;
    if (N > 0 && (!workspace || workspace_bytes < b200yolo_target_loss_workspace_bytes(N)))
        return fail(B200YOLO_EINVAL, "target_loss: workspace too small (%zu < %zu)", workspace_bytes,
                    b200yolo_target_loss_workspace_bytes(N));
    p.partial = (double *)workspace;
    p.wait_inputs = g_inputs_ready.load() ? 0 : 1;
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0;
    if (int rc = current_device(&dev)) return rc;
    // shared-memory staging sized for the largest image (the caller's bound, else min(G, 1024))
    int gcap = (max_gt_per_image > 0) ? max_gt_per_image : G;
    if (gcap > kTLMaxGT) gcap = kTLMaxGT;
    if (gcap < 1) gcap = 1;
    p.gcap = (gcap + 3) / 4 * 4;
    const uint32_t smem = tl_smem_bytes(p.cells, p.gcap, A);
    const int lim = smem_optin(dev);
    if ((int)smem > lim)
        return fail(B200YOLO_EUNSUPPORTED, "target_loss: %d cells per image need %u B of shared memory (limit %d B)",
                    p.cells, smem, lim);
    int nsm = 148;
    {
        static std::mutex mu;
        static bool configured[64] = {false};
        static int sm_count[64] = {0};
        std::lock_guard<std::mutex> g(mu);
        if (dev < 64 && !configured[dev]) {
            CUDA_TRY(cudaFuncSetAttribute(target_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
            CUDA_TRY(cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev));
            configured[dev] = true;
        }
        if (dev < 64 && sm_count[dev] > 0) nsm = sm_count[dev];
    }
    // CTAs per image: enough to put ~3 CTAs on every SM when the batch alone cannot, slices of >= 256 cells
    int S = 1;
    if (N > 0) {
        S = (3 * nsm + N - 1) / N;
        const int smax = (p.cells + kTLThreads - 1) / kTLThreads;
        if (S > smax) S = smax;
        if (S > kTLMaxSplit) S = kTLMaxSplit;
        if (S < 1) S = 1;
        static const int env_split = [] { const char *e = getenv("B200YOLO_TL_SPLIT"); return e ? atoi(e) : 0; }();
        if (env_split > 0 && env_split <= smax && env_split <= kTLMaxSplit) S = env_split;  // tuning knob
    }
    p.S = S;
    p.chunk = ((p.cells + S - 1) / S + 31) / 32 * 32;
    // both kernels use programmatic dependent launch: the next call's target_loss_kernel overlaps this call's tail
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = (g_flags.load() & 2) ? 0 : 1;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (N > 0) {
        cfg.gridDim = dim3((unsigned)(N * S));
        cfg.blockDim = dim3(kTLThreads);
        cfg.dynamicSmemBytes = smem;
        CUDA_TRY(cudaLaunchKernelEx(&cfg, target_loss_kernel, p));
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(kTLSums * 32);
    cfg.dynamicSmemBytes = 0;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, target_loss_reduce_kernel, (const double *)p.partial, N * S, sums, status));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

int b200yolo_target_loss_backward(const float *head, int N, int A, int C, int H, int W, const float *anchors_all, int NA,
                                  const int *mask, const float *gt, const int *gt_off, int G, float iou_thr,
                                  int max_gt_per_image, const unsigned char *cell_state, const double *sums,
                                  float iou_weighting, const float *grad_out, float *grad_input, void *stream) {
    if (!head || !anchors_all || !mask || !gt_off || !sums || !cell_state || !grad_input)
        return fail(B200YOLO_EINVAL, "target_loss_backward: null pointer");
    if (G > 0 && !gt) return fail(B200YOLO_EINVAL, "target_loss_backward: null gt");
    if (N < 0 || A < 1 || A > kMaxAnchors || NA < A || NA > B200YOLO_MAX_ALL_ANCHORS || C < 1 || C > 4096 || H < 1 ||
        W < 1 || G < 0)
        return fail(B200YOLO_EINVAL, "target_loss_backward: bad shape");
    for (int k = 0; k < A; ++k)
        if (mask[k] < 0 || mask[k] >= NA) return fail(B200YOLO_EINVAL, "target_loss_backward: mask[%d]=%d out of range", k, mask[k]);
    if (N == 0) return 0;
    TLParams p;
    memset(&p, 0, sizeof(p));
    p.head = head;
    p.N = N; p.A = A; p.C = C; p.attrs = 5 + C; p.H = H; p.W = W; p.HW = H * W; p.cells = A * H * W;
    p.NA = NA;
    p.invHW = 1.0f / (float)(H * W);
    p.invW = 1.0f / (float)W;
    p.fW = (float)W; p.fH = (float)H;
    for (int a = 0; a < NA; ++a) { p.aw_all[a] = anchors_all[2 * a]; p.ah_all[a] = anchors_all[2 * a + 1]; }
    for (int k = 0; k < A; ++k) p.mask[k] = mask[k];
    p.gt = gt; p.gt_off = gt_off; p.G = G;
    p.iou_thr = iou_thr;
    p.sums = const_cast<double *>(sums);
    p.cell_state_in = cell_state;
    p.grad_out = grad_out;
    p.grad_input = grad_input;
    p.iou_weighting = iou_weighting;
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0;
    if (int rc = current_device(&dev)) return rc;
    int gcap = (max_gt_per_image > 0) ? max_gt_per_image : G;
    if (gcap > kTLMaxGT) gcap = kTLMaxGT;
    if (gcap < 1) gcap = 1;
    p.gcap = (gcap + 3) / 4 * 4;
    const uint32_t smem = tl_smem_bytes(p.cells, p.gcap, A, true);
    const int lim = smem_optin(dev);
    if ((int)smem > lim)
        return fail(B200YOLO_EUNSUPPORTED, "target_loss_backward: %d cells per image need %u B of shared memory (limit %d B)",
                    p.cells, smem, lim);
    int nsm = 148;
    {
        static std::mutex mu;
        static bool configured[64] = {false};
        static int sm_count[64] = {0};
        std::lock_guard<std::mutex> g(mu);
        if (dev < 64 && !configured[dev]) {
            CUDA_TRY(cudaFuncSetAttribute(target_loss_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
            CUDA_TRY(cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev));
            configured[dev] = true;
        }
        if (dev < 64 && sm_count[dev] > 0) nsm = sm_count[dev];
    }
    int S = (4 * nsm + N - 1) / N;
    const int smax = (p.cells + kTLThreads - 1) / kTLThreads;
    if (S > smax) S = smax;
    if (S > kTLMaxSplit) S = kTLMaxSplit;
    if (S < 1) S = 1;
    p.S = S;
    p.chunk = ((p.cells + S - 1) / S + 31) / 32 * 32;
    // programmatic dependent launch, as in the forward: the kernel after this one may start during its tail
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = (g_flags.load() & 2) ? 0 : 1;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(N * S));
    cfg.blockDim = dim3(kTLThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, target_loss_backward_kernel, p));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static bool seg_vectorisable(const float *input, const float *truth, int C, int HW) {
    return C >= 1 && C <= 4 && (HW % 4) == 0 && !(((uintptr_t)input | (uintptr_t)truth) & 15);
}

static int seg_grid(long long total, int dev) {
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    long long g = (total + kSegThreads - 1) / kSegThreads;
    const long long cap = (long long)nsm * 8;   // 8 CTAs of 256 threads per SM, grid-stride above that
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

size_t b200yolo_seg_loss_workspace_bytes(void) { return (size_t)148 * 8 * 4 * kSegSums * sizeof(double); }

int b200yolo_seg_loss(const float *input, const float *truth, int N, int C, int H, int W, double *sums, void *workspace,
                      size_t workspace_bytes, void *stream) {
    if (!input || !truth || !sums || !workspace) return fail(B200YOLO_EINVAL, "seg_loss: null pointer");
    if (N < 1 || C < 1 || H < 1 || W < 1 || (long long)N * ((H * W + kSegThreads - 1) / kSegThreads) > 0x7fffffffLL)
        return fail(B200YOLO_EINVAL, "seg_loss: bad shape");
    int dev = 0;
    if (int rc = current_device(&dev)) return rc;
    SegParams p;
    memset(&p, 0, sizeof(p));
    p.input = input; p.truth = truth; p.C = C; p.HW = H * W;
    p.total = (long long)N * C * H * W;
    p.partial = (double *)workspace; p.sums = sums;
    const bool vec = seg_vectorisable(input, truth, C, H * W);
    const int tile = kSegThreads * (vec ? 4 : 1);
    p.items = N * ((H * W + tile - 1) / tile);
    const int grid = seg_grid((long long)p.items * kSegThreads, dev);
    if (workspace_bytes < (size_t)grid * kSegSums * sizeof(double)) return fail(B200YOLO_EINVAL, "seg_loss: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (!vec) seg_loss_kernel<0><<<grid, kSegThreads, 0, st>>>(p);
    else if (C == 1) seg_loss_kernel<1><<<grid, kSegThreads, 0, st>>>(p);
    else if (C == 2) seg_loss_kernel<2><<<grid, kSegThreads, 0, st>>>(p);
    else if (C == 3) seg_loss_kernel<3><<<grid, kSegThreads, 0, st>>>(p);
    else seg_loss_kernel<4><<<grid, kSegThreads, 0, st>>>(p);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
    seg_loss_reduce_kernel<<<1, kSegSums * 32, 0, (cudaStream_t)stream>>>(p.partial, grid, sums);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int b200yolo_seg_loss_backward(const float *input, const float *truth, int N, int C, int H, int W, const float *grad_out,
                               float *grad_input, void *stream) {
    if (!input || !truth || !grad_input) return fail(B200YOLO_EINVAL, "seg_loss_backward: null pointer");
    if (N < 1 || C < 1 || H < 1 || W < 1) return fail(B200YOLO_EINVAL, "seg_loss_backward: bad shape");
    int dev = 0;
    if (int rc = current_device(&dev)) return rc;
    SegParams p;
    memset(&p, 0, sizeof(p));
    p.input = input; p.truth = truth; p.C = C; p.HW = H * W;
    p.total = (long long)N * C * H * W;
    p.grad_out = grad_out; p.grad_input = grad_input;
    const bool vec = seg_vectorisable(input, truth, C, H * W) && !((uintptr_t)grad_input & 15);
    const int tile = kSegThreads * (vec ? 4 : 1);
    p.items = N * ((H * W + tile - 1) / tile);
    const int grid = seg_grid((long long)p.items * kSegThreads, dev);
    cudaStream_t st = (cudaStream_t)stream;
    if (!vec) seg_loss_backward_kernel<0><<<grid, kSegThreads, 0, st>>>(p);
    else if (C == 1) seg_loss_backward_kernel<1><<<grid, kSegThreads, 0, st>>>(p);
    else if (C == 2) seg_loss_backward_kernel<2><<<grid, kSegThreads, 0, st>>>(p);
    else if (C == 3) seg_loss_backward_kernel<3><<<grid, kSegThreads, 0, st>>>(p);
    else seg_loss_backward_kernel<4><<<grid, kSegThreads, 0, st>>>(p);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int b200yolo_seg_sigmoid(const float *input, long long count, float *out, void *stream) {
    if (count < 0) return fail(B200YOLO_EINVAL, "seg_sigmoid: bad count");
    if (count == 0) return 0;
    if (!input || !out) return fail(B200YOLO_EINVAL, "seg_sigmoid: null pointer");
    int dev = 0;
    if (int rc = current_device(&dev)) return rc;
    SegParams p;
    memset(&p, 0, sizeof(p));
    p.input = input; p.out = out; p.total = count;
    seg_sigmoid_kernel<<<seg_grid(count, dev), kSegThreads, 0, (cudaStream_t)stream>>>(p);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static size_t map_align(size_t v) { return (v + 255) / 256 * 256; }

size_t b200yolo_map_eval_workspace_bytes(int D, int T, int N, int n_classes) {
    const size_t Cf = (size_t)(n_classes > 1 ? n_classes - 1 : 1);
    return map_align((size_t)(T > 0 ? T : 1)) + map_align(2 * Cf * sizeof(int)) + map_align((size_t)(D > 0 ? D : 1)) +
           map_align(Cf * (size_t)(N > 0 ? N : 1) * sizeof(int)) + map_align((2 * (size_t)(D > 0 ? D : 1) + Cf) * sizeof(unsigned long long));
}

int b200yolo_map_eval(const float *det_boxes, const int *det_labels, const float *det_scores, const int *det_off, int D,
                      const float *true_boxes, const int *true_labels, const unsigned char *true_difficult,
                      const int *true_off, int T, int N, int n_classes, float iou_thr, const float *recall_thresholds,
                      int n_thresholds, float *ap, float *tp_sum, float *fp_sum, void *workspace, size_t workspace_bytes,
                      void *stream) {
    if (!det_off || !true_off || !ap || !tp_sum || !fp_sum || !workspace || !recall_thresholds)
        return fail(B200YOLO_EINVAL, "map_eval: null pointer");
    if ((D > 0 && (!det_boxes || !det_labels || !det_scores)) || (T > 0 && (!true_boxes || !true_labels || !true_difficult)))
        return fail(B200YOLO_EINVAL, "map_eval: null data pointer");
    if (N < 0 || D < 0 || T < 0 || n_classes < 2 || n_thresholds < 1 || n_thresholds > kMapMaxThr)
        return fail(B200YOLO_EINVAL, "map_eval: bad argument");
    if (D >= (1 << 30)) return fail(B200YOLO_EUNSUPPORTED, "map_eval: more than 2^30 detections");
    if ((D > 0 && ((uintptr_t)det_boxes & 15)) || (T > 0 && ((uintptr_t)true_boxes & 15)))
        return fail(B200YOLO_EINVAL, "map_eval: boxes must be 16-byte aligned");
    if (workspace_bytes < b200yolo_map_eval_workspace_bytes(D, T, N, n_classes))
        return fail(B200YOLO_EINVAL, "map_eval: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    MapParams p;
    memset(&p, 0, sizeof(p));
    p.det_boxes = det_boxes; p.det_labels = det_labels; p.det_scores = det_scores; p.det_off = det_off;
    p.true_boxes = true_boxes; p.true_labels = true_labels; p.true_difficult = true_difficult; p.true_off = true_off;
    p.N = N; p.Cf = n_classes - 1; p.iou_thr = iou_thr;
    p.n_thr = n_thresholds;
    for (int i = 0; i < n_thresholds; ++i) p.thr[i] = recall_thresholds[i];
    const size_t Cf = (size_t)p.Cf;
    unsigned char *w = (unsigned char *)workspace;
    p.detected = w; w += map_align((size_t)(T > 0 ? T : 1));
    p.M = (int *)w; p.n_easy = p.M + Cf; w += map_align(2 * Cf * sizeof(int));
    const size_t zero_bytes = (size_t)(w - (unsigned char *)workspace);   // detected[], M[], n_easy[] start at zero
    p.flags = w; w += map_align((size_t)(D > 0 ? D : 1));
    p.cnt = (int *)w; w += map_align(Cf * (size_t)(N > 0 ? N : 1) * sizeof(int));
    p.keys = (unsigned long long *)w;
    p.ap = ap; p.tp_sum = tp_sum; p.fp_sum = fp_sum;
    CUDA_TRY(cudaMemsetAsync(workspace, 0, zero_bytes, st));
    if (N > 0) {
        map_match_kernel<<<N, kMapMatchThreads, 0, st>>>(p);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        CUDA_TRY(cudaGetLastError());
    }
    map_class_kernel<<<p.Cf, kMapClassThreads, 0, st>>>(p);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

size_t b200yolo_decode_nms_large_workspace_bytes(int N, int cells_per_image) {
    if (N < 0 || cells_per_image < 0) return 0;
    return (size_t)N * (size_t)cells_per_image * 2 * sizeof(float4);
}

int b200yolo_decode_nms_large(const float *head0, const float *head1, int N, int A, int C, int H0, int W0, int H1,
                              int W1, const float *anchor_wh, float conf_thr, double iou_thr, float *out,
                              int *out_count, int *out_idx, void *workspace, size_t workspace_bytes, void *stream) {
    if (!head0 || !head1 || !anchor_wh || !out || !out_count) return fail(B200YOLO_EINVAL, "decode_nms_large: null pointer");
    if (N < 0 || A < 1 || A > kMaxAnchors || C < 1 || C > 4096 || H0 < 1 || W0 < 1 || H1 < 1 || W1 < 1)
        return fail(B200YOLO_EINVAL, "decode_nms_large: bad shape");
    if (!(iou_thr == iou_thr)) return fail(B200YOLO_EINVAL, "decode_nms_large: NaN threshold");
    const long long cells = (long long)A * H0 * W0 + (long long)A * H1 * W1;
    if (cells > kLargeMaxKeys)
        return fail(B200YOLO_EUNSUPPORTED, "decode_nms_large: %lld cells per image (limit %d)", cells, kLargeMaxKeys);
    if (N == 0) return 0;
    if (!workspace || ((uintptr_t)workspace & 15) || workspace_bytes < b200yolo_decode_nms_large_workspace_bytes(N, (int)cells))
        return fail(B200YOLO_EINVAL, "decode_nms_large: workspace missing, misaligned or too small (%zu < %zu)", workspace_bytes,
                    b200yolo_decode_nms_large_workspace_bytes(N, (int)cells));
    LargeParams p;
    memset(&p, 0, sizeof(p));
    fill_head(p.head[0], head0, A, H0, W0, anchor_wh);
    fill_head(p.head[1], head1, A, H1, W1, anchor_wh + 2 * A);
    p.N = N; p.A = A; p.C = C; p.attrs = 5 + C;
    p.K = (int)cells;
    p.conf_thr = conf_thr;
    p.iou = make_thr(iou_thr);
    p.rec = (float4 *)workspace;
    p.out = out; p.out_count = out_count; p.out_idx = out_idx;
    return launch_large<0>(p, (cudaStream_t)stream, "decode_nms_large");
}

int b200yolo_compact_rows(const float *dets, const int *count, int N, int K, float *packed, int *offsets, void *stream) {
    if (N < 0 || K < 1) return fail(B200YOLO_EINVAL, "compact_rows: bad shape");
    if (!offsets || (N > 0 && (!dets || !count || !packed))) return fail(B200YOLO_EINVAL, "compact_rows: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    compact_offsets_kernel<<<1, 1024, 0, st>>>(count, N, K, offsets);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
    if (N > 0) {
        compact_rows_kernel<<<N, 128, 0, st>>>(dets, count, offsets, K, packed);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        CUDA_TRY(cudaGetLastError());
    }
    return 0;
}

int b200yolo_loss_finalize_dev(const double *sums, float iou_weighting, float *result, void *stream) {
    if (!sums || !result) return fail(B200YOLO_EINVAL, "loss_finalize_dev: null pointer");
    loss_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sums, iou_weighting, result);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

size_t b200yolo_target_loss_workspace_bytes(int N) { return (size_t)(N > 0 ? N : 1) * kTLMaxSplit * kTLSums * sizeof(double); }

int b200yolo_loss_finalize(const double *s, float iou_weighting, double *r) {
    if (!s || !r) return fail(B200YOLO_EINVAL, "loss_finalize: null pointer");
    const double count = s[B200YOLO_S_NASSIGN];
    const double l_dense = s[B200YOLO_S_SQW] / s[B200YOLO_S_W];                 // yolo_loss.py:54-58
    const double l_iou = count > 0 ? s[B200YOLO_S_IOU_SQ] / count : 0.0;        // :222-224 (quirk Q6)
    r[0] = l_dense + l_iou * (double)iou_weighting;                             // :234
    if (count > 0) {                                                            // :170-175
        r[1] = s[B200YOLO_S_RECALL] / count;
        r[2] = s[B200YOLO_S_IOU] / count;
        r[3] = s[B200YOLO_S_OBJ] / count;
        r[4] = (s[B200YOLO_S_CONF_ALL] - s[B200YOLO_S_OBJ]) / (s[B200YOLO_S_NCELLS] - count);
        r[5] = s[B200YOLO_S_CLS] / count;
    } else {
        r[1] = r[2] = r[3] = r[4] = r[5] = 0.0;                                 // :176-177
    }
    r[6] = s[B200YOLO_S_NIMG] > 0 ? count / s[B200YOLO_S_NIMG] : 0.0;           // :178
    return 0;
}

}  // extern "C"

#include "host_pipeline.inl"
