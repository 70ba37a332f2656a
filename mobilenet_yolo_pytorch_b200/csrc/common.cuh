// common.cuh -- shared device helpers for the b200yolo kernels (sm_100a).
//
// Arithmetic rules: every reference op is one fp32
// rounding, so this translation unit is compiled with -fmad=false and without
// --use_fast_math; FMA is used only where written explicitly (guard-band
// filters whose exact fallback decides the close calls).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200yolo {

constexpr unsigned kFullMask = 0xffffffffu;

// yolo_loss.py:19 / :187-189 -- 1/(1+exp(-x)) with the accurate expf and an
// IEEE (round-to-nearest) divide.
__device__ __forceinline__ float sigmoid_f(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

// Inference decode (YOLOLoss.get_pred_boxes) uses the SFU forms directly: ex2.approx.ftz and
// rcp.approx.ftz are one MUFU instruction each (no range fix-ups; results below 2^-126 flush to
// zero), <= ~4 ulp + |x|*6e-8 relative, well inside the 1e-5 contract on decoded floats.  The decode
// phase is issue-bound, and the IEEE forms above cost ~20 instructions more per transcendental.
// The training path (target_loss.cuh) keeps the IEEE forms.
__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_fast(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float exp_fast(float x) { return ex2_fast(__fmul_rn(x, 1.4426950408889634f)); }
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_fast(__fadd_rn(1.0f, ex2_fast(__fmul_rn(x, -1.4426950408889634f)))); }

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// Order-preserving map float -> uint32 (ascending).  NaN sorts above +inf (torch
// sort treats NaN as the largest value); -0.0 == +0.0.
__device__ __forceinline__ uint32_t float_order_key(float s) {
    uint32_t b = __float_as_uint(s);
    if (s != s) return 0xffffffffu;
    if (b == 0x80000000u) b = 0u;
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// xyxy box area exactly as torchvision nms / utils/iou.py:39-40: (x2-x1)*(y2-y1)
__device__ __forceinline__ float box_area(const float4 &b) { return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y)); }

// torchvision nms_kernel semantics (call site utils/box.py:28):
//   inter = max(0, min(x2)-max(x1)) * max(0, min(y2)-max(y1))
//   iou   = inter / ((area_a + area_b) - inter)        (fp32, IEEE divide)
//   suppress iff (double)iou > thr
// The hot loop (decode_nms.cuh, pair masks) decides most pairs without the divide:
// with t = thr/(1+thr), iou > thr <=> inter > t*(area_a+area_b) in exact arithmetic,
// and every rounding on either side is < 5e-7 relative, so when the two sides differ
// by more than 1e-5 relative the cheap comparison IS the exact answer; closer calls,
// degenerate boxes (area not in [1e-20, 1e20], |coordinate| >= 4096, NaN) and thr
// outside [0.01, 1] are decided by this exact routine.
struct IouThr {
    double thr;
    float ts;     // thr/(1+thr) * 2^-13 (large_nms.cuh: the fp32 pair loop works on widths scaled into [0,1))
    float tf;     // thr/(1+thr) rounded to fp32 (decode_nms.cuh, pair_decide)
    float th;     // thr/(1+thr) * 64 * (1 - 2^-8): scale and safety margin of the fp16 prefilter (h16_store)
    int fast_ok;  // thr in [0.01, 1]
};

__device__ __forceinline__ bool nms_suppress_exact(const float4 &a, float area_a, const float4 &b, float area_b,
                                                   double thr) {
    float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
    float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    float w = fmaxf(0.0f, __fsub_rn(xx2, xx1)), h = fmaxf(0.0f, __fsub_rn(yy2, yy1));
    float inter = __fmul_rn(w, h);
    float u = __fsub_rn(__fadd_rn(area_a, area_b), inter);
    float ovr = __fdiv_rn(inter, u);
    return (double)ovr > thr;
}

// utils/iou.py:4-13 find_intersection for one pair (clamp(min=0) keeps NaN)
__device__ __forceinline__ float pair_inter(const float4 &a, const float4 &b) {
    float lx = fmaxf(a.x, b.x), ly = fmaxf(a.y, b.y);
    float ux = fminf(a.z, b.z), uy = fminf(a.w, b.w);
    float dx = __fsub_rn(ux, lx), dy = __fsub_rn(uy, ly);
    dx = (dx < 0.0f) ? 0.0f : dx;
    dy = (dy < 0.0f) ? 0.0f : dy;
    return __fmul_rn(dx, dy);
}
// utils/iou.py:44  union = area1 + area2 - inter
__device__ __forceinline__ float pair_union(float area_a, float area_b, float inter) {
    return __fsub_rn(__fadd_rn(area_a, area_b), inter);
}

// warp inclusive scan (int)
__device__ __forceinline__ int warp_inclusive_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(kFullMask, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

}  // namespace b200yolo
