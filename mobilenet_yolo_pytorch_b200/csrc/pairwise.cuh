// pairwise.cuh -- utils/iou.py: find_intersection / find_union /
// find_jaccard_overlap as one (n1, n2) kernel; modes 3 / 4 evaluate YOLOLoss.box_giou / box_ciou
// (models/yolo_loss.py:295-317 / 257-293; value = iou - term, box1 = row, box2 = column).  set_2 is staged in shared memory
// in tiles; each thread owns one set_1 row and streams a coalesced output row
// segment (consecutive threads -> consecutive columns).
#pragma once
#include "common.cuh"
#include "target_loss.cuh"

namespace b200yolo {

constexpr int kPairTile = 256;  // columns per CTA (threads.x), rows per CTA = kPairRows
constexpr int kPairRows = 16;

// grid: (ceil(n2/kPairTile), ceil(n1/kPairRows)); block: kPairTile threads
__global__ void __launch_bounds__(kPairTile) pairwise_kernel(const float4 *__restrict__ s1, int n1,
                                                             const float4 *__restrict__ s2, int n2, int mode,
                                                             float *__restrict__ out) {
    __shared__ float4 rows[kPairRows];
    __shared__ float rarea[kPairRows];
    const int col = blockIdx.x * kPairTile + threadIdx.x;
    const int r0 = blockIdx.y * kPairRows;
    if (threadIdx.x < kPairRows && r0 + threadIdx.x < n1) {
        const float4 a = __ldg(s1 + r0 + threadIdx.x);
        rows[threadIdx.x] = a;
        rarea[threadIdx.x] = box_area(a);  // iou.py:39
    }
    __syncthreads();
    if (col >= n2) return;
    const float4 b = __ldg(s2 + col);
    const float barea = box_area(b);  // iou.py:40
    const int nr = min(kPairRows, n1 - r0);
#pragma unroll 4
    for (int r = 0; r < nr; ++r) {
        const float inter = pair_inter(rows[r], b);  // iou.py:4-13
        float v = inter;
        if (mode == 1 || mode == 2) {
            const float u = pair_union(rarea[r], barea, inter);  // iou.py:44
            v = (mode == 1) ? u : __fdiv_rn(inter, u);           // iou.py:49
        } else if (mode >= 3) {
            float iou;
            v = (mode == 3) ? tl_box_giou(rows[r], b, &iou) : tl_box_ciou(rows[r], b, &iou);
        }
        out[(size_t)(r0 + r) * n2 + col] = v;
    }
}

}  // namespace b200yolo
