// iou.cuh -- util.bbox_iou (src/util.py:120-153) with the reference's exact fp32 arithmetic: every tensor
// op of the reference is a separate rounding, so each step here is an explicitly rounded intrinsic
// (no fused multiply-add contraction), and min / max / clamp propagate NaN like torch's.
#pragma once
#include "common.cuh"

namespace rtod {

__device__ __forceinline__ float nan_max(float a, float b) {   // torch.max propagates NaN
    return (a != a || b != b) ? __int_as_float(0x7fc00000) : fmaxf(a, b);
}
__device__ __forceinline__ float nan_min(float a, float b) {
    return (a != a || b != b) ? __int_as_float(0x7fc00000) : fminf(a, b);
}
__device__ __forceinline__ float clamp_min0(float v) {         // torch.clamp(min=0) keeps NaN
    return (v != v) ? v : fmaxf(v, 0.0f);
}
__device__ __forceinline__ float box_area(float x1, float y1, float x2, float y2) {
    return __fmul_rn(__fadd_rn(__fsub_rn(x2, x1), 1.0f), __fadd_rn(__fsub_rn(y2, y1), 1.0f));
}
// src/util.py:138-151
__device__ __forceinline__ float iou_exact(float ax1, float ay1, float ax2, float ay2, float aarea,
                                           float bx1, float by1, float bx2, float by2,
                                           float barea) {
    const float left = nan_max(ax1, bx1), top = nan_max(ay1, by1);
    const float right = nan_min(ax2, bx2), bottom = nan_min(ay2, by2);
    const float iw = clamp_min0(__fadd_rn(__fsub_rn(right, left), 1.0f));
    const float ih = clamp_min0(__fadd_rn(__fsub_rn(bottom, top), 1.0f));
    const float inter = __fmul_rn(iw, ih);
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(aarea, barea), inter));
}

}  // namespace rtod
