// decode.cuh -- parameter blocks of the YOLO decode kernels (decode.cu)
#pragma once
#include "common.cuh"

namespace rtod {

constexpr int kMaxHeads = 4;

struct DecodeAnchors {                 // anchors already divided by the stride (fp32(a / stride))
    float w[RTOD_MAX_ANCHORS];
    float h[RTOD_MAX_ANCHORS];
};

struct DecodeHeads {                   // all yolo heads of one plan, in cfg order
    int count;
    const float* raw[kMaxHeads];       // fp32 logits, pixel-major [B*G*G, pitch]
    int pitch[kMaxHeads];              // floats per pixel row (>= A*(5+C))
    int grid[kMaxHeads];               // G
    int num_anchors[kMaxHeads];        // A
    int row_base[kMaxHeads];           // first row of this head in the [N] axis
    float stride[kMaxHeads];           // inp_dim / G
    float anchor_w[kMaxHeads][RTOD_MAX_ANCHORS];
    float anchor_h[kMaxHeads][RTOD_MAX_ANCHORS];
    int* err_flag;                     // the plan's device failure block (tc_ptx.cuh: raise_device_error), may be null
};

int launch_decode_heads(const DecodeHeads& heads, int B, int N, int L, int train, float* pred,
                        cudaStream_t stream);

}  // namespace rtod
