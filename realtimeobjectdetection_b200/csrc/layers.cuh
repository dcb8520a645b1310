// layers.cuh -- activation views and launchers of the non-GEMM layer kernels (layers.cu) and the
// two convolution back ends (conv_tc.cu = tcgen05 implicit GEMM, conv_simt.cu = validation).
#pragma once
#include "common.cuh"

namespace rtod {

// NHWC activation view.  `ptr` points at channel 0 of the view's channel slice inside a buffer
// whose pixels are `pitch` elements apart (pitch > C when the view is a slice of a route/concat
// buffer).  Elements are bf16, or fp32 for detection-head logits.
struct Act {
    void* ptr;
    int C, pitch, H, W;
    int fp32;
    int f16;                    // 16-bit elements are fp16 (1) or bf16 (0); ignored when fp32
};

// Packed weight layout (fold_pack_kernel): K-BLOCK-major, [K / PK][rows][PK] 16-bit elements, rows = Cout_pad (two-term
// weights: hi rows then lo rows, 2 * Cout_pad), K = (ky*ks + kx)*Cin + c, PK = weight_pack_k(Cin, K) = the K tile of the
// tcgen05 kernels.  Any run of consecutive rows of one K block -- every weight tile the kernels load -- is ONE contiguous
// block of memory: a TMA box whose 128-byte rows are adjacent streams at ~74 B/clk/SM from L2, rows that lie a pitch
// apart (the plain [rows][K] layout) at ~49 (tools/tma_multicast_probe.cu).
__host__ __device__ inline int weight_pack_k(int cin, int K) { return cin % 64 == 0 ? 64 : (cin % 32 == 0 ? 32 : (cin % 16 == 0 ? 16 : K)); }
__host__ __device__ inline size_t packed_weight_index(int n, int k, int rows, int PK) {
    return ((size_t)(k / PK) * (size_t)rows + (size_t)n) * (size_t)PK + (size_t)(k % PK);
}

struct ConvArgs {
    Act in, out;
    const void* w;              // packed weights (see packed_weight_index), BN folded
    const float* bias;          // [Cout_pad] fp32 (BN folded)
    const void* res;            // optional shortcut operand, same pixels/channels as out
    int res_pitch;
    int B, Cin, Cout, Cout_pad, ks, stride, pad, leaky;
    int K;                      // ks*ks*Cin
    int w_split;                // weights stored as hi rows [0, Cout_pad) + lo rows [Cout_pad, 2*Cout_pad)
    float* split_scratch;       // split-K scratch shared by all layers of a plan (or null)
    size_t split_scratch_bytes;
    int* split_count;           // zeroed counters, split_count_n entries
    int split_count_n;
};

int launch_stem_conv(const float* x_nchw, int B, int Cin, int H, int W, const float* w_f32,
                     const float* bias, int Cout, int ks, int stride, int pad, int leaky, Act out,
                     cudaStream_t stream);
int launch_stem_tma(const float* x_nchw, int B, int H, int W, const float* w_f32, const float* bias, int Cout,
                    int leaky, Act out, cudaStream_t stream);   // stem.cu
// stem_tc.cu: tcgen05 stem for fp16 storage; fp32 NCHW frames or uint8 planes (value / 255 folded into the weights)
int launch_stem_tc(const float* x_f32, const unsigned char* x_u8, int B, int H, int W, const float* w_f32, const float* bias,
                   int Cout, int leaky, Act out, cudaStream_t stream);
int launch_nchw_to_nhwc(const float* x_nchw, int B, Act out, cudaStream_t stream);
int launch_nhwc_to_nchw(Act in, int B, float* out_nchw, cudaStream_t stream);
int launch_maxpool(Act in, Act out, int B, int size, int stride, cudaStream_t stream);
int launch_upsample2x(Act in, Act out, int B, cudaStream_t stream);
int launch_copy(Act in, Act out, int B, cudaStream_t stream);
int launch_add(Act a, Act b, Act out, int B, cudaStream_t stream);
// BN fold + K-major bf16 re-layout (and the fp32 [Cout][Cin*ks*ks] copy the stem kernel reads)
int launch_fold_pack(const float* w, const float* bias, const float* gamma, const float* beta,
                     const float* mean, const float* var, float eps, int Cout, int Cin, int ks,
                     int Cout_pad, int f16, int w_split, void* w_packed, float* w_f32_or_null,
                     float* bias_out, cudaStream_t stream);
int launch_conv_simt(const ConvArgs& a, cudaStream_t stream);

}  // namespace rtod
