// conv_tc.cuh -- host interface of the tcgen05 implicit-GEMM convolution (conv_tc.cu)
#pragma once
#include <cuda.h>

#include "layers.cuh"

namespace rtod {

// Everything one launch needs; built once at plan-bind time (the tensor maps embed addresses).
struct alignas(64) ConvTcParams {
    CUtensorMap tmA;            // activations: 2-D tiled {C, M} (1x1) or 4-D im2col {C, W, H, N} (3x3)
    CUtensorMap tmB;            // weights: 3-D tiled {PK, rows, K / PK}, K-block-major (layers.cuh)
    CUtensorMap tmOut;          // output: 2-D tiled {Cout, M}, NHWC bf16 (fp32 for head logits)
    CUtensorMap tmRes;          // shortcut operand: 2-D tiled {Cout, M} bf16 (valid iff has_res)
    const float* bias;          // [Cout_pad]
    int* err_flag;              // device-side failure flag (pipeline time-out)
    int out_fp32, has_res;
    int b_resident;             // whole [BN x K] weight matrix lives in smem (single N tile, small K*BN)
    int stage_bufs;             // staging slices per epilogue warp (2; 1 when shared memory is tight)
    int epi_warps;              // 4 or 8 epilogue warps (kernel template argument)
    int a_producers;            // 1 or 2 activation producer threads (2: they alternate k-blocks)
    int dbg;                    // bring-up switches (RTOD_PAIR_MODE / RTOD_CLK_DBG), 0 in production
    int ecols;                  // channels per epilogue chunk (one staging row: <= 64 bf16 / 32 fp32)
    int M, Cout;                // output pixels, real channels
    int leaky;
    int ks, cchunks;            // kernel size, Cin / BK
    int BK, BN, stages;         // K tile (16/32/64 -> 32B/64B/128B swizzle), N tile, pipeline depth
    int Ho, Wo, stride, pad;    // im2col traversal
    int tmem_cols;              // 2 * BN rounded up to a power of two
    int m_tiles, total_tiles;   // tile = m_tile + m_tiles * n_tile
    uint32_t idesc;             // tcgen05 instruction descriptor (bf16 x bf16 -> fp32, M=128, N=BN)
    const void* pf_ptr;         // next convolution's weights: prefetched into L2 while this layer runs (small batches)
    unsigned long long pf_bytes;
    int f16;                    // 16-bit storage type of activations / weights: 1 = fp16, 0 = bf16
    int w_split;                // 1: weights are hi + lo (two fp16 terms, rows [Cout_pad, 2*Cout_pad) of tmB hold lo):
                                // two MMAs per K step, for layers whose tensor time hides behind their HBM time
    int cout_pad;               // row offset of the lo half in tmB
    int w_cat;                  // two-term weights as ONE MMA of N = 2*BN per K step (hi and lo accumulate in separate
                                // TMEM columns, added by the epilogue): half the MMA instructions; needs 2*BN <= 256
    int acc_cols;               // TMEM columns of one accumulator buffer: BN << w_cat
    int subs;                   // pipelines (warp sets) per CTA: 1, or 2 sharing the resident weights
    // row mode (3x3 / stride 1, Cin == BK): a tile is a 128-pixel segment of ONE image row; per filter row ky one TILED
    // TMA load brings the segment's 130 input pixels (halo and image border zero-filled) and the three kx taps are
    // three row-shifted views of that slab -- a third of the TMA rows and of the L2 -> SM bytes of the im2col gather
    int row_mode, segs;         // segs = ceil(Wo / 128) segments per image row
    uint32_t slab_bytes;        // one staged slab: 130 rows, rounded up to 1024
    uint32_t fd_segs[3], fd_ho[3];
    int a_prefetch;             // activation tiles prefetched into L2 this many of the CTA's tiles ahead (0 = off): layers that
                                // stream their input from HBM are bound by latency x bytes in flight, and the ring is small
    int pack_k, pack_shift;     // K block of the packed weight layout (16 / 32 / 64) and its log2
    int lo_col;                 // column distance from a value's hi term to its lo term (w_cat): BN, pair kernel BN / 2
    int split_k, split_shift;   // K slices per output tile (power of two, 1 = off) and log2 of it
    float* split_scratch;       // [total_tiles][split_k][128][BN] fp32 partial accumulators
    int* split_count;           // [total_tiles] arrival counters (zero between forwards)
    uint32_t fd_mtiles[3], fd_wo[3], fd_howo[3];   // FastDiv {mul, shift, d} for m_tiles, Wo, Ho*Wo (tc_ptx.cuh)
};

struct ConvTcChoice {           // one launch configuration of conv_tc_kernel (or the CTA-pair kernel)
    int ctas, resident, sbufs, pair;
    int bn;                     // N tile (0 = heuristic)
    int split;                  // K slices per tile (0 = by shape)
    int ew, ap;                 // epilogue warps (4 / 8), activation producer threads (1 / 2); 0 = heuristic
    int subs;                   // 2 = dual pipeline (two warp sets sharing resident weights in one CTA); 0 / 1 = single
    int row;                    // 1 = row mode (ConvTcParams::row_mode)
};

struct ConvTcLaunch {           // host side: kernel parameters + launch geometry
    ConvTcChoice choice;
    ConvTcParams p;
    int patch;                  // 0: conv_tc_kernel(p), 2: conv_pair_kernel(p)
    dim3 grid;
    uint32_t smem_bytes;
};

// conv_pair.cu: cta_group::2 tiles (256 x 256 per CTA pair) for Cout % 256 == 0
bool conv_pair_eligible(const ConvArgs& a);
int conv_pair_prepare(const ConvArgs& a, int* err_flag, ConvTcLaunch* launch);
int conv_pair_launch(const ConvTcLaunch& launch, cudaStream_t stream);

// K slices per output tile for this shape (1 = none); deterministic
int conv_split_factor(const ConvArgs& a);
// true if the tensor-core kernel tiles this convolution
bool conv_tc_supported(const ConvArgs& a);
bool conv_tc_row_eligible(const ConvArgs& a);
// fills `p` (encodes the TMA descriptors); `err_flag` is a device int
int conv_tc_prepare(const ConvArgs& a, int* err_flag, ConvTcLaunch* launch, const ConvTcChoice* force = nullptr);
// conv_tc_prepare + timing of every configuration that fits (plan-bind time); keeps the fastest
int conv_tc_autotune(const ConvArgs& a, int* err_flag, ConvTcLaunch* launch, cudaStream_t stream);
int conv_tc_launch(const ConvTcLaunch& launch, cudaStream_t stream);

}  // namespace rtod
