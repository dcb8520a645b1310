// decode.cu -- util.predict_transform (src/util.py:175-239) on the GPU.
//
// Two entry points share one element transform:
//   * yolo_decode_nchw_kernel: the drop-in for predict_transform(): an NCHW fp32 head
//     [B, A*(5+C), G, G] is transposed through shared memory (coalesced 128-byte reads along
//     the cell axis, coalesced writes along the attribute axis) into [B, G*G*A, 5+C];
//   * yolo_decode_heads_kernel: used inside the forward plan: the head convolutions leave
//     their fp32 logits in NHWC (pixel-major) buffers, which is already the row order of
//     the prediction tensor, so ONE launch decodes all yolo heads element-wise and writes
//     the concatenated [B, N, 5+C] tensor (replaces predict_transform x3 + torch.cat x2,
//     src/darknet.py:226-247).
//
// Exact op order of the reference: cx = (sigmoid(tx) + x) * stride;
// w = (exp(tw) * f32(anchor_w / stride)) * stride; sigmoid on objectness and classes.
#include <cstdlib>

#include "decode.cuh"

namespace rtod {

namespace {

__device__ __forceinline__ float decode_value(float v, int attr, int cell_x, int cell_y, float aw,
                                              float ah, float stride, int train) {
    if (attr >= 4) return sigmoid_f32(v);
    if (attr < 2) {
        const float s = sigmoid_f32(v);
        if (train) return s;
        return __fmul_rn(__fadd_rn(s, (float)(attr == 0 ? cell_x : cell_y)), stride);
    }
    if (train) return v;
    return __fmul_rn(__fmul_rn(expf(v), attr == 2 ? aw : ah), stride);
}

constexpr int kTileCells = 32;
constexpr int kTileCh = 256;

__global__ void __launch_bounds__(256)
yolo_decode_nchw_kernel(const float* __restrict__ head, int G, int A, int L, float stride, int train, int exact,
                        DecodeAnchors anchors, float* __restrict__ out) {
    __shared__ float tile[kTileCh][kTileCells + 1];
    const int GG = G * G, Ch = A * L;
    const int cell0 = blockIdx.x * kTileCells, ch0 = blockIdx.z * kTileCh, b = blockIdx.y;
    const int n_ch = min(kTileCh, Ch - ch0), n_cell = min(kTileCells, GG - cell0);
    const float* src = head + ((long long)b * Ch + ch0) * GG + cell0;
    for (int i = threadIdx.x; i < n_ch * kTileCells; i += 256) {
        const int ch = i / kTileCells, cl = i % kTileCells;
        tile[ch][cl] = cl < n_cell ? src[(long long)ch * GG + cl] : 0.0f;
    }
    __syncthreads();
    // one warp per cell: its channels are contiguous in the output.  Main pass: every element that is not a box
    // attribute is a plain sigmoid (one uniform instruction stream); fix-up pass: the 4 box attributes of each
    // anchor, one lane each, in the reference's op order (same scheme as yolo_decode_heads_fast_kernel)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned plain = 0;
#pragma unroll
    for (int i = 0; i < kTileCh / 32; ++i) {
        const int chl = lane + 32 * i;
        if (chl < n_ch && (ch0 + chl) % L >= 4) plain |= 1u << i;
    }
    for (int cl = warp; cl < n_cell; cl += 8) {
        const int cell = cell0 + cl;
        const int cy = cell / G, cx = cell - cy * G;
        float* dst = out + ((long long)b * GG + cell) * Ch + ch0;
#pragma unroll
        for (int i = 0; i < kTileCh / 32; ++i) {
            const float r = exact ? sigmoid_f32(tile[lane + 32 * i][cl]) : sigmoid_fast_f32(tile[lane + 32 * i][cl]);
            if ((plain >> i) & 1u) __stcs(dst + lane + 32 * i, r);
        }
        if (lane < 4 * A) {
            const int a = lane >> 2, attr = lane & 3, chl = a * L + attr - ch0;
            if (chl >= 0 && chl < n_ch) {
                const float v = tile[chl][cl];
                float r;
                if (attr < 2) {
                    r = sigmoid_f32(v);
                    if (!train) r = __fmul_rn(__fadd_rn(r, (float)(attr == 0 ? cx : cy)), stride);
                } else {
                    r = train ? v : __fmul_rn(__fmul_rn(expf(v), attr == 2 ? anchors.w[a] : anchors.h[a]), stride);
                }
                __stcs(dst + chl, r);
            }
        }
    }
}

// One warp per (image, head, cell): the A*(5+C) logits of a cell are contiguous in the NHWC head
// buffer and its A prediction rows are contiguous in the output, so both sides are coalesced and
// the index arithmetic is done once per warp instead of once per element.
__global__ void __launch_bounds__(256)
yolo_decode_heads_kernel(DecodeHeads heads, int B, int N, int L, int train, int cells_per_image,
                         float* __restrict__ pred) {
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long total = (long long)B * cells_per_image;
    for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < total; w += warps) {
        const int b = (int)(w / cells_per_image);
        int cell = (int)(w - (long long)b * cells_per_image);
        int h = 0;
#pragma unroll
        for (int k = 0; k < kMaxHeads - 1; ++k) {
            const int gg = heads.grid[h] * heads.grid[h];
            if (h + 1 < heads.count && cell >= gg) {
                cell -= gg;
                ++h;
            }
        }
        const int G = heads.grid[h], A = heads.num_anchors[h];
        const int cx = cell % G, cy = cell / G;
        const float stride = heads.stride[h];
        const float* src = heads.raw[h] + ((long long)b * G * G + cell) * heads.pitch[h];
        float* dst = pred + ((long long)b * N + heads.row_base[h] + (long long)cell * A) * L;
        const int n = A * L;
        int a = lane / L, attr = lane - a * L;               // (anchor, attribute) advance by additions
        const int step_a = 32 / L, step_attr = 32 - step_a * L;
        for (int e0 = 0; e0 < n; e0 += 256) {                // 8 independent loads in flight per lane
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int e = e0 + lane + 32 * i;
                v[i] = e < n ? __ldcs(src + e) : 0.0f;       // read once: streaming load
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int e = e0 + lane + 32 * i;
                // one exponential for every lane (no divergence): exp(v) for w/h, exp(-v) inside the sigmoid
                const bool is_wh = attr == 2 || attr == 3;
                const float t = expf(is_wh ? v[i] : -v[i]);
                float r;
                if (is_wh) {
                    r = train ? v[i] : __fmul_rn(__fmul_rn(t, attr == 2 ? heads.anchor_w[h][a] : heads.anchor_h[h][a]), stride);
                } else {
                    r = __frcp_rn(__fadd_rn(1.0f, t));       // == sigmoid_f32(v)
                    if (attr < 2 && !train) r = __fmul_rn(__fadd_rn(r, (float)(attr == 0 ? cx : cy)), stride);
                }
                if (e < n) __stcs(dst + e, r);
                a += step_a;
                attr += step_attr;
                if (attr >= L) {
                    attr -= L;
                    ++a;
                }
            }
        }
    }
}

// Fast path of the same kernel for A*L <= 256 (YOLOv3: 255): which of a lane's eight elements are box
// attributes (attr < 4) depends on the lane only, so the plain-sigmoid elements (249 of 255) run a short
// uniform path -- exp(-v), 1 + t, correctly rounded reciprocal -- and the twelve special ones take the general
// formula in the three iterations where they occur.  32-bit index arithmetic (B * cells < 2^31).
template <bool kTrain, bool kExact, bool kYolo>
__global__ void __launch_bounds__(256)
yolo_decode_heads_fast_kernel(DecodeHeads heads, unsigned total, int N, int L, unsigned cells_per_image,
                              float* __restrict__ pred) {
    const int lane = threadIdx.x & 31;
    const unsigned warps = gridDim.x * (blockDim.x >> 5);
    const int A = heads.num_anchors[0], n = A * L;
    unsigned special = 0, valid = 0;                     // bit i: element lane + 32 i is a box attribute / exists
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int e = lane + 32 * i;
        if (e < n) {
            valid |= 1u << i;
            if (e % L < 4) special |= 1u << i;
        }
    }
    const unsigned plain = valid & ~special;
    // every warp owns a contiguous run of cells: image, head and cell coordinates are located once (two
    // integer divides) and then advanced by additions; a new head or image re-locates
    const unsigned wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const unsigned per_warp = (total + warps - 1) / warps;
    unsigned w = wid * per_warp;
    const unsigned w_end = min(total, w + per_warp);
    const float* src = nullptr;
    float* dst = nullptr;
    float stride = 0.f;
    unsigned cx = 0, cy = 0, G = 1, left_in_head = 0;     // cells left in this (image, head) including the current
    int h = 0, pitch = 0;
    for (; w < w_end; ++w) {
        if (left_in_head == 0) {
            const unsigned b = w / cells_per_image;
            unsigned cell = w - b * cells_per_image;
            h = 0;
#pragma unroll
            for (int k = 0; k < kMaxHeads - 1; ++k) {
                const unsigned gg = (unsigned)(heads.grid[h] * heads.grid[h]);
                if (h + 1 < heads.count && cell >= gg) {
                    cell -= gg;
                    ++h;
                }
            }
            G = (unsigned)heads.grid[h];
            cy = cell / G;
            cx = cell - cy * G;
            stride = heads.stride[h];
            pitch = heads.pitch[h];
            left_in_head = G * G - cell;
            src = heads.raw[h] + ((size_t)b * G * G + cell) * pitch + lane;
            dst = pred + ((size_t)b * N + heads.row_base[h] + (size_t)cell * A) * L + lane;
        }
        // main pass: every element that is not a box attribute is a plain sigmoid -- one uniform instruction
        // stream for the whole warp (`plain` = valid & ~special decides the store only)
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            // kYolo (3 anchors x 85 attributes): groups 0-6 are complete, only lane 31 of group 7 is missing
            if (kYolo && i < 7) v[i] = __ldcs(src + 32 * i);                        // read once: streaming
            else v[i] = (valid >> i) & 1u ? __ldcs(src + 32 * i) : 0.0f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float r = kExact ? sigmoid_f32(v[i]) : sigmoid_fast_f32(v[i]);
            // kYolo: box attributes sit in groups 0 (0-3), 2 (85-88) and 5 (170-173) only
            if (kYolo && (i == 1 || i == 3 || i == 4 || i == 6)) __stcs(dst + 32 * i, r);
            else if ((plain >> i) & 1u) __stcs(dst + 32 * i, r);
        }
        // fix-up pass: the 4 box attributes of each anchor (A * 4 elements per cell), one lane each
        if (lane < 4 * A) {
            const int a = lane >> 2, attr = lane & 3, e = a * L + attr;
            const float x = __ldg(src - lane + e);           // (src / dst carry the + lane offset)
            float r;
            if (attr < 2) {
                r = sigmoid_f32(x);
                if (!kTrain) r = __fmul_rn(__fadd_rn(r, (float)(attr == 0 ? cx : cy)), stride);
            } else {
                r = kTrain ? x : __fmul_rn(__fmul_rn(expf(x), attr == 2 ? heads.anchor_w[h][a] : heads.anchor_h[h][a]), stride);
            }
            __stcs(dst - lane + e, r);
        }
        src += pitch;
        dst += n;
        --left_in_head;
        if (++cx == G) {
            cx = 0;
            ++cy;
        }
    }
}

}  // namespace

int launch_decode_heads(const DecodeHeads& heads, int B, int N, int L, int train, float* pred,
                        cudaStream_t stream) {
    int cells = 0;
    for (int h = 0; h < heads.count; ++h) cells += heads.grid[h] * heads.grid[h];
    const long long total = (long long)B * cells;                  // one warp each
    if (total == 0 || N == 0) return RTOD_OK;
    long long blocks = (total + 7) / 8;
    if (blocks > (long long)kNumSMs * 16) blocks = (long long)kNumSMs * 16;
    bool fast = total < (1ll << 31) && heads.num_anchors[0] * L <= 256 && heads.num_anchors[0] <= 8 && L >= 4;
    for (int h = 1; h < heads.count; ++h) fast = fast && heads.num_anchors[h] == heads.num_anchors[0];
    static const bool exact = getenv("RTOD_DECODE_EXACT") != nullptr;    // full-precision expf for every element
    const bool yolo = fast && heads.num_anchors[0] == 3 && L == 85;
    const unsigned nb = (unsigned)blocks, tot = (unsigned)total, cpi = (unsigned)cells;
    if (fast && train && exact)
        yolo_decode_heads_fast_kernel<true, true, false><<<nb, 256, 0, stream>>>(heads, tot, N, L, cpi, pred);
    else if (fast && train)                              // same sigmoid as the inference decode: TRAIN only changes the box attributes
        yolo_decode_heads_fast_kernel<true, false, false><<<nb, 256, 0, stream>>>(heads, tot, N, L, cpi, pred);
    else if (fast && exact)
        yolo_decode_heads_fast_kernel<false, true, false><<<nb, 256, 0, stream>>>(heads, tot, N, L, cpi, pred);
    else if (yolo)
        yolo_decode_heads_fast_kernel<false, false, true><<<nb, 256, 0, stream>>>(heads, tot, N, L, cpi, pred);
    else if (fast)
        yolo_decode_heads_fast_kernel<false, false, false><<<nb, 256, 0, stream>>>(heads, tot, N, L, cpi, pred);
    else
        yolo_decode_heads_kernel<<<(unsigned)blocks, 256, 0, stream>>>(heads, B, N, L, train, cells, pred);
    RTOD_LAUNCH_OK("yolo_decode_heads_kernel");
    return RTOD_OK;
}

}  // namespace rtod

using namespace rtod;

extern "C" int rtod_yolo_decode(const float* head_nchw, int B, int G, int A, int C, int inp_dim,
                                const float* anchors_host, int train, float* out, void* stream_) {
    if (B < 0 || G <= 0 || A <= 0 || A > RTOD_MAX_ANCHORS || C < 0 || inp_dim <= 0)
        return fail(RTOD_ERR_BAD_ARG, "rtod_yolo_decode: bad shape B=%d G=%d A=%d C=%d inp_dim=%d", B, G,
                    A, C, inp_dim);
    if (B == 0) return RTOD_OK;
    if (!head_nchw || !out || !anchors_host)
        return fail(RTOD_ERR_BAD_ARG, "rtod_yolo_decode: null pointer");
    const int stride = inp_dim / G;                       // src/util.py:194
    if (stride <= 0 || inp_dim / stride != G)             // src/util.py:195 grid_size must match
        return fail(RTOD_ERR_BAD_ARG, "rtod_yolo_decode: inp_dim %d does not fit grid %d", inp_dim, G);
    DecodeAnchors anc;
    for (int a = 0; a < RTOD_MAX_ANCHORS; ++a) {
        // python evaluates anchor/stride in double and FloatTensor() rounds it to fp32
        anc.w[a] = a < A ? (float)((double)anchors_host[2 * a] / (double)stride) : 0.0f;
        anc.h[a] = a < A ? (float)((double)anchors_host[2 * a + 1] / (double)stride) : 0.0f;
    }
    const int L = 5 + C, Ch = A * L, GG = G * G;
    dim3 grid(ceil_div(GG, kTileCells), B, ceil_div(Ch, kTileCh));
    if (A > 8 || L < 4) return fail(RTOD_ERR_UNSUPPORTED, "rtod_yolo_decode: at most 8 anchors per head");
    static const int exact = getenv("RTOD_DECODE_EXACT") != nullptr;      // full-precision expf for every element
    yolo_decode_nchw_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>(head_nchw, G, A, L, (float)stride,
                                                                    train, exact, anc, out);
    RTOD_LAUNCH_OK("yolo_decode_nchw_kernel");
    return RTOD_OK;
}
