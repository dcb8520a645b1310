// decode.cu -- util.predict_transform (src/util.py:175-239) on the GPU.
//
// Two entry points share one element transform:
//   * yolo_decode_nchw_tma_kernel / yolo_decode_nchw_kernel: the drop-in for predict_transform(): an
//     NCHW fp32 head [B, A*(5+C), G, G] is transposed through shared memory (128-byte reads along the
//     cell axis, coalesced writes along the attribute axis) into [B, G*G*A, 5+C]; persistent CTAs fed
//     by a TMA tile ring when the channel rows are 16-byte aligned (G*G % 4 == 0), one tile per CTA
//     otherwise;
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
#include "tc_ptx.cuh"

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

// tile: kTileCh channels x kTileCells cells (32 KB of shared memory)
template <int kTileCh, int kTileCells>
__global__ void __launch_bounds__(256)
yolo_decode_nchw_kernel(const float* __restrict__ head, int G, int A, int L, float stride, int train, int exact, int vec,
                        DecodeAnchors anchors, float* __restrict__ out) {
    extern __shared__ float tile_raw[];
    float (*tile)[kTileCells + 1] = reinterpret_cast<float (*)[kTileCells + 1]>(tile_raw);
    const int GG = G * G, Ch = A * L;
    const int cell0 = blockIdx.x * kTileCells, ch0 = blockIdx.z * kTileCh, b = blockIdx.y;
    const int n_ch = min(kTileCh, Ch - ch0), n_cell = min(kTileCells, GG - cell0);
    const float* src = head + ((long long)b * Ch + ch0) * GG + cell0;
    if (vec && n_cell == kTileCells) {
        // whole tile, 16-byte aligned rows (G*G % 4 == 0): every thread has its eight 16-byte loads in flight at once
        // (32 KB per CTA; the scalar loop below keeps a few hundred bytes per warp in flight and is latency-bound);
        // a warp covers four channels x 128 B, its shared-memory writes (row pitch 33 words) are conflict-free
        constexpr int kPerRow = kTileCells / 4, kLoads = kTileCh * kPerRow / 256;
        float4 v[kLoads];
#pragma unroll
        for (int k = 0; k < kLoads; ++k) {
            const int j = threadIdx.x + 256 * k, ch = j / kPerRow;
            v[k] = ch < n_ch ? __ldcs(reinterpret_cast<const float4*>(src + (long long)ch * GG) + (j % kPerRow)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < kLoads; ++k) {
            const int j = threadIdx.x + 256 * k, ch = j / kPerRow, cl = (j % kPerRow) * 4;
            tile[ch][cl] = v[k].x; tile[ch][cl + 1] = v[k].y; tile[ch][cl + 2] = v[k].z; tile[ch][cl + 3] = v[k].w;
        }
    } else
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
    {   // attribute of channel ch0 + lane + 32 i: one modulo, then carried additions (an integer modulo by a run-time
        // value costs ~30 instructions; eight of them per thread were a fifth of this kernel's instruction stream)
        int attr = (int)((unsigned)(ch0 + lane) % (unsigned)L);
#pragma unroll
        for (int i = 0; i < kTileCh / 32; ++i) {
            if (lane + 32 * i < n_ch && attr >= 4) plain |= 1u << i;
            attr += 32;
            while (attr >= L) attr -= L;
        }
    }
    constexpr int kWarps = 8;
    for (int cl = warp; cl < n_cell; cl += kWarps) {
        float* dst = out + ((long long)b * GG + cell0 + cl) * Ch + ch0;
#pragma unroll
        for (int i = 0; i < kTileCh / 32; ++i) {
            const float r = exact ? sigmoid_f32(tile[lane + 32 * i][cl]) : sigmoid_fast_f32(tile[lane + 32 * i][cl]);
            if ((plain >> i) & 1u) __stcs(dst + lane + 32 * i, r);
        }
    }
    // fix-up: the 4 box attributes of each anchor in the reference's op order, one THREAD per (cell, anchor, attribute)
    // over the whole tile -- the full-precision expf / division chain is ~80 dependent instructions; run per cell by
    // twelve lanes of a warp it took longer than the main pass
    const int per_cell = 4 * A;
    for (int it = threadIdx.x; it < n_cell * per_cell; it += 256) {
        const int cl = it / per_cell, rem = it - cl * per_cell, a = rem >> 2, attr = rem & 3, chl = a * L + attr - ch0;
        if (chl < 0 || chl >= n_ch) continue;
        const int cell = cell0 + cl, cy = cell / G, cx = cell - cy * G;
        const float v = tile[chl][cl];
        float r;
        if (attr < 2) {
            r = sigmoid_f32(v);
            if (!train) r = __fmul_rn(__fadd_rn(r, (float)(attr == 0 ? cx : cy)), stride);
        } else {
            float anchor = 0.0f;
#pragma unroll
            for (int q = 0; q < RTOD_MAX_ANCHORS; ++q)           // (a run-time index into the by-value arrays would go through local memory)
                if (q == a) anchor = attr == 2 ? anchors.w[q] : anchors.h[q];
            r = train ? v : __fmul_rn(__fmul_rn(expf(v), anchor), stride);
        }
        __stcs(out + ((long long)b * GG + cell) * Ch + ch0 + chl, r);
    }
}

// TMA-pipelined variant of yolo_decode_nchw_kernel (G*G % 4 == 0, A*(5+C) <= 256): persistent CTAs, one producer thread
// keeps `stages` 32 KB tiles (256 channel rows x 32 cells, 128-byte rows, SWIZZLE_128B) in flight with 3-D tiled TMA loads
// of the NCHW head {cells, channels, images}; eight consumer warps decode tile n while tiles n+1 .. n+stages-1 are landing.
// The load -> barrier -> store phases of the one-tile-per-CTA kernel never overlap inside a CTA; here the stores of one
// tile always run under the loads of the next ones.  Channel row 255 and cells beyond G*G are zero-filled by the TMA
// bounds check.  Reading column `cl` of 32 different rows from the swizzled tile is a 4-way bank conflict (8 distinct
// 16-byte slots per 128-byte row): 256 such reads per tile, far from the shared-memory limit.
constexpr int kTmaTileCells = 32, kTmaTileCh = 256, kTmaConsumers = 256;
constexpr uint32_t kTmaTileBytes = kTmaTileCells * kTmaTileCh * 4;

__device__ __forceinline__ bool decode_wait(uint64_t* bar, uint32_t parity) {     // bounded: never hangs the GPU
    for (unsigned spins = 0; spins < (1u << 26); ++spins)
        if (mbar_try_wait(bar, parity)) return true;
    return false;
}
__device__ __forceinline__ float swz_tile(const uint8_t* tile, int ch, int cl) {  // element (ch, cl) of a SWIZZLE_128B tile
    return *reinterpret_cast<const float*>(tile + ch * 128 + ((((cl >> 2) ^ (ch & 7))) << 4) + ((cl & 3) << 2));
}

template <bool kExact>
__global__ void __launch_bounds__(kTmaConsumers + 32)
yolo_decode_nchw_tma_kernel(const __grid_constant__ CUtensorMap tm, int G, int A, int L, float stride, int train, DecodeAnchors anchors,
                            int stages, unsigned tiles_per_image, unsigned total_tiles, float* __restrict__ out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * kTmaTileBytes);
    uint64_t* empty = full + stages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int GG = G * G, Ch = A * L;
    if (threadIdx.x == 0) {
        prefetch_tmap(&tm);
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kTmaConsumers / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == kTmaConsumers / 32) {
        if (lane == 0) {                                     // producer
            int stage = 0;
            uint32_t phase = 0;
            for (unsigned t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const unsigned b = t / tiles_per_image;
                const int cell0 = (int)(t - b * tiles_per_image) * kTmaTileCells;
                if (!decode_wait(&empty[stage], phase ^ 1u)) __trap();       // no plan, no failure flag: fail loudly
                mbar_expect_tx(&full[stage], kTmaTileBytes);
                tma_load_3d(smem + (size_t)stage * kTmaTileBytes, &tm, &full[stage], cell0, 0, (int)b);
                if (++stage == stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
        return;
    }
    unsigned plain = 0;                                      // bit i: channel lane + 32 i is a plain sigmoid
    {
        int attr = (int)((unsigned)lane % (unsigned)L);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (lane + 32 * i < Ch && attr >= 4) plain |= 1u << i;
            attr += 32;
            while (attr >= L) attr -= L;
        }
    }
    const int per_cell = 4 * A;
    int stage = 0;
    uint32_t phase = 0;
    for (unsigned t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const unsigned b = t / tiles_per_image;
        const int cell0 = (int)(t - b * tiles_per_image) * kTmaTileCells;
        const int n_cell = min(kTmaTileCells, GG - cell0);
        const uint8_t* tile = smem + (size_t)stage * kTmaTileBytes;
        if (!decode_wait(&full[stage], phase)) __trap();
        // main pass: one warp per cell, plain sigmoids in one uniform instruction stream
        for (int cl = warp; cl < n_cell; cl += kTmaConsumers / 32) {
            float* dst = out + ((size_t)b * GG + cell0 + cl) * Ch + lane;
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = swz_tile(tile, lane + 32 * i, cl);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float r = kExact ? sigmoid_f32(v[i]) : sigmoid_fast_f32(v[i]);
                if ((plain >> i) & 1u) __stcs(dst + 32 * i, r);
            }
        }
        // fix-up: box attributes in the reference's op order, one thread per (cell, anchor, attribute) of the tile
        for (int it = threadIdx.x; it < n_cell * per_cell; it += kTmaConsumers) {
            const int cl = it / per_cell, rem = it - cl * per_cell, a = rem >> 2, attr = rem & 3, ch = a * L + attr;
            const int cell = cell0 + cl, cy = cell / G, cx = cell - cy * G;
            const float x = swz_tile(tile, ch, cl);
            float r;
            if (attr < 2) {
                r = sigmoid_f32(x);
                if (!train) r = __fmul_rn(__fadd_rn(r, (float)(attr == 0 ? cx : cy)), stride);
            } else {
                float anchor = 0.0f;
#pragma unroll
                for (int q = 0; q < RTOD_MAX_ANCHORS; ++q)
                    if (q == a) anchor = attr == 2 ? anchors.w[q] : anchors.h[q];
                r = train ? x : __fmul_rn(__fmul_rn(expf(x), anchor), stride);
            }
            __stcs(out + ((size_t)b * GG + cell) * Ch + ch, r);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);           // this warp has read the tile
        if (++stage == stages) {
            stage = 0;
            phase ^= 1u;
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

// Ring variant of the fast kernel for the YOLOv3 layout (3 anchors x 85 attributes, inference, SFU sigmoids): persistent
// CTAs, one producer thread keeps `stages` tiles of up to 32 consecutive cells of one (image, head) -- 32 KB of contiguous
// NHWC logits -- in flight with 1-D bulk copies (cp.async.bulk, mbarrier transaction count), eight consumer warps decode a
// tile (one warp per cell, conflict-free row reads, 128-byte streaming stores; box attributes by a tile-wide fix-up pass)
// while the next ones land: the same pipeline that took the NCHW kernel from 0.54 to 0.79 of the HBM peak.
constexpr int kRingCells = 32, kRingConsumers = 256;

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

struct RingTile {                      // tile t of the step -> (image, head, first cell, cells)
    unsigned b;
    int h, cell0, n_cell;
};
__device__ __forceinline__ RingTile ring_tile(const DecodeHeads& heads, unsigned t, unsigned tiles_per_image) {
    RingTile r;
    r.b = t / tiles_per_image;
    int rem = (int)(t - r.b * tiles_per_image);
    r.h = 0;
#pragma unroll
    for (int k = 0; k < kMaxHeads - 1; ++k) {
        const int th = (heads.grid[r.h] * heads.grid[r.h] + kRingCells - 1) / kRingCells;
        if (r.h + 1 < heads.count && rem >= th) {
            rem -= th;
            ++r.h;
        }
    }
    r.cell0 = rem * kRingCells;
    r.n_cell = min(kRingCells, heads.grid[r.h] * heads.grid[r.h] - r.cell0);
    return r;
}

__global__ void __launch_bounds__(kRingConsumers + 32)
yolo_decode_heads_ring_kernel(DecodeHeads heads, int N, int stages, unsigned tiles_per_image, unsigned total_tiles,
                              float* __restrict__ pred) {
    constexpr int L = 85, A = 3, n = A * L;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 127u) & ~127u) - raw_addr);
    const uint32_t stage_bytes = (uint32_t)kRingCells * 256u * 4u;                  // pitch <= 256 floats (checked by the host)
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
    uint64_t* empty = full + stages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kRingConsumers / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == kRingConsumers / 32) {
        if (lane == 0) {                                     // producer
            int stage = 0;
            uint32_t phase = 0;
            for (unsigned t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const RingTile rt = ring_tile(heads, t, tiles_per_image);
                const int G = heads.grid[rt.h], pitch = heads.pitch[rt.h];
                if (!decode_wait(&empty[stage], phase ^ 1u)) {
                    if (heads.err_flag) raise_device_error(heads.err_flag, 3);
                    break;
                }
                const uint32_t bytes = (uint32_t)rt.n_cell * (uint32_t)pitch * 4u;
                mbar_expect_tx(&full[stage], bytes);
                bulk_load_1d(smem + (size_t)stage * stage_bytes, heads.raw[rt.h] + ((size_t)rt.b * G * G + rt.cell0) * pitch, bytes,
                             &full[stage]);
                if (++stage == stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
        return;
    }
    int stage = 0;
    uint32_t phase = 0;
    for (unsigned t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const RingTile rt = ring_tile(heads, t, tiles_per_image);
        const int G = heads.grid[rt.h], pitch = heads.pitch[rt.h];
        const float stride = heads.stride[rt.h];
        const float* tile = reinterpret_cast<const float*>(smem + (size_t)stage * stage_bytes);
        float* out0 = pred + ((size_t)rt.b * N + heads.row_base[rt.h] + (size_t)rt.cell0 * A) * L;
        if (!decode_wait(&full[stage], phase)) {             // (never in practice: reported like the convolutions' time-outs)
            if (heads.err_flag && lane == 0) raise_device_error(heads.err_flag, 3);
            break;
        }
        // main pass: groups 0-6 are complete, lane 31 of group 7 (element 255) does not exist; box attributes sit in
        // groups 0 (0-3), 2 (85-88) and 5 (170-173) and are left to the fix-up pass
        for (int cl = warp; cl < rt.n_cell; cl += kRingConsumers / 32) {
            const float* row = tile + (size_t)cl * pitch + lane;
            float* dst = out0 + (size_t)cl * n + lane;
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = row[32 * i];                          // (element 255 is padding inside the pitch)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float r = sigmoid_fast_f32(v[i]);
                const int e = lane + 32 * i;
                bool st = true;
                if (i == 0) st = lane >= 4;
                if (i == 2) st = e < L || e >= L + 4;
                if (i == 5) st = e < 2 * L || e >= 2 * L + 4;
                if (i == 7) st = lane < 31;
                if (st) __stcs(dst + 32 * i, r);
            }
        }
        for (int it = threadIdx.x; it < rt.n_cell * 4 * A; it += kRingConsumers) {
            const int cl = it / (4 * A), rem = it - cl * 4 * A, a = rem >> 2, attr = rem & 3, e = a * L + attr;
            const int cell = rt.cell0 + cl, cy = cell / G, cx = cell - cy * G;
            const float x = tile[(size_t)cl * pitch + e];
            float r;
            if (attr < 2) r = __fmul_rn(__fadd_rn(sigmoid_f32(x), (float)(attr == 0 ? cx : cy)), stride);
            else r = __fmul_rn(__fmul_rn(expf(x), attr == 2 ? heads.anchor_w[rt.h][a] : heads.anchor_h[rt.h][a]), stride);
            __stcs(out0 + (size_t)cl * n + e, r);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == stages) {
            stage = 0;
            phase ^= 1u;
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
    bool ring_ok = yolo && !train && !exact;             // bulk copies: 16-byte aligned rows of at most 256 floats
    for (int h = 0; h < heads.count; ++h)
        ring_ok = ring_ok && heads.pitch[h] <= 256 && heads.pitch[h] % 4 == 0 && (reinterpret_cast<uintptr_t>(heads.raw[h]) & 15u) == 0;
    const unsigned nb = (unsigned)blocks, tot = (unsigned)total, cpi = (unsigned)cells;
    if (fast && train && exact)
        yolo_decode_heads_fast_kernel<true, true, false><<<nb, 256, 0, stream>>>(heads, tot, N, L, cpi, pred);
    else if (fast && train)                              // same sigmoid as the inference decode: TRAIN only changes the box attributes
        yolo_decode_heads_fast_kernel<true, false, false><<<nb, 256, 0, stream>>>(heads, tot, N, L, cpi, pred);
    else if (fast && exact)
        yolo_decode_heads_fast_kernel<false, true, false><<<nb, 256, 0, stream>>>(heads, tot, N, L, cpi, pred);
    else if (yolo && ring_ok && !getenv("RTOD_DECODE_NO_RING")) {
        const int stages = getenv("RTOD_DECODE_STAGES") && atoi(getenv("RTOD_DECODE_STAGES")) >= 2 && atoi(getenv("RTOD_DECODE_STAGES")) <= 6 ? atoi(getenv("RTOD_DECODE_STAGES")) : 3;
        const int smem = stages * kRingCells * 256 * 4 + 128 + 256;
        unsigned per_image = 0;
        for (int h = 0; h < heads.count; ++h) per_image += (unsigned)((heads.grid[h] * heads.grid[h] + kRingCells - 1) / kRingCells);
        const unsigned tiles = per_image * (unsigned)B;
        const int per_sm = (227 * 1024) / (smem + 1024);
        unsigned grid = (unsigned)(kNumSMs * per_sm);
        if (grid > tiles) grid = tiles;
        {   // once per device (a plan's first forward is never inside a stream capture)
            static unsigned long long attr_set = 0;
            int dev = 0;
            cudaGetDevice(&dev);
            if (dev >= 64 || !((attr_set >> dev) & 1ull)) {
                RTOD_CUDA_OK(cudaFuncSetAttribute(yolo_decode_heads_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * kRingCells * 256 * 4 + 128 + 256));
                if (dev < 64) attr_set |= 1ull << dev;
            }
        }
        yolo_decode_heads_ring_kernel<<<grid, kRingConsumers + 32, smem, stream>>>(heads, N, stages, per_image, tiles, pred);
    } else if (yolo)
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
    if (A > 8 || L < 4) return fail(RTOD_ERR_UNSUPPORTED, "rtod_yolo_decode: at most 8 anchors per head");
    static const int exact = getenv("RTOD_DECODE_EXACT") != nullptr;      // full-precision expf for every element
    const int vec = GG % 4 == 0 && (reinterpret_cast<uintptr_t>(head_nchw) & 15u) == 0;      // 16-byte aligned channel rows
    const bool no_tma = getenv("RTOD_DECODE_NO_TMA") != nullptr;       // tests: the fallback kernel on every shape
    if (vec && Ch <= kTmaTileCh && !no_tma && (long long)B * GG < (1ll << 31)) {   // TMA-pipelined tiles (see the kernel)
        static EncodeTiledFn encode_tiled = nullptr;
        if (!encode_tiled) {
            const int rc = driver_fn("cuTensorMapEncodeTiled", reinterpret_cast<void**>(&encode_tiled));
            if (rc) return rc;
        }
        CUtensorMap tm;
        const cuuint64_t dims[3] = {(cuuint64_t)GG, (cuuint64_t)Ch, (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)GG * 4, (cuuint64_t)GG * 4 * Ch};
        const cuuint32_t box[3] = {(cuuint32_t)kTmaTileCells, (cuuint32_t)kTmaTileCh, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = encode_tiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(head_nchw), dims, strides, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(RTOD_ERR_CUDA, "cuTensorMapEncodeTiled (decode head) failed: %d", (int)r);
        const int stages_env = getenv("RTOD_DECODE_STAGES") ? atoi(getenv("RTOD_DECODE_STAGES")) : 0;
        const int stages = stages_env >= 2 && stages_env <= 6 ? stages_env : 2;
        const int smem = stages * (int)kTmaTileBytes + 1024 + 256;
        const int per_sm = (227 * 1024) / (smem + 1024);
        const unsigned per_image = (unsigned)ceil_div(GG, kTmaTileCells), total = (unsigned)B * per_image;
        unsigned blocks = (unsigned)(kNumSMs * (per_sm < 1 ? 1 : per_sm));
        if (blocks > total) blocks = total;
        if (exact) {
            RTOD_CUDA_OK(cudaFuncSetAttribute(yolo_decode_nchw_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            yolo_decode_nchw_tma_kernel<true><<<blocks, kTmaConsumers + 32, smem, (cudaStream_t)stream_>>>(tm, G, A, L, (float)stride, train, anc, stages, per_image, total, out);
        } else {
            RTOD_CUDA_OK(cudaFuncSetAttribute(yolo_decode_nchw_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            yolo_decode_nchw_tma_kernel<false><<<blocks, kTmaConsumers + 32, smem, (cudaStream_t)stream_>>>(tm, G, A, L, (float)stride, train, anc, stages, per_image, total, out);
        }
        RTOD_LAUNCH_OK("yolo_decode_nchw_tma_kernel");
        return RTOD_OK;
    }
    {   // fallback (odd G*G, > 256 channels): one 256-channel x 32-cell tile per CTA through padded shared memory
        const int smem = 256 * 33 * 4;
        yolo_decode_nchw_kernel<256, 32><<<dim3(ceil_div(GG, 32), B, ceil_div(Ch, 256)), 256, smem, (cudaStream_t)stream_>>>(
            head_nchw, G, A, L, (float)stride, train, exact, vec, anc, out);
    }
    RTOD_LAUNCH_OK("yolo_decode_nchw_kernel");
    return RTOD_OK;
}
