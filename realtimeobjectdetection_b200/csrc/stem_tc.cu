// stem_tc.cu -- the first convolution block (src/darknet.py:488-501 with Cin = 3: 3x3, stride 1, pad 1, BatchNorm
// folded, leaky 0.1) on the tcgen05 tensor cores, fp16 storage mode.
//
// K = 27 is one thin K block, so there is nothing for TMA to gather: the operand is BUILT in shared memory -- and only
// a third of the im2col matrix is: per 128-pixel strip of one image row
//   * a 3-D TMA tiled load stages the 3 rows x 3 channels x (128 + halo) columns neighbourhood (fp32 NCHW planes, or
//     uint8 planes of the pre-processing kernel; halo and image border zero-filled by the TMA bounds check) three
//     strips ahead;
//   * 128 "builder" threads write ONE operand row per input COLUMN x0-1 .. x0+128 (130 rows): its nine (channel, ky)
//     values plus a constant 1 (K = 16, the first half of a 64-byte-swizzled row), each value split into two fp16 terms
//     hi + lo.  The three kx taps are that slab shifted by 0 / 1 / 2 rows: a UMMA descriptor may start at any row of a
//     swizzled tile (tools/umma_shift64_probe.cu);
//   * one thread issues, per tap kx, D[128 x 2C] += A_hi(kx) * [W_hi | W_lo](kx)^T and D[:, :C] += A_lo(kx) * W_hi(kx)^T --
//     the three significant products of (a_hi + a_lo)(w_hi + w_lo), ~22 bits of both operands (uint8 frames are exact
//     in one fp16 term: one MMA per tap; they enter as value / 256, the weights carry the other 256 / 255 of
//     prep_image's / 255).  The bias rides on the constant-1 column of the centre tap;
//   * 128 "epilogue" threads read the accumulator (TMEM), add the two column halves, apply the leaky slope, round to
//     fp16 and store their pixel's C channels: a warp writes 32 pixels x 2C bytes of contiguous NHWC through TMA.
// Operand tiles and accumulators are double-buffered; builders of strip i+1 overlap the MMAs and the epilogue of
// strip i.  Three CTAs per SM.  (History: stem.cu -- mma.sync fragments built in registers, still used for bf16 storage --
// ran at 0.3 of the layer's HBM bound; the first tcgen05 version built all 27 taps per pixel: three times the builder
// work of this one.)
#include <cstdlib>

#include "layers.cuh"
#include "tc_ptx.cuh"

namespace rtod {

namespace {

constexpr int kStages = 3;            // staged input strips in flight
constexpr int kStrip = 128;           // output pixels per strip = MMA M
constexpr int kBoxW = kStrip + 8;     // staged columns: 4 left (16-byte alignment of the TMA start), 1 + 3 right
constexpr int kSlabRows = kStrip + 2; // operand rows per strip: input columns x0-1 .. x0+128
constexpr uint32_t kATileBytes = ((uint32_t)kSlabRows * 64u + 511u) & ~511u;   // 8704: keeps every tile on a swizzle-pattern boundary

struct StemTcParams {
    CUtensorMap tmX;                  // {W, H, 3*B} fp32 or uint8, box {kBoxW (uint8: kStrip + 32), 3, 3}
    CUtensorMap tmOut;                // NHWC fp16 output as {C, W, B*H}, box {C, 32, 1}: columns beyond W are clipped
    const float* w;                   // [C][27] fp32, BN folded
    const float* bias;                // [C]
    int B, H, W, C, leaky;
    int strips_per_row, total_strips;
    int no_relay;                     // bring-up switch RTOD_STEM_NO_RELAY: every builder warp polls its barriers itself
    float in_scale;                   // 1 (fp32 frames) or 256/255 (uint8 frames, fed as value / 256): folded into the weights
};

// byte offset of 16-byte chunk j of row r in a K-major tile of 64-byte rows, SWIZZLE_64B
__device__ __forceinline__ uint32_t a_row_offset(int r, int j) { return (uint32_t)r * 64u + (((uint32_t)j ^ ((uint32_t)(r >> 1) & 3u)) << 4); }
// ... of 32-byte rows, SWIZZLE_32B (the weight tiles: K = 16)
__device__ __forceinline__ uint32_t b_row_offset(int r, int j) { return (uint32_t)r * 32u + (((uint32_t)j ^ ((uint32_t)(r >> 2) & 1u)) << 4); }

// mbarrier wait that parks the warp in hardware for up to ~20 us per attempt instead of spinning through the issue
// slots the working warps need (128 threads wait at a time here)
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    for (int tries = 0; tries < (1 << 20) && !done; ++tries)            // bounded: never in practice, avoids a hard hang
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
            : "memory");
}

struct StripPos {                     // (segment of a row, row, image) of a strip, advanced by the grid size per step
    int seg, y, b;
};

template <bool kU8>
__global__ void __launch_bounds__(288, 3) stem_tc_kernel(const __grid_constant__ StemTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    constexpr uint32_t kElem = kU8 ? 1u : 4u;
    constexpr uint32_t kBox = kU8 ? (uint32_t)kStrip + 32u : (uint32_t)kBoxW;      // uint8: 16 columns left, 16 right of the strip
    constexpr uint32_t kLeft = kU8 ? 16u : 4u;
    constexpr uint32_t kStageBytes = 9u * kBox * kElem;                              // 3 channels x 3 rows
    constexpr uint32_t kStagePitch = (kStageBytes + 127u) & ~127u;
    constexpr uint32_t kATile = kATileBytes;                                         // 130 rows x 64 B, rounded up
    // layout: [2 bufs][A_hi | A_lo] | B [3 taps][2C rows x 32 B] | output staging | staged strips | barriers
    uint8_t* a_tiles = smem;
    uint8_t* b_tile = smem + 4 * kATile;
    uint8_t* out_stage = b_tile + (size_t)p.C * 192;               // [4 warps][2][32 pixels x 2C bytes], swizzled like the store map
    uint8_t* stage0 = out_stage + 8 * (size_t)p.C * 64;
    uint64_t* in_full = reinterpret_cast<uint64_t*>(stage0 + kStages * kStagePitch);
    uint64_t* in_empty = in_full + kStages;        // [kStages] the builders have read staged strip s
    uint64_t* a_ready = in_empty + kStages;        // [2] operand buffer b is built
    uint64_t* a_free = a_ready + 2;                // [2] MMAs that read operand buffer b have completed
    uint64_t* acc_full = a_free + 2;               // [2] accumulator b complete
    uint64_t* acc_empty = acc_full + 2;            // [2] epilogue has drained accumulator b
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(in_full) + 128);   // [C <= 64], 16-byte aligned

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int C = p.C;
    if (tid == 0) {
        prefetch_tmap(&p.tmX);
        prefetch_tmap(&p.tmOut);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&in_full[s], 1);
            mbar_init(&in_empty[s], 4);            // one arrive per builder warp
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&a_ready[b], 4);
            mbar_init(&a_free[b], 1);
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 4);           // one arrive per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) tmem_alloc(tmem_slot, (uint32_t)(4 * C < 32 ? 32 : 4 * C));       // 2 buffers x 2C columns
    (void)s_bias;
    // weights, one tile per tap kx: rows [0, C) = hi terms, [C, 2C) = lo terms of w * in_scale; 16 K values per row:
    // k = c*3 + ky (9 taps), k = 9 = the bias (centre tap only: it meets the operand's constant-1 column), zeros
    for (int i = tid; i < 3 * 2 * C * 2; i += 288) {
        const int j = i & 1, row = (i >> 1) % (2 * C), kx = i / (4 * C), n = row < C ? row : row - C;
        uint32_t packed[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float wv[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int k = j * 8 + q * 2 + e;
                wv[e] = k < 9 ? __ldg(p.w + n * 27 + k * 3 + kx) * p.in_scale : ((k == 9 && kx == 1) ? __ldg(p.bias + n) : 0.0f);
            }
            const uint32_t hi = pack_f16x2(wv[0], wv[1]);
            packed[q] = row < C ? hi : pack_f16x2(wv[0] - f16_lo(hi), wv[1] - f16_hi(hi));
        }
        *reinterpret_cast<uint4*>(b_tile + (size_t)kx * 2 * C * 32 + b_row_offset(row, j)) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_mine = ((int)blockIdx.x < p.total_strips) ? (p.total_strips - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    // strip -> (segment, row, image) advances by gridDim.x per step: carried additions instead of integer divides
    const int g_seg = (int)(gridDim.x % (unsigned)p.strips_per_row);
    const int g_y = (int)((gridDim.x / (unsigned)p.strips_per_row) % (unsigned)p.H);
    const int g_b = (int)(gridDim.x / (unsigned)(p.strips_per_row * p.H));
    auto advance = [&](StripPos& q) {
        q.seg += g_seg;
        const int c0 = q.seg >= p.strips_per_row;
        q.seg -= c0 * p.strips_per_row;
        q.y += g_y + c0;
        const int c1 = q.y >= p.H;
        q.y -= c1 * p.H;
        q.b += g_b + c1;
    };
    StripPos first;
    first.seg = (int)(blockIdx.x % (unsigned)p.strips_per_row);
    first.y = (int)((blockIdx.x / (unsigned)p.strips_per_row) % (unsigned)p.H);
    first.b = (int)(blockIdx.x / (unsigned)(p.strips_per_row * p.H));

    if (warp == 8) {
        // ================= control thread: stages strips (TMA) ahead and issues the MMAs =================
        if (lane == 0) {
            StripPos lpos = first;
            int lnext = 0;
            auto issue_load = [&]() {                      // stage this CTA's next strip (after its stage was read)
                if (lnext >= n_mine) return;
                const int s = lnext % kStages;
                if (lnext >= kStages) mbar_wait_parked(&in_empty[s], (uint32_t)(lnext / kStages - 1) & 1u);
                mbar_expect_tx(&in_full[s], kStageBytes);
                tma_load_3d(stage0 + s * kStagePitch, &p.tmX, &in_full[s], lpos.seg * kStrip - (int)kLeft, lpos.y - 1, 3 * lpos.b);
                advance(lpos);
                ++lnext;
            };
            for (int l = 0; l < kStages - 1; ++l) issue_load();
            const uint32_t idesc_cat = umma_idesc(1, kStrip, 2 * C), idesc_hi = umma_idesc(1, kStrip, C);
            const uint64_t desc_tmpl = smem_desc(0u, 64u), descb_tmpl = smem_desc(0u, 32u);
            const uint32_t b_addr = smem_u32(b_tile);
            for (int local = 0; local < n_mine; ++local) {
                const int buf = local & 1;
                const uint32_t use = (uint32_t)(local >> 1);
                issue_load();
                mbar_wait_parked(&a_ready[buf], use & 1u);                 // the builders have written this operand buffer
                mbar_wait_parked(&acc_empty[buf], (use & 1u) ^ 1u);        // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)(buf * 2 * C);
                const uint32_t a_hi = smem_u32(a_tiles + (size_t)buf * 2 * kATile);
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {                           // tap kx = the slab shifted by kx rows
                    const uint64_t da = desc_tmpl | (uint64_t)(((a_hi + (uint32_t)kx * 64u) & 0x3FFFFu) >> 4);
                    const uint64_t dl = desc_tmpl | (uint64_t)(((a_hi + kATile + (uint32_t)kx * 64u) & 0x3FFFFu) >> 4);
                    const uint64_t db = descb_tmpl | (uint64_t)(((b_addr + (uint32_t)(kx * 2 * C * 32)) & 0x3FFFFu) >> 4);
                    umma_bf16(acc, da, db, idesc_cat, kx ? 1u : 0u);
                    if (!kU8) umma_bf16(acc, dl, db, idesc_hi, 1u);
                }
                umma_commit(&a_free[buf]);
                umma_commit(&acc_full[buf]);
            }
        }
    } else if (warp < 4) {
        const bool no_relay = p.no_relay != 0;
        // ================= builders: one operand row (= one input column) per thread; threads 0, 1 also rows 128, 129 ====
        for (int local = 0; local < n_mine; ++local) {
            const int s = local % kStages, buf = local & 1;
            const uint32_t use = (uint32_t)(local >> 1);           // how often this buffer has been used before
            // the staged strip has landed; the MMAs that read this operand buffer two strips ago have completed
            // ONE builder warp polls the two barriers, the other three sleep in a named barrier: a polling warp re-issues
            // its try_wait loop every time the hardware wakes it (ncu: those loops were 18 % of the instructions this
            // issue-bound kernel executes)
            if (warp == 0 || no_relay) {
                mbar_wait_parked(&in_full[s], (uint32_t)(local / kStages) & 1u);
                if (local >= 2) mbar_wait_parked(&a_free[buf], (use - 1u) & 1u);
            }
            if (!no_relay) asm volatile("bar.sync 1, 128;" ::: "memory");
            const uint8_t* st = stage0 + s * kStagePitch;
            uint8_t* a_hi = a_tiles + (size_t)buf * 2 * kATile;
            uint8_t* a_lo = a_hi + kATile;
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const int t = pass == 0 ? tid : kStrip + tid;      // slab row = input column x0 - 1 + t
                if (pass == 1 && tid >= kSlabRows - kStrip) break;
                // k = c*3 + ky: the nine staged rows at this column, then the constant 1 that meets the bias, then zeros
                float v[9];
#pragma unroll
                for (int r = 0; r < 9; ++r) {
                    const uint32_t idx = (uint32_t)r * kBox + (uint32_t)t + kLeft - 1u;
                    v[r] = kU8 ? (float)st[idx] * (1.0f / 256.0f) : reinterpret_cast<const float*>(st)[idx];
                }
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float x0 = 2 * q < 9 ? v[2 * q < 9 ? 2 * q : 0] : 0.0f;
                    const float x1 = 2 * q + 1 < 9 ? v[2 * q + 1 < 9 ? 2 * q + 1 : 0] : (2 * q + 1 == 9 ? 1.0f : 0.0f);
                    hi[q] = pack_f16x2(x0, x1);
                    if (!kU8) lo[q] = pack_f16x2(x0 - f16_lo(hi[q]), x1 - f16_hi(hi[q]));
                }
                *reinterpret_cast<uint4*>(a_hi + a_row_offset(t, 0)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(a_hi + a_row_offset(t, 1)) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                if (!kU8) {
                    *reinterpret_cast<uint4*>(a_lo + a_row_offset(t, 0)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    *reinterpret_cast<uint4*>(a_lo + a_row_offset(t, 1)) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&in_empty[s]);              // this warp has read its part of the staged strip
            fence_async_smem();                                    // generic-proxy writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_ready[buf]);
        }
    } else {
        // ================= epilogue: warp w reads TMEM lanes [32*(w%4), +32) = 32 pixels of the strip =================
        // accumulator (hi + lo column halves; the bias came through K slot 27) -> leaky -> fp16 -> the warp's swizzled
        // staging slice -> one TMA store of 32 pixels x C channels (columns beyond the image width are clipped)
        const int ew = warp & 3;
        const uint32_t row_bytes = (uint32_t)C * 2u;                  // 32 / 64 / 128 bytes per pixel
        uint8_t* my_stage = out_stage + (size_t)ew * 2 * 32 * row_bytes;
        StripPos pos = first;
        for (int local = 0; local < n_mine; ++local, advance(pos)) {
            const int buf = local & 1;
            mbar_wait_parked(&acc_full[buf], (uint32_t)(local >> 1) & 1u);
            tc_fence_after();
            const uint32_t acc = tmem_base + (uint32_t)(buf * 2 * C) + ((uint32_t)(ew * 32) << 16);
            uint8_t* slice = my_stage + (size_t)buf * 32 * row_bytes;
            if (lane == 0) bulk_wait_read_1();                      // the store that last read this slice has drained
            __syncwarp();
            // 16 channels at a time: the hi and the lo accumulator columns are read with ONE wait (this warp's chain of
            // TMEM round trips is what bounds the kernel), added, activated, rounded and staged
            for (int c0 = 0; c0 < C; c0 += 16) {
                uint32_t vh[16], vl[16];
                tmem_ld_32x16_issue(acc + (uint32_t)c0, vh);
                tmem_ld_32x16_issue(acc + (uint32_t)(C + c0), vl);
                tmem_ld_wait();
                if (c0 + 16 >= C) {                                // last TMEM read of this strip: hand the buffer back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[buf]);
                }
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    float f[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        f[j] = __uint_as_float(vh[q * 8 + j]) + __uint_as_float(vl[q * 8 + j]);
                        if (p.leaky) f[j] = fmaxf(f[j], 0.1f * f[j]);
                    }
                    // 16-byte chunk (c0 / 8 + q) of row `lane`, swizzled like the store's tensor map
                    const uint32_t chunk = (uint32_t)(c0 >> 3) + (uint32_t)q;
                    const uint32_t sw = row_bytes == 128 ? (uint32_t)(lane & 7) : (row_bytes == 64 ? (uint32_t)((lane >> 1) & 3) : (uint32_t)((lane >> 2) & 1));
                    *reinterpret_cast<uint4*>(slice + (uint32_t)lane * row_bytes + ((chunk ^ sw) << 4)) =
                        make_uint4(pack_f16x2(f[0], f[1]), pack_f16x2(f[2], f[3]), pack_f16x2(f[4], f[5]), pack_f16x2(f[6], f[7]));
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&p.tmOut),
                             "r"(smem_u32(slice)), "r"(0), "r"(pos.seg * kStrip + ew * 32), "r"(pos.b * p.H + pos.y)
                             : "memory");
                bulk_commit();
            }
        }
        if (lane == 0) bulk_wait_read_0();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)(4 * C < 32 ? 32 : 4 * C));
    }
}

}  // namespace

// fp32 NCHW frames (x_u8 == null) or uint8 planes [B, 3, H, W] (x_f32 == null, values scaled by 1/255 like prep_image)
int launch_stem_tc(const float* x_f32, const unsigned char* x_u8, int B, int H, int W, const float* w, const float* bias,
                   int Cout, int leaky, Act out, cudaStream_t stream) {
    static EncodeTiledFn encode_tiled = nullptr;
    if (!encode_tiled) {
        const int rc = driver_fn("cuTensorMapEncodeTiled", reinterpret_cast<void**>(&encode_tiled));
        if (rc) return rc;
    }
    const bool u8 = x_u8 != nullptr;
    if (!out.f16 || out.fp32 || (Cout != 16 && Cout != 32 && Cout != 64) || W < kStrip + 32 || W % (u8 ? 16 : 4) != 0 ||
        out.pitch % 8 != 0 || (reinterpret_cast<uintptr_t>(out.ptr) & 15u))
        return fail(RTOD_ERR_UNSUPPORTED, "stem_tc: fp16 output, 16/32/64 filters, width >= %d (multiple of %d)", kBoxW + 8, u8 ? 16 : 4);
    StemTcParams p{};
    p.w = w; p.bias = bias;
    p.B = B; p.H = H; p.W = W; p.C = Cout; p.leaky = leaky;
    p.strips_per_row = (W + kStrip - 1) / kStrip;
    p.total_strips = B * H * p.strips_per_row;
    // uint8 frames enter as value / 256 (exact in fp16); the weights carry the remaining 256 / 255 of prep_image's
    // value / 255 (scaling them by 1 / 255 instead would push their lo terms into the fp16 subnormals)
    p.in_scale = u8 ? 256.0f / 255.0f : 1.0f;
    p.no_relay = getenv("RTOD_STEM_NO_RELAY") != nullptr;
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)3 * B};
    const cuuint64_t strides[2] = {(cuuint64_t)W * (u8 ? 1 : 4), (cuuint64_t)W * H * (u8 ? 1 : 4)};
    const cuuint32_t box[3] = {(cuuint32_t)(u8 ? kStrip + 32 : kBoxW), 3, 3};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode_tiled(&p.tmX, u8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                                    u8 ? (void*)x_u8 : (void*)x_f32, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(RTOD_ERR_CUDA, "cuTensorMapEncodeTiled (stem_tc image) failed: %d", (int)r);
    {
        const cuuint64_t odims[3] = {(cuuint64_t)Cout, (cuuint64_t)W, (cuuint64_t)B * H};
        const cuuint64_t ostrides[2] = {(cuuint64_t)out.pitch * 2, (cuuint64_t)out.pitch * 2 * W};
        const cuuint32_t obox[3] = {(cuuint32_t)Cout, 32, 1};
        const CUtensorMapSwizzle sw = Cout == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (Cout == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
        const CUresult ro = encode_tiled(&p.tmOut, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, out.ptr, odims, ostrides, obox, estr,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (ro != CUDA_SUCCESS) return fail(RTOD_ERR_CUDA, "cuTensorMapEncodeTiled (stem_tc output) failed: %d", (int)ro);
    }
    const size_t stage_pitch = ((size_t)9 * (u8 ? kStrip + 32 : kBoxW) * (u8 ? 1 : 4) + 127) & ~(size_t)127;
    const size_t smem = 1024 + 4 * (size_t)kATileBytes + (size_t)Cout * 192 + 8 * (size_t)Cout * 64 + kStages * stage_pitch + 512;
    // persistent CTAs: exactly as many as are resident at once (a second wave would double the time).  Residency is
    // computed here from shared memory (64 registers x 256 threads and 4C TMEM columns allow four):
    // cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for every kernel that allocates tensor memory.
    // (Without the carve-out preference the driver may size shared memory for a single block per SM.)
    RTOD_CUDA_OK(cudaFuncSetAttribute(stem_tc_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    RTOD_CUDA_OK(cudaFuncSetAttribute(stem_tc_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    RTOD_CUDA_OK(cudaFuncSetAttribute(stem_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RTOD_CUDA_OK(cudaFuncSetAttribute(stem_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = (int)((227u * 1024u) / (smem + 1024));
    if (per_sm > 3) per_sm = 3;                                    // __launch_bounds__(288, 3)
    if (const char* e = getenv("RTOD_STEM_CTAS")) per_sm = atoi(e) >= 1 && atoi(e) <= 4 ? atoi(e) : per_sm;
    if (per_sm < 1) return fail(RTOD_ERR_CUDA, "stem_tc: kernel does not fit an SM");
    int grid = kNumSMs * per_sm;
    if (grid > p.total_strips) grid = p.total_strips;
    if (u8) stem_tc_kernel<true><<<grid, 288, smem, stream>>>(p);
    else stem_tc_kernel<false><<<grid, 288, smem, stream>>>(p);
    RTOD_LAUNCH_OK("stem_tc_kernel");
    return RTOD_OK;
}

}  // namespace rtod
