// stem.cu -- the first convolution block (src/darknet.py:488-501 with Cin = 3: 3x3, stride 1, pad 1,
// BatchNorm folded, leaky 0.1) straight from the NCHW fp32 image to NHWC bf16.
//
// HBM-bound by construction (it writes the largest activation of the network: B*H*W*Cout bf16), so
// the design is about never waiting on memory: persistent CTAs walk "strips" (one output row segment
// of <= 240 pixels of one image); a 3-D TMA tiled load (box {SW+8, 3 rows, 3 channels}, halo and
// image border zero-filled by the TMA bounds check) stages the 3x3 neighbourhood of a whole strip in
// shared memory four strips ahead; each warp then builds the im2col fragments of 16 pixels from
// shared memory (no bounds checks, no global latency) and multiplies with register-resident weights
// on the warp-level tensor-core path (mma.sync m16n8k16, fp32 accumulate).  K = 27 is one K-step:
// the tcgen05/TMEM pipeline has nothing to pipeline here.  The image is split into bf16 hi + lo parts
// (two MMAs), the weights are bf16 like every other layer's.
#include <cstdlib>

#include "layers.cuh"
#include "tc_ptx.cuh"

namespace rtod {

namespace {

constexpr int kStemStages = 4;

struct StemParams {
    CUtensorMap tmX;          // {W, H, 3*B} fp32, box {box_w, 3, 3}
    const float* w;           // [Cout][27] fp32, BN folded
    const float* bias;
    int* err_flag_unused;
    Act out;
    int B, H, W, leaky;
    int SW, box_w, strips_per_row, total_strips, tiles_per_strip;
    int dbg;
};

template <bool kF16>
__device__ __forceinline__ void stem_mma(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    if constexpr (kF16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// kF16: fp16 operands, weights AND image as two-term sums (three products: the stem's weight rounding would
// otherwise be the largest single error source of the network, it touches the most output elements);
// bf16: image hi + lo, single bf16 weights (the north-star wording's arithmetic)
template <int NT, bool kF16>
__global__ void __launch_bounds__(224, 4) stem_tma_kernel(const __grid_constant__ StemParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 127u) & ~127u) - raw_addr);
    const uint32_t stage_bytes = (uint32_t)9 * p.box_w * 4;            // multiple of 16 (box_w % 4 == 0)
    const uint32_t stage_pitch = (stage_bytes + 127u) & ~127u;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStemStages * stage_pitch);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int nwarps = blockDim.x >> 5;
    if (threadIdx.x == 0) {
        prefetch_tmap(&p.tmX);
        for (int s = 0; s < kStemStages; ++s) mbar_init(&full[s], 1);
        *reinterpret_cast<uint32_t*>(full + kStemStages) = 0u;         // the zero word padding taps read
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // tap offsets inside a stage for the 8 K indices this thread feeds: k = 16*s + 2*t + d + 8*hh
    // byte offsets inside a stage; the five padding taps (k >= 27) read a zero word kept behind the barriers
    uint32_t koff[8];
    uint32_t* zero_word = reinterpret_cast<uint32_t*>(full + kStemStages);
    const uint32_t smem_base = smem_u32(smem), zero_addr = smem_u32(zero_word);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int k = 16 * (i >> 2) + 2 * t + (i & 1) + 8 * ((i >> 1) & 1);
        koff[i] = k < 27 ? (uint32_t)((k / 3) * p.box_w + (k % 3) + 3) * 4u : 0xFFFFFFFFu;   // (c*3+ky)*box_w + kx + 3
    }
    uint32_t bfr[NT][2][2];                              // B fragments: W[n = 8j + g][k]
    uint32_t blo[kF16 ? NT : 1][2][2];                   // second term of the weights (fp16 mode)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int k0 = 16 * s + 2 * t + 8 * hh;
                const float w0 = k0 < 27 ? __ldg(p.w + (8 * j + g) * 27 + k0) : 0.0f;
                const float w1 = k0 + 1 < 27 ? __ldg(p.w + (8 * j + g) * 27 + k0 + 1) : 0.0f;
                bfr[j][s][hh] = pack_h2<kF16>(w0, w1);
                if constexpr (kF16)
                    blo[j][s][hh] = pack_h2<kF16>(w0 - h2_lo<kF16>(bfr[j][s][hh]), w1 - h2_hi<kF16>(bfr[j][s][hh]));
            }
    float bia[NT][2];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        bia[j][0] = __ldg(p.bias + 8 * j + 2 * t);
        bia[j][1] = __ldg(p.bias + 8 * j + 2 * t + 1);
    }

    // strip -> (segment, row, image) advances by gridDim.x per iteration: carried additions instead of
    // three integer divides (~450 cycles) on every warp's critical path
    struct Pos { int seg, y, b; };
    const int g_seg = (int)(gridDim.x % (unsigned)p.strips_per_row);
    const int g_y = (int)((gridDim.x / (unsigned)p.strips_per_row) % (unsigned)p.H);
    const int g_b = (int)(gridDim.x / (unsigned)(p.strips_per_row * p.H));
    auto advance = [&](Pos& q) {
        q.seg += g_seg;
        const int c0 = q.seg >= p.strips_per_row;
        q.seg -= c0 * p.strips_per_row;
        q.y += g_y + c0;
        const int c1 = q.y >= p.H;
        q.y -= c1 * p.H;
        q.b += g_b + c1;
    };
    Pos first;
    first.seg = (int)(blockIdx.x % (unsigned)p.strips_per_row);
    first.y = (int)((blockIdx.x / (unsigned)p.strips_per_row) % (unsigned)p.H);
    first.b = (int)(blockIdx.x / (unsigned)(p.strips_per_row * p.H));
    Pos ipos = first;                                 // thread 0: position of the next strip to load
    int iloc = 0;
    auto issue = [&]() {                              // thread 0: this CTA's next strip -> stage iloc % S
        if (ipos.b >= p.B || p.dbg == 1) return;
        const int s = iloc % kStemStages;
        mbar_expect_tx(&full[s], stage_bytes);
        // the innermost start coordinate must be 16-byte aligned: start 4 columns (not 1) left of the strip,
        // so shared-memory column j holds image column seg*SW - 4 + j
        tma_load_3d(smem + s * stage_pitch, &p.tmX, &full[s], ipos.seg * p.SW - 4, ipos.y - 1, 3 * ipos.b);
        advance(ipos);
        ++iloc;
    };
    if (threadIdx.x == 0)
        for (int l = 0; l < kStemStages - 1; ++l) issue();

    unsigned short* obase = reinterpret_cast<unsigned short*>(p.out.ptr);
    Pos pos = first;
    int s = 0;
    uint32_t parity = 0;
    for (; pos.b < p.B; advance(pos)) {
        // the stage consumed in the previous iteration is free (all warps passed its barrier)
        if (threadIdx.x == 0) issue();
        if (p.dbg != 1) {
            unsigned spins = 0;
            while (!mbar_try_wait(&full[s], parity))
                if (++spins > (1u << 26)) break;     // never in practice; avoids a hard hang
        }
        const int y = pos.y, b = pos.b;
        const int x_first = pos.seg * p.SW;
        for (int tile = warp; tile < p.tiles_per_strip && p.dbg != 2; tile += nwarps) {
            const int px0 = tile * 16;
            if (x_first + px0 >= p.W) break;
            uint32_t ahi[2][4], alo[2][4];
            const uint32_t stage_addr = smem_base + (uint32_t)s * stage_pitch + (uint32_t)(px0 + g) * 4u;
#pragma unroll
            for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                for (int i = 0; i < 8; i += 2) {
                    float v0, v1;
                    const uint32_t a0 = koff[i] == 0xFFFFFFFFu ? zero_addr : stage_addr + rr * 32u + koff[i];
                    const uint32_t a1 = koff[i + 1] == 0xFFFFFFFFu ? zero_addr : stage_addr + rr * 32u + koff[i + 1];
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v0) : "r"(a0));
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v1) : "r"(a1));
                    // image = hi + lo: two MMAs keep ~16 (bf16) / ~22 (fp16) mantissa bits of the pixels
                    const uint32_t hi = pack_h2<kF16>(v0, v1);
                    ahi[i >> 2][((i >> 1) & 1) * 2 + rr] = hi;
                    alo[i >> 2][((i >> 1) & 1) * 2 + rr] = pack_h2<kF16>(v0 - h2_lo<kF16>(hi), v1 - h2_hi<kF16>(hi));
                }
            float acc[NT][4];
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                acc[j][0] = bia[j][0]; acc[j][1] = bia[j][1]; acc[j][2] = bia[j][0]; acc[j][3] = bia[j][1];
#pragma unroll
                for (int s2 = 0; s2 < 2; ++s2) {
                    stem_mma<kF16>(acc[j], alo[s2], bfr[j][s2]);
                    if constexpr (kF16) stem_mma<kF16>(acc[j], ahi[s2], blo[j][s2]);
                    stem_mma<kF16>(acc[j], ahi[s2], bfr[j][s2]);
                }
            }
            const long long row_pix = ((long long)b * p.H + y) * p.W;
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int x = x_first + px0 + g + 8 * rr;
                if (x >= p.W) continue;
                unsigned short* dst = obase + (row_pix + x) * p.out.pitch + 2 * t;
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    float v0 = acc[j][2 * rr], v1 = acc[j][2 * rr + 1];
                    if (p.leaky) {
                        v0 = fmaxf(v0, 0.1f * v0);
                        v1 = fmaxf(v1, 0.1f * v1);
                    }
                    *reinterpret_cast<uint32_t*>(dst + 8 * j) = pack_h2<kF16>(v0, v1);
                }
            }
        }
        __syncthreads();                              // every warp is done with stage s
        if (++s == kStemStages) {
            s = 0;
            parity ^= 1u;
        }
    }
}

}  // namespace

// returns RTOD_ERR_UNSUPPORTED (without touching the error text semantics) when the shape is not covered
int launch_stem_tma(const float* x, int B, int H, int W, const float* w, const float* bias, int Cout, int leaky,
                    Act out, cudaStream_t stream) {
    static EncodeTiledFn encode_tiled = nullptr;
    if (!encode_tiled) {
        const int rc = driver_fn("cuTensorMapEncodeTiled", reinterpret_cast<void**>(&encode_tiled));
        if (rc) return rc;
    }
    StemParams p{};
    int n_seg = (W + 239) / 240;
    if (n_seg < 2) n_seg = 2;                                  // the box (SW + 4 columns) must not exceed the image width
    p.SW = ((W + n_seg - 1) / n_seg + 15) / 16 * 16;            // strip width, multiple of 16, <= 240
    p.box_w = p.SW + 8;                                        // 4 columns left (alignment), 1 + 3 right
    if (p.box_w > W) return fail(RTOD_ERR_UNSUPPORTED, "stem_tma: image width %d too small", W);
    p.strips_per_row = (W + p.SW - 1) / p.SW;
    p.total_strips = B * H * p.strips_per_row;
    p.tiles_per_strip = p.SW / 16;
    p.dbg = getenv("RTOD_STEM_DBG") ? atoi(getenv("RTOD_STEM_DBG")) : 0;
    p.w = w; p.bias = bias; p.out = out; p.B = B; p.H = H; p.W = W; p.leaky = leaky;
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)3 * B};
    const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    const cuuint32_t box[3] = {(cuuint32_t)p.box_w, 3, 3};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode_tiled(&p.tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(x), dims, strides, box,
                                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(RTOD_ERR_CUDA, "cuTensorMapEncodeTiled (stem image) failed: %d", (int)r);
    int warps = p.tiles_per_strip;
    if (warps > 7) warps = (warps + 1) / 2;                     // two tiles per warp and strip
    if (warps > 7) warps = 7;
    if (warps < 2) warps = 2;
    const size_t smem = (size_t)kStemStages * (((size_t)9 * p.box_w * 4 + 127) & ~(size_t)127) + kStemStages * 8 + 16 + 256;
    const int per_sm = 4;                                        // <= 72 registers/thread: four CTAs of <= 7 warps per SM
    int grid = kNumSMs * per_sm;
    if (grid > p.total_strips) grid = p.total_strips;
    if (out.f16) {
        if (Cout == 32) stem_tma_kernel<4, true><<<grid, warps * 32, smem, stream>>>(p);
        else if (Cout == 16) stem_tma_kernel<2, true><<<grid, warps * 32, smem, stream>>>(p);
        else stem_tma_kernel<8, true><<<grid, warps * 32, smem, stream>>>(p);
    } else {
        if (Cout == 32) stem_tma_kernel<4, false><<<grid, warps * 32, smem, stream>>>(p);
        else if (Cout == 16) stem_tma_kernel<2, false><<<grid, warps * 32, smem, stream>>>(p);
        else stem_tma_kernel<8, false><<<grid, warps * 32, smem, stream>>>(p);
    }
    RTOD_LAUNCH_OK("stem_tma_kernel");
    return RTOD_OK;
}

}  // namespace rtod
