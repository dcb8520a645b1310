// layers.cu -- the non-GEMM layers of Darknet.forward (src/darknet.py:199-295) on NHWC fp16/bf16
// activations: stem convolution (Cin = 3, fused NCHW fp32 -> NHWC 16-bit), max-pool
// (src/darknet.py:17-46, 547-555), bilinear x2 upsample (src/darknet.py:591-592), the
// copy/add fall-backs for route/shortcut (src/darknet.py:263-290) when they cannot be fused
// into a convolution, and the weight ingest (BatchNorm fold + K-major fp16/bf16 re-layout,
// src/darknet.py:316-410).  All of them are HBM-bound: 16-byte vector accesses, one pass.
#include "layers.cuh"

#include <cstdlib>

namespace rtod {

namespace {

__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

template <bool kF16> __device__ __forceinline__ uint32_t hmax2_u32(uint32_t a, uint32_t b) {
    if constexpr (kF16) {
        const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
        return *reinterpret_cast<const uint32_t*>(&r);
    } else {
        const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
        return *reinterpret_cast<const uint32_t*>(&r);
    }
}

// warp-level tensor-core step D += A(16x16) * B(16x8), fp32 accumulate, fp16 or bf16 operands
template <bool kF16>
__device__ __forceinline__ void mma_m16n8k16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    if constexpr (kF16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// ---------------------------------------------------------------------------------------------
// stem: 3x3 convolution over a 3-channel NCHW fp32 image, bias + leaky fused, NHWC bf16 out
// ---------------------------------------------------------------------------------------------
template <bool kF16>
__global__ void __launch_bounds__(128)
stem_conv3x3_kernel(const float* __restrict__ x, int B, int H, int W, const float* __restrict__ w,
                    const float* __restrict__ bias, int Cout, int stride, int pad, int leaky, Act out) {
    extern __shared__ float stem_w[];                     // [Cout][27] then [Cout] bias
    for (int i = threadIdx.x; i < Cout * 27; i += 128) stem_w[i] = w[i];
    for (int i = threadIdx.x; i < Cout; i += 128) stem_w[Cout * 27 + i] = bias[i];
    __syncthreads();
    const int Ho = out.H, Wo = out.W;
    const long long pix = blockIdx.x * 128ll + threadIdx.x;
    if (pix >= (long long)B * Ho * Wo) return;
    const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho), b = (int)(pix / ((long long)Wo * Ho));
    float in[27];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
                in[(c * 3 + ky) * 3 + kx] =
                    (iy >= 0 && iy < H && ix >= 0 && ix < W)
                        ? __ldg(x + (((long long)b * 3 + c) * H + iy) * W + ix)
                        : 0.0f;
            }
    unsigned short* dst = reinterpret_cast<unsigned short*>(out.ptr) + pix * out.pitch;
    for (int n0 = 0; n0 < Cout; n0 += 8) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = stem_w[Cout * 27 + n0 + j];
#pragma unroll
        for (int k = 0; k < 27; ++k)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(in[k], stem_w[(n0 + j) * 27 + k], acc[j]);
        if (leaky)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = leaky01(acc[j]);
        uint4 v;
        v.x = pack_h2<kF16>(acc[0], acc[1]);
        v.y = pack_h2<kF16>(acc[2], acc[3]);
        v.z = pack_h2<kF16>(acc[4], acc[5]);
        v.w = pack_h2<kF16>(acc[6], acc[7]);
        *reinterpret_cast<uint4*>(dst + n0) = v;
    }
}

// ---------------------------------------------------------------------------------------------
// stem on the warp-level tensor-core path: K = 27 is far too thin for the tcgen05/TMA pipeline
// (one K step), and the layer is HBM-bound (it writes the largest activation of the network), so
// each warp builds the im2col fragments of 16 output pixels directly in registers from the NCHW
// fp32 image, multiplies with register-resident weights via mma.sync.m16n8k16 (fp32 accumulate;
// image and weights are each split into bf16 hi + lo parts, three products, so the result is
// fp32-accurate) and stores bf16 NHWC.  The gather of the next tile is issued before the MMAs of
// the current one.  NT = Cout / 8.
// ---------------------------------------------------------------------------------------------
template <int NT, bool kF16>
__global__ void __launch_bounds__(128)
stem_conv3x3_mma_kernel(const float* __restrict__ x, int B, int H, int W, const float* __restrict__ w,
                        const float* __restrict__ bias, int stride, int pad, int leaky, Act out) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    // the 8 K indices this thread feeds (A and B fragments share them): k = 16*s + 2*t + d + 8*hh,
    // i = 4*s + 2*hh + d.  krel = offset of the tap relative to the pixel's (c=0, iy0, ix0) element.
    int krel[8], kdy[8], kdx[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int k = 16 * (i >> 2) + 2 * t + (i & 1) + 8 * ((i >> 1) & 1);
        const int c = k / 9, r = k % 9;
        kdy[i] = k < 27 ? r / 3 - pad : -100000;             // invalid taps never pass the bounds test
        kdx[i] = r % 3 - pad;
        krel[i] = c * H * W + kdy[i] * W + kdx[i];
    }
    // B fragments (W[n][k], n = 8*j + g), split w = hi + lo in bf16 so that with the same split of the
    // image the products hi*hi + lo*hi + hi*lo carry ~16 mantissa bits (the stem stays fp32-accurate)
    uint32_t bhi[NT][2][2], blo[NT][2][2];
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int k0 = 16 * s + 2 * t + 8 * hh;
                const float w0 = k0 < 27 ? __ldg(w + (8 * j + g) * 27 + k0) : 0.0f;
                const float w1 = k0 + 1 < 27 ? __ldg(w + (8 * j + g) * 27 + k0 + 1) : 0.0f;
                bhi[j][s][hh] = pack_h2<kF16>(w0, w1);
                blo[j][s][hh] = pack_h2<kF16>(w0 - h2_lo<kF16>(bhi[j][s][hh]), w1 - h2_hi<kF16>(bhi[j][s][hh]));
            }
    float bia[NT][2];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        bia[j][0] = __ldg(bias + 8 * j + 2 * t);
        bia[j][1] = __ldg(bias + 8 * j + 2 * t + 1);
    }

    const int Ho = out.H, Wo = out.W;
    const long long P = (long long)B * Ho * Wo;
    const long long n_tiles = (P + 15) / 16;
    const long long warps = (long long)gridDim.x * 4;
    unsigned short* obase = reinterpret_cast<unsigned short*>(out.ptr);

    auto gather = [&](long long tile, float (&v)[16]) {     // rows g and g + 8 of the tile, 8 taps each
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const long long pix = tile * 16 + g + 8 * rr;
            const bool pv = pix < P;
            const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho);
            const long long b = pix / ((long long)Wo * Ho);
            const int iy0 = oy * stride, ix0 = ox * stride;
            const float* base = x + b * 3 * H * W + (long long)iy0 * W + ix0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int iy = iy0 + kdy[i], ix = ix0 + kdx[i];
                v[rr * 8 + i] = (pv && iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(base + krel[i]) : 0.0f;
            }
        }
    };

    long long tile = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    float cur[16], nxt[16];
    if (tile < n_tiles) gather(tile, cur);
    for (; tile < n_tiles; tile += warps) {
        if (tile + warps < n_tiles) gather(tile + warps, nxt);            // in flight during the MMAs/stores
        uint32_t ahi[2][4], alo[2][4];                       // [k step][a0a1, a2a3, a4a5, a6a7]
#pragma unroll
        for (int rr = 0; rr < 2; ++rr)
#pragma unroll
            for (int s = 0; s < 2; ++s)
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const float v0 = cur[rr * 8 + s * 4 + hh * 2], v1 = cur[rr * 8 + s * 4 + hh * 2 + 1];
                    const uint32_t hi = pack_h2<kF16>(v0, v1);
                    ahi[s][hh * 2 + rr] = hi;
                    alo[s][hh * 2 + rr] = pack_h2<kF16>(v0 - h2_lo<kF16>(hi), v1 - h2_hi<kF16>(hi));
                }
        float acc[NT][4];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            acc[j][0] = bia[j][0]; acc[j][1] = bia[j][1]; acc[j][2] = bia[j][0]; acc[j][3] = bia[j][1];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
#define RTOD_MMA(A, Bf) mma_m16n8k16<kF16>(acc[j], A[s], Bf[j][s])
                RTOD_MMA(alo, bhi);
                RTOD_MMA(ahi, blo);
                RTOD_MMA(ahi, bhi);
#undef RTOD_MMA
            }
        }
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const long long pix = tile * 16 + g + 8 * rr;
            if (pix >= P) continue;
            unsigned short* dst = obase + pix * out.pitch + 2 * t;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                float v0 = acc[j][2 * rr], v1 = acc[j][2 * rr + 1];
                if (leaky) {
                    v0 = leaky01(v0);
                    v1 = leaky01(v1);
                }
                *reinterpret_cast<uint32_t*>(dst + 8 * j) = pack_h2<kF16>(v0, v1);
            }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) cur[i] = nxt[i];
    }
}

// ---------------------------------------------------------------------------------------------
// layout converters (plan input when the first layer is not a stem; debug read-back)
// ---------------------------------------------------------------------------------------------
template <bool kF16>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, int B, Act out) {
    const int HW = out.H * out.W, groups = out.C / 8;
    const long long total = (long long)B * HW * groups;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(i % HW);
        const int g = (int)((i / HW) % groups);
        const int b = (int)(i / ((long long)HW * groups));
        const float* src = x + ((long long)b * out.C + g * 8) * HW + p;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldg(src + (long long)j * HW);
        uint4 o;
        o.x = pack_h2<kF16>(v[0], v[1]);
        o.y = pack_h2<kF16>(v[2], v[3]);
        o.z = pack_h2<kF16>(v[4], v[5]);
        o.w = pack_h2<kF16>(v[6], v[7]);
        *reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(out.ptr) +
                                  ((long long)b * HW + p) * out.pitch + g * 8) = o;
    }
}

template <bool kF16>
__global__ void nhwc_to_nchw_kernel(Act in, int B, float* __restrict__ out) {
    const int HW = in.H * in.W;
    const long long total = (long long)B * in.C * HW;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(i % HW);
        const int c = (int)((i / HW) % in.C);
        const int b = (int)(i / ((long long)HW * in.C));
        const long long src = ((long long)b * HW + p) * in.pitch + c;
        out[i] = in.fp32 ? reinterpret_cast<const float*>(in.ptr)[src]
                         : h1_to_float<kF16>(reinterpret_cast<const unsigned short*>(in.ptr)[src]);
    }
}

// ---------------------------------------------------------------------------------------------
// max-pool: MaxPool2d(size, stride) (floor mode, no padding) or, for stride 1, the reference's
// MaxPoolStride1: replicate-pad right/bottom by size-1, then pool with stride size-1
// ---------------------------------------------------------------------------------------------
template <typename Index, bool kF16>                        // 32-bit index arithmetic where the tensor allows
__global__ void maxpool_kernel(Act in, Act out, int B, int size, int step, int clamp_edge) {
    const int groups = out.C / 8;
    const Index total = (Index)B * out.H * out.W * groups;
    for (Index i = blockIdx.x * (Index)blockDim.x + threadIdx.x; i < total; i += (Index)gridDim.x * blockDim.x) {
        const Index pix = i / (Index)groups;
        const int g = (int)(i - pix * (Index)groups);
        const Index line = pix / (Index)out.W;
        const int ox = (int)(pix - line * (Index)out.W);
        const int b = (int)(line / (Index)out.H);
        const int oy = (int)(line - (Index)b * (Index)out.H);
        const unsigned short* base = reinterpret_cast<const unsigned short*>(in.ptr) + g * 8;
        uint32_t best[4];
        bool first = true;
        for (int dy = 0; dy < size; ++dy) {
            int iy = oy * step + dy;
            if (iy >= in.H) {
                if (!clamp_edge) continue;
                iy = in.H - 1;
            }
            for (int dx = 0; dx < size; ++dx) {
                int ix = ox * step + dx;
                if (ix >= in.W) {
                    if (!clamp_edge) continue;
                    ix = in.W - 1;
                }
                const uint4 v = ldg16(base + (((long long)b * in.H + iy) * in.W + ix) * in.pitch);
                const uint32_t* h = &v.x;
                if (first) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) best[j] = h[j];
                    first = false;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) best[j] = hmax2_u32<kF16>(best[j], h[j]);
                }
            }
        }
        *reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(out.ptr) +
                                  (((long long)b * out.H + oy) * out.W + ox) * out.pitch + g * 8) =
            make_uint4(best[0], best[1], best[2], best[3]);
    }
}

// ---------------------------------------------------------------------------------------------
// bilinear x2, align_corners=False: even rows .25/.75 of (i-1, i), odd rows .75/.25 of (i, i+1),
// indices clamped at the border; evaluated like ATen: h0*(w0*a + w1*b) + h1*(w0*c + w1*d)
// ---------------------------------------------------------------------------------------------
// Index: 32-bit where the tensor allows (64-bit divides cost ~100 instructions each; the kernel was issue-bound)
template <typename Index, bool kF16>
__global__ void upsample2x_kernel(Act in, Act out, int B) {
    const int groups = out.C / 8;
    const Index total = (Index)B * out.H * out.W * groups;
    for (Index i = blockIdx.x * (Index)blockDim.x + threadIdx.x; i < total; i += (Index)gridDim.x * blockDim.x) {
        const Index pix = i / (Index)groups;
        const int g = (int)(i - pix * (Index)groups);
        const Index line = pix / (Index)out.W;
        const int ox = (int)(pix - line * (Index)out.W);
        const int b = (int)(line / (Index)out.H);
        const int oy = (int)(line - (Index)b * (Index)out.H);
        const float sy = fmaxf((oy + 0.5f) * 0.5f - 0.5f, 0.0f), sx = fmaxf((ox + 0.5f) * 0.5f - 0.5f, 0.0f);
        const int y0 = (int)sy, x0 = (int)sx;
        const int y1 = min(y0 + 1, in.H - 1), x1 = min(x0 + 1, in.W - 1);
        const float hy1 = sy - y0, hy0 = 1.0f - hy1, wx1 = sx - x0, wx0 = 1.0f - wx1;
        const unsigned short* base = reinterpret_cast<const unsigned short*>(in.ptr) + g * 8;
        const long long row0 = ((long long)b * in.H + y0) * in.W, row1 = ((long long)b * in.H + y1) * in.W;
        const uint4 a = ldg16(base + (row0 + x0) * in.pitch), bb = ldg16(base + (row0 + x1) * in.pitch);
        const uint4 c = ldg16(base + (row1 + x0) * in.pitch), d = ldg16(base + (row1 + x1) * in.pitch);
        const uint32_t* pa = &a.x; const uint32_t* pb = &bb.x; const uint32_t* pc = &c.x; const uint32_t* pd = &d.x;
        uint4 o;
        uint32_t* po = &o.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float lo = hy0 * (wx0 * h2_lo<kF16>(pa[j]) + wx1 * h2_lo<kF16>(pb[j])) +
                             hy1 * (wx0 * h2_lo<kF16>(pc[j]) + wx1 * h2_lo<kF16>(pd[j]));
            const float hi = hy0 * (wx0 * h2_hi<kF16>(pa[j]) + wx1 * h2_hi<kF16>(pb[j])) +
                             hy1 * (wx0 * h2_hi<kF16>(pc[j]) + wx1 * h2_hi<kF16>(pd[j]));
            po[j] = pack_h2<kF16>(lo, hi);
        }
        *reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(out.ptr) +
                                  (((long long)b * out.H + oy) * out.W + ox) * out.pitch + g * 8) = o;
    }
}

// Same arithmetic, one thread per INPUT pixel and 8-channel group: its 3 x 3 neighbourhood (nine 16-byte loads, clamped at
// the border) yields the 2 x 2 output block (2y..2y+1, 2x..2x+1) -- 2.25 loads and a quarter of the index arithmetic per
// output vector instead of 4 loads and three integer divides.  Each output is evaluated with exactly the expression of
// upsample2x_kernel (weights 0 / .25 / .75 are what its float formula produces): bit-identical results.
template <typename Index, bool kF16>
__global__ void upsample2x_block_kernel(Act in, Act out, int B) {
    const int groups = in.C / 8;
    const Index total = (Index)B * in.H * in.W * groups;
    for (Index i = blockIdx.x * (Index)blockDim.x + threadIdx.x; i < total; i += (Index)gridDim.x * blockDim.x) {
        const Index pix = i / (Index)groups;
        const int g = (int)(i - pix * (Index)groups);
        const Index line = pix / (Index)in.W;
        const int x = (int)(pix - line * (Index)in.W);
        const int b = (int)(line / (Index)in.H);
        const int y = (int)(line - (Index)b * (Index)in.H);
        const int ys[3] = {max(y - 1, 0), y, min(y + 1, in.H - 1)}, xs[3] = {max(x - 1, 0), x, min(x + 1, in.W - 1)};
        const unsigned short* base = reinterpret_cast<const unsigned short*>(in.ptr) + g * 8;
        uint4 v[3][3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) v[r][c] = ldg16(base + (((long long)b * in.H + ys[r]) * in.W + xs[c]) * in.pitch);
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
            // output row 2y + dy: (first source row, second source row, weight of the second) as upsample2x_kernel derives them
            const int r0 = dy == 0 ? (y > 0 ? 0 : 1) : 1, r1 = r0 + 1;
            const float hy1 = dy == 0 ? (y > 0 ? 0.75f : 0.0f) : 0.25f, hy0 = 1.0f - hy1;
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int c0 = dx == 0 ? (x > 0 ? 0 : 1) : 1, c1 = c0 + 1;
                const float wx1 = dx == 0 ? (x > 0 ? 0.75f : 0.0f) : 0.25f, wx0 = 1.0f - wx1;
                uint4 va, vb, vc, vd;                      // (select by run-time row / column without indexing the register array)
                va = r0 == 0 ? (c0 == 0 ? v[0][0] : v[0][1]) : (c0 == 0 ? v[1][0] : v[1][1]);
                vb = r0 == 0 ? (c1 == 1 ? v[0][1] : v[0][2]) : (c1 == 1 ? v[1][1] : v[1][2]);
                vc = r1 == 1 ? (c0 == 0 ? v[1][0] : v[1][1]) : (c0 == 0 ? v[2][0] : v[2][1]);
                vd = r1 == 1 ? (c1 == 1 ? v[1][1] : v[1][2]) : (c1 == 1 ? v[2][1] : v[2][2]);
                const uint32_t* pa = &va.x; const uint32_t* pb = &vb.x; const uint32_t* pc = &vc.x; const uint32_t* pd = &vd.x;
                uint4 o;
                uint32_t* po = &o.x;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float lo = hy0 * (wx0 * h2_lo<kF16>(pa[j]) + wx1 * h2_lo<kF16>(pb[j])) +
                                     hy1 * (wx0 * h2_lo<kF16>(pc[j]) + wx1 * h2_lo<kF16>(pd[j]));
                    const float hi = hy0 * (wx0 * h2_hi<kF16>(pa[j]) + wx1 * h2_hi<kF16>(pb[j])) +
                                     hy1 * (wx0 * h2_hi<kF16>(pc[j]) + wx1 * h2_hi<kF16>(pd[j]));
                    po[j] = pack_h2<kF16>(lo, hi);
                }
                *reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(out.ptr) +
                                          (((long long)b * out.H + 2 * y + dy) * out.W + 2 * x + dx) * out.pitch + g * 8) = o;
            }
        }
    }
}

__global__ void copy_kernel(Act in, Act out, int B) {
    const int groups = out.C / 8;
    const long long total = (long long)B * out.H * out.W * groups;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i % groups);
        const long long p = i / groups;
        *reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(out.ptr) + p * out.pitch + g * 8) =
            ldg16(reinterpret_cast<const unsigned short*>(in.ptr) + p * in.pitch + g * 8);
    }
}

template <bool kF16>
__global__ void add_kernel(Act a, Act b, Act out, int B) {
    const int groups = out.C / 8;
    const long long total = (long long)B * out.H * out.W * groups;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i % groups);
        const long long p = i / groups;
        const uint4 va = ldg16(reinterpret_cast<const unsigned short*>(a.ptr) + p * a.pitch + g * 8);
        const uint4 vb = ldg16(reinterpret_cast<const unsigned short*>(b.ptr) + p * b.pitch + g * 8);
        const uint32_t* pa = &va.x; const uint32_t* pb = &vb.x;
        uint4 o;
        uint32_t* po = &o.x;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            po[j] = pack_h2<kF16>(h2_lo<kF16>(pa[j]) + h2_lo<kF16>(pb[j]), h2_hi<kF16>(pa[j]) + h2_hi<kF16>(pb[j]));
        *reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(out.ptr) + p * out.pitch + g * 8) = o;
    }
}

// ---------------------------------------------------------------------------------------------
// weight ingest: BatchNorm fold (eval semantics) + [Cout,Cin,k,k] fp32 -> [Cout][(ky,kx,c)] bf16
// ---------------------------------------------------------------------------------------------
template <bool kF16>
__global__ void fold_pack_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                 const float* __restrict__ mean, const float* __restrict__ var,
                                 float eps, int Cout, int Cin, int ks, int Cout_pad, int w_split,
                                 unsigned short* __restrict__ wp, float* __restrict__ wf,
                                 float* __restrict__ bias_out) {
    const int K = Cin * ks * ks;
    const long long total = (long long)Cout * K;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int o = (int)(i / K), r = (int)(i % K);        // r indexes [c][ky][kx]
        const int c = r / (ks * ks), t = r % (ks * ks);
        const float scale = gamma ? gamma[o] * (1.0f / sqrtf(var[o] + eps)) : 1.0f;
        const float v = w[i] * scale;
        const unsigned short hi = float_to_h1<kF16>(v);
        const int rows = Cout_pad << w_split, PK = weight_pack_k(Cin, K), k = t * Cin + c;
        wp[packed_weight_index(o, k, rows, PK)] = hi;
        if (w_split)                                         // second term of the two-term weight (rows Cout_pad ..)
            wp[packed_weight_index(Cout_pad + o, k, rows, PK)] = float_to_h1<kF16>(v - h1_to_float<kF16>(hi));
        if (wf) wf[i] = v;
        if (r == 0) {
            float bv = bias ? bias[o] : 0.0f;
            if (gamma) bv = (bv - mean[o]) * scale + beta[o];
            bias_out[o] = bv;
        }
    }
}

int grid_for(long long total, int threads) {
    long long blocks = (total + threads - 1) / threads;
    const long long cap = (long long)kNumSMs * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace

int launch_stem_conv(const float* x, int B, int Cin, int H, int W, const float* w, const float* bias,
                     int Cout, int ks, int stride, int pad, int leaky, Act out, cudaStream_t stream) {
    if (Cin != 3 || ks != 3 || Cout % 8 != 0 || Cout > 256 || out.fp32)
        return fail(RTOD_ERR_UNSUPPORTED, "stem conv supports 3x3, Cin=3, Cout%%8==0, Cout<=256");
    // fp16 storage, stride-1 stems of 16/32/64 filters (both reference networks): tcgen05 with the im2col operand
    // built in shared memory, see stem_tc.cu
    if (getenv("RTOD_STEM_NO_TC") == nullptr && out.f16 && stride == 1 && pad == 1 && W % 4 == 0 && W >= 160 &&
        (Cout == 16 || Cout == 32 || Cout == 64) && (reinterpret_cast<uintptr_t>(x) & 15u) == 0 && out.H == H && out.W == W)
        return launch_stem_tc(x, nullptr, B, H, W, w, bias, Cout, leaky, out, stream);
    // stride-1 stems, bf16 storage: TMA-staged strips + warp-level MMA, see stem.cu
    if (getenv("RTOD_STEM_NO_TMA") == nullptr && stride == 1 && pad == 1 && W % 4 == 0 && W >= 64 && (Cout == 16 || Cout == 32 || Cout == 64) &&
        (reinterpret_cast<uintptr_t>(x) & 15u) == 0 && out.H == H && out.W == W)
        return launch_stem_tma(x, B, H, W, w, bias, Cout, leaky, out, stream);
    const long long pixels = (long long)B * out.H * out.W;
    long long blocks = (pixels + 63) / 64;                       // 4 warps x 16 pixels per pass
    if (blocks > (long long)kNumSMs * 16) blocks = (long long)kNumSMs * 16;
#define RTOD_STEM_MMA(NT)                                                                                              \
    do {                                                                                                               \
        if (out.f16) stem_conv3x3_mma_kernel<NT, true><<<(unsigned)blocks, 128, 0, stream>>>(x, B, H, W, w, bias, stride, pad, leaky, out);  \
        else stem_conv3x3_mma_kernel<NT, false><<<(unsigned)blocks, 128, 0, stream>>>(x, B, H, W, w, bias, stride, pad, leaky, out);         \
    } while (0)
    if (Cout == 32) RTOD_STEM_MMA(4);
    else if (Cout == 16) RTOD_STEM_MMA(2);
    else if (Cout == 64) RTOD_STEM_MMA(8);
    else if (out.f16)
        stem_conv3x3_kernel<true><<<ceil_div(pixels, 128), 128, (size_t)Cout * 28 * sizeof(float), stream>>>(
            x, B, H, W, w, bias, Cout, stride, pad, leaky, out);
    else
        stem_conv3x3_kernel<false><<<ceil_div(pixels, 128), 128, (size_t)Cout * 28 * sizeof(float), stream>>>(
            x, B, H, W, w, bias, Cout, stride, pad, leaky, out);
#undef RTOD_STEM_MMA
    RTOD_LAUNCH_OK("stem_conv3x3 kernel");
    return RTOD_OK;
}

int launch_nchw_to_nhwc(const float* x, int B, Act out, cudaStream_t stream) {
    if (out.C % 8 != 0 || out.fp32) return fail(RTOD_ERR_UNSUPPORTED, "input channels must be a multiple of 8");
    const long long total = (long long)B * out.H * out.W * (out.C / 8);
    if (out.f16) nchw_to_nhwc_kernel<true><<<grid_for(total, 256), 256, 0, stream>>>(x, B, out);
    else nchw_to_nhwc_kernel<false><<<grid_for(total, 256), 256, 0, stream>>>(x, B, out);
    RTOD_LAUNCH_OK("nchw_to_nhwc_kernel");
    return RTOD_OK;
}

int launch_nhwc_to_nchw(Act in, int B, float* out, cudaStream_t stream) {
    const long long total = (long long)B * in.C * in.H * in.W;
    if (in.f16) nhwc_to_nchw_kernel<true><<<grid_for(total, 256), 256, 0, stream>>>(in, B, out);
    else nhwc_to_nchw_kernel<false><<<grid_for(total, 256), 256, 0, stream>>>(in, B, out);
    RTOD_LAUNCH_OK("nhwc_to_nchw_kernel");
    return RTOD_OK;
}

#define RTOD_BY_INDEX_AND_TYPE(kernel, total, f16, ...)                                                        \
    do {                                                                                                       \
        if ((total) < (1ll << 31)) {                                                                           \
            if (f16) kernel<unsigned, true><<<grid_for(total, 256), 256, 0, stream>>>(__VA_ARGS__);            \
            else kernel<unsigned, false><<<grid_for(total, 256), 256, 0, stream>>>(__VA_ARGS__);               \
        } else {                                                                                               \
            if (f16) kernel<long long, true><<<grid_for(total, 256), 256, 0, stream>>>(__VA_ARGS__);           \
            else kernel<long long, false><<<grid_for(total, 256), 256, 0, stream>>>(__VA_ARGS__);              \
        }                                                                                                      \
    } while (0)

int launch_maxpool(Act in, Act out, int B, int size, int stride, cudaStream_t stream) {
    if (in.C % 8 != 0 || in.fp32 || out.fp32) return fail(RTOD_ERR_UNSUPPORTED, "maxpool needs 16-bit activations, C%%8==0");
    const long long total = (long long)B * out.H * out.W * (out.C / 8);
    const int step = stride != 1 ? stride : size - 1;           // src/darknet.py:35,45
    RTOD_BY_INDEX_AND_TYPE(maxpool_kernel, total, in.f16, in, out, B, size, step < 1 ? 1 : step, stride == 1);
    RTOD_LAUNCH_OK("maxpool_kernel");
    return RTOD_OK;
}

int launch_upsample2x(Act in, Act out, int B, cudaStream_t stream) {
    if (in.C % 8 != 0 || in.fp32 || out.fp32) return fail(RTOD_ERR_UNSUPPORTED, "upsample needs 16-bit activations, C%%8==0");
    if (out.H == 2 * in.H && out.W == 2 * in.W && getenv("RTOD_UPSAMPLE_PER_OUTPUT") == nullptr) {
        const long long blocks = (long long)B * in.H * in.W * (in.C / 8);      // one thread per input pixel and channel group
        RTOD_BY_INDEX_AND_TYPE(upsample2x_block_kernel, blocks, in.f16, in, out, B);
        RTOD_LAUNCH_OK("upsample2x_block_kernel");
        return RTOD_OK;
    }
    const long long total = (long long)B * out.H * out.W * (out.C / 8);
    RTOD_BY_INDEX_AND_TYPE(upsample2x_kernel, total, in.f16, in, out, B);
    RTOD_LAUNCH_OK("upsample2x_kernel");
    return RTOD_OK;
}
#undef RTOD_BY_INDEX_AND_TYPE

int launch_copy(Act in, Act out, int B, cudaStream_t stream) {
    if (in.C % 8 != 0 || in.fp32 || out.fp32) return fail(RTOD_ERR_UNSUPPORTED, "copy needs 16-bit activations, C%%8==0");
    const long long total = (long long)B * out.H * out.W * (out.C / 8);
    copy_kernel<<<grid_for(total, 256), 256, 0, stream>>>(in, out, B);
    RTOD_LAUNCH_OK("copy_kernel");
    return RTOD_OK;
}

int launch_add(Act a, Act b, Act out, int B, cudaStream_t stream) {
    if (out.C % 8 != 0 || a.fp32 || b.fp32 || out.fp32)
        return fail(RTOD_ERR_UNSUPPORTED, "shortcut add needs 16-bit activations, C%%8==0");
    const long long total = (long long)B * out.H * out.W * (out.C / 8);
    if (out.f16) add_kernel<true><<<grid_for(total, 256), 256, 0, stream>>>(a, b, out, B);
    else add_kernel<false><<<grid_for(total, 256), 256, 0, stream>>>(a, b, out, B);
    RTOD_LAUNCH_OK("add_kernel");
    return RTOD_OK;
}

int launch_fold_pack(const float* w, const float* bias, const float* gamma, const float* beta,
                     const float* mean, const float* var, float eps, int Cout, int Cin, int ks,
                     int Cout_pad, int f16, int w_split, void* wp, float* wf, float* bias_out,
                     cudaStream_t stream) {
    const long long total = (long long)Cout * Cin * ks * ks;
    unsigned short* wp16 = static_cast<unsigned short*>(wp);
    if (f16)
        fold_pack_kernel<true><<<grid_for(total, 256), 256, 0, stream>>>(w, bias, gamma, beta, mean, var, eps, Cout, Cin, ks,
                                                                         Cout_pad, w_split, wp16, wf, bias_out);
    else
        fold_pack_kernel<false><<<grid_for(total, 256), 256, 0, stream>>>(w, bias, gamma, beta, mean, var, eps, Cout, Cin, ks,
                                                                          Cout_pad, w_split, wp16, wf, bias_out);
    RTOD_LAUNCH_OK("fold_pack_kernel");
    return RTOD_OK;
}

}  // namespace rtod
