// conv_simt.cu -- CUDA-core implicit-GEMM convolution over the same NHWC fp16/bf16 tensors, K-major
// weights and fused epilogue (bias, leaky 0.1, shortcut add) as the tcgen05 kernel in
// conv_tc.cu.  It is NOT the product path: it exists (a) as the on-device cross-check the
// parity tests run next to the tcgen05 kernel (plan flag RTOD_PLAN_CONV_SIMT) and (b) for
// shapes the tensor-core kernel does not tile (Cin not a multiple of 16, kernels other than
// 1x1/3x3).  Any shape, no alignment requirements beyond the 8-channel vector width.
#include "layers.cuh"

namespace rtod {

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

template <bool kF16>
__global__ void __launch_bounds__(256) conv_simt_kernel(ConvArgs a) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int Ho = a.out.H, Wo = a.out.W, H = a.in.H, W = a.in.W;
    const long long M = (long long)a.B * Ho * Wo;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const unsigned short* in = reinterpret_cast<const unsigned short*>(a.in.ptr);
    const unsigned short* w16 = reinterpret_cast<const unsigned short*>(a.w);
    const unsigned short* res16 = reinterpret_cast<const unsigned short*>(a.res);

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    for (int k0 = 0; k0 < a.K; k0 += BK) {
        for (int e = tid; e < BM * BK; e += 256) {
            const int kk = e % BK, mm = e / BK;
            const int k = k0 + kk;
            const long long m = m0 + mm;
            float v = 0.0f;
            if (k < a.K && m < M) {
                const int tap = k / a.Cin, c = k % a.Cin;
                const int ky = tap / a.ks, kx = tap % a.ks;
                const int ox = (int)(m % Wo), oy = (int)((m / Wo) % Ho);
                const long long b = m / ((long long)Wo * Ho);
                const int iy = oy * a.stride - a.pad + ky, ix = ox * a.stride - a.pad + kx;
                if (iy >= 0 && iy < H && ix >= 0 && ix < W)
                    v = h1_to_float<kF16>(in[((b * H + iy) * W + ix) * a.in.pitch + c]);
            }
            As[kk][mm] = v;
        }
        for (int e = tid; e < BN * BK; e += 256) {
            const int kk = e % BK, nn = e / BK;
            const int k = k0 + kk, n = n0 + nn;
            float wv = 0.0f;
            if (k < a.K && n < a.Cout) {
                const int rows = a.Cout_pad << a.w_split, PK = weight_pack_k(a.Cin, a.K);
                wv = h1_to_float<kF16>(w16[packed_weight_index(n, k, rows, PK)]);
                if (a.w_split) wv += h1_to_float<kF16>(w16[packed_weight_index(a.Cout_pad + n, k, rows, PK)]);      // hi + lo
            }
            Bs[kk][nn] = wv;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= a.Cout) continue;
            float v = acc[i][j] + a.bias[n];
            if (a.leaky) v = leaky01(v);
            if (a.res) v += h1_to_float<kF16>(res16[m * a.res_pitch + n]);
            if (a.out.fp32) reinterpret_cast<float*>(a.out.ptr)[m * a.out.pitch + n] = v;
            else reinterpret_cast<unsigned short*>(a.out.ptr)[m * a.out.pitch + n] = float_to_h1<kF16>(v);
        }
    }
}

}  // namespace

int launch_conv_simt(const ConvArgs& a, cudaStream_t stream) {
    const long long M = (long long)a.B * a.out.H * a.out.W;
    dim3 grid(ceil_div(M, BM), ceil_div(a.Cout, BN));
    if (a.in.f16) conv_simt_kernel<true><<<grid, 256, 0, stream>>>(a);
    else conv_simt_kernel<false><<<grid, 256, 0, stream>>>(a);
    RTOD_LAUNCH_OK("conv_simt_kernel");
    return RTOD_OK;
}

}  // namespace rtod
