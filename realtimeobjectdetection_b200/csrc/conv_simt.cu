// conv_simt.cu -- CUDA-core implicit-GEMM convolution over the same NHWC bf16 tensors, K-major
// bf16 weights and fused epilogue (bias, leaky 0.1, shortcut add) as the tcgen05 kernel in
// conv_tc.cu.  It is NOT the product path: it exists (a) as the on-device cross-check the
// parity tests run next to the tcgen05 kernel (plan flag RTOD_PLAN_CONV_SIMT) and (b) for
// shapes the tensor-core kernel does not tile (Cin not a multiple of 16, kernels other than
// 1x1/3x3).  Any shape, no alignment requirements beyond the 8-channel vector width.
#include "layers.cuh"

namespace rtod {

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(256) conv_simt_kernel(ConvArgs a) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int Ho = a.out.H, Wo = a.out.W, H = a.in.H, W = a.in.W;
    const long long M = (long long)a.B * Ho * Wo;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const __nv_bfloat16* in = reinterpret_cast<const __nv_bfloat16*>(a.in.ptr);

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
                    v = __bfloat162float(in[((b * H + iy) * W + ix) * a.in.pitch + c]);
            }
            As[kk][mm] = v;
        }
        for (int e = tid; e < BN * BK; e += 256) {
            const int kk = e % BK, nn = e / BK;
            const int k = k0 + kk, n = n0 + nn;
            Bs[kk][nn] = (k < a.K && n < a.Cout) ? __bfloat162float(a.w[(long long)n * a.K + k]) : 0.0f;
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
            if (a.res) v += __bfloat162float(a.res[m * a.res_pitch + n]);
            if (a.out.fp32) reinterpret_cast<float*>(a.out.ptr)[m * a.out.pitch + n] = v;
            else reinterpret_cast<__nv_bfloat16*>(a.out.ptr)[m * a.out.pitch + n] = __float2bfloat16_rn(v);
        }
    }
}

}  // namespace

int launch_conv_simt(const ConvArgs& a, cudaStream_t stream) {
    const long long M = (long long)a.B * a.out.H * a.out.W;
    dim3 grid(ceil_div(M, BM), ceil_div(a.Cout, BN));
    conv_simt_kernel<<<grid, 256, 0, stream>>>(a);
    RTOD_LAUNCH_OK("conv_simt_kernel");
    return RTOD_OK;
}

}  // namespace rtod
