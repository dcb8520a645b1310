// tma_rate_probe.cu -- bring-up measurement (not part of the library): sustained throughput of one
// SM's TMA unit for the box shapes the convolution kernels use, with all 148 SMs loading at once
// from an L2-resident tensor.  Each CTA keeps `depth` loads in flight (ring of mbarriers) and
// issues `iters` loads; prints cycles per load and bytes/clk/SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_rate_probe tma_rate_probe.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

enum { MODE_2D = 0, MODE_IM2COL = 1, MODE_4D = 2 };

struct Args {
    CUtensorMap map;
    int mode, iters, depth;
    uint32_t bytes;          // per load
    int W, H, N, C;          // tensor dims for coordinate generation
    int rows;                // 2-D: rows of the matrix
    int box_w, box_h;        // 4-D
    int cbox;                // channels per box
    int variant;             // 0: lone thread, 1: converged warp + elect.sync
};

__global__ void __launch_bounds__(32) probe(const __grid_constant__ Args a, long long* cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = raw + (((smem_u32(raw) + 1023u) & ~1023u) - smem_u32(raw));
    const uint32_t slot = (a.bytes + 1023u) & ~1023u;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)a.depth * slot);
    if (threadIdx.x == 0) {
        for (int s = 0; s < a.depth; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (a.variant != 1 && threadIdx.x != 0) return;
    uint32_t seed = blockIdx.x * 7919u + 13u;
    const long long t0 = clock64();
    if (a.variant == 2) {                                    // bursts: `depth` loads back to back, then wait for all
        const int rounds = a.iters / a.depth;
        for (int r = 0; r < rounds; ++r) {
            for (int s = 0; s < a.depth; ++s) {
                seed = seed * 1664525u + 1013904223u;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(a.bytes) : "memory");
                const int r0 = (int)((seed >> 8) & 0x7FFFu);
                const int cc0 = (int)((seed >> 4) & (uint32_t)(a.C / a.cbox - 1)) * a.cbox;
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                                 smem_u32(smem + (size_t)s * slot)), "l"(&a.map), "r"(smem_u32(&bars[s])), "r"(cc0), "r"(r0) : "memory");
            }
            for (int s = 0; s < a.depth; ++s) {
                uint32_t done = 0;
                while (!done)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(done) : "r"(smem_u32(&bars[s])), "r"((uint32_t)(r & 1)) : "memory");
            }
        }
        cycles[blockIdx.x] = clock64() - t0;
        return;
    }
    long long t_wait = 0, t_arm = 0, t_issue = 0;
    for (int i = 0; i < a.iters + a.depth; ++i) {
        const int s = i % a.depth;
        const long long c0 = clock64();
        if (i >= a.depth) {                                   // wait for the load that used this slot
            const uint32_t parity = ((i / a.depth) - 1) & 1;
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(smem_u32(&bars[s])), "r"(parity) : "memory");
        }
        const long long c1 = clock64();
        t_wait += c1 - c0;
        if (i >= a.iters) continue;
        seed = seed * 1664525u + 1013904223u;
        if (a.variant == 1) {
            uint32_t is_leader = 0;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(is_leader));
            if (!is_leader) continue;
        }
        const long long c2 = clock64();
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(a.bytes) : "memory");
        const long long c3 = clock64();
        t_arm += c3 - c2;
        void* dst = smem + (size_t)s * slot;
        if (a.mode == MODE_2D) {
            const int r0 = (int)((seed >> 8) & 0x7FFFu);
            const int cc0 = (int)((seed >> 4) & (uint32_t)(a.C / a.cbox - 1)) * a.cbox;
            const long long c4 = clock64();
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                             smem_u32(dst)), "l"(&a.map), "r"(smem_u32(&bars[s])), "r"(cc0), "r"(r0) : "memory");
            t_issue += clock64() - c4;
        } else {
            const int n = (int)((seed >> 20) & (uint32_t)(a.N - 1));
            const int y = (int)((seed >> 10) & 31u);
            const int x = (int)((seed >> 2) & 15u);
            const int c0 = (int)((seed >> 6) & (uint32_t)(a.C / a.cbox - 1)) * a.cbox;
            if (a.mode == MODE_IM2COL)
                asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
                             " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(smem_u32(dst)), "l"(&a.map),
                             "r"(smem_u32(&bars[s])), "r"(c0), "r"(x - 1), "r"(y - 1), "r"(n), "h"((uint16_t)(seed & 1)),
                             "h"((uint16_t)((seed >> 1) & 1)) : "memory");
            else
                asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
                             " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)), "l"(&a.map), "r"(smem_u32(&bars[s])),
                             "r"(c0), "r"(x - 1), "r"(y - 1), "r"(n) : "memory");
        }
    }
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
    if (threadIdx.x == 0 && blockIdx.x == 0) { cycles[1000] = t_wait; cycles[1001] = t_arm; cycles[1002] = t_issue; }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    EncodeTiledFn enc = nullptr;
    EncodeIm2colFn enci = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
    cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", (void**)&enci, cudaEnableDefault, &q);
    // tensor [N=16][H=52][W=52][C=256] bf16 = 22 MB (L2 resident); also viewed as a [43264, 256] matrix
    const int N = 16, H = 52, W = 52, C = 256;
    void* d;
    cudaMalloc(&d, (size_t)N * H * W * C * 2);
    cudaMemset(d, 0, (size_t)N * H * W * C * 2);
    long long* dc;
    cudaMalloc(&dc, 2048 * 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct Case { const char* name; int mode, cbox, rows_or_w, h, depth, mult, variant; };
    const Case cases[] = {
        {"2d 64x128 d1 x1", MODE_2D, 64, 128, 0, 1, 1, 0}, {"2d 64x128 d2 x1", MODE_2D, 64, 128, 0, 2, 1, 0},
        {"2d 64x128 d4 x1", MODE_2D, 64, 128, 0, 4, 1, 0}, {"2d 64x128 d8 x1", MODE_2D, 64, 128, 0, 8, 1, 0},
        {"2d 64x128 d12 x1", MODE_2D, 64, 128, 0, 12, 1, 0},
        {"im2col 64 d1 x1", MODE_IM2COL, 64, 0, 0, 1, 1, 0}, {"im2col 64 d4 x1", MODE_IM2COL, 64, 0, 0, 4, 1, 0},
        {"im2col 64 d8 x1", MODE_IM2COL, 64, 0, 0, 8, 1, 0}, {"im2col 64 d12 x1", MODE_IM2COL, 64, 0, 0, 12, 1, 0},
        {"im2col 64 d4 x3", MODE_IM2COL, 64, 0, 0, 4, 3, 0},
        {"im2col 32 d8 x1", MODE_IM2COL, 32, 0, 0, 8, 1, 0}, {"im2col 32 d4 x3", MODE_IM2COL, 32, 0, 0, 4, 3, 0},
        {"2d 64x256 d6 x1", MODE_2D, 64, 256, 0, 6, 1, 0},
    };
    for (const Case& c : cases) {
        Args a{};
        a.mode = c.mode; a.iters = 400; a.depth = c.depth; a.W = W; a.H = H; a.N = N; a.C = C; a.cbox = c.cbox;
        a.rows = N * H * W; a.variant = c.variant;
        const CUtensorMapSwizzle sw = c.cbox == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
        CUresult r;
        if (c.mode == MODE_2D) {
            const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)a.rows};
            const cuuint64_t st[1] = {(cuuint64_t)C * 2};
            const cuuint32_t box[2] = {(cuuint32_t)c.cbox, (cuuint32_t)c.rows_or_w};
            const cuuint32_t es[2] = {1, 1};
            r = enc(&a.map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            a.bytes = (uint32_t)c.cbox * 2 * c.rows_or_w;
        } else {
            const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
            const cuuint64_t st[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * W, (cuuint64_t)C * 2 * W * H};
            const cuuint32_t es[4] = {1, 1, 1, 1};
            if (c.mode == MODE_IM2COL) {
                const int lo[2] = {-1, -1}, up[2] = {-1, -1};
                r = enci(&a.map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, st, lo, up, (cuuint32_t)c.cbox, 128, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                a.bytes = (uint32_t)c.cbox * 2 * 128;
            } else {
                const cuuint32_t box[4] = {(cuuint32_t)c.cbox, (cuuint32_t)c.rows_or_w, (cuuint32_t)c.h, 1};
                r = enc(&a.map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                a.bytes = (uint32_t)c.cbox * 2 * c.rows_or_w * c.h;
                a.box_w = c.rows_or_w; a.box_h = c.h;
            }
        }
        if (r) { printf("%s: encode failed %d\n", c.name, (int)r); continue; }
        const size_t smem = (size_t)c.depth * ((a.bytes + 1023) & ~1023u) + 1024 + 256;
        for (int rep = 0; rep < 2; ++rep) {
            probe<<<148 * c.mult, 32, smem>>>(a, dc);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: CUDA error %s\n", c.name, cudaGetErrorString(e)); return 1; }
        }
        std::vector<long long> hc(1003);
        cudaMemcpy(hc.data(), dc, 1003 * 8, cudaMemcpyDeviceToHost);
        if (c.variant != 2) printf("   per-iteration clk (block 0): wait %.0f  arm %.0f  tma-issue %.0f\n", hc[1000] / (double)a.iters, hc[1001] / (double)a.iters, hc[1002] / (double)a.iters);
        double avg = 0;
        for (int b = 0; b < 148 * c.mult && b < 1000; ++b) avg += (double)hc[b] / (148 * c.mult);
        printf("%-30s depth %d  %7u B/load  %8.1f clk/load/cta  %6.1f B/clk/SM\n", c.name, c.depth, a.bytes, avg / a.iters,
               c.mult * a.bytes * (double)a.iters / avg);
    }
    return 0;
}
