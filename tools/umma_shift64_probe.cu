// umma_shift_probe.cu -- hardware-semantics probe (bring-up tool, not part of the library):
// can a K-major SWIZZLE_128B UMMA operand start at an arbitrary ROW of a TMA-written tile,
// i.e. at base + r0*128 bytes (not 1024-byte aligned)?  This is what re-using one halo patch
// for the nine taps of a 3x3 convolution needs.  For every r0 and both settings of the
// descriptor's base_offset field the result D = A[r0 : r0+128, :] * B^T is compared with a CPU
// reference.     nvcc -gencode arch=compute_100a,code=sm_100a -o umma_shift_probe umma_shift_probe.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int kRowsA = 320, kK = 32, kN = 64, kM = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tmA,
                                            const __grid_constant__ CUtensorMap tmB, int r0, int use_base_offset,
                                            float* out, int* status) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = raw + (((smem_u32(raw) + 1023u) & ~1023u) - smem_u32(raw));
    uint8_t* sA = smem;                          // 320 rows x 128 B = 40960
    uint8_t* sB = smem + 40960;                  // 64 rows x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 40960 + 8192);
    uint64_t* mma_bar = bar + 1;
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mma_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(20480 + 4096) : "memory");
        // A in two boxes of 160 rows (box rows <= 256)
        for (int h = 0; h < 2; ++h)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                             smem_u32(sA + h * 160 * 64)),
                         "l"(&tmA), "r"(smem_u32(bar)), "r"(0), "r"(h * 160)
                         : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                         smem_u32(sB)),
                     "l"(&tmB), "r"(smem_u32(bar)), "r"(0), "r"(0)
                     : "memory");
        uint32_t done = 0;
        for (int spin = 0; spin < (1 << 22) && !done; ++spin)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(bar)) : "memory");
        if (!done) *status = 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_addr = smem_u32(sA) + (uint32_t)r0 * 64u, b_addr = smem_u32(sB);
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
        for (int k = 0; k < 2; ++k) {
            auto desc = [&](uint32_t addr, bool shifted) {
                uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(512u >> 4) << 32) | (1ull << 46) | (4ull << 61);
                if (shifted && use_base_offset) d |= (uint64_t)((addr >> 7) & 7u) << 49;
                return d;
            };
            const uint64_t da = desc(a_addr + k * 32, true), db = desc(b_addr + k * 32, false);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                         "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)k)
                         : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mma_bar)) : "memory");
    }
    {
        uint32_t done = 0;
        for (int spin = 0; spin < (1 << 22) && !done; ++spin)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(mma_bar)) : "memory");
        if (!done && threadIdx.x == 0) *status = 2;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c * 32));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * kN + c * 32 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    std::vector<__nv_bfloat16> hA(kRowsA * kK), hB(kN * kK);
    std::vector<float> fA(kRowsA * kK), fB(kN * kK);
    srand(1);
    for (size_t i = 0; i < hA.size(); ++i) { hA[i] = __float2bfloat16((rand() % 17 - 8) / 8.0f); fA[i] = __bfloat162float(hA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { hB[i] = __float2bfloat16((rand() % 13 - 6) / 4.0f); fB[i] = __bfloat162float(hB[i]); }
    __nv_bfloat16 *dA, *dB;
    float* dO;
    int* dS;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dO, kM * kN * 4); cudaMalloc(&dS, 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    EncodeTiledFn enc = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
    CUtensorMap tmA, tmB;
    const cuuint32_t es[2] = {1, 1};
    { const cuuint64_t d[2] = {kK, kRowsA}; const cuuint64_t s[1] = {kK * 2}; const cuuint32_t b[2] = {kK, 160};
      CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                       CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); if (r) { printf("encode A %d\n", r); return 1; } }
    { const cuuint64_t d[2] = {kK, kN}; const cuuint64_t s[1] = {kK * 2}; const cuuint32_t b[2] = {kK, kN};
      CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                       CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); if (r) { printf("encode B %d\n", r); return 1; } }
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    std::vector<float> hO(kM * kN);
    const int shifts[] = {0, 8, 1, 2, 3, 5, 7, 13, 54, 55, 109, 110, 111};
    for (int mode = 0; mode < 2; ++mode)
        for (int r0 : shifts) {
            cudaMemset(dS, 0, 4);
            cudaMemset(dO, 0, kM * kN * 4);
            probe<<<1, 128, 64 * 1024>>>(tmA, tmB, r0, mode, dO, dS);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("r0=%d mode=%d CUDA error %s\n", r0, mode, cudaGetErrorString(e)); return 2; }
            int st = 0;
            cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(hO.data(), dO, kM * kN * 4, cudaMemcpyDeviceToHost);
            double worst = 0;
            for (int m = 0; m < kM; ++m)
                for (int n = 0; n < kN; ++n) {
                    double ref = 0;
                    for (int k = 0; k < kK; ++k) ref += (double)fA[(r0 + m) * kK + k] * fB[n * kK + k];
                    worst = fmax(worst, fabs(ref - hO[m * kN + n]));
                }
            printf("r0=%3d base_offset_field=%d status=%d max|err|=%g  %s\n", r0, mode, st, worst, worst < 1e-3 ? "OK" : "MISMATCH");
        }
    return 0;
}
