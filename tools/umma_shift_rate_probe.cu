// umma_shift_rate_probe.cu -- bring-up measurement (not part of the library): does a tcgen05.mma whose K-major A
// operand starts at a row that is NOT a multiple of 8 (a row-shifted view of a swizzled tile, what re-using one
// staged slab for the three kx taps of a 3x3 convolution needs) run at the full rate?  Operands resident in
// shared memory (contents irrelevant), cta_group::1, M = 128; prints cycles per MMA for 64-byte and 128-byte rows.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I../realtimeobjectdetection_b200/csrc -o umma_shift_rate_probe umma_shift_rate_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

#include "tc_ptx.cuh"
using namespace rtod;

struct Args { int N, iters, row_bytes, shift_rows, b_shift_rows; };

__global__ void __launch_bounds__(128, 1) probe(Args a, long long* cycles, int* err) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = raw + (((smem_u32(raw) + 1023u) & ~1023u) - smem_u32(raw));
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 4);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) tmem_alloc(slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (warp == 0 && elect_one()) {
        const uint32_t idesc = umma_idesc(1, 128, a.N);
        const uint64_t tmpl = smem_desc(0u, (uint32_t)a.row_bytes);
        const uint32_t base = smem_u32(smem);
        const int ksteps = a.row_bytes / 32;
        const long long t0 = clock64();
        for (int it = 0; it < a.iters; ++it) {
            const uint32_t a_addr = base + (uint32_t)(it & 1) * 24576u + (uint32_t)(a.shift_rows * a.row_bytes);
            const uint32_t b_addr = base + 49152u + (uint32_t)(a.b_shift_rows * a.row_bytes);
            uint64_t da = tmpl | (uint64_t)((a_addr & 0x3FFFFu) >> 4);
            uint64_t db = tmpl | (uint64_t)((b_addr & 0x3FFFFu) >> 4);
            for (int k = 0; k < ksteps; ++k, da += 2, db += 2) umma_bf16(tmem + (uint32_t)((it & 1) * 256), da, db, idesc, 1u);
        }
        umma_commit(&bar[0]);
        mbar_wait(&bar[0], 0, err);
        cycles[blockIdx.x] = clock64() - t0;
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

int main() {
    long long* dc;
    cudaMalloc(&dc, 4096 * 8);
    int* derr;
    cudaMalloc(&derr, 64);
    cudaMemset(derr, 0, 64);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int rb : {64, 128})
        for (int N : {64, 128, 256})
            for (int sh : {0, 1, 2}) {
                Args a{N, 2000, rb, sh, 0};
                for (int rep = 0; rep < 2; ++rep) {
                    probe<<<148, 128, 98 * 1024>>>(a, dc, derr);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
                }
                std::vector<long long> h(148);
                cudaMemcpy(h.data(), dc, 148 * 8, cudaMemcpyDeviceToHost);
                double all = 0;
                for (long long v : h) all += v;
                const double mmas = 2000.0 * (rb / 32);
                printf("row_bytes %3d N %3d A shifted by %d rows: %.1f clk/MMA (nominal %d)\n", rb, N, sh, all / 148 / mmas, N / 2);
            }
    return 0;
}
