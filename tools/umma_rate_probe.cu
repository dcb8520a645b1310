// umma_rate_probe.cu -- bring-up measurement (not part of the library): issue rate of tcgen05.mma
// kind::f16 on all SMs with operands resident in shared memory (contents irrelevant), for
// cta_group::1 (M=128) and cta_group::2 (M=256 per CTA pair), N in {64,128,256}; optionally with a
// concurrent stream of TMA-free bulk copies disabled.  Prints cycles per MMA instruction.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I../realtimeobjectdetection_b200/csrc -o umma_rate_probe umma_rate_probe.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

#include "tc_ptx.cuh"
using namespace rtod;

struct Args { int pair, N, iters, elect, commit_every; };

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

template <int kPair>
__global__ void __launch_bounds__(128, 1) probe(Args a, long long* cycles, int* err) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = raw + (((smem_u32(raw) + 1023u) & ~1023u) - smem_u32(raw));
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 4);
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = kPair ? cluster_ctarank() : 0;
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_init(&bar[2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (kPair) cluster_sync_all();
    if (warp == 1) {
        if (kPair) tmem_alloc_pair(slot, 512);
        else tmem_alloc(slot, 512);
    }
    tc_fence_before();
    __syncthreads();
    if (kPair) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *slot;
    const bool issuer = a.elect ? (warp == 0 && elect_one()) : threadIdx.x == 0;
    if (issuer && rank == 0) {
        const int M = kPair ? 256 : 128;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint64_t tmpl = smem_desc(0u, 128u);
        const uint32_t base = smem_u32(smem);
        const long long t0 = clock64();
        uint32_t phase = 0;
        for (int it = 0; it < a.iters; ++it) {
            const uint32_t a_addr = base + (uint32_t)(it & 1) * 49152u;
            uint64_t da = tmpl | (uint64_t)((a_addr & 0x3FFFFu) >> 4);
            uint64_t db = tmpl | (uint64_t)(((a_addr + 16384u) & 0x3FFFFu) >> 4);
            for (int k = 0; k < 4; ++k, da += 2, db += 2) {
                if (kPair) umma_bf16_pair(tmem + (uint32_t)((it & 1) * 256), da, db, idesc, 1u);
                else umma_bf16(tmem + (uint32_t)((it & 1) * 256), da, db, idesc, 1u);
            }
            if (a.commit_every && (it % a.commit_every) == a.commit_every - 1) {     // like releasing a smem stage
                if (kPair) umma_commit_pair(&bar[1 + (it & 1)]);
                else umma_commit(&bar[1 + (it & 1)]);
            }
        }
        if (kPair) umma_commit_pair(&bar[0]);
        else umma_commit(&bar[0]);
        const long long t_issue = clock64() - t0;
        mbar_wait(&bar[0], phase, err);
        const long long t_all = clock64() - t0;
        cycles[2 * (blockIdx.x >> (kPair ? 1 : 0))] = t_issue;
        cycles[2 * (blockIdx.x >> (kPair ? 1 : 0)) + 1] = t_all;
    } else if (issuer && kPair) {
        mbar_wait(&bar[0], 0, err);          // the multicast commit also arrives here
    }
    __syncthreads();
    if (kPair) cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        if (kPair) tmem_dealloc_pair(tmem, 512);
        else tmem_dealloc(tmem, 512);
    }
}

int main() {
    long long* dc;
    cudaMalloc(&dc, 4096 * 8);
    int* derr;
    cudaMalloc(&derr, 4);
    cudaMemset(derr, 0, 4);
    cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int Ns[] = {256, 128};
    for (int ce = 0; ce < 3; ++ce)
    for (int pair = 0; pair < 2; ++pair)
        for (int N : Ns) {
            const int el = 0;
            Args a{pair, N, 2000, el, ce};
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(148, 1, 1);
            cfg.blockDim = dim3(128, 1, 1);
            cfg.dynamicSmemBytes = 98 * 1024;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = pair ? 2 : 1;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = pair ? 1 : 0;
            for (int rep = 0; rep < 2; ++rep) {
                cudaError_t le = pair ? cudaLaunchKernelEx(&cfg, probe<1>, a, dc, derr) : cudaLaunchKernelEx(&cfg, probe<0>, a, dc, derr);
                if (le != cudaSuccess) { printf("launch error %s\n", cudaGetErrorString(le)); return 1; }
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
            }
            const int n = pair ? 74 : 148;
            std::vector<long long> h(2 * n);
            cudaMemcpy(h.data(), dc, 2 * n * 8, cudaMemcpyDeviceToHost);
            double issue = 0, all = 0;
            for (int i = 0; i < n; ++i) { issue += h[2 * i]; all += h[2 * i + 1]; }
            const double mmas = 2000.0 * 4;
            printf("commit every %d k-blocks, cta_group::%d M=%d N=%3d K=16: %.1f clk/MMA to completion (%.1f clk/MMA issue), %.0f MAC/clk/SM\n", ce, pair + 1,
                   pair ? 256 : 128, N, all / n / mmas, issue / n / mmas, 128.0 * N * 16 / (all / n / mmas));
        }
    return 0;
}
