// tma_multicast_probe.cu -- bring-up measurement (not part of the library): how many bytes per clock does an SM receive
// from L2 when the CTAs of a cluster want the SAME tile -- each issuing its own TMA load (does the L2 merge them?) or one
// CTA multicasting it (cp.async.bulk.tensor ... .multicast::cluster) -- compared with every CTA loading different tiles?
// This decides whether multicasting the weight tile across two CTA pairs would lift the L2 -> SM bound of the 3x3 layers.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_multicast_probe tma_multicast_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}

constexpr int kDepth = 6;
constexpr uint32_t kTileBytes = 16384;            // 128 rows x 128 B
constexpr int kTiles = 512;                       // 8 MB tensor: L2 resident

struct Args { CUtensorMap map; int mode, iters, csz, col_blocks; };

__global__ void __launch_bounds__(64) probe(const __grid_constant__ Args a, long long* cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = raw + (((smem_u32(raw) + 1023u) & ~1023u) - smem_u32(raw));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)kDepth * kTileBytes);
    uint64_t* empty = full + kDepth;
    const uint32_t rank = cluster_rank();
    const bool mc = a.mode == 2;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kDepth; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[s])), "r"(mc ? a.csz : 1));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (a.csz > 1) cluster_sync();
    const uint32_t base = a.mode == 0 ? blockIdx.x * 977u : (a.mode == 3 ? 0u : cluster_id() * 977u);
    const long long t0 = clock64();
    if (threadIdx.x == 0) {                                   // producer
        if (!mc || rank == 0) {
            for (int i = 0; i < a.iters; ++i) {
                const int s = i % kDepth;
                if (i >= kDepth) while (!try_wait(&empty[s], ((i / kDepth) - 1) & 1)) {}
                const int tile = (int)((base + (uint32_t)i * 31u) % kTiles);
                if (!mc) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(kTileBytes) : "memory");
                    // (wide tensor: the 128 rows of a box lie `pitch` bytes apart -- the weight-tile pattern)
                    const int cb = tile % a.col_blocks, rb = tile / a.col_blocks;
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                                     smem_u32(smem + (size_t)s * kTileBytes)), "l"(&a.map), "r"(smem_u32(&full[s])), "r"(cb * 64), "r"(rb * 128) : "memory");
                } else {
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
                                     smem_u32(smem + (size_t)s * kTileBytes)), "l"(&a.map), "r"(smem_u32(&full[s])), "r"(0), "r"(tile * 128),
                                 "h"((uint16_t)((1u << a.csz) - 1u)) : "memory");
                }
            }
        }
    } else if (threadIdx.x == 32) {                            // consumer
        uint32_t remote_empty[kDepth];
        for (int s = 0; s < kDepth; ++s)
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote_empty[s]) : "r"(smem_u32(&empty[s])), "r"(0u));
        for (int i = 0; i < a.iters; ++i) {
            const int s = i % kDepth;
            if (mc)                                            // every destination arms its own barrier for the multicast bytes
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(kTileBytes) : "memory");
            while (!try_wait(&full[s], (i / kDepth) & 1)) {}
            if (mc) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_empty[s]) : "memory");
            else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
        }
        cycles[blockIdx.x] = clock64() - t0;
    }
    __syncthreads();
    if (a.csz > 1) cluster_sync();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    EncodeTiledFn encode = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", reinterpret_cast<void**>(&encode), cudaEnableDefault, &q);
    if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    void* data;
    cudaMalloc(&data, (size_t)kTiles * kTileBytes);
    cudaMemset(data, 1, (size_t)kTiles * kTileBytes);
    long long* dc;
    cudaMalloc(&dc, 4096 * 8);
    Args a{};
    const cuuint32_t box[2] = {64, 128}, estr[2] = {1, 1};
    auto make_map = [&](int col_blocks) {
        // the same 8 MB as [kTiles / col_blocks * 128 rows][col_blocks * 64 columns]: col_blocks = 1 -> every box is one
        // contiguous 16 KB block, col_blocks = 16 -> its 128 rows lie 2048 bytes apart
        const cuuint64_t dims[2] = {(cuuint64_t)col_blocks * 64, (cuuint64_t)(kTiles / col_blocks) * 128};
        const cuuint64_t strides[1] = {(cuuint64_t)col_blocks * 128};
        a.col_blocks = col_blocks;
        return encode(&a.map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, data, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    if (!make_map(1)) { printf("encode failed\n"); return 1; }
    const size_t smem = 1024 + (size_t)kDepth * kTileBytes + 256;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const char* names[4] = {"unicast, every CTA a different tile", "unicast, the CTAs of a cluster the same tile", "multicast from rank 0 to the cluster",
                            "unicast, EVERY CTA of the grid the same tile"};
    for (int pass = 0; pass < 2; ++pass)
    for (int csz : {1, 2, 4})
        for (int mode = 0; mode < 4; ++mode) {
            if (csz == 1 && (mode == 1 || mode == 2)) continue;
            if (pass == 1) {                       // second pass: scattered rows, unicast only
                if (csz != 1 || mode != 0) continue;
                for (int cb : {2, 4, 16}) {
                    if (!make_map(cb)) { printf("encode failed\n"); return 1; }
                    a.mode = 0; a.iters = 4000; a.csz = 1;
                    for (int rep = 0; rep < 2; ++rep) { probe<<<148, 64, smem>>>(a, dc); cudaDeviceSynchronize(); }
                    std::vector<long long> h2(148);
                    cudaMemcpy(h2.data(), dc, 148 * 8, cudaMemcpyDeviceToHost);
                    double m2 = 0;
                    for (long long v : h2) m2 += (double)v / 148;
                    printf("unicast, different tiles, box rows %5d bytes apart: %6.1f B/clk/SM delivered\n", cb * 128, (double)a.iters * kTileBytes / m2);
                }
                make_map(1);
                continue;
            }
            a.mode = mode; a.iters = 4000; a.csz = csz;
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(148, 1, 1);
            cfg.blockDim = dim3(64, 1, 1);
            cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = csz; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = csz > 1 ? 1 : 0;
            for (int rep = 0; rep < 2; ++rep) {
                cudaError_t le = cudaLaunchKernelEx(&cfg, probe, a, dc);
                if (le != cudaSuccess) { printf("launch error %s\n", cudaGetErrorString(le)); return 1; }
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
            }
            std::vector<long long> h(148);
            cudaMemcpy(h.data(), dc, 148 * 8, cudaMemcpyDeviceToHost);
            double mean = 0, worst = 0;
            for (long long v : h) { mean += (double)v / 148; if ((double)v > worst) worst = (double)v; }
            printf("cluster %d, %-48s: %6.1f B/clk/SM delivered (slowest CTA %6.1f)\n", csz, names[mode], (double)a.iters * kTileBytes / mean,
                   (double)a.iters * kTileBytes / worst);
        }
    return 0;
}
