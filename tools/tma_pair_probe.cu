// tma_pair_probe.cu -- bring-up measurement: how fast can ONE producer thread (or two) feed
// {im2col A box 64ch x 128px, tiled B box 64 x BN rows} pairs, consumer = a thread that frees
// the stage at once.  Lean loops (no divisions, no clock reads inside).
//   nvcc -gencode arch=compute_100a,code=sm_100a -I../realtimeobjectdetection_b200/csrc -o tma_pair_probe tma_pair_probe.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>
#include "tc_ptx.cuh"
namespace rtod { char* last_error_buffer() { static char b[512]; return b; } }
using namespace rtod;

struct Args { CUtensorMap tmA, tmB; int iters, stages, bn_rows, producers, use_b, stream, nimg; int* err; };

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ Args a, long long* cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = raw + (((smem_u32(raw) + 1023u) & ~1023u) - smem_u32(raw));
    const uint32_t stage_bytes = 16384u + (a.use_b ? (uint32_t)a.bn_rows * 128u : 0u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)a.stages * stage_bytes);
    uint64_t* empty = full + a.stages;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < a.stages; ++s) { mbar_init(&full[s], a.producers == 2 && a.use_b ? 2 : 1); mbar_init(&empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long t0 = clock64();
    if (a.producers == 3 && warp < 2) {
        if (elect_one()) {                       // two threads alternate k-blocks: is the im2col issue cost per thread?
            int ow = -1 + 7 * warp, oh = -1, on = 0, k0 = 64 * warp;
            for (int it = warp; it < a.iters; it += 2) {
                const int stage = it % a.stages;
                const uint32_t phase = (uint32_t)(it / a.stages) & 1u;
                if (!mbar_wait(&empty[stage], phase ^ 1u, a.err)) break;
                uint8_t* dst = smem + (size_t)stage * stage_bytes;
                mbar_expect_tx(&full[stage], stage_bytes);
                tma_load_im2col_4d(dst, &a.tmA, &full[stage], (it & 3) * 64, ow, oh, on, (uint16_t)(it % 3), (uint16_t)1);
                if (a.use_b) tma_load_2d(dst + 16384, &a.tmB, &full[stage], k0, 0);
                k0 = (k0 + 128) & 2047;
                ow += 14; if (ow > 40) { ow = -1 + 7 * warp; oh += 1; if (oh > 40) { oh = -1; on = (on + 1) & 15; } }
            }
        }
    } else if (warp == 0 || (warp == 1 && a.producers == 2)) {
        if (elect_one()) {
            const bool doA = warp == 0, doB = a.use_b && (a.producers == 1 || warp == 1);
            int stage = 0; uint32_t phase = 0;
            int ow = -1, oh = -1, on = a.stream ? (int)((blockIdx.x * 3u) % (uint32_t)a.nimg) : 0, k0 = 0;
            int cc = 0;
            for (int it = 0; it < a.iters; ++it) {
                if (!mbar_wait(&empty[stage], phase ^ 1u, a.err)) break;
                uint8_t* dst = smem + (size_t)stage * stage_bytes;
                const uint32_t bytes = (doA ? 16384u : 0u) + (doB ? (uint32_t)a.bn_rows * 128u : 0u);
                mbar_expect_tx(&full[stage], bytes);
                if (doA) tma_load_im2col_4d(dst, &a.tmA, &full[stage], a.stream ? cc : (it & 3) * 64, ow, oh, on, (uint16_t)(a.stream ? 1 : it % 3), (uint16_t)1);
                if (doB) tma_load_2d(dst + 16384, &a.tmB, &full[stage], k0, 0);
                k0 = (k0 + 64) & 2047;
                if (a.stream) {          // every load touches new memory: 4 channel chunks, then the next 128 pixels
                    cc += 64;
                    if (cc == 256) { cc = 0; ow += 24; if (ow > 40) { ow = -1; oh += 3; if (oh > 45) { oh = -1; on += 1; if (on >= a.nimg) on = 0; } } }
                } else {
                    ow += 7; if (ow > 40) { ow = -1; oh += 1; if (oh > 40) { oh = -1; on = (on + 1) & 15; } }
                }
                if (++stage == a.stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 2) {
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int it = 0; it < a.iters; ++it) {
                if (!mbar_wait(&full[stage], phase, a.err)) break;
                mbar_arrive(&empty[stage]);
                if (++stage == a.stages) { stage = 0; phase ^= 1u; }
            }
            cycles[blockIdx.x] = clock64() - t0;
        }
    }
}

int main() {
    EncodeTiledFn enc = nullptr; EncodeIm2colFn enci = nullptr;
    driver_fn("cuTensorMapEncodeTiled", (void**)&enc); driver_fn("cuTensorMapEncodeIm2col", (void**)&enci);
    const int N = 512, H = 52, W = 52, C = 256;      // 708 MB: streams through L2 when stream = 1
    void *dA, *dB; cudaMalloc(&dA, (size_t)N * H * W * C * 2); cudaMemset(dA, 0, (size_t)N * H * W * C * 2);
    cudaMalloc(&dB, (size_t)256 * 2304 * 2); cudaMemset(dB, 0, (size_t)256 * 2304 * 2);
    long long* dc; cudaMalloc(&dc, 4096 * 8); int* derr; cudaMalloc(&derr, 4); cudaMemset(derr, 0, 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    for (int stream = 0; stream < 1; ++stream)
    for (int use_b = 0; use_b < 2; ++use_b)
    for (int bn : {128, 256})
    for (int producers = 1; producers <= 3; ++producers)
    for (int mult = 1; mult <= 2; ++mult) {
        if (!use_b && (bn != 128 || producers == 2)) continue;
        if (producers == 2) continue;
        Args a{}; a.iters = 2000; a.bn_rows = bn; a.producers = producers; a.use_b = use_b; a.err = derr; a.stream = stream; a.nimg = stream ? N : 16;
        const uint32_t stage_bytes = 16384u + (use_b ? bn * 128u : 0u);
        a.stages = (int)((mult == 1 ? 200u * 1024u : 100u * 1024u) / stage_bytes); if (a.stages > 8) a.stages = 8;
        const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        const cuuint64_t st[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * W, (cuuint64_t)C * 2 * W * H};
        const cuuint32_t es[4] = {1, 1, 1, 1}; const int lo[2] = {-1, -1}, up[2] = {-1, -1};
        CUresult r = enci(&a.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dA, dims, st, lo, up, 64, 128, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const cuuint64_t bd[2] = {2304, 256}; const cuuint64_t bs[1] = {2304 * 2}; const cuuint32_t bb[2] = {64, (cuuint32_t)bn};
        CUresult r2 = enc(&a.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, bd, bs, bb, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r || r2) { printf("encode failed\n"); return 1; }
        const size_t smem = (size_t)a.stages * stage_bytes + 2048;
        for (int rep = 0; rep < 2; ++rep) {
            probe<<<148 * mult, 128, smem>>>(a, dc);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        }
        std::vector<long long> h(148 * mult); cudaMemcpy(h.data(), dc, h.size() * 8, cudaMemcpyDeviceToHost);
        double avg = 0; for (long long v : h) avg += (double)v / h.size();
        printf("%s A im2col 16KB%s, %d producer thread(s), %d CTA/SM, %d stages: %.0f clk per k-block per CTA, %.1f B/clk/SM\n",
               stream ? "[HBM stream]" : "[L2 resident]", use_b ? (bn == 128 ? " + B 16KB" : " + B 32KB") : "", producers, mult, a.stages, avg / a.iters, mult * stage_bytes * a.iters / avg);
    }
    return 0;
}
