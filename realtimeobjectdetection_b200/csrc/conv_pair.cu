// conv_pair.cu -- the convolution block as a tcgen05 implicit GEMM on CTA PAIRS (cta_group::2).
//
// Same algorithm as conv_tc.cu (TMA im2col / tiled operand feed, persistent tiles, double-buffered TMEM
// accumulators, swizzled staging + TMA store epilogue, TMA-loaded shortcut operand) for layers with
// Cout % 256 == 0, but two CTAs on neighbouring SMs share one 256(M) x 256(N) tile: each CTA loads its
// own 128 rows of A and only HALF of the weight tile (128 of the 256 N rows); one tcgen05.mma
// .cta_group::2 issued by the leader CTA multiplies both halves (the tensor cores exchange the B
// halves), each CTA keeps the accumulator of its 128 rows in its own TMEM and runs its own epilogue.
// Per SM and k-block that is 16 KB (A) + 16 KB (B) instead of 16 + 32 KB: the weight stream, which
// bounds the single-CTA kernel at ~50 % tensor utilisation, is halved.
//
// Synchronisation: both producers' TMA loads (.cta_group::2) complete on the LEADER's full barrier, which
// the leader's producer arms for the bytes of both CTAs (the peer's bytes cannot reach it a phase early:
// the peer refills a stage only after the leader's commit released it); the leader's tcgen05.commit
// is multicast to the barriers at the same offset in both CTAs (stage empty, accumulator full); the
// peer's epilogue warps arrive remotely on the leader's accumulator-empty barrier.
#include <cstdlib>

#include "conv_epilogue.cuh"
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

namespace rtod {

namespace {

constexpr int kEpilogueWarps = 8;
constexpr int kThreads = 64 + 32 * kEpilogueWarps;

template <bool kF16>
__global__ void __launch_bounds__(kThreads, 1) conv_pair_kernel(const __grid_constant__ ConvTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef RTOD_TC_TRACE
    long long dbg_c0 = 0;
    unsigned long long dbg_t0 = 0;
    if ((p.dbg & 8) && blockIdx.x == 0 && threadIdx.x == 0) {
        dbg_c0 = clock64();
        dbg_t0 = global_timer_ns();
    }
#endif
    const uint32_t rank = cluster_ctarank();             // 0 = leader (issues the MMAs), 1 = peer
    const int tile_first = blockIdx.x >> 1, tile_step = gridDim.x >> 1;
    const uint32_t row_bytes = (uint32_t)p.BK * 2u;
    // half of the weight tile per CTA; two-term weights: the hi half-tile, then the lo half-tile
    const uint32_t a_bytes = kBM * row_bytes, b_half = (uint32_t)(p.BN / 2) * row_bytes, b_bytes = b_half << p.w_split;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    uint8_t* epi_stage = smem + (size_t)p.stages * stage_bytes;          // [kEpilogueWarps][stage_bufs][kEpiSlice]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_stage + (size_t)kEpilogueWarps * p.stage_bufs * kEpiSlice);
    uint64_t* empty_bar = full_bar + p.stages;
    uint64_t* acc_full = empty_bar + p.stages;
    uint64_t* acc_empty = acc_full + 2;
    uint64_t* res_full = acc_empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full + 2 * kEpilogueWarps);
    const int num_kb = p.ks * p.ks * p.cchunks;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&p.tmA);
        prefetch_tmap(&p.tmB);
        prefetch_tmap(&p.tmOut);
        if (p.has_res) prefetch_tmap(&p.tmRes);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);                  // the leader's arrive.expect_tx covers both CTAs' loads
            mbar_init(&empty_bar[s], 1);                 // the leader's multicast commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 2 * kEpilogueWarps);   // epilogue warps of both CTAs
        }
        for (int b = 0; b < 2 * kEpilogueWarps; ++b) mbar_init(&res_full[b], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster_sync_all();                                  // peer barriers exist before any remote arrive
    if (warp == 2) tmem_alloc_pair(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (p.pf_bytes && warp == 2 && elect_one()) {
        // small batches stream every weight from HBM once per forward and the k-loops are latency-bound: pull
        // the NEXT layer's weights into L2 now (they do not depend on the previous layer, so before the wait)
        const unsigned long long per = ((p.pf_bytes + gridDim.x - 1) / gridDim.x + 127ull) & ~127ull;
        const unsigned long long off = per * blockIdx.x;
        if (off < p.pf_bytes) {
            unsigned long long left = p.pf_bytes - off < per ? p.pf_bytes - off : per;
            const char* ptr = static_cast<const char*>(p.pf_ptr) + off;
            while (left) {
                const uint32_t n = left > 32768ull ? 32768u : (uint32_t)left;
                bulk_prefetch_l2(ptr, n);
                ptr += n;
                left -= n;
            }
        }
    }
    pdl_wait();                          // everything above overlapped the previous layer's tail
    pdl_launch_dependents();

    if (warp == 0) {
        // ================= TMA producer (both CTAs) =================
        if (!(p.dbg & 1) && elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            bool ok = true;
            TRACE_DECL(dbg_wait);
            TRACE_T0(dbg_start);
            for (int tile = tile_first; ok && tile < p.total_tiles; tile += tile_step) {
                const int n_tile = (int)fast_div((uint32_t)tile, p.fd_mtiles);
                const int m0 = (2 * (tile - n_tile * p.m_tiles) + (int)rank) * kBM;
                const int nrow0 = n_tile * p.BN + (int)rank * (p.BN / 2);                 // this CTA's half of B
                int ow = 0, oh = 0, on = 0;
                if (p.ks > 1) {
                    const int prow = (int)fast_div((uint32_t)m0, p.fd_wo);       // m0 / Wo
                    on = (int)fast_div((uint32_t)m0, p.fd_howo);                 // m0 / (Ho * Wo)
                    ow = (m0 - prow * p.Wo) * p.stride - p.pad;
                    oh = (prow - on * p.Ho) * p.stride - p.pad;
                }
                int k0 = 0;
                const int cin = p.cchunks * p.BK;
                for (int ky = 0; ok && ky < p.ks; ++ky)
                for (int kx = 0; ok && kx < p.ks; ++kx) {               // (no tap / ks: a divide costs ~150 cycles here)
                    const uint16_t off_w = (uint16_t)kx, off_h = (uint16_t)ky;
                    for (int c0 = 0; c0 < cin; c0 += p.BK, k0 += p.BK) {
                        TRACE_T0(w0);
                        if (!mbar_wait(&empty_bar[stage], phase ^ 1u, p.err_flag)) { ok = false; break; }
                        TRACE_ADD(dbg_wait, w0);
                        uint8_t* a_dst = smem + (size_t)stage * stage_bytes;
                        const uint32_t lead_full = mapa_u32(&full_bar[stage], 0);
                        if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * stage_bytes);   // both CTAs' bytes
                        if (p.ks > 1) tma_load_im2col_4d_pair(a_dst, &p.tmA, lead_full, c0, ow, oh, on, off_w, off_h);
                        else tma_load_2d_pair(a_dst, &p.tmA, lead_full, c0, m0);
                        // (K-block-major packed weights, BK == PK: block k0 / BK, column 0)
                        tma_load_3d_pair(a_dst + a_bytes, &p.tmB, lead_full, 0, nrow0, k0 >> p.pack_shift);
                        if (p.w_split) tma_load_3d_pair(a_dst + a_bytes + b_half, &p.tmB, lead_full, 0, p.cout_pad + nrow0, k0 >> p.pack_shift);
                        if (++stage == p.stages) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
#ifdef RTOD_TC_TRACE
            if ((p.dbg & 8) && blockIdx.x < 2)
                printf("  pair producer cta %d: total %lld clk, waiting for empty %lld\n", blockIdx.x, clock64() - dbg_start, dbg_wait);
#endif
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA only) =================
        if (rank == 0 && elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            bool ok = true;
            int local = 0;
            const uint64_t desc_tmpl = smem_desc(0u, row_bytes);
            const uint32_t ring_base = smem_u32(smem);
            const int ksteps = p.BK / 16;
            TRACE_DECL(dbg_wacc);
            TRACE_DECL(dbg_wfull);
            TRACE_T0(dbg_start);
            for (int tile = tile_first; ok && tile < p.total_tiles; tile += tile_step, ++local) {
                const int buf = local & 1;
                TRACE_T0(w0);
                if (!mbar_wait(&acc_empty[buf], ((uint32_t)(local >> 1) & 1u) ^ 1u, p.err_flag)) break;
                TRACE_ADD(dbg_wacc, w0);
                tc_fence_after();
                const uint32_t tmem_acc = tmem_base + (uint32_t)(buf * p.acc_cols);
                for (int kb = 0; kb < num_kb; ++kb) {
                    TRACE_T0(w1);
                    if (!(p.dbg & 1) && !mbar_wait(&full_bar[stage], phase, p.err_flag)) { ok = false; break; }
                    TRACE_ADD(dbg_wfull, w1);
                    tc_fence_after();
                    const uint32_t a_addr = ring_base + (uint32_t)stage * stage_bytes;
                    uint64_t da = desc_tmpl | (uint64_t)((a_addr & 0x3FFFFu) >> 4);
                    uint64_t db = desc_tmpl | (uint64_t)(((a_addr + a_bytes) & 0x3FFFFu) >> 4);
                    // (two-term weights: each CTA's B tile is [BN/2 hi rows | BN/2 lo rows], one MMA of N = 2*BN)
                    for (int k = 0; k < ((p.dbg & 2) ? 0 : ksteps); ++k, da += 2, db += 2)
                        umma_bf16_pair(tmem_acc, da, db, p.idesc, (uint32_t)(kb | k));
                    umma_commit_pair(&empty_bar[stage]);     // both CTAs' producers may refill the stage
                    if (++stage == p.stages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit_pair(&acc_full[buf]);            // both CTAs' epilogues may read their half
            }
#ifdef RTOD_TC_TRACE
            if ((p.dbg & 8) && blockIdx.x < 2)
                printf("  pair mma cta %d: total %lld clk, waiting for full %lld, for acc_empty %lld\n", blockIdx.x, clock64() - dbg_start, dbg_wfull, dbg_wacc);
#endif
        }
    } else {
        // ================= epilogue (conv_epilogue.cuh): each CTA stores its own 128 rows =================
        conv_epilogue<kEpilogueWarps, kF16, true>(
            p, tmem_base, acc_full, epi_stage, res_full, warp - 2, lane, tile_first, tile_step,
            [&](int tile, int& m0, int& n0, int& row) {
                const int n_tile = (int)fast_div((uint32_t)tile, p.fd_mtiles);
                m0 = (2 * (tile - n_tile * p.m_tiles) + (int)rank) * kBM;
                n0 = n_tile * p.BN;
                row = -1;
            },
            [&](int buf) {                               // the leader's MMA issuer waits for both CTAs' epilogues
                if (rank == 0) mbar_arrive(&acc_empty[buf]);
                else mbar_arrive_cluster(mapa_u32(&acc_empty[buf], 0));
            });
    }

#ifdef RTOD_TC_TRACE
    if ((p.dbg & 8) && blockIdx.x == 0 && threadIdx.x == 0) {
        const long long dc = clock64() - dbg_c0;
        const unsigned long long dt = global_timer_ns() - dbg_t0;
        printf("%s M %d Cout %d ks %d: %lld clk in %llu ns = %.0f MHz\n", "conv_pair", p.M, p.Cout, p.ks, dc, dt, (double)dc * 1e3 / (double)dt);
    }
#endif
    __syncthreads();
    cluster_sync_all();                                  // the peer may still be reading operands via the MMA
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, (uint32_t)p.tmem_cols);
    }
}

}  // namespace

bool conv_pair_eligible(const ConvArgs& a) {
    if (getenv("RTOD_TC_NO_PAIR")) return false;
    if (a.Cout_pad % 128 != 0 || a.out.fp32) return false;
    const int BN = (a.Cout_pad % 256 == 0 && !a.w_split) ? 256 : 128; // N tile of the pair (each CTA loads half of it)
    if (BN == 128 && getenv("RTOD_TC_NO_PAIR128")) return false;
    const long long m_tiles = ((long long)a.B * a.out.H * a.out.W + kBM - 1) / kBM;
    return ((m_tiles + 1) / 2) * (a.Cout_pad / BN) >= 60;           // enough pair tiles to fill 74 SM pairs
}

int conv_pair_prepare(const ConvArgs& a, int* err_flag, ConvTcLaunch* launch) {
    static EncodeTiledFn encode_tiled = nullptr;
    static EncodeIm2colFn encode_im2col = nullptr;
    if (!encode_tiled) {
        int rc = driver_fn("cuTensorMapEncodeTiled", reinterpret_cast<void**>(&encode_tiled));
        if (rc) return rc;
        rc = driver_fn("cuTensorMapEncodeIm2col", reinterpret_cast<void**>(&encode_im2col));
        if (rc) return rc;
    }
    ConvTcParams& p = launch->p;
    const int BK = pick_bk(a.Cin), BN = (a.Cout_pad % 256 == 0 && !a.w_split) ? 256 : 128;   // (two-term weights: N = 2*BN)
    const long long M = (long long)a.B * a.out.H * a.out.W;
    p.bias = a.bias; p.err_flag = err_flag; p.out_fp32 = 0; p.M = (int)M; p.Cout = a.Cout; p.leaky = a.leaky;
    p.ks = a.ks; p.cchunks = a.Cin / BK; p.BK = BK; p.BN = BN;
    p.Ho = a.out.H; p.Wo = a.out.W; p.stride = a.stride; p.pad = a.pad;
    p.w_cat = a.w_split ? 1 : 0;
    p.acc_cols = BN << p.w_cat; p.lo_col = BN / 2; p.subs = 1; p.row_mode = 0;
    p.tmem_cols = 2 * p.acc_cols; p.has_res = a.res != nullptr; p.ecols = 64; p.b_resident = 0; p.stage_bufs = 2;
    p.dbg = (getenv("RTOD_PAIR_MODE") ? atoi(getenv("RTOD_PAIR_MODE")) : 0) | (getenv("RTOD_CLK_DBG") ? 8 : 0);
    p.f16 = a.in.f16; p.w_split = a.w_split; p.cout_pad = a.Cout_pad;
    p.idesc = umma_idesc(p.f16, 256, p.acc_cols);       // M = 256 per pair
    const uint32_t stage_bytes = (uint32_t)(kBM + ((BN / 2) << a.w_split)) * BK * 2;
    p.epi_warps = kEpilogueWarps;
    if (const char* e = getenv("RTOD_TC_SBUFS")) p.stage_bufs = atoi(e) == 1 ? 1 : 2;
    const uint32_t fixed = 1024 + kEpilogueWarps * p.stage_bufs * kEpiSlice + 512;
    int stages = (int)((kSmemLimit - fixed) / stage_bytes);
    if (stages > 8) stages = 8;
    if (stages < 2) return fail(RTOD_ERR_UNSUPPORTED, "conv_pair: shared memory budget exceeded");
    p.stages = stages;
    const int m_tiles = (int)((M + kBM - 1) / kBM);
    p.m_tiles = (m_tiles + 1) / 2;                        // tile PAIRS along M
    p.total_tiles = p.m_tiles * (a.Cout_pad / BN);
    store_fastdiv(p.fd_mtiles, (uint32_t)p.m_tiles);
    store_fastdiv(p.fd_wo, (uint32_t)a.out.W);
    store_fastdiv(p.fd_howo, (uint32_t)(a.out.W * a.out.H));
    launch->patch = 2;
    launch->choice = ConvTcChoice{1, 0, p.stage_bufs, 1, BN, 1, 8, 1};
    p.split_k = 1; p.split_shift = 0; p.split_scratch = nullptr; p.split_count = nullptr;
    launch->smem_bytes = stages * stage_bytes + fixed;
    RTOD_CUDA_OK(cudaFuncSetAttribute(conv_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    RTOD_CUDA_OK(cudaFuncSetAttribute(conv_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    // clusters that can be resident at once (pairs must share a TPC): the persistent tile loop
    // strides by the grid, so a cluster that only starts in a second wave would double the time
    static int max_clusters = 0;                         // same for every B200 of a box: queried once
    if (!max_clusters) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(kNumSMs, 1, 1);
        cfg.blockDim = dim3(kThreads, 1, 1);
        cfg.dynamicSmemBytes = launch->smem_bytes;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        RTOD_CUDA_OK(cudaOccupancyMaxActiveClusters(&max_clusters, conv_pair_kernel<true>, &cfg));
        if (const char* e = getenv("RTOD_PAIR_CLUSTERS")) max_clusters = atoi(e);
        if (getenv("RTOD_PAIR_DBG")) fprintf(stderr, "conv_pair: %d resident clusters\n", max_clusters);
        if (max_clusters < 1) return fail(RTOD_ERR_CUDA, "conv_pair: no resident cluster fits");
    }
    const int pairs = p.total_tiles < max_clusters ? p.total_tiles : max_clusters;
    launch->grid = dim3((unsigned)(2 * pairs), 1, 1);

    const cuuint32_t estr1[4] = {1, 1, 1, 1};
    CUresult r;
    if (a.ks == 1) {
        const cuuint64_t dims[2] = {(cuuint64_t)a.Cin, (cuuint64_t)M};
        const cuuint64_t strides[1] = {(cuuint64_t)a.in.pitch * 2};
        const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)kBM};
        r = encode_tiled(&p.tmA, h16_tmap_type(a.in.f16), 2, a.in.ptr, dims, strides, box, estr1,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(BK), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        const cuuint64_t dims[4] = {(cuuint64_t)a.Cin, (cuuint64_t)a.in.W, (cuuint64_t)a.in.H, (cuuint64_t)a.B};
        const cuuint64_t strides[3] = {(cuuint64_t)a.in.pitch * 2, (cuuint64_t)a.in.pitch * 2 * a.in.W,
                                       (cuuint64_t)a.in.pitch * 2 * a.in.W * a.in.H};
        const int lower[2] = {-a.pad, -a.pad};
        const int upper[2] = {a.pad - (a.ks - 1), a.pad - (a.ks - 1)};
        const cuuint32_t estr[4] = {1, (cuuint32_t)a.stride, (cuuint32_t)a.stride, 1};
        r = encode_im2col(&p.tmA, h16_tmap_type(a.in.f16), 4, a.in.ptr, dims, strides, lower, upper,
                          (cuuint32_t)BK, (cuuint32_t)kBM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(BK),
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS && (unsigned long long)a.in.pitch * 2ull * a.in.W * a.in.H * a.B < 131072ull)
            reinterpret_cast<uint64_t*>(&p.tmA)[1] &= ~(1ull << 21);      // see conv_tc.cu
    }
    if (r != CUDA_SUCCESS) return fail(RTOD_ERR_CUDA, "cuTensorMapEncode (pair activations) failed: %d", (int)r);
    {
        p.pack_k = BK;                                     // (= weight_pack_k(Cin, K): the pair kernel's K tile is pick_bk(Cin))
        p.pack_shift = BK == 64 ? 6 : (BK == 32 ? 5 : 4);
        const cuuint64_t rows = (cuuint64_t)(a.Cout_pad << a.w_split);
        const cuuint64_t dims[3] = {(cuuint64_t)BK, rows, (cuuint64_t)(a.K / BK)};
        const cuuint64_t strides[2] = {(cuuint64_t)BK * 2, rows * BK * 2};
        const cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)(BN / 2), 1};
        r = encode_tiled(&p.tmB, h16_tmap_type(a.in.f16), 3, const_cast<void*>(a.w), dims, strides,
                         box, estr1, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(BK), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(RTOD_ERR_CUDA, "cuTensorMapEncodeTiled (pair weights) failed: %d", (int)r);
    }
    {
        const cuuint64_t dims[2] = {(cuuint64_t)a.Cout, (cuuint64_t)M};
        const cuuint64_t strides[1] = {(cuuint64_t)a.out.pitch * 2};
        const cuuint32_t box[2] = {64, 32};                       // one epilogue warp's rows
        r = encode_tiled(&p.tmOut, h16_tmap_type(a.in.f16), 2, a.out.ptr, dims, strides, box, estr1,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(RTOD_ERR_CUDA, "cuTensorMapEncodeTiled (pair output) failed: %d", (int)r);
        if (p.has_res) {
            const cuuint64_t rstrides[1] = {(cuuint64_t)a.res_pitch * 2};
            r = encode_tiled(&p.tmRes, h16_tmap_type(a.in.f16), 2, const_cast<void*>(a.res), dims,
                             rstrides, box, estr1, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(RTOD_ERR_CUDA, "cuTensorMapEncodeTiled (pair shortcut) failed: %d", (int)r);
        }
    }
    return RTOD_OK;
}

int conv_pair_launch(const ConvTcLaunch& launch, cudaStream_t stream) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = launch.grid;
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = launch.smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    if (launch.p.f16) RTOD_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_pair_kernel<true>, launch.p));
    else RTOD_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_pair_kernel<false>, launch.p));
    return RTOD_OK;
}

}  // namespace rtod
