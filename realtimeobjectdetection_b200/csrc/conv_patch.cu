// conv_patch.cu -- 3x3 / stride 1 / pad 1 convolution blocks (src/darknet.py:292-295, 467-501; 38 of
// YOLOv3's 75 convolutions, 85 % of its FLOPs) as a tcgen05 implicit GEMM that fetches every
// activation ONCE per output tile instead of once per filter tap.
//
// The im2col kernel (conv_tc.cu) asks the TMA unit for nine 128-pixel gathers per channel slice;
// measured on B200 the im2col gather costs ~5.5 cycles per pixel, which bounds every 3x3 layer.
// Here an output tile is TH x TW pixels of one image.  One TILED 4-D TMA load brings the
// (TH+2) x (TW+2) halo patch of a 64-channel slice into 128B-swizzled shared memory, rows =
// patch pixels in raster order (image borders are zero-filled by the TMA bounds check = the
// convolution's zero padding).  Filter tap (ky, kx) is then simply the same tile read from row
// ky*(TW+2) + kx on: the UMMA shared-memory descriptor may start at any 128-byte row of a swizzled
// tile (the swizzle is a function of the absolute address; verified by tools/umma_shift_probe.cu),
// so nine descriptors with shifted start addresses feed nine groups of tcgen05.mma from one load.
// Accumulator row i corresponds to patch position i (padded width), i.e. output pixel
// (i / PW, i % PW); the two halo columns per row produce rows that are simply never stored
// (81 % of the 128 MMA rows are real outputs for 52x52 / 104x104 / 208x208 maps).
//
// Everything else matches conv_tc.cu: persistent CTAs, double-buffered TMEM accumulators, weights
// resident in shared memory when they fit (else a TMA ring), epilogue through swizzled staging
// tiles + TMA store (4-D box {64 ch, TW, TH, 1}), shortcut operand TMA-loaded two chunks ahead.
#include <cstdlib>

#include "conv_tc.cuh"
#include "tc_ptx.cuh"

namespace rtod {

namespace {

constexpr int kThreads = 224;               // warps: 0 = weights TMA, 1 = MMA, 2-5 = epilogue, 6 = patch TMA
constexpr int kEpilogueWarps = 4;
constexpr int kMaxABufs = 6;
constexpr int kMaxBStages = 8;
constexpr uint32_t kResidentLimit = 100 * 1024;

__global__ void __launch_bounds__(kThreads, 1) conv_patch_kernel(const __grid_constant__ ConvPatchParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t row_bytes = (uint32_t)p.BK * 2u;
    const uint32_t b_bytes = (uint32_t)p.BN * row_bytes;
    const int num_kb = 9 * p.cchunks;
    uint8_t* a_buf = smem;                                               // [a_bufs][a_buf_bytes]
    uint8_t* b_buf = a_buf + (size_t)p.a_bufs * p.a_buf_bytes;           // ring or resident [num_kb]
    uint8_t* out_stage = b_buf + (size_t)(p.b_resident ? num_kb : p.b_stages) * b_bytes;
    uint8_t* res_stage = out_stage + 2 * kStageTile;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(res_stage + (p.has_res ? 2 * kStageTile : 0));
    uint64_t* a_empty = a_full + kMaxABufs;
    uint64_t* b_full = a_empty + kMaxABufs;
    uint64_t* b_empty = b_full + kMaxBStages;
    uint64_t* acc_full = b_empty + kMaxBStages;
    uint64_t* acc_empty = acc_full + 2;
    uint64_t* res_full = acc_empty + 2;
    uint64_t* wres_bar = res_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wres_bar + 1);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&p.tmA);
        prefetch_tmap(&p.tmB);
        prefetch_tmap(&p.tmOut);
        if (p.has_res) prefetch_tmap(&p.tmRes);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kMaxABufs; ++s) {
            mbar_init(&a_full[s], 1);
            mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < kMaxBStages; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], kEpilogueWarps);
            mbar_init(&res_full[b], 1);
        }
        mbar_init(wres_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int my_tiles = (int)blockIdx.x < p.total_tiles
                             ? (p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    // tile -> (image n, first output row y0, first output column x0, first output channel n0)
    auto tile_coords = [&](int local, int& n, int& y0, int& x0, int& n0) {
        const int tile = blockIdx.x + local * gridDim.x;
        const int mt = tile % p.m_tiles;
        n0 = (tile / p.m_tiles) * p.BN;
        x0 = (mt % p.tiles_x) * p.TW;
        y0 = ((mt / p.tiles_x) % p.tiles_y) * p.TH;
        n = mt / (p.tiles_x * p.tiles_y);
    };

    if (warp == 6) {
        // ================= TMA producer A: one halo patch per (tile, channel slice) ====================
        // (its own warp: a wait for a free patch buffer must never hold back the weight stream)
        if (elect_one()) {
            const int Q = my_tiles * p.cchunks;              // global sequence of (tile, channel slice)
            const uint32_t a_tx = (uint32_t)p.PH * p.PW * row_bytes;
            int buf = 0;
            uint32_t aphase = 0;
            int n = 0, y0 = 0, x0 = 0, n0 = 0, cc = p.cchunks, local = -1;
            for (int q = 0; q < Q; ++q, ++cc) {
                if (cc == p.cchunks) {                       // next tile
                    cc = 0;
                    tile_coords(++local, n, y0, x0, n0);
                }
                if (!mbar_wait(&a_empty[buf], aphase ^ 1u, p.err_flag)) break;
                mbar_expect_tx(&a_full[buf], a_tx);
                tma_load_4d(a_buf + (size_t)buf * p.a_buf_bytes, &p.tmA, &a_full[buf], cc * p.BK, x0 - 1, y0 - 1, n);
                if (++buf == p.a_bufs) {
                    buf = 0;
                    aphase ^= 1u;
                }
            }
        }
    } else if (warp == 0) {
        // ================= TMA producer B: weight tiles (ring), or the whole matrix once ===============
        if (elect_one()) {
            if (p.b_resident) {
                mbar_expect_tx(wres_bar, (uint32_t)num_kb * b_bytes);
                for (int kb = 0; kb < num_kb; ++kb)
                    tma_load_2d(b_buf + (size_t)kb * b_bytes, &p.tmB, wres_bar, kb * p.BK, 0);
            } else {
                int bs = 0;
                uint32_t bphase = 0;
                bool ok = true;
                for (int local = 0; ok && local < my_tiles; ++local) {
                    int n, y0, x0, n0;
                    tile_coords(local, n, y0, x0, n0);
                    const int cin = p.cchunks * p.BK;
                    for (int c0 = 0; ok && c0 < cin; c0 += p.BK)
                        for (int tap = 0, k0 = c0; tap < 9; ++tap, k0 += cin) {      // k0 = tap*Cin + c0
                            if (!mbar_wait(&b_empty[bs], bphase ^ 1u, p.err_flag)) { ok = false; break; }
                            mbar_expect_tx(&b_full[bs], b_bytes);
                            tma_load_2d(b_buf + (size_t)bs * b_bytes, &p.tmB, &b_full[bs], k0, n0);
                            if (++bs == p.b_stages) {
                                bs = 0;
                                bphase ^= 1u;
                            }
                        }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (elect_one()) {
            bool ok = true;
            int bs = 0, ab = 0;
            uint32_t bphase = 0, aphase = 0;
            const uint64_t desc_tmpl = smem_desc(0u, row_bytes);
            const uint32_t a_base = smem_u32(a_buf), b_base = smem_u32(b_buf);
            const int ksteps = p.BK / 16;
            const uint32_t row_step = (uint32_t)(p.PW - 3) * row_bytes;     // from tap (ky,2) to tap (ky+1,0)
            if (p.b_resident) ok = mbar_wait(wres_bar, 0u, p.err_flag);
            for (int local = 0; ok && local < my_tiles; ++local) {
                const int buf = local & 1;
                if (!mbar_wait(&acc_empty[buf], ((uint32_t)(local >> 1) & 1u) ^ 1u, p.err_flag)) break;
                tc_fence_after();
                const uint32_t tmem_acc = tmem_base + (uint32_t)(buf * p.BN);
                for (int cc = 0; ok && cc < p.cchunks; ++cc) {
                    if (!mbar_wait(&a_full[ab], aphase, p.err_flag)) { ok = false; break; }
                    // tap (ky, kx) = the patch viewed from row ky*PW + kx on: the A descriptor's start
                    // address walks through the patch by additions only
                    uint32_t a_addr = a_base + (uint32_t)ab * p.a_buf_bytes;
                    uint32_t b_res = b_base + (uint32_t)cc * b_bytes;       // resident: tile (tap*cchunks + cc)
                    for (int tap = 0, kx = 0; tap < 9; ++tap) {
                        uint32_t b_addr;
                        if (p.b_resident) {
                            b_addr = b_res;
                            b_res += (uint32_t)p.cchunks * b_bytes;
                        } else {
                            if (!mbar_wait(&b_full[bs], bphase, p.err_flag)) { ok = false; break; }
                            b_addr = b_base + (uint32_t)bs * b_bytes;
                        }
                        tc_fence_after();
                        uint64_t da = desc_tmpl | (uint64_t)((a_addr & 0x3FFFFu) >> 4);
                        uint64_t db = desc_tmpl | (uint64_t)((b_addr & 0x3FFFFu) >> 4);
                        for (int k = 0; k < ksteps; ++k, da += 2, db += 2)
                            umma_bf16(tmem_acc, da, db, p.idesc, (uint32_t)(cc | tap | k));
                        if (!p.b_resident) {
                            umma_commit(&b_empty[bs]);
                            if (++bs == p.b_stages) {
                                bs = 0;
                                bphase ^= 1u;
                            }
                        }
                        if (++kx == 3) {
                            kx = 0;
                            a_addr += row_bytes + row_step;
                        } else {
                            a_addr += row_bytes;
                        }
                    }
                    umma_commit(&a_empty[ab]);               // patch buffer free once these MMAs retire
                    if (++ab == p.a_bufs) {
                        ab = 0;
                        aphase ^= 1u;
                    }
                }
                umma_commit(&acc_full[buf]);
            }
        }
    } else {
        // ================= epilogue (see conv_tc.cu; rows are patch positions here) ================
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;                 // accumulator row = patch position
        const int oy = row / p.PW, ox = row - oy * p.PW;
        const bool valid = ox < p.TW && oy < p.TH;           // halo columns / rows past the tile: dropped
        const int srow = oy * p.TW + ox;                     // row of the dense TH x TW staging tile
        const bool leader = warp == 2 && lane == 0;
        const int ecols = p.ecols;
        const uint32_t erow = (uint32_t)ecols * 2u;
        const int n_chunks = p.BN / ecols;
        const uint32_t res_tx = (uint32_t)p.TH * p.TW * erow;
        auto issue_res = [&](uint32_t g) {
            const int tl = (int)(g / (uint32_t)n_chunks), c = (int)(g - (uint32_t)tl * n_chunks);
            if (tl >= my_tiles) return;
            int n, y0, x0, n0;
            tile_coords(tl, n, y0, x0, n0);
            mbar_expect_tx(&res_full[g & 1], res_tx);
            tma_load_4d(res_stage + (g & 1) * kStageTile, &p.tmRes, &res_full[g & 1], n0 + c * ecols, x0, y0, n);
        };
        if (p.has_res && leader) {
            issue_res(0);
            issue_res(1);
        }
        // no early exit in this role (named barriers); after a time-out every wait returns at once
        uint32_t g = 0;
        for (int local = 0; local < my_tiles; ++local) {
            const int buf = local & 1;
            int n, y0, x0, n0;
            tile_coords(local, n, y0, x0, n0);
            mbar_wait(&acc_full[buf], (uint32_t)(local >> 1) & 1u, p.err_flag);
            tc_fence_after();
            const uint32_t tmem_acc = tmem_base + (uint32_t)(buf * p.BN) + ((uint32_t)(quarter * 32) << 16);
            for (int c = 0; c < n_chunks; ++c, ++g) {
                uint8_t* ostage = out_stage + (g & 1) * kStageTile;
                const uint8_t* rstage = res_stage + (g & 1) * kStageTile;
                if (leader) bulk_wait_read_1();
                epi_barrier(1);
                if (p.has_res) mbar_wait(&res_full[g & 1], (g >> 1) & 1u, p.err_flag);
                const int halves = (ecols + 31) / 32;
                for (int h = 0; h < halves; ++h) {
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_acc + (uint32_t)(c * ecols + h * 32), v);
                    if (c == n_chunks - 1 && h == halves - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&acc_empty[buf]);
                    }
                    if (!valid) continue;
                    const int nbase = n0 + c * ecols + h * 32;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + nbase + q * 8));
                        const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + nbase + q * 8 + 4));
                        float f[8];
                        f[0] = __uint_as_float(v[q * 8 + 0]) + b0.x;
                        f[1] = __uint_as_float(v[q * 8 + 1]) + b0.y;
                        f[2] = __uint_as_float(v[q * 8 + 2]) + b0.z;
                        f[3] = __uint_as_float(v[q * 8 + 3]) + b0.w;
                        f[4] = __uint_as_float(v[q * 8 + 4]) + b1.x;
                        f[5] = __uint_as_float(v[q * 8 + 5]) + b1.y;
                        f[6] = __uint_as_float(v[q * 8 + 6]) + b1.z;
                        f[7] = __uint_as_float(v[q * 8 + 7]) + b1.w;
                        if (p.leaky) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) f[j] = leaky01(f[j]);
                        }
                        const uint32_t off = staged_offset(srow, h * 4 + q, erow);
                        if (p.has_res) {
                            const uint4 r = *reinterpret_cast<const uint4*>(rstage + off);
                            f[0] += bf16_lo(r.x); f[1] += bf16_hi(r.x);
                            f[2] += bf16_lo(r.y); f[3] += bf16_hi(r.y);
                            f[4] += bf16_lo(r.z); f[5] += bf16_hi(r.z);
                            f[6] += bf16_lo(r.w); f[7] += bf16_hi(r.w);
                        }
                        uint4 o;
                        o.x = pack_bf16x2(f[0], f[1]);
                        o.y = pack_bf16x2(f[2], f[3]);
                        o.z = pack_bf16x2(f[4], f[5]);
                        o.w = pack_bf16x2(f[6], f[7]);
                        *reinterpret_cast<uint4*>(ostage + off) = o;
                    }
                }
                fence_async_smem();
                epi_barrier(2);
                if (leader) {
                    tma_store_4d(&p.tmOut, ostage, n0 + c * ecols, x0, y0, n);   // pixels outside the image are clipped
                    bulk_commit();
                    if (p.has_res) issue_res(g + 2);
                }
            }
        }
        if (leader) bulk_wait_all();
        tc_fence_before();
    }

    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// best TH x TW output tile for an H x W map: (TH-1)*(TW+2) + TW <= 128 accumulator rows
struct PatchTile { int TW, TH; double eff; };

PatchTile pick_tile(int H, int W) {
    PatchTile best{0, 0, 0.0};
    for (int tw = 4; tw <= W && tw <= 126; ++tw) {
        int th = (kBM - tw) / (tw + 2) + 1;
        if (th > H) th = H;
        const long long tiles = (long long)((W + tw - 1) / tw) * ((H + th - 1) / th);
        const double eff = (double)H * W / (double)(tiles * kBM);
        // ties: the squarer tile fetches fewer halo pixels per output
        const bool tie = eff > best.eff - 1e-9 && best.TW > 0 &&
                         (long long)(th + 2) * (tw + 2) * best.TH * best.TW < (long long)(best.TH + 2) * (best.TW + 2) * th * tw;
        if (eff > best.eff + 1e-9 || tie) best = PatchTile{tw, th, eff};
    }
    return best;
}

}  // namespace

bool conv_patch_eligible(const ConvArgs& a) {
    // Opt-in (RTOD_TC_PATCH=1 for resident-weight layers, RTOD_TC_PATCH_ALL=1 for every 3x3/s1 layer):
    // on B200 the weight stream, not the activation gather, bounds these layers, so with one CTA per
    // pair of tiles the 19 % of MMA rows lost to the halo columns outweigh the saved traffic
    // (measured: 129 vs 96 us on the 26x26x512 layers).  It becomes the default once the weight
    // stream is halved by cta_group::2 pairs.
    if (!getenv("RTOD_TC_PATCH") && !getenv("RTOD_TC_PATCH_ALL")) return false;
    if (a.ks != 3 || a.stride != 1 || a.pad != 1 || a.out.fp32) return false;
    if (a.out.H != a.in.H || a.out.W != a.in.W) return false;
    if (pick_tile(a.out.H, a.out.W).eff < 0.70) return false;   // small maps tile badly: im2col kernel
    // until the weight stream is halved (2-CTA pairs) the patch kernel only wins where the weights are
    // resident in shared memory; RTOD_TC_PATCH_ALL=1 forces it for every eligible layer
    const int BN = a.Cout_pad < 256 ? a.Cout_pad : 256;
    const bool resident = a.Cout_pad == BN && (uint32_t)BN * a.K * 2 <= kResidentLimit;
    return resident || getenv("RTOD_TC_PATCH_ALL") != nullptr;
}

int conv_patch_prepare(const ConvArgs& a, int* err_flag, ConvTcLaunch* launch) {
    static EncodeTiledFn encode_tiled = nullptr;
    if (!encode_tiled) {
        const int rc = driver_fn("cuTensorMapEncodeTiled", reinterpret_cast<void**>(&encode_tiled));
        if (rc) return rc;
    }
    ConvPatchParams& p = launch->pp;
    const int BK = pick_bk(a.Cin);
    const PatchTile t = pick_tile(a.out.H, a.out.W);
    p.TW = t.TW; p.TH = t.TH; p.PW = t.TW + 2; p.PH = t.TH + 2;
    p.tiles_x = (a.out.W + t.TW - 1) / t.TW;
    p.tiles_y = (a.out.H + t.TH - 1) / t.TH;
    p.m_tiles = a.B * p.tiles_x * p.tiles_y;
    const uint32_t row_bytes = (uint32_t)BK * 2;
    p.a_buf_bytes = (uint32_t)align_up((size_t)(kBM + 2 * p.PW + 2) * row_bytes, 1024);
    p.has_res = a.res != nullptr;
    p.leaky = a.leaky;
    p.bias = a.bias;
    p.err_flag = err_flag;
    p.BK = BK;
    p.cchunks = a.Cin / BK;

    int BN = a.Cout_pad < 256 ? a.Cout_pad : 256;
    if (a.Cout_pad % BN != 0) BN = 128;
    uint32_t fixed = 0, b_bytes = 0;
    for (;; BN = 128) {                                   // second pass: narrower tile if smem is short
        if (a.Cout_pad % BN != 0 || BN % 32 != 0)
            return fail(RTOD_ERR_UNSUPPORTED, "conv_patch: Cout_pad %d not tileable", a.Cout_pad);
        b_bytes = (uint32_t)BN * row_bytes;
        p.a_bufs = 2;
        fixed = 1024 + p.a_bufs * p.a_buf_bytes + 2 * kStageTile + (p.has_res ? 2 * kStageTile : 0) + 1024;
        const uint32_t w_bytes = (uint32_t)BN * a.K * 2;
        p.b_resident = (a.Cout_pad == BN && w_bytes <= kResidentLimit && fixed + w_bytes <= kSmemLimit &&
                        getenv("RTOD_TC_NO_RESIDENT") == nullptr) ? 1 : 0;
        if (p.b_resident) {
            fixed += w_bytes;
            p.b_stages = 0;
            // small patches: keep more of them in flight (these layers are latency/HBM-bound)
            while (p.a_bufs < kMaxABufs && fixed + p.a_buf_bytes <= kSmemLimit && p.a_bufs * p.a_buf_bytes < 96 * 1024) {
                ++p.a_bufs;
                fixed += p.a_buf_bytes;
            }
            break;
        }
        const int stages = fixed < kSmemLimit ? (int)((kSmemLimit - fixed) / b_bytes) : 0;
        if (stages >= 3 || BN <= 128) {
            if (stages < 2) return fail(RTOD_ERR_UNSUPPORTED, "conv_patch: shared memory budget exceeded");
            p.b_stages = stages > kMaxBStages ? kMaxBStages : stages;
            fixed += (uint32_t)p.b_stages * b_bytes;
            break;
        }
    }
    p.BN = BN;
    p.ecols = BN < 64 ? BN : 64;
    p.total_tiles = p.m_tiles * (a.Cout_pad / BN);
    int cols = 32;
    while (cols < 2 * BN) cols <<= 1;
    p.tmem_cols = cols;
    p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
    launch->patch = 1;
    launch->smem_bytes = fixed;
    launch->grid = dim3((unsigned)(p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs), 1, 1);

    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r;
    {   // input patches
        const cuuint64_t dims[4] = {(cuuint64_t)a.Cin, (cuuint64_t)a.in.W, (cuuint64_t)a.in.H, (cuuint64_t)a.B};
        const cuuint64_t strides[3] = {(cuuint64_t)a.in.pitch * 2, (cuuint64_t)a.in.pitch * 2 * a.in.W,
                                       (cuuint64_t)a.in.pitch * 2 * a.in.W * a.in.H};
        const cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)p.PW, (cuuint32_t)p.PH, 1};
        r = encode_tiled(&p.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a.in.ptr, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(BK), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(RTOD_ERR_CUDA, "cuTensorMapEncodeTiled (patch input) failed: %d", (int)r);
    }
    {   // weights
        const cuuint64_t dims[2] = {(cuuint64_t)a.K, (cuuint64_t)a.Cout_pad};
        const cuuint64_t strides[1] = {(cuuint64_t)a.K * 2};
        const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
        r = encode_tiled(&p.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(a.w), dims, strides,
                         box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(BK), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(RTOD_ERR_CUDA, "cuTensorMapEncodeTiled (patch weights) failed: %d", (int)r);
    }
    {   // output tile / shortcut operand
        const CUtensorMapSwizzle sw = p.ecols * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
        const cuuint64_t dims[4] = {(cuuint64_t)a.Cout, (cuuint64_t)a.out.W, (cuuint64_t)a.out.H, (cuuint64_t)a.B};
        const cuuint64_t strides[3] = {(cuuint64_t)a.out.pitch * 2, (cuuint64_t)a.out.pitch * 2 * a.out.W,
                                       (cuuint64_t)a.out.pitch * 2 * a.out.W * a.out.H};
        const cuuint32_t box[4] = {(cuuint32_t)p.ecols, (cuuint32_t)p.TW, (cuuint32_t)p.TH, 1};
        r = encode_tiled(&p.tmOut, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a.out.ptr, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(RTOD_ERR_CUDA, "cuTensorMapEncodeTiled (patch output) failed: %d", (int)r);
        if (p.has_res) {
            const cuuint64_t rstrides[3] = {(cuuint64_t)a.res_pitch * 2, (cuuint64_t)a.res_pitch * 2 * a.out.W,
                                            (cuuint64_t)a.res_pitch * 2 * a.out.W * a.out.H};
            r = encode_tiled(&p.tmRes, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(a.res), dims,
                             rstrides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(RTOD_ERR_CUDA, "cuTensorMapEncodeTiled (patch shortcut) failed: %d", (int)r);
        }
    }
    RTOD_CUDA_OK(cudaFuncSetAttribute(conv_patch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    return RTOD_OK;
}

int conv_patch_launch(const ConvTcLaunch& launch, cudaStream_t stream) {
    conv_patch_kernel<<<launch.grid, kThreads, launch.smem_bytes, stream>>>(launch.pp);
    RTOD_LAUNCH_OK("conv_patch_kernel");
    return RTOD_OK;
}

}  // namespace rtod
