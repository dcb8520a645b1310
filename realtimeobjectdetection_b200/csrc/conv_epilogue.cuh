// conv_epilogue.cuh -- epilogue shared by the tcgen05 convolution kernels (conv_tc.cu, conv_pair.cu):
// accumulator (TMEM) -> + folded bias -> leaky 0.1 -> + shortcut operand -> fp16 / bf16 (fp32 for head logits)
// -> swizzled shared memory -> TMA store (src/darknet.py:292-295 BN/LeakyReLU, :263-268 shortcut).
//
// Every epilogue warp is its own pipeline -- no CTA-level barrier: warp w may only read TMEM lanes
// [32*(w%4), +32), i.e. 32 rows of the 128-row tile; it owns `stage_bufs` staging slices of 32 rows x
// 128 B, stores each finished slice with its own TMA store (box {ecols, 32 rows}; rows >= M and
// channels >= Cout are clipped by the descriptor) and, for shortcut layers, TMA-loads the operand of
// its NEXT chunk into the slice it is about to reuse (the sum is formed in place).  With 8 epilogue
// warps the two warps that share a TMEM lane quarter split the tile's column chunks (even / odd).
#pragma once
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

namespace rtod {

constexpr uint32_t kEpiSlice = 4096;            // one warp's staging slice: 32 rows x 128 B

// TMEM column of output channel n (relative to the tile) inside an accumulator buffer.  Plain tiles: n.
// Concatenated two-term weights on a CTA pair (each CTA supplies BN/2 hi rows followed by BN/2 lo rows of B):
// the accumulator holds [hi 0..BN/2 | lo 0..BN/2 | hi BN/2..BN | lo BN/2..BN].
static __device__ __forceinline__ uint32_t acc_column(const ConvTcParams& p, int n) {
    if (!p.w_cat || p.lo_col == p.BN) return (uint32_t)n;
    const int half = p.BN >> 1;
    return (uint32_t)(n < half ? n : n + half);
}

// ew: epilogue warp index (0 .. kEpiWarps-1; warps ew and ew+4 share a lane quarter)
// origin(tile, m0, n0, row): first row / first channel of the tile in the output matrix [M, Cout]; in row mode
//   (p.row_mode: a tile is a segment of ONE image row) m0 is the segment's first column and `row` the image row
//   b * Ho + y: output and shortcut operand are then addressed as {Cout, W, B*Ho}, columns beyond W are clipped
// release(buf): arrive on the accumulator-empty barrier the MMA issuer waits on (called by one lane)
// The TMA / bulk-group / mbarrier-arrive instructions of a warp are always issued by its elect.sync lane
// (deterministic for a full mask), so the per-thread bulk async-groups stay with one thread.
// kAhead: registers to spare (one CTA per SM): the folded bias of a 32-column group is loaded BEFORE the accumulator
// read is awaited and both terms of a two-term accumulator are read with one wait -- the epilogue of a thin tile is a
// single warp's dependent chain (measured ~1100 cycles per 32 columns), every round trip taken out of it counts
// kMode: -1 = every layer flag read from `p` at run time (rare combinations, bring-up switches); otherwise the flags
// are compile-time constants (bit 0 leaky, bit 1 shortcut operand, bit 2 two-term accumulator, bit 3 fp32 head logits)
// and the loop body is straight-line code: the generic body spends most of its issue slots on uniform branches and
// instruction-cache misses (ncu source view of a 1x1 layer: stall_branch_resolving / stall_no_inst dominate)
template <int kEpiWarps, bool kF16, bool kAhead, int kMode, class Origin, class Release>
static __device__ __forceinline__ void conv_epilogue_body(const ConvTcParams& p, uint32_t tmem_base, uint64_t* acc_full,
                                                     uint8_t* epi_stage, uint64_t* res_bar, int ew, int lane,
                                                     int tile_first, int tile_step, Origin origin, Release release) {
    constexpr int kColGroups = kEpiWarps / 4;
    const bool f_leaky = kMode < 0 ? p.leaky != 0 : (kMode & 1) != 0;
    const bool f_res = kMode < 0 ? p.has_res != 0 : (kMode & 2) != 0;
    const bool f_cat = kMode < 0 ? p.w_cat != 0 : (kMode & 4) != 0;
    const bool f_fp32 = kMode < 0 ? p.out_fp32 != 0 : (kMode & 8) != 0;
    const int f_dbg = kMode < 0 ? p.dbg : 0;
    const int quarter = (int)(threadIdx.x >> 5) & 3;     // TMEM lanes [32*quarter, +32) = rows of the tile
    const int cg = ew >> 2;                              // this warp's chunks: cg, cg + kColGroups, ...
    const int ecols = p.ecols;
    const uint32_t erow = (uint32_t)ecols * (f_fp32 ? 4u : 2u);
    const int n_chunks = p.BN / ecols;
    const uint32_t sbufs = (uint32_t)p.stage_bufs;
    uint8_t* my_stage = epi_stage + (size_t)ew * sbufs * kEpiSlice;
    uint64_t* my_res = res_bar + ew * 2;
    const int halves = f_fp32 ? 1 : (ecols + 31) / 32;              // 32 accumulator columns each

    auto issue_res = [&](int tile, int c, uint32_t sb) {                 // elected lane only
        int m0, n0, row;
        origin(tile, m0, n0, row);
        mbar_expect_tx(&my_res[sb], 32u * erow);
        if (p.row_mode) tma_load_3d(my_stage + sb * kEpiSlice, &p.tmRes, &my_res[sb], n0 + c * ecols, m0 + quarter * 32, row);
        else tma_load_2d(my_stage + sb * kEpiSlice, &p.tmRes, &my_res[sb], n0 + c * ecols, m0 + quarter * 32);
    };
    if (f_res && cg < n_chunks && tile_first < p.total_tiles && elect_one()) issue_res(tile_first, cg, 0u);

    // NOTE: no early exit: after a time-out (*err_flag != 0) every wait returns at once, the loop
    // drains quickly and the host sees the flag.
    uint32_t g = 0;                                      // chunks this warp has handled
    int local = 0;
    TRACE_DECL(t_acc);
    TRACE_DECL(t_slice);
    TRACE_DECL(t_tmem);
    TRACE_DECL(t_math);
    TRACE_DECL(t_store);
    TRACE_T0(t_begin);
    for (int tile = tile_first; tile < p.total_tiles; tile += tile_step, ++local) {
        const int buf = local & 1;
        int m0, n0, row;
        origin(tile, m0, n0, row);
        TRACE_T0(ta);
        mbar_wait(&acc_full[buf], (uint32_t)(local >> 1) & 1u, p.err_flag);
        TRACE_ADD(t_acc, ta);
        tc_fence_after();
        if (cg >= n_chunks) {                            // tile narrower than the column groups: nothing to do
            if (elect_one()) release(buf);
            continue;
        }
        const uint32_t tmem_acc = tmem_base + (uint32_t)(buf * p.acc_cols) + ((uint32_t)(quarter * 32) << 16);
        for (int c = cg; c < n_chunks; c += kColGroups, ++g) {
            const uint32_t sb = sbufs == 2 ? (g & 1u) : 0u;
            uint8_t* slice = my_stage + sb * kEpiSlice;
            TRACE_T0(ts);
            if (!f_res) {                            // the store that last read this slice has drained
                if (elect_one()) {
                    if (sbufs == 2) bulk_wait_read_1();
                    else bulk_wait_read_0();
                }
                __syncwarp();
            }
            TRACE_ADD(t_slice, ts);
            for (int h = 0; h < halves; ++h) {
                uint32_t v[32];
                TRACE_T0(tt);
                const int nbase = n0 + c * ecols + h * 32;
                float4 bq[kAhead ? 8 : 1];
                if (kAhead) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) bq[q] = __ldg(reinterpret_cast<const float4*>(p.bias + nbase) + q);
                }
                if (!(f_dbg & 32)) {
                    const uint32_t col = f_cat ? acc_column(p, c * ecols + h * 32) : (uint32_t)(c * ecols + h * 32);
                    if (kAhead && f_cat) {                 // hi + lo term, one wait
                        uint32_t w[32];
                        tmem_ld_32x32_issue(tmem_acc + col, v);
                        tmem_ld_32x32_issue(tmem_acc + col + (uint32_t)p.lo_col, w);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
                    } else {
                        tmem_ld_32x32(tmem_acc + col, v);
                        if (f_cat) tmem_ld_add_32x32(tmem_acc + col + (uint32_t)p.lo_col, v);      // hi + lo term
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = (uint32_t)(j + lane);
                }
                if (c + kColGroups >= n_chunks && h == halves - 1) {
                    // this warp's last TMEM read of the tile: hand the accumulator back to the MMA issuer
                    tc_fence_before();
                    __syncwarp();
                    if (elect_one()) release(buf);
                }
                TRACE_ADD(t_tmem, tt);
                TRACE_T0(tm);
                if (f_res && h == 0)                 // shortcut operand of this chunk has landed in `slice`
                    mbar_wait(&my_res[sb], (sbufs == 2 ? (g >> 1) : g) & 1u, p.err_flag);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 b0 = kAhead ? bq[kAhead ? 2 * q : 0] : __ldg(reinterpret_cast<const float4*>(p.bias + nbase + q * 8));
                    const float4 b1 = kAhead ? bq[kAhead ? 2 * q + 1 : 0] : __ldg(reinterpret_cast<const float4*>(p.bias + nbase + q * 8 + 4));
                    float f[8];
                    f[0] = __uint_as_float(v[q * 8 + 0]) + b0.x;
                    f[1] = __uint_as_float(v[q * 8 + 1]) + b0.y;
                    f[2] = __uint_as_float(v[q * 8 + 2]) + b0.z;
                    f[3] = __uint_as_float(v[q * 8 + 3]) + b0.w;
                    f[4] = __uint_as_float(v[q * 8 + 4]) + b1.x;
                    f[5] = __uint_as_float(v[q * 8 + 5]) + b1.y;
                    f[6] = __uint_as_float(v[q * 8 + 6]) + b1.z;
                    f[7] = __uint_as_float(v[q * 8 + 7]) + b1.w;
                    if (f_leaky) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) f[j] = leaky01(f[j]);
                    }
                    if (f_fp32) {                    // 8 fp32 = two 16-byte chunks
                        *reinterpret_cast<float4*>(slice + staged_offset(lane, q * 2, erow)) =
                            make_float4(f[0], f[1], f[2], f[3]);
                        *reinterpret_cast<float4*>(slice + staged_offset(lane, q * 2 + 1, erow)) =
                            make_float4(f[4], f[5], f[6], f[7]);
                    } else {
                        const uint32_t off = staged_offset(lane, h * 4 + q, erow);
                        if (f_res) {
                            const uint4 r = *reinterpret_cast<const uint4*>(slice + off);
                            f[0] += h2_lo<kF16>(r.x); f[1] += h2_hi<kF16>(r.x);
                            f[2] += h2_lo<kF16>(r.y); f[3] += h2_hi<kF16>(r.y);
                            f[4] += h2_lo<kF16>(r.z); f[5] += h2_hi<kF16>(r.z);
                            f[6] += h2_lo<kF16>(r.w); f[7] += h2_hi<kF16>(r.w);
                        }
                        uint4 o;
                        o.x = pack_h2<kF16>(f[0], f[1]);
                        o.y = pack_h2<kF16>(f[2], f[3]);
                        o.z = pack_h2<kF16>(f[4], f[5]);
                        o.w = pack_h2<kF16>(f[6], f[7]);
                        *reinterpret_cast<uint4*>(slice + off) = o;
                    }
                }
                TRACE_ADD(t_math, tm);
            }
            TRACE_T0(tst);
            fence_async_smem();                          // generic-proxy writes -> visible to the TMA
            __syncwarp();
            if (elect_one()) {
                if (p.row_mode) tma_store_3d(&p.tmOut, slice, n0 + c * ecols, m0 + quarter * 32, row);
                else if (!(f_dbg & 16)) tma_store_2d(&p.tmOut, slice, n0 + c * ecols, m0 + quarter * 32);
                bulk_commit();
                if (f_res) {                         // shortcut operand of this warp's next chunk
                    int nt = tile, nc = c + kColGroups;
                    if (nc >= n_chunks) {
                        nt += tile_step;
                        nc = cg;
                    }
                    if (nt < p.total_tiles) {
                        if (sbufs == 2) bulk_wait_read_1();      // the store that last read the other slice
                        else bulk_wait_read_0();
                        issue_res(nt, nc, sbufs == 2 ? ((g + 1u) & 1u) : 0u);
                    }
                }
            }
            TRACE_ADD(t_store, tst);
        }
    }
#ifdef RTOD_TC_TRACE
    if ((p.dbg & 64) && blockIdx.x == 0 && ew == 0 && lane == 0)
        printf("  epilogue warp 0: total %lld clk for %d tiles (%u chunks): acc_full wait %lld, slice wait %lld, tmem ld %lld, "
               "math+res+st.shared %lld, fence+store+res issue %lld\n", clock64() - t_begin, local, g, t_acc, t_slice, t_tmem, t_math, t_store);
#endif
    // shared memory must stay valid until the last stores have READ it; their global writes are complete and
    // visible at kernel completion, which is what the next layer (stream order / griddepcontrol.wait) waits for
    if (elect_one()) bulk_wait_read_0();
    tc_fence_before();
}

// dispatch on the layer flags once per kernel (they are uniform): see conv_epilogue_body
template <int kEpiWarps, bool kF16, bool kAhead, class Origin, class Release>
static __device__ __forceinline__ void conv_epilogue(const ConvTcParams& p, uint32_t tmem_base, uint64_t* acc_full,
                                                     uint8_t* epi_stage, uint64_t* res_bar, int ew, int lane,
                                                     int tile_first, int tile_step, Origin origin, Release release) {
#define RTOD_EPI(mode) conv_epilogue_body<kEpiWarps, kF16, kAhead, mode>(p, tmem_base, acc_full, epi_stage, res_bar, ew, lane, \
                                                                         tile_first, tile_step, origin, release)
    if (p.dbg & (16 | 32)) RTOD_EPI(-1);
    else if (p.out_fp32) {
        if (!p.leaky && !p.has_res && !p.w_cat) RTOD_EPI(8);
        else RTOD_EPI(-1);
    } else if (p.leaky) {
        if (p.w_cat) {
            if (p.has_res) RTOD_EPI(1 | 2 | 4);
            else RTOD_EPI(1 | 4);
        } else {
            if (p.has_res) RTOD_EPI(1 | 2);
            else RTOD_EPI(1);
        }
    } else RTOD_EPI(-1);
#undef RTOD_EPI
}

// ------------------------------------------------------------------------------------------------------
// Split-K variant (small batches: a 13x13x1024 layer at batch 1 has two M tiles and 72 serial k-blocks).
// `split` CTAs share one output tile, each accumulating a slice of K.  Every CTA dumps its raw fp32
// accumulator to scratch[tile][slice][128][BN]; a per-tile counter elects the LAST CTA to arrive, which
// sums the `split` partials in slice order (deterministic, independent of arrival order) and runs the
// normal bias / leaky / shortcut / store path.  work(i, tile, slice) enumerates this CTA's work items.
template <int kEpiWarps, bool kF16, class Work, class Origin, class Release>
static __device__ __forceinline__ void conv_epilogue_split(const ConvTcParams& p, uint32_t tmem_base,
                                                           uint64_t* acc_full, uint8_t* epi_stage, uint64_t* res_bar,
                                                           int* s_last, int ew, int lane, Work work, Origin origin,
                                                           Release release) {
    constexpr int kColGroups = kEpiWarps / 4;
    const int quarter = (int)(threadIdx.x >> 5) & 3;
    const int cg = ew >> 2;
    const int ecols = p.ecols;
    const uint32_t erow = (uint32_t)ecols * (p.out_fp32 ? 4u : 2u);
    const int n_chunks = p.BN / ecols;
    uint8_t* slice = epi_stage + (size_t)ew * p.stage_bufs * kEpiSlice;   // one staging slice is enough here
    uint64_t* my_res = res_bar + ew * 2;
    const int halves = p.out_fp32 ? 1 : (ecols + 31) / 32;
    const int S = p.split_k;
    const int row = quarter * 32 + lane;
    uint32_t res_uses = 0;                               // completed phases of my_res[0]
    int tile, slice_k;
    for (int local = 0; work(local, tile, slice_k); ++local) {
        const int buf = local & 1;
        int m0, n0, image_row;
        origin(tile, m0, n0, image_row);                    // (split-K is never combined with row mode)
        mbar_wait(&acc_full[buf], (uint32_t)(local >> 1) & 1u, p.err_flag);
        tc_fence_after();
        // ---- pass 1: raw accumulator -> scratch ----------------------------------------------------------
        if (cg < n_chunks) {
            const uint32_t tmem_acc = tmem_base + (uint32_t)(buf * p.acc_cols) + ((uint32_t)(quarter * 32) << 16);
            float* part = p.split_scratch + ((size_t)(tile * S + slice_k) * kBM + row) * p.BN;
            for (int c = cg; c < n_chunks; c += kColGroups)
                for (int h = 0; h < halves; ++h) {
                    uint32_t v[32];
                    const uint32_t col = acc_column(p, c * ecols + h * 32);
                    tmem_ld_32x32(tmem_acc + col, v);
                    if (p.w_cat) tmem_ld_add_32x32(tmem_acc + col + (uint32_t)p.lo_col, v);
                    float4* dst = reinterpret_cast<float4*>(part + c * ecols + h * 32);
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        __stcg(dst + q, make_float4(__uint_as_float(v[q * 4]), __uint_as_float(v[q * 4 + 1]),
                                                    __uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3])));
                }
        }
        tc_fence_before();
        __syncwarp();
        if (elect_one()) release(buf);
        // ---- which CTA finishes the tile? ----------------------------------------------------------------
        __threadfence();                                 // partials visible before the counter moves
        asm volatile("bar.sync 1, %0;" ::"r"(32 * kEpiWarps) : "memory");
        if (ew == 0 && lane == 0) {
            const int old = atomicAdd(&p.split_count[tile], 1);
            *s_last = old == S - 1;
            if (old == S - 1) p.split_count[tile] = 0;   // ready for the next forward
        }
        asm volatile("bar.sync 1, %0;" ::"r"(32 * kEpiWarps) : "memory");
        const bool last = *s_last != 0;
        asm volatile("bar.sync 1, %0;" ::"r"(32 * kEpiWarps) : "memory");       // s_last may be rewritten
        if (!last || cg >= n_chunks) continue;
        __threadfence();
        // ---- pass 2 (last arriver): sum the partials in slice order, then the normal epilogue ------------
        const float* part0 = p.split_scratch + ((size_t)(tile * S) * kBM + row) * p.BN;
        for (int c = cg; c < n_chunks; c += kColGroups) {
            if (elect_one()) {
                bulk_wait_read_0();                      // the store that last read `slice` has drained
                if (p.has_res) {
                    mbar_expect_tx(&my_res[0], 32u * erow);
                    tma_load_2d(slice, &p.tmRes, &my_res[0], n0 + c * ecols, m0 + quarter * 32);
                }
            }
            __syncwarp();
            if (p.has_res) {
                mbar_wait(&my_res[0], res_uses & 1u, p.err_flag);
                ++res_uses;
            }
            for (int h = 0; h < halves; ++h) {
                const int nbase = n0 + c * ecols + h * 32;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + nbase + q * 8));
                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + nbase + q * 8 + 4));
                    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    for (int s = 0; s < S; ++s) {
                        const float4* src =
                            reinterpret_cast<const float4*>(part0 + (size_t)s * kBM * p.BN + c * ecols + h * 32 + q * 8);
                        const float4 x0 = __ldcg(src), x1 = __ldcg(src + 1);
                        f[0] += x0.x; f[1] += x0.y; f[2] += x0.z; f[3] += x0.w;
                        f[4] += x1.x; f[5] += x1.y; f[6] += x1.z; f[7] += x1.w;
                    }
                    f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
                    f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
                    if (p.leaky) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) f[j] = leaky01(f[j]);
                    }
                    if (p.out_fp32) {
                        *reinterpret_cast<float4*>(slice + staged_offset(lane, q * 2, erow)) =
                            make_float4(f[0], f[1], f[2], f[3]);
                        *reinterpret_cast<float4*>(slice + staged_offset(lane, q * 2 + 1, erow)) =
                            make_float4(f[4], f[5], f[6], f[7]);
                    } else {
                        const uint32_t off = staged_offset(lane, h * 4 + q, erow);
                        if (p.has_res) {
                            const uint4 r = *reinterpret_cast<const uint4*>(slice + off);
                            f[0] += h2_lo<kF16>(r.x); f[1] += h2_hi<kF16>(r.x);
                            f[2] += h2_lo<kF16>(r.y); f[3] += h2_hi<kF16>(r.y);
                            f[4] += h2_lo<kF16>(r.z); f[5] += h2_hi<kF16>(r.z);
                            f[6] += h2_lo<kF16>(r.w); f[7] += h2_hi<kF16>(r.w);
                        }
                        uint4 o;
                        o.x = pack_h2<kF16>(f[0], f[1]);
                        o.y = pack_h2<kF16>(f[2], f[3]);
                        o.z = pack_h2<kF16>(f[4], f[5]);
                        o.w = pack_h2<kF16>(f[6], f[7]);
                        *reinterpret_cast<uint4*>(slice + off) = o;
                    }
                }
            }
            fence_async_smem();
            __syncwarp();
            if (elect_one()) {
                tma_store_2d(&p.tmOut, slice, n0 + c * ecols, m0 + quarter * 32);
                bulk_commit();
            }
        }
    }
    if (elect_one()) bulk_wait_read_0();
    tc_fence_before();
}

}  // namespace rtod
