// nms.cu -- util.write_results / confidence_mask / bbox_iou (src/util.py:106-153, 242-346)
// as three HBM-bound kernels:
//
//   scan   : streams the [B*N, 5+C] prediction tensor through shared memory with 1-D TMA bulk
//            copies (cp.async.bulk + mbarrier, double buffered), applies the strict
//            objectness threshold, finds the first-max class per surviving row with a warp
//            shuffle reduction and appends a 64-bit sort key per candidate
//            (class:12 | ~orderable(objectness):32 | row:20) to a per-image list.
//   image  : one CTA per image: bitonic sort of the keys in shared memory (global-memory
//            path for > 16384 candidates), then warp-per-class greedy suppression in sorted
//            order with bit-exact IoU arithmetic, then an ordered compaction of the kept rows.
//   emit   : prefix over the per-image counts and assembly of the [D, 8] output rows.
//
// All float arithmetic that decides a comparison uses explicitly rounded intrinsics
// (__fadd_rn/__fsub_rn/__fmul_rn/__fdiv_rn): the reference evaluates every tensor op with a
// separate rounding, so fused multiply-add contraction must not happen here.
#include "common.cuh"
#include "iou.cuh"

#include <cstdlib>

namespace rtod {

namespace {

constexpr unsigned long long kDeadKey = ~0ull;
constexpr int kScanThreads = 256;
constexpr int kScanMaxRows = 64;
constexpr int kScanStageBytes = 24576;
constexpr int kScanStages = 3;
constexpr int kImageThreads = 1024;
constexpr int kSortSmemCap = 16384;          // keys sortable in shared memory per image (heavy pass)
constexpr int kLightCap = 4096;              // ... in the light pass

struct Box {
    float x1, y1, x2, y2, area;
};

// confidence_mask (src/util.py:116-117) then centre/size -> corners (src/util.py:263-268)
__device__ __forceinline__ Box load_box(const float* __restrict__ row, float conf) {
    const float cx = row[0], cy = row[1], w = row[2], h = row[3], obj = row[4];
    const float m = obj > conf ? 1.0f : 0.0f;
    const float hw = __fmul_rn(__fmul_rn(w, m), 0.5f), hh = __fmul_rn(__fmul_rn(h, m), 0.5f);
    const float mx = __fmul_rn(cx, m), my = __fmul_rn(cy, m);
    Box b;
    b.x1 = __fsub_rn(mx, hw);
    b.y1 = __fsub_rn(my, hh);
    b.x2 = __fadd_rn(mx, hw);
    b.y2 = __fadd_rn(my, hh);
    b.area = box_area(b.x1, b.y1, b.x2, b.y2);
    return b;
}

__device__ __forceinline__ uint32_t orderable(float f) {      // monotone float -> uint
    const uint32_t u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}

// ---- mbarrier / bulk-copy PTX ----------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                         unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// =============================================================================================
// kernel 1: threshold + class argmax + key append
// =============================================================================================
__global__ void __launch_bounds__(kScanThreads)
nms_scan_kernel(const float* __restrict__ pred, long long total_rows, int N, int L, int C,
                float conf, int P, int rows_per_chunk, int use_bulk, int* __restrict__ cand_count,
                unsigned long long* __restrict__ keys, int* __restrict__ err_flag) {
    extern __shared__ __align__(128) unsigned char scan_smem[];
    __shared__ __align__(8) unsigned long long bars[kScanStages];
    __shared__ unsigned keep_words[2];
    __shared__ int slot_base[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");     // the image pass may be scheduled early
    const int stage_floats = rows_per_chunk * L;
    float* stage_buf[kScanStages];
    for (int s = 0; s < kScanStages; ++s)
        stage_buf[s] = reinterpret_cast<float*>(scan_smem) + (size_t)s * (((stage_floats + 31) / 32) * 32);
    const long long n_chunks = (total_rows + rows_per_chunk - 1) / rows_per_chunk;

    if (tid == 0) {
        for (int s = 0; s < kScanStages; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // issue the load of one chunk into a stage: full aligned chunks go through the TMA bulk
    // engine, the ragged tail (or an unaligned tensor) through plain coalesced loads
    auto chunk_is_bulk = [&](long long c) {
        return use_bulk && (c + 1) * (long long)rows_per_chunk <= total_rows;
    };
    auto prefetch = [&](long long c, int s) {
        const long long first = c * rows_per_chunk;
        if (chunk_is_bulk(c)) {
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&bars[s], (uint32_t)stage_floats * 4u);
                bulk_g2s(stage_buf[s], pred + first * L, (uint32_t)stage_floats * 4u, &bars[s]);
            }
        } else {
            long long rows = total_rows - first;
            if (rows > rows_per_chunk) rows = rows_per_chunk;
            const float* src = pred + first * L;
            for (int i = tid; i < (int)rows * L; i += kScanThreads) stage_buf[s][i] = __ldg(src + i);
        }
    };

    // chunk `it` of this CTA lives in stage it % kScanStages; two chunks are in flight while one is
    // processed.  Two CTA barriers per chunk: (A) the chunk has landed and everybody is done with the
    // previous one, (B) warp 0 has published the keep mask and the reserved slots.
    uint32_t phase_bits = 0u;
    long long c = blockIdx.x;
    for (int pre = 0; pre < kScanStages - 1; ++pre)
        if (c + (long long)pre * gridDim.x < n_chunks) prefetch(c + (long long)pre * gridDim.x, pre);
    const bool grouped = N >= rows_per_chunk;                 // a chunk touches at most two images
    // (image, row) of the chunk's first row advance by carried additions: no 64-bit divide per chunk and thread
    const long long adv = (long long)gridDim.x * rows_per_chunk;
    const long long step_img = adv / N;
    const int step_row = (int)(adv - step_img * N);
    long long img_a = (c * rows_per_chunk) / N;
    int row_a0 = (int)(c * rows_per_chunk - img_a * N);
    bool first_iter = true;
    for (int it = 0; c < n_chunks; ++it, c += gridDim.x) {
        if (!first_iter) {
            img_a += step_img;
            row_a0 += step_row;
            if (row_a0 >= N) {
                row_a0 -= N;
                ++img_a;
            }
        }
        first_iter = false;
        const int s = it % kScanStages;
        if (chunk_is_bulk(c)) {
            if (tid == 0) {                               // one waiter; barrier (A) publishes the data
                const unsigned long long t0 = global_timer_ns();
                while (!mbar_try_wait(&bars[s], (phase_bits >> s) & 1u)) {
                    if (global_timer_ns() - t0 > 2000000000ull) {     // 2 s: report, do not hang
                        atomicExch(err_flag, 1);
                        break;
                    }
                }
            }
            phase_bits ^= 1u << s;
        }
        __syncthreads();                                  // (A)
        const long long ahead = c + (long long)(kScanStages - 1) * gridDim.x;
        if (ahead < n_chunks) prefetch(ahead, (it + kScanStages - 1) % kScanStages);   // stage of chunk it-1
        if (use_bulk & 4) continue;                           // timing experiment only: stream, do not process

        const float* buf = stage_buf[s];
        const long long first = c * rows_per_chunk;
        const int rows = (int)((total_rows - first) < rows_per_chunk ? (total_rows - first) : rows_per_chunk);
        const int split = N - row_a0;                     // rows of image a in this chunk

        // (a) warp 0: objectness threshold (strict '>' in fp32, then "masked objectness != 0") for all
        //     rows (two per lane), and one atomicAdd per image touched to reserve contiguous slots
        if (warp == 0) {
            bool k0 = false, k1 = false;
            if (lane < rows) {
                const float obj = buf[lane * L + 4];
                k0 = __fmul_rn(obj, obj > conf ? 1.0f : 0.0f) != 0.0f;
            }
            if (lane + 32 < rows) {
                const float obj = buf[(lane + 32) * L + 4];
                k1 = __fmul_rn(obj, obj > conf ? 1.0f : 0.0f) != 0.0f;
            }
            const unsigned w0 = __ballot_sync(0xffffffffu, k0), w1 = __ballot_sync(0xffffffffu, k1);
            if (lane == 0) {
                keep_words[0] = w0;
                keep_words[1] = w1;
                if (grouped) {
                    const unsigned long long km = (unsigned long long)w0 | ((unsigned long long)w1 << 32);
                    const unsigned long long lo_mask = split >= 64 ? ~0ull : ((1ull << split) - 1ull);
                    const int cnt_a = __popcll(km & lo_mask), cnt_b = __popcll(km & ~lo_mask);
                    slot_base[0] = cnt_a ? atomicAdd(&cand_count[img_a], cnt_a) : 0;
                    slot_base[1] = cnt_b ? atomicAdd(&cand_count[img_a + 1], cnt_b) : 0;
                }
            }
        }
        __syncthreads();                                  // (B)
        const unsigned long long kmask =
            (unsigned long long)keep_words[0] | ((unsigned long long)keep_words[1] << 32);

        // (b) surviving rows, four at a time per warp (8 lanes each): survivor i belongs to warp (i / 4) % 8,
        //     lane group i % 4.  First-max class over the C scores (ties -> lowest index, NaN as torch.max),
        //     64-bit sort key, append to the image's candidate list
        const int n_keep = __popcll(kmask);
        const int sub = lane >> 3, l8 = lane & 7;
        for (int base = warp * 4; base < n_keep; base += 4 * (kScanThreads / 32)) {
            const int ordinal = base + sub;
            const bool have = ordinal < n_keep;
            int r = 0;
            if (have) {
                const int c0 = __popc(keep_words[0]);
                r = ordinal < c0 ? (int)__fns(keep_words[0], 0, ordinal + 1)
                                 : 32 + (int)__fns(keep_words[1], 0, ordinal - c0 + 1);
            }
            const float* row = buf + r * L;
            const float obj = have ? row[4] : 0.0f;
            const float m = obj > conf ? 1.0f : 0.0f;
            float best = -INFINITY;
            int best_idx = 0x7fffffff;
            if (have)
                for (int j = l8; j < C; j += 8) {
                    const float v = __fmul_rn(row[5 + j], m);
                    const bool take = (best_idx == 0x7fffffff) || (v > best) || (v != v && best == best);
                    if (take) {
                        best = v;
                        best_idx = j;
                    }
                }
#pragma unroll
            for (int off = 4; off > 0; off >>= 1) {                  // stays inside the 8-lane group
                const float ov = __shfl_xor_sync(0xffffffffu, best, off);
                const int oi = __shfl_xor_sync(0xffffffffu, best_idx, off);
                const bool a_nan = best != best, b_nan = ov != ov;
                bool other;
                if (oi == 0x7fffffff) other = false;
                else if (best_idx == 0x7fffffff) other = true;
                else if (a_nan || b_nan) other = (a_nan && b_nan) ? (oi < best_idx) : b_nan;
                else other = (ov > best) || (ov == best && oi < best_idx);
                if (other) {
                    best = ov;
                    best_idx = oi;
                }
            }
            if (have && l8 == 0) {
                long long img;
                int row_in_img, slot;
                if (grouped) {
                    const bool in_a = r < split;
                    img = in_a ? img_a : img_a + 1;
                    row_in_img = in_a ? row_a0 + r : r - split;
                    slot = in_a ? slot_base[0] + ordinal
                                : slot_base[1] + ordinal - __popcll(kmask & ((1ull << split) - 1ull));
                } else {
                    const long long grow = first + r;
                    img = grow / N;
                    row_in_img = (int)(grow - img * N);
                    slot = atomicAdd(&cand_count[img], 1);
                }
                unsigned long long key = kDeadKey;
                if (C > 0 && best != 0.0f) {                      // src/util.py:305 cls_conf != 0
                    const float mobj = __fmul_rn(obj, m);
                    key = ((unsigned long long)best_idx << 52) |
                          ((unsigned long long)(~orderable(mobj)) << 20) |
                          (unsigned long long)row_in_img;
                }
                keys[img * (long long)P + slot] = key;
            }
        }
    }
}

// =============================================================================================
// kernel 1': the same scan without streaming the tensor.  Only the objectness column decides whether a row
// matters (src/util.py:116-117 multiplies everything else by the mask), so every lane reads the 4 bytes of ONE row
// (one 32-byte sector of the 340-byte row reaches the SM) and only the surviving rows -- ~1 % in a detector's
// output -- are read in full, eight lanes per row.  DRAM traffic falls from B*N*(5+C)*4 bytes to about a fifth
// (sector granularity); the result (candidate counts and keys, in any order: the image pass sorts them) is the same.
// =============================================================================================
constexpr int kSparseThreads = 256;
constexpr int kSparseGroups = 4;                 // 32-row groups per warp iteration: four independent loads per lane

__global__ void __launch_bounds__(kSparseThreads)
nms_scan_sparse_kernel(const float* __restrict__ pred, long long total_rows, int N, int L, int C, float conf, int P,
                       int* __restrict__ cand_count, unsigned long long* __restrict__ keys) {
    const int lane = threadIdx.x & 31, sub = lane >> 3, l8 = lane & 7;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");     // the image pass may be scheduled early
    const long long warps_total = (long long)gridDim.x * (kSparseThreads / 32);
    const long long n_blocks = (total_rows + 32 * kSparseGroups - 1) / (32 * kSparseGroups);
    for (long long blk = (long long)blockIdx.x * (kSparseThreads / 32) + (threadIdx.x >> 5); blk < n_blocks; blk += warps_total) {
        const long long base = blk * (32 * kSparseGroups);
        float obj[kSparseGroups];
#pragma unroll
        for (int g = 0; g < kSparseGroups; ++g) {
            const long long row = base + g * 32 + lane;
            obj[g] = row < total_rows ? __ldg(pred + row * L + 4) : 0.0f;
        }
#pragma unroll
        for (int g = 0; g < kSparseGroups; ++g) {
            const long long r0 = base + g * 32;
            // strict '>' in fp32, then "masked objectness != 0" (src/util.py:116-117, :275)
            const bool keep = r0 + lane < total_rows && __fmul_rn(obj[g], obj[g] > conf ? 1.0f : 0.0f) != 0.0f;
            const unsigned mask = __ballot_sync(0xffffffffu, keep);
            if (mask == 0u) continue;
            const int n_keep = __popc(mask);
            // N >= 32: the group touches at most two images -> one atomicAdd per image reserves contiguous slots
            const bool grouped = N >= 32;
            long long img_a = 0;
            int row_a0 = 0, split = 32, slot_a = 0, slot_b = 0, cnt_a = n_keep;
            if (grouped) {
                img_a = r0 / N;
                row_a0 = (int)(r0 - img_a * N);
                split = N - row_a0;                           // rows of image a in this group (may exceed 32)
                const unsigned lo_mask = split >= 32 ? 0xffffffffu : ((1u << split) - 1u);
                cnt_a = __popc(mask & lo_mask);
                const int cnt_b = n_keep - cnt_a;
                if (lane == 0) {
                    slot_a = cnt_a ? atomicAdd(&cand_count[img_a], cnt_a) : 0;
                    slot_b = cnt_b ? atomicAdd(&cand_count[img_a + 1], cnt_b) : 0;
                }
                slot_a = __shfl_sync(0xffffffffu, slot_a, 0);
                slot_b = __shfl_sync(0xffffffffu, slot_b, 0);
            }
            // surviving rows, four at a time (8 lanes each): first-max class over the C scores (ties -> lowest index,
            // NaN as torch.max), 64-bit sort key, append to the image's candidate list
            for (int o0 = 0; o0 < n_keep; o0 += 4) {
                const int ordinal = o0 + sub;
                const bool have = ordinal < n_keep;
                const int r = have ? (int)__fns(mask, 0, ordinal + 1) : 0;
                const float o = __shfl_sync(0xffffffffu, obj[g], r);
                const float* row = pred + (r0 + r) * L;
                const float m = o > conf ? 1.0f : 0.0f;
                float best = -INFINITY;
                int best_idx = 0x7fffffff;
                if (have)
                    for (int j0 = l8; j0 < C; j0 += 32) {      // four independent loads, then the ordered compares
                        float v[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) v[u] = j0 + 8 * u < C ? __ldg(row + 5 + j0 + 8 * u) : 0.0f;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int j = j0 + 8 * u;
                            if (j >= C) break;
                            const float x = __fmul_rn(v[u], m);
                            const bool take = (best_idx == 0x7fffffff) || (x > best) || (x != x && best == best);
                            if (take) {
                                best = x;
                                best_idx = j;
                            }
                        }
                    }
#pragma unroll
                for (int off = 4; off > 0; off >>= 1) {                  // stays inside the 8-lane group
                    const float ov = __shfl_xor_sync(0xffffffffu, best, off);
                    const int oi = __shfl_xor_sync(0xffffffffu, best_idx, off);
                    const bool a_nan = best != best, b_nan = ov != ov;
                    bool other;
                    if (oi == 0x7fffffff) other = false;
                    else if (best_idx == 0x7fffffff) other = true;
                    else if (a_nan || b_nan) other = (a_nan && b_nan) ? (oi < best_idx) : b_nan;
                    else other = (ov > best) || (ov == best && oi < best_idx);
                    if (other) {
                        best = ov;
                        best_idx = oi;
                    }
                }
                if (have && l8 == 0) {
                    long long img;
                    int row_in_img, slot;
                    if (grouped) {
                        const bool in_a = r < split;
                        img = in_a ? img_a : img_a + 1;
                        row_in_img = in_a ? row_a0 + r : r - split;
                        slot = in_a ? slot_a + ordinal : slot_b + ordinal - cnt_a;
                    } else {
                        const long long grow = r0 + r;
                        img = grow / N;
                        row_in_img = (int)(grow - img * N);
                        slot = atomicAdd(&cand_count[img], 1);
                    }
                    unsigned long long key = kDeadKey;
                    if (C > 0 && best != 0.0f) {                      // src/util.py:305 cls_conf != 0
                        const float mobj = __fmul_rn(o, m);
                        key = ((unsigned long long)best_idx << 52) |
                              ((unsigned long long)(~orderable(mobj)) << 20) |
                              (unsigned long long)row_in_img;
                    }
                    keys[img * (long long)P + slot] = key;
                }
            }
        }
    }
}

// =============================================================================================
// kernel 2: per-image sort + per-class greedy suppression + ordered compaction
// =============================================================================================
__device__ __forceinline__ int upper_bound_class(const unsigned long long* keys, int lo, int hi,
                                                 unsigned long long cls) {
    // first position in [lo, hi) whose class field is > cls
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((keys[mid] >> 52) <= cls) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// Launched twice per call: a light configuration (512 threads, 4096 keys in shared memory, four CTAs per
// SM) that takes the images with n_lo < n <= n_hi candidates -- the normal case -- and a heavy one
// (1024 threads, 16384 keys, one CTA per SM; > 16384 candidates sort in global memory) for dense
// images; CTAs whose image is out of their range exit at once.
__global__ void __launch_bounds__(kImageThreads)
nms_image_kernel(const float* __restrict__ pred, int N, int L, float conf, float nms_thr, int P,
                 const int* __restrict__ cand_count, unsigned long long* __restrict__ keys_g,
                 uint32_t* __restrict__ klist_g, uint32_t* __restrict__ kbits_g,
                 uint32_t* __restrict__ kept_pair, int* __restrict__ kept_count, int smem_cap, int n_lo,
                 int n_hi, int box_cap, int B, int* __restrict__ kept_flag, float* __restrict__ out_rows, int out_cap,
                 int* __restrict__ out_count, int* __restrict__ err_flag) {
    // kept_flag != null ("resident" launch: one CTA per image, all of them co-resident): the emit step is fused --
    // every CTA publishes its kept count (+1) in kept_flag[img], sums its predecessors' and writes its own rows
    extern __shared__ __align__(16) unsigned char image_smem[];
    __shared__ int s_live, s_cursor, s_total, s_offset;
    __shared__ int s_part[kImageThreads / 32];
    __shared__ int s_warp_sums[kImageThreads / 32];

    const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthreads = blockDim.x;
    // programmatic dependent launch: this grid may start while the scan (or the light pass) drains
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    int n = cand_count[img];
    if (n > N) n = N;
    if (n <= 0 && !kept_flag) {
        if (n_lo < 0 && tid == 0) kept_count[img] = 0;     // the light pass owns the empty images
        return;
    }
    if (!kept_flag && (n <= n_lo || n > n_hi)) return;       // the other pass handles this image
    if (n < 0) n = 0;
    int np = 32;
    while (np < n) np <<= 1;

    unsigned long long* keys;
    uint32_t *klist, *kbits;
    unsigned long long* src = keys_g + (long long)img * P;
    if (np <= smem_cap) {
        keys = reinterpret_cast<unsigned long long*>(image_smem);
        klist = reinterpret_cast<uint32_t*>(image_smem + (size_t)smem_cap * 8);
        kbits = reinterpret_cast<uint32_t*>(image_smem + (size_t)smem_cap * 12);
        for (int i = tid; i < np; i += nthreads) keys[i] = i < n ? src[i] : kDeadKey;
    } else {
        keys = src;
        klist = klist_g + (long long)img * P;
        kbits = kbits_g + (long long)img * (P / 32);
        for (int i = n + tid; i < np; i += nthreads) keys[i] = kDeadKey;
    }
    for (int i = tid; i < np / 32; i += nthreads) kbits[i] = 0u;
    if (tid == 0) {
        s_live = 0;
        s_cursor = 0;
    }
    __syncthreads();

    // ---- bitonic sort, ascending: class up, objectness down, row index up --------------------
    for (int k = 2; k <= np; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (np >> 1); t += nthreads) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const unsigned long long a = keys[i], b = keys[l];
                const bool up = (i & k) == 0;
                if ((a > b) == up) {
                    keys[i] = b;
                    keys[l] = a;
                }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < np; i += nthreads)
        if (keys[i] != kDeadKey && (i + 1 == np || keys[i + 1] == kDeadKey)) s_live = i + 1;
    __syncthreads();
    const int n_live = s_live;
    const float* img_pred = pred + (long long)img * N * L;
    // the candidates' corner boxes, gathered ONCE into shared memory (all lanes in parallel) when they fit: the
    // suppression loops below then never wait on global memory
    float* sbox = reinterpret_cast<float*>(image_smem + (size_t)smem_cap * 12 + (size_t)smem_cap / 8);
    const bool staged = np <= smem_cap && n_live <= box_cap;
    if (staged) {
        for (int i = tid; i < n_live; i += nthreads) {
            const Box b = load_box(img_pred + (long long)(keys[i] & 0xFFFFFull) * L, conf);
            sbox[i] = b.x1;
            sbox[box_cap + i] = b.y1;
            sbox[2 * box_cap + i] = b.x2;
            sbox[3 * box_cap + i] = b.y2;
            sbox[4 * box_cap + i] = b.area;
        }
        __syncthreads();
    }
    auto box_at = [&](int pos) {
        if (staged) {
            Box b;
            b.x1 = sbox[pos]; b.y1 = sbox[box_cap + pos]; b.x2 = sbox[2 * box_cap + pos]; b.y2 = sbox[3 * box_cap + pos];
            b.area = sbox[4 * box_cap + pos];
            return b;
        }
        return load_box(img_pred + (long long)(keys[pos] & 0xFFFFFull) * L, conf);
    };

    // ---- greedy suppression: class c belongs to warp c % nwarps.  Each lane finds the segment of one of
    //      its warp's classes by binary search in the sorted keys (no claiming, no contention), then the
    //      warp walks through the non-empty segments one after the other.
    const int nwarps = nthreads >> 5;
    const int n_classes = L - 5;
    auto lower_bound_class = [&](unsigned long long cls) {     // first position with class field >= cls
        int lo = 0, hi = n_live;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((keys[mid] >> 52) < cls) lo = mid + 1;
            else hi = mid;
        }
        return lo;
    };
    for (int cbase = warp; cbase < n_classes; cbase += nwarps * 32) {
        const int my_cls = cbase + lane * nwarps;
        int my_lo = 0, my_hi = 0;
        if (my_cls < n_classes) {
            my_lo = lower_bound_class((unsigned long long)my_cls);
            my_hi = lower_bound_class((unsigned long long)my_cls + 1ull);
        }
        unsigned nonempty = __ballot_sync(0xffffffffu, my_hi > my_lo);
      while (nonempty) {
        const int src_lane = __ffs(nonempty) - 1;
        nonempty &= nonempty - 1;
        const int seg_lo = __shfl_sync(0xffffffffu, my_lo, src_lane);
        const int seg_hi = __shfl_sync(0xffffffffu, my_hi, src_lane);

        int kept_in_seg = 0;
        for (int t0 = seg_lo; t0 < seg_hi; t0 += 32) {
            const int p = t0 + lane;
            const bool valid = p < seg_hi;
            Box mine = {0.f, 0.f, 0.f, 0.f, 0.f};
            if (valid) mine = box_at(p);
            bool alive = valid;

            // boxes kept in earlier tiles of this class suppress first
            for (int base = 0; base < kept_in_seg; base += 32) {
                Box kb = {0.f, 0.f, 0.f, 0.f, 0.f};
                if (base + lane < kept_in_seg) {
                    kb = box_at((int)klist[seg_lo + base + lane]);
                }
                const int cnt = min(32, kept_in_seg - base);
                for (int j = 0; j < cnt; ++j) {
                    const float bx1 = __shfl_sync(0xffffffffu, kb.x1, j);
                    const float by1 = __shfl_sync(0xffffffffu, kb.y1, j);
                    const float bx2 = __shfl_sync(0xffffffffu, kb.x2, j);
                    const float by2 = __shfl_sync(0xffffffffu, kb.y2, j);
                    const float bar = __shfl_sync(0xffffffffu, kb.area, j);
                    if (alive) {
                        const float v = iou_exact(bx1, by1, bx2, by2, bar, mine.x1, mine.y1,
                                                  mine.x2, mine.y2, mine.area);
                        if (!(v < nms_thr)) alive = false;
                    }
                }
            }
            // then the tile resolves itself in score order
            unsigned todo = __ballot_sync(0xffffffffu, alive);
            while (todo) {
                const int i = __ffs(todo) - 1;
                const float bx1 = __shfl_sync(0xffffffffu, mine.x1, i);
                const float by1 = __shfl_sync(0xffffffffu, mine.y1, i);
                const float bx2 = __shfl_sync(0xffffffffu, mine.x2, i);
                const float by2 = __shfl_sync(0xffffffffu, mine.y2, i);
                const float bar = __shfl_sync(0xffffffffu, mine.area, i);
                if (lane == i) {
                    klist[seg_lo + kept_in_seg] = (uint32_t)p;
                    atomicOr(&kbits[p >> 5], 1u << (p & 31));
                }
                ++kept_in_seg;
                if (alive && lane > i) {
                    const float v = iou_exact(bx1, by1, bx2, by2, bar, mine.x1, mine.y1, mine.x2,
                                              mine.y2, mine.area);
                    if (!(v < nms_thr)) alive = false;
                }
                const unsigned still = __ballot_sync(0xffffffffu, alive);
                todo = still & ~((2u << i) - 1u);
            }
            __syncwarp();
        }
      }
    }
    __syncthreads();

    // ---- ordered compaction of kept positions -> (row | class << 20) -----------------------------
    int running = 0;
    const int n_words = np / 32;
    for (int w0 = 0; w0 < n_words; w0 += nthreads) {
        const int w = w0 + tid;
        const uint32_t word = w < n_words ? kbits[w] : 0u;
        const int cnt = __popc(word);
        int incl = cnt;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += v;
        }
        if (lane == 31) s_warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int v = lane < (nthreads >> 5) ? s_warp_sums[lane] : 0;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, v, off);
                if (lane >= off) v += u;
            }
            s_warp_sums[lane] = v;                                  // inclusive over warps
            if (lane == 31) s_total = v;
        }
        __syncthreads();
        int rank = running + (warp ? s_warp_sums[warp - 1] : 0) + incl - cnt;
        uint32_t bits = word;
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            const unsigned long long key = keys[w * 32 + b];
            kept_pair[(long long)img * N + rank] =
                (uint32_t)(key & 0xFFFFFull) | ((uint32_t)(key >> 52) << 20);
            ++rank;
        }
        running += s_total;
        __syncthreads();
    }
    if (tid == 0) kept_count[img] = running;
    if (!kept_flag) return;

    // ---- fused emit (src/util.py:332-341): publish the count, add up the predecessors', write the rows --------------
    __syncthreads();                                       // kept_pair rows of this image are written
    if (tid == 0) {
        __threadfence();
        atomicExch(&kept_flag[img], running + 1);
    }
    int acc = 0;
    for (int i = tid; i < img; i += nthreads) {
        int v = 0;
        const unsigned long long t0 = global_timer_ns();
        while ((v = *reinterpret_cast<volatile int*>(&kept_flag[i])) == 0) {
            if (global_timer_ns() - t0 > 2000000000ull) {          // 2 s: report, do not hang
                atomicExch(err_flag, 3);
                v = 1;
                break;
            }
        }
        acc += v - 1;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) s_part[warp] = acc;
    __syncthreads();
    if (tid == 0) {
        int total = 0;
        for (int i = 0; i < (nthreads >> 5); ++i) total += s_part[i];
        s_offset = total;
        // a pipeline time-out anywhere in this call turns the count negative: the host raises (util.py)
        if (img == B - 1) *out_count = *reinterpret_cast<volatile int*>(err_flag) ? -1 : total + running;
    }
    __syncthreads();
    const int offset = s_offset;
    for (int k = tid; k < running; k += nthreads) {
        const long long dst = (long long)offset + k;
        if (dst >= out_cap) break;
        const uint32_t pair = kept_pair[(long long)img * N + k];
        const int row = (int)(pair & 0xFFFFFu), cls = (int)(pair >> 20);
        const float* src = img_pred + (long long)row * L;
        const Box b = load_box(src, conf);
        const float m = src[4] > conf ? 1.0f : 0.0f;
        float* o = out_rows + dst * 8;
        o[0] = (float)img;
        o[1] = b.x1;
        o[2] = b.y1;
        o[3] = b.x2;
        o[4] = b.y2;
        o[5] = __fmul_rn(src[4], m);
        o[6] = __fmul_rn(src[5 + cls], m);
        o[7] = (float)cls;
    }
}

// =============================================================================================
// kernel 3: output rows [img, x1, y1, x2, y2, obj, cls_conf, cls]   (src/util.py:332-341)
// =============================================================================================
__global__ void __launch_bounds__(256)
nms_emit_kernel(const float* __restrict__ pred, int B, int N, int L, float conf,
                const uint32_t* __restrict__ kept_pair, const int* __restrict__ kept_count,
                float* __restrict__ out_rows, int cap, int* __restrict__ out_count, const int* __restrict__ err_flag) {
    __shared__ int s_part[8];
    __shared__ int s_offset;
    const int img = blockIdx.x, tid = threadIdx.x;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    int acc = 0;
    for (int i = tid; i < img; i += 256) acc += kept_count[i];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((tid & 31) == 0) s_part[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        int total = 0;
        for (int i = 0; i < 8; ++i) total += s_part[i];
        s_offset = total;
        if (img == B - 1) *out_count = *reinterpret_cast<const volatile int*>(err_flag) ? -1 : total + kept_count[img];
    }
    __syncthreads();
    const int offset = s_offset, mine = kept_count[img];
    const float* img_pred = pred + (long long)img * N * L;
    for (int k = tid; k < mine; k += 256) {
        const long long dst = (long long)offset + k;
        if (dst >= cap) break;
        const uint32_t pair = kept_pair[(long long)img * N + k];
        const int row = (int)(pair & 0xFFFFFu), cls = (int)(pair >> 20);
        const float* src = img_pred + (long long)row * L;
        const Box b = load_box(src, conf);
        const float m = src[4] > conf ? 1.0f : 0.0f;
        float* o = out_rows + dst * 8;
        o[0] = (float)img;
        o[1] = b.x1;
        o[2] = b.y1;
        o[3] = b.x2;
        o[4] = b.y2;
        o[5] = __fmul_rn(src[4], m);
        o[6] = __fmul_rn(src[5 + cls], m);
        o[7] = (float)cls;
    }
}

// ---- standalone helpers ----------------------------------------------------------------------
__global__ void confidence_mask_kernel(const float* __restrict__ pred, long long rows, int attrs,
                                       float conf, float* __restrict__ out) {
    const long long total = rows * attrs;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / attrs;
        const float m = pred[r * attrs + 4] > conf ? 1.0f : 0.0f;
        out[i] = __fmul_rn(pred[i], m);
    }
}

__global__ void bbox_iou_kernel(const float* __restrict__ b1, int n1, int s1,
                                const float* __restrict__ b2, int n2, int s2, float* __restrict__ out,
                                int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* a = b1 + (long long)(n1 == 1 ? 0 : i) * s1;
    const float* b = b2 + (long long)(n2 == 1 ? 0 : i) * s2;
    out[i] = iou_exact(a[0], a[1], a[2], a[3], box_area(a[0], a[1], a[2], a[3]), b[0], b[1], b[2],
                       b[3], box_area(b[0], b[1], b[2], b[3]));
}

int pow2_ceil(int v) {
    int p = 32;
    while (p < v) p <<= 1;
    return p;
}

struct NmsLayout {
    size_t off_cand, off_kept, off_err, off_flag, off_keys, off_pair, off_klist, off_kbits, total;
    int P;
};

NmsLayout nms_layout(int B, int N) {
    NmsLayout l;
    l.P = pow2_ceil(N > 0 ? N : 1);
    size_t o = 0;
    l.off_cand = o;  o = align_up(o + sizeof(int) * (size_t)B, 256);
    l.off_kept = o;  o = align_up(o + sizeof(int) * (size_t)B, 256);
    l.off_err = o;   o = align_up(o + sizeof(int), 256);
    l.off_flag = o;  o = align_up(o + sizeof(int) * (size_t)B, 256);
    l.off_keys = o;  o = align_up(o + 8ull * B * l.P, 256);
    l.off_pair = o;  o = align_up(o + 4ull * B * (size_t)(N > 0 ? N : 1), 256);
    l.off_klist = o; l.off_kbits = o;
    if (l.P > kSortSmemCap) {
        o = align_up(o + 4ull * B * l.P, 256);
        l.off_kbits = o;
        o = align_up(o + 4ull * B * (l.P / 32), 256);
    }
    l.total = o;
    return l;
}

}  // namespace

}  // namespace rtod

using namespace rtod;

extern "C" size_t rtod_write_results_workspace_bytes(int B, int N, int C) {
    (void)C;
    if (B <= 0 || N < 0) return 0;
    return nms_layout(B, N).total;
}

extern "C" int rtod_write_results(const float* pred, int B, int N, int C, float confidence,
                                  float nms_conf, float* out_rows, int cap, int* out_count,
                                  void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!out_count) return fail(RTOD_ERR_BAD_ARG, "rtod_write_results: out_count is null");
    if (B < 0 || N < 0 || C < 0 || cap < 0)
        return fail(RTOD_ERR_BAD_ARG, "rtod_write_results: negative size (B=%d N=%d C=%d cap=%d)", B, N,
                    C, cap);
    if (B == 0 || N == 0) {
        RTOD_CUDA_OK(cudaMemsetAsync(out_count, 0, sizeof(int), stream));
        return RTOD_OK;
    }
    if (!pred || !workspace || (cap > 0 && !out_rows))
        return fail(RTOD_ERR_BAD_ARG, "rtod_write_results: null pointer");
    if (C > 4096 || N > (1 << 20))
        return fail(RTOD_ERR_UNSUPPORTED, "rtod_write_results: C=%d (max 4096) or N=%d (max 2^20)", C, N);
    const NmsLayout lay = nms_layout(B, N);
    if (workspace_bytes < lay.total)
        return fail(RTOD_ERR_CAPACITY, "rtod_write_results: workspace %zu < required %zu",
                    workspace_bytes, lay.total);
    if ((reinterpret_cast<uintptr_t>(workspace) & 255u) != 0)
        return fail(RTOD_ERR_BAD_ARG, "rtod_write_results: workspace must be 256-byte aligned");

    unsigned char* ws = static_cast<unsigned char*>(workspace);
    int* cand_count = reinterpret_cast<int*>(ws + lay.off_cand);
    int* kept_count = reinterpret_cast<int*>(ws + lay.off_kept);
    int* err_flag = reinterpret_cast<int*>(ws + lay.off_err);
    int* kept_flag = reinterpret_cast<int*>(ws + lay.off_flag);
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(ws + lay.off_keys);
    uint32_t* pair = reinterpret_cast<uint32_t*>(ws + lay.off_pair);
    uint32_t* klist = reinterpret_cast<uint32_t*>(ws + lay.off_klist);
    uint32_t* kbits = reinterpret_cast<uint32_t*>(ws + lay.off_kbits);
    const int L = 5 + C;

    // counters (cand, kept, err, flags) are contiguous at the start of the workspace
    RTOD_CUDA_OK(cudaMemsetAsync(ws, 0, lay.off_keys, stream));

    // ---- scan -----------------------------------------------------------------------------
    const long long total_rows = (long long)B * N;
    const bool stream_scan = getenv("RTOD_NMS_STREAM") != nullptr;            // the tensor-streaming scan (comparison runs, tests)
    if (!stream_scan) {
        const long long n_blocks = (total_rows + 32 * kSparseGroups - 1) / (32 * kSparseGroups);
        long long grid = (n_blocks + kSparseThreads / 32 - 1) / (kSparseThreads / 32);
        if (grid > (long long)kNumSMs * 8) grid = (long long)kNumSMs * 8;
        nms_scan_sparse_kernel<<<(unsigned)grid, kSparseThreads, 0, stream>>>(pred, total_rows, N, L, C, confidence, lay.P,
                                                                            cand_count, keys);
        RTOD_LAUNCH_OK("nms_scan_sparse_kernel");
    } else {
        int rows_per_chunk = kScanStageBytes / (L * 4);
        if (rows_per_chunk > kScanMaxRows) rows_per_chunk = kScanMaxRows;
        if (rows_per_chunk >= 4) rows_per_chunk &= ~3;
        if (rows_per_chunk < 1) rows_per_chunk = 1;
        int use_bulk = ((reinterpret_cast<uintptr_t>(pred) & 15u) == 0) &&
                             (((long long)rows_per_chunk * L) % 4 == 0);
        if (const char* e = getenv("RTOD_NMS_DBG")) use_bulk |= atoi(e) << 1;
        const size_t stage_bytes = (size_t)(((rows_per_chunk * L + 31) / 32) * 32) * 4;
        const size_t scan_smem = kScanStages * stage_bytes;
        // per device (a process may drive several GPUs) and cheap: set on every call
        RTOD_CUDA_OK(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          kScanStages * (kScanStageBytes + 128)));
        const long long n_chunks = (total_rows + rows_per_chunk - 1) / rows_per_chunk;
        int per_sm = (int)(200 * 1024 / (scan_smem + 1024));
        if (per_sm > 8) per_sm = 8;
        if (per_sm < 1) per_sm = 1;
        long long grid = (long long)kNumSMs * per_sm;
        if (grid > n_chunks) grid = n_chunks;
        nms_scan_kernel<<<(unsigned)grid, kScanThreads, scan_smem, stream>>>(
            pred, total_rows, N, L, C, confidence, lay.P, rows_per_chunk, use_bulk, cand_count, keys,
            err_flag);
        RTOD_LAUNCH_OK("nms_scan_kernel");
    }

    // ---- per-image sort + suppression (+ emit) ---------------------------------------------
    constexpr int kBoxCap = 1024;                            // candidates per image whose boxes are staged in shared memory
    RTOD_CUDA_OK(cudaFuncSetAttribute(nms_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      kSortSmemCap * 12 + kSortSmemCap / 8 + kBoxCap * 20));
    cudaLaunchAttribute pdl[1];
    pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pdl[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)B, 1, 1);
    cfg.stream = stream;
    cfg.attrs = pdl;
    cfg.numAttrs = 1;
    const bool no_fused = getenv("RTOD_NMS_NO_FUSED_EMIT") != nullptr;        // light / heavy pass + emit kernel for every batch size
    if (B <= kNumSMs && !no_fused) {
        // resident launch: one CTA per image and SM, every image size in one configuration, emit fused (the CTAs wait
        // for their predecessors' counts, which needs all of them on the machine at once)
        const int hcap = lay.P < kSortSmemCap ? lay.P : kSortSmemCap;
        const int threads = getenv("RTOD_NMS_THREADS") ? atoi(getenv("RTOD_NMS_THREADS")) : kImageThreads;
        cfg.blockDim = dim3((unsigned)(threads >= 128 && threads <= kImageThreads && threads % 32 == 0 ? threads : kImageThreads), 1, 1);
        cfg.dynamicSmemBytes = (size_t)hcap * 12 + hcap / 8 + kBoxCap * 20;
        RTOD_CUDA_OK(cudaLaunchKernelEx(&cfg, nms_image_kernel, pred, N, L, confidence, nms_conf, lay.P,
                                        (const int*)cand_count, keys, klist, kbits, pair, kept_count, hcap, -1, 0x7fffffff,
                                        (int)kBoxCap, B, kept_flag, out_rows, cap, out_count, err_flag));
        return RTOD_OK;
    }
    {   // light pass: images with at most kLightCap candidates (and the empty ones)
        const int lcap = lay.P < kLightCap ? lay.P : kLightCap;
        cfg.blockDim = dim3(512, 1, 1);
        cfg.dynamicSmemBytes = (size_t)lcap * 12 + lcap / 8 + kBoxCap * 20;
        RTOD_CUDA_OK(cudaLaunchKernelEx(&cfg, nms_image_kernel, pred, N, L, confidence, nms_conf, lay.P,
                                        (const int*)cand_count, keys, klist, kbits, pair, kept_count, lcap, -1, lcap,
                                        (int)kBoxCap, B, (int*)nullptr, (float*)nullptr, 0, (int*)nullptr, err_flag));
        if (lay.P > kLightCap) {   // heavy pass: dense images
            const int hcap = lay.P < kSortSmemCap ? lay.P : kSortSmemCap;
            cfg.blockDim = dim3(kImageThreads, 1, 1);
            cfg.dynamicSmemBytes = (size_t)hcap * 12 + hcap / 8 + kBoxCap * 20;
            RTOD_CUDA_OK(cudaLaunchKernelEx(&cfg, nms_image_kernel, pred, N, L, confidence, nms_conf, lay.P,
                                            (const int*)cand_count, keys, klist, kbits, pair, kept_count, hcap,
                                            (int)kLightCap, 0x7fffffff, (int)kBoxCap, B, (int*)nullptr, (float*)nullptr, 0,
                                            (int*)nullptr, err_flag));
        }
    }

    // ---- emit ---------------------------------------------------------------------------------
    cfg.blockDim = dim3(256, 1, 1);
    cfg.dynamicSmemBytes = 0;
    RTOD_CUDA_OK(cudaLaunchKernelEx(&cfg, nms_emit_kernel, pred, B, N, L, confidence, (const uint32_t*)pair,
                                    (const int*)kept_count, out_rows, cap, out_count, (const int*)err_flag));
    return RTOD_OK;
}

extern "C" int rtod_confidence_mask(const float* pred, long long rows, int attrs, float confidence,
                                    float* out, void* stream_) {
    if (rows < 0 || attrs < 5) return fail(RTOD_ERR_BAD_ARG, "rtod_confidence_mask: bad shape");
    if (rows == 0) return RTOD_OK;
    if (!pred || !out) return fail(RTOD_ERR_BAD_ARG, "rtod_confidence_mask: null pointer");
    const long long total = rows * attrs;
    long long blocks = (total + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    confidence_mask_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(pred, rows, attrs,
                                                                                confidence, out);
    RTOD_LAUNCH_OK("confidence_mask_kernel");
    return RTOD_OK;
}

extern "C" int rtod_bbox_iou(const float* box1, int n1, int stride1, const float* box2, int n2,
                             int stride2, float* out, void* stream_) {
    if (n1 < 0 || n2 < 0 || stride1 < 4 || stride2 < 4)
        return fail(RTOD_ERR_BAD_ARG, "rtod_bbox_iou: bad shape");
    if (n1 == 0 || n2 == 0) return RTOD_OK;
    if (n1 != n2 && n1 != 1 && n2 != 1)
        return fail(RTOD_ERR_BAD_ARG, "rtod_bbox_iou: n1=%d and n2=%d do not broadcast", n1, n2);
    if (!box1 || !box2 || !out) return fail(RTOD_ERR_BAD_ARG, "rtod_bbox_iou: null pointer");
    const int n = n1 > n2 ? n1 : n2;
    bbox_iou_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream_>>>(box1, n1, stride1, box2, n2,
                                                                         stride2, out, n);
    RTOD_LAUNCH_OK("bbox_iou_kernel");
    return RTOD_OK;
}
