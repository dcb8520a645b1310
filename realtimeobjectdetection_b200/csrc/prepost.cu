// prepost.cu -- the callers either side of the hot path:
//
//   rtod_prep_image       util.prep_image / letterbox_image (src/util.py:349-397): uint8 HWC BGR frames ->
//                         aspect-preserving bicubic resize (cv2.INTER_CUBIC semantics) onto a 128-filled
//                         canvas, BGR->RGB, HWC->CHW, /255 -- one kernel, frames never visit the host as fp32
//   rtod_rescale_boxes    detect.py:120-136: letterbox coordinates -> source-image coordinates + clamp
//   rtod_bbox_iou_matrix  test.py:139-151: [P, T] IoU matrix of predictions against targets
//
// All three are tiny next to the forward pass (HBM-bound, a few MB); what matters is that their arithmetic is
// the reference's: every float step that the reference rounds separately is an explicitly rounded intrinsic.
#include <cstdlib>

#include "common.cuh"
#include "iou.cuh"

namespace rtod {

namespace {

// ---- cv2.resize(..., INTER_CUBIC) on 8-bit images (OpenCV modules/imgproc/src/resize.cpp) -----------------
// Keys cubic kernel, A = -0.75, evaluated in fp32 with one rounding per operation (interpolateCubic)
__device__ __forceinline__ void cubic_coefficients(float x, float (&c)[4]) {
    const float A = -0.75f;
    const float x1 = __fadd_rn(x, 1.0f);
    c[0] = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(A, x1), __fmul_rn(5.0f, A)), x1), __fmul_rn(8.0f, A)), x1),
                     __fmul_rn(4.0f, A));
    c[1] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(A, 2.0f), x), __fadd_rn(A, 3.0f)), x), x), 1.0f);
    const float y = __fsub_rn(1.0f, x);
    c[2] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(A, 2.0f), y), __fadd_rn(A, 3.0f)), y), y), 1.0f);
    c[3] = __fsub_rn(__fsub_rn(__fsub_rn(1.0f, c[0]), c[1]), c[2]);
}

// destination index -> first tap's source index + 1 and the four weights:
// f = float((d + 0.5) * scale - 0.5) (double arithmetic, then one rounding), s = floor(f), weights of f - s
__device__ __forceinline__ int axis_entry(int d, double scale, float (&c)[4]) {
    const float f = (float)(__dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5));
    const float fl = floorf(f);
    cubic_coefficients(__fsub_rn(f, fl), c);
    return (int)fl;
}

struct PrepParams {
    const unsigned char* src;      // [B, src_h, src_w, 3]
    void* out;                     // [B, 3, dim, dim] fp32 or uint8
    int B, src_h, src_w, dim;
    int new_w, new_h, left, top;   // letterbox geometry (src/util.py:360-369)
    double scale_x, scale_y;       // source pixels per destination pixel
    int reverse;                   // 1: output channel c = source channel 2 - c (BGR -> RGB)
    int identity;                  // the frame already has the network's size: the cubic weights are exactly {0, 1, 0, 0}
};

// kFixed: OpenCV's own path (11-bit fixed-point weights, int32 horizontal pass, fp32 vertical pass added
// from tap 3 down to tap 0) -- bit-identical to cv2 built without IPP; otherwise all-float weights
// (the IPP-backed stock wheel, to within one grey level on < 0.03 % of the values).
template <bool kFixed, bool kOutU8>
__global__ void __launch_bounds__(256) prep_image_kernel(const PrepParams p) {
    const long long total = (long long)p.B * p.dim * p.dim;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int x = (int)(i % p.dim), y = (int)((i / p.dim) % p.dim), b = (int)(i / ((long long)p.dim * p.dim));
        int v[3] = {128, 128, 128};                                           // np.full(..., 128)
        const int dx = x - p.left, dy = y - p.top;
        if (p.identity) {
            const unsigned char* px = p.src + (((long long)b * p.src_h + y) * p.src_w + x) * 3;
            v[0] = px[0]; v[1] = px[1]; v[2] = px[2];
        } else if (dx >= 0 && dx < p.new_w && dy >= 0 && dy < p.new_h) {
            float cx[4], cy[4];
            const int sx = axis_entry(dx, p.scale_x, cx), sy = axis_entry(dy, p.scale_y, cy);
            const unsigned char* img = p.src + (long long)b * p.src_h * p.src_w * 3;
            int xs[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) xs[k] = min(max(sx - 1 + k, 0), p.src_w - 1) * 3;       // border: replicate
            if (kFixed) {
                int ia[4];
                float fb[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    ia[k] = max(-32768, min(32767, __float2int_rn(__fmul_rn(cx[k], 2048.0f))));
                    const int ib = max(-32768, min(32767, __float2int_rn(__fmul_rn(cy[k], 2048.0f))));
                    fb[k] = __fmul_rn((float)ib, 1.0f / (2048.0f * 2048.0f));
                }
                float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 3; k >= 0; --k) {
                    const unsigned char* row = img + (long long)min(max(sy - 1 + k, 0), p.src_h - 1) * p.src_w * 3;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const int h = row[xs[0] + c] * ia[0] + row[xs[1] + c] * ia[1] + row[xs[2] + c] * ia[2] + row[xs[3] + c] * ia[3];
                        const float t = __fmul_rn((float)h, fb[k]);
                        acc[c] = k == 3 ? t : __fadd_rn(t, acc[c]);
                    }
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) v[c] = max(0, min(255, __float2int_rn(acc[c])));
            } else {
                float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const unsigned char* row = img + (long long)min(max(sy - 1 + k, 0), p.src_h - 1) * p.src_w * 3;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        float h = __fmul_rn((float)row[xs[0] + c], cx[0]);
                        h = __fadd_rn(h, __fmul_rn((float)row[xs[1] + c], cx[1]));
                        h = __fadd_rn(h, __fmul_rn((float)row[xs[2] + c], cx[2]));
                        h = __fadd_rn(h, __fmul_rn((float)row[xs[3] + c], cx[3]));
                        acc[c] = __fadd_rn(acc[c], __fmul_rn(h, cy[k]));
                    }
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) v[c] = max(0, min(255, __float2int_rn(acc[c])));
            }
        }
        const long long plane = (long long)p.dim * p.dim;
        const long long o = (long long)b * 3 * plane + (long long)y * p.dim + x;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int sv = v[p.reverse ? 2 - c : c];
            if (kOutU8) static_cast<unsigned char*>(p.out)[o + c * plane] = (unsigned char)sv;
            else static_cast<float*>(p.out)[o + c * plane] = __fdiv_rn((float)sv, 255.0f);     // .float().div(255.0)
        }
    }
}

// ---- detect.py:120-136 ------------------------------------------------------------------------------------------
__global__ void rescale_boxes_kernel(const float* __restrict__ rows, int D, const float* __restrict__ im_dims, int n_img,
                                     float inp_dim, float ref_dim, float* __restrict__ out_rows, float* __restrict__ out_dims) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= D) return;
    const float* r = rows + (long long)j * 8;
    long long img = (long long)r[0];                                        // .long(): truncation
    img = img < 0 ? 0 : (img >= n_img ? n_img - 1 : img);                   // (index_select would raise; stay in range)
    const float w = im_dims[img * 4 + 0], h = im_dims[img * 4 + 1];
    // torch.min(ref / dims, 1)[0] over (w, h, w, h); int / Tensor is Tensor.__rtruediv__ = reciprocal() * int
    const float sf = nan_min(__fmul_rn(__frcp_rn(w), ref_dim), __fmul_rn(__frcp_rn(h), ref_dim));
    const float ox = __fdiv_rn(__fsub_rn(inp_dim, __fmul_rn(sf, w)), 2.0f);
    const float oy = __fdiv_rn(__fsub_rn(inp_dim, __fmul_rn(sf, h)), 2.0f);
    float* o = out_rows + (long long)j * 8;
    o[0] = r[0];
    auto clampf = [](float v, float hi) {                                   // torch.clamp(v, 0.0, hi), NaN kept
        v = v < 0.0f ? 0.0f : v;
        return v > hi ? hi : v;
    };
    o[1] = clampf(__fdiv_rn(__fsub_rn(r[1], ox), sf), w);
    o[2] = clampf(__fdiv_rn(__fsub_rn(r[2], oy), sf), h);
    o[3] = clampf(__fdiv_rn(__fsub_rn(r[3], ox), sf), w);
    o[4] = clampf(__fdiv_rn(__fsub_rn(r[4], oy), sf), h);
    o[5] = r[5]; o[6] = r[6]; o[7] = r[7];
    if (out_dims) {
        out_dims[(long long)j * 4 + 0] = w; out_dims[(long long)j * 4 + 1] = h;
        out_dims[(long long)j * 4 + 2] = im_dims[img * 4 + 2]; out_dims[(long long)j * 4 + 3] = im_dims[img * 4 + 3];
    }
}

// ---- fixed-capacity gather payload of one rank (sharding.gather_detections_async) ------------------------------------
// payload row 0 = (count, 0, ...); row 1 + i = detection i with the image column shifted by first_frame for i < count,
// zeros beyond: one launch instead of the eight small tensor operations it replaces on every step of every rank
__global__ void pack_detections_kernel(const float* __restrict__ rows, int n_rows, const int* __restrict__ count, float first_frame,
                                       int capacity, float* __restrict__ payload) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;                     // payload row
    if (i > capacity) return;
    const int cnt = *count;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (i == 0) a.x = (float)cnt;
    else if (i - 1 < cnt && i - 1 < n_rows) {
        const float4* r = reinterpret_cast<const float4*>(rows + (long long)(i - 1) * 8);
        a = r[0];
        b = r[1];
        a.x = __fadd_rn(a.x, first_frame);
    }
    float4* o = reinterpret_cast<float4*>(payload + (long long)i * 8);
    o[0] = a;
    o[1] = b;
}

// ---- test.py:139-151 ---------------------------------------------------------------------------------------------
// one thread per target and eight predictions; the eight prediction boxes of a CTA sit in shared memory
__global__ void __launch_bounds__(256) iou_matrix_kernel(const float* __restrict__ pred, int P, int ps, const float* __restrict__ target,
                                                         int T, int ts, int use_thr, double thr, float* __restrict__ out) {
    __shared__ float sbox[8][5];
    const int p0 = blockIdx.y * 8;
    if (threadIdx.x < 8 && p0 + threadIdx.x < P) {
        const float* a = pred + (long long)(p0 + threadIdx.x) * ps;
        sbox[threadIdx.x][0] = a[0]; sbox[threadIdx.x][1] = a[1]; sbox[threadIdx.x][2] = a[2]; sbox[threadIdx.x][3] = a[3];
        sbox[threadIdx.x][4] = box_area(a[0], a[1], a[2], a[3]);
    }
    __syncthreads();
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= T) return;
    const float* b = target + (long long)t * ts;
    const float bx1 = b[0], by1 = b[1], bx2 = b[2], by2 = b[3], barea = box_area(bx1, by1, bx2, by2);
    for (int k = 0; k < 8 && p0 + k < P; ++k) {
        float v = iou_exact(sbox[k][0], sbox[k][1], sbox[k][2], sbox[k][3], sbox[k][4], bx1, by1, bx2, by2, barea);
        if (use_thr && !((double)v > thr)) v = 0.0f;                       // iou.item() > threshold, as Python floats
        out[(long long)(p0 + k) * T + t] = v;
    }
}

}  // namespace

}  // namespace rtod

using namespace rtod;

extern "C" int rtod_letterbox_geometry(int src_w, int src_h, int inp_dim, int* new_w, int* new_h, int* left, int* top) {
    if (src_w <= 0 || src_h <= 0 || inp_dim <= 0) return fail(RTOD_ERR_BAD_ARG, "rtod_letterbox_geometry: bad size");
    // Python: int(img_w * min(w / img_w, h / img_h)) -- double arithmetic, truncation
    const double a = (double)inp_dim / (double)src_w, b = (double)inp_dim / (double)src_h;
    const double m = a < b ? a : b;
    const int nw = (int)((double)src_w * m), nh = (int)((double)src_h * m);
    if (new_w) *new_w = nw;
    if (new_h) *new_h = nh;
    // Python floor division (the operands are non-negative: the resized image never exceeds the canvas)
    if (left) *left = (inp_dim - nw) / 2;
    if (top) *top = (inp_dim - nh) / 2;
    return RTOD_OK;
}

// Frames that already have the network's size and uint8 planes out (the streaming pipeline's normal case): the resize is
// the identity and the kernel is a pure HWC -> CHW de-interleave.  16 pixels per thread: three 16-byte loads, one 16-byte
// store per plane (the per-pixel kernel above moves single bytes: 0.9 TB/s).
__global__ void __launch_bounds__(256) prep_identity_u8_kernel(const unsigned char* __restrict__ src, unsigned char* __restrict__ out,
                                                              long long groups, int hw16, int reverse) {
    for (long long g = blockIdx.x * 256ll + threadIdx.x; g < groups; g += (long long)gridDim.x * 256) {
        const uint4* in = reinterpret_cast<const uint4*>(src) + g * 3;
        const uint4 w0 = __ldcs(in), w1 = __ldcs(in + 1), w2 = __ldcs(in + 2);
        const uint32_t w[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
        const long long b = g / hw16, q = g - b * hw16;                      // image, 16-pixel group inside it
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {                                     // pixels 4j .. 4j+3 of the group
                uint32_t v = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int byte = 3 * (4 * j + i) + c;
                    v |= ((w[byte >> 2] >> ((byte & 3) * 8)) & 0xFFu) << (8 * i);
                }
                o[j] = v;
            }
            const int plane = reverse ? 2 - c : c;
            reinterpret_cast<uint4*>(out)[(b * 3 + plane) * hw16 + q] = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

extern "C" int rtod_prep_image(const unsigned char* src, int B, int src_h, int src_w, int inp_dim, int keep_order,
                               int resize_mode, int out_u8, void* out, void* stream) {
    if (B < 0 || src_h <= 0 || src_w <= 0 || inp_dim <= 0)
        return fail(RTOD_ERR_BAD_ARG, "rtod_prep_image: bad shape (B=%d %dx%d -> %d)", B, src_h, src_w, inp_dim);
    if (B == 0) return RTOD_OK;
    if (!src || !out) return fail(RTOD_ERR_BAD_ARG, "rtod_prep_image: null pointer");
    if (resize_mode != 0 && resize_mode != 1)
        return fail(RTOD_ERR_BAD_ARG, "rtod_prep_image: resize_mode is 0 (float) or 1 (OpenCV fixed point)");
    PrepParams p{};
    p.src = src; p.out = out; p.B = B; p.src_h = src_h; p.src_w = src_w; p.dim = inp_dim; p.reverse = keep_order ? 0 : 1;
    rtod_letterbox_geometry(src_w, src_h, inp_dim, &p.new_w, &p.new_h, &p.left, &p.top);
    if (p.new_w < 1 || p.new_h < 1)
        return fail(RTOD_ERR_UNSUPPORTED, "rtod_prep_image: %dx%d collapses at %d", src_w, src_h, inp_dim);
    p.identity = (p.new_w == src_w && p.new_h == src_h && src_w == inp_dim && src_h == inp_dim) ? 1 : 0;
    p.scale_x = 1.0 / ((double)p.new_w / (double)src_w);                  // resize.cpp: scale = 1 / (dsize / ssize)
    p.scale_y = 1.0 / ((double)p.new_h / (double)src_h);
    const long long total = (long long)B * inp_dim * inp_dim;
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)kNumSMs * 32) blocks = (long long)kNumSMs * 32;
    cudaStream_t s = (cudaStream_t)stream;
    if (p.identity && out_u8 && ((long long)inp_dim * inp_dim) % 16 == 0 && (reinterpret_cast<uintptr_t>(src) & 15u) == 0 &&
        (reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
        const long long groups = total / 16;
        long long gb = (groups + 255) / 256;
        if (gb > (long long)kNumSMs * 16) gb = (long long)kNumSMs * 16;
        prep_identity_u8_kernel<<<(unsigned)gb, 256, 0, s>>>(src, static_cast<unsigned char*>(out), groups, inp_dim * inp_dim / 16,
                                                             p.reverse);
        RTOD_LAUNCH_OK("prep_identity_u8_kernel");
        return RTOD_OK;
    }
    if (resize_mode == 1) {
        if (out_u8) prep_image_kernel<true, true><<<(unsigned)blocks, 256, 0, s>>>(p);
        else prep_image_kernel<true, false><<<(unsigned)blocks, 256, 0, s>>>(p);
    } else {
        if (out_u8) prep_image_kernel<false, true><<<(unsigned)blocks, 256, 0, s>>>(p);
        else prep_image_kernel<false, false><<<(unsigned)blocks, 256, 0, s>>>(p);
    }
    RTOD_LAUNCH_OK("prep_image_kernel");
    return RTOD_OK;
}

extern "C" int rtod_rescale_boxes(const float* rows, int D, const float* im_dims, int n_img, int inp_dim, int ref_dim,
                                  float* out_rows, float* out_dims, void* stream) {
    if (D < 0 || n_img < 0 || inp_dim <= 0 || ref_dim <= 0) return fail(RTOD_ERR_BAD_ARG, "rtod_rescale_boxes: bad size");
    if (D == 0) return RTOD_OK;
    if (!rows || !im_dims || !out_rows || n_img == 0)
        return fail(RTOD_ERR_BAD_ARG, "rtod_rescale_boxes: null pointer / no images");
    rescale_boxes_kernel<<<ceil_div(D, 128), 128, 0, (cudaStream_t)stream>>>(rows, D, im_dims, n_img, (float)inp_dim,
                                                                             (float)ref_dim, out_rows, out_dims);
    RTOD_LAUNCH_OK("rescale_boxes_kernel");
    return RTOD_OK;
}

extern "C" int rtod_pack_detections(const float* rows, int n_rows, const int* count, float first_frame, int capacity,
                                    float* payload, void* stream) {
    if (n_rows < 0 || capacity < 0) return fail(RTOD_ERR_BAD_ARG, "rtod_pack_detections: bad size");
    if (!count || !payload || (n_rows > 0 && !rows)) return fail(RTOD_ERR_BAD_ARG, "rtod_pack_detections: null pointer");
    if ((reinterpret_cast<uintptr_t>(rows) | reinterpret_cast<uintptr_t>(payload)) & 15u)
        return fail(RTOD_ERR_BAD_ARG, "rtod_pack_detections: rows / payload must be 16-byte aligned");
    pack_detections_kernel<<<ceil_div(capacity + 1, 256), 256, 0, (cudaStream_t)stream>>>(rows, n_rows, count, first_frame, capacity,
                                                                                         payload);
    RTOD_LAUNCH_OK("pack_detections_kernel");
    return RTOD_OK;
}

extern "C" int rtod_bbox_iou_matrix(const float* pred_boxes, int P, int pred_stride, const float* target_boxes, int T,
                                    int target_stride, int use_threshold, double threshold, float* out, void* stream) {
    if (P < 0 || T < 0 || pred_stride < 4 || target_stride < 4) return fail(RTOD_ERR_BAD_ARG, "rtod_bbox_iou_matrix: bad shape");
    if (P == 0 || T == 0) return RTOD_OK;
    if (!pred_boxes || !target_boxes || !out) return fail(RTOD_ERR_BAD_ARG, "rtod_bbox_iou_matrix: null pointer");
    if ((P + 7) / 8 > 65535) return fail(RTOD_ERR_UNSUPPORTED, "rtod_bbox_iou_matrix: more than 524280 predictions");
    dim3 grid((unsigned)ceil_div(T, 256), (unsigned)((P + 7) / 8));
    iou_matrix_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pred_boxes, P, pred_stride, target_boxes, T, target_stride,
                                                             use_threshold, threshold, out);
    RTOD_LAUNCH_OK("iou_matrix_kernel");
    return RTOD_OK;
}
