/*
 * rtod.h -- C ABI of librtod.so, the B200-native (sm_100a) YOLOv3 detection hot path.
 *
 * The reference (uguryagmur/RealTimeObjectDetection) has no FFI: its boundary is the Python
 * surface of src/darknet.py and src/util.py.  Each entry point below names the reference
 * interface it replaces; the Python mirror of that interface
 * (realtimeobjectdetection_b200/{darknet,util}.py) binds these symbols with ctypes and keeps
 * the reference's names, argument order and return conventions.
 *
 * Conventions
 *   - plain pointers and sizes only; every data pointer is a DEVICE pointer owned by the
 *     caller unless the name says _host; the library allocates no device memory;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); all work is
 *     enqueued asynchronously on it, no hidden synchronisation, CUDA-graph capturable;
 *   - return value 0 = RTOD_OK, negative = error; rtod_last_error() returns a thread-local
 *     human-readable message for the last failing call on this thread;
 *   - never throws, exits or prints.
 */
#ifndef RTOD_H_
#define RTOD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTOD_ABI_VERSION 2

enum {
    RTOD_OK = 0,
    RTOD_ERR_BAD_ARG = -1,      /* null pointer, negative size, inconsistent shapes          */
    RTOD_ERR_UNSUPPORTED = -2,  /* a shape/option the kernels do not implement               */
    RTOD_ERR_CAPACITY = -3,     /* caller-provided buffer/workspace too small                */
    RTOD_ERR_CUDA = -4,         /* a CUDA runtime/driver call failed (see rtod_last_error)   */
    RTOD_ERR_STATE = -5,        /* call order violated (e.g. forward before bind/weights)    */
    RTOD_ERR_DEVICE = -6        /* kernel-side failure flag (pipeline time-out, bad tile)    */
};

/* Block types of a Darknet cfg (src/darknet.py:460-526). */
enum {
    RTOD_LAYER_CONV = 0,
    RTOD_LAYER_SHORTCUT = 1,
    RTOD_LAYER_ROUTE = 2,
    RTOD_LAYER_UPSAMPLE = 3,
    RTOD_LAYER_MAXPOOL = 4,
    RTOD_LAYER_YOLO = 5
};

#define RTOD_MAX_ANCHORS 8

/* One parsed cfg block, already resolved the way Darknet.create_modules does
 * (src/darknet.py:449-603): `pad` is the effective padding ((size-1)/2 iff cfg pad != 0),
 * route/shortcut sources are ABSOLUTE layer indices (-1 = unused). */
typedef struct RtodLayerDesc {
    int32_t type;             /* RTOD_LAYER_*                                              */
    int32_t filters;          /* conv: output channels                                     */
    int32_t size;             /* conv / maxpool: kernel size                               */
    int32_t stride;           /* conv / maxpool: stride                                    */
    int32_t pad;              /* conv: effective zero padding                              */
    int32_t batch_normalize;  /* conv: 1 = BatchNorm2d follows (folded, eval semantics)    */
    int32_t leaky;            /* conv: 1 = LeakyReLU(0.1) follows, 0 = linear              */
    int32_t src0;             /* shortcut: i-1; route: first source                        */
    int32_t src1;             /* shortcut: i+from; route: second source or -1              */
    int32_t num_anchors;      /* yolo: anchors selected by mask                            */
    int32_t classes;          /* yolo: number of classes                                   */
    float anchors[2 * RTOD_MAX_ANCHORS]; /* yolo: (w,h) pixel pairs in mask order          */
} RtodLayerDesc;

typedef struct RtodPlan RtodPlan;

/* plan flags */
#define RTOD_PLAN_KEEP_ALL 1u     /* no buffer reuse: every layer output stays readable    */
#define RTOD_PLAN_CONV_SIMT 2u    /* validation only: CUDA-core convs instead of tcgen05   */
#define RTOD_PLAN_NO_AUTOTUNE 4u  /* bind: heuristic launch configurations, no timing runs */
#define RTOD_PLAN_BF16 8u         /* bf16 activations/weights (default: fp16, same tensor rate, 8x finer rounding) */
#define RTOD_PLAN_NO_WSPLIT 16u   /* fp16 only: no two-term weights in the HBM-bound early layers           */

int rtod_abi_version(void);
const char* rtod_last_error(void);

/* ---- Darknet.__init__ / create_modules (src/darknet.py:176-189, 449-603) --------------
 * Builds the execution plan (shape inference, NHWC buffer liveness plan, concat aliasing,
 * per-layer kernel selection) for a fixed input shape [batch, in_c, in_h, in_w];
 * inp_dim is int(net_info["height"]) used by the decode (src/darknet.py:258). */
int rtod_plan_create(const RtodLayerDesc* layers, int n_layers, int batch, int in_c, int in_h,
                     int in_w, int inp_dim, unsigned flags, RtodPlan** out_plan);
void rtod_plan_destroy(RtodPlan* plan);
size_t rtod_plan_workspace_bytes(const RtodPlan* plan);
/* part of rtod_plan_workspace_bytes that is not activations: split-K partial tiles + arrival counters */
size_t rtod_plan_scratch_bytes(const RtodPlan* plan);   /* activations + head logits     */
size_t rtod_plan_weight_bytes(const RtodPlan* plan);      /* packed bf16 weights + biases  */
int rtod_plan_num_rows(const RtodPlan* plan);             /* N of the [B, N, 5+C] output   */
int rtod_plan_num_attrs(const RtodPlan* plan);            /* 5+C (0 if no yolo layer)      */
int rtod_plan_layer_shape(const RtodPlan* plan, int layer, int* c, int* h, int* w);
int rtod_plan_launch_count(const RtodPlan* plan);         /* kernels per forward           */
double rtod_plan_conv_flops(const RtodPlan* plan);        /* 2*M*N*K summed over convs     */

/* Binds caller-owned device arenas (256-byte aligned) and encodes the TMA descriptors. */
int rtod_plan_bind(RtodPlan* plan, void* workspace, size_t workspace_bytes, void* weight_arena,
                   size_t weight_bytes);

/* ---- Darknet.load_weights / load_state_dict (src/darknet.py:316-410) -------------------
 * Folds BatchNorm (eval: w' = w*g/sqrt(var+eps), b' = beta - mean*g/sqrt(var+eps)) and
 * re-lays the fp32 [Cout,Cin,k,k] weight out as K-major fp16 / bf16 [Cout_pad][k*k*Cin] on device (two-term
 * layers: hi rows then lo rows).
 * bias may be null (BN convs); the four BN pointers are null for convs without BN. */
int rtod_plan_set_conv_weights(RtodPlan* plan, int layer, const float* weight, const float* bias,
                               const float* bn_gamma, const float* bn_beta, const float* bn_mean,
                               const float* bn_var, float bn_eps, void* stream);

/* ---- Darknet.forward (src/darknet.py:199-253) ------------------------------------------
 * x: [batch, in_c, in_h, in_w] fp32 NCHW; pred: [batch, N, 5+C] fp32.  train != 0 selects the
 * TRAIN=True decode (sigmoids only, src/util.py:211).  pred may be null when the cfg has no
 * yolo layer. */
int rtod_plan_forward(RtodPlan* plan, const float* x_nchw, float* pred, int train, void* stream);

/* Same forward for uint8 frames: x_planes is [batch, 3, in_h, in_w] uint8 (what rtod_prep_image writes with
 * out_u8), interpreted as value / 255 like prep_image (src/util.py:396) -- the scale is folded into the stem's
 * weights, the pixels enter the tensor cores exactly.  fp16 storage plans whose first layer is a 3x3 / stride-1
 * convolution of 16 / 32 / 64 filters over 3 channels (both reference networks); RTOD_ERR_UNSUPPORTED otherwise. */
int rtod_plan_forward_u8(RtodPlan* plan, const unsigned char* x_planes, float* pred, int train, void* stream);

/* Measurement: same as rtod_plan_forward, but brackets every layer with CUDA events on `stream`,
 * synchronises, and returns the device time of each layer's launch in layer_ms_host[0..n_layers)
 * and of the decode launch in layer_ms_host[n_layers] (HOST arrays; layer_kind_host may be null:
 * 0 = no kernel (alias/fused), 1 = tcgen05 conv, 2 = CUDA-core conv, 3 = stem conv, 4 = other). */
int rtod_plan_forward_profile(RtodPlan* plan, const float* x_nchw, float* pred, int train, void* stream,
                              float* layer_ms_host, int* layer_kind_host);
/* Same forward, timed in two classes with events only where the stream switches between them: the
 * tcgen05 convolution launches (conv_ms) and every other kernel (other_ms: stem, upsample, decode ...). */
int rtod_plan_forward_segments(RtodPlan* plan, const float* x_nchw, float* pred, int train, void* stream,
                               float* conv_ms, float* other_ms);
double rtod_plan_layer_flops(const RtodPlan* plan, int layer); /* 2*M*N*K of one convolution  */
/* which kernel a bound plan runs for convolution `layer` (tests and profiling; not a reference call) */
enum {
    RTOD_CONV_NONE = 0,      /* not a convolution / plan not bound                               */
    RTOD_CONV_STEM = 1,      /* stem.cu: 3-channel fp32 NCHW image -> NHWC bf16                  */
    RTOD_CONV_TC = 2,        /* conv_tc.cu: tcgen05 implicit GEMM, one CTA per tile              */
    RTOD_CONV_TC_PAIR = 3,   /* conv_pair.cu: tcgen05 cta_group::2, one CTA pair per 256x256 tile */
    RTOD_CONV_TC_PATCH = 4,  /* retired (halo-patch 3x3 variant); never returned                 */
    RTOD_CONV_SIMT = 5       /* conv_simt.cu: CUDA-core validation path (RTOD_PLAN_CONV_SIMT)     */
};
int rtod_plan_conv_backend(const RtodPlan* plan, int layer);

/* launch configuration of convolution `layer` of a bound plan (tests, profiling): out12 = {backend, N tile, CTAs per
 * SM, resident weights, staging slices, split-K, epilogue warps, activation producers, pipelines per CTA, ring
 * stages, two-term weights, grid size} */
int rtod_plan_conv_config(const RtodPlan* plan, int layer, int* out12);

/* Debug/validation: copy one layer's output to fp32 NCHW [batch, c, h, w].  Meaningful after a
 * forward of a plan created with RTOD_PLAN_KEEP_ALL (or for the last layer). */
int rtod_plan_read_layer(RtodPlan* plan, int layer, float* out_nchw, void* stream);
/* Polls the device-side failure flag of the last forward (synchronises the stream). */
int rtod_plan_check(RtodPlan* plan, void* stream);
/* Failure reporting without a synchronisation: `host_flag` is a pinned (device-mapped, UVA) host int owned by
 * the caller; a kernel that times out stores its failure code there, the caller polls it between calls.
 * Null detaches.  Enqueued on `stream`. */
int rtod_plan_set_error_sink(RtodPlan* plan, int* host_flag, void* stream);
/* After a reported failure: clears the device flag and the split-K arrival counters (a timed-out layer leaves
 * them mid-count), so that the next forward starts clean.  Enqueued on `stream`. */
int rtod_plan_reset_errors(RtodPlan* plan, void* stream);
/* 1 = the plan stores activations/weights as fp16, 0 = bf16 */
int rtod_plan_is_f16(const RtodPlan* plan);
/* 1 = convolution `layer` keeps two-term (hi + lo) weights */
int rtod_plan_conv_w_split(const RtodPlan* plan, int layer);
/* 1 if convolution `layer` of a bound plan runs in row mode (3x3 / stride 1 on the tcgen05 kernel: one tiled TMA load
 * per filter row, the three kx taps as row-shifted views of it; csrc/conv_tc.cuh) */
int rtod_plan_conv_row_mode(const RtodPlan* plan, int layer);

/* ---- util.predict_transform (src/util.py:175-239) -------------------------------------
 * head: [B, A*(5+C), G, G] fp32 NCHW -> out: [B, G*G*A, 5+C] fp32; anchors_host: A (w,h) pixel
 * pairs in HOST memory; stride = inp_dim / G. */
int rtod_yolo_decode(const float* head_nchw, int B, int G, int A, int C, int inp_dim,
                     const float* anchors_host, int train, float* out, void* stream);

/* ---- util.write_results (+confidence_mask, bbox_iou inside it; src/util.py:242-346) -----
 * pred: [B, N, 5+C] fp32 (not modified).  out_rows: [cap, 8] fp32 rows
 * [img, x1, y1, x2, y2, obj, cls_conf, cls] ordered image, class ascending, objectness
 * descending (ties: lower row index first).  *out_count (device int) receives the number of
 * detections D; rows beyond cap are dropped (D may exceed cap: compare after the copy back).
 * Limits: C <= 4096, N <= 2^20. */
size_t rtod_write_results_workspace_bytes(int B, int N, int C);
int rtod_write_results(const float* pred, int B, int N, int C, float confidence, float nms_conf,
                       float* out_rows, int cap, int* out_count, void* workspace,
                       size_t workspace_bytes, void* stream);

/* ---- util.confidence_mask (src/util.py:106-117) -- out = pred * (pred[...,4] > conf) */
int rtod_confidence_mask(const float* pred, long long rows, int attrs, float confidence, float* out,
                         void* stream);

/* ---- util.bbox_iou (src/util.py:120-153) --------------------------------------------------
 * box1: n1 boxes, box2: n2 boxes, rows of `stride` floats whose first four are x1,y1,x2,y2;
 * n1 == n2, or one of them == 1 (broadcast).  out: max(n1,n2) fp32 IoUs (+1 pixel convention,
 * every operation individually rounded like the reference's tensor ops). */
int rtod_bbox_iou(const float* box1, int n1, int stride1, const float* box2, int n2, int stride2,
                  float* out, void* stream);

/* ---- util.prep_image / letterbox_image (src/util.py:349-397) -------------------------------
 * src: [B, src_h, src_w, 3] uint8 HWC frames (BGR, as cv2.imread delivers them), all of one size.
 * out: [B, 3, inp_dim, inp_dim], fp32 (value / 255, what prep_image returns) or uint8 planes (out_u8 != 0).
 * The frame is resized with cv2.INTER_CUBIC semantics to (new_w, new_h) = rtod_letterbox_geometry and pasted
 * on a canvas filled with 128; channels are reversed (BGR -> RGB, prep_image's default mode) unless keep_order.
 * resize_mode 0: fp32 cubic weights (the stock opencv-python wheel routes 8-bit cubic resizes to IPP; equal to it
 * to within one grey level on < 0.03 % of the values); 1: OpenCV's own fixed-point path, bit-identical to
 * cv2.resize built/run without IPP. */
int rtod_letterbox_geometry(int src_w, int src_h, int inp_dim, int* new_w, int* new_h, int* left, int* top);
int rtod_prep_image(const unsigned char* src, int B, int src_h, int src_w, int inp_dim, int keep_order,
                    int resize_mode, int out_u8, void* out, void* stream);

/* ---- Darknetv3Detector.convert_box_dims_to_original_image + clamp_box_dims (detect.py:120-136) --------
 * rows: [D, 8] write_results rows whose column 0 indexes im_dims; im_dims: [n_img, 4] fp32 (w, h, w, h)
 * (detect.py:247-248).  out_rows: [D, 8] with columns 1..4 mapped to source-image pixels and clamped to
 * [0, w] / [0, h]; out_dims (may be null): [D, 4] the im_dims row of every detection (the reference returns it).
 * The scale is min(ref_dim / w, ref_dim / h) -- the reference hard-codes ref_dim = 416 (detect.py:130) --
 * the centring offset uses inp_dim (:131-134). */
int rtod_rescale_boxes(const float* rows, int D, const float* im_dims, int n_img, int inp_dim, int ref_dim,
                       float* out_rows, float* out_dims, void* stream);

/* ---- DarknetValidator.create_iou_matrix_for_predictions_and_targets (test.py:139-151) ------------------
 * out[p * T + t] = bbox_iou(pred_boxes[p], target_boxes[t]) (rows of `stride` floats starting with
 * x1,y1,x2,y2: pass pred + 1 for write_results rows), zeroed unless (double)iou > threshold when
 * use_threshold != 0 (the reference compares Python floats). */
int rtod_bbox_iou_matrix(const float* pred_boxes, int P, int pred_stride, const float* target_boxes, int T,
                         int target_stride, int use_threshold, double threshold, float* out, void* stream);

/* ---- multi-GPU: payload of the per-step detection gather (no reference counterpart: the reference is
 * single-process) ---------------------------------------------------------
 * payload: [capacity + 1, 8] fp32.  Row 0 = (count, 0, ...); row 1 + i = rows[i] with column 0 (the image
 * index) shifted by first_frame for i < min(*count, n_rows), zeros beyond.  `count` is a device int (the
 * count rtod_write_results stored): nothing is read on the host. */
int rtod_pack_detections(const float* rows, int n_rows, const int* count, float first_frame, int capacity,
                         float* payload, void* stream);

/* ---- measurement aid (bench.py "clocks"; no reference counterpart) ------------------------
 * One thread samples the SM clock it runs on: `samples` windows of `interval_us` microseconds,
 * out_mhz[i] = SM cycles / wall time of window i.  Launch it on a side stream next to the work
 * being measured; it ends by itself after samples * interval_us. */
int rtod_sm_clock_probe(float* out_mhz, int samples, int interval_us, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RTOD_H_ */
