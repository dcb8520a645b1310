// plan.cu -- host runtime of Darknet.forward (src/darknet.py:199-303): turns the parsed cfg into
// a fixed launch sequence over pre-planned NHWC buffers.
//
//   * shape inference mirrors create_modules (src/darknet.py:449-603);
//   * the reference caches EVERY layer output (outputs{}, :301); here only tensors with a later
//     reader stay alive: buffers are placed in one arena by lifetime (largest first, first fit);
//   * shortcut (:263-268) is folded into the epilogue of the convolution that feeds it, route with
//     one source (:277-278) is an alias, route with two sources (:285-288) is a shared buffer its
//     producers write channel slices of (zero-copy torch.cat), yolo (:226-247) is one decode launch
//     over all heads at the end;
//   * convolutions whose only reader is a yolo layer keep fp32 logits (no bf16 rounding of the
//     values the sigmoid/exp see).
#include <algorithm>
#include <new>
#include <vector>

#include <cstdlib>

#include "conv_tc.cuh"
#include "decode.cuh"
#include "layers.cuh"

namespace rtod {

constexpr size_t kArenaAlign = 1024;
constexpr size_t kSplitScratchBytes = 8u << 20;   // split-K partial accumulators (small batches only)
constexpr int kSplitCounters = 16384;
constexpr int kInputLayer = -1;          // pseudo index of the network input

struct Buf {
    size_t bytes = 0, offset = 0;
    int pitch = 0, H = 0, W = 0, fp32 = 0;
    int first = 0, last = 0;             // lifetime in layer indices (inclusive)
};

struct Node {                            // one cfg block
    RtodLayerDesc d{};
    int C = 0, H = 0, W = 0;             // output shape
    int alias_of = -2;                   // >= -1: this output IS that layer's output (no kernel)
    int buf = -1, ch_off = 0;            // physical placement (roots only)
    std::vector<int> readers;            // layers reading this output
    // convolution
    int Cin = 0, Cout_pad = 0, K = 0;
    bool stem = false, head = false, use_tc = false, weights_set = false;
    int w_split = 0;                     // weights kept as hi + lo (two MMAs per K step), see choose_w_split
    int res_src = -2;                    // fused shortcut operand (layer index) or -2
    size_t w_off = 0, wf_off = 0, bias_off = 0;
    ConvArgs args{};
    ConvTcLaunch tc{};
    // route fall-back (copy) / shortcut fall-back (add)
    bool copy_concat = false;
};

}  // namespace rtod

struct RtodPlan {
    std::vector<rtod::Node> nodes;
    std::vector<rtod::Buf> bufs;
    int batch = 0, in_c = 0, in_h = 0, in_w = 0, inp_dim = 0;
    unsigned flags = 0;
    int f16 = 1;                         // 16-bit storage type: fp16 (default) or bf16 (RTOD_PLAN_BF16)
    int input_buf = -1;                  // NHWC copy of the input when layer 0 is not a stem conv
    size_t workspace_bytes = 0, weight_bytes = 0;
    size_t split_off = 0;                // offset of the split-K scratch inside the workspace
    unsigned char* ws = nullptr;
    unsigned char* wa = nullptr;
    int* err_flag = nullptr;
    int* err_sink = nullptr;             // device view of the caller's pinned failure int (or null)
    bool bound = false;
    // decode
    rtod::DecodeHeads heads{};
    int n_rows = 0, n_attrs = 0;
    int launches = 0;
    double conv_flops = 0.0;
};

namespace rtod {

namespace {

int root_of(const RtodPlan& p, int i) {
    while (i >= 0 && p.nodes[i].alias_of >= -1) i = p.nodes[i].alias_of;
    return i;
}

// shape/pitch view of layer i's output (i == -1: the NHWC input copy)
Act act_of(const RtodPlan& p, int i) {
    Act a{};
    if (i == kInputLayer) {
        const Buf& b = p.bufs[p.input_buf];
        a.ptr = p.ws ? p.ws + b.offset : nullptr;
        a.C = p.in_c; a.pitch = b.pitch; a.H = p.in_h; a.W = p.in_w; a.fp32 = 0; a.f16 = p.f16;
        return a;
    }
    const Node& n = p.nodes[i];
    const int r = root_of(p, i);
    if (r == kInputLayer) return act_of(p, kInputLayer);
    const Node& rn = p.nodes[r];
    const Buf& b = p.bufs[rn.buf];
    const size_t esz = b.fp32 ? 4 : 2;
    a.ptr = p.ws ? p.ws + b.offset + (size_t)rn.ch_off * esz : nullptr;
    a.C = n.C; a.pitch = b.pitch; a.H = n.H; a.W = n.W; a.fp32 = b.fp32; a.f16 = p.f16;
    return a;
}

int new_buf(RtodPlan& p, int pitch, int H, int W, int fp32, int first) {
    Buf b;
    b.pitch = pitch; b.H = H; b.W = W; b.fp32 = fp32;
    b.bytes = align_up((size_t)p.batch * H * W * pitch * (fp32 ? 4 : 2), kArenaAlign);
    b.first = first; b.last = first;
    p.bufs.push_back(b);
    return (int)p.bufs.size() - 1;
}

int infer_shapes(RtodPlan& p) {
    int C = p.in_c, H = p.in_h, W = p.in_w;
    bool after_yolo = false;
    const int n = (int)p.nodes.size();
    for (int i = 0; i < n; ++i) {
        Node& nd = p.nodes[i];
        const RtodLayerDesc& d = nd.d;
        if (after_yolo && d.type != RTOD_LAYER_ROUTE)
            return fail(RTOD_ERR_UNSUPPORTED, "layer %d: only a route may follow a yolo layer", i);
        after_yolo = false;
        switch (d.type) {
        case RTOD_LAYER_CONV:
            if (d.filters <= 0 || d.size <= 0 || d.stride <= 0 || d.pad < 0)
                return fail(RTOD_ERR_BAD_ARG, "layer %d: bad convolution parameters", i);
            if (H + 2 * d.pad < d.size || W + 2 * d.pad < d.size)
                return fail(RTOD_ERR_BAD_ARG, "layer %d: kernel larger than padded input", i);
            nd.Cin = C;
            C = d.filters;
            H = (H + 2 * d.pad - d.size) / d.stride + 1;
            W = (W + 2 * d.pad - d.size) / d.stride + 1;
            break;
        case RTOD_LAYER_SHORTCUT: {
            if (d.src0 != i - 1 || d.src1 < 0 || d.src1 >= i)
                return fail(RTOD_ERR_BAD_ARG, "layer %d: bad shortcut sources (%d, %d)", i, d.src0, d.src1);
            const Node& o = p.nodes[d.src1];
            if (o.C != C || o.H != H || o.W != W)
                return fail(RTOD_ERR_BAD_ARG, "layer %d: shortcut shapes differ", i);
            break;
        }
        case RTOD_LAYER_ROUTE: {
            if (d.src0 < 0 || d.src0 >= i || d.src1 >= i)
                return fail(RTOD_ERR_BAD_ARG, "layer %d: bad route sources (%d, %d)", i, d.src0, d.src1);
            const Node& a = p.nodes[d.src0];
            C = a.C; H = a.H; W = a.W;
            if (d.src1 >= 0) {
                const Node& b = p.nodes[d.src1];
                if (b.H != H || b.W != W)
                    return fail(RTOD_ERR_BAD_ARG, "layer %d: route sources have different sizes", i);
                C += b.C;
            }
            break;
        }
        case RTOD_LAYER_UPSAMPLE:
            H *= 2; W *= 2;                                          // src/darknet.py:591 (always x2)
            break;
        case RTOD_LAYER_MAXPOOL:
            if (d.size < 1 || d.stride < 1 || (d.stride == 1 && d.size < 2) || H < d.size || W < d.size)
                return fail(RTOD_ERR_BAD_ARG, "layer %d: bad maxpool parameters", i);
            if (d.stride != 1) {
                H = (H - d.size) / d.stride + 1;
                W = (W - d.size) / d.stride + 1;
            } else {                                                 // MaxPoolStride1, :37-46
                H = (H - 1) / (d.size - 1) + 1;
                W = (W - 1) / (d.size - 1) + 1;
            }
            break;
        case RTOD_LAYER_YOLO:
            if (i == 0) return fail(RTOD_ERR_BAD_ARG, "yolo layer cannot be first");
            after_yolo = true;
            break;
        default:
            return fail(RTOD_ERR_BAD_ARG, "layer %d: unknown type %d", i, d.type);
        }
        nd.C = C; nd.H = H; nd.W = W;
        if (H <= 0 || W <= 0) return fail(RTOD_ERR_BAD_ARG, "layer %d: empty output", i);
    }
    return RTOD_OK;
}

// Two-term weights (w = hi + lo in fp16, two MMAs per K step) where the extra tensor work hides behind the
// layer's own memory time: arithmetic intensity 2*M*N*K / bytes below the machine's ridge (211 FLOP/B
// measured), and only for layers with many output elements -- every rounded element contributes equally to the
// prediction error of a deep network, so the large early layers dominate it ("numerics").
// A function of the layer shape only.
int choose_w_split(const RtodPlan& p, const Node& nd) {
    if (!p.f16 || (p.flags & RTOD_PLAN_NO_WSPLIT) || nd.stem || nd.head) return 0;
    const double in_px = (double)(nd.d.stride * nd.d.stride);            // input pixels read per output pixel
    const double bytes = 2.0 * (in_px * nd.Cin + nd.d.filters * (nd.res_src >= -1 ? 2.0 : 1.0));
    const double ai = 2.0 * nd.d.filters * nd.K / bytes;
    double ai_max = 200.0, min_elems = 300000.0;
    if (const char* e = getenv("RTOD_WSPLIT_AI")) ai_max = atof(e);
    if (const char* e = getenv("RTOD_WSPLIT_ELEMS")) min_elems = atof(e);
    return ai < ai_max && (double)nd.H * nd.W * nd.d.filters >= min_elems ? 1 : 0;
}

void add_reader(RtodPlan& p, int src, int reader) {
    if (src >= 0) p.nodes[src].readers.push_back(reader);
}

int build(RtodPlan& p) {
    const int n = (int)p.nodes.size();
    int rc = infer_shapes(p);
    if (rc) return rc;

    // ---- who reads what (explicit sources + the implicit x = previous output) ----------------
    for (int i = 0; i < n; ++i) {
        const RtodLayerDesc& d = p.nodes[i].d;
        switch (d.type) {
        case RTOD_LAYER_CONV: case RTOD_LAYER_UPSAMPLE: case RTOD_LAYER_MAXPOOL: case RTOD_LAYER_YOLO:
            add_reader(p, i - 1, i);
            break;
        case RTOD_LAYER_SHORTCUT:
            add_reader(p, d.src0, i);
            add_reader(p, d.src1, i);
            break;
        case RTOD_LAYER_ROUTE:
            add_reader(p, d.src0, i);
            add_reader(p, d.src1, i);
            break;
        }
    }

    // ---- aliases and fusions ----------------------------------------------------------------------
    const bool stem_input = n > 0 && p.nodes[0].d.type == RTOD_LAYER_CONV && p.in_c == 3 &&
                            p.nodes[0].d.size == 3 && p.nodes[0].d.filters % 8 == 0 &&
                            p.nodes[0].d.filters <= 256;
    for (int i = 0; i < n; ++i) {
        Node& nd = p.nodes[i];
        const RtodLayerDesc& d = nd.d;
        if (d.type == RTOD_LAYER_YOLO) nd.alias_of = i - 1;
        if (d.type == RTOD_LAYER_ROUTE && d.src1 < 0) nd.alias_of = d.src0;
        if (d.type == RTOD_LAYER_CONV) {
            nd.stem = (i == 0 && stem_input);
            nd.head = nd.readers.size() == 1 && p.nodes[nd.readers[0]].d.type == RTOD_LAYER_YOLO;
            nd.K = d.size * d.size * nd.Cin;
            nd.Cout_pad = d.filters <= 128 ? (d.filters + 31) / 32 * 32 : (d.filters + 127) / 128 * 128;
            if (!nd.head && d.filters % 8 != 0)
                return fail(RTOD_ERR_UNSUPPORTED, "layer %d: %d filters (bf16 NHWC needs a multiple of 8)", i,
                            d.filters);
        }
    }
    for (int i = 0; i < n; ++i) {
        Node& nd = p.nodes[i];
        if (nd.d.type == RTOD_LAYER_YOLO && !(p.nodes[i - 1].d.type == RTOD_LAYER_CONV && p.nodes[i - 1].head))
            return fail(RTOD_ERR_UNSUPPORTED, "layer %d: yolo must directly follow its own convolution", i);
        if (nd.d.type != RTOD_LAYER_SHORTCUT) continue;
        Node& prev = p.nodes[i - 1];
        // outputs[i-1] is read by the shortcut only -> fold the add into that convolution
        if (prev.d.type == RTOD_LAYER_CONV && !prev.stem && !prev.head && prev.readers.size() == 1 &&
            root_of(p, nd.d.src1) != i - 1) {
            prev.res_src = nd.d.src1;
            nd.alias_of = i - 1;
        }
    }

    // ---- buffers: roots first, then concat groups --------------------------------------------
    if (!stem_input) {
        if (p.in_c % 8 != 0)
            return fail(RTOD_ERR_UNSUPPORTED, "input with %d channels needs a 3x3 stem convolution first", p.in_c);
        p.input_buf = new_buf(p, p.in_c, p.in_h, p.in_w, 0, kInputLayer);
    }
    std::vector<int> in_group(n, 0);
    for (int i = 0; i < n; ++i) {                      // two-source routes: try zero-copy
        Node& nd = p.nodes[i];
        if (nd.d.type != RTOD_LAYER_ROUTE || nd.d.src1 < 0) continue;
        const int r0 = root_of(p, nd.d.src0), r1 = root_of(p, nd.d.src1);
        const bool ok = r0 >= 0 && r1 >= 0 && r0 != r1 && !in_group[r0] && !in_group[r1] &&
                        !p.nodes[r0].head && !p.nodes[r1].head &&
                        p.nodes[r0].d.type != RTOD_LAYER_ROUTE && p.nodes[r1].d.type != RTOD_LAYER_ROUTE &&
                        p.nodes[nd.d.src0].C % 8 == 0 && p.nodes[nd.d.src1].C % 8 == 0;
        nd.buf = new_buf(p, nd.C, nd.H, nd.W, 0, ok ? std::min(r0, r1) : i);
        if (ok) {
            p.nodes[r0].buf = nd.buf; p.nodes[r0].ch_off = 0;
            p.nodes[r1].buf = nd.buf; p.nodes[r1].ch_off = p.nodes[nd.d.src0].C;
            in_group[r0] = in_group[r1] = 1;
        } else {
            nd.copy_concat = true;
        }
    }
    for (int i = 0; i < n; ++i) {
        Node& nd = p.nodes[i];
        if (nd.alias_of >= -1 || nd.buf >= 0) continue;
        if (nd.d.type == RTOD_LAYER_CONV && nd.head)
            nd.buf = new_buf(p, (nd.C + 7) / 8 * 8, nd.H, nd.W, 1, i);
        else
            nd.buf = new_buf(p, nd.C, nd.H, nd.W, 0, i);
    }
    // ---- lifetimes --------------------------------------------------------------------------------
    const int forever = n + 1;
    auto touch = [&](int layer, int at) {              // layer's data is needed at step `at`
        const int r = root_of(p, layer);
        if (r == kInputLayer) { if (p.input_buf >= 0) p.bufs[p.input_buf].last = std::max(p.bufs[p.input_buf].last, at); return; }
        if (r < 0) return;
        Buf& b = p.bufs[p.nodes[r].buf];
        b.last = std::max(b.last, at);
        b.first = std::min(b.first, r);
    };
    for (int i = 0; i < n; ++i) {
        const Node& nd = p.nodes[i];
        touch(i, i);
        for (int r : nd.readers) touch(i, r);
        if (nd.res_src >= 0) touch(nd.res_src, i);
        if (nd.d.type == RTOD_LAYER_CONV && nd.head) touch(i, forever);      // read by the decode launch
        if (nd.d.type == RTOD_LAYER_CONV || nd.d.type == RTOD_LAYER_UPSAMPLE || nd.d.type == RTOD_LAYER_MAXPOOL)
            touch(i - 1, i);
    }
    if (n > 0) touch(n - 1, forever);
    if (p.flags & RTOD_PLAN_KEEP_ALL)
        for (Buf& b : p.bufs) b.last = forever;

    // ---- arena placement: largest first, lowest non-conflicting offset ---------------------------
    std::vector<int> order(p.bufs.size());
    for (size_t k = 0; k < order.size(); ++k) order[k] = (int)k;
    std::sort(order.begin(), order.end(), [&](int a, int b) {
        return p.bufs[a].bytes != p.bufs[b].bytes ? p.bufs[a].bytes > p.bufs[b].bytes : a < b;
    });
    std::vector<int> placed;
    size_t arena = kArenaAlign;                        // first block: device failure flag
    for (int id : order) {
        Buf& b = p.bufs[id];
        std::vector<std::pair<size_t, size_t>> busy;
        for (int q : placed) {
            const Buf& o = p.bufs[q];
            if (o.first <= b.last && b.first <= o.last) busy.push_back({o.offset, o.offset + o.bytes});
        }
        std::sort(busy.begin(), busy.end());
        size_t off = kArenaAlign;
        for (auto& iv : busy) {
            if (off + b.bytes <= iv.first) break;
            off = std::max(off, iv.second);
        }
        b.offset = off;
        arena = std::max(arena, off + b.bytes);
        placed.push_back(id);
    }
    // split-K scratch (fp32 partial tiles) + per-tile arrival counters, shared by all layers (they run in order)
    p.split_off = align_up(arena, 256);
    p.workspace_bytes = p.split_off + kSplitScratchBytes + kSplitCounters * sizeof(int);

    // ---- weight arena ---------------------------------------------------------------------------
    size_t woff = 0;
    for (int i = 0; i < n; ++i) {
        Node& nd = p.nodes[i];
        if (nd.d.type != RTOD_LAYER_CONV) continue;
        nd.w_split = choose_w_split(p, nd);
        nd.w_off = woff;  woff = align_up(woff + ((size_t)nd.Cout_pad << nd.w_split) * nd.K * 2, 256);
        nd.bias_off = woff; woff = align_up(woff + (size_t)nd.Cout_pad * 4, 256);
        if (nd.stem) { nd.wf_off = woff; woff = align_up(woff + (size_t)nd.d.filters * nd.K * 4, 256); }
        const double M = (double)p.batch * nd.H * nd.W;
        p.conv_flops += 2.0 * M * nd.d.filters * nd.K;
    }
    p.weight_bytes = woff > 0 ? woff : 256;

    // ---- yolo heads ---------------------------------------------------------------------------------
    p.heads.count = 0;
    p.n_rows = 0;
    p.n_attrs = 0;
    for (int i = 0; i < n; ++i) {
        const Node& nd = p.nodes[i];
        if (nd.d.type != RTOD_LAYER_YOLO) continue;
        const RtodLayerDesc& d = nd.d;
        if (p.heads.count >= kMaxHeads) return fail(RTOD_ERR_UNSUPPORTED, "more than %d yolo layers", kMaxHeads);
        if (d.num_anchors <= 0 || d.num_anchors > RTOD_MAX_ANCHORS || d.classes < 0)
            return fail(RTOD_ERR_BAD_ARG, "layer %d: bad yolo parameters", i);
        const int L = 5 + d.classes;
        if (p.n_attrs && p.n_attrs != L)
            return fail(RTOD_ERR_UNSUPPORTED, "yolo layers with different class counts");   // torch.cat would fail too
        if (nd.H != nd.W) return fail(RTOD_ERR_UNSUPPORTED, "layer %d: non-square yolo grid %dx%d", i, nd.H, nd.W);
        if (nd.C != d.num_anchors * L)
            return fail(RTOD_ERR_BAD_ARG, "layer %d: %d channels feed a yolo layer that needs %d", i, nd.C,
                        d.num_anchors * L);
        const int G = nd.H;
        const int stride = p.inp_dim / G;                              // src/util.py:194
        if (stride <= 0 || p.inp_dim / stride != G)
            return fail(RTOD_ERR_BAD_ARG, "layer %d: net_info height %d does not match a %dx%d grid", i,
                        p.inp_dim, G, G);
        const int h = p.heads.count++;
        p.n_attrs = L;
        p.heads.grid[h] = G;
        p.heads.num_anchors[h] = d.num_anchors;
        p.heads.row_base[h] = p.n_rows;
        p.heads.stride[h] = (float)stride;
        for (int a = 0; a < d.num_anchors; ++a) {
            p.heads.anchor_w[h][a] = (float)((double)d.anchors[2 * a] / (double)stride);
            p.heads.anchor_h[h][a] = (float)((double)d.anchors[2 * a + 1] / (double)stride);
        }
        p.n_rows += G * G * d.num_anchors;
    }

    // ---- launches per forward ------------------------------------------------------------------------
    p.launches = 0;
    if (p.input_buf >= 0) ++p.launches;
    for (int i = 0; i < n; ++i) {
        const Node& nd = p.nodes[i];
        if (nd.alias_of >= -1) continue;
        if (nd.d.type == RTOD_LAYER_ROUTE) p.launches += nd.copy_concat ? 2 : 0;
        else ++p.launches;
    }
    if (p.heads.count) ++p.launches;
    return RTOD_OK;
}

int bind_layers(RtodPlan& p) {
    const int n = (int)p.nodes.size();
    int head_idx = 0;
    for (int i = 0; i < n; ++i) {
        Node& nd = p.nodes[i];
        if (nd.d.type == RTOD_LAYER_YOLO) {
            const Act raw = act_of(p, i - 1);
            p.heads.raw[head_idx] = reinterpret_cast<const float*>(raw.ptr);
            p.heads.pitch[head_idx] = raw.pitch;
            ++head_idx;
            continue;
        }
        if (nd.d.type != RTOD_LAYER_CONV) continue;
        ConvArgs& a = nd.args;
        a = ConvArgs{};
        if (!nd.stem) a.in = act_of(p, i - 1);
        a.out = act_of(p, i);
        a.out.C = nd.d.filters;
        a.w = p.wa + nd.w_off;
        a.w_split = nd.w_split;
        a.bias = reinterpret_cast<const float*>(p.wa + nd.bias_off);
        if (nd.res_src >= -1) {
            const Act r = act_of(p, nd.res_src);
            if (r.fp32) return fail(RTOD_ERR_UNSUPPORTED, "layer %d: fp32 shortcut operand", i);
            a.res = r.ptr;
            a.res_pitch = r.pitch;
        }
        a.split_scratch = reinterpret_cast<float*>(p.ws + p.split_off);
        a.split_scratch_bytes = kSplitScratchBytes;
        a.split_count = reinterpret_cast<int*>(p.ws + p.split_off + kSplitScratchBytes);
        a.split_count_n = kSplitCounters;
        a.B = p.batch; a.Cin = nd.Cin; a.Cout = nd.d.filters; a.Cout_pad = nd.Cout_pad;
        a.ks = nd.d.size; a.stride = nd.d.stride; a.pad = nd.d.pad; a.leaky = nd.d.leaky; a.K = nd.K;
        nd.use_tc = false;
        if (!nd.stem && !(p.flags & RTOD_PLAN_CONV_SIMT) && conv_tc_supported(a)) {
            const int rc = (p.flags & RTOD_PLAN_NO_AUTOTUNE) ? conv_tc_prepare(a, p.err_flag, &nd.tc)
                                                             : conv_tc_autotune(a, p.err_flag, &nd.tc, nullptr);
            if (rc) return rc;
            nd.use_tc = true;
        }
    }
    // small batches: every tcgen05 convolution prefetches the next one's weights into L2 (conv_tc.cu)
    if ((long long)p.batch * p.in_h * p.in_w <= 4ll * 416 * 416 && getenv("RTOD_NO_WEIGHT_PREFETCH") == nullptr) {
        Node* prev = nullptr;
        for (int i = 0; i < n; ++i) {
            Node& nd = p.nodes[i];
            if (nd.d.type != RTOD_LAYER_CONV || nd.alias_of >= -1 || !nd.use_tc) continue;
            if (prev) {
                prev->tc.p.pf_ptr = p.wa + nd.w_off;
                prev->tc.p.pf_bytes = ((unsigned long long)nd.Cout_pad << nd.w_split) * nd.K * 2ull;     // multiple of 16
            }
            prev = &nd;
        }
    }
    return RTOD_OK;
}

}  // namespace

}  // namespace rtod

using namespace rtod;

extern "C" int rtod_plan_create(const RtodLayerDesc* layers, int n_layers, int batch, int in_c, int in_h,
                                int in_w, int inp_dim, unsigned flags, RtodPlan** out_plan) {
    if (!out_plan) return fail(RTOD_ERR_BAD_ARG, "rtod_plan_create: out_plan is null");
    *out_plan = nullptr;
    if (!layers || n_layers <= 0 || batch <= 0 || in_c <= 0 || in_h <= 0 || in_w <= 0 || inp_dim <= 0)
        return fail(RTOD_ERR_BAD_ARG, "rtod_plan_create: bad arguments (n=%d B=%d C=%d H=%d W=%d inp_dim=%d)",
                    n_layers, batch, in_c, in_h, in_w, inp_dim);
    RtodPlan* p = new (std::nothrow) RtodPlan();
    if (!p) return fail(RTOD_ERR_BAD_ARG, "rtod_plan_create: out of host memory");
    p->batch = batch; p->in_c = in_c; p->in_h = in_h; p->in_w = in_w; p->inp_dim = inp_dim; p->flags = flags;
    p->f16 = (flags & RTOD_PLAN_BF16) ? 0 : 1;
    p->nodes.resize(n_layers);
    for (int i = 0; i < n_layers; ++i) p->nodes[i].d = layers[i];
    const int rc = build(*p);
    if (rc) { delete p; return rc; }
    *out_plan = p;
    return RTOD_OK;
}

extern "C" void rtod_plan_destroy(RtodPlan* plan) { delete plan; }
extern "C" size_t rtod_plan_workspace_bytes(const RtodPlan* p) { return p ? p->workspace_bytes : 0; }
extern "C" size_t rtod_plan_scratch_bytes(const RtodPlan* p) { return p ? p->workspace_bytes - p->split_off : 0; }
extern "C" size_t rtod_plan_weight_bytes(const RtodPlan* p) { return p ? p->weight_bytes : 0; }
extern "C" int rtod_plan_num_rows(const RtodPlan* p) { return p ? p->n_rows : 0; }
extern "C" int rtod_plan_num_attrs(const RtodPlan* p) { return p ? p->n_attrs : 0; }
extern "C" int rtod_plan_launch_count(const RtodPlan* p) { return p ? p->launches : 0; }
extern "C" double rtod_plan_conv_flops(const RtodPlan* p) { return p ? p->conv_flops : 0.0; }

extern "C" int rtod_plan_layer_shape(const RtodPlan* p, int layer, int* c, int* h, int* w) {
    if (!p || layer < 0 || layer >= (int)p->nodes.size())
        return fail(RTOD_ERR_BAD_ARG, "rtod_plan_layer_shape: bad plan or layer %d", layer);
    if (c) *c = p->nodes[layer].C;
    if (h) *h = p->nodes[layer].H;
    if (w) *w = p->nodes[layer].W;
    return RTOD_OK;
}

extern "C" int rtod_plan_bind(RtodPlan* p, void* workspace, size_t workspace_bytes, void* weight_arena,
                              size_t weight_bytes) {
    if (!p || !workspace || !weight_arena) return fail(RTOD_ERR_BAD_ARG, "rtod_plan_bind: null argument");
    if (workspace_bytes < p->workspace_bytes || weight_bytes < p->weight_bytes)
        return fail(RTOD_ERR_CAPACITY, "rtod_plan_bind: arenas too small (%zu < %zu or %zu < %zu)",
                    workspace_bytes, p->workspace_bytes, weight_bytes, p->weight_bytes);
    if ((reinterpret_cast<uintptr_t>(workspace) & 255u) || (reinterpret_cast<uintptr_t>(weight_arena) & 255u))
        return fail(RTOD_ERR_BAD_ARG, "rtod_plan_bind: arenas must be 256-byte aligned");
    p->ws = static_cast<unsigned char*>(workspace);
    p->wa = static_cast<unsigned char*>(weight_arena);
    p->err_flag = reinterpret_cast<int*>(p->ws);
    // padded weight rows / bias entries must read as zero
    RTOD_CUDA_OK(cudaMemset(p->wa, 0, p->weight_bytes));
    RTOD_CUDA_OK(cudaMemset(p->ws, 0, kArenaAlign));
    RTOD_CUDA_OK(cudaMemset(p->ws + p->split_off + kSplitScratchBytes, 0, kSplitCounters * sizeof(int)));
    RTOD_CUDA_OK(cudaDeviceSynchronize());         // bind is rare; later work may use any stream
    for (Node& nd : p->nodes) nd.weights_set = false;
    const int rc = bind_layers(*p);
    if (rc) return rc;
    p->bound = true;
    return RTOD_OK;
}

extern "C" int rtod_plan_set_conv_weights(RtodPlan* p, int layer, const float* weight, const float* bias,
                                          const float* bn_gamma, const float* bn_beta, const float* bn_mean,
                                          const float* bn_var, float bn_eps, void* stream) {
    if (!p || !p->bound) return fail(RTOD_ERR_STATE, "rtod_plan_set_conv_weights: plan is not bound");
    if (layer < 0 || layer >= (int)p->nodes.size() || p->nodes[layer].d.type != RTOD_LAYER_CONV)
        return fail(RTOD_ERR_BAD_ARG, "rtod_plan_set_conv_weights: layer %d is not a convolution", layer);
    if (!weight) return fail(RTOD_ERR_BAD_ARG, "rtod_plan_set_conv_weights: weight is null");
    const bool any_bn = bn_gamma || bn_beta || bn_mean || bn_var;
    if (any_bn && !(bn_gamma && bn_beta && bn_mean && bn_var))
        return fail(RTOD_ERR_BAD_ARG, "rtod_plan_set_conv_weights: all four BatchNorm tensors or none");
    Node& nd = p->nodes[layer];
    const int rc = launch_fold_pack(weight, bias, bn_gamma, bn_beta, bn_mean, bn_var, bn_eps, nd.d.filters,
                                    nd.Cin, nd.d.size, nd.Cout_pad, p->f16, nd.w_split, p->wa + nd.w_off,
                                    nd.stem ? reinterpret_cast<float*>(p->wa + nd.wf_off) : nullptr,
                                    reinterpret_cast<float*>(p->wa + nd.bias_off), (cudaStream_t)stream);
    if (rc) return rc;
    nd.weights_set = true;
    return RTOD_OK;
}

// one forward; when `ev` is non-null it holds n_layers + 2 events: ev[0] before the first launch,
// ev[i + 1] after layer i, ev[n + 1] after the decode launch
// `seg` (optional): events only where the stream switches between tcgen05 convolution launches (class 1) and
// everything else (class 0) -- per-launch events would put a ~2-5 us gap after every one of the 74 convolutions
struct Segments {
    std::vector<cudaEvent_t> ev;
    std::vector<int> cls;
};
static int mark_segment(Segments* seg, int cls, cudaStream_t stream) {
    if (!seg || (!seg->cls.empty() && seg->cls.back() == cls)) return RTOD_OK;
    cudaEvent_t e;
    RTOD_CUDA_OK(cudaEventCreate(&e));
    RTOD_CUDA_OK(cudaEventRecord(e, stream));
    seg->ev.push_back(e);
    seg->cls.push_back(cls);
    return RTOD_OK;
}

static int run_forward(RtodPlan* p, const float* x, float* pred, int train, cudaStream_t stream,
                       cudaEvent_t* ev, Segments* seg = nullptr, const unsigned char* x_u8 = nullptr) {
    if (!p || !p->bound) return fail(RTOD_ERR_STATE, "rtod_plan_forward: plan is not bound");
    if (!x && !x_u8) return fail(RTOD_ERR_BAD_ARG, "rtod_plan_forward: input is null");
    if (x_u8 && !(p->nodes.size() && p->nodes[0].stem && p->f16 && p->nodes[0].d.stride == 1 && p->nodes[0].d.pad == 1))
        return fail(RTOD_ERR_UNSUPPORTED, "rtod_plan_forward_u8: needs fp16 storage and a 3x3 / stride 1 stem convolution");
    if (p->heads.count && !pred) return fail(RTOD_ERR_BAD_ARG, "rtod_plan_forward: pred is null");
    const int n = (int)p->nodes.size();
    for (int i = 0; i < n; ++i)
        if (p->nodes[i].d.type == RTOD_LAYER_CONV && !p->nodes[i].weights_set)
            return fail(RTOD_ERR_STATE, "rtod_plan_forward: convolution %d has no weights", i);
    int rc;
    if (p->input_buf >= 0 && (rc = launch_nchw_to_nhwc(x, p->batch, act_of(*p, kInputLayer), stream))) return rc;
    if (ev) RTOD_CUDA_OK(cudaEventRecord(ev[0], stream));              // (behind the layout conversion of a non-stem input)
    for (int i = 0; i < n; ++i) {
        Node& nd = p->nodes[i];
        if (ev && i > 0) RTOD_CUDA_OK(cudaEventRecord(ev[i], stream));
        if (nd.alias_of >= -1) continue;
        const RtodLayerDesc& d = nd.d;
        if (d.type != RTOD_LAYER_ROUTE || nd.copy_concat)
            if ((rc = mark_segment(seg, d.type == RTOD_LAYER_CONV && !nd.stem && nd.use_tc ? 1 : 0, stream))) return rc;
        rc = RTOD_OK;
        switch (d.type) {
        case RTOD_LAYER_CONV:
            if (nd.stem && x_u8)
                rc = launch_stem_tc(nullptr, x_u8, p->batch, p->in_h, p->in_w, reinterpret_cast<const float*>(p->wa + nd.wf_off),
                                    nd.args.bias, d.filters, d.leaky, nd.args.out, stream);
            else if (nd.stem)
                rc = launch_stem_conv(x, p->batch, p->in_c, p->in_h, p->in_w,
                                      reinterpret_cast<const float*>(p->wa + nd.wf_off), nd.args.bias, d.filters,
                                      d.size, d.stride, d.pad, d.leaky, nd.args.out, stream);
            else if (nd.use_tc) rc = conv_tc_launch(nd.tc, stream);
            else rc = launch_conv_simt(nd.args, stream);
            break;
        case RTOD_LAYER_SHORTCUT:
            rc = launch_add(act_of(*p, d.src0), act_of(*p, d.src1), act_of(*p, i), p->batch, stream);
            break;
        case RTOD_LAYER_ROUTE:
            if (nd.copy_concat) {
                Act dst = act_of(*p, i);
                Act a = act_of(*p, d.src0), b = act_of(*p, d.src1);
                Act d0 = dst; d0.C = a.C;
                Act d1 = dst; d1.C = b.C;
                d1.ptr = static_cast<unsigned char*>(dst.ptr) + (size_t)a.C * 2;
                rc = launch_copy(a, d0, p->batch, stream);
                if (!rc) rc = launch_copy(b, d1, p->batch, stream);
            }
            break;
        case RTOD_LAYER_UPSAMPLE:
            rc = launch_upsample2x(act_of(*p, i - 1), act_of(*p, i), p->batch, stream);
            break;
        case RTOD_LAYER_MAXPOOL:
            rc = launch_maxpool(act_of(*p, i - 1), act_of(*p, i), p->batch, d.size, d.stride, stream);
            break;
        }
        if (rc) return rc;
    }
    if (ev) RTOD_CUDA_OK(cudaEventRecord(ev[n], stream));
    if (p->heads.count && (rc = mark_segment(seg, 0, stream))) return rc;
    if (p->heads.count) {
        p->heads.err_flag = p->err_flag;
        rc = launch_decode_heads(p->heads, p->batch, p->n_rows, p->n_attrs, train, pred, stream);
        if (rc) return rc;
    }
    if (ev) RTOD_CUDA_OK(cudaEventRecord(ev[n + 1], stream));
    if (seg && (rc = mark_segment(seg, -1, stream))) return rc;          // closing event
    return RTOD_OK;
}

extern "C" int rtod_plan_forward(RtodPlan* p, const float* x, float* pred, int train, void* stream) {
    return run_forward(p, x, pred, train, (cudaStream_t)stream, nullptr);
}

extern "C" int rtod_plan_forward_u8(RtodPlan* p, const unsigned char* x_planes, float* pred, int train, void* stream) {
    return run_forward(p, nullptr, pred, train, (cudaStream_t)stream, nullptr, nullptr, x_planes);
}

extern "C" int rtod_plan_forward_profile(RtodPlan* p, const float* x, float* pred, int train, void* stream_,
                                         float* layer_ms_host, int* layer_kind_host) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!p || !layer_ms_host) return fail(RTOD_ERR_BAD_ARG, "rtod_plan_forward_profile: null argument");
    const int n = (int)p->nodes.size();
    std::vector<cudaEvent_t> ev(n + 2);
    for (auto& e : ev) RTOD_CUDA_OK(cudaEventCreate(&e));
    int rc = run_forward(p, x, pred, train, stream, ev.data());
    if (!rc) {
        cudaError_t err = cudaStreamSynchronize(stream);
        if (err != cudaSuccess) rc = fail(RTOD_ERR_CUDA, "profile sync failed: %s", cudaGetErrorString(err));
    }
    for (int i = 0; i <= n && !rc; ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
        layer_ms_host[i] = ms;                               // [n] = decode launch
        if (layer_kind_host && i < n) {
            const Node& nd = p->nodes[i];
            // 0 = no kernel (alias / fused), 1 = tcgen05 conv, 2 = CUDA-core conv, 3 = stem, 4 = other
            layer_kind_host[i] = nd.alias_of >= -1 ? 0
                                 : nd.d.type != RTOD_LAYER_CONV ? (nd.d.type == RTOD_LAYER_ROUTE && !nd.copy_concat ? 0 : 4)
                                 : nd.stem ? 3 : (nd.use_tc ? 1 : 2);
        }
    }
    for (auto& e : ev) cudaEventDestroy(e);
    return rc;
}

extern "C" int rtod_plan_forward_segments(RtodPlan* p, const float* x, float* pred, int train, void* stream_,
                                          float* conv_ms, float* other_ms) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!p || !conv_ms || !other_ms) return fail(RTOD_ERR_BAD_ARG, "rtod_plan_forward_segments: null argument");
    Segments seg;
    int rc = run_forward(p, x, pred, train, stream, nullptr, &seg);
    if (!rc) {
        cudaError_t err = cudaStreamSynchronize(stream);
        if (err != cudaSuccess) rc = fail(RTOD_ERR_CUDA, "segment timing sync failed: %s", cudaGetErrorString(err));
    }
    *conv_ms = 0.f;
    *other_ms = 0.f;
    for (size_t i = 0; i + 1 < seg.ev.size() && !rc; ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, seg.ev[i], seg.ev[i + 1]);
        (seg.cls[i] == 1 ? *conv_ms : *other_ms) += ms;
    }
    for (auto& e : seg.ev) cudaEventDestroy(e);
    return rc;
}

extern "C" double rtod_plan_layer_flops(const RtodPlan* p, int layer) {
    if (!p || layer < 0 || layer >= (int)p->nodes.size() || p->nodes[layer].d.type != RTOD_LAYER_CONV) return 0.0;
    const Node& nd = p->nodes[layer];
    return 2.0 * (double)p->batch * nd.H * nd.W * nd.d.filters * nd.K;
}

extern "C" int rtod_plan_conv_backend(const RtodPlan* p, int layer) {
    if (!p || !p->bound || layer < 0 || layer >= (int)p->nodes.size() || p->nodes[layer].d.type != RTOD_LAYER_CONV)
        return RTOD_CONV_NONE;
    const Node& nd = p->nodes[layer];
    if (nd.stem) return RTOD_CONV_STEM;
    if (!nd.use_tc) return RTOD_CONV_SIMT;
    return nd.tc.patch == 2 ? RTOD_CONV_TC_PAIR : RTOD_CONV_TC;
}

extern "C" int rtod_plan_conv_config(const RtodPlan* p, int layer, int* out12) {
    if (!p || !p->bound || !out12 || layer < 0 || layer >= (int)p->nodes.size() || p->nodes[layer].d.type != RTOD_LAYER_CONV)
        return fail(RTOD_ERR_BAD_ARG, "rtod_plan_conv_config: bad plan, layer %d or output", layer);
    const Node& nd = p->nodes[layer];
    for (int i = 0; i < 12; ++i) out12[i] = 0;
    out12[0] = rtod_plan_conv_backend(p, layer);
    out12[10] = nd.w_split;
    if (!nd.use_tc) return RTOD_OK;
    const ConvTcChoice& c = nd.tc.choice;
    out12[1] = nd.tc.p.BN; out12[2] = c.ctas; out12[3] = nd.tc.p.b_resident; out12[4] = nd.tc.p.stage_bufs;
    out12[5] = nd.tc.p.split_k; out12[6] = nd.tc.p.epi_warps; out12[7] = nd.tc.p.a_producers; out12[8] = nd.tc.p.subs;
    out12[9] = nd.tc.p.stages; out12[11] = (int)nd.tc.grid.x;
    return RTOD_OK;
}

extern "C" int rtod_plan_read_layer(RtodPlan* p, int layer, float* out_nchw, void* stream) {
    if (!p || !p->bound) return fail(RTOD_ERR_STATE, "rtod_plan_read_layer: plan is not bound");
    if (layer < 0 || layer >= (int)p->nodes.size() || !out_nchw)
        return fail(RTOD_ERR_BAD_ARG, "rtod_plan_read_layer: bad layer %d or null output", layer);
    return launch_nhwc_to_nchw(act_of(*p, layer), p->batch, out_nchw, (cudaStream_t)stream);
}

extern "C" int rtod_plan_set_error_sink(RtodPlan* p, int* host_flag, void* stream) {
    if (!p || !p->bound) return fail(RTOD_ERR_STATE, "rtod_plan_set_error_sink: plan is not bound");
    int* dev_view = nullptr;
    if (host_flag) {
        void* d = nullptr;
        RTOD_CUDA_OK(cudaHostGetDevicePointer(&d, host_flag, 0));     // fails unless the memory is pinned and mapped
        dev_view = static_cast<int*>(d);
    }
    p->err_sink = dev_view;
    RTOD_CUDA_OK(cudaMemcpyAsync(p->err_flag + 2, &p->err_sink, sizeof(int*), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return RTOD_OK;
}

extern "C" int rtod_plan_reset_errors(RtodPlan* p, void* stream) {
    if (!p || !p->bound) return fail(RTOD_ERR_STATE, "rtod_plan_reset_errors: plan is not bound");
    RTOD_CUDA_OK(cudaMemsetAsync(p->err_flag, 0, sizeof(int), (cudaStream_t)stream));
    RTOD_CUDA_OK(cudaMemsetAsync(p->ws + p->split_off + kSplitScratchBytes, 0, kSplitCounters * sizeof(int), (cudaStream_t)stream));
    return RTOD_OK;
}

extern "C" int rtod_plan_is_f16(const RtodPlan* p) { return p ? p->f16 : 0; }
extern "C" int rtod_plan_conv_w_split(const RtodPlan* p, int layer) {
    if (!p || layer < 0 || layer >= (int)p->nodes.size() || p->nodes[layer].d.type != RTOD_LAYER_CONV) return 0;
    return p->nodes[layer].w_split;
}

extern "C" int rtod_plan_conv_row_mode(const RtodPlan* p, int layer) {
    if (!p || !p->bound || layer < 0 || layer >= (int)p->nodes.size() || p->nodes[layer].d.type != RTOD_LAYER_CONV) return 0;
    return p->nodes[layer].use_tc ? p->nodes[layer].tc.p.row_mode : 0;
}

extern "C" int rtod_plan_check(RtodPlan* p, void* stream) {
    if (!p || !p->bound) return fail(RTOD_ERR_STATE, "rtod_plan_check: plan is not bound");
    int flag = 0;
    RTOD_CUDA_OK(cudaMemcpyAsync(&flag, p->err_flag, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    RTOD_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    if (flag != 0) {
        rtod_plan_reset_errors(p, stream);
        return fail(RTOD_ERR_DEVICE, "device-side failure flag %d (tcgen05/TMA pipeline time-out)", flag);
    }
    return RTOD_OK;
}
