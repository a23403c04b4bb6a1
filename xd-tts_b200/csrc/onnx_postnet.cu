// Reads the weights of the Tacotron2 postnet out of `postnet.onnx` -- the file the reference opens
// with `Session::builder()?.commit_from_file(path.join("postnet.onnx"))`
// (/root/reference src/tacotron2/mod.rs:256-259; 17.4 MB git-LFS object, models/tacotron2/postnet.onnx:1-3)
// -- without ONNX Runtime or protobuf: a ~200-line reader of the protobuf wire format for the handful
// of ModelProto / GraphProto / NodeProto / TensorProto fields that matter.
//
// The graph is matched structurally, not by initializer names (exporters rename them), and it is VALIDATED: the
// nodes are walked along the dataflow from the graph input and must spell Conv [-> BatchNormalization] -> Tanh for
// every layer but the last, then the residual Add of the graph input -- the function the device kernels compute.
// Anything else (another activation, no residual, an extra operator) is rejected with XDTTS_ERR_UNSUPPORTED instead
// of loading into a network that computes something different.  Host-only code: usable (and tested) without a GPU.
// Read against two encoders: tests/onnx_writer.py (hand-rolled) and PyTorch's own serializer
// (tests/make_foreign_onnx.py -> tests/golden/postnet_torch_export_*.onnx).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "api_internal.h"
#include "onnx_wire.h"

using namespace xdtts;
#define fail xdtts::set_error

namespace {
using namespace xdtts_onnx;

struct Layer {
    int cout = 0, cin = 0, k = 0;
    std::vector<float> w, b, gamma, beta, mean, var;
    float eps = 1e-5f;
};

}  // namespace

struct xdtts_onnx_postnet {
    std::vector<Layer> layers;
    float eps = 1e-5f;
};

static bool all_equal(const std::vector<int64_t>& v, int64_t x) {
    for (int64_t a : v)
        if (a != x) return false;
    return true;
}

extern "C" void xdtts_onnx_postnet_close(xdtts_onnx_postnet* m) { delete m; }

extern "C" int xdtts_onnx_postnet_open(const char* path, xdtts_onnx_postnet** out) {
    if (!out) return fail(XDTTS_ERR_BAD_ARG, "onnx_postnet_open: out is null");
    *out = nullptr;
    if (!path) return fail(XDTTS_ERR_BAD_ARG, "onnx_postnet_open: path is null");
    FILE* fp = fopen(path, "rb");
    if (!fp) return fail(XDTTS_ERR_BAD_ARG, "onnx_postnet_open: cannot open %s", path);
    std::vector<uint8_t> buf;
    fseek(fp, 0, SEEK_END);
    const long size = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    if (size > 0) {
        buf.resize((size_t)size);
        if (fread(buf.data(), 1, (size_t)size, fp) != (size_t)size) buf.clear();
    }
    fclose(fp);
    if (buf.empty()) return fail(XDTTS_ERR_BAD_ARG, "onnx_postnet_open: %s is empty or unreadable", path);
    if (buf.size() < 200 && memcmp(buf.data(), "version https://git-lfs", 23) == 0)
        return fail(XDTTS_ERR_BAD_ARG, "onnx_postnet_open: %s is a git-LFS pointer, not the model (run `git lfs pull`)", path);

    // ModelProto: graph = field 7
    Span model{buf.data(), buf.data() + buf.size()};
    Span graph{nullptr, nullptr, false};
    uint32_t f, w;
    while (model.field(&f, &w)) {
        if (f == 7 && w == 2) graph = model.bytes();
        else model.skip(w);
    }
    if (!model.ok || !graph.ok || !graph.p) return fail(XDTTS_ERR_BAD_ARG, "onnx_postnet_open: %s is not an ONNX ModelProto with a graph", path);

    std::map<std::string, Tensor> init;
    std::vector<Node> nodes;
    std::vector<std::string> g_in, g_out;   // GraphProto.input / .output (ValueInfoProto.name)
    auto value_name = [](Span v) {
        std::string name;
        uint32_t vf, vw;
        while (v.field(&vf, &vw)) {
            if (vf == 1 && vw == 2) name = v.str();
            else v.skip(vw);
        }
        return name;
    };
    while (graph.field(&f, &w)) {
        if (f == 1 && w == 2) {
            Node n;
            if (!parse_node(graph.bytes(), &n)) return fail(XDTTS_ERR_BAD_ARG, "onnx_postnet_open: malformed NodeProto");
            nodes.push_back(std::move(n));
        } else if (f == 5 && w == 2) {
            std::string name;
            Tensor t;
            if (!parse_tensor(graph.bytes(), &name, &t)) return fail(XDTTS_ERR_BAD_ARG, "onnx_postnet_open: malformed TensorProto");
            init[name] = std::move(t);
        } else if (f == 11 && w == 2) {
            g_in.push_back(value_name(graph.bytes()));
        } else if (f == 12 && w == 2) {
            g_out.push_back(value_name(graph.bytes()));
        } else {
            graph.skip(w);
        }
    }
    if (!graph.ok) return fail(XDTTS_ERR_BAD_ARG, "onnx_postnet_open: malformed GraphProto");

    auto get = [&](const std::string& name, size_t want, const char* what, std::vector<float>* dst) -> int {
        auto it = init.find(name);
        if (it == init.end()) return fail(XDTTS_ERR_BAD_ARG, "onnx_postnet_open: %s '%s' is not an initializer", what, name.c_str());
        const Tensor& t = it->second;
        if (t.external) return fail(XDTTS_ERR_UNSUPPORTED, "onnx_postnet_open: %s '%s' uses external data", what, name.c_str());
        if (t.data_type != 1) return fail(XDTTS_ERR_UNSUPPORTED, "onnx_postnet_open: %s '%s' has data type %d, need FLOAT", what, name.c_str(), t.data_type);
        if (t.data.size() != want || t.count() != want)
            return fail(XDTTS_ERR_SHAPE, "onnx_postnet_open: %s '%s' has %zu values, expected %zu", what, name.c_str(), t.data.size(), want);
        *dst = t.data;
        return XDTTS_OK;
    };

    // The device path is fixed: out = x + L_{n-1}(tanh(L_{n-2}(... tanh(L_0(x))))), L = Conv [-> BatchNormalization].  The
    // graph must BE that -- walked along its dataflow from the (single) non-initializer input: Conv, then optionally the
    // BatchNormalization of its output, then Tanh after every layer but the last, a final Add of the graph input, and
    // nothing else but Identity / Dropout (inference: identity).  Any other operator, a missing activation or a missing
    // residual would load without complaint otherwise and silently produce a different mel_outputs_postnet.
    std::string x_name;
    for (const std::string& nm : g_in)
        if (init.find(nm) == init.end()) {   // old exporters list the initializers as graph inputs too
            if (!x_name.empty()) return fail(XDTTS_ERR_UNSUPPORTED, "onnx_postnet_open: more than one graph input ('%s', '%s')", x_name.c_str(), nm.c_str());
            x_name = nm;
        }
    if (x_name.empty()) return fail(XDTTS_ERR_BAD_ARG, "onnx_postnet_open: the graph has no input");
    xdtts_onnx_postnet* m = new xdtts_onnx_postnet();
    int rc = XDTTS_OK;
    std::string cur = x_name;          // tensor the walk stands on
    std::vector<bool> has_tanh;        // per layer
    bool bn_allowed = false, residual = false;
    for (size_t i = 0; i < nodes.size() && rc == XDTTS_OK; i++) {
        const Node& n = nodes[i];
        if (residual) { rc = fail(XDTTS_ERR_UNSUPPORTED, "onnx_postnet_open: operator %s after the residual Add", n.op.c_str()); break; }
        if (n.out.empty() || n.in.empty()) { rc = fail(XDTTS_ERR_BAD_ARG, "onnx_postnet_open: %s node without input or output", n.op.c_str()); break; }
        if (n.op == "Identity" || n.op == "Dropout") {
            if (n.in[0] != cur) { rc = fail(XDTTS_ERR_UNSUPPORTED, "onnx_postnet_open: %s node off the main path ('%s')", n.op.c_str(), n.in[0].c_str()); break; }
            cur = n.out[0];
            continue;
        }
        if (n.op == "Tanh") {
            if (n.in[0] != cur || m->layers.empty() || has_tanh.back()) { rc = fail(XDTTS_ERR_UNSUPPORTED, "onnx_postnet_open: Tanh that does not follow a convolution layer"); break; }
            has_tanh.back() = true;
            bn_allowed = false;
            cur = n.out[0];
            continue;
        }
        if (n.op == "Add") {
            const bool ok = n.in.size() == 2 && ((n.in[0] == x_name && n.in[1] == cur) || (n.in[1] == x_name && n.in[0] == cur));
            if (!ok || m->layers.empty() || has_tanh.back()) { rc = fail(XDTTS_ERR_UNSUPPORTED, "onnx_postnet_open: Add is not the residual (graph input + last layer)"); break; }
            residual = true;
            cur = n.out[0];
            continue;
        }
        if (n.op == "BatchNormalization") {
            if (n.in.size() < 5 || n.in[0] != cur || !bn_allowed) { rc = fail(XDTTS_ERR_UNSUPPORTED, "onnx_postnet_open: BatchNormalization that does not directly follow a Conv"); break; }
            Layer& L = m->layers.back();
            rc = get(n.in[1], (size_t)L.cout, "BatchNormalization scale", &L.gamma);
            if (rc == XDTTS_OK) rc = get(n.in[2], (size_t)L.cout, "BatchNormalization bias", &L.beta);
            if (rc == XDTTS_OK) rc = get(n.in[3], (size_t)L.cout, "BatchNormalization mean", &L.mean);
            if (rc == XDTTS_OK) rc = get(n.in[4], (size_t)L.cout, "BatchNormalization var", &L.var);
            L.eps = n.epsilon;
            bn_allowed = false;
            cur = n.out[0];
            continue;
        }
        if (n.op != "Conv") { rc = fail(XDTTS_ERR_UNSUPPORTED, "onnx_postnet_open: operator %s is not part of a Tacotron2 postnet (Conv, BatchNormalization, Tanh, Add, Identity, Dropout)", n.op.c_str()); break; }
        if (n.in[0] != cur) { rc = fail(XDTTS_ERR_UNSUPPORTED, "onnx_postnet_open: Conv reads '%s', not the previous layer's output", n.in[0].c_str()); break; }
        if (!m->layers.empty() && !has_tanh.back()) { rc = fail(XDTTS_ERR_UNSUPPORTED, "onnx_postnet_open: layer %zu is not followed by Tanh (only the last layer may be linear)", m->layers.size() - 1); break; }
        if (n.in.size() < 2) { rc = fail(XDTTS_ERR_BAD_ARG, "onnx_postnet_open: Conv node without weights"); break; }
        auto wt = init.find(n.in[1]);
        if (wt == init.end() || wt->second.dims.size() != 3) { rc = fail(XDTTS_ERR_UNSUPPORTED, "onnx_postnet_open: Conv weight '%s' is not a 3-D initializer (Conv1d expected)", n.in[1].c_str()); break; }
        Layer L;
        L.cout = (int)wt->second.dims[0]; L.cin = (int)wt->second.dims[1]; L.k = (int)wt->second.dims[2];
        if (n.group != 1 || (n.pads.empty() && L.k != 1) || !all_equal(n.strides, 1) || !all_equal(n.dilations, 1) || !all_equal(n.pads, L.k / 2) || (L.k & 1) == 0) {
            rc = fail(XDTTS_ERR_UNSUPPORTED, "onnx_postnet_open: Conv '%s' is not a stride-1, dilation-1, same-padded, ungrouped convolution", n.in[1].c_str());
            break;
        }
        rc = get(n.in[1], (size_t)L.cout * L.cin * L.k, "Conv weight", &L.w);
        if (rc == XDTTS_OK && n.in.size() > 2 && !n.in[2].empty()) rc = get(n.in[2], (size_t)L.cout, "Conv bias", &L.b);
        if (rc != XDTTS_OK) break;
        m->layers.push_back(std::move(L));
        has_tanh.push_back(false);
        bn_allowed = true;
        cur = n.out[0];
    }
    if (rc == XDTTS_OK && !m->layers.empty() && !residual)
        rc = fail(XDTTS_ERR_UNSUPPORTED, "onnx_postnet_open: %s has no residual Add of the graph input (mel_outputs_postnet = mel + postnet(mel))", path);
    if (rc == XDTTS_OK && !m->layers.empty() && (g_out.size() != 1 || g_out[0] != cur))
        rc = fail(XDTTS_ERR_UNSUPPORTED, "onnx_postnet_open: the graph output is not the residual sum");
    if (rc == XDTTS_OK && m->layers.empty()) rc = fail(XDTTS_ERR_BAD_ARG, "onnx_postnet_open: %s has no Conv node", path);
    for (size_t l = 1; rc == XDTTS_OK && l < m->layers.size(); l++) {
        if (m->layers[l].cin != m->layers[l - 1].cout) rc = fail(XDTTS_ERR_SHAPE, "onnx_postnet_open: layer %zu takes %d channels, layer %zu makes %d", l, m->layers[l].cin, l - 1, m->layers[l - 1].cout);
        if (m->layers[l].k != m->layers[0].k) rc = fail(XDTTS_ERR_UNSUPPORTED, "onnx_postnet_open: layers with different kernel sizes");
    }
    if (rc == XDTTS_OK) {
        bool have = false;
        for (const Layer& L : m->layers)
            if (!L.gamma.empty()) {
                if (have && L.eps != m->eps) rc = fail(XDTTS_ERR_UNSUPPORTED, "onnx_postnet_open: BatchNormalization nodes with different epsilon");
                m->eps = L.eps;
                have = true;
            }
    }
    if (rc != XDTTS_OK) {
        delete m;
        return rc;
    }
    *out = m;
    return XDTTS_OK;
}

extern "C" int xdtts_onnx_postnet_n_layers(const xdtts_onnx_postnet* m) {
    if (!m) return fail(XDTTS_ERR_BAD_ARG, "onnx_postnet_n_layers: null");
    return (int)m->layers.size();
}

extern "C" int xdtts_onnx_postnet_layer_info(const xdtts_onnx_postnet* m, int layer, int* cout, int* cin, int* ksize,
                                             int* has_bias, int* has_bn, float* eps) {
    if (!m || layer < 0 || layer >= (int)m->layers.size()) return fail(XDTTS_ERR_BAD_ARG, "onnx_postnet_layer_info: bad argument");
    const Layer& L = m->layers[layer];
    if (cout) *cout = L.cout;
    if (cin) *cin = L.cin;
    if (ksize) *ksize = L.k;
    if (has_bias) *has_bias = !L.b.empty();
    if (has_bn) *has_bn = !L.gamma.empty();
    if (eps) *eps = L.eps;
    return XDTTS_OK;
}

// which: 0 conv weight [cout, cin, k], 1 conv bias, 2 gamma, 3 beta, 4 running mean, 5 running var
extern "C" int xdtts_onnx_postnet_layer_copy(const xdtts_onnx_postnet* m, int layer, int which, float* out) {
    if (!m || !out || layer < 0 || layer >= (int)m->layers.size() || which < 0 || which > 5)
        return fail(XDTTS_ERR_BAD_ARG, "onnx_postnet_layer_copy: bad argument");
    const Layer& L = m->layers[layer];
    const std::vector<float>* src[6] = {&L.w, &L.b, &L.gamma, &L.beta, &L.mean, &L.var};
    if (src[which]->empty()) return fail(XDTTS_ERR_BAD_ARG, "onnx_postnet_layer_copy: layer %d has no tensor %d", layer, which);
    memcpy(out, src[which]->data(), src[which]->size() * 4);
    return XDTTS_OK;
}

// Tacotron2::load for the postnet session (src/tacotron2/mod.rs:256-259) straight onto the device
extern "C" int xdtts_postnet_create_from_onnx(const char* path, const xdtts_postnet_opts* opts, int device, xdtts_postnet** out) {
    if (!out) return fail(XDTTS_ERR_BAD_ARG, "postnet_create_from_onnx: out is null");
    *out = nullptr;
    xdtts_onnx_postnet* m = nullptr;
    int rc = xdtts_onnx_postnet_open(path, &m);
    if (rc) return rc;
    const int n = (int)m->layers.size();
    std::vector<int> ch(n + 1);
    std::vector<const float*> w(n), b(n), g(n), be(n), mu(n), var(n);
    ch[0] = m->layers[0].cin;
    for (int l = 0; l < n; l++) {
        const Layer& L = m->layers[l];
        ch[l + 1] = L.cout;
        w[l] = L.w.data();
        b[l] = L.b.empty() ? nullptr : L.b.data();
        g[l] = L.gamma.empty() ? nullptr : L.gamma.data();
        be[l] = L.beta.empty() ? nullptr : L.beta.data();
        mu[l] = L.mean.empty() ? nullptr : L.mean.data();
        var[l] = L.var.empty() ? nullptr : L.var.data();
    }
    rc = xdtts_postnet_create(n, ch.data(), m->layers[0].k, w.data(), b.data(), g.data(), be.data(), mu.data(), var.data(),
                              m->eps, opts, device, out);
    delete m;
    return rc;
}
