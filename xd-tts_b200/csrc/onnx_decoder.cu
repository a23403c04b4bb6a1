// Reads the weights of the Tacotron2 decoder step out of `decoder_iter.onnx` -- the file the reference opens with
// `commit_from_file(path.join("decoder_iter.onnx"))` (/root/reference src/tacotron2/mod.rs:251-254; 72.8 MB git-LFS
// object) and runs once per output frame (:302-342) -- and hands them to the persistent decoder kernel in the layouts of
// `xdtts_decoder_weights` (PyTorch layouts, LSTM gate order i, f, g, o).
//
// The graph is NVIDIA Tacotron2's `Decoder.decode` wrapped by `DecoderIter` of export_tacotron2_onnx.py (named at
// src/tacotron2/mod.rs:137-138).  Its tensors are found by ROLE, not by name (exporters rename folded constants to
// onnx::MatMul_123): the graph inputs and outputs are the names the reference feeds and reads (src/tacotron2/mod.rs:285-341),
// and every weight is identified by the dataflow between them:
//     prenet 1, 2          the two constant-weight products on the path from "decoder_input"
//     attention LSTM       the LSTM (or its Gemm/Gemm/Add/Split decomposition) whose state is "attention_hidden"
//     query / location / v the constant-weight products inside the attention, in graph order, + the one Conv
//     decoder LSTM         ... whose state is "decoder_hidden"
//     projection, gate     the products that write "decoder_output" / "gate_prediction"
// Both encodings of an LSTM cell are read: the ONNX `LSTM` operator NVIDIA's script exports (W/R/B with gate order
// i, o, f, c -- silently wrong if taken as PyTorch's i, f, g, o) and the Gemm + Gemm + Add + Split decomposition of today's
// PyTorch exporter.  Every shape is checked against its neighbours; anything that does not fit is XDTTS_ERR_UNSUPPORTED.
// Read against a foreign encoder: tests/make_foreign_decoder_onnx.py (PyTorch's own serializer, both encodings).
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "api_internal.h"
#include "onnx_wire.h"

using namespace xdtts;
using namespace xdtts_onnx;
#define fail xdtts::set_error

struct xdtts_onnx_decoder {
    int n_mel = 0, prenet = 0, enc = 0, att_rnn = 0, dec_rnn = 0, att_dim = 0, loc_f = 0, loc_k = 0, lstm_form = 0, has_dropout = 0;
    std::vector<float> t[18];   // in the field order of xdtts_decoder_weights
};

namespace {

enum { W_PRENET1, W_PRENET2, W_ATT_IH, W_ATT_HH, W_ATT_BIH, W_ATT_BHH, W_QUERY, W_V, W_LOC_CONV, W_LOC_DENSE, W_DEC_IH, W_DEC_HH,
       W_DEC_BIH, W_DEC_BHH, W_PROJ_W, W_PROJ_B, W_GATE_W, W_GATE_B };

struct Linear {       // y = x W^T (+ b): W stored [out][in]
    int node = -1, out = 0, in = 0;
    std::vector<float> w, b;
    std::string x_name, y_name;
};

const Tensor* const_of(const Graph& g, const std::string& name) {
    auto it = g.init.find(name);
    if (it == g.init.end() || it->second.external || it->second.data_type != 1 || it->second.data.size() != it->second.count()) return nullptr;
    return &it->second;
}

// MatMul(x, B) with constant B [in, out], or Gemm(x, B, C) with constant B ([out, in] when transB) -> Linear
bool as_linear(const Graph& g, int idx, Linear* L) {
    const Node& n = g.nodes[idx];
    if (n.in.size() < 2 || n.out.empty()) return false;
    const Tensor* B = const_of(g, n.in[1]);
    if (!B || B->dims.size() != 2) return false;
    if (n.op == "MatMul") {
        L->in = (int)B->dims[0]; L->out = (int)B->dims[1];
        L->w.resize(B->data.size());
        for (int o = 0; o < L->out; o++)
            for (int i = 0; i < L->in; i++) L->w[(size_t)o * L->in + i] = B->data[(size_t)i * L->out + o];
    } else if (n.op == "Gemm") {
        if (n.trans_a != 0 || n.alpha != 1.f) return false;
        if (n.trans_b) { L->out = (int)B->dims[0]; L->in = (int)B->dims[1]; L->w = B->data; }
        else {
            L->in = (int)B->dims[0]; L->out = (int)B->dims[1];
            L->w.resize(B->data.size());
            for (int o = 0; o < L->out; o++)
                for (int i = 0; i < L->in; i++) L->w[(size_t)o * L->in + i] = B->data[(size_t)i * L->out + o];
        }
        if (n.in.size() > 2 && !n.in[2].empty()) {
            const Tensor* C = const_of(g, n.in[2]);
            if (!C || (int)C->count() != L->out || n.beta != 1.f) return false;
            L->b = C->data;
        }
    } else {
        return false;
    }
    L->node = idx; L->x_name = n.in[0]; L->y_name = n.out[0];
    return true;
}

// ONNX gate order i, o, f, c (rows of W / R, entries of the two halves of B) -> PyTorch's i, f, g, o
void reorder_iofc(const float* src, int H, int cols, float* dst) {
    const int from[4] = {0, 2, 3, 1};   // torch block q comes from onnx block from[q]
    for (int q = 0; q < 4; q++) memcpy(dst + (size_t)q * H * cols, src + (size_t)from[q] * H * cols, sizeof(float) * (size_t)H * cols);
}

}  // namespace

extern "C" void xdtts_onnx_decoder_close(xdtts_onnx_decoder* m) { delete m; }

extern "C" int xdtts_onnx_decoder_open(const char* path, xdtts_onnx_decoder** out) {
    if (!out) return fail(XDTTS_ERR_BAD_ARG, "onnx_decoder_open: out is null");
    *out = nullptr;
    if (!path) return fail(XDTTS_ERR_BAD_ARG, "onnx_decoder_open: path is null");
    Graph g;
    std::string err;
    if (!load_graph(path, &g, &err)) return fail(XDTTS_ERR_BAD_ARG, "onnx_decoder_open: %s", err.c_str());

    // the contract with the caller is by NAME (src/tacotron2/mod.rs:285-341)
    static const char* kIn[] = {"decoder_input", "attention_hidden", "attention_cell", "decoder_hidden", "decoder_cell", "attention_weights",
                                "attention_weights_cum", "attention_context", "memory", "processed_memory", "mask"};
    static const char* kOut[] = {"decoder_output", "gate_prediction", "out_attention_hidden", "out_attention_cell", "out_decoder_hidden",
                                 "out_decoder_cell", "out_attention_weights", "out_attention_weights_cum", "out_attention_context"};
    for (const char* nm : kIn)
        if (std::find(g.inputs.begin(), g.inputs.end(), nm) == g.inputs.end())
            return fail(XDTTS_ERR_UNSUPPORTED, "onnx_decoder_open: the graph has no input '%s' (not a Tacotron2 decoder_iter.onnx)", nm);
    for (const char* nm : kOut)
        if (std::find(g.outputs.begin(), g.outputs.end(), nm) == g.outputs.end())
            return fail(XDTTS_ERR_UNSUPPORTED, "onnx_decoder_open: the graph has no output '%s' (not a Tacotron2 decoder_iter.onnx)", nm);

    // tensor -> the tensor it is a reshaped copy of (Unsqueeze / Squeeze / Identity / Cast / Reshape of one input)
    std::map<std::string, std::string> alias;
    for (const Node& n : g.nodes)
        if ((n.op == "Unsqueeze" || n.op == "Squeeze" || n.op == "Identity" || n.op == "Cast" || n.op == "Reshape") && !n.in.empty() && !n.out.empty())
            alias[n.out[0]] = n.in[0];
    auto root = [&](std::string s) {
        for (int hop = 0; hop < 16; hop++) {
            auto it = alias.find(s);
            if (it == alias.end()) break;
            s = it->second;
        }
        return s;
    };
    auto producer = [&](const std::string& t) -> int {
        for (size_t i = 0; i < g.nodes.size(); i++)
            for (const std::string& o : g.nodes[i].out)
                if (o == t) return (int)i;
        return -1;
    };

    xdtts_onnx_decoder* m = new xdtts_onnx_decoder();
    auto bail = [&](int code, const char* msg) {
        delete m;
        return fail(code, "onnx_decoder_open: %s", msg);
    };
    std::vector<bool> used(g.nodes.size(), false);

    // ---- the two LSTM cells
    struct Cell { int H = 0, in = 0; std::vector<float> w_ih, w_hh, b_ih, b_hh; bool found = false; };
    Cell cells[2];
    const char* state_name[2] = {"attention_hidden", "decoder_hidden"};
    for (size_t i = 0; i < g.nodes.size(); i++) {
        const Node& n = g.nodes[i];
        if (n.op != "LSTM") continue;
        if (n.in.size() < 6) return bail(XDTTS_ERR_UNSUPPORTED, "LSTM operator without initial state");
        const std::string h0 = root(n.in[5]);
        int which = h0 == state_name[0] ? 0 : (h0 == state_name[1] ? 1 : -1);
        if (which < 0 || cells[which].found) return bail(XDTTS_ERR_UNSUPPORTED, "LSTM operator whose state is neither attention_hidden nor decoder_hidden");
        const Tensor *W = const_of(g, n.in[1]), *R = const_of(g, n.in[2]), *B = n.in[3].empty() ? nullptr : const_of(g, n.in[3]);
        if (!W || !R || W->dims.size() != 3 || R->dims.size() != 3 || W->dims[0] != 1 || R->dims[0] != 1 || W->dims[1] % 4 || R->dims[1] != W->dims[1] ||
            R->dims[2] * 4 != R->dims[1])
            return bail(XDTTS_ERR_UNSUPPORTED, "LSTM operator with unexpected W / R shapes (one direction, one layer expected)");
        Cell& c = cells[which];
        c.H = (int)R->dims[2]; c.in = (int)W->dims[2];
        c.w_ih.resize(W->data.size()); c.w_hh.resize(R->data.size());
        reorder_iofc(W->data.data(), c.H, c.in, c.w_ih.data());
        reorder_iofc(R->data.data(), c.H, c.H, c.w_hh.data());
        c.b_ih.assign((size_t)4 * c.H, 0.f); c.b_hh.assign((size_t)4 * c.H, 0.f);
        if (B) {
            if ((int)B->count() != 8 * c.H) return bail(XDTTS_ERR_UNSUPPORTED, "LSTM operator with a bias that is not [1, 8 H]");
            reorder_iofc(B->data.data(), c.H, 1, c.b_ih.data());
            reorder_iofc(B->data.data() + 4 * c.H, c.H, 1, c.b_hh.data());
        }
        c.found = true;
        used[i] = true;
        m->lstm_form = 1;
    }
    for (int which = 0; which < 2; which++) {   // the decomposition: Gemm(state, W_hh, b_hh) + Gemm(input, W_ih, b_ih) -> Add -> Split(4)
        if (cells[which].found) continue;
        int hh = -1;
        Linear Lhh, Lih;
        for (size_t i = 0; i < g.nodes.size() && hh < 0; i++)
            if (g.nodes[i].op == "Gemm" && !g.nodes[i].in.empty() && root(g.nodes[i].in[0]) == state_name[which] && as_linear(g, (int)i, &Lhh)) hh = (int)i;
        if (hh < 0) return bail(XDTTS_ERR_UNSUPPORTED, which == 0 ? "no LSTM (operator or Gemm decomposition) on attention_hidden" : "no LSTM (operator or Gemm decomposition) on decoder_hidden");
        int ih = -1, add = -1;
        for (size_t i = 0; i < g.nodes.size() && ih < 0; i++) {
            const Node& a = g.nodes[i];
            if (a.op != "Add" || a.in.size() != 2) continue;
            const int other = a.in[0] == Lhh.y_name ? 1 : (a.in[1] == Lhh.y_name ? 0 : -1);
            if (other < 0) continue;
            const int pi = producer(a.in[other]);
            if (pi >= 0 && as_linear(g, pi, &Lih)) { ih = pi; add = (int)i; }
        }
        if (ih < 0) return bail(XDTTS_ERR_UNSUPPORTED, "LSTM decomposition: the input-to-hidden Gemm was not found");
        bool split4 = false;
        for (const Node& s : g.nodes)
            if (s.op == "Split" && !s.in.empty() && s.in[0] == g.nodes[add].out[0] && s.out.size() == 4) split4 = true;
        if (!split4 || Lhh.out != Lih.out || Lhh.out % 4 || Lhh.in * 4 != Lhh.out)
            return bail(XDTTS_ERR_UNSUPPORTED, "LSTM decomposition with unexpected shapes (gates must be one Split into 4)");
        Cell& c = cells[which];
        c.H = Lhh.in; c.in = Lih.in;
        c.w_ih = Lih.w; c.w_hh = Lhh.w;     // the decomposition keeps PyTorch's gate order: Split -> sigmoid, sigmoid, tanh, sigmoid
        c.b_ih = Lih.b.empty() ? std::vector<float>((size_t)4 * c.H, 0.f) : Lih.b;
        c.b_hh = Lhh.b.empty() ? std::vector<float>((size_t)4 * c.H, 0.f) : Lhh.b;
        c.found = true;
        used[hh] = used[ih] = true;
    }

    // ---- the location convolution
    int n_conv = 0;
    for (size_t i = 0; i < g.nodes.size(); i++) {
        const Node& n = g.nodes[i];
        if (n.op != "Conv") continue;
        const Tensor* W = n.in.size() >= 2 ? const_of(g, n.in[1]) : nullptr;
        if (!W || W->dims.size() != 3 || W->dims[1] != 2 || (W->dims[2] & 1) == 0 || (n.in.size() > 2 && !n.in[2].empty()))
            return bail(XDTTS_ERR_UNSUPPORTED, "Conv that is not the bias-free [filters, 2, odd k] location convolution");
        for (int64_t p_ : n.pads)
            if (p_ != W->dims[2] / 2) return bail(XDTTS_ERR_UNSUPPORTED, "location convolution is not same-padded");
        m->loc_f = (int)W->dims[0]; m->loc_k = (int)W->dims[2];
        m->t[W_LOC_CONV] = W->data;
        n_conv++;
    }
    if (n_conv != 1) return bail(XDTTS_ERR_UNSUPPORTED, "expected exactly one Conv (the location layer)");

    // ---- the remaining constant-weight products, in graph order: prenet 1, prenet 2, query, location dense, v, then the
    // two that write the named outputs
    std::vector<Linear> lin;
    for (size_t i = 0; i < g.nodes.size(); i++) {
        if (used[i]) continue;
        Linear L;
        if (as_linear(g, (int)i, &L)) lin.push_back(std::move(L));
    }
    int i_proj = -1, i_gate = -1;
    for (size_t i = 0; i < lin.size(); i++) {
        std::string y = lin[i].y_name;
        for (const Node& a : g.nodes)   // MatMul + Add(bias) form
            if (a.op == "Add" && a.in.size() == 2 && (a.in[0] == y || a.in[1] == y)) {
                const Tensor* b = const_of(g, a.in[0] == y ? a.in[1] : a.in[0]);
                if (b && (int)b->count() == lin[i].out && lin[i].b.empty() && (a.out[0] == "decoder_output" || a.out[0] == "gate_prediction")) {
                    lin[i].b = b->data;
                    y = a.out[0];
                }
            }
        if (y == "decoder_output") i_proj = (int)i;
        if (y == "gate_prediction") i_gate = (int)i;
    }
    if (i_proj < 0 || i_gate < 0) return bail(XDTTS_ERR_UNSUPPORTED, "no constant-weight product writes decoder_output / gate_prediction");
    std::vector<Linear*> rest;
    for (size_t i = 0; i < lin.size(); i++)
        if ((int)i != i_proj && (int)i != i_gate) rest.push_back(&lin[i]);
    if (rest.size() != 5) return bail(XDTTS_ERR_UNSUPPORTED, "expected five more constant-weight products (prenet x2, query, location dense, v)");
    Linear &p1 = *rest[0], &p2 = *rest[1], &q = *rest[2], &ld = *rest[3], &v = *rest[4], &pr = lin[i_proj], &gt = lin[i_gate];
    const Cell &ca = cells[0], &cd = cells[1];
    m->n_mel = p1.in; m->prenet = p1.out; m->att_rnn = ca.H; m->dec_rnn = cd.H; m->att_dim = q.out;
    m->enc = ca.in - m->prenet;
    const bool ok = root(p1.x_name) == "decoder_input" && p2.in == m->prenet && p2.out == m->prenet && m->enc > 0 && q.in == m->att_rnn &&
                    ld.in == m->loc_f && ld.out == m->att_dim && v.in == m->att_dim && v.out == 1 && cd.in == m->att_rnn + m->enc &&
                    pr.in == m->dec_rnn + m->enc && pr.out == m->n_mel && gt.in == pr.in && gt.out == 1 && p1.b.empty() && p2.b.empty() &&
                    q.b.empty() && ld.b.empty() && v.b.empty();
    if (!ok) return bail(XDTTS_ERR_UNSUPPORTED, "the weight shapes do not chain like a Tacotron2 decoder step (prenet -> attention LSTM -> attention -> decoder LSTM -> projection, gate)");
    for (const Node& n : g.nodes)
        if (n.op == "RandomUniformLike" || n.op == "RandomUniform" || n.op == "Bernoulli") m->has_dropout = 1;
    m->t[W_PRENET1] = p1.w; m->t[W_PRENET2] = p2.w;
    m->t[W_ATT_IH] = ca.w_ih; m->t[W_ATT_HH] = ca.w_hh; m->t[W_ATT_BIH] = ca.b_ih; m->t[W_ATT_BHH] = ca.b_hh;
    m->t[W_QUERY] = q.w; m->t[W_V] = v.w; m->t[W_LOC_DENSE] = ld.w;
    m->t[W_DEC_IH] = cd.w_ih; m->t[W_DEC_HH] = cd.w_hh; m->t[W_DEC_BIH] = cd.b_ih; m->t[W_DEC_BHH] = cd.b_hh;
    m->t[W_PROJ_W] = pr.w; m->t[W_PROJ_B] = pr.b.empty() ? std::vector<float>((size_t)pr.out, 0.f) : pr.b;
    m->t[W_GATE_W] = gt.w; m->t[W_GATE_B] = gt.b.empty() ? std::vector<float>(1, 0.f) : gt.b;
    *out = m;
    return XDTTS_OK;
}

extern "C" int xdtts_onnx_decoder_dims(const xdtts_onnx_decoder* m, int* dims10) {
    if (!m || !dims10) return fail(XDTTS_ERR_BAD_ARG, "onnx_decoder_dims: null argument");
    const int d[10] = {m->n_mel, m->prenet, m->enc, m->att_rnn, m->dec_rnn, m->att_dim, m->loc_f, m->loc_k, m->lstm_form, m->has_dropout};
    memcpy(dims10, d, sizeof(d));
    return XDTTS_OK;
}

extern "C" long long xdtts_onnx_decoder_tensor(const xdtts_onnx_decoder* m, int which, float* out, long long capacity) {
    if (!m || which < 0 || which > 17) return fail(XDTTS_ERR_BAD_ARG, "onnx_decoder_tensor: bad argument");
    const long long n = (long long)m->t[which].size();
    if (!out) return n;
    if (capacity < n) return fail(XDTTS_ERR_SHAPE, "onnx_decoder_tensor: tensor %d has %lld values, buffer holds %lld", which, n, capacity);
    memcpy(out, m->t[which].data(), (size_t)n * 4);
    return n;
}

// Tacotron2::load for the decoder session (src/tacotron2/mod.rs:251-254) straight onto the device
extern "C" int xdtts_decoder_create_from_onnx(const char* path, const xdtts_decoder_opts* opts, int device, xdtts_decoder** out) {
    if (!out) return fail(XDTTS_ERR_BAD_ARG, "decoder_create_from_onnx: out is null");
    *out = nullptr;
    xdtts_onnx_decoder* m = nullptr;
    int rc = xdtts_onnx_decoder_open(path, &m);
    if (rc) return rc;
    if (m->n_mel != 80 || m->prenet != 256 || m->enc != 512 || m->att_rnn != 1024 || m->dec_rnn != 1024 || m->att_dim != 128 || m->loc_f != 32 ||
        m->loc_k != 31) {
        rc = fail(XDTTS_ERR_UNSUPPORTED,
                  "decoder_create_from_onnx: dimensions (mel %d, prenet %d, encoder %d, attention rnn %d, decoder rnn %d, attention %d, location %d x %d) "
                  "are not Tacotron2's (80, 256, 512, 1024, 1024, 128, 32 x 31) the device kernel is built for",
                  m->n_mel, m->prenet, m->enc, m->att_rnn, m->dec_rnn, m->att_dim, m->loc_f, m->loc_k);
        delete m;
        return rc;
    }
    xdtts_decoder_weights w;
    const float** fields = reinterpret_cast<const float**>(&w);
    for (int i = 0; i < 18; i++) fields[i] = m->t[i].data();
    rc = xdtts_decoder_create(&w, opts, device, out);
    delete m;
    return rc;
}
