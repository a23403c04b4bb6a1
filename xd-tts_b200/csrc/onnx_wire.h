// Minimal reader of the protobuf wire format for the handful of ONNX ModelProto / GraphProto / NodeProto / TensorProto
// fields the library needs (no ONNX Runtime, no protobuf): shared by the postnet reader (onnx_postnet.cu) and the decoder
// reader (onnx_decoder.cu).  Host-only code.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace xdtts_onnx {

struct Span {
    const uint8_t* p;
    const uint8_t* end;
    bool ok = true;
    bool done() const { return p >= end; }
    uint64_t varint() {
        uint64_t v = 0;
        for (int shift = 0; shift < 64 && p < end; shift += 7) {
            const uint8_t b = *p++;
            v |= (uint64_t)(b & 0x7F) << shift;
            if (!(b & 0x80)) return v;
        }
        ok = false;
        return 0;
    }
    // reads one field header; returns false at the end or on a malformed stream
    bool field(uint32_t* number, uint32_t* wire) {
        if (done() || !ok) return false;
        const uint64_t key = varint();
        *number = (uint32_t)(key >> 3);
        *wire = (uint32_t)(key & 7);
        return ok;
    }
    Span bytes() {   // length-delimited payload
        const uint64_t n = varint();
        if (!ok || n > (uint64_t)(end - p)) {
            ok = false;
            return Span{end, end, false};
        }
        Span s{p, p + n};
        p += n;
        return s;
    }
    void skip(uint32_t wire) {
        switch (wire) {
            case 0: varint(); break;
            case 1: if (end - p >= 8) p += 8; else ok = false; break;
            case 2: bytes(); break;
            case 5: if (end - p >= 4) p += 4; else ok = false; break;
            default: ok = false;
        }
    }
    std::string str() {
        Span s = bytes();
        return s.ok ? std::string((const char*)s.p, (size_t)(s.end - s.p)) : std::string();
    }
};

struct Tensor {
    std::vector<int64_t> dims;
    int data_type = 0;   // 1 = FLOAT
    std::vector<float> data;
    bool external = false;
    size_t count() const {
        size_t n = 1;
        for (int64_t d : dims) n *= (size_t)d;
        return n;
    }
};

struct Node {
    std::string op;
    std::vector<std::string> in, out;
    float epsilon = 1e-5f;
    int64_t group = 1;
    std::vector<int64_t> pads, strides, dilations, kernel_shape;
    int64_t trans_a = 0, trans_b = 0, hidden_size = 0, axis = 0;
    float alpha = 1.f, beta = 1.f;
};

inline bool parse_tensor(Span s, std::string* name, Tensor* t) {
    uint32_t f, w;
    const uint8_t* raw = nullptr;
    size_t raw_n = 0;
    while (s.field(&f, &w)) {
        if (f == 1 && w == 0) t->dims.push_back((int64_t)s.varint());
        else if (f == 1 && w == 2) { Span d = s.bytes(); while (!d.done() && d.ok) t->dims.push_back((int64_t)d.varint()); }
        else if (f == 2 && w == 0) t->data_type = (int)s.varint();
        else if (f == 4 && w == 2) {   // packed float_data
            Span d = s.bytes();
            const size_t n = (size_t)(d.end - d.p) / 4;
            t->data.resize(n);
            if (n) memcpy(t->data.data(), d.p, n * 4);
        } else if (f == 4 && w == 5) { float v; if (s.end - s.p < 4) return false; memcpy(&v, s.p, 4); s.p += 4; t->data.push_back(v); }
        else if (f == 8 && w == 2) *name = s.str();
        else if (f == 9 && w == 2) { Span d = s.bytes(); raw = d.p; raw_n = (size_t)(d.end - d.p); }
        else if (f == 14 && w == 0) t->external = s.varint() == 1;   // data_location = EXTERNAL
        else s.skip(w);
    }
    if (!s.ok) return false;
    if (raw && t->data_type == 1) {
        t->data.resize(raw_n / 4);
        if (raw_n) memcpy(t->data.data(), raw, raw_n / 4 * 4);
    }
    return true;
}

inline bool parse_attribute(Span s, Node* n) {
    uint32_t f, w;
    std::string name;
    float fv = 0.f;
    int64_t iv = 0;
    std::vector<int64_t> ints;
    while (s.field(&f, &w)) {
        if (f == 1 && w == 2) name = s.str();
        else if (f == 2 && w == 5) { if (s.end - s.p < 4) return false; memcpy(&fv, s.p, 4); s.p += 4; }
        else if (f == 3 && w == 0) iv = (int64_t)s.varint();
        else if (f == 8 && w == 0) ints.push_back((int64_t)s.varint());
        else if (f == 8 && w == 2) { Span d = s.bytes(); while (!d.done() && d.ok) ints.push_back((int64_t)d.varint()); }
        else s.skip(w);
    }
    if (!s.ok) return false;
    if (name == "epsilon") n->epsilon = fv;
    else if (name == "group") n->group = iv;
    else if (name == "pads") n->pads = ints;
    else if (name == "strides") n->strides = ints;
    else if (name == "dilations") n->dilations = ints;
    else if (name == "kernel_shape") n->kernel_shape = ints;
    else if (name == "transA") n->trans_a = iv;
    else if (name == "transB") n->trans_b = iv;
    else if (name == "hidden_size") n->hidden_size = iv;
    else if (name == "axis") n->axis = iv;
    else if (name == "alpha") n->alpha = fv;
    else if (name == "beta") n->beta = fv;
    return true;
}

inline bool parse_node(Span s, Node* n) {
    uint32_t f, w;
    while (s.field(&f, &w)) {
        if (f == 1 && w == 2) n->in.push_back(s.str());
        else if (f == 2 && w == 2) n->out.push_back(s.str());
        else if (f == 4 && w == 2) n->op = s.str();
        else if (f == 5 && w == 2) { if (!parse_attribute(s.bytes(), n)) return false; }
        else s.skip(w);
    }
    return s.ok;
}


struct Graph {
    std::map<std::string, Tensor> init;
    std::vector<Node> nodes;
    std::vector<std::string> inputs, outputs;   // GraphProto.input / .output names (inputs may include initializers)
};

// 0 ok; otherwise a message in *err.  Reads the whole file.
inline bool load_graph(const char* path, Graph* g, std::string* err) {
    FILE* fp = fopen(path, "rb");
    if (!fp) { *err = std::string("cannot open ") + path; return false; }
    std::vector<uint8_t> buf;
    fseek(fp, 0, SEEK_END);
    const long size = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    if (size > 0) {
        buf.resize((size_t)size);
        if (fread(buf.data(), 1, (size_t)size, fp) != (size_t)size) buf.clear();
    }
    fclose(fp);
    if (buf.empty()) { *err = std::string(path) + " is empty or unreadable"; return false; }
    if (buf.size() < 200 && memcmp(buf.data(), "version https://git-lfs", 23) == 0) {
        *err = std::string(path) + " is a git-LFS pointer, not the model (run `git lfs pull`)";
        return false;
    }
    Span model{buf.data(), buf.data() + buf.size()};
    Span graph{nullptr, nullptr, false};
    uint32_t f, w;
    while (model.field(&f, &w)) {
        if (f == 7 && w == 2) graph = model.bytes();
        else model.skip(w);
    }
    if (!model.ok || !graph.ok || !graph.p) { *err = std::string(path) + " is not an ONNX ModelProto with a graph"; return false; }
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
            if (!parse_node(graph.bytes(), &n)) { *err = "malformed NodeProto"; return false; }
            g->nodes.push_back(std::move(n));
        } else if (f == 5 && w == 2) {
            std::string name;
            Tensor t;
            if (!parse_tensor(graph.bytes(), &name, &t)) { *err = "malformed TensorProto"; return false; }
            g->init[name] = std::move(t);
        } else if (f == 11 && w == 2) {
            g->inputs.push_back(value_name(graph.bytes()));
        } else if (f == 12 && w == 2) {
            g->outputs.push_back(value_name(graph.bytes()));
        } else {
            graph.skip(w);
        }
    }
    if (!graph.ok) { *err = "malformed GraphProto"; return false; }
    return true;
}

}  // namespace xdtts_onnx
