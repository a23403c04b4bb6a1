// .npy reader / writer for 2-D float32 arrays: the wire format of the reference's spectrogram dump
// (`ndarray_npy::write_npy(path, &spectrogram)`, /root/reference src/lib.rs:125-139, CLI flag
// --output-spectrogram src/bin/app.rs:12-14) -- the hook for feeding real Tacotron2 mels to this
// library and diffing results as files.  Host-only code.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "api_internal.h"

using namespace xdtts;
#define fail xdtts::set_error

extern "C" int xdtts_npy_write_f32(const char* path, const float* data, int rows, int cols) {
    if (!path || !data || rows < 0 || cols < 0) return fail(XDTTS_ERR_BAD_ARG, "npy_write: bad argument");
    char dict[128];
    int n = snprintf(dict, sizeof(dict), "{'descr': '<f4', 'fortran_order': False, 'shape': (%d, %d), }", rows, cols);
    std::string header(dict, (size_t)n);
    // magic (6) + version (2) + header length (2) + header, padded with spaces to a multiple of 64, '\n' last
    size_t total = 10 + header.size() + 1;
    total = (total + 63) / 64 * 64;
    header.append(total - 10 - header.size() - 1, ' ');
    header.push_back('\n');
    FILE* fp = fopen(path, "wb");
    if (!fp) return fail(XDTTS_ERR_BAD_ARG, "npy_write: cannot create %s", path);
    const unsigned char magic[8] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0};
    const uint16_t hlen = (uint16_t)header.size();
    const unsigned char len2[2] = {(unsigned char)(hlen & 0xFF), (unsigned char)(hlen >> 8)};
    const size_t count = (size_t)rows * (size_t)cols;
    bool ok = fwrite(magic, 1, 8, fp) == 8 && fwrite(len2, 1, 2, fp) == 2 && fwrite(header.data(), 1, header.size(), fp) == header.size() &&
              (count == 0 || fwrite(data, 4, count, fp) == count);
    ok = (fclose(fp) == 0) && ok;
    return ok ? XDTTS_OK : fail(XDTTS_ERR_BAD_ARG, "npy_write: short write to %s", path);
}

// Two-call protocol: out == null returns the shape only; otherwise capacity (floats) must hold rows*cols.
// A 1-D array of n elements reads as [1, n]; Fortran-ordered files are transposed into C order.
extern "C" int xdtts_npy_read_f32(const char* path, float* out, long long capacity, int* rows, int* cols) {
    if (!path || !rows || !cols) return fail(XDTTS_ERR_BAD_ARG, "npy_read: bad argument");
    FILE* fp = fopen(path, "rb");
    if (!fp) return fail(XDTTS_ERR_BAD_ARG, "npy_read: cannot open %s", path);
    unsigned char pre[12];
    size_t hlen = 0, hoff = 0;
    if (fread(pre, 1, 10, fp) != 10 || memcmp(pre, "\x93NUMPY", 6) != 0) { fclose(fp); return fail(XDTTS_ERR_BAD_ARG, "npy_read: %s is not a .npy file", path); }
    if (pre[6] == 1) { hlen = pre[8] | (pre[9] << 8); hoff = 10; }
    else if (pre[6] == 2 || pre[6] == 3) {
        if (fread(pre + 10, 1, 2, fp) != 2) { fclose(fp); return fail(XDTTS_ERR_BAD_ARG, "npy_read: truncated header"); }
        hlen = (size_t)pre[8] | ((size_t)pre[9] << 8) | ((size_t)pre[10] << 16) | ((size_t)pre[11] << 24); hoff = 12;
    } else { fclose(fp); return fail(XDTTS_ERR_UNSUPPORTED, "npy_read: format version %d", pre[6]); }
    if (hlen > (1u << 20)) { fclose(fp); return fail(XDTTS_ERR_BAD_ARG, "npy_read: implausible header length"); }
    std::string h(hlen, '\0');
    if (fread(&h[0], 1, hlen, fp) != hlen) { fclose(fp); return fail(XDTTS_ERR_BAD_ARG, "npy_read: truncated header"); }
    (void)hoff;
    auto value_of = [&](const char* key) -> std::string {
        size_t p = h.find(key);
        if (p == std::string::npos) return "";
        p = h.find(':', p);
        if (p == std::string::npos) return "";
        p++;
        while (p < h.size() && h[p] == ' ') p++;
        size_t e = p;
        if (h[p] == '(') e = h.find(')', p) + 1;
        else while (e < h.size() && h[e] != ',' && h[e] != '}') e++;
        return h.substr(p, e - p);
    };
    const std::string descr = value_of("'descr'"), order = value_of("'fortran_order'"), shape = value_of("'shape'");
    if (descr.find("<f4") == std::string::npos && descr.find("|f4") == std::string::npos && descr.find("=f4") == std::string::npos) {
        fclose(fp);
        return fail(XDTTS_ERR_UNSUPPORTED, "npy_read: dtype %s, need little-endian float32", descr.c_str());
    }
    const bool fortran = order.find("True") != std::string::npos;
    std::vector<long long> dims;
    for (size_t i = 0; i < shape.size();) {
        if (shape[i] >= '0' && shape[i] <= '9') {
            char* end = nullptr;
            dims.push_back(strtoll(shape.c_str() + i, &end, 10));
            i = (size_t)(end - shape.c_str());
        } else i++;
    }
    if (dims.empty() || dims.size() > 2) { fclose(fp); return fail(XDTTS_ERR_SHAPE, "npy_read: %zu-D array, need 1-D or 2-D", dims.size()); }
    const long long r = dims.size() == 2 ? dims[0] : 1, c = dims.size() == 2 ? dims[1] : dims[0];
    if (r > 0x7fffffffLL || c > 0x7fffffffLL) { fclose(fp); return fail(XDTTS_ERR_SHAPE, "npy_read: array too large"); }
    *rows = (int)r;
    *cols = (int)c;
    if (!out) { fclose(fp); return XDTTS_OK; }
    if (capacity < r * c) { fclose(fp); return fail(XDTTS_ERR_SHAPE, "npy_read: buffer holds %lld floats, the array has %lld", capacity, r * c); }
    const size_t count = (size_t)(r * c);
    int rc = XDTTS_OK;
    if (!fortran) {
        if (count && fread(out, 4, count, fp) != count) rc = fail(XDTTS_ERR_BAD_ARG, "npy_read: truncated data in %s", path);
    } else {
        std::vector<float> tmp(count);
        if (count && fread(tmp.data(), 4, count, fp) != count) rc = fail(XDTTS_ERR_BAD_ARG, "npy_read: truncated data in %s", path);
        else
            for (long long i = 0; i < r; i++)
                for (long long j = 0; j < c; j++) out[i * c + j] = tmp[(size_t)(j * r + i)];
    }
    fclose(fp);
    return rc;
}
