// bin_loader.cu -- C-level reader of libwavernn `.bin` checkpoints: the equivalent of WaveRNNVocoder::loadWeights(path)
// (vocoder/libwavernn/fatchord_version/src/WaveRNNVocoder.cpp:22-31) for hosts that bind the C ABI without Python.
// Host code only.  Wire format (vocoder/libwavernn/convert.py:55 file header 4 x int32 = res_blocks, #upsample layers, total
// scale, pad; :170-175 layer header int32 type + 64-byte name; payloads by type -- Conv1d :98-108, Conv2d :110-119, BatchNorm1d
// :121-133, Linear :87-96, GRU :135-162, Stretch2d :164-167; compressed matrix :61-84 / wavernn.h:23-92: int32 nW, float[nW] kept
// 1x4 groups row-major, int32 nIdx, uint8[nIdx] group columns per row, 255 ends a row).  Layer order :57-59, 302-352: resnet,
// upsample, I, rnn1, rnn2, fc1, fc2, fc3 (the fatchord topology).  Same decoder as vocoder/libwavernn_bin.py (indices read as
// UNSIGNED bytes: the reference's own C++ reader mis-reads matrices with more than 127 groups per row, Q12).
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/wavernn_b200.h"

namespace {

enum { CONV1D = 1, CONV2D = 2, BATCHNORM1D = 3, LINEAR = 4, GRU = 5, STRETCH2D = 6 };

struct Reader {
    const std::vector<unsigned char>& buf;
    size_t pos = 0;
    std::string err;
    explicit Reader(const std::vector<unsigned char>& b) : buf(b) {}
    bool take(void* dst, size_t n) {
        if (!err.empty()) return false;
        if (pos + n > buf.size()) { err = "truncated libwavernn file"; return false; }
        memcpy(dst, buf.data() + pos, n);
        pos += n;
        return true;
    }
    int i32() { int v = 0; take(&v, 4); return v; }
    bool floats(std::vector<float>& out, long long count) {
        if (count < 0 || (size_t)count > buf.size()) { if (err.empty()) err = "truncated libwavernn file"; return false; }
        out.resize((size_t)count);
        return take(out.data(), (size_t)count * 4);
    }
    bool header(int want) {
        const int kind = i32();
        char name[64];
        take(name, 64);
        if (err.empty() && kind != want) err = "unexpected layer type: not a fatchord libwavernn export";
        return err.empty();
    }
    bool el4() { if (i32() != 4 && err.empty()) err = "element size is not 4: only float32 exports exist"; return err.empty(); }
    // compressed matrix -> dense rows x cols appended to `W`
    bool compressed(int rows, int cols, std::vector<float>& W) {
        const int nw = i32();
        std::vector<float> w;
        if (!floats(w, nw)) return false;
        const int nidx = i32();
        if (nidx < 0 || pos + (size_t)nidx > buf.size()) { if (err.empty()) err = "truncated libwavernn file"; return false; }
        const unsigned char* idx = buf.data() + pos;
        pos += (size_t)nidx;
        const size_t base = W.size();
        W.resize(base + (size_t)rows * cols, 0.f);
        int k = 0, at = 0;
        for (int r = 0; r < rows; ++r) {
            for (; at < nidx && idx[at] != 255; ++at) {
                const int c = idx[at];
                if ((c + 1) * 4 > cols || k + 4 > nw) { err = "compressed matrix index out of range"; return false; }
                memcpy(&W[base + (size_t)r * cols + c * 4], &w[k], 16);
                k += 4;
            }
            if (at >= nidx) { err = "compressed matrix does not describe all rows"; return false; }
            ++at;
        }
        if (k != nw) { err = "compressed matrix has stray weights"; return false; }
        return true;
    }
};

int put(wrnn_engine* e, const std::string& name, const std::vector<float>& v, std::initializer_list<int64_t> shape) {
    std::vector<int64_t> s(shape);
    return wrnn_set_tensor(e, name.c_str(), v.data(), s.data(), (int)s.size());
}

}  // namespace

extern "C" int wrnn_create_from_bin(const char* path, int device, wrnn_engine** out, char* err, int err_len) {
    auto fail = [&](int code, const std::string& msg) {
        if (err && err_len > 0) { strncpy(err, msg.c_str(), (size_t)err_len - 1); err[err_len - 1] = 0; }
        return code;
    };
    if (!path || !out) return fail(WRNN_ERR_INVALID, "wrnn_create_from_bin: bad argument");
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) return fail(WRNN_ERR_INVALID, "Cannot open file.");                      // WaveRNNVocoder.cpp:24-26
    std::vector<unsigned char> buf;
    {
        unsigned char chunk[1 << 16];
        size_t n;
        while ((n = fread(chunk, 1, sizeof(chunk), f)) > 0) buf.insert(buf.end(), chunk, chunk + n);
        fclose(f);
    }
    Reader r(buf);
    const int res_blocks = r.i32(), n_up = r.i32(), total = r.i32(), pad = r.i32();
    if (!r.err.empty() || res_blocks != 10 || n_up != 3 || total != 200 || pad != 2)
        return fail(WRNN_ERR_SHAPE, r.err.empty() ? "libwavernn file was exported with other hparams than the fatchord vocoder" : r.err);
    struct T { std::string name; std::vector<float> v; std::vector<int64_t> shape; };
    std::vector<T> ts;
    auto conv1d = [&](const std::string& name) {
        if (!r.header(CONV1D) || !r.el4()) return;
        const int has_bias = r.i32(), cin = r.i32(), cout = r.i32(), k = r.i32();
        T w; w.name = name + ".weight"; w.shape = {cout, cin, k};
        if (!r.floats(w.v, (long long)cout * cin * k)) return;
        ts.push_back(std::move(w));
        if (has_bias) { T b; b.name = name + ".bias"; b.shape = {cout}; if (r.floats(b.v, cout)) ts.push_back(std::move(b)); }
    };
    auto batchnorm = [&](const std::string& name) {
        if (!r.header(BATCHNORM1D) || !r.el4()) return;
        const int n = r.i32();
        float eps; r.take(&eps, 4);
        for (const char* part : {".weight", ".bias", ".running_mean", ".running_var"}) {
            T t; t.name = name + part; t.shape = {n};
            if (!r.floats(t.v, n)) return;
            ts.push_back(std::move(t));
        }
    };
    auto linear = [&](const std::string& name) -> int {
        if (!r.header(LINEAR) || !r.el4()) return 0;
        const int rows = r.i32(), cols = r.i32();
        if (rows <= 0 || cols <= 0 || rows > 65536 || cols > 65536) { if (r.err.empty()) r.err = "implausible Linear shape"; return 0; }
        T w; w.name = name + ".weight"; w.shape = {rows, cols};
        if (!r.compressed(rows, cols, w.v)) return 0;
        T b; b.name = name + ".bias"; b.shape = {rows};
        if (!r.floats(b.v, rows)) return 0;
        ts.push_back(std::move(w)); ts.push_back(std::move(b));
        return rows;
    };
    auto gru = [&](const std::string& name) {
        if (!r.header(GRU) || !r.el4()) return;
        const int hidden = r.i32(), inp = r.i32();
        if (hidden <= 0 || inp <= 0 || hidden > 65536 || inp > 65536) { if (r.err.empty()) r.err = "implausible GRU shape"; return; }
        T wi; wi.name = name + ".weight_ih_l0"; wi.shape = {3 * hidden, inp};
        T wh; wh.name = name + ".weight_hh_l0"; wh.shape = {3 * hidden, hidden};
        for (int g = 0; g < 3; ++g) if (!r.compressed(hidden, inp, wi.v)) return;
        for (int g = 0; g < 3; ++g) if (!r.compressed(hidden, hidden, wh.v)) return;
        T bi; bi.name = name + ".bias_ih_l0"; bi.shape = {3 * hidden};
        T bh; bh.name = name + ".bias_hh_l0"; bh.shape = {3 * hidden};
        std::vector<float> part;
        for (int g = 0; g < 6; ++g) {
            if (!r.floats(part, hidden)) return;
            (g < 3 ? bi.v : bh.v).insert((g < 3 ? bi.v : bh.v).end(), part.begin(), part.end());
        }
        ts.push_back(std::move(wi)); ts.push_back(std::move(wh)); ts.push_back(std::move(bi)); ts.push_back(std::move(bh));
    };
    const std::string rn = "upsample.resnet";
    conv1d(rn + ".conv_in");
    batchnorm(rn + ".batch_norm");
    for (int i = 0; i < res_blocks && r.err.empty(); ++i) {
        const std::string p = rn + ".layers." + std::to_string(i);
        conv1d(p + ".conv1"); batchnorm(p + ".batch_norm1");
        conv1d(p + ".conv2"); batchnorm(p + ".batch_norm2");
    }
    conv1d(rn + ".conv_out");
    if (r.header(STRETCH2D)) { const int sx = r.i32(); r.i32(); if (r.err.empty() && sx != total) r.err = "resnet stretch != total scale"; }
    const int want_scale[3] = {5, 5, 8};
    for (int i = 0; i < n_up && r.err.empty(); ++i) {
        if (!r.header(STRETCH2D)) break;
        const int sx = r.i32(); r.i32();
        if (r.err.empty() && sx != want_scale[i]) r.err = "upsample factors are not (5, 5, 8)";
        if (!r.header(CONV2D) || !r.el4()) break;
        const int k = r.i32();
        T t; t.name = "upsample.up_layers." + std::to_string(2 * i + 1) + ".weight"; t.shape = {1, 1, 1, k};
        if (r.floats(t.v, k)) ts.push_back(std::move(t));
    }
    if (r.err.empty()) linear("I");
    if (r.err.empty()) gru("rnn1");
    if (r.err.empty()) gru("rnn2");
    if (r.err.empty()) linear("fc1");
    if (r.err.empty()) linear("fc2");
    int C = 0;
    if (r.err.empty()) C = linear("fc3");
    if (r.err.empty() && r.pos != buf.size()) r.err = "trailing bytes: not a fatchord libwavernn export";
    if (!r.err.empty()) return fail(WRNN_ERR_SHAPE, r.err);
    // the file carries no hparams: mode / bits follow from fc3's row count (30 -> MOL, 2**bits -> RAW)
    int mode = WRNN_MODE_MOL, bits = 9;
    if (C != 30) {
        if (C < 2 || (C & (C - 1)) != 0) return fail(WRNN_ERR_SHAPE, "fc3 has neither 30 (MOL) nor 2**bits (RAW) rows");
        mode = WRNN_MODE_RAW;
        bits = 0;
        while ((1 << bits) < C) ++bits;
    }
    wrnn_engine* e = nullptr;
    int rc = wrnn_create(device, bits, mode, &e);
    if (rc != WRNN_OK) return fail(rc, "wrnn_create failed (no usable CUDA device, or unsupported class count)");
    for (const T& t : ts) {
        rc = wrnn_set_tensor(e, t.name.c_str(), t.v.data(), t.shape.data(), (int)t.shape.size());
        if (rc != WRNN_OK) break;
    }
    if (rc == WRNN_OK) rc = wrnn_finalize(e);
    if (rc != WRNN_OK) {
        const std::string msg = wrnn_last_error(e);
        wrnn_destroy(e);
        return fail(rc, msg);
    }
    *out = e;
    return WRNN_OK;
}
