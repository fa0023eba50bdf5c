// libwavernn_port.cpp -- plain C++17 restatement of the reference's CPU engine vocoder/libwavernn
// (fatchord_version/src): block-sparse 1x4 CompMatrix matvec (wavernn.h:23-92, wavernn.cpp:162-184), GRU
// (wavernn.cpp:112-159), Linear (:94-110), Conv1d/Conv2d/BatchNorm1d/Stretch2d (:186-327), Resnet / UpsampleNetwork
// / Model::apply (net_impl.cpp:19-27,38-92,129-210), reading the real `.bin` format (convert.py).
//
// TEST / BENCH INFRASTRUCTURE ONLY: the CPU comparator of BASELINE config 4.  libwavernn itself cannot be built here
// (Eigen 3.4.0, cnpy, pybind11 2.2.3 are network FetchContent, CMakeLists.txt:12-32), hence this Eigen-free port,
// compiled with the reference's flags (-O2 -ffast-math -march=native, CMakeLists.txt:42-43).
// Differences, stated: column indices are unsigned (the reference's int8_t colIdx overflows for 544-column
// matrices); the number of classes is read from fc3 instead of a compile-time constant (Q15); the
// categorical draw takes injected uniforms so it can be replayed against the oracle (the reference's static
// std::ranlux24 is used when none are given).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

namespace {

struct Mat { int rows = 0, cols = 0; std::vector<float> d; float& at(int r, int c) { return d[(size_t)r * cols + c]; }
             float at(int r, int c) const { return d[(size_t)r * cols + c]; } };

bool rd(FILE* f, void* p, size_t n) { return fread(p, 1, n, f) == n; }

struct CompMatrix {     // wavernn.h:23-92
    std::vector<float> w; std::vector<int> rowIdx; std::vector<int> colIdx; int nRows = 0, nCols = 0;
    bool read(FILE* f, int rows, int cols) {
        nRows = rows; nCols = cols;
        int nW = 0, nI = 0;
        if (!rd(f, &nW, 4)) return false;
        w.resize(nW); if (nW && !rd(f, w.data(), (size_t)nW * 4)) return false;
        if (!rd(f, &nI, 4)) return false;
        std::vector<uint8_t> idx(nI); if (nI && !rd(f, idx.data(), nI)) return false;
        int row = 0;
        for (uint8_t v : idx) { if (v == 255) ++row; else { colIdx.push_back((int)v); rowIdx.push_back(row); } }
        return (int)colIdx.size() * 4 == nW;
    }
    void mul(const float* x, float* y) const {        // wavernn.cpp:162-184
        std::fill(y, y + nRows, 0.f);
        const float* wp = w.data();
        const int n = (int)colIdx.size();
        for (int i = 0; i < n; ++i, wp += 4) {
            const float* xp = x + 4 * colIdx[i];
            y[rowIdx[i]] += wp[0] * xp[0] + wp[1] * xp[1] + wp[2] * xp[2] + wp[3] * xp[3];
        }
    }
};

struct Linear { CompMatrix m; std::vector<float> b;
    bool read(FILE* f) { int h[3]; if (!rd(f, h, 12)) return false; if (!m.read(f, h[1], h[2])) return false;
                         b.resize(h[1]); return rd(f, b.data(), (size_t)h[1] * 4); }
    void apply(const float* x, float* y) const { m.mul(x, y); for (int i = 0; i < m.nRows; ++i) y[i] += b[i]; } };

struct GRU { CompMatrix W[6]; std::vector<float> b[6]; int H = 0, I = 0;
    bool read(FILE* f) { int h[3]; if (!rd(f, h, 12)) return false; H = h[1]; I = h[2];
        for (int i = 0; i < 6; ++i) if (!W[i].read(f, H, i < 3 ? I : H)) return false;
        for (int i = 0; i < 6; ++i) { b[i].resize(H); if (!rd(f, b[i].data(), (size_t)H * 4)) return false; }
        return true; }
    void apply(const float* x, float* h, float* tmp) const {   // wavernn.cpp:150-159; tmp: 6*H
        for (int i = 0; i < 3; ++i) W[i].mul(x, tmp + i * H);
        for (int i = 3; i < 6; ++i) W[i].mul(h, tmp + i * H);
        for (int j = 0; j < H; ++j) {
            const float r = 1.f / (1.f + std::exp(-(tmp[j] + b[0][j] + tmp[3 * H + j] + b[3][j])));
            const float z = 1.f / (1.f + std::exp(-(tmp[H + j] + b[1][j] + tmp[4 * H + j] + b[4][j])));
            const float n = std::tanh(tmp[2 * H + j] + b[2][j] + r * (tmp[5 * H + j] + b[5][j]));
            h[j] = (1.f - z) * n + z * h[j];
        } } };

struct Conv1d { int in = 0, out = 0, k = 0; bool hasBias = false; std::vector<float> w, b;
    bool read(FILE* f) { int h[5]; if (!rd(f, h, 20)) return false; hasBias = h[1]; in = h[2]; out = h[3]; k = h[4];
        w.resize((size_t)in * out * k); if (!rd(f, w.data(), w.size() * 4)) return false;
        if (hasBias) { b.resize(out); if (!rd(f, b.data(), (size_t)out * 4)) return false; } return true; }
    Mat apply(const Mat& x) const {                             // wavernn.cpp:217-239
        Mat y; y.rows = out; y.cols = x.cols - k + 1; y.d.assign((size_t)y.rows * y.cols, 0.f);
        for (int o = 0; o < out; ++o) for (int c = 0; c < in; ++c) for (int j = 0; j < k; ++j) {
            const float ww = w[((size_t)o * in + c) * k + j];
            const float* xr = &x.d[(size_t)c * x.cols + j]; float* yr = &y.d[(size_t)o * y.cols];
            for (int t = 0; t < y.cols; ++t) yr[t] += ww * xr[t]; }
        if (hasBias) for (int o = 0; o < out; ++o) for (int t = 0; t < y.cols; ++t) y.at(o, t) += b[o];
        return y; } };

struct BatchNorm { int n = 0; float eps = 1e-5f; std::vector<float> g, be, mu, var;
    bool read(FILE* f) { int h[2]; if (!rd(f, h, 8) || !rd(f, &eps, 4)) return false; n = h[1];
        for (auto* v : {&g, &be, &mu, &var}) { v->resize(n); if (!rd(f, v->data(), (size_t)n * 4)) return false; } return true; }
    void apply(Mat& x, bool relu) const {                        // wavernn.cpp:294-304
        for (int c = 0; c < n; ++c) { const float inv = 1.f / std::sqrt(var[c] + eps);
            for (int t = 0; t < x.cols; ++t) { float v = (x.at(c, t) - mu[c]) * inv * g[c] + be[c]; x.at(c, t) = relu ? std::max(v, 0.f) : v; } } } };

struct Model {
    int resBlocks = 0, nUp = 0, totalScale = 0, pad = 0;
    Conv1d convIn, convOut; BatchNorm bnIn; std::vector<Conv1d> rc1, rc2; std::vector<BatchNorm> rb1, rb2;
    std::vector<int> upScale; std::vector<std::vector<float>> upW;
    Linear I, fc1, fc2, fc3; GRU rnn1, rnn2;
    bool skipHdr(FILE* f, int want) { int t; char name[64]; return rd(f, &t, 4) && rd(f, name, 64) && t == want; }
    bool load(const char* path) {
        FILE* f = fopen(path, "rb"); if (!f) return false;
        int h[4]; bool ok = rd(f, h, 16); resBlocks = h[0]; nUp = h[1]; totalScale = h[2]; pad = h[3];
        ok = ok && skipHdr(f, 1) && convIn.read(f) && skipHdr(f, 3) && bnIn.read(f);
        rc1.resize(resBlocks); rc2.resize(resBlocks); rb1.resize(resBlocks); rb2.resize(resBlocks);
        for (int i = 0; ok && i < resBlocks; ++i)
            ok = skipHdr(f, 1) && rc1[i].read(f) && skipHdr(f, 3) && rb1[i].read(f) && skipHdr(f, 1) && rc2[i].read(f) && skipHdr(f, 3) && rb2[i].read(f);
        ok = ok && skipHdr(f, 1) && convOut.read(f);
        int s2[2]; ok = ok && skipHdr(f, 6) && rd(f, s2, 8);
        for (int i = 0; ok && i < nUp; ++i) {
            ok = skipHdr(f, 6) && rd(f, s2, 8); upScale.push_back(s2[0]);
            int c2[2]; ok = ok && skipHdr(f, 2) && rd(f, c2, 8); std::vector<float> w(ok ? c2[1] : 0);
            ok = ok && rd(f, w.data(), w.size() * 4); upW.push_back(w); }
        ok = ok && skipHdr(f, 4) && I.read(f) && skipHdr(f, 5) && rnn1.read(f) && skipHdr(f, 5) && rnn2.read(f)
                && skipHdr(f, 4) && fc1.read(f) && skipHdr(f, 4) && fc2.read(f) && skipHdr(f, 4) && fc3.read(f);
        fclose(f); return ok; }

    // Model::apply, net_impl.cpp:150-210.  mel: (80, T) row-major.  out: T*totalScale samples in [-1, 1].
    void apply(const float* mel, int T, const float* uniforms, float* out) const {
        const int F = convIn.in, Tp = T + 2 * pad;
        Mat mp; mp.rows = F; mp.cols = Tp; mp.d.assign((size_t)F * Tp, 0.f);
        for (int c = 0; c < F; ++c) std::memcpy(&mp.d[(size_t)c * Tp + pad], mel + (size_t)c * T, (size_t)T * 4);
        // upsample network (net_impl.cpp:85-92, wavernn.cpp:253-270,315-327)
        Mat m = mp;
        for (int l = 0; l < nUp; ++l) {
            const int s = upScale[l], k = (int)upW[l].size(), np = (k - 1) / 2;
            Mat st; st.rows = F; st.cols = m.cols * s; st.d.resize((size_t)F * st.cols);
            for (int c = 0; c < F; ++c) for (int t = 0; t < st.cols; ++t) st.at(c, t) = m.at(c, t / s);
            Mat y = st;
            for (int c = 0; c < F; ++c) for (int t = 0; t < st.cols; ++t) { float a = 0.f;
                for (int j = 0; j < k; ++j) { const int q = t + j - np; if (q >= 0 && q < st.cols) a += upW[l][j] * st.at(c, q); }
                y.at(c, t) = a; }
            m = y; }
        const int indent = pad * totalScale, N = m.cols - 2 * indent;
        // resnet (net_impl.cpp:38-75)
        Mat a = convIn.apply(mp); bnIn.apply(a, true);
        for (int i = 0; i < resBlocks; ++i) { Mat r = a; Mat y = rc1[i].apply(a); rb1[i].apply(y, true); y = rc2[i].apply(y); rb2[i].apply(y, false);
            for (size_t q = 0; q < y.d.size(); ++q) y.d[q] += r.d[q]; a = y; }
        a = convOut.apply(a);                                      // (128, T); stretch by totalScale is implicit below
        const int nAux = a.rows, d = nAux / 4, H = rnn1.H, C = fc3.m.nRows;
        std::vector<float> xin(I.m.nCols), y(H), h1(H, 0.f), h2(H, 0.f), inp(H + d), tmp(6 * H), f1(H), f2(H), lg(C), in2(H + d);
        float x = 0.f;
        static std::ranlux24 rnd;
        for (int i = 0; i < N; ++i) {
            const int fr = i / totalScale;
            xin[0] = x;
            for (int c = 0; c < F; ++c) xin[1 + c] = m.at(c, indent + i);
            for (int c = 0; c < d - 1; ++c) xin[1 + F + c] = a.at(c, fr);          // a1 without its last row (Q5)
            I.apply(xin.data(), y.data());
            rnn1.apply(y.data(), h1.data(), tmp.data());
            for (int j = 0; j < H; ++j) y[j] += h1[j];
            std::copy(y.begin(), y.end(), inp.begin());
            for (int c = 0; c < d; ++c) inp[H + c] = a.at(d + c, fr);
            rnn2.apply(inp.data(), h2.data(), tmp.data());
            for (int j = 0; j < H; ++j) y[j] += h2[j];
            std::copy(y.begin(), y.end(), in2.begin());
            for (int c = 0; c < d; ++c) in2[H + c] = a.at(2 * d + c, fr);
            fc1.apply(in2.data(), f1.data()); for (auto& v : f1) v = std::max(v, 0.f);
            std::copy(f1.begin(), f1.end(), in2.begin());
            for (int c = 0; c < d; ++c) in2[H + c] = a.at(3 * d + c, fr);
            fc2.apply(in2.data(), f2.data()); for (auto& v : f2) v = std::max(v, 0.f);
            fc3.apply(f2.data(), lg.data());
            float mx = lg[0]; for (int c = 1; c < C; ++c) mx = std::max(mx, lg[c]);     // softmax, net_impl.cpp:19-27
            float sum = 0.f; for (int c = 0; c < C; ++c) { lg[c] = std::exp(lg[c] - mx); sum += lg[c]; }
            const float u = uniforms ? uniforms[i] : static_cast<float>(rnd()) / rnd.max();
            float cdf = 0.f; int k = C - 1;                                                // sampleCategorical, :129-140
            for (int c = 0; c < C; ++c) { cdf += lg[c] / sum; if (cdf >= u) { k = c; break; } }
            x = (2.f * k) / (C - 1.f) - 1.f;
            out[i] = x;
        }
    }
};
}  // namespace

extern "C" {
void* lwr_load(const char* path) { Model* m = new Model(); if (!m->load(path)) { delete m; return nullptr; } return m; }
void lwr_free(void* h) { delete static_cast<Model*>(h); }
int lwr_mel_to_wav(void* h, const float* mel, int T, const float* uniforms, float* out) {
    if (!h) return -1; static_cast<Model*>(h)->apply(mel, T, uniforms, out); return 0; }
int lwr_total_scale(void* h) { return h ? static_cast<Model*>(h)->totalScale : 0; }
}
