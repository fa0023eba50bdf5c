// engine.cu -- host side of the engine and the C ABI of include/wavernn_b200.h.
// Owns device memory, streams and launches; no torch, no CPU arithmetic on the data path (the only host
// math is one-off weight preparation at wrnn_finalize and O(folds) index planning).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../../include/wavernn_b200.h"
#include "engine_internal.h"

using namespace wrnn;

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct HostTensor {
    std::vector<float> data;
    std::vector<int64_t> shape;
};

constexpr int kResBlocks = 10;
constexpr int kCondLayers = 2 + 2 * kResBlocks;   // conv_in, 20 1x1, conv_out

// device pointers of the geneing topology (engine_gn.inc)
struct GnDev {
    float *Whh = nullptr, *Wfc1a = nullptr, *Wfc3 = nullptr, *u1 = nullptr, *u2 = nullptr, *bhn = nullptr, *bfc3 = nullptr, *coef = nullptr;
    float* CW[2 + 2 * kGnResBlocks] = {};      // front end (64 channels, 3 residual blocks): conv_in, 6 1x1, conv_out
    float* CB[2 + 2 * kGnResBlocks] = {};
    float *MA = nullptr, *bA = nullptr, *MQ = nullptr;      // table projections [1024][64], [1024], [1024][80]
};
// device pointers of the runtimeracer topology (engine_rr.inc)
struct RrDev {
    float *Whh[4] = {}, *Wih[3] = {}, *M12 = nullptr, *M34 = nullptr, *Wfc5 = nullptr, *u = nullptr, *bhn = nullptr, *bfc5 = nullptr, *coef = nullptr;
    float *MA = nullptr, *bA = nullptr, *MQ = nullptr;      // table projections [4096][128], [4096], [4096][80]
};

}  // namespace

struct wrnn_engine {
    int device = 0, bits = 9, mode = 0, C = 512, Cpad = 512, CR = 4;
    int rs_samplers = 0;         // RAW on the role-specialised loop: sampler CTAs per group (wrnn_finalize)
    int topology = 0;            // WRNN_TOPO_FATCHORD / WRNN_TOPO_RUNTIMERACER (wrnn_set_topology, before wrnn_finalize)
    RrDev rr;
    GnDev gn;
    std::string err;
    std::map<std::string, HostTensor> tensors;
    int64_t step = 0;
    bool finalized = false;
    double sparsity = 0.0;
    int64_t launches = 0;
    cudaStream_t stream = nullptr;
    int n_sms = 0;
    size_t smem_limit = 0;
    // weights on device
    DevBuf wLoop;    // all loop weights, one allocation
    float *dWhh1 = nullptr, *dWih2a = nullptr, *dWhh2 = nullptr, *dWfc1a = nullptr, *dWfc2a = nullptr, *dWfc3 = nullptr;
    float *dv1 = nullptr, *dv2 = nullptr, *dv3 = nullptr, *dbhn1 = nullptr, *dbhn2 = nullptr, *dbfc3 = nullptr, *dcoef = nullptr;
    DevBuf wCond;    // conditioning weights
    float* dCW[kCondLayers] = {};
    float* dCB[kCondLayers] = {};
    float *dMA1 = nullptr, *dbA1 = nullptr, *dMA2 = nullptr, *dbA2 = nullptr, *dMQ1 = nullptr, *dMQ2 = nullptr;
    std::vector<float> hcoef;
    // grow-only work buffers
    DevBuf bMel, bUtt, bX0, bMP, bH[3], bAux, bTA1, bTA2, bTQ1, bTQ2, bFolds, bExch, bSamples, bLogits, bForced;
    DevBuf bPostUtt, bFade, bScratch, bWav, bFloor, bCsDone;
    DevBuf wTc, wTcS, bTcExch, bCS;
    DevBuf wTc2;                // cluster-local tensor-core loop (MOL): 16 per-CTA weight tile streams
    DevBuf wRsX[3], bFR, bM16, bRsPlace;  // inline conditioning of the role-specialised loop: W_q tiles per role, per-frame rows, upsampled mel (fp16)
    DevBuf wRs[5], bRsExch;     // role-specialised tensor-core loop: per-role weight images ([4]: RAW sampler CTAs), exchange matrices + sample words
    DevBuf wSp[2];              // block-sparse cluster loop: per-CTA compressed images for cluster sizes 16 and 8
    int spStride[2] = {0, 0};
    DevBuf wCondTc, bCondH;     // tensor-core front end: hi/lo fp16 weights (scaled by 2^8) and activation pairs
    size_t oCondTc[kCondLayers + 4] = {};   // element offsets of each layer's W_hi inside wCondTc (W_lo follows)   // tensor-core loop: per-CTA fp16 weight images, exchange buffers, per-sample conditioning
    int* dAbort = nullptr;
    int* hProgress = nullptr;   // mapped pinned
    int* dProgress = nullptr;
    int fade_overlap = -1;
    cudaEvent_t ev[8] = {};
    cudaEvent_t evx[2] = {};     // brackets the per-sample conditioning expansion of a wave
    cudaEvent_t evc[2] = {};     // brackets one trial of the role-specialised loop's layout calibration
    std::map<int, std::pair<int, int>> rs_layout;     // (mode, samplers, fold bucket) -> (groups, grid padded): measured once per engine
};

namespace {

int fail(wrnn_engine* e, int code, const std::string& msg) {
    if (e) e->err = msg;
    return code;
}
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t _err = (call);                                                                 \
        if (_err != cudaSuccess)                                                                   \
            return fail(e, WRNN_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_err));   \
    } while (0)

const HostTensor* get(wrnn_engine* e, const char* name, std::initializer_list<int64_t> shape) {
    auto it = e->tensors.find(name);
    if (it == e->tensors.end()) { e->err = std::string("missing tensor '") + name + "'"; return nullptr; }
    const HostTensor& t = it->second;
    if (t.shape.size() != shape.size() || !std::equal(shape.begin(), shape.end(), t.shape.begin())) {
        e->err = std::string("shape mismatch for '") + name + "'";
        return nullptr;
    }
    return &t;
}

// numpy.linspace(start, stop, num) in float64 (the reference builds its fades with it, :385, :252)
void np_linspace(double start, double stop, int num, std::vector<double>& out) {
    out.resize(num);
    if (num == 1) { out[0] = start; return; }
    const double step = (stop - start) / (double)(num - 1);
    for (int i = 0; i < num; ++i) {
        volatile double prod = (double)i * step;   // no FMA contraction: numpy does mul then add
        out[i] = prod + start;
    }
    out[num - 1] = stop;
}

// composite interpolation weights of the three Stretch2d+Conv2d layers (fatchord_version.py:66-75,82-84):
// coef[phase][d] multiplies padded mel frame (n/200 + d) for output sample n with n%200 == phase.
void build_coef(const float* w5a, const float* w5b, const float* w8, std::vector<float>& coef, int s0 = 5, int s1 = 5, int s2 = 8) {
    const int L = 8, scales[3] = {s0, s1, s2};       // (geneing: 4, 5, 10 -- the same hop of 200 and the same reach of +-2 frames)
    const float* ws[3] = {w5a, w5b, w8};
    std::vector<double> x(L, 0.0);
    x[3] = 1.0;
    for (int s = 0; s < 3; ++s) {
        const int sc = scales[s], n = (int)x.size() * sc, k = 2 * sc + 1;
        std::vector<double> st(n), y(n, 0.0);
        for (int i = 0; i < n; ++i) st[i] = x[i / sc];
        for (int i = 0; i < n; ++i) {
            double a = 0.0;
            for (int j = 0; j < k; ++j) {
                const int q = i + j - sc;
                if (q >= 0 && q < n) a += (double)ws[s][j] * st[q];
            }
            y[i] = a;
        }
        x.swap(y);
    }
    coef.assign(kHop * kTaps, 0.f);
    for (int ph = 0; ph < kHop; ++ph)
        for (int d = 0; d < kTaps; ++d) coef[ph * kTaps + d] = (float)x[1000 - 200 * d + ph];
}

int python_floordiv(int64_t a, int64_t b) { int64_t q = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) --q; return (int)q; }

#include "engine_rr.inc"
#include "engine_gn.inc"

}  // namespace

extern "C" {

int wrnn_fold_plan(int64_t total_len, int64_t target, int64_t overlap, int64_t* num_folds, int64_t* padded_len) {
    if (target + overlap <= 0 || target < 0 || overlap < 0) return WRNN_ERR_INVALID;
    int64_t nf = python_floordiv(total_len - overlap, target + overlap);
    const int64_t extended = nf * (overlap + target) + overlap;
    const int64_t remaining = total_len - extended;
    int64_t padded = total_len;
    if (remaining != 0) {
        nf += 1;
        padded = total_len + target + 2 * overlap - remaining;
    }
    if (num_folds) *num_folds = nf;
    if (padded_len) *padded_len = padded;
    return WRNN_OK;
}

int wrnn_create(int device, int bits, int mode, wrnn_engine** out) {
    if (!out) return WRNN_ERR_INVALID;
    *out = nullptr;
    if (mode != WRNN_MODE_RAW && mode != WRNN_MODE_MOL) return WRNN_ERR_INVALID;
    if (mode == WRNN_MODE_RAW && (bits < 8 || bits > 10)) return WRNN_ERR_INVALID;      // (header: RAW 8..10 bits; the reference trains 9 / 10)
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return WRNN_ERR_CUDA;
    wrnn_engine* e = new wrnn_engine();
    e->device = device;
    e->bits = bits;
    e->mode = mode;
    e->C = (mode == WRNN_MODE_RAW) ? (1 << bits) : 30;
    e->Cpad = (e->C + 1) & ~1;
    e->CR = (e->C + kCtasF32 - 1) / kCtasF32;
    cudaError_t err = cudaSetDevice(device);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaDeviceGetAttribute(&e->n_sms, cudaDevAttrMultiProcessorCount, device);
    int smem = 0;
    if (err == cudaSuccess) err = cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    e->smem_limit = (size_t)smem;
    if (err == cudaSuccess) err = cudaMalloc(&e->dAbort, 2 * sizeof(int));   // [0] loop abort flag, [1] front-end status
    if (err == cudaSuccess) err = cudaHostAlloc(&e->hProgress, sizeof(int), cudaHostAllocMapped);
    if (err == cudaSuccess) err = cudaHostGetDevicePointer(&e->dProgress, e->hProgress, 0);
    for (int i = 0; i < 8 && err == cudaSuccess; ++i) err = cudaEventCreate(&e->ev[i]);
    for (int i = 0; i < 2 && err == cudaSuccess; ++i) err = cudaEventCreate(&e->evx[i]);
    for (int i = 0; i < 2 && err == cudaSuccess; ++i) err = cudaEventCreate(&e->evc[i]);
    if (const char* dl = getenv("WRNN_SPIN_DEADLINE_MS"))   // profilers slow the loop down: let them widen the guard
        if (err == cudaSuccess) err = set_spin_deadline((long long)(atof(dl) * 1.9e6));
    if (const char* dl = getenv("WRNN_SPIN_DEADLINE_MS"))
        if (err == cudaSuccess) err = set_tc_deadline((long long)(atof(dl) * 1.9e6));
    if (const char* dl = getenv("WRNN_SPIN_DEADLINE_MS"))
        if (err == cudaSuccess) err = set_tc2_deadline((long long)(atof(dl) * 1.9e6));
    if (const char* dl = getenv("WRNN_SPIN_DEADLINE_MS"))
        if (err == cudaSuccess) err = set_rs_deadline((long long)(atof(dl) * 1.9e6));
    if (const char* dl = getenv("WRNN_SPIN_DEADLINE_MS"))
        if (err == cudaSuccess) err = set_rr_deadline((long long)(atof(dl) * 1.9e6));
    if (const char* dl = getenv("WRNN_SPIN_DEADLINE_MS"))
        if (err == cudaSuccess) err = set_gn_deadline((long long)(atof(dl) * 1.9e6));
    if (err != cudaSuccess) { delete e; return WRNN_ERR_CUDA; }
    *out = e;
    return WRNN_OK;
}

int wrnn_destroy(wrnn_engine* e) {
    if (!e) return WRNN_OK;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    DevBuf* bufs[] = {&e->wLoop, &e->wCond, &e->bMel, &e->bUtt, &e->bX0, &e->bMP, &e->bH[0], &e->bH[1], &e->bH[2], &e->bAux,
                      &e->bTA1, &e->bTA2, &e->bTQ1, &e->bTQ2, &e->bFolds, &e->bExch, &e->bSamples, &e->bLogits, &e->bForced,
                      &e->bPostUtt, &e->bFade, &e->bScratch, &e->bWav, &e->bFloor, &e->bCsDone, &e->wTc, &e->wTcS, &e->bTcExch, &e->bCS, &e->wCondTc, &e->bCondH, &e->wSp[0], &e->wSp[1], &e->wTc2, &e->wRs[0], &e->wRs[1], &e->wRs[2], &e->wRs[3], &e->wRs[4], &e->bRsExch, &e->wRsX[0], &e->wRsX[1], &e->wRsX[2], &e->bFR, &e->bM16, &e->bRsPlace};
    for (DevBuf* b : bufs) b->release();
    if (e->dAbort) cudaFree(e->dAbort);
    if (e->hProgress) cudaFreeHost(e->hProgress);
    for (auto& ev : e->ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : e->evx) if (ev) cudaEventDestroy(ev);
    for (auto& ev : e->evc) if (ev) cudaEventDestroy(ev);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
    return WRNN_OK;
}

const char* wrnn_last_error(const wrnn_engine* e) { return e ? e->err.c_str() : "null engine"; }

int wrnn_set_tensor(wrnn_engine* e, const char* name, const float* data, const int64_t* shape, int ndim) {
    if (!e || !name || !data || ndim < 0 || ndim > 4) return fail(e, WRNN_ERR_INVALID, "wrnn_set_tensor: bad argument");
    HostTensor t;
    int64_t n = 1;
    for (int i = 0; i < ndim; ++i) { t.shape.push_back(shape[i]); n *= shape[i]; }
    t.data.assign(data, data + n);
    e->tensors[name] = std::move(t);
    e->finalized = false;
    return WRNN_OK;
}
int wrnn_set_step(wrnn_engine* e, int64_t step) { if (!e) return WRNN_ERR_INVALID; e->step = step; return WRNN_OK; }
int64_t wrnn_get_step(const wrnn_engine* e) { return e ? e->step : -1; }
double wrnn_sparsity(const wrnn_engine* e) { return e ? e->sparsity : 0.0; }
int wrnn_sparse_available(const wrnn_engine* e) { return (e && (e->spStride[0] || e->spStride[1])) ? 1 : 0; }
int64_t wrnn_launch_count(const wrnn_engine* e) { return e ? e->launches : 0; }

int wrnn_set_topology(wrnn_engine* e, int topology) {
    if (!e || (topology != WRNN_TOPO_FATCHORD && topology != WRNN_TOPO_RUNTIMERACER && topology != WRNN_TOPO_GENEING))
        return fail(e, WRNN_ERR_INVALID, "unknown topology");
    e->topology = topology;
    e->finalized = false;
    return WRNN_OK;
}

int wrnn_finalize(wrnn_engine* e) {
    if (!e) return WRNN_ERR_INVALID;
    if (e->topology == WRNN_TOPO_RUNTIMERACER) return finalize_rr(e, e->rr);
    if (e->topology == WRNN_TOPO_GENEING) return finalize_gn(e, e->gn);
    CU(cudaSetDevice(e->device));
    const int H = kRnn, C = e->C;
#define GET(var, name, ...)                                   \
    const HostTensor* var = get(e, name, {__VA_ARGS__});      \
    if (!var) return WRNN_ERR_SHAPE;
    GET(tI, "I.weight", H, 112)
    GET(tIb, "I.bias", H)
    GET(r1ih, "rnn1.weight_ih_l0", 3 * H, H)
    GET(r1hh, "rnn1.weight_hh_l0", 3 * H, H)
    GET(r1bi, "rnn1.bias_ih_l0", 3 * H)
    GET(r1bh, "rnn1.bias_hh_l0", 3 * H)
    GET(r2ih, "rnn2.weight_ih_l0", 3 * H, H + kAux)
    GET(r2hh, "rnn2.weight_hh_l0", 3 * H, H)
    GET(r2bi, "rnn2.bias_ih_l0", 3 * H)
    GET(r2bh, "rnn2.bias_hh_l0", 3 * H)
    GET(f1w, "fc1.weight", H, H + kAux)
    GET(f1b, "fc1.bias", H)
    GET(f2w, "fc2.weight", H, H + kAux)
    GET(f2b, "fc2.bias", H)
    GET(f3w, "fc3.weight", C, H)
    GET(f3b, "fc3.bias", C)
    GET(cin, "upsample.resnet.conv_in.weight", 128, kFeat, 5)
    GET(cout, "upsample.resnet.conv_out.weight", 128, 128, 1)
    GET(coutb, "upsample.resnet.conv_out.bias", 128)
    GET(up1, "upsample.up_layers.1.weight", 1, 1, 1, 11)
    GET(up3, "upsample.up_layers.3.weight", 1, 1, 1, 11)
    GET(up5, "upsample.up_layers.5.weight", 1, 1, 1, 17)

    // ---- loop weights ------------------------------------------------------------------------------------
    std::vector<float> hl;
    size_t off = 0;
    auto take = [&](size_t n) { size_t o = off; off += (n + 3) & ~(size_t)3; return o; };
    const size_t oWhh1 = take((size_t)3 * H * H), oWih2a = take((size_t)3 * H * H), oWhh2 = take((size_t)3 * H * H);
    const size_t oWfc1a = take((size_t)H * H), oWfc2a = take((size_t)H * H), oWfc3 = take((size_t)C * H);
    const size_t ov1 = take(3 * H), ov2 = take(3 * H), ov3 = take(H), obhn1 = take(H), obhn2 = take(H), obfc3 = take(C),
                 ocoef = take(kHop * kTaps);
    hl.assign(off, 0.f);
    std::copy(r1hh->data.begin(), r1hh->data.end(), hl.begin() + oWhh1);
    std::copy(r2hh->data.begin(), r2hh->data.end(), hl.begin() + oWhh2);
    for (int r = 0; r < 3 * H; ++r)
        std::copy(r2ih->data.begin() + (size_t)r * (H + kAux), r2ih->data.begin() + (size_t)r * (H + kAux) + H,
                  hl.begin() + oWih2a + (size_t)r * H);
    for (int r = 0; r < H; ++r) {
        std::copy(f1w->data.begin() + (size_t)r * (H + kAux), f1w->data.begin() + (size_t)r * (H + kAux) + H,
                  hl.begin() + oWfc1a + (size_t)r * H);
        std::copy(f2w->data.begin() + (size_t)r * (H + kAux), f2w->data.begin() + (size_t)r * (H + kAux) + H,
                  hl.begin() + oWfc2a + (size_t)r * H);
    }
    std::copy(f3w->data.begin(), f3w->data.end(), hl.begin() + oWfc3);
    std::copy(f3b->data.begin(), f3b->data.end(), hl.begin() + obfc3);
    for (int j = 0; j < H; ++j) {
        hl[obhn1 + j] = r1bh->data[2 * H + j];
        hl[obhn2 + j] = r2bh->data[2 * H + j];
    }
    // The I layer feeds rnn1 (W_ih1), rnn2 (through the residual x1 = xI + h1) and fc1 (x2 = xI + h1 + h2),
    // always linearly, so its contribution is pre-multiplied in float64:
    //   P1 = W_ih1 @ [I.weight | I.bias],  P2 = W_ih2[:, :512] @ [...],  P3 = fc1[:, :512] @ [...]
    // column 0 (the previous sample) becomes the rank-1 coefficients v1/v2/v3 used inside the loop,
    // columns 1..80 (mel) and 81..111 (aux a1[:31]) go into the per-frame conditioning projections.
    std::vector<double> P1((size_t)3 * H * 113, 0.0), P2((size_t)3 * H * 113, 0.0), P3((size_t)H * 113, 0.0);
    auto premul = [&](const float* W, int ldw, int rows, std::vector<double>& P) {
        for (int r = 0; r < rows; ++r) {
            double* pr = &P[(size_t)r * 113];
            for (int k = 0; k < H; ++k) {
                const double w = W[(size_t)r * ldw + k];
                const float* ik = &tI->data[(size_t)k * 112];
                for (int c = 0; c < 112; ++c) pr[c] += w * (double)ik[c];
                pr[112] += w * (double)tIb->data[k];
            }
        }
    };
    premul(r1ih->data.data(), H, 3 * H, P1);
    premul(r2ih->data.data(), H + kAux, 3 * H, P2);
    premul(f1w->data.data(), H + kAux, H, P3);
    for (int r = 0; r < 3 * H; ++r) {
        hl[ov1 + r] = (float)P1[(size_t)r * 113];
        hl[ov2 + r] = (float)P2[(size_t)r * 113];
    }
    for (int j = 0; j < H; ++j) hl[ov3 + j] = (float)P3[(size_t)j * 113];
    build_coef(up1->data.data(), up3->data.data(), up5->data.data(), e->hcoef);
    std::copy(e->hcoef.begin(), e->hcoef.end(), hl.begin() + ocoef);

    // sparsity of a pruned checkpoint (vocoder/pruner.py leaves zeros in 1x4 column groups)
    {
        size_t zero = 0, total = 0;
        auto scan = [&](const std::vector<float>& w, int cols, int c0, int c1) {
            const size_t rows = w.size() / cols;
            for (size_t r = 0; r < rows; ++r)
                for (int c = c0; c + 3 < c1; c += 4) {
                    const float* q = &w[r * cols + c];
                    ++total;
                    if (q[0] == 0.f && q[1] == 0.f && q[2] == 0.f && q[3] == 0.f) ++zero;
                }
        };
        scan(r1hh->data, H, 0, H); scan(r2ih->data, H + kAux, 0, H + kAux); scan(r2hh->data, H, 0, H);
        scan(f1w->data, H + kAux, 0, H + kAux); scan(f2w->data, H + kAux, 0, H + kAux); scan(f3w->data, H, 0, H);
        e->sparsity = total ? (double)zero / (double)total : 0.0;
    }
    CU(e->wLoop.ensure(hl.size() * sizeof(float)));
    CU(cudaMemcpy(e->wLoop.p, hl.data(), hl.size() * sizeof(float), cudaMemcpyHostToDevice));
    float* base = e->wLoop.as<float>();
    e->dWhh1 = base + oWhh1; e->dWih2a = base + oWih2a; e->dWhh2 = base + oWhh2;
    e->dWfc1a = base + oWfc1a; e->dWfc2a = base + oWfc2a; e->dWfc3 = base + oWfc3;
    e->dv1 = base + ov1; e->dv2 = base + ov2; e->dv3 = base + ov3; e->dbhn1 = base + obhn1; e->dbhn2 = base + obhn2;
    e->dbfc3 = base + obfc3; e->dcoef = base + ocoef;

    // ---- tensor-core loop: per-CTA fp16 weight images, already in the shared-memory layout of loop_tc.cu ----
    // Stage tiles are K-major SWIZZLE_128B: per 64-column k-block, rows of 128 bytes, 16-byte chunk c of row r
    // stored at chunk (c ^ (r & 7)).  Row order inside a stage groups the two hidden units of an epilogue thread.
    if (C == 30 || C == 512 || C == 1024) {
        const size_t img = loop_tc_weight_image_bytes();
        std::vector<unsigned char> hw((size_t)kTcCtas * img, 0);
        auto put = [&](unsigned char* stage, int nrows, int r, const float* src) {
            for (int kb = 0; kb < kRnn / 64; ++kb)
                for (int c = 0; c < 8; ++c) {
                    __half* dst = reinterpret_cast<__half*>(stage + (size_t)kb * nrows * 128 + r * 128 + ((c ^ (r & 7)) << 4));
                    for (int i = 0; i < 8; ++i) dst[i] = __float2half_rn(src[kb * 64 + c * 8 + i]);
                }
        };
        const float* Wih2a = hl.data() + oWih2a;
        const float* Wfc1a = hl.data() + oWfc1a;
        const float* Wfc2a = hl.data() + oWfc2a;
        for (int cta = 0; cta < kTcCtas; ++cta) {
            unsigned char* base_img = hw.data() + (size_t)cta * img;
            unsigned char* sB = base_img;                           // 64 rows
            unsigned char* sC = sB + (size_t)64 * 128 * 8;          // 32 rows
            unsigned char* sD = sC + (size_t)32 * 128 * 8;          // 16 rows
            unsigned char* sE = sD + (size_t)16 * 128 * 8;          // 16 (RAW) or 32 (MOL) rows
            for (int up = 0; up < 4; ++up)
                for (int u = 0; u < 2; ++u) {
                    const int j = cta * kTcUnits + 2 * up + u;
                    for (int gt = 0; gt < 3; ++gt) {
                        put(sB, 64, 16 * up + 2 * gt + u, Wih2a + (size_t)(gt * H + j) * H);              // p2
                        put(sB, 64, 16 * up + 8 + 2 * gt + u, r1hh->data.data() + (size_t)(gt * H + j) * H);  // gh1'
                        put(sC, 32, 8 * up + 2 * gt + u, r2hh->data.data() + (size_t)(gt * H + j) * H);       // gh2'
                    }
                    put(sB, 64, 16 * up + 6 + u, Wfc1a + (size_t)j * H);                                   // p3
                    put(sC, 32, 8 * up + 6 + u, Wfc1a + (size_t)j * H);                                    // q3
                    put(sD, 16, 2 * up + u, Wfc2a + (size_t)j * H);
                }
            if (C == 30) {
                if (cta == 0)
                    for (int c = 0; c < 30; ++c) put(sE, 32, c, f3w->data.data() + (size_t)c * H);
            } else {
                const int cpu = C / (kTcCtas * 4);
                for (int up = 0; up < 4; ++up)
                    for (int i = 0; i < cpu; ++i)
                        put(sE, 16, cpu * up + i, f3w->data.data() + (size_t)(cta * cpu * 4 + cpu * up + i) * H);
            }
        }
        CU(e->wTc.ensure(hw.size()));
        CU(cudaMemcpy(e->wTc.p, hw.data(), hw.size(), cudaMemcpyHostToDevice));
        if (C == 512) {     // RAW sampler CTAs: quarter q of fc3 (rows 128 q ..) as eight [128 x 64] tiles
            const size_t simg = loop_tc_raw_sampler_image_bytes();
            const int nq = loop_tc_raw_sampler_ctas();
            std::vector<unsigned char> hs((size_t)nq * simg, 0);
            for (int q = 0; q < nq; ++q)
                for (int r = 0; r < C / nq; ++r) put(hs.data() + (size_t)q * simg, C / nq, r, f3w->data.data() + (size_t)(q * (C / nq) + r) * H);
            CU(e->wTcS.ensure(hs.size()));
            CU(cudaMemcpy(e->wTcS.p, hs.data(), hs.size(), cudaMemcpyHostToDevice));
        }
    }

    // ---- role-specialised tensor-core loop (loop_rs.cu): one image per CTA and role, K-major SWIZZLE_128B tiles ---------------
    //   T1 (GRU1, units 32c..32c+31): [W_hh1 rows gate*32+u (96)] [MOL: fc3 rows (30 of 32)]  T2 (GRU2): [W_ih2a (96)] [W_hh2 (96)]
    //   T3 (fc1, units 64c..64c+63): [fc1[:, :512] (64)]                                       T4 (fc2): [fc2[:, :512] (64)]
    //   T5 (RAW with 512 / 1024 classes: sampler CTAs): [fc3 rows 128c..128c+127]
    const bool rs_raw = e->mode == WRNN_MODE_RAW && (C == 512 || C == 1024);          // 8 sampler CTAs with 64 / 128 classes each
    if (C == 30 || rs_raw) {
        auto put_tile = [&](unsigned char* tile, int nrows, int r, const float* src) {
            for (int kb = 0; kb < 8; ++kb)
                for (int c = 0; c < 8; ++c) {
                    __half* dst = reinterpret_cast<__half*>(tile + (size_t)kb * nrows * 128 + r * 128 + ((c ^ (r & 7)) << 4));
                    for (int i = 0; i < 8; ++i) dst[i] = __float2half_rn(src[kb * 64 + c * 8 + i]);
                }
        };
        const float* Wih2a = hl.data() + oWih2a;
        const float* Wfc1a = hl.data() + oWfc1a;
        const float* Wfc2a = hl.data() + oWfc2a;
        const int nct[4] = {kRsT1, kRsT2, kRsT3, kRsT4};
        for (int role = 0; role < 4; ++role) {
            const size_t img = loop_rs_image_bytes(role);
            std::vector<unsigned char> hw((size_t)nct[role] * img, 0);
            for (int c = 0; c < nct[role]; ++c) {
                unsigned char* t0 = hw.data() + (size_t)c * img;
                unsigned char* t1 = t0 + (size_t)96 * 1024;
                if (role < 2) {
                    for (int u = 0; u < 32; ++u)
                        for (int gt = 0; gt < 3; ++gt) {
                            const size_t r = (size_t)(gt * H + 32 * c + u) * H;
                            if (role == 0) put_tile(t0, 96, 32 * gt + u, r1hh->data.data() + r);
                            else { put_tile(t0, 96, 32 * gt + u, Wih2a + r); put_tile(t1, 96, 32 * gt + u, r2hh->data.data() + r); }
                        }
                    if (role == 0 && C == 30)
                        for (int k = 0; k < 30; ++k) put_tile(t1, 32, k, f3w->data.data() + (size_t)k * H);
                } else {
                    const int FU = H / kRsT3;
                    for (int u = 0; u < FU; ++u) put_tile(t0, FU, u, (role == 2 ? Wfc1a : Wfc2a) + (size_t)(FU * c + u) * H);
                }
            }
            CU(e->wRs[role].ensure(hw.size()));
            CU(cudaMemcpy(e->wRs[role].p, hw.data(), hw.size(), cudaMemcpyHostToDevice));
        }
        // inline conditioning: the mel columns (1..80) of the folded input-side matrices P1 / P2 / P3 as fp16 tiles
        // [k-block 2][rows][128 B] (k-block 0: mel 0..63, k-block 1: mel 64..79 in the first two 16-byte columns).
        // T1's tile has 128 rows -- r, z, 32 zero rows, n -- so that the candidate's input side lands next to W_hn h, not on it.
        {
            auto put_x = [&](unsigned char* tile, int nrows, int r, const double* src) {
                for (int k = 0; k < kFeat; ++k) {
                    const int kb = k >> 6, c = (k & 63) >> 3;
                    reinterpret_cast<__half*>(tile + (size_t)kb * nrows * 128 + r * 128 + ((c ^ (r & 7)) << 4))[k & 7] = __float2half_rn((float)src[k]);
                }
            };
            for (int role = 0; role < 3; ++role) {
                const size_t img = loop_rs_ximage_bytes(role);
                std::vector<unsigned char> hw((size_t)nct[role] * img, 0);
                for (int c = 0; c < nct[role]; ++c) {
                    unsigned char* t0 = hw.data() + (size_t)c * img;
                    if (role < 2) {
                        const std::vector<double>& P = role == 0 ? P1 : P2;
                        for (int u = 0; u < 32; ++u)
                            for (int gt = 0; gt < 3; ++gt)
                                put_x(t0, role == 0 ? 128 : 96, (role == 0 && gt == 2 ? 96 : 32 * gt) + u, &P[(size_t)(gt * H + 32 * c + u) * 113 + 1]);
                    } else {
                        const int FU = H / kRsT3;
                        for (int u = 0; u < FU; ++u) put_x(t0, FU, u, &P3[(size_t)(FU * c + u) * 113 + 1]);
                    }
                }
                CU(e->wRsX[role].ensure(hw.size()));
                CU(cudaMemcpy(e->wRsX[role].p, hw.data(), hw.size(), cudaMemcpyHostToDevice));
            }
        }
        if (rs_raw) {
            // sampler CTAs per group: classes / 128 (4 for 9 bits, 8 for 10) -- fewer, fatter samplers exchange fewer partials;
            // WRNN_RS_QCOLS=64 builds 8 x 64 for 9 bits (measured below)
            int qc = 128;
            if (const char* ev = getenv("WRNN_RS_QCOLS")) qc = (atoi(ev) == 64 && C == 512) ? 64 : 128;
            const int nq = C / qc;
            e->rs_samplers = nq;
            const size_t img = (size_t)qc * 1024;
            std::vector<unsigned char> hw((size_t)nq * img, 0);
            for (int c = 0; c < nq; ++c)
                for (int k = 0; k < qc; ++k) put_tile(hw.data() + (size_t)c * img, qc, k, f3w->data.data() + (size_t)(qc * c + k) * H);
            CU(e->wRs[4].ensure(hw.size()));
            CU(cudaMemcpy(e->wRs[4].p, hw.data(), hw.size(), cudaMemcpyHostToDevice));
        }
    }

    // ---- cluster-local tensor-core loop (MOL): per-CTA streams of weight tiles, gate-major rows (loop_tc2.cu) -------------
    if (C == 30) {
        const size_t img = loop_tc2_image_bytes();
        std::vector<unsigned char> hw((size_t)16 * img, 0);
        const float* Wih2a = hl.data() + oWih2a;
        const float* Wfc1a = hl.data() + oWfc1a;
        const float* Wfc2a = hl.data() + oWfc2a;
        auto put_tile = [&](unsigned char* tile, int nrows, int r, const float* src) {
            for (int kb = 0; kb < 8; ++kb)
                for (int c = 0; c < 8; ++c) {
                    __half* dst = reinterpret_cast<__half*>(tile + (size_t)kb * nrows * 128 + r * 128 + ((c ^ (r & 7)) << 4));
                    for (int i = 0; i < 8; ++i) dst[i] = __float2half_rn(src[kb * 64 + c * 8 + i]);
                }
        };
        for (int cr = 0; cr < 16; ++cr) {
            unsigned char* t0 = hw.data() + (size_t)cr * img;
            unsigned char* t1 = t0 + (size_t)128 * 1024;
            unsigned char* t2 = t1 + (size_t)96 * 1024;
            unsigned char* t3 = t2 + (size_t)128 * 1024;
            unsigned char* t4 = t3 + (size_t)32 * 1024;
            for (int u = 0; u < 32; ++u) {
                const int j = cr * 32 + u;
                for (int gt = 0; gt < 3; ++gt) {
                    put_tile(t0, 128, 32 * gt + u, Wih2a + (size_t)(gt * H + j) * H);
                    put_tile(t1, 96, 32 * gt + u, r1hh->data.data() + (size_t)(gt * H + j) * H);
                    put_tile(t2, 128, 32 * gt + u, r2hh->data.data() + (size_t)(gt * H + j) * H);
                }
                put_tile(t0, 128, 96 + u, Wfc1a + (size_t)j * H);
                put_tile(t2, 128, 96 + u, Wfc1a + (size_t)j * H);
                put_tile(t3, 32, u, Wfc2a + (size_t)j * H);
            }
            for (int c = 0; c < 30; ++c) put_tile(t4, 32, c, f3w->data.data() + (size_t)c * H);   // padded to 32 rows (1024-aligned k-blocks)
        }
        CU(e->wTc2.ensure(hw.size()));
        CU(cudaMemcpy(e->wTc2.p, hw.data(), hw.size(), cudaMemcpyHostToDevice));
    }

    // ---- block-sparse cluster loop: compressed per-CTA images (only for pruned checkpoints) -----------------------
    e->spStride[0] = e->spStride[1] = 0;
    if (e->sparsity >= 0.5) {
        const float* Wih2a = hl.data() + oWih2a;
        const float* Wfc1a = hl.data() + oWfc1a;
        const float* Wfc2a = hl.data() + oWfc2a;
        for (int v = 0; v < 2; ++v) {
            const int CL = v == 0 ? 16 : 8, U = H / CL, CRs = (C + CL - 1) / CL;
            std::vector<std::vector<unsigned char>> imgs(CL);
            size_t stride = 0;
            for (int cr = 0; cr < CL; ++cr) {
                std::vector<const float*> rows[4];
                const int j0 = cr * U;
                for (int gt = 0; gt < 3; ++gt) for (int u = 0; u < U; ++u) rows[0].push_back(r1hh->data.data() + (size_t)(gt * H + j0 + u) * H);
                for (int gt = 0; gt < 3; ++gt) for (int u = 0; u < U; ++u) rows[0].push_back(Wih2a + (size_t)(gt * H + j0 + u) * H);
                for (int u = 0; u < U; ++u) rows[0].push_back(Wfc1a + (size_t)(j0 + u) * H);
                for (int gt = 0; gt < 3; ++gt) for (int u = 0; u < U; ++u) rows[1].push_back(r2hh->data.data() + (size_t)(gt * H + j0 + u) * H);
                for (int u = 0; u < U; ++u) rows[1].push_back(Wfc1a + (size_t)(j0 + u) * H);
                for (int u = 0; u < U; ++u) rows[2].push_back(Wfc2a + (size_t)(j0 + u) * H);
                for (int r = 0; r < CRs; ++r) rows[3].push_back(cr * CRs + r < C ? f3w->data.data() + (size_t)(cr * CRs + r) * H : nullptr);
                std::vector<int> rowptr[4];
                std::vector<unsigned char> col[4];
                std::vector<float> w[4];
                for (int s4 = 0; s4 < 4; ++s4) {
                    rowptr[s4].push_back(0);
                    for (const float* r : rows[s4]) {
                        if (r)
                            for (int gcol = 0; gcol < H / 4; ++gcol) {
                                const float* q = r + 4 * gcol;
                                if (q[0] != 0.f || q[1] != 0.f || q[2] != 0.f || q[3] != 0.f) {
                                    col[s4].push_back((unsigned char)gcol);
                                    w[s4].insert(w[s4].end(), q, q + 4);
                                }
                            }
                        rowptr[s4].push_back((int)col[s4].size());
                    }
                }
                std::vector<unsigned char>& im = imgs[cr];
                im.assign(64, 0);
                int hdr[16] = {0};
                auto align = [&](size_t a) { im.resize((im.size() + a - 1) / a * a, 0); };
                for (int s4 = 0; s4 < 4; ++s4) {
                    align(4); hdr[4 + 3 * s4 + 0] = (int)im.size();
                    im.insert(im.end(), (unsigned char*)rowptr[s4].data(), (unsigned char*)(rowptr[s4].data() + rowptr[s4].size()));
                    hdr[4 + 3 * s4 + 1] = (int)im.size();
                    im.insert(im.end(), col[s4].begin(), col[s4].end());
                    align(16); hdr[4 + 3 * s4 + 2] = (int)im.size();
                    im.insert(im.end(), (unsigned char*)w[s4].data(), (unsigned char*)(w[s4].data() + w[s4].size()));
                    hdr[s4] = (int)col[s4].size();
                }
                align(16);
                memcpy(im.data(), hdr, sizeof(hdr));
                stride = std::max(stride, im.size());
            }
            if (stride + 16384 > e->smem_limit) continue;      // not sparse enough for this cluster size
            std::vector<unsigned char> all((size_t)CL * stride, 0);
            for (int cr = 0; cr < CL; ++cr) memcpy(all.data() + (size_t)cr * stride, imgs[cr].data(), imgs[cr].size());
            CU(e->wSp[v].ensure(all.size()));
            CU(cudaMemcpy(e->wSp[v].p, all.data(), all.size(), cudaMemcpyHostToDevice));
            e->spStride[v] = (int)stride;
        }
    }

    // ---- conditioning weights: BatchNorm folded (eval: (x-mean)*rsqrt(var+eps)*gamma+beta, eps=1e-5) ------
    std::vector<float> hc;
    std::vector<size_t> oW(kCondLayers), oB(kCondLayers);
    auto push = [&](size_t n) { size_t o = hc.size(); hc.resize(o + ((n + 3) & ~(size_t)3), 0.f); return o; };
    auto bn_fold = [&](const std::string& conv, const std::string& bn, int K, size_t& ow, size_t& ob) -> bool {
        const HostTensor* w = nullptr;
        auto it = e->tensors.find(conv + ".weight");
        if (it == e->tensors.end() || (int64_t)it->second.data.size() != (int64_t)128 * K) { e->err = "missing/bad " + conv; return false; }
        w = &it->second;
        const char* parts[4] = {".weight", ".bias", ".running_mean", ".running_var"};
        const HostTensor* q[4];
        for (int i = 0; i < 4; ++i) {
            auto jt = e->tensors.find(bn + parts[i]);
            if (jt == e->tensors.end() || jt->second.data.size() != 128) { e->err = "missing/bad " + bn + parts[i]; return false; }
            q[i] = &jt->second;
        }
        ow = push((size_t)128 * K);
        ob = push(128);
        for (int o = 0; o < 128; ++o) {
            const double sc = (double)q[0]->data[o] / std::sqrt((double)q[3]->data[o] + 1e-5);
            for (int k = 0; k < K; ++k) hc[ow + (size_t)o * K + k] = (float)((double)w->data[(size_t)o * K + k] * sc);
            hc[ob + o] = (float)((double)q[1]->data[o] - (double)q[2]->data[o] * sc);
        }
        return true;
    };
    (void)cin;
    if (!bn_fold("upsample.resnet.conv_in", "upsample.resnet.batch_norm", kFeat * 5, oW[0], oB[0])) return WRNN_ERR_SHAPE;
    for (int i = 0; i < kResBlocks; ++i) {
        const std::string p = "upsample.resnet.layers." + std::to_string(i);
        if (!bn_fold(p + ".conv1", p + ".batch_norm1", 128, oW[1 + 2 * i], oB[1 + 2 * i])) return WRNN_ERR_SHAPE;
        if (!bn_fold(p + ".conv2", p + ".batch_norm2", 128, oW[2 + 2 * i], oB[2 + 2 * i])) return WRNN_ERR_SHAPE;
    }
    oW[kCondLayers - 1] = push(128 * 128);
    oB[kCondLayers - 1] = push(128);
    std::copy(cout->data.begin(), cout->data.end(), hc.begin() + oW[kCondLayers - 1]);
    std::copy(coutb->data.begin(), coutb->data.end(), hc.begin() + oB[kCondLayers - 1]);

    // Projection matrices, unit-major with 4 values per hidden unit j (engine_internal.h, LoopParams):
    //   TA1[f][j] = {c1_r, c1_z, c1_n, c3}   TA2[f][j] = {c2_r, c2_z, c2_n, c4}      from aux[f] (128) + biases
    //   TQ1[q][j] = {q1_r, q1_z, q1_n, q3}   TQ2[q][j] = {q2_r, q2_z, q2_n, 0}       from padded mel frame q (80)
    // c1/q1: rnn1 input gates, c2/q2: rnn2 input gates, c3/q3: fc1, c4: fc2.  b_hh of the r,z gates is folded
    // in (it is added outside the r* product); b_hn stays in the loop.  a1 drops its last channel (Q5).
    const size_t oMA1 = push((size_t)4 * H * 128), obA1 = push(4 * H), oMA2 = push((size_t)4 * H * 128), obA2 = push(4 * H);
    const size_t oMQ1 = push((size_t)4 * H * kFeat), oMQ2 = push((size_t)4 * H * kFeat);
    for (int j = 0; j < H; ++j) {
        for (int g = 0; g < 4; ++g) {
            const size_t rowA1 = oMA1 + (size_t)(j * 4 + g) * 128, rowA2 = oMA2 + (size_t)(j * 4 + g) * 128;
            const size_t rowQ1 = oMQ1 + (size_t)(j * 4 + g) * kFeat, rowQ2 = oMQ2 + (size_t)(j * 4 + g) * kFeat;
            if (g < 3) {
                const int r = g * H + j;
                const double* p1 = &P1[(size_t)r * 113];
                const double* p2 = &P2[(size_t)r * 113];
                for (int c = 0; c < kAux - 1; ++c) { hc[rowA1 + c] = (float)p1[81 + c]; hc[rowA2 + c] = (float)p2[81 + c]; }
                for (int c = 0; c < kFeat; ++c) { hc[rowQ1 + c] = (float)p1[1 + c]; hc[rowQ2 + c] = (float)p2[1 + c]; }
                for (int c = 0; c < kAux; ++c) hc[rowA2 + kAux + c] = r2ih->data[(size_t)r * (H + kAux) + H + c];   // a2
                hc[obA1 + j * 4 + g] = (float)(p1[112] + (double)r1bi->data[r] + (g < 2 ? (double)r1bh->data[r] : 0.0));
                hc[obA2 + j * 4 + g] = (float)(p2[112] + (double)r2bi->data[r] + (g < 2 ? (double)r2bh->data[r] : 0.0));
            } else {
                const double* p3 = &P3[(size_t)j * 113];
                for (int c = 0; c < kAux - 1; ++c) hc[rowA1 + c] = (float)p3[81 + c];
                for (int c = 0; c < kFeat; ++c) hc[rowQ1 + c] = (float)p3[1 + c];
                for (int c = 0; c < kAux; ++c) hc[rowA1 + 2 * kAux + c] = f1w->data[(size_t)j * (H + kAux) + H + c];   // a3
                hc[obA1 + j * 4 + g] = (float)(p3[112] + (double)f1b->data[j]);
                for (int c = 0; c < kAux; ++c) hc[rowA2 + 3 * kAux + c] = f2w->data[(size_t)j * (H + kAux) + H + c];   // a4
                hc[obA2 + j * 4 + g] = f2b->data[j];
            }
        }
    }
    // hi/lo fp16 pairs of every front-end weight matrix, scaled by 2^8, for cond_tc.cu
    {
        std::vector<__half> ht;
        auto pack = [&](const float* W, int N, int K, int Kpad, int idx) {
            e->oCondTc[idx] = ht.size();
            const size_t n = (size_t)N * Kpad;
            ht.resize(ht.size() + 2 * n, __float2half_rn(0.f));
            __half* hi = ht.data() + e->oCondTc[idx];
            __half* lo = hi + n;
            for (int r = 0; r < N; ++r)
                for (int k = 0; k < K; ++k) {
                    const float v = W[(size_t)r * K + k] * 256.0f;
                    const __half h = __float2half_rn(v);
                    hi[(size_t)r * Kpad + k] = h;
                    lo[(size_t)r * Kpad + k] = __float2half_rn(v - __half2float(h));
                }
        };
        {   // conv_in: [128][80][5] -> tap-major [128][5*128]
            std::vector<float> w((size_t)128 * 640, 0.f);
            for (int o = 0; o < 128; ++o)
                for (int c = 0; c < kFeat; ++c)
                    for (int j = 0; j < 5; ++j) w[(size_t)o * 640 + j * 128 + c] = hc[oW[0] + ((size_t)o * kFeat + c) * 5 + j];
            pack(w.data(), 128, 640, 640, 0);
        }
        for (int i = 1; i < kCondLayers; ++i) pack(hc.data() + oW[i], 128, 128, 128, i);
        pack(hc.data() + oMA1, 4 * H, 128, 128, kCondLayers + 0);
        pack(hc.data() + oMA2, 4 * H, 128, 128, kCondLayers + 1);
        pack(hc.data() + oMQ1, 4 * H, kFeat, 128, kCondLayers + 2);
        pack(hc.data() + oMQ2, 4 * H, kFeat, 128, kCondLayers + 3);
        CU(e->wCondTc.ensure(ht.size() * sizeof(__half)));
        CU(cudaMemcpy(e->wCondTc.p, ht.data(), ht.size() * sizeof(__half), cudaMemcpyHostToDevice));
    }
    CU(e->wCond.ensure(hc.size() * sizeof(float)));
    CU(cudaMemcpy(e->wCond.p, hc.data(), hc.size() * sizeof(float), cudaMemcpyHostToDevice));
    float* cb = e->wCond.as<float>();
    for (int i = 0; i < kCondLayers; ++i) { e->dCW[i] = cb + oW[i]; e->dCB[i] = cb + oB[i]; }
    e->dMA1 = cb + oMA1; e->dbA1 = cb + obA1; e->dMA2 = cb + oMA2; e->dbA2 = cb + obA2;
    e->dMQ1 = cb + oMQ1; e->dMQ2 = cb + oMQ2;
    e->finalized = true;
    return WRNN_OK;
#undef GET
}

}  // extern "C"

namespace {

// Runs the conditioning front end for the utterances described by `utts` (already on device in bUtt).
// Tensor-core front end (cond_tc.cu): same contractions, operands as hi/lo fp16 pairs, fp32 accumulation in TMEM.
int run_conditioning_tc(wrnn_engine* e, int n_utts, int rows) {
    const size_t n = (size_t)rows * 128;
    CU(e->bCondH.ensure(8 * n * sizeof(__half) + (size_t)rows * sizeof(float) + 256));
    for (int i = 0; i < 2; ++i) CU(e->bH[i].ensure(n * sizeof(float)));
    CU(e->bAux.ensure(n * sizeof(float)));
    CU(e->bTA1.ensure((size_t)rows * 4 * kRnn * sizeof(float)));
    CU(e->bTA2.ensure((size_t)rows * 4 * kRnn * sizeof(float)));
    CU(e->bTQ1.ensure((size_t)rows * 4 * kRnn * sizeof(float)));
    CU(e->bTQ2.ensure((size_t)rows * 4 * kRnn * sizeof(float)));
    cudaStream_t st = e->stream;
    __half* hb = e->bCondH.as<__half>();
    __half *MPhi = hb, *MPlo = hb + n, *Xhi = hb + 2 * n, *Xlo = hb + 3 * n, *Yhi = hb + 4 * n, *Ylo = hb + 5 * n, *Ahi = hb + 6 * n, *Alo = hb + 7 * n;
    float* rowmask = reinterpret_cast<float*>(hb + 8 * n);
    float* xf[2] = {e->bH[0].as<float>(), e->bH[1].as<float>()};
    const __half* W = e->wCondTc.as<__half>();
    auto whi = [&](int idx) { return W + e->oCondTc[idx]; };
    auto wlo = [&](int idx, size_t cnt) { return W + e->oCondTc[idx] + cnt; };
    CU(launch_mel_split(e->bMel.as<float>(), e->bUtt.as<UttDesc>(), n_utts, rows, MPhi, MPlo, rowmask, st));
    GemmTcArgs g;
    memset(&g, 0, sizeof(g));
    g.M = rows; g.scale = 1.0f / 256.0f; g.status = e->dAbort + 1;
    // conv_in (k=5 as five row-shifted K blocks pairs) + BN + ReLU
    g.N = 128; g.nkb = 10; g.kb_per_tap = 2; g.row_shift = 1; g.relu = 1; g.bias = e->dCB[0]; g.C = xf[0]; g.Chi = Xhi; g.Clo = Xlo;
    CU(launch_gemm_tc_split(MPhi, MPlo, rows, 128, whi(0), wlo(0, (size_t)128 * 640), g, st));
    g.nkb = 2; g.kb_per_tap = 2; g.row_shift = 0;
    int cur = 0;
    for (int i = 0; i < kResBlocks; ++i) {
        g.relu = 1; g.bias = e->dCB[1 + 2 * i]; g.R = nullptr; g.C = nullptr; g.Chi = Yhi; g.Clo = Ylo;
        CU(launch_gemm_tc_split(Xhi, Xlo, rows, 128, whi(1 + 2 * i), wlo(1 + 2 * i, 128 * 128), g, st));
        g.relu = 0; g.bias = e->dCB[2 + 2 * i]; g.R = xf[cur]; g.C = xf[cur ^ 1]; g.Chi = Xhi; g.Clo = Xlo;
        CU(launch_gemm_tc_split(Yhi, Ylo, rows, 128, whi(2 + 2 * i), wlo(2 + 2 * i, 128 * 128), g, st));
        cur ^= 1;
    }
    g.relu = 0; g.bias = e->dCB[kCondLayers - 1]; g.R = nullptr; g.C = e->bAux.as<float>(); g.Chi = Ahi; g.Clo = Alo; g.rowmask = rowmask;
    CU(launch_gemm_tc_split(Xhi, Xlo, rows, 128, whi(kCondLayers - 1), wlo(kCondLayers - 1, 128 * 128), g, st));
    g.rowmask = nullptr; g.Chi = nullptr; g.Clo = nullptr; g.N = 4 * kRnn;
    const size_t pn = (size_t)4 * kRnn * 128;
    g.bias = e->dbA1; g.C = e->bTA1.as<float>();
    CU(launch_gemm_tc_split(Ahi, Alo, rows, 128, whi(kCondLayers + 0), wlo(kCondLayers + 0, pn), g, st));
    g.bias = e->dbA2; g.C = e->bTA2.as<float>();
    CU(launch_gemm_tc_split(Ahi, Alo, rows, 128, whi(kCondLayers + 1), wlo(kCondLayers + 1, pn), g, st));
    g.bias = nullptr; g.C = e->bTQ1.as<float>();
    CU(launch_gemm_tc_split(MPhi, MPlo, rows, 128, whi(kCondLayers + 2), wlo(kCondLayers + 2, pn), g, st));
    g.C = e->bTQ2.as<float>();
    CU(launch_gemm_tc_split(MPhi, MPlo, rows, 128, whi(kCondLayers + 3), wlo(kCondLayers + 3, pn), g, st));
    e->launches += 1 + 2 + 2 * kResBlocks + 4;
    return WRNN_OK;
}

int run_conditioning(wrnn_engine* e, int n_utts, int ta_rows, int tq_rows) {
    CU(e->bX0.ensure((size_t)ta_rows * 400 * sizeof(float)));
    CU(e->bMP.ensure((size_t)tq_rows * kFeat * sizeof(float)));
    for (int i = 0; i < 3; ++i) CU(e->bH[i].ensure((size_t)ta_rows * 128 * sizeof(float)));
    CU(e->bAux.ensure((size_t)ta_rows * 128 * sizeof(float)));
    CU(e->bTA1.ensure((size_t)ta_rows * 4 * kRnn * sizeof(float)));
    CU(e->bTA2.ensure((size_t)ta_rows * 4 * kRnn * sizeof(float)));
    CU(e->bTQ1.ensure((size_t)tq_rows * 4 * kRnn * sizeof(float)));
    CU(e->bTQ2.ensure((size_t)tq_rows * 4 * kRnn * sizeof(float)));
    cudaStream_t st = e->stream;
    const UttDesc* du = e->bUtt.as<UttDesc>();
    CU(launch_im2col(e->bMel.as<float>(), du, n_utts, ta_rows, tq_rows, e->bX0.as<float>(), e->bMP.as<float>(), st));
    float* h[3] = {e->bH[0].as<float>(), e->bH[1].as<float>(), e->bH[2].as<float>()};
    CU(launch_gemm_f32(e->bX0.as<float>(), e->dCW[0], e->dCB[0], nullptr, h[0], ta_rows, 128, 400, 1, st));
    int cur = 0;
    for (int i = 0; i < kResBlocks; ++i) {
        const int mid = (cur + 1) % 3, nxt = (cur + 2) % 3;
        CU(launch_gemm_f32(h[cur], e->dCW[1 + 2 * i], e->dCB[1 + 2 * i], nullptr, h[mid], ta_rows, 128, 128, 1, st));
        CU(launch_gemm_f32(h[mid], e->dCW[2 + 2 * i], e->dCB[2 + 2 * i], h[cur], h[nxt], ta_rows, 128, 128, 0, st));
        cur = nxt;
    }
    CU(launch_gemm_f32(h[cur], e->dCW[kCondLayers - 1], e->dCB[kCondLayers - 1], nullptr, e->bAux.as<float>(), ta_rows, 128, 128, 0, st));
    CU(launch_zero_rows(e->bAux.as<float>(), du, n_utts, st));
    CU(launch_gemm_f32(e->bAux.as<float>(), e->dMA1, e->dbA1, nullptr, e->bTA1.as<float>(), ta_rows, 4 * kRnn, 128, 0, st));
    CU(launch_gemm_f32(e->bAux.as<float>(), e->dMA2, e->dbA2, nullptr, e->bTA2.as<float>(), ta_rows, 4 * kRnn, 128, 0, st));
    CU(launch_gemm_f32(e->bMP.as<float>(), e->dMQ1, nullptr, nullptr, e->bTQ1.as<float>(), tq_rows, 4 * kRnn, kFeat, 0, st));
    CU(launch_gemm_f32(e->bMP.as<float>(), e->dMQ2, nullptr, nullptr, e->bTQ2.as<float>(), tq_rows, 4 * kRnn, kFeat, 0, st));
    e->launches += 3 + 2 * kResBlocks + 1 + 4;   // im2col, conv_in, 20 res convs, conv_out, zero_rows, 4 projections
    return WRNN_OK;
}

int upload_utts(wrnn_engine* e, const float* const* mels, const int32_t* T, int n_utts, int mels_on_device,
                std::vector<UttDesc>& utts, int& ta_rows, int& tq_rows) {
    utts.resize(n_utts);
    long long mel_off = 0;
    ta_rows = 0;
    tq_rows = 0;
    for (int i = 0; i < n_utts; ++i) {
        utts[i].mel_off = mel_off;
        utts[i].T = T[i];
        utts[i].ta_row0 = ta_rows;
        utts[i].tq_row0 = tq_rows;
        utts[i].pad_ = 0;
        mel_off += (long long)kFeat * T[i];
        ta_rows += T[i] + 2 * kPad;   // one row space for frames and padded frames: T frames, then the bias-only row T
        tq_rows += T[i] + 2 * kPad;
    }
    CU(e->bMel.ensure((size_t)mel_off * sizeof(float)));
    CU(e->bUtt.ensure(utts.size() * sizeof(UttDesc)));
    for (int i = 0; i < n_utts; ++i)
        CU(cudaMemcpyAsync(e->bMel.as<float>() + utts[i].mel_off, mels[i], (size_t)kFeat * T[i] * sizeof(float),
                           mels_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->bUtt.p, utts.data(), utts.size() * sizeof(UttDesc), cudaMemcpyHostToDevice, e->stream));
    return WRNN_OK;
}

int ensure_fades(wrnn_engine* e, int overlap) {
    if (e->fade_overlap == overlap) return WRNN_OK;
    // fatchord_version.py:381-391
    const int silence_len = overlap / 2, fade_len = overlap - silence_len;
    std::vector<double> t, fin(overlap, 0.0), fout(overlap, 0.0);
    if (fade_len > 0) np_linspace(-1.0, 1.0, fade_len, t);
    for (int i = 0; i < fade_len; ++i) {
        fin[silence_len + i] = std::sqrt(0.5 * (1.0 + t[i]));
        fout[i] = std::sqrt(0.5 * (1.0 - t[i]));
    }
    CU(e->bFade.ensure((size_t)std::max(1, 2 * overlap) * sizeof(double)));
    if (overlap > 0) {
        CU(cudaMemcpyAsync(e->bFade.p, fin.data(), overlap * sizeof(double), cudaMemcpyHostToDevice, e->stream));
        CU(cudaMemcpyAsync(e->bFade.as<double>() + overlap, fout.data(), overlap * sizeof(double), cudaMemcpyHostToDevice, e->stream));
        CU(cudaStreamSynchronize(e->stream));   // host vectors go out of scope
    }
    e->fade_overlap = overlap;
    return WRNN_OK;
}

float elapsed(cudaEvent_t a, cudaEvent_t b) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

}  // namespace

extern "C" {

int wrnn_generate(wrnn_engine* e, wrnn_request* rq) {
    if (!e || !rq) return WRNN_ERR_INVALID;
    if (!e->finalized) return fail(e, WRNN_ERR_NOT_LOADED, "Please load Wave-RNN in memory before using it");
    if (rq->n_utts < 1 || !rq->mels || !rq->T) return fail(e, WRNN_ERR_INVALID, "wrnn_generate: no utterances");
    if (rq->precision != WRNN_PREC_F32 && rq->precision != WRNN_PREC_F16 && rq->precision != WRNN_PREC_SPARSE_F32)
        return fail(e, WRNN_ERR_INVALID, "wrnn_generate: unknown precision");
    const bool use_sparse = rq->precision == WRNN_PREC_SPARSE_F32;
    if (use_sparse && !e->spStride[0] && !e->spStride[1])
        return fail(e, WRNN_ERR_INVALID, "block-sparse loop needs a pruned checkpoint (>= 50 % zero 1x4 groups that fit one cluster)");
    const bool use_tc = rq->precision == WRNN_PREC_F16;
    const bool is_rr = e->topology == WRNN_TOPO_RUNTIMERACER, is_gn = e->topology == WRNN_TOPO_GENEING;
    if ((is_rr || is_gn) && rq->precision != WRNN_PREC_F32)
        return fail(e, WRNN_ERR_INVALID, "the runtimeracer / geneing topologies run the fp32 loop only");
    if (use_tc && !e->wTc.p) return fail(e, WRNN_ERR_INVALID, "tensor-core loop supports RAW 9/10-bit and MOL only");
    CU(cudaSetDevice(e->device));
    const int n_utts = rq->n_utts;
    const bool partial = (rq->fold_begin != 0 || rq->fold_end != 0);
    if ((partial || rq->forced || rq->samples || rq->logits) && n_utts != 1)
        return fail(e, WRNN_ERR_INVALID, "fold ranges / forced / samples / logits need n_utts == 1");
    const bool want_wav = rq->wav != nullptr && !partial && rq->max_steps == 0;
    for (int i = 0; i < n_utts; ++i) {
        if (rq->T[i] < 1) return fail(e, WRNN_ERR_INVALID, "empty mel");
        if (want_wav && rq->T[i] <= 20)
            return fail(e, WRNN_ERR_TOO_SHORT, "operands could not be broadcast together: mel has <= 20 frames "
                                                "(fade-out needs 4000 samples, fatchord_version.py:253-255)");
    }
    const int target = rq->target, overlap = rq->overlap;
    if (rq->batched) {
        if (target < 0 || overlap < 0 || target + overlap <= 0) return fail(e, WRNN_ERR_INVALID, "bad target/overlap");
        if (overlap == 0 && want_wav)
            return fail(e, WRNN_ERR_INVALID, "overlap == 0: xfade_and_unfold raises ValueError (fatchord_version.py:394)");
    }
    const int mu_law = (e->mode == WRNN_MODE_RAW) ? rq->mu_law : 0;   // fatchord_version.py:156

    cudaStream_t st = e->stream;
    CU(cudaEventRecord(e->ev[0], st));
    std::vector<UttDesc> utts;
    int ta_rows = 0, tq_rows = 0;
    int rc = upload_utts(e, rq->mels, rq->T, n_utts, rq->mels_on_device, utts, ta_rows, tq_rows);
    if (rc) return rc;
    CU(cudaEventRecord(e->ev[1], st));
    CU(cudaMemsetAsync(e->dAbort, 0, 2 * sizeof(int), st));
    rc = is_rr ? run_conditioning_rr(e, e->rr, n_utts, ta_rows, tq_rows)
       : is_gn ? run_conditioning_gn(e, e->gn, n_utts, ta_rows, tq_rows)
               : (use_tc ? run_conditioning_tc(e, n_utts, ta_rows) : run_conditioning(e, n_utts, ta_rows, tq_rows));
    if (rc) return rc;
    CU(cudaEventRecord(e->ev[2], st));

    // ---- fold plan (fatchord_version.py:315-340): integer arithmetic on the host, O(folds) -------------------
    std::vector<FoldDesc> folds;
    std::vector<PostUtt> putts(n_utts);
    int S = 0;
    long long wav_total = 0;
    int max_wave_len = 0;
    for (int i = 0; i < n_utts; ++i) {
        const int N = utts[i].T * kHop;
        int64_t F = 1, padded = N;
        if (rq->batched) {
            wrnn_fold_plan(N, target, overlap, &F, &padded);
            if (F < 1) return fail(e, WRNN_ERR_INVALID, "utterance shorter than the fold overlap");
            S = target + 2 * overlap;
        } else {
            S = std::max(S, N);
        }
        int f0 = 0, f1 = (int)F;
        if (partial) { f0 = std::max(0, rq->fold_begin); f1 = std::min((int)F, rq->fold_end); }
        putts[i].samp_off = (long long)folds.size();   // fold index for now; scaled by S below
        putts[i].F = (int)F;
        putts[i].wave_len = (utts[i].T - 1) * kHop;
        putts[i].wav_off = wav_total;
        wav_total += putts[i].wave_len;
        max_wave_len = std::max(max_wave_len, putts[i].wave_len);
        for (int f = f0; f < f1; ++f) {
            FoldDesc d;
            d.ta_row0 = utts[i].ta_row0; d.tq_row0 = utts[i].tq_row0; d.T = utts[i].T; d.N = N;
            d.n0 = rq->batched ? f * (target + overlap) : 0;
            d.utt = rq->utt_index0 + i; d.fold = f; d.pad_ = 0;
            folds.push_back(d);
        }
    }
    const int S_full = S;
    if (rq->max_steps > 0) S = std::min(S, rq->max_steps);
    const int Btot = (int)folds.size();
    if (Btot < 1) return fail(e, WRNN_ERR_INVALID, "empty fold range");
    for (int i = 0; i < n_utts; ++i) putts[i].samp_off *= S;
    if (rq->wav_offsets) {
        for (int i = 0; i < n_utts; ++i) rq->wav_offsets[i] = putts[i].wav_off;
        rq->wav_offsets[n_utts] = wav_total;
    }
    if (want_wav && rq->wav_capacity < wav_total) return fail(e, WRNN_ERR_INVALID, "wav buffer too small");

    CU(e->bFolds.ensure(folds.size() * sizeof(FoldDesc)));
    CU(cudaMemcpyAsync(e->bFolds.p, folds.data(), folds.size() * sizeof(FoldDesc), cudaMemcpyHostToDevice, st));
    CU(e->bSamples.ensure((size_t)Btot * S * sizeof(float)));
    if (rq->logits) CU(e->bLogits.ensure((size_t)Btot * S * e->C * sizeof(float)));
    if (rq->forced) {
        CU(e->bForced.ensure((size_t)Btot * S_full * sizeof(float)));
        // forced is (F,S_full) on the host; the kernel indexes [b][S] with the launch S
        if (S == S_full) {
            CU(cudaMemcpyAsync(e->bForced.p, rq->forced, (size_t)Btot * S * sizeof(float), cudaMemcpyHostToDevice, st));
        } else {
            CU(cudaMemcpy2DAsync(e->bForced.p, (size_t)S * sizeof(float), rq->forced, (size_t)S_full * sizeof(float),
                                 (size_t)S * sizeof(float), Btot, cudaMemcpyHostToDevice, st));
        }
    }

    // ---- the loop: waves of <= kMaxFoldsPerLaunch folds -----------------------------------------------------
    rq->n_launches = 0;
    const size_t words_per_fold = (size_t)4 * kRnn + e->Cpad + 2;
    *e->hProgress = 0;
    auto t_start = std::chrono::steady_clock::now();
    float ms_expand = 0.f;
    int wave = is_rr ? kRrMaxFolds : is_gn ? kGnMaxFolds : (use_tc ? kTcMaxFolds : (use_sparse ? 4096 : kMaxFoldsPerLaunch));
    // role-specialised loop: as many 48-CTA groups as the device holds; WRNN_RS=0 keeps loop_tc.cu for every fold count
    const int rs_samplers = (e->mode == WRNN_MODE_RAW && e->wRs[4].p) ? e->rs_samplers : 0;      // RAW: sampler CTAs per group
    const int rs_ctas = kRsCtas + rs_samplers;
    const int rs_groups_max = std::max(0, e->n_sms / rs_ctas);
    const int rs_max_folds = (getenv("WRNN_RS") && atoi(getenv("WRNN_RS")) == 0) ? 0 : rs_groups_max * kRsMaxFoldsPerGroup;
    // MOL, one to two role-specialised launches' worth of folds: two balanced waves of loop_rs beat one launch of loop_tc
    // (measured, us per step over all folds: 385 folds 2 x 12.8 vs 27.0, 512: 2 x 13.0 vs 28.6, 766: 2 x 14.1 vs 29.9; three waves
    // lose at 1024: 3 x 13.9 vs 36.6 -- tools/wave_crossover.py), and loop_rs keeps no per-sample conditioning records
    if (use_tc && !is_rr && !is_gn && !use_sparse && e->mode == WRNN_MODE_MOL && e->wRs[0].p && e->wRsX[0].p && rs_max_folds > 0 &&
        Btot > rs_max_folds && Btot <= 2 * rs_max_folds && !(getenv("WRNN_RS_WAVES") && atoi(getenv("WRNN_RS_WAVES")) == 0))
        wave = (Btot + 1) / 2;
    bool rs_tables = false;
    for (int w0 = 0; w0 < Btot; w0 += wave) {
        const int B = std::min(wave, Btot - w0);
        CU(cudaMemsetAsync(e->dAbort, 0, sizeof(int), st));
        bool expanded = false;
        if (is_rr) {
            // ---- runtimeracer topology: fp32 weight-stationary loop (loop_rr.cu), <= 64 folds per launch ------------------------
            const size_t words = (size_t)B * (6 * kRrH + e->Cpad + 2);
            CU(e->bExch.ensure(words * sizeof(unsigned long long)));
            CU(cudaMemsetAsync(e->bExch.p, 0, words * sizeof(unsigned long long), st));
            RrLoopParams p;
            memset(&p, 0, sizeof(p));
            for (int k = 0; k < 4; ++k) p.Whh[k] = e->rr.Whh[k];
            for (int k = 0; k < 3; ++k) p.Wih[k] = e->rr.Wih[k];
            p.M12 = e->rr.M12; p.M34 = e->rr.M34; p.Wfc5 = e->rr.Wfc5; p.u = e->rr.u; p.bhn = e->rr.bhn; p.bfc5 = e->rr.bfc5;
            p.TA = e->bTA1.as<float4>(); p.TQ = e->bTQ1.as<float4>(); p.coef = e->rr.coef;
            p.folds = e->bFolds.as<FoldDesc>() + w0;
            p.B = B; p.S = S; p.C = e->C; p.Cpad = e->Cpad; p.CR = e->CR; p.mode = e->mode; p.seed = rq->seed;
            unsigned long long* x = e->bExch.as<unsigned long long>();
            for (int k = 0; k < 4; ++k) { p.bH[k] = x; x += (size_t)B * kRrH; }
            p.bY2 = x; x += (size_t)B * kRrH; p.bY4 = x; x += (size_t)B * kRrH;
            p.bLG = x; x += (size_t)B * e->Cpad; p.bX = x;
            p.samples = e->bSamples.as<float>() + (size_t)w0 * S;
            p.logits_out = rq->logits ? e->bLogits.as<float>() + (size_t)w0 * S * e->C : nullptr;
            p.forced = rq->forced ? e->bForced.as<float>() + (size_t)w0 * S : nullptr;
            p.progress = e->dProgress;
            p.abort_flag = e->dAbort;
            if (loop_rr_smem_bytes(B, e->CR) > e->smem_limit) return fail(e, WRNN_ERR_INVALID, "shared memory budget exceeded");
            CU(launch_loop_rr(p, st));
            rq->loop_kernel = WRNN_LOOP_RR;
            e->launches += 1;
        } else if (is_gn) {
            // ---- geneing topology: fp32 weight-stationary loop (loop_gn.cu) ----------------------------------------------------
            const size_t words = (size_t)B * (kGnH + kGnFc + e->Cpad + 2);
            CU(e->bExch.ensure(words * sizeof(unsigned long long)));
            CU(cudaMemsetAsync(e->bExch.p, 0, words * sizeof(unsigned long long), st));
            GnLoopParams p;
            memset(&p, 0, sizeof(p));
            p.Whh = e->gn.Whh; p.Wfc1a = e->gn.Wfc1a; p.Wfc3 = e->gn.Wfc3; p.u1 = e->gn.u1; p.u2 = e->gn.u2; p.bhn = e->gn.bhn; p.bfc3 = e->gn.bfc3;
            p.TA = e->bTA1.as<float4>(); p.TQ = e->bTQ1.as<float4>(); p.coef = e->gn.coef;
            p.folds = e->bFolds.as<FoldDesc>() + w0;
            p.B = B; p.S = S; p.C = e->C; p.Cpad = e->Cpad; p.CR = e->CR; p.mode = e->mode; p.seed = rq->seed;
            unsigned long long* x = e->bExch.as<unsigned long long>();
            p.bH = x; x += (size_t)B * kGnH; p.bF = x; x += (size_t)B * kGnFc; p.bLG = x; x += (size_t)B * e->Cpad; p.bX = x;
            p.samples = e->bSamples.as<float>() + (size_t)w0 * S;
            p.logits_out = rq->logits ? e->bLogits.as<float>() + (size_t)w0 * S * e->C : nullptr;
            p.forced = rq->forced ? e->bForced.as<float>() + (size_t)w0 * S : nullptr;
            p.progress = e->dProgress;
            p.abort_flag = e->dAbort;
            if (loop_gn_smem_bytes(B, e->CR) > e->smem_limit) return fail(e, WRNN_ERR_INVALID, "shared memory budget exceeded");
            CU(launch_loop_gn(p, st));
            rq->loop_kernel = WRNN_LOOP_GN;
            e->launches += 1;
        } else if (use_sparse) {
            // ---- block-sparse cluster loop: folds partitioned over independent 16-CTA (or 8-CTA) clusters ------------------
            SparseParams sp;
            memset(&sp, 0, sizeof(sp));
            sp.v1 = e->dv1; sp.v2 = e->dv2; sp.v3 = e->dv3; sp.bhn1 = e->dbhn1; sp.bhn2 = e->dbhn2; sp.bfc3 = e->dbfc3;
            sp.TA1 = e->bTA1.as<float4>(); sp.TA2 = e->bTA2.as<float4>(); sp.TQ1 = e->bTQ1.as<float4>(); sp.TQ2 = e->bTQ2.as<float4>();
            sp.coef = e->dcoef;
            sp.folds = e->bFolds.as<FoldDesc>() + w0;
            sp.B = B; sp.S = S; sp.C = e->C; sp.Cpad = (e->C + 3) & ~3; sp.mode = e->mode; sp.seed = rq->seed;
            sp.samples = e->bSamples.as<float>() + (size_t)w0 * S;
            sp.logits_out = rq->logits ? e->bLogits.as<float>() + (size_t)w0 * S * e->C : nullptr;
            sp.forced = rq->forced ? e->bForced.as<float>() + (size_t)w0 * S : nullptr;
            sp.progress = e->dProgress;
            cudaError_t lerr = cudaErrorInvalidValue;
            for (int v = 0; v < 2 && lerr != cudaSuccess; ++v) {
                if (!e->spStride[v]) continue;
                const int CL = v == 0 ? 16 : 8;
                sp.wimg = e->wSp[v].as<unsigned char>(); sp.img_stride = e->spStride[v]; sp.CRs = (e->C + CL - 1) / CL;
                int Bc = std::min(8, std::max(1, (B + 8) / 9));
                while (Bc > 1 && loop_sparse_smem_bytes(CL, sp.img_stride, Bc, sp.Cpad, sp.CRs) > e->smem_limit) --Bc;
                if (loop_sparse_smem_bytes(CL, sp.img_stride, Bc, sp.Cpad, sp.CRs) > e->smem_limit) continue;
                sp.Bc = Bc;
                lerr = launch_loop_sparse(sp, CL, (B + Bc - 1) / Bc, st);
                if (lerr != cudaSuccess) cudaGetLastError();      // e.g. 16-CTA clusters not schedulable: try 8
            }
            if (lerr != cudaSuccess) return fail(e, WRNN_ERR_CUDA, std::string("block-sparse loop launch: ") + cudaGetErrorString(lerr));
            e->launches += 1;
            rq->loop_kernel = WRNN_LOOP_SPARSE;
        } else if (use_tc && (e->mode == WRNN_MODE_MOL || rs_samplers > 0) && e->wRs[0].p && B <= rs_max_folds) {
            // ---- role-specialised tensor-core loop (loop_rs.cu): the latency-bound regime, <= 128 folds per 48-CTA group ---------
            // groups: two leave 52 SMs to the expanders (three groups run the loop 4 % faster but starve them)
            // inline conditioning (default; WRNN_RS_INLINE=0: records from expander CTAs): no per-sample records at all -- the aux share is
            // a per-frame row, the mel share one more K = 80 slab of the on-path MMAs -- so every SM can belong to a group
            const bool inl = e->wRsX[0].p && !(getenv("WRNN_RS_INLINE") && atoi(getenv("WRNN_RS_INLINE")) == 0);
            const int G_min = (B + kRsMaxFoldsPerGroup - 1) / kRsMaxFoldsPerGroup, G_max = std::max(G_min, std::min(rs_groups_max, inl ? 3 : 2));
            int G_def = std::max(G_min, std::min(G_max, (B + 63) / 64));
            if (G_max >= 3 && B > 96) G_def = std::max(G_def, 3);      // measured (profiles/r2_probes/cal_trials_2gpu.txt): 106 folds 11.96 us in three groups, 12.44 in two; 68 folds 12.18 vs 11.99
            if (const char* ev = getenv("WRNN_RS_GROUPS")) G_def = std::max(G_min, std::min(rs_groups_max, atoi(ev)));
            const int pad_def = (getenv("WRNN_RS_PAD") && atoi(getenv("WRNN_RS_PAD")) == 0) ? 0 : 1;
            // Layout calibration.  Which SMs a group gets is the block scheduler's choice, it depends on the GPU and on the grid size, and
            // the step time follows it: the same binary runs the 137-fold RAW step in 15.3 us or in 17.6 us.  So the first
            // call of a shape on an engine times kRsCalSteps steps of up to four layouts -- the default group count and its neighbour,
            // each with the grid padded to the SM count and not -- and the engine keeps the fastest.  The result does not depend on the
            // layout (tests: bit-identical samples across groups and padding).  WRNN_RS_CALIBRATE=0: the default layout always.
            constexpr int kRsCalSteps = 256;
            struct RsLayout { int G, pad; };
            std::vector<RsLayout> cands;
            const bool can_cal = inl && S >= 8 * kRsCalSteps && !getenv("WRNN_RS_GROUPS") && !getenv("WRNN_RS_PAD") && !getenv("WRNN_RS_TRACE") &&
                                 !getenv("WRNN_RS_DEBUG") && !getenv("WRNN_RS_PLACE") && !(getenv("WRNN_RS_CALIBRATE") && atoi(getenv("WRNN_RS_CALIBRATE")) == 0);
            const int cal_key = (e->mode << 24) | (rs_samplers << 16) | ((B + 15) / 16);
            RsLayout chosen{G_def, pad_def};
            bool have_choice = !can_cal;
            if (can_cal) {
                auto it = e->rs_layout.find(cal_key);
                if (it != e->rs_layout.end()) { chosen = RsLayout{it->second.first, it->second.second}; have_choice = true; }
                else {
                    const int G_alt = G_def + 1 <= G_max ? G_def + 1 : (G_def - 1 >= G_min ? G_def - 1 : G_def);
                    for (int k = 0; k < (G_alt == G_def ? 1 : 2); ++k)
                        for (int pad : {1, 0}) {
                            const int g = k == 0 ? G_def : G_alt;
                            if (pad == 0 && e->n_sms - g * rs_ctas < 8) continue;        // (padding a handful of CTAs changes nothing)
                            cands.push_back(RsLayout{g, pad});
                        }
                    if (cands.size() < 2) have_choice = true;
                }
            }
            // (trial 0 is an unscored warm-up of the default layout: the first launch of the kernel in a process, cold tables and cold
            //  instruction caches cost the first trial several per cent -- measured: a 2-GPU run kept two groups for 213 folds because of it)
            const int n_cal = have_choice ? 0 : (int)cands.size() + 1;
            float best_ms = 0.f;
            for (int trial = 0; trial <= n_cal; ++trial) {
            const bool cal = trial < n_cal;
            const RsLayout lay = cal ? cands[trial == 0 ? 0 : trial - 1] : chosen;
            const int G = lay.G;
            const int S_run = cal ? kRsCalSteps : S;
            const int Ng = (B + G - 1) / G;
            // Conditioning records (16 KB per fold and step) are produced INSIDE the loop kernel by expander CTAs on the SMs the
            // groups leave free, into a ring sized to stay resident in L2 (WRNN_RS_RING_MB, default 48 MB: three 4-step chunks at 213 folds): produced / consumed
            // counters per chunk of kRsChunk steps order the two sides.  WRNN_RS_EXPAND=0 (or no SM left): the whole table is
            // expanded before the launch.
            const int n_exp = e->n_sms - G * rs_ctas;
            const bool ring = !inl && n_exp >= 1 && !(getenv("WRNN_RS_EXPAND") && atoi(getenv("WRNN_RS_EXPAND")) == 0);
            const int nchunks = (S + kRsChunk - 1) / kRsChunk;
            int cs_steps = nchunks * kRsChunk;
            if (ring) {
                const size_t chunk_bytes = (size_t)G * Ng * 4096 * sizeof(float) * kRsChunk;
                const size_t budget = (size_t)(getenv("WRNN_RS_RING_MB") ? atoll(getenv("WRNN_RS_RING_MB")) : 48) << 20;
                const int ring_chunks = (int)std::min<size_t>((size_t)nchunks, std::max<size_t>(3, budget / chunk_bytes));
                cs_steps = ring_chunks * kRsChunk;
            }
            const size_t cs_bytes = inl ? 0 : (size_t)G * cs_steps * Ng * 4096 * sizeof(float);
            if (cs_bytes > ((size_t)96 << 30)) return fail(e, WRNN_ERR_INVALID, "per-sample conditioning table would exceed 96 GiB");
            if (!inl) CU(e->bCS.ensure(cs_bytes));
            CU(e->bCsDone.ensure((size_t)2 * nchunks * sizeof(unsigned int)));
            CU(cudaMemsetAsync(e->bCsDone.p, 0, (size_t)2 * nchunks * sizeof(unsigned int), st));
            CU(cudaEventRecord(e->evx[0], st));
            if (inl) {
                if (!rs_tables) {      // once per call (the waves of a long batch share them)
                    CU(e->bFR.ensure((size_t)ta_rows * 4096 * sizeof(float)));
                    CU(e->bM16.ensure(((size_t)tq_rows * kHop + 16) * kFeat * sizeof(__half)));
                    CU(launch_rs_inline_tables(e->bTA1.as<float4>(), e->bTA2.as<float4>(), e->bMel.as<float>(), e->bUtt.as<UttDesc>(), n_utts, e->dcoef,
                                               ta_rows, e->bFR.as<float>(), e->bM16.as<__half>(), st));
                    rs_tables = true;
                    e->launches += 1;
                }
            } else if (!ring)
                CU(launch_expand_cond_rs(e->bTA1.as<float4>(), e->bTA2.as<float4>(), e->bTQ1.as<float4>(), e->bTQ2.as<float4>(), e->dcoef,
                                         e->bFolds.as<FoldDesc>() + w0, B, S, Ng, cs_steps, e->bCS.as<float>(), st));
            CU(cudaEventRecord(e->evx[1], st));
            expanded = true;
            // exchange matrices (all bytes 0xFF), then the sample words and (RAW) the soft-max partial words (zero)
            const size_t xbytes = loop_rs_exchange_bytes(G);
            const size_t wbytes = (size_t)G * 128 * sizeof(unsigned long long) * (1 + 2 * kRsMaxSamplers * 2);
            CU(e->bRsExch.ensure(xbytes + wbytes));
            CU(cudaMemsetAsync(e->bRsExch.p, 0xFF, xbytes, st));
            CU(cudaMemsetAsync(e->bRsExch.as<unsigned char>() + xbytes, 0, wbytes, st));
            RsParams rp;
            memset(&rp, 0, sizeof(rp));
            rp.w1 = e->wRs[0].as<unsigned char>(); rp.w2 = e->wRs[1].as<unsigned char>();
            rp.w3 = e->wRs[2].as<unsigned char>(); rp.w4 = e->wRs[3].as<unsigned char>();
            rp.v1 = e->dv1; rp.v2 = e->dv2; rp.v3 = e->dv3; rp.bhn1 = e->dbhn1; rp.bhn2 = e->dbhn2; rp.bfc3 = e->dbfc3;
            rp.CS = e->bCS.as<float>(); rp.cs_steps = cs_steps; rp.Ng = Ng; rp.G = G;
            if (inl) {
                rp.inl = 1;
                rp.wx1 = e->wRsX[0].as<unsigned char>(); rp.wx2 = e->wRsX[1].as<unsigned char>(); rp.wx3 = e->wRsX[2].as<unsigned char>();
                rp.FR = e->bFR.as<float>(); rp.M16 = e->bM16.as<__half>(); rp.m16_zero = (long long)tq_rows * kHop;
                rp.n_expanders = lay.pad ? std::max(0, n_exp) : 0;
            }
            if (ring) {
                rp.cs_done = e->bCsDone.as<unsigned int>(); rp.cs_consumed = e->bCsDone.as<unsigned int>() + nchunks;
                rp.CSw = e->bCS.as<float>();
                rp.TA1 = e->bTA1.as<float4>(); rp.TA2 = e->bTA2.as<float4>(); rp.TQ1 = e->bTQ1.as<float4>(); rp.TQ2 = e->bTQ2.as<float4>();
                rp.coef = e->dcoef;
                rp.n_expanders = n_exp;
                if (const char* ev = getenv("WRNN_RS_EXPANDERS")) rp.n_expanders = std::max(1, std::min(n_exp, atoi(ev)));
            }
            rp.offpath_delay_ns = getenv("WRNN_RS_DELAY_NS") ? atoi(getenv("WRNN_RS_DELAY_NS")) : 2500;
            rp.canary_all = getenv("WRNN_RS_CANARY_ALL") ? atoi(getenv("WRNN_RS_CANARY_ALL")) : 0;
            rp.folds = e->bFolds.as<FoldDesc>() + w0;
            rp.B = B; rp.S = S_run; rp.seed = rq->seed;
            rp.X = e->bRsExch.as<uint4>();
            rp.bX = reinterpret_cast<unsigned long long*>(e->bRsExch.as<unsigned char>() + xbytes);
            rp.bP = rp.bX + (size_t)G * 128;
            rp.w5 = e->wRs[4].as<unsigned char>();
            rp.mode = e->mode; rp.C = e->C; rp.n_samplers = rs_samplers; rp.ctas = rs_ctas; rp.qcols = rs_samplers ? e->C / rs_samplers : 0;
            rp.samples = e->bSamples.as<float>() + (size_t)w0 * S;
            rp.logits_out = (rq->logits && !cal) ? e->bLogits.as<float>() + (size_t)w0 * S * e->C : nullptr;
            rp.forced = rq->forced ? e->bForced.as<float>() + (size_t)w0 * S : nullptr;
            rp.progress = cal ? nullptr : e->dProgress;
            rp.abort_flag = e->dAbort;
            int* hdbg = nullptr;
            if (getenv("WRNN_RS_DEBUG") && atoi(getenv("WRNN_RS_DEBUG"))) {     // checkpoints in mapped host memory: readable while the kernel hangs
                CU(cudaHostAlloc(&hdbg, (size_t)G * rs_ctas * 32 * sizeof(int), cudaHostAllocMapped));
                memset(hdbg, 0, (size_t)G * rs_ctas * 32 * sizeof(int));
                CU(cudaHostGetDevicePointer(&rp.dbg, hdbg, 0));
            }
            const char* rs_trace = getenv("WRNN_RS_TRACE");
            const size_t trace_n = (size_t)G * rs_ctas * 8 * 48;
            if (rs_trace) {
                CU(e->bFloor.ensure(trace_n * sizeof(unsigned long long)));
                CU(cudaMemsetAsync(e->bFloor.p, 0, trace_n * sizeof(unsigned long long), st));
                rp.trace = e->bFloor.as<unsigned long long>();
            }
            // the ring is the only large buffer the kernel rewrites: pin it in L2 for the launch (persisting access window)
            bool l2_window = false;
            if (ring && !(getenv("WRNN_RS_L2_PERSIST") && atoi(getenv("WRNN_RS_L2_PERSIST")) == 0)) {
                int max_win = 0, max_persist = 0;
                cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, e->device);
                cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, e->device);
                if (max_win > 0 && max_persist > 0) {
                    const size_t want = std::min<size_t>(cs_bytes, (size_t)max_persist);
                    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
                        cudaStreamAttrValue av;
                        memset(&av, 0, sizeof(av));
                        av.accessPolicyWindow.base_ptr = e->bCS.p;
                        av.accessPolicyWindow.num_bytes = std::min<size_t>(cs_bytes, (size_t)max_win);
                        av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)want / (double)av.accessPolicyWindow.num_bytes);
                        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                        av.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
                        l2_window = cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av) == cudaSuccess;
                    }
                    cudaGetLastError();
                }
            }
            // physical placement (loop_rs.cu: p.place): WRNN_RS_PLACE=1 [WRNN_RS_ROT=k]; not with the timeline / checkpoints (indexed by blockIdx)
            if (getenv("WRNN_RS_PLACE") && atoi(getenv("WRNN_RS_PLACE")) && !rp.trace && !rp.dbg) {
                const int grid = G * rs_ctas + ((rp.cs_done || rp.inl) ? rp.n_expanders : 0);
                CU(e->bRsPlace.ensure((size_t)grid * sizeof(unsigned int)));
                CU(cudaMemsetAsync(e->bRsPlace.p, 0, (size_t)grid * sizeof(unsigned int), st));
                rp.place = e->bRsPlace.as<unsigned int>();
                rp.rot = getenv("WRNN_RS_ROT") ? std::max(0, atoi(getenv("WRNN_RS_ROT"))) % grid : 0;
            }
            if (cal) CU(cudaEventRecord(e->evc[0], st));
            CU(launch_loop_rs(rp, st));
            if (l2_window) {
                cudaStreamAttrValue av;
                memset(&av, 0, sizeof(av));
                av.accessPolicyWindow.num_bytes = 0;
                cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av);
                cudaGetLastError();
            }
            if (cal) {        // one trial: its time, and -- after the last -- the layout this engine keeps for the shape
                CU(cudaEventRecord(e->evc[1], st));
                int ab = 0;
                CU(cudaMemcpyAsync(&ab, e->dAbort, sizeof(int), cudaMemcpyDeviceToHost, st));
                CU(cudaStreamSynchronize(st));
                if (ab) return fail(e, WRNN_ERR_TIMEOUT, "sample loop deadlock guard fired (layout calibration)");
                const float ms = elapsed(e->evc[0], e->evc[1]);
                if (getenv("WRNN_VERBOSE")) fprintf(stderr, "[wrnn] rs layout %s: %d folds, %d groups, padded %d: %.2f us per step\n", trial == 0 ? "warm-up" : "trial", B, lay.G, lay.pad, ms * 1e3 / kRsCalSteps);
                if (trial == 1 || (trial > 1 && ms < 0.975f * best_ms)) { best_ms = ms; chosen = lay; }      // (the default layout -- trial 1 -- unless another is clearly faster)
                e->launches += 1;
                if (trial == n_cal - 1) e->rs_layout[cal_key] = std::make_pair(chosen.G, chosen.pad);
                continue;
            }
            rq->loop_kernel = WRNN_LOOP_RS;
            if (rs_trace) {
                std::vector<unsigned long long> tr(trace_n);
                CU(cudaMemcpyAsync(tr.data(), e->bFloor.p, trace_n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
                CU(cudaStreamSynchronize(st));
                if (FILE* f = fopen(rs_trace, "w")) {
                    for (int c = 0; c < G * rs_ctas; ++c)
                        for (int k = 0; k < 8; ++k) {
                            fprintf(f, "%d %d", c, k);
                            for (int j = 0; j < 48; ++j) fprintf(f, " %llu", tr[((size_t)c * 8 + k) * 48 + j]);
                            fprintf(f, "\n");
                        }
                    fclose(f);
                }
            }
            if (hdbg) {
                CU(cudaEventRecord(e->ev[7], st));
                const auto t_dbg = std::chrono::steady_clock::now();
                bool dumped = false;
                while (cudaEventQuery(e->ev[7]) == cudaErrorNotReady) {
                    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_dbg).count();
                    if (dt > 3.0 && !dumped) {
                        dumped = true;
                        for (int c = 0; c < G * rs_ctas; ++c) {
                            fprintf(stderr, "[rs dbg] cta %3d:", c);
                            for (int w = 0; w < 20; ++w) fprintf(stderr, " %d:%02x", hdbg[c * 32 + w] >> 8, hdbg[c * 32 + w] & 0xFF);
                            fprintf(stderr, "\n");
                        }
                        fflush(stderr);
                    }
                    std::this_thread::sleep_for(std::chrono::milliseconds(10));
                }
                cudaFreeHost(hdbg);
            }
            e->launches += 2;
            }     // (layout trials, then the run)
        } else if (use_tc && e->mode == WRNN_MODE_MOL && e->wTc2.p && getenv("WRNN_TC_V2") && atoi(getenv("WRNN_TC_V2"))) {
            // (opt-in: measured equal to loop_tc.cu in round 1 -- 27.4 vs 26.6 us/step -- see DESIGN.md section 4.3)
            // ---- cluster-local tensor-core loop (MOL): folds partitioned over independent 16-CTA clusters -----------------
            const int max_cl = std::max(1, loop_tc2_max_clusters());       // 16-CTA clusters that can be co-resident
            if (getenv("WRNN_VERBOSE")) fprintf(stderr, "[wrnn] tc2: max co-resident 16-CTA clusters = %d\n", max_cl);
            const int ncl = std::max(std::min(max_cl, (B + 7) / 8), (B + 31) / 32);
            const int Bc = (B + ncl - 1) / ncl;                    // <= 32 because B <= 256
            const size_t cs_bytes = (size_t)ncl * S * 16 * 4 * 32 * 32 * 2 * sizeof(float);
            if (cs_bytes > ((size_t)96 << 30)) return fail(e, WRNN_ERR_INVALID, "per-sample conditioning table would exceed 96 GiB");
            CU(e->bCS.ensure(cs_bytes));
            CU(cudaEventRecord(e->evx[0], st));
            CU(launch_expand_cond2(e->bTA1.as<float4>(), e->bTA2.as<float4>(), e->bTQ1.as<float4>(), e->bTQ2.as<float4>(), e->dcoef,
                                   e->bFolds.as<FoldDesc>() + w0, B, Bc, ncl, S, e->bCS.as<float>(), st));
            CU(cudaEventRecord(e->evx[1], st));
            expanded = true;
            Tc2Params tp;
            memset(&tp, 0, sizeof(tp));
            tp.wimg = e->wTc2.as<unsigned char>(); tp.img_bytes = (int)loop_tc2_image_bytes();
            tp.v1 = e->dv1; tp.v2 = e->dv2; tp.v3 = e->dv3; tp.bhn1 = e->dbhn1; tp.bhn2 = e->dbhn2; tp.bfc3 = e->dbfc3;
            tp.CS = e->bCS.as<float>();
            tp.folds = e->bFolds.as<FoldDesc>() + w0;
            tp.B = B; tp.Bc = Bc; tp.S = S; tp.seed = rq->seed;
            tp.samples = e->bSamples.as<float>() + (size_t)w0 * S;
            tp.logits_out = rq->logits ? e->bLogits.as<float>() + (size_t)w0 * S * e->C : nullptr;
            tp.forced = rq->forced ? e->bForced.as<float>() + (size_t)w0 * S : nullptr;
            tp.progress = e->dProgress;
            tp.abort_flag = e->dAbort;
            const bool want_trace2 = getenv("WRNN_TC_TRACE") != nullptr;
            if (want_trace2) {
                CU(e->bFloor.ensure(16 * 32 * sizeof(long long)));
                CU(cudaMemsetAsync(e->bFloor.p, 0, 16 * 32 * sizeof(long long), st));
                tp.trace = e->bFloor.as<long long>();
            }
            CU(launch_loop_tc2(tp, ncl, st));
            rq->loop_kernel = WRNN_LOOP_TC2;
            if (want_trace2) {
                std::vector<long long> tr(16 * 32);
                CU(cudaMemcpyAsync(tr.data(), e->bFloor.p, tr.size() * sizeof(long long), cudaMemcpyDeviceToHost, st));
                CU(cudaStreamSynchronize(st));
                if (FILE* f = fopen(getenv("WRNN_TC_TRACE"), "w")) {
                    for (int i = 0; i < 16; ++i) {
                        for (int j = 0; j < 32; ++j) fprintf(f, "%lld ", tr[i * 32 + j] ? tr[i * 32 + j] - tr[i * 32] : -1LL);
                        fprintf(f, "\n");
                    }
                    fclose(f);
                }
            }
            e->launches += 2;
        } else if (use_tc) {
            // ---- tensor-core loop: expand the conditioning per sample, then one cooperative launch ----------------
            // up to 256 folds: one set per group; more: two or three sets per group, pipelined through the same CTAs
            int nsets = std::min(kTcSets, (B + kTcGroups * 128 - 1) / (kTcGroups * 128));
            if (const char* ev = getenv("WRNN_TC_SETS")) nsets = std::max(nsets, std::min(kTcSets, atoi(ev)));
            // RAW with 512 classes: fc3 + the draw on 4 sampler CTAs per group (WRNN_TC_RAWSAMP=0: classes spread over the unit-owning CTAs)
            const bool raw_samplers = e->mode != WRNN_MODE_MOL && e->C == 512 && !(getenv("WRNN_TC_RAWSAMP") && atoi(getenv("WRNN_TC_RAWSAMP")) == 0);
            // CTA pairs (tcgen05 cta_group::2), possible whenever fc3 has its own CTAs: an even number of sets per group, set 2s+r
            // on the rank-r CTA of every pair
            // (measured: pairs win once a group pipelines three or more sets, i.e. > 512 folds; below that the longer pair MMA
            //  and the coupling of two CTAs cost ~1 us per stage; WRNN_TC_PAIR=1 / 0 forces the choice)
            const char* pair_env = getenv("WRNN_TC_PAIR");
            bool pair = (e->mode == WRNN_MODE_MOL || raw_samplers) && (pair_env ? atoi(pair_env) != 0 : nsets >= 3);
            if (pair) {     // every CTA spins on the others: all 2-CTA clusters must fit the device at once (else: single CTAs)
                int grid = kTcGroups * kTcCtas + loop_tc_sampler_ctas(e->mode, raw_samplers ? 1 : 0, 1);
                grid += std::max(0, (e->n_sms - grid) & ~1);
                if (!loop_tc_pair_fits(nsets, std::min(grid, e->n_sms & ~1))) pair = false;
            }
            if (pair) nsets = nsets <= 2 ? 2 : 4;
            const int nvg = kTcGroups * nsets, Mg = (B + nvg - 1) / nvg;
            // The expansion is HBM-write-bound (~0.5 ms per GB; 17 GB for a 60 s utterance).  Kernels on different streams
            // do not overlap on this platform (tools/probes/concurrent.cu), so it runs INSIDE the loop kernel: the loop
            // leaves 18-20 SMs free, expander CTAs on them produce the records a few 16-step chunks ahead of the loop
            // (which needs <= 0.35 TB/s of them); per-chunk counters order the two, and CS is a ring of cs_steps steps per
            // fold (<= 24 GiB whatever the fold length).  WRNN_TC_OVERLAP=0: expand the whole table first.
            const bool overlap_cs = !(getenv("WRNN_TC_OVERLAP") && atoi(getenv("WRNN_TC_OVERLAP")) == 0);
            const int nchunks = (S + kExpandSteps - 1) / kExpandSteps;
            int cs_steps = nchunks * kExpandSteps;
            if (overlap_cs) {
                const size_t per_step = (size_t)nvg * Mg * 256 * 64;
                size_t budget = (size_t)24 << 30;
                if (const char* ev = getenv("WRNN_TC_CS_BUDGET_MB")) budget = (size_t)atoll(ev) << 20;     // (tests: force the ring)
                const size_t fit = (budget / per_step) / kExpandSteps * kExpandSteps;
                cs_steps = (int)std::min<size_t>((size_t)cs_steps, std::max<size_t>(fit, 8 * kExpandSteps));
            }
            const size_t cs_bytes = (size_t)nvg * cs_steps * Mg * 256 * 64;
            if (cs_bytes > ((size_t)96 << 30)) return fail(e, WRNN_ERR_INVALID, "per-sample conditioning table would exceed 96 GiB");
            CU(e->bCS.ensure(cs_bytes));
            CU(e->bCsDone.ensure((size_t)2 * nchunks * sizeof(unsigned int)));
            CU(cudaMemsetAsync(e->bCsDone.p, 0, (size_t)2 * nchunks * sizeof(unsigned int), st));
            CU(cudaEventRecord(e->evx[0], st));
            if (!overlap_cs)
                CU(launch_expand_cond(e->bTA1.as<float4>(), e->bTA2.as<float4>(), e->bTQ1.as<float4>(), e->bTQ2.as<float4>(), e->dcoef,
                                      e->bFolds.as<FoldDesc>() + w0, B, S, Mg, e->bCS.as<float4>(), st));
            CU(cudaEventRecord(e->evx[1], st));
            expanded = true;
            const size_t xrows = (size_t)kTcGroups * kTcSets * 128;
            const size_t actb = xrows * kRnn * sizeof(__half);
            const size_t lgb = xrows * e->Cpad * sizeof(unsigned long long);
            const size_t xb = xrows * sizeof(unsigned long long);
            const size_t exch = 4 * actb + lgb + xb + 256;
            CU(e->bTcExch.ensure(exch));
            CU(cudaMemsetAsync(e->bTcExch.p, 0, exch, st));
            unsigned char* xb0 = e->bTcExch.as<unsigned char>();
            TcParams tp;
            memset(&tp, 0, sizeof(tp));
            tp.wimg = e->wTc.as<unsigned char>();
            tp.v1 = e->dv1; tp.v2 = e->dv2; tp.v3 = e->dv3; tp.bhn1 = e->dbhn1; tp.bhn2 = e->dbhn2; tp.bfc3 = e->dbfc3;
            tp.CS = e->bCS.as<float4>(); tp.Mg = Mg; tp.nsets = nsets; tp.cs_steps = cs_steps;
            tp.pair = pair ? 1 : 0;
            if (raw_samplers) {
                tp.raw_samplers = 1;
                tp.wimg_s = e->wTcS.as<unsigned char>();
            }
            if (overlap_cs) {
                tp.cs_done = e->bCsDone.as<unsigned int>(); tp.CSw = e->bCS.as<float4>();
                if (cs_steps < nchunks * kExpandSteps) tp.cs_consumed = e->bCsDone.as<unsigned int>() + nchunks;
                tp.TA1 = e->bTA1.as<float4>(); tp.TA2 = e->bTA2.as<float4>(); tp.TQ1 = e->bTQ1.as<float4>(); tp.TQ2 = e->bTQ2.as<float4>();
                tp.coef = e->dcoef;
                tp.n_expanders = std::max(0, e->n_sms - (kTcGroups * kTcCtas + loop_tc_sampler_ctas(e->mode, tp.raw_samplers, tp.pair)));
                if (const char* ev = getenv("WRNN_TC_EXPANDERS")) tp.n_expanders = std::max(1, std::min(tp.n_expanders, atoi(ev)));   // (diagnosis: fewer expander CTAs)
                if (pair && ((kTcGroups * kTcCtas + loop_tc_sampler_ctas(e->mode, tp.raw_samplers, tp.pair) + tp.n_expanders) & 1)) --tp.n_expanders;   // 2-CTA clusters
                if (tp.n_expanders <= 0) return fail(e, WRNN_ERR_INVALID, "no SM left for the conditioning expanders (WRNN_TC_OVERLAP=0 expands first)");
            }
            if (const char* ev = getenv("WRNN_TC_FLAGS")) tp.flags = atoi(ev);
            tp.folds = e->bFolds.as<FoldDesc>() + w0;
            tp.B = B; tp.S = S; tp.C = e->C; tp.Cpad = e->Cpad; tp.mode = e->mode;
            tp.seed = rq->seed;
            tp.H1 = reinterpret_cast<__half*>(xb0); tp.H2 = reinterpret_cast<__half*>(xb0 + actb);
            tp.F1 = reinterpret_cast<__half*>(xb0 + 2 * actb); tp.F2 = reinterpret_cast<__half*>(xb0 + 3 * actb);
            tp.bLG = reinterpret_cast<unsigned long long*>(xb0 + 4 * actb);
            tp.bX = reinterpret_cast<unsigned long long*>(xb0 + 4 * actb + lgb);
            tp.counters = reinterpret_cast<unsigned int*>(xb0 + 4 * actb + lgb + xb);
            tp.samples = e->bSamples.as<float>() + (size_t)w0 * S;
            tp.logits_out = rq->logits ? e->bLogits.as<float>() + (size_t)w0 * S * e->C : nullptr;
            tp.forced = rq->forced ? e->bForced.as<float>() + (size_t)w0 * S : nullptr;
            tp.progress = e->dProgress;
            tp.abort_flag = e->dAbort;
            alignas(64) unsigned char tmaps[4][128];
            __half* acts[4] = {tp.H1, tp.H2, tp.F1, tp.F2};
            const int box_rows = std::min(128, (Mg + 7) & ~7);     // only the live folds travel
            tp.tile_bytes = box_rows * 128;
            for (int i = 0; i < 4; ++i) CU(make_tmap_f16_kblocks(tmaps[i], acts[i], (uint64_t)xrows, kRnn, box_rows, kTcKbPerOp));
            const bool want_trace = getenv("WRNN_TC_TRACE") != nullptr;
            if (want_trace) {
                CU(e->bFloor.ensure((16 * 192 + 148 * 16) * sizeof(long long)));
                CU(cudaMemsetAsync(e->bFloor.p, 0, (16 * 192 + 148 * 16) * sizeof(long long), st));
                tp.trace = e->bFloor.as<long long>();
            }
            CU(launch_loop_tc(tp, tmaps, st));
            rq->loop_kernel = WRNN_LOOP_TC;
            if (want_trace) {
                std::vector<long long> tr(16 * 192);
                CU(cudaMemcpyAsync(tr.data(), e->bFloor.p, tr.size() * sizeof(long long), cudaMemcpyDeviceToHost, st));
                CU(cudaStreamSynchronize(st));
                if (FILE* f = fopen(getenv("WRNN_TC_TRACE"), "w")) {
                    for (int i = 0; i < 16; ++i) {
                        for (int j = 0; j < 192; ++j) fprintf(f, "%lld ", tr[i * 192 + j] ? tr[i * 192 + j] - tr[i * 192] : -1LL);
                        fprintf(f, "\n");
                    }
                    fclose(f);
                }
                {   // cross-CTA stamps (globaltimer ns) of one step: one line per unit-owning CTA
                    std::vector<long long> xs(148 * 16);
                    CU(cudaMemcpyAsync(xs.data(), e->bFloor.as<long long>() + 16 * 192, xs.size() * sizeof(long long), cudaMemcpyDeviceToHost, st));
                    CU(cudaStreamSynchronize(st));
                    if (FILE* f = fopen((std::string(getenv("WRNN_TC_TRACE")) + ".skew").c_str(), "w")) {
                        for (int i = 0; i < kTcGroups * kTcCtas; ++i) {
                            for (int j = 0; j < 16; ++j) fprintf(f, "%lld ", xs[i * 16 + j]);
                            fprintf(f, "\n");
                        }
                        fclose(f);
                    }
                }
                if (FILE* f = fopen((std::string(getenv("WRNN_TC_TRACE")) + ".abs").c_str(), "w")) {   // step starts (absolute)
                    for (int i = 0; i < 16; ++i) fprintf(f, "%lld\n", tr[i * 192]);
                    fclose(f);
                }
            }
            e->launches += 2;
        } else {
        const int FB = loop_f32_pick_fb(B, e->CR, e->smem_limit);
        if (FB < 1) return fail(e, WRNN_ERR_INVALID, "shared memory budget exceeded");
        CU(e->bExch.ensure(words_per_fold * B * sizeof(unsigned long long)));
        CU(cudaMemsetAsync(e->bExch.p, 0, words_per_fold * B * sizeof(unsigned long long), st));
        LoopParams p;
        memset(&p, 0, sizeof(p));
        p.Whh1 = e->dWhh1; p.Wih2a = e->dWih2a; p.Whh2 = e->dWhh2; p.Wfc1a = e->dWfc1a; p.Wfc2a = e->dWfc2a; p.Wfc3 = e->dWfc3;
        p.v1 = e->dv1; p.v2 = e->dv2; p.v3 = e->dv3; p.bhn1 = e->dbhn1; p.bhn2 = e->dbhn2; p.bfc3 = e->dbfc3;
        p.TA1 = e->bTA1.as<float4>(); p.TA2 = e->bTA2.as<float4>(); p.TQ1 = e->bTQ1.as<float4>(); p.TQ2 = e->bTQ2.as<float4>();
        p.coef = e->dcoef;
        p.folds = e->bFolds.as<FoldDesc>() + w0;
        p.B = B; p.S = S; p.C = e->C; p.Cpad = e->Cpad; p.CR = e->CR; p.FB = FB; p.mode = e->mode;
        p.seed = rq->seed;
        unsigned long long* x = e->bExch.as<unsigned long long>();
        p.bH1 = x; x += (size_t)B * kRnn; p.bH2 = x; x += (size_t)B * kRnn;
        p.bF1 = x; x += (size_t)B * kRnn; p.bF2 = x; x += (size_t)B * kRnn;
        p.bLG = x; x += (size_t)B * e->Cpad; p.bX = x;
        p.samples = e->bSamples.as<float>() + (size_t)w0 * S;
        p.logits_out = rq->logits ? e->bLogits.as<float>() + (size_t)w0 * S * e->C : nullptr;
        p.forced = rq->forced ? e->bForced.as<float>() + (size_t)w0 * S : nullptr;
        p.progress = e->dProgress;
        p.abort_flag = e->dAbort;
        CU(launch_loop_f32(p, st));
        rq->loop_kernel = WRNN_LOOP_F32;
        e->launches += 1;
        }
        CU(cudaEventRecord(e->ev[7], st));
        rq->n_launches += 1;
        // progress_callback(i, seq_len, b_size, gen_rate) -- fatchord_version.py:234-236
        if (rq->progress) {
            int last = -1;
            while (cudaEventQuery(e->ev[7]) == cudaErrorNotReady) {
                const int i = *reinterpret_cast<volatile int*>(e->hProgress);
                if (i != last) {
                    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
                    rq->progress(i, S, B, (double)(i + 1) / std::max(dt, 1e-9) * B / 1000.0, rq->progress_user);
                    last = i;
                }
                std::this_thread::sleep_for(std::chrono::microseconds(100));
            }
        }
        int aborted[2] = {0, 0};
        CU(cudaMemcpyAsync(aborted, e->dAbort, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (expanded) ms_expand += elapsed(e->evx[0], e->evx[1]);
        if (aborted[1]) return fail(e, WRNN_ERR_TIMEOUT, "front-end tensor-core GEMM pipeline timed out");
        if (aborted[0]) return fail(e, WRNN_ERR_TIMEOUT, "sample loop deadlock guard fired (an exchange word never arrived)");
        if (rq->progress) {
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
            rq->progress(S - 1, S, B, (double)S / std::max(dt, 1e-9) * B / 1000.0, rq->progress_user);
        }
    }
    CU(cudaEventRecord(e->ev[3], st));

    // ---- post chain + copies back -----------------------------------------------------------------------------
    if (want_wav) {
        rc = ensure_fades(e, rq->batched ? overlap : 0);
        if (rc) return rc;
        CU(e->bPostUtt.ensure(putts.size() * sizeof(PostUtt)));
        CU(cudaMemcpyAsync(e->bPostUtt.p, putts.data(), putts.size() * sizeof(PostUtt), cudaMemcpyHostToDevice, st));
        CU(e->bScratch.ensure((size_t)wav_total * sizeof(double)));
        double* dwav = rq->wav;
        if (!rq->wav_on_device) {
            CU(e->bWav.ensure((size_t)wav_total * sizeof(double)));
            dwav = e->bWav.as<double>();
        }
        const int ov = rq->batched ? overlap : 0;
        CU(launch_post(e->bSamples.as<float>(), e->bPostUtt.as<PostUtt>(), n_utts, max_wave_len, S, rq->batched, target, ov,
                       e->bFade.as<double>(), e->bFade.as<double>() + ov, mu_law, e->C, rq->apply_preemphasis,
                       e->bScratch.as<double>(), dwav, st));
        e->launches += 2;
        CU(cudaEventRecord(e->ev[4], st));
        if (!rq->wav_on_device)
            CU(cudaMemcpyAsync(rq->wav, dwav, (size_t)wav_total * sizeof(double), cudaMemcpyDeviceToHost, st));
    } else {
        CU(cudaEventRecord(e->ev[4], st));
    }
    if (rq->samples) CU(cudaMemcpyAsync(rq->samples, e->bSamples.p, (size_t)Btot * S * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (rq->logits) CU(cudaMemcpyAsync(rq->logits, e->bLogits.p, (size_t)Btot * S * e->C * sizeof(float), cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(e->ev[5], st));
    CU(cudaStreamSynchronize(st));
    rq->ms_h2d = elapsed(e->ev[0], e->ev[1]);
    rq->ms_cond = elapsed(e->ev[1], e->ev[2]) + ms_expand;     // front end + per-sample expansion
    rq->ms_loop = elapsed(e->ev[2], e->ev[3]) - ms_expand;     // the loop kernels only
    rq->ms_post = elapsed(e->ev[3], e->ev[4]);
    rq->ms_d2h = elapsed(e->ev[4], e->ev[5]);
    rq->n_folds = Btot;
    rq->n_steps = S;
    return WRNN_OK;
}

static int condition_impl(wrnn_engine* e, const float* mel, int32_t T, float* aux_frames, float* mels_up, bool tc) {
    if (!e || !mel || T < 1) return WRNN_ERR_INVALID;
    if (!e->finalized) return fail(e, WRNN_ERR_NOT_LOADED, "Please load Wave-RNN in memory before using it");
    CU(cudaSetDevice(e->device));
    std::vector<UttDesc> utts;
    int ta_rows = 0, tq_rows = 0;
    const float* mels[1] = {mel};
    int rc = upload_utts(e, mels, &T, 1, 0, utts, ta_rows, tq_rows);
    if (rc) return rc;
    CU(cudaMemsetAsync(e->dAbort, 0, 2 * sizeof(int), e->stream));
    rc = tc ? run_conditioning_tc(e, 1, ta_rows) : run_conditioning(e, 1, ta_rows, tq_rows);
    if (rc) return rc;
    int st2[2] = {0, 0};
    CU(cudaMemcpyAsync(st2, e->dAbort, 2 * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    if (aux_frames) CU(cudaMemcpyAsync(aux_frames, e->bAux.p, (size_t)T * 128 * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    if (st2[1]) return fail(e, WRNN_ERR_TIMEOUT, "front-end tensor-core GEMM pipeline timed out");
    if (mels_up) {
        // debug view of the interpolation table: mels_up[n][c] = sum_d coef[n%200][d] * melpad[c][n/200 + d]
        for (int n = 0; n < T * kHop; ++n) {
            const int q0 = n / kHop, ph = n % kHop;
            for (int c = 0; c < kFeat; ++c) {
                float a = 0.f;
                for (int d = 0; d < kTaps; ++d) {
                    const int t = q0 + d - kPad;
                    if (t >= 0 && t < T) a = fmaf(e->hcoef[ph * kTaps + d], mel[(size_t)c * T + t], a);
                }
                mels_up[(size_t)n * kFeat + c] = a;
            }
        }
    }
    return WRNN_OK;
}

int wrnn_condition(wrnn_engine* e, const float* mel, int32_t T, float* aux_frames, float* mels_up) {
    return condition_impl(e, mel, T, aux_frames, mels_up, false);
}
int wrnn_condition_tc(wrnn_engine* e, const float* mel, int32_t T, float* aux_frames) {
    return condition_impl(e, mel, T, aux_frames, nullptr, true);
}

int wrnn_postprocess(wrnn_engine* e, const float* samples, int64_t num_folds, int64_t S, int32_t batched, int32_t overlap,
                     int32_t T, int32_t mu_law, int32_t apply_preemphasis, double* wav) {
    if (!e || !samples || !wav || num_folds < 1 || S < 1) return WRNN_ERR_INVALID;
    if (T <= 20) return fail(e, WRNN_ERR_TOO_SHORT, "mel has <= 20 frames (fatchord_version.py:253-255)");
    if (batched && overlap <= 0) return fail(e, WRNN_ERR_INVALID, "overlap == 0 (fatchord_version.py:394)");
    if (batched && S < 2 * (int64_t)overlap) return fail(e, WRNN_ERR_INVALID, "folds shorter than 2 * overlap (fatchord_version.py:376)");
    CU(cudaSetDevice(e->device));
    const int target = batched ? (int)S - 2 * overlap : 0;
    PostUtt u;
    u.samp_off = 0; u.wav_off = 0; u.F = (int)num_folds; u.wave_len = (T - 1) * kHop;
    const int64_t total_len = batched ? num_folds * (target + overlap) + overlap : S;
    if (u.wave_len > total_len) return fail(e, WRNN_ERR_INVALID, "samples shorter than (T-1)*hop");
    int rc = ensure_fades(e, batched ? overlap : 0);
    if (rc) return rc;
    cudaStream_t st = e->stream;
    CU(e->bSamples.ensure((size_t)num_folds * S * sizeof(float)));
    CU(cudaMemcpyAsync(e->bSamples.p, samples, (size_t)num_folds * S * sizeof(float), cudaMemcpyHostToDevice, st));
    CU(e->bPostUtt.ensure(sizeof(PostUtt)));
    CU(cudaMemcpyAsync(e->bPostUtt.p, &u, sizeof(PostUtt), cudaMemcpyHostToDevice, st));
    CU(e->bScratch.ensure((size_t)u.wave_len * sizeof(double)));
    CU(e->bWav.ensure((size_t)u.wave_len * sizeof(double)));
    const int ov = batched ? overlap : 0;
    CU(launch_post(e->bSamples.as<float>(), e->bPostUtt.as<PostUtt>(), 1, u.wave_len, (int)S, batched, target, ov,
                   e->bFade.as<double>(), e->bFade.as<double>() + ov, (e->mode == WRNN_MODE_RAW) ? mu_law : 0, e->C,
                   apply_preemphasis, e->bScratch.as<double>(), e->bWav.as<double>(), st));
    e->launches += 2;
    CU(cudaMemcpyAsync(wav, e->bWav.p, (size_t)u.wave_len * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return WRNN_OK;
}

int wrnn_xfade_unfold(wrnn_engine* e, const double* y, int64_t num_folds, int64_t S, int32_t overlap, double* out) {
    if (!e || !y || !out || num_folds < 1 || S < 1) return WRNN_ERR_INVALID;
    if (overlap <= 0 || 2 * (int64_t)overlap > S)
        return fail(e, WRNN_ERR_INVALID, "overlap must be in [1, S/2] (overlap == 0 raises ValueError at fatchord_version.py:394)");
    CU(cudaSetDevice(e->device));
    int rc = ensure_fades(e, overlap);
    if (rc) return rc;
    const long long total_len = num_folds * (S - overlap) + overlap;
    cudaStream_t st = e->stream;
    CU(e->bScratch.ensure((size_t)num_folds * S * sizeof(double)));
    CU(e->bWav.ensure((size_t)total_len * sizeof(double)));
    CU(cudaMemcpyAsync(e->bScratch.p, y, (size_t)num_folds * S * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(launch_xfade_unfold_f64(e->bScratch.as<double>(), (int)num_folds, (int)S, overlap, e->bFade.as<double>(),
                               e->bFade.as<double>() + overlap, total_len, e->bWav.as<double>(), st));
    e->launches += 1;
    CU(cudaMemcpyAsync(out, e->bWav.p, (size_t)total_len * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return WRNN_OK;
}

int wrnn_debug_tc_gemm(wrnn_engine* e, const uint16_t* A, const uint16_t* W, int32_t N, float* C) {
    if (!e || !A || !W || !C) return WRNN_ERR_INVALID;
    CU(cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    CU(e->bScratch.ensure((size_t)128 * 512 * 2 + (size_t)N * 512 * 2 + (size_t)128 * N * 4 + 1024));
    uint8_t* b = e->bScratch.as<uint8_t>();
    uint8_t *dA = b, *dW = b + 128 * 512 * 2, *dC = dW + (size_t)N * 512 * 2;
    CU(cudaMemcpyAsync(dA, A, 128 * 512 * 2, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(dW, W, (size_t)N * 512 * 2, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(dC, 0, (size_t)128 * N * 4, st));
    CU(cudaMemsetAsync(e->dAbort, 0, sizeof(int), st));
    CU(run_tc_gemm_test(dA, dW, N, reinterpret_cast<float*>(dC), e->dAbort, st));
    e->launches += 1;
    int status = 0;
    CU(cudaMemcpyAsync(C, dC, (size_t)128 * N * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&status, e->dAbort, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (status) return fail(e, WRNN_ERR_TIMEOUT, "tc gemm self-test: pipeline stage " + std::to_string(status) + " timed out");
    return WRNN_OK;
}

int wrnn_debug_tc_gemm2(wrnn_engine* e, const uint16_t* A, const uint16_t* W, int32_t N, float* C) {
    if (!e || !A || !W || !C) return WRNN_ERR_INVALID;
    CU(cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    CU(e->bScratch.ensure((size_t)256 * 512 * 2 + (size_t)N * 512 * 2 + (size_t)256 * N * 4 + 1024));
    uint8_t* b = e->bScratch.as<uint8_t>();
    uint8_t *dA = b, *dW = b + 256 * 512 * 2, *dC = dW + (size_t)N * 512 * 2;
    CU(cudaMemcpyAsync(dA, A, 256 * 512 * 2, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(dW, W, (size_t)N * 512 * 2, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(dC, 0, (size_t)256 * N * 4, st));
    CU(cudaMemsetAsync(e->dAbort, 0, sizeof(int), st));
    CU(run_tc_gemm2_test(dA, dW, N, reinterpret_cast<float*>(dC), e->dAbort, st));
    e->launches += 1;
    int status = 0;
    CU(cudaMemcpyAsync(C, dC, (size_t)256 * N * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&status, e->dAbort, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (status) return fail(e, WRNN_ERR_TIMEOUT, "tc pair gemm self-test: pipeline stage " + std::to_string(status) + " timed out");
    return WRNN_OK;
}

int wrnn_debug_umma_rate(wrnn_engine* e, int32_t N, int32_t iters, int32_t mode, int64_t* cycles_issue, int64_t* cycles_total) {
    if (!e || !cycles_issue || !cycles_total) return WRNN_ERR_INVALID;
    CU(cudaSetDevice(e->device));
    CU(e->bFloor.ensure(4 * kRnn * sizeof(unsigned long long) + 64));
    CU(run_umma_rate(N, iters, mode, e->bFloor.as<long long>(), e->stream));
    long long h[2] = {0, 0};
    CU(cudaMemcpyAsync(h, e->bFloor.p, sizeof(h), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    e->launches += 1;
    *cycles_issue = h[0];
    *cycles_total = h[1];
    return WRNN_OK;
}

int wrnn_cluster_floor(wrnn_engine* e, int32_t cluster_size, int32_t rounds, float* us) {
    if (!e || rounds < 1 || !us) return WRNN_ERR_INVALID;
    CU(cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    CU(e->bFloor.ensure(4 * kRnn * sizeof(unsigned long long) + 64));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CU(cudaEventRecord(e->ev[0], st));
        CU(launch_floor_cluster(cluster_size, rounds, e->bFloor.as<float>(), st));
        CU(cudaEventRecord(e->ev[1], st));
        CU(cudaStreamSynchronize(st));
        e->launches += 1;
        best = std::min(best, elapsed(e->ev[0], e->ev[1]) * 1000.f / rounds);
    }
    *us = best;
    return WRNN_OK;
}

int wrnn_barrier_floor(wrnn_engine* e, int32_t rounds, float* ll_us, float* counter_us) {
    if (!e || rounds < 1) return WRNN_ERR_INVALID;
    CU(cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    CU(e->bFloor.ensure(4 * kRnn * sizeof(unsigned long long) + 64));
    unsigned long long* buf = e->bFloor.as<unsigned long long>();
    for (int variant = 0; variant < 2; ++variant) {
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            CU(cudaMemsetAsync(e->bFloor.p, 0, e->bFloor.cap, st));
            CU(cudaMemsetAsync(e->dAbort, 0, sizeof(int), st));
            CU(cudaEventRecord(e->ev[0], st));
            if (variant == 0) CU(launch_floor_ll(buf, rounds, e->dAbort, st));
            else CU(launch_floor_counter(reinterpret_cast<unsigned int*>(buf + 3 * kRnn), reinterpret_cast<float*>(buf), rounds, e->dAbort, st));
            CU(cudaEventRecord(e->ev[1], st));
            int aborted = 0;
            CU(cudaMemcpyAsync(&aborted, e->dAbort, sizeof(int), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            e->launches += 1;
            if (aborted) return fail(e, WRNN_ERR_TIMEOUT, "exchange-floor kernel timed out");
            best = std::min(best, elapsed(e->ev[0], e->ev[1]) * 1000.f / rounds);
        }
        if (variant == 0 && ll_us) *ll_us = best;
        if (variant == 1 && counter_us) *counter_us = best;
    }
    return WRNN_OK;
}

}  // extern "C"
