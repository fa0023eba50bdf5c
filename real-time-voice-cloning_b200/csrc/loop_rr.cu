// loop_rr.cu -- the autoregressive sample loop of the `runtimeracer-wavernn` topology (reference: WaveRNN.generate body,
// vocoder/models/runtimeracer_version.py:248-288; layers :119-132: four GRU-256 cells and five FC layers) as ONE persistent
// cooperative kernel, fp32 -- SURVEY.md section 8(f) row 1.  Same construction as loop_f32.cu (the fatchord parity loop):
//  * weight-stationary: 128 CTAs (one per SM) each own 2 of the 256 hidden units of every layer; their rows stay in shared
//    memory (fp32, ~50 KB) for the whole sequence;
//  * everything linear in the conditioning is folded into per-frame tables by the front end (engine.cu: finalize_rr): the I
//    layer, every W_ih applied to it through the residual chain v_k = v_0 + h_1 + .. + h_k, the aux columns, the biases; the two
//    pairs of FC layers without an activation between them are ONE matrix each (M12 = fc2 fc1[:, :256], M34 = fc4 fc3[:, :256]),
//    so a step is 4 GRU stations + 3 FC stations:
//      x -> GRU1 -> h1 -> [W_hh1 | W_ih2] h1 -> GRU2 -> h2 -> W_ih3 (h1+h2), W_hh2 h2 -> GRU3 -> h3 -> W_ih4 (h1+h2+h3), W_hh3 h3
//        -> GRU4 -> h4 -> M12 (h1+..+h4), W_hh4 h4 -> relu -> y2 -> M34 y2 -> relu -> y4 -> fc5 y4 -> logits -> sample x';
//  * activations travel between SMs as {value, step-tag} 8-byte words through L2 (common.cuh): consumers spin on the data
//    itself; no grid barrier, fence or atomic on the path.  Every h_k is gathered once per CTA and serves both its recurrent
//    product (for the next step) and the running sum the next layer reads;
//  * sampling (softmax + inverse CDF, or mixture of logistics) fused, Philox noise as in loop_f32.cu (sampling.cuh);
//  * every spin has a deadline; a miss raises the abort flag and all CTAs leave.
// One launch serves <= kRrMaxFolds (64) folds (the engine runs longer batches in waves).
#include "engine_internal.h"
#include "sampling.cuh"
#include "chain_f32.cuh"

namespace wrnn {

namespace {

constexpr int H = kRrH;           // 256
using chain::NT;
using chain::NW;
constexpr int U = kRrH / kRrCtas; // hidden units per CTA (2)
constexpr int G = 3 * U;          // gate rows per CTA and GRU matrix (6)
__device__ long long g_rr_deadline = 1500000000LL;

struct Smem {
    float* W1;      // [2G][H]  W_hh1 | W_ih2                 on h1
    float* W2s;     // [G][H]   W_ih3[:, :H]                  on h1 + h2
    float* W2h;     // [G][H]   W_hh2                         on h2
    float* W3s;     // [G][H]   W_ih4                         on h1 + h2 + h3
    float* W3h;     // [G][H]   W_hh3                         on h3
    float* W4s;     // [U][H]   M12                           on h1 + .. + h4
    float* W4h;     // [G][H]   W_hh4                         on h4
    float* W5;      // [U][H]   M34                           on y2
    float* W6;      // [CR][H]  fc5                           on y4
    float* act;     // [B][H]   the vector just gathered
    float* sum;     // [B][H]   running sum h1 + .. + hk
    float* tmp;     // [B][2G]  GEMV results
    float* h;       // [4][B][U]
    float* gh;      // [4][B][G]   W_hh h of the previous step
    float4* c;      // [B][U][4]   conditioning of this step (see RrParams::TA)
    float* x;       // [B]
    float* u;       // [13][U]  rank-1 coefficients of the previous sample: GRU1..4 r,z,n | M12
    float* bhn;     // [4][U]
    float* b5;      // [CR]
    float* coef;    // [200][kTaps]
};

__device__ __forceinline__ Smem carve(float* base, int B, int CR) {
    Smem s;
    float* p = base;
    s.W1 = p;  p += 2 * G * H;
    s.W2s = p; p += G * H;
    s.W2h = p; p += G * H;
    s.W3s = p; p += G * H;
    s.W3h = p; p += G * H;
    s.W4s = p; p += U * H;
    s.W4h = p; p += G * H;
    s.W5 = p;  p += U * H;
    s.W6 = p;  p += CR * H;
    s.act = p; p += B * H;
    s.sum = p; p += B * H;
    s.c = reinterpret_cast<float4*>(p); p += B * U * 16;
    s.tmp = p; p += B * 2 * G;
    s.h = p;   p += 4 * B * U;
    s.gh = p;  p += 4 * B * G;
    s.x = p;   p += (B + 3) & ~3;
    s.u = p;   p += 13 * U + 2;
    s.bhn = p; p += 4 * U;
    s.b5 = p;  p += (CR + 3) & ~3;
    s.coef = p; p += kHop * kTaps;
    return s;
}

}  // namespace

__global__ void __launch_bounds__(NT, 1) wrnn_loop_rr_kernel(RrParams p) {
    extern __shared__ __align__(16) float smem_f[];
    const int cta = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int B = p.B, C = p.C, CR = p.CR;
    const int j0 = cta * U;
    Smem s = carve(smem_f, B, CR);
    constexpr int LDT = 2 * G;

    // ---- one-time: my weight rows into shared memory (gate-major: row = gate * U + unit) --------------------------------
    auto load_gru = [&](float* dst, const float* W) {         // rows (g, u) of a [3H][H] matrix
        for (int i = tid; i < G * (H / 4); i += NT) {
            const int row = i / (H / 4), k4 = i % (H / 4), g = row / U, u = row % U;
            reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(W)[(size_t)(g * H + j0 + u) * (H / 4) + k4];
        }
    };
    auto load_fc = [&](float* dst, const float* W) {          // rows j0 .. j0 + U of a [H][H] matrix
        for (int i = tid; i < U * (H / 4); i += NT) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(W)[(size_t)j0 * (H / 4) + i];
    };
    load_gru(s.W1, p.Whh[0]); load_gru(s.W1 + G * H, p.Wih[0]);
    load_gru(s.W2s, p.Wih[1]); load_gru(s.W2h, p.Whh[1]);
    load_gru(s.W3s, p.Wih[2]); load_gru(s.W3h, p.Whh[2]);
    load_fc(s.W4s, p.M12);     load_gru(s.W4h, p.Whh[3]);
    load_fc(s.W5, p.M34);
    for (int i = tid; i < CR * (H / 4); i += NT) {
        const int row = i / (H / 4), cls = cta * CR + row;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cls < C) v = reinterpret_cast<const float4*>(p.Wfc5)[(size_t)cls * (H / 4) + i % (H / 4)];
        reinterpret_cast<float4*>(s.W6)[i] = v;
    }
    if (tid < 13 * U) s.u[tid] = p.u[(tid / U) * H + j0 + tid % U];
    if (tid < 4 * U) s.bhn[tid] = p.bhn[(tid / U) * H + j0 + tid % U];
    if (tid < CR) s.b5[tid] = (cta * CR + tid < C) ? p.bfc5[cta * CR + tid] : 0.f;
    for (int i = tid; i < kHop * kTaps; i += NT) s.coef[i] = p.coef[i];
    for (int i = tid; i < 4 * B * U; i += NT) s.h[i] = 0.f;
    for (int i = tid; i < 4 * B * G; i += NT) s.gh[i] = 0.f;     // W_hh * 0
    for (int i = tid; i < B; i += NT) s.x[i] = 0.f;              // x_0 = 0, runtimeracer_version.py:236
    __syncthreads();

    const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32));
    // one GRU cell update for my units of layer k (0..3): gi = on-path product (tmp columns gi0.., or none) + rank-1 + table
    auto gru_station = [&](int k, bool have_gi, int gi0, unsigned long long* out, uint32_t tag) {
        for (int e = tid; e < B * U; e += NT) {
            const int b = e / U, uu = e % U;
            const float x = s.x[b];
            const float4 cc = s.c[e * 4 + k];
            const float* gh = s.gh + ((size_t)k * B + b) * G;
            const float* gi = s.tmp + b * LDT + gi0;
            const float* uk = s.u + 3 * k * U;
            const float ir = (have_gi ? gi[0 * U + uu] : 0.f) + fmaf(uk[0 * U + uu], x, cc.x);
            const float iz = (have_gi ? gi[1 * U + uu] : 0.f) + fmaf(uk[1 * U + uu], x, cc.y);
            const float in = (have_gi ? gi[2 * U + uu] : 0.f) + fmaf(uk[2 * U + uu], x, cc.z);
            const float r = sigmoid_acc(ir + gh[0 * U + uu]);
            const float z = sigmoid_acc(iz + gh[1 * U + uu]);
            const float nn = tanhf(in + r * (gh[2 * U + uu] + s.bhn[k * U + uu]));
            float* hp = s.h + ((size_t)k * B + b) * U + uu;
            const float hn = (1.0f - z) * nn + z * *hp;
            *hp = hn;
            ll_store(out + (size_t)b * H + j0 + uu, hn, tag);
        }
    };
    auto save_gh = [&](int k, int c0) {       // tmp columns [c0, c0 + G) -> gh of layer k (used by the NEXT step)
        for (int e = tid; e < B * G; e += NT) s.gh[((size_t)k * B + e / G) * G + e % G] = s.tmp[(e / G) * LDT + c0 + e % G];
    };
#define RR_FAIL() do { if (tid == 0) atomicExch(p.abort_flag, 1); return; } while (0)

    for (int t = 0; t < p.S; ++t) {
        const uint32_t tag = (uint32_t)t + 1u;

        // ---- conditioning of this step for my units (independent of the exchange: issued before the wait) ------------------
        for (int e = tid; e < B * U * 4; e += NT) {
            const int k = e & 3, bu = e >> 2, b = bu / U, j = j0 + bu % U;
            const FoldDesc fd = p.folds[b];
            const int n = fd.n0 + t;
            const bool valid = n < fd.N;                 // positions past the utterance are fold tail padding
            const int q0 = valid ? n / kHop : 0;
            float4 a = __ldg(p.TA + ((size_t)(fd.ta_row0 + (valid ? q0 : fd.T)) * H + j) * 4 + k);
            if (valid) {
                const float* cf = s.coef + (n - q0 * kHop) * kTaps;
#pragma unroll
                for (int d = 0; d < kTaps; ++d) {
                    const float cw = cf[d];
                    if (cw != 0.f) {
                        const float4 q = __ldg(p.TQ + ((size_t)(fd.tq_row0 + q0 + d) * H + j) * 4 + k);
                        a.x = fmaf(cw, q.x, a.x); a.y = fmaf(cw, q.y, a.y); a.z = fmaf(cw, q.z, a.z); a.w = fmaf(cw, q.w, a.w);
                    }
                }
            }
            s.c[e] = a;
        }

        // ---- 1: wait x_{t-1}; GRU1 for my units; publish h1 ------------------------------------------------------------------
        {
            int failed = 0;
            if (t > 0) {
                for (int b = tid; b < B; b += NT) {
                    float v;
                    if (!wait_word(p.bX + b, (uint32_t)t, v, p.abort_flag, g_rr_deadline)) failed = 1;
                    s.x[b] = v;
                }
            }
            if (__syncthreads_or(failed)) RR_FAIL();
            gru_station(0, false, 0, p.bH[0], tag);
        }
        // ---- 2: h1 -> [W_hh1 h1 (next step) | W_ih2 h1]; GRU2; publish h2 ---------------------------------------------------
        if (chain::gather<H>(p.bH[0], B, s.act, s.sum, 1, tag, p.abort_flag, g_rr_deadline)) RR_FAIL();
        chain::dots<H>(s.W1, 2 * G, s.act, B, s.tmp, LDT, 0);
        __syncthreads();
        gru_station(1, true, G, p.bH[1], tag);
        save_gh(0, 0);
        __syncthreads();
        // ---- 3: h2 -> W_ih3 (h1 + h2), W_hh2 h2; GRU3; publish h3 -----------------------------------------------------------
        if (chain::gather<H>(p.bH[1], B, s.act, s.sum, 2, tag, p.abort_flag, g_rr_deadline)) RR_FAIL();
        chain::dots<H>(s.W2h, G, s.act, B, s.tmp, LDT, 0);
        chain::dots<H>(s.W2s, G, s.sum, B, s.tmp, LDT, G);
        __syncthreads();
        gru_station(2, true, G, p.bH[2], tag);
        save_gh(1, 0);
        __syncthreads();
        // ---- 4: h3 -> W_ih4 (h1 + h2 + h3), W_hh3 h3; GRU4; publish h4 ------------------------------------------------------
        if (chain::gather<H>(p.bH[2], B, s.act, s.sum, 2, tag, p.abort_flag, g_rr_deadline)) RR_FAIL();
        chain::dots<H>(s.W3h, G, s.act, B, s.tmp, LDT, 0);
        chain::dots<H>(s.W3s, G, s.sum, B, s.tmp, LDT, G);
        __syncthreads();
        gru_station(3, true, G, p.bH[3], tag);
        save_gh(2, 0);
        __syncthreads();
        // ---- 5: h4 -> y2 = relu(M12 (h1 + .. + h4) + u5 x + c5), W_hh4 h4; publish y2 ----------------------------------------
        if (chain::gather<H>(p.bH[3], B, s.act, s.sum, 2, tag, p.abort_flag, g_rr_deadline)) RR_FAIL();
        chain::dots<H>(s.W4h, G, s.act, B, s.tmp, LDT, 0);
        chain::dots<H>(s.W4s, U, s.sum, B, s.tmp, LDT, G);
        __syncthreads();
        for (int e = tid; e < B * U; e += NT) {
            const int b = e / U, uu = e % U;
            const float v = s.tmp[b * LDT + G + uu] + fmaf(s.u[12 * U + uu], s.x[b], s.c[e * 4 + 0].w);
            ll_store(p.bY2 + (size_t)b * H + j0 + uu, fmaxf(v, 0.f), tag);
        }
        save_gh(3, 0);
        __syncthreads();
        // ---- 6: y4 = relu(M34 y2 + c6); publish ---------------------------------------------------------------------------------
        if (chain::gather<H>(p.bY2, B, s.act, s.sum, 0, tag, p.abort_flag, g_rr_deadline)) RR_FAIL();
        chain::dots<H>(s.W5, U, s.act, B, s.tmp, LDT, 0);
        __syncthreads();
        for (int e = tid; e < B * U; e += NT) {
            const int b = e / U, uu = e % U;
            ll_store(p.bY4 + (size_t)b * H + j0 + uu, fmaxf(s.tmp[b * LDT + uu] + s.c[e * 4 + 1].w, 0.f), tag);
        }
        __syncthreads();
        // ---- 7: my classes of logits = fc5 y4 + b; publish ------------------------------------------------------------------
        if (chain::gather<H>(p.bY4, B, s.act, s.sum, 0, tag, p.abort_flag, g_rr_deadline)) RR_FAIL();
        if (cta * CR < C) {
            chain::dots<H>(s.W6, CR, s.act, B, s.tmp, LDT, 0);
            __syncthreads();
            for (int e = tid; e < B * CR; e += NT) {
                const int b = e / CR, r = e % CR, cls = cta * CR + r;
                if (cls < C) {
                    const float v = s.tmp[b * LDT + r] + s.b5[r];
                    ll_store(p.bLG + (size_t)b * p.Cpad + cls, v, tag);
                    if (p.logits_out) p.logits_out[((size_t)b * p.S + t) * C + cls] = v;
                }
            }
        }
        __syncthreads();
        // ---- 8: sample the folds assigned to this CTA; publish x_t ----------------------------------------------------------
        {
            int failed = 0;
            for (int b = cta + gridDim.x * warp; b < B; b += gridDim.x * NW) {
                const FoldDesc fd = p.folds[b];
                const unsigned long long* row = p.bLG + (size_t)b * p.Cpad;
                float xs;
                if (p.mode == 1) {
                    bool ok;
                    xs = sample_mol_warp(row, tag, key, (uint32_t)t, fd, p.abort_flag, g_rr_deadline, ok);
                    if (!ok) { failed = 1; break; }
                } else {
                    uint4 r = philox4x32_10(make_uint4((uint32_t)t, (uint32_t)fd.fold, (uint32_t)fd.utt, 0u), key);
                    const float uu = u01(r.x);
                    int k;
                    if (C == 256) k = sample_raw_warp<8>(row, tag, uu, p.abort_flag, g_rr_deadline);
                    else if (C == 512) k = sample_raw_warp<16>(row, tag, uu, p.abort_flag, g_rr_deadline);
                    else k = sample_raw_warp<32>(row, tag, uu, p.abort_flag, g_rr_deadline);
                    if (k < 0) { failed = 1; break; }
                    xs = 2.0f * (float)k / ((float)C - 1.0f) - 1.0f;     // runtimeracer_version.py:284 (fp32)
                }
                if (lane == 0) {
                    p.samples[(size_t)b * p.S + t] = xs;
                    const float fed = p.forced ? p.forced[(size_t)b * p.S + t] : xs;
                    ll_store(p.bX + b, fed, tag);
                }
            }
            if (__syncthreads_or(failed)) RR_FAIL();
        }
        if (cta == 0 && tid == 0 && (t % 100) == 0 && p.progress) {
            *reinterpret_cast<volatile int*>(p.progress) = t;
            __threadfence_system();
        }
    }
#undef RR_FAIL
}

size_t loop_rr_smem_bytes(int B, int CR) {
    size_t f = (size_t)(2 * G + 5 * G + 2 * U) * H + (size_t)CR * H + (size_t)2 * B * H + (size_t)B * U * 16 + (size_t)B * 2 * G +
               (size_t)4 * B * U + (size_t)4 * B * G + ((B + 3) & ~3) + 13 * U + 2 + 4 * U + ((CR + 3) & ~3) + kHop * kTaps;
    return f * sizeof(float);
}

cudaError_t set_rr_deadline(long long cycles) { return cudaMemcpyToSymbol(g_rr_deadline, &cycles, sizeof(cycles)); }

cudaError_t launch_loop_rr(const RrParams& p, cudaStream_t stream) {
    const size_t smem = loop_rr_smem_bytes(p.B, p.CR);
    cudaError_t err = cudaFuncSetAttribute(wrnn_loop_rr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    RrParams pp = p;
    void* args[] = {&pp};
    // cooperative launch: guarantees the 128 CTAs are co-resident (they wait on one another)
    return cudaLaunchCooperativeKernel((const void*)wrnn_loop_rr_kernel, dim3(kRrCtas), dim3(NT), args, smem, stream);
}

}  // namespace wrnn
