// loop_gn.cu -- the autoregressive sample loop of the `geneing-wavernn` topology (reference: WaveRNN.generate body,
// vocoder/models/geneing_version.py:199-232; layers :107-113: I, one GRU-256, fc1 (288 -> 128, ReLU), fc3 (128 -> classes)) as
// ONE persistent cooperative kernel, fp32 -- SURVEY.md section 8(f) row 3.  Same construction as loop_rr.cu / loop_f32.cu:
//  * weight-stationary: 128 CTAs (one per SM) own 2 GRU units, 1 fc1 unit and C/128 classes each; rows resident in shared memory;
//  * I, W_ih1 and every mel/aux column are folded into per-frame tables (engine_gn.inc: finalize_gn); the residual
//    v + h1 is expanded algebraically, so a step is three stations:
//      x -> GRU1 -> h1 -> [W_hh1 h1 (next step) | fc1[:, :256] h1] -> relu -> f -> fc3 f -> logits -> sample x';
//  * {value, step-tag} exchange words through L2, fused sampling (mode 'BITS' = softmax + inverse CDF on one Philox uniform per
//    (step, fold), the same rule as the fatchord RAW mode; 'MOL' as fatchord), deadline on every spin.
#include "engine_internal.h"
#include "sampling.cuh"
#include "chain_f32.cuh"

namespace wrnn {

namespace {

constexpr int H = kGnH;            // 256
constexpr int F = kGnFc;           // 128
using chain::NT;
using chain::NW;
constexpr int U = kGnH / kGnCtas;  // GRU units per CTA (2)
constexpr int G = 3 * U;
static_assert(kGnFc == kGnCtas, "one fc1 unit per CTA");
__device__ long long g_gn_deadline = 1500000000LL;

}  // namespace

__global__ void __launch_bounds__(NT, 1) wrnn_loop_gn_kernel(GnLoopParams p) {
    extern __shared__ __align__(16) float smem_f[];
    const int cta = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int B = p.B, C = p.C, CR = p.CR;
    const int j0 = cta * U;
    constexpr int LDT = G + 2;
    float* sp = smem_f;
    float* W1 = sp;  sp += (G + 2) * H;          // W_hh1 rows (6) | fc1[:, :256] row (1) | one zero row (the blocked GEMV takes row pairs)
    float* W3 = sp;  sp += CR * F;               // fc3 rows
    float* act = sp; sp += B * H;
    float4* cc = reinterpret_cast<float4*>(sp); sp += B * U * 4;      // {c1 r, z, n, c2} per (fold, GRU unit); c2 of unit 0 only is used
    float* tmp = sp; sp += B * LDT;
    float* h = sp;   sp += B * U;
    float* gh = sp;  sp += B * G;
    float* xs_ = sp; sp += (B + 3) & ~3;
    float* su = sp;  sp += 4 * U;                // u1 r,z,n per unit (6) | u2 (1) | pad
    float* sbhn = sp; sp += 4;
    float* sb3 = sp; sp += (CR + 3) & ~3;
    float* coef = sp; sp += kHop * kTaps;

    for (int i = tid; i < G * (H / 4); i += NT) {
        const int row = i / (H / 4), k4 = i % (H / 4), g = row / U, u = row % U;
        reinterpret_cast<float4*>(W1)[i] = reinterpret_cast<const float4*>(p.Whh)[(size_t)(g * H + j0 + u) * (H / 4) + k4];
    }
    for (int i = tid; i < H / 4; i += NT) {
        reinterpret_cast<float4*>(W1 + G * H)[i] = reinterpret_cast<const float4*>(p.Wfc1a)[(size_t)cta * (H / 4) + i];
        reinterpret_cast<float4*>(W1 + (G + 1) * H)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int i = tid; i < CR * (F / 4); i += NT) {
        const int row = i / (F / 4), cls = cta * CR + row;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cls < C) v = reinterpret_cast<const float4*>(p.Wfc3)[(size_t)cls * (F / 4) + i % (F / 4)];
        reinterpret_cast<float4*>(W3)[i] = v;
    }
    if (tid < 3 * U) su[tid] = p.u1[(tid / U) * H + j0 + tid % U];
    if (tid == 3 * U) su[tid] = p.u2[cta];
    if (tid < U) sbhn[tid] = p.bhn[j0 + tid];
    if (tid < CR) sb3[tid] = (cta * CR + tid < C) ? p.bfc3[cta * CR + tid] : 0.f;
    for (int i = tid; i < kHop * kTaps; i += NT) coef[i] = p.coef[i];
    for (int i = tid; i < B * U; i += NT) h[i] = 0.f;
    for (int i = tid; i < B * G; i += NT) gh[i] = 0.f;
    for (int i = tid; i < B; i += NT) xs_[i] = 0.f;              // x_0 = 0, geneing_version.py:188
    __syncthreads();

    const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32));
#define GN_FAIL() do { if (tid == 0) atomicExch(p.abort_flag, 1); return; } while (0)

    for (int t = 0; t < p.S; ++t) {
        const uint32_t tag = (uint32_t)t + 1u;
        // ---- conditioning of this step: TA[frame] + sum_d coef[phase][d] TQ[frame + d] for my GRU units (and my fc1 unit in .w) ----
        for (int e = tid; e < B * U; e += NT) {
            const int b = e / U, uu = e % U;
            const FoldDesc fd = p.folds[b];
            const int n = fd.n0 + t;
            const bool valid = n < fd.N;
            const int q0 = valid ? n / kHop : 0;
            float4 a = __ldg(p.TA + (size_t)(fd.ta_row0 + (valid ? q0 : fd.T)) * H + j0 + uu);
            if (valid) {
                const float* cf = coef + (n - q0 * kHop) * kTaps;
#pragma unroll
                for (int d = 0; d < kTaps; ++d) {
                    const float cw = cf[d];
                    if (cw != 0.f) {
                        const float4 q = __ldg(p.TQ + (size_t)(fd.tq_row0 + q0 + d) * H + j0 + uu);
                        a.x = fmaf(cw, q.x, a.x); a.y = fmaf(cw, q.y, a.y); a.z = fmaf(cw, q.z, a.z); a.w = fmaf(cw, q.w, a.w);
                    }
                }
            }
            cc[e] = a;
        }
        // ---- 1: wait x_{t-1}; GRU1 for my units; publish h1 ------------------------------------------------------------------
        {
            int failed = 0;
            if (t > 0) {
                for (int b = tid; b < B; b += NT) {
                    float v;
                    if (!wait_word(p.bX + b, (uint32_t)t, v, p.abort_flag, g_gn_deadline)) failed = 1;
                    xs_[b] = v;
                }
            }
            if (__syncthreads_or(failed)) GN_FAIL();
            for (int e = tid; e < B * U; e += NT) {
                const int b = e / U, uu = e % U;
                const float x = xs_[b];
                const float4 c = cc[e];
                const float* g_ = gh + b * G;
                const float r = sigmoid_acc(fmaf(su[0 * U + uu], x, c.x) + g_[0 * U + uu]);
                const float z = sigmoid_acc(fmaf(su[1 * U + uu], x, c.y) + g_[1 * U + uu]);
                const float nn = tanhf(fmaf(su[2 * U + uu], x, c.z) + r * (g_[2 * U + uu] + sbhn[uu]));
                const float hn = (1.0f - z) * nn + z * h[e];
                h[e] = hn;
                ll_store(p.bH + (size_t)b * H + j0 + uu, hn, tag);
            }
        }
        // ---- 2: h1 -> W_hh1 h1 (next step), f = relu(fc1[:, :256] h1 + u2 x + c2); publish f --------------------------------
        if (chain::gather<H>(p.bH, B, act, nullptr, 0, tag, p.abort_flag, g_gn_deadline)) GN_FAIL();
        chain::dots<H>(W1, G + 2, act, B, tmp, LDT, 0);
        __syncthreads();
        for (int b = tid; b < B; b += NT) {
            // the fc1 unit of this CTA is unit `cta`; its table entry sits in .w of GRU unit 2 cta (= j0) -- see GnLoopParams::TA
            const float v = tmp[b * LDT + G] + fmaf(su[3 * U], xs_[b], cc[b * U].w);
            ll_store(p.bF + (size_t)b * F + cta, fmaxf(v, 0.f), tag);
        }
        for (int e = tid; e < B * G; e += NT) gh[e] = tmp[(e / G) * LDT + e % G];
        __syncthreads();
        // ---- 3: my classes of logits = fc3 f + b; publish ---------------------------------------------------------------------
        if (chain::gather<F>(p.bF, B, act, nullptr, 0, tag, p.abort_flag, g_gn_deadline)) GN_FAIL();
        if (cta * CR < C) {
            chain::dots<F>(W3, CR, act, B, tmp, LDT, 0);
            __syncthreads();
            for (int e = tid; e < B * CR; e += NT) {
                const int b = e / CR, r = e % CR, cls = cta * CR + r;
                if (cls < C) {
                    const float v = tmp[b * LDT + r] + sb3[r];
                    ll_store(p.bLG + (size_t)b * p.Cpad + cls, v, tag);
                    if (p.logits_out) p.logits_out[((size_t)b * p.S + t) * C + cls] = v;
                }
            }
        }
        __syncthreads();
        // ---- 4: sample the folds assigned to this CTA; publish x_t ------------------------------------------------------------
        {
            int failed = 0;
            for (int b = cta + gridDim.x * warp; b < B; b += gridDim.x * NW) {
                const FoldDesc fd = p.folds[b];
                const unsigned long long* row = p.bLG + (size_t)b * p.Cpad;
                float xs;
                if (p.mode == 1) {
                    bool ok;
                    xs = sample_mol_warp(row, tag, key, (uint32_t)t, fd, p.abort_flag, g_gn_deadline, ok);
                    if (!ok) { failed = 1; break; }
                } else {
                    uint4 r = philox4x32_10(make_uint4((uint32_t)t, (uint32_t)fd.fold, (uint32_t)fd.utt, 0u), key);
                    const float uu = u01(r.x);
                    int k;
                    if (C == 256) k = sample_raw_warp<8>(row, tag, uu, p.abort_flag, g_gn_deadline);
                    else if (C == 512) k = sample_raw_warp<16>(row, tag, uu, p.abort_flag, g_gn_deadline);
                    else k = sample_raw_warp<32>(row, tag, uu, p.abort_flag, g_gn_deadline);
                    if (k < 0) { failed = 1; break; }
                    xs = 2.0f * (float)k / ((float)C - 1.0f) - 1.0f;     // geneing_version.py:222 (fp32)
                }
                if (lane == 0) {
                    p.samples[(size_t)b * p.S + t] = xs;
                    const float fed = p.forced ? p.forced[(size_t)b * p.S + t] : xs;
                    ll_store(p.bX + b, fed, tag);
                }
            }
            if (__syncthreads_or(failed)) GN_FAIL();
        }
        if (cta == 0 && tid == 0 && (t % 100) == 0 && p.progress) {
            *reinterpret_cast<volatile int*>(p.progress) = t;
            __threadfence_system();
        }
    }
#undef GN_FAIL
}

size_t loop_gn_smem_bytes(int B, int CR) {
    size_t f = (size_t)(G + 2) * H + (size_t)CR * F + (size_t)B * H + (size_t)B * U * 4 + (size_t)B * (G + 2) + (size_t)B * U + (size_t)B * G +
               ((B + 3) & ~3) + 4 * U + 4 + ((CR + 3) & ~3) + kHop * kTaps;
    return f * sizeof(float);
}

cudaError_t set_gn_deadline(long long cycles) { return cudaMemcpyToSymbol(g_gn_deadline, &cycles, sizeof(cycles)); }

cudaError_t launch_loop_gn(const GnLoopParams& p, cudaStream_t stream) {
    const size_t smem = loop_gn_smem_bytes(p.B, p.CR);
    cudaError_t err = cudaFuncSetAttribute(wrnn_loop_gn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    GnLoopParams pp = p;
    void* args[] = {&pp};
    return cudaLaunchCooperativeKernel((const void*)wrnn_loop_gn_kernel, dim3(kGnCtas), dim3(NT), args, smem, stream);
}

}  // namespace wrnn
