// loop_f32.cu -- the autoregressive sample loop (reference: WaveRNN.generate body,
// vocoder/models/fatchord_version.py:192-236) as ONE persistent cooperative kernel, fp32 parity mode.
//
// Design:
//  * weight-stationary: the 128 CTAs (one per SM) each own 4 of the 512 hidden units of every layer;
//    their rows of W_hh1, W_ih2[:, :512], W_hh2, fc1[:, :512], fc2[:, :512] and fc3 stay in shared
//    memory (fp32, ~100 KB/SM) for the whole sequence -- weights are never re-read from HBM/L2.
//  * the I layer, W_ih1 and every mel/aux column of the other layers are folded into per-frame
//    conditioning tables by the front end (cond.cu); the residuals x1 = xI + h1, x2 = x1 + h2 are
//    expanded algebraically (W x1 = W xI + W h1), so only h1, h2, f1, f2 ever travel between SMs and a
//    step is five exchanges: h1 -> [W_hh1;W_ih2;W_fc1] -> h2 -> [W_hh2;W_fc1] -> f1 -> fc2 -> f2 -> fc3 -> x.
//  * activations travel between SMs as {value, step-tag} 8-byte words through L2 (common.cuh):
//    consumers spin on the data itself; there is no grid barrier, fence or atomic on the path.
//  * sampling (softmax + inverse CDF, or mixture-of-logistics) is fused; noise is Philox(step, fold,
//    utterance) so the run can be replayed on the reference (oracle/philox.py).
//  * every spin has a deadline; a miss raises a global abort flag and all CTAs leave (no GPU hang).
#include "engine_internal.h"
#include "sampling.cuh"

namespace wrnn {

namespace {

constexpr int NT = 512;         // threads per CTA
constexpr int NW = NT / 32;     // warps
constexpr int U = kUnitsF32;    // hidden units per CTA (4)
constexpr int G = 3 * U;        // gate rows per CTA per GRU matrix (12)
__device__ long long g_spin_deadline = 1500000000LL;   // SM clocks (~0.8 s); host-settable
#define kSpinDeadline g_spin_deadline

constexpr int RB = 2 * G + U;   // rows of stage B: W_hh1 (gh1 for the next step) | W_ih2a | W_fc1a        (28)
constexpr int RC = G + U;       // rows of stage C: W_hh2 (gh2 for the next step) | W_fc1a                  (16)
constexpr int LDT = RB;         // leading dimension of the per-chunk GEMV result buffer

struct Smem {
    float* WB;     // [RB][512]
    float* WC;     // [RC][512]
    float* WD;     // [U][512]   fc2[:, :512]
    float* WE;     // [CR][512]  fc3
    float* act;    // [FB][512]  gathered activations of the current chunk
    float* tmp;    // [FB][LDT]  GEMV results of the current chunk
    float* h1;     // [B][U]
    float* h2;     // [B][U]
    float* p3;     // [B][U]     fc1 partial  W_fc1a h1
    float* gh1;    // [B][G]     W_hh1 h1 for the next step
    float* gh2;    // [B][G]
    float4* c1;    // [B][U]     {gi1_r, gi1_z, gi1_n, fc1} conditioning of this step
    float4* c2;    // [B][U]     {gi2_r, gi2_z, gi2_n, fc2}
    float* x;      // [B]        previous sample of every fold
    float* v1;     // [3][U]
    float* v2;     // [3][U]
    float* v3;     // [U]
    float* bhn1;   // [U]
    float* bhn2;   // [U]
    float* bfc3;   // [8]
    float* coef;   // [200][kTaps]
};

__device__ __forceinline__ Smem carve(float* base, int B, int FB, int CR) {
    Smem s;
    float* p = base;
    s.WB = p;   p += RB * kRnn;
    s.WC = p;   p += RC * kRnn;
    s.WD = p;   p += U * kRnn;
    s.WE = p;   p += CR * kRnn;
    s.act = p;  p += FB * kRnn;
    s.tmp = p;  p += FB * LDT;
    s.c1 = reinterpret_cast<float4*>(p); p += B * U * 4;
    s.c2 = reinterpret_cast<float4*>(p); p += B * U * 4;
    s.h1 = p;   p += B * U;
    s.h2 = p;   p += B * U;
    s.p3 = p;   p += B * U;
    s.gh1 = p;  p += B * G;
    s.gh2 = p;  p += B * G;
    s.x = p;    p += (B + 3) & ~3;
    s.v1 = p;   p += 3 * U;
    s.v2 = p;   p += 3 * U;
    s.v3 = p;   p += U;
    s.bhn1 = p; p += U;
    s.bhn2 = p; p += U;
    s.bfc3 = p; p += 8;
    s.coef = p; p += kHop * kTaps;
    return s;
}

// Spin until every {value, tag} word of rows [b0, b0+nb) of `buf` carries `tag`; copy the values into
// act[nb][512].  Returns nonzero (CTA-uniform) if the deadline passed or another CTA aborted.
__device__ __noinline__ int gather(const unsigned long long* __restrict__ buf, int b0, int nb, float* __restrict__ act,
                                   uint32_t tag, int* abort_flag) {
    constexpr int MAXI = 8;
    const int tid = threadIdx.x;
    const int npairs = nb * (kRnn / 2);
    const unsigned long long* src = buf + (size_t)b0 * kRnn;
    int failed = 0;
    for (int base = 0; base < npairs; base += NT * MAXI) {
        uint32_t pending = 0;
#pragma unroll
        for (int i = 0; i < MAXI; ++i)
            if (base + tid + i * NT < npairs) pending |= 1u << i;
        long long t0 = 0;
        int spins = 0;
        while (pending) {
            unsigned long long a[MAXI], b[MAXI];
#pragma unroll
            for (int i = 0; i < MAXI; ++i)
                if ((pending >> i) & 1u) ll_load2(src + 2 * (size_t)(base + tid + i * NT), a[i], b[i]);
#pragma unroll
            for (int i = 0; i < MAXI; ++i)
                if ((pending >> i) & 1u) {
                    if (ll_tag(a[i]) == tag && ll_tag(b[i]) == tag) {
                        *reinterpret_cast<float2*>(act + 2 * (size_t)(base + tid + i * NT)) =
                            make_float2(ll_val(a[i]), ll_val(b[i]));
                        pending &= ~(1u << i);
                    }
                }
            if (pending && ((++spins) & 63) == 0) {
                if (t0 == 0) t0 = clock64();
                if (clock64() - t0 > kSpinDeadline || ld_volatile_i32(abort_flag) != 0) { failed = 1; break; }
            }
        }
        if (failed) break;
    }
    return __syncthreads_or(failed);
}

// out[f][r] = sum_k W[r][k] * act[f][k] for r < R, f < nb.  Warp tasks of RT rows x FT folds; lanes
// split K (float4, conflict-free) and reduce with shuffles.  Ends with __syncthreads().
template <int RT, int FT>
__device__ __noinline__ void dots(const float* __restrict__ sW, int R, const float* __restrict__ act, int nb,
                     float* __restrict__ out, int ldo) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nrg = (R + RT - 1) / RT, nfg = (nb + FT - 1) / FT;
    for (int task = warp; task < nrg * nfg; task += NW) {
        const int r0 = (task % nrg) * RT, f0 = (task / nrg) * FT;
        float acc[RT][FT];
#pragma unroll
        for (int r = 0; r < RT; ++r)
#pragma unroll
            for (int f = 0; f < FT; ++f) acc[r][f] = 0.f;
#pragma unroll
        for (int kk = 0; kk < kRnn / 128; ++kk) {
            const int k = kk * 128 + lane * 4;
            float4 w[RT], a[FT];
#pragma unroll
            for (int r = 0; r < RT; ++r) w[r] = *reinterpret_cast<const float4*>(sW + min(r0 + r, R - 1) * kRnn + k);
#pragma unroll
            for (int f = 0; f < FT; ++f) a[f] = *reinterpret_cast<const float4*>(act + min(f0 + f, nb - 1) * kRnn + k);
#pragma unroll
            for (int r = 0; r < RT; ++r)
#pragma unroll
                for (int f = 0; f < FT; ++f) {
                    acc[r][f] = fmaf(w[r].x, a[f].x, acc[r][f]);
                    acc[r][f] = fmaf(w[r].y, a[f].y, acc[r][f]);
                    acc[r][f] = fmaf(w[r].z, a[f].z, acc[r][f]);
                    acc[r][f] = fmaf(w[r].w, a[f].w, acc[r][f]);
                }
        }
#pragma unroll
        for (int r = 0; r < RT; ++r)
#pragma unroll
            for (int f = 0; f < FT; ++f) {
                float v = warp_sum(acc[r][f]);
                if (lane == r * FT + f && r0 + r < R && f0 + f < nb) out[(f0 + f) * ldo + r0 + r] = v;
            }
    }
    __syncthreads();
}

__device__ __forceinline__ void dots_any(bool single, const float* sW, int R, const float* act, int nb, float* out, int ldo) {
    if (single) dots<1, 1>(sW, R, act, nb, out, ldo);
    else dots<4, 4>(sW, R, act, nb, out, ldo);
}

}  // namespace

__global__ void __launch_bounds__(NT, 1) wrnn_loop_f32_kernel(LoopParams p) {
    extern __shared__ __align__(16) float smem_f[];
    const int cta = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int B = p.B, FB = p.FB, C = p.C, CR = p.CR;
    const int j0 = cta * U;
    Smem s = carve(smem_f, B, FB, CR);
    const bool single = (B == 1);

    // ---- one-time: my weight rows into shared memory ------------------------------------------------
    for (int i = tid; i < G * (kRnn / 4); i += NT) {
        const int row = i / (kRnn / 4), k4 = i % (kRnn / 4);
        const int g = row / U, u = row % U;
        const size_t src = ((size_t)(g * kRnn + j0 + u) * kRnn) / 4 + k4;
        reinterpret_cast<float4*>(s.WB)[i] = reinterpret_cast<const float4*>(p.Whh1)[src];
        reinterpret_cast<float4*>(s.WB)[G * (kRnn / 4) + i] = reinterpret_cast<const float4*>(p.Wih2a)[src];
        reinterpret_cast<float4*>(s.WC)[i] = reinterpret_cast<const float4*>(p.Whh2)[src];
    }
    for (int i = tid; i < U * (kRnn / 4); i += NT) {
        const size_t src = ((size_t)j0 * kRnn) / 4 + i;
        const float4 w1 = reinterpret_cast<const float4*>(p.Wfc1a)[src];
        reinterpret_cast<float4*>(s.WB)[2 * G * (kRnn / 4) + i] = w1;
        reinterpret_cast<float4*>(s.WC)[G * (kRnn / 4) + i] = w1;
        reinterpret_cast<float4*>(s.WD)[i] = reinterpret_cast<const float4*>(p.Wfc2a)[src];
    }
    for (int i = tid; i < CR * (kRnn / 4); i += NT) {
        const int row = i / (kRnn / 4), cls = cta * CR + row;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cls < C) v = reinterpret_cast<const float4*>(p.Wfc3)[(size_t)cls * (kRnn / 4) + i % (kRnn / 4)];
        reinterpret_cast<float4*>(s.WE)[i] = v;
    }
    if (tid < 3 * U) {
        s.v1[tid] = p.v1[(tid / U) * kRnn + j0 + tid % U];
        s.v2[tid] = p.v2[(tid / U) * kRnn + j0 + tid % U];
    }
    if (tid < U) {
        s.v3[tid] = p.v3[j0 + tid];
        s.bhn1[tid] = p.bhn1[j0 + tid];
        s.bhn2[tid] = p.bhn2[j0 + tid];
    }
    if (tid < CR) s.bfc3[tid] = (cta * CR + tid < C) ? p.bfc3[cta * CR + tid] : 0.f;
    for (int i = tid; i < kHop * kTaps; i += NT) s.coef[i] = p.coef[i];
    for (int i = tid; i < B * U; i += NT) { s.h1[i] = 0.f; s.h2[i] = 0.f; s.p3[i] = 0.f; }
    for (int i = tid; i < B * G; i += NT) { s.gh1[i] = 0.f; s.gh2[i] = 0.f; }   // W_hh * 0
    for (int i = tid; i < B; i += NT) s.x[i] = 0.f;                              // x_0 = 0, fatchord_version.py:183
    __syncthreads();

    const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32));
    const int nchunks = (B + FB - 1) / FB;

    for (int t = 0; t < p.S; ++t) {
        const uint32_t tag = (uint32_t)t + 1u;

        // ---- conditioning of this step for my units (independent of the exchange: issued before the wait) ----
        for (int e = tid; e < B * U; e += NT) {
            const int b = e / U, j = j0 + e % U;
            const FoldDesc fd = p.folds[b];
            const int n = fd.n0 + t;
            const bool valid = n < fd.N;                 // positions past the utterance are fold tail padding (Q9)
            const int q0 = valid ? n / kHop : 0;
            const size_t ra = (size_t)(fd.ta_row0 + (valid ? q0 : fd.T)) * kRnn + j;
            float4 a1 = __ldg(p.TA1 + ra), a2 = __ldg(p.TA2 + ra);
            if (valid) {
                const float* cf = s.coef + (n - q0 * kHop) * kTaps;
#pragma unroll
                for (int d = 0; d < kTaps; ++d) {
                    const float c = cf[d];
                    if (c != 0.f) {
                        const size_t rq = (size_t)(fd.tq_row0 + q0 + d) * kRnn + j;
                        const float4 q1 = __ldg(p.TQ1 + rq), q2 = __ldg(p.TQ2 + rq);
                        a1.x = fmaf(c, q1.x, a1.x); a1.y = fmaf(c, q1.y, a1.y); a1.z = fmaf(c, q1.z, a1.z); a1.w = fmaf(c, q1.w, a1.w);
                        a2.x = fmaf(c, q2.x, a2.x); a2.y = fmaf(c, q2.y, a2.y); a2.z = fmaf(c, q2.z, a2.z);
                    }
                }
            }
            s.c1[e] = a1;
            s.c2[e] = a2;
        }

        // ---- A: wait x_{t-1}; GRU1 update for my units; publish h1 -------------------------------------------
        {
            int failed = 0;
            if (t > 0) {
                for (int b = tid; b < B; b += NT) {
                    float v;
                    if (!wait_word(p.bX + b, (uint32_t)t, v, p.abort_flag, kSpinDeadline)) failed = 1;
                    s.x[b] = v;
                }
            }
            if (__syncthreads_or(failed)) { if (tid == 0) atomicExch(p.abort_flag, 1); return; }
            for (int e = tid; e < B * U; e += NT) {
                const int b = e / U, u = e % U;
                const float x = s.x[b];
                const float4 c = s.c1[e];
                const float* gh = s.gh1 + b * G;
                const float r = sigmoid_acc(fmaf(s.v1[0 * U + u], x, c.x) + gh[0 * U + u]);
                const float z = sigmoid_acc(fmaf(s.v1[1 * U + u], x, c.y) + gh[1 * U + u]);
                const float nn = tanhf(fmaf(s.v1[2 * U + u], x, c.z) + r * (gh[2 * U + u] + s.bhn1[u]));
                const float h = (1.0f - z) * nn + z * s.h1[e];
                s.h1[e] = h;
                ll_store(p.bH1 + (size_t)b * kRnn + j0 + u, h, tag);
            }
        }

        // ---- B: [gh1' ; W_ih2a h1 ; W_fc1a h1] ; GRU2 update ; publish h2 --------------------------------------
        for (int c = 0; c < nchunks; ++c) {
            const int b0 = c * FB, nb = min(FB, B - b0);
            if (gather(p.bH1, b0, nb, s.act, tag, p.abort_flag)) { if (tid == 0) atomicExch(p.abort_flag, 1); return; }
            dots_any(single, s.WB, RB, s.act, nb, s.tmp, LDT);
            for (int e = tid; e < nb * U; e += NT) {
                const int bl = e / U, u = e % U, b = b0 + bl;
                const float* r_ = s.tmp + bl * LDT;
                const float x = s.x[b];
                const float4 cc = s.c2[b * U + u];
                const float* gh = s.gh2 + b * G;
                const float r = sigmoid_acc(r_[G + 0 * U + u] + fmaf(s.v2[0 * U + u], x, cc.x) + gh[0 * U + u]);
                const float z = sigmoid_acc(r_[G + 1 * U + u] + fmaf(s.v2[1 * U + u], x, cc.y) + gh[1 * U + u]);
                const float nn = tanhf(r_[G + 2 * U + u] + fmaf(s.v2[2 * U + u], x, cc.z) + r * (gh[2 * U + u] + s.bhn2[u]));
                const float h = (1.0f - z) * nn + z * s.h2[b * U + u];
                s.h2[b * U + u] = h;
                s.p3[b * U + u] = r_[2 * G + u];
                ll_store(p.bH2 + (size_t)b * kRnn + j0 + u, h, tag);
            }
            for (int e = tid; e < nb * G; e += NT) s.gh1[(b0 + e / G) * G + e % G] = s.tmp[(e / G) * LDT + e % G];
            __syncthreads();
        }
        // ---- C: [gh2' ; W_fc1a h2] ; f1 = relu(fc1) ; publish ----------------------------------------------------
        for (int c = 0; c < nchunks; ++c) {
            const int b0 = c * FB, nb = min(FB, B - b0);
            if (gather(p.bH2, b0, nb, s.act, tag, p.abort_flag)) { if (tid == 0) atomicExch(p.abort_flag, 1); return; }
            dots_any(single, s.WC, RC, s.act, nb, s.tmp, LDT);
            for (int e = tid; e < nb * U; e += NT) {
                const int bl = e / U, u = e % U, b = b0 + bl;
                const float v = s.p3[b * U + u] + s.tmp[bl * LDT + G + u] + fmaf(s.v3[u], s.x[b], s.c1[b * U + u].w);
                ll_store(p.bF1 + (size_t)b * kRnn + j0 + u, fmaxf(v, 0.f), tag);
            }
            for (int e = tid; e < nb * G; e += NT) s.gh2[(b0 + e / G) * G + e % G] = s.tmp[(e / G) * LDT + e % G];
            __syncthreads();
        }
        // ---- D: f2 = relu(fc2[:, :512] f1 + c4) ; publish -------------------------------------------------------
        for (int c = 0; c < nchunks; ++c) {
            const int b0 = c * FB, nb = min(FB, B - b0);
            if (gather(p.bF1, b0, nb, s.act, tag, p.abort_flag)) { if (tid == 0) atomicExch(p.abort_flag, 1); return; }
            dots_any(single, s.WD, U, s.act, nb, s.tmp, LDT);
            for (int e = tid; e < nb * U; e += NT) {
                const int bl = e / U, u = e % U, b = b0 + bl;
                ll_store(p.bF2 + (size_t)b * kRnn + j0 + u, fmaxf(s.tmp[bl * LDT + u] + s.c2[b * U + u].w, 0.f), tag);
            }
            __syncthreads();
        }
        // ---- E: my classes of logits = fc3 f2 + b ; publish ---------------------------------------------------
        for (int c = 0; c < nchunks; ++c) {
            const int b0 = c * FB, nb = min(FB, B - b0);
            if (gather(p.bF2, b0, nb, s.act, tag, p.abort_flag)) { if (tid == 0) atomicExch(p.abort_flag, 1); return; }
            if (cta * CR < C) {
                dots_any(single, s.WE, CR, s.act, nb, s.tmp, LDT);
                for (int e = tid; e < nb * CR; e += NT) {
                    const int bl = e / CR, r = e % CR, b = b0 + bl, cls = cta * CR + r;
                    if (cls < C) {
                        const float v = s.tmp[bl * LDT + r] + s.bfc3[r];
                        ll_store(p.bLG + (size_t)b * p.Cpad + cls, v, tag);
                        if (p.logits_out) p.logits_out[((size_t)b * p.S + t) * C + cls] = v;
                    }
                }
            }
            __syncthreads();
        }
        // ---- F: sample the folds assigned to this CTA ; publish x_t ----------------------------------------
        {
            int failed = 0;
            for (int b = cta + gridDim.x * warp; b < B; b += gridDim.x * NW) {
                const FoldDesc fd = p.folds[b];
                const unsigned long long* row = p.bLG + (size_t)b * p.Cpad;
                float xs;
                if (p.mode == 1) {
                    bool ok;
                    xs = sample_mol_warp(row, tag, key, (uint32_t)t, fd, p.abort_flag, kSpinDeadline, ok);
                    if (!ok) { failed = 1; break; }
                } else {
                    uint4 r = philox4x32_10(make_uint4((uint32_t)t, (uint32_t)fd.fold, (uint32_t)fd.utt, 0u), key);
                    const float u = u01(r.x);
                    int k;
                    if (C == 256) k = sample_raw_warp<8>(row, tag, u, p.abort_flag, kSpinDeadline);
                    else if (C == 512) k = sample_raw_warp<16>(row, tag, u, p.abort_flag, kSpinDeadline);
                    else k = sample_raw_warp<32>(row, tag, u, p.abort_flag, kSpinDeadline);
                    if (k < 0) { failed = 1; break; }
                    xs = 2.0f * (float)k / ((float)C - 1.0f) - 1.0f;     // fatchord_version.py:228 (fp32, Q10)
                }
                if (lane == 0) {
                    p.samples[(size_t)b * p.S + t] = xs;
                    const float fed = p.forced ? p.forced[(size_t)b * p.S + t] : xs;
                    ll_store(p.bX + b, fed, tag);
                }
            }
            if (__syncthreads_or(failed)) { if (tid == 0) atomicExch(p.abort_flag, 1); return; }
        }
        if (cta == 0 && tid == 0 && (t % 100) == 0 && p.progress) {
            *reinterpret_cast<volatile int*>(p.progress) = t;
            __threadfence_system();
        }
    }
}

size_t loop_f32_smem_bytes(int B, int FB, int CR) {
    size_t f = (size_t)(RB + RC + U) * kRnn + (size_t)CR * kRnn + (size_t)FB * kRnn + (size_t)FB * LDT +
               (size_t)B * (8 * U + 3 * U + 2 * G) + ((B + 3) & ~3) + 6 * U + 3 * U + 8 + kHop * kTaps;
    return f * sizeof(float);
}

int loop_f32_pick_fb(int B, int CR, size_t smem_limit) {
    int fb = B < 32 ? B : 32;
    while (fb > 1 && loop_f32_smem_bytes(B, fb, CR) > smem_limit) --fb;
    return loop_f32_smem_bytes(B, fb, CR) <= smem_limit ? fb : 0;
}

cudaError_t set_spin_deadline(long long cycles) { return cudaMemcpyToSymbol(g_spin_deadline, &cycles, sizeof(cycles)); }

cudaError_t launch_loop_f32(const LoopParams& p, cudaStream_t stream) {
    const size_t smem = loop_f32_smem_bytes(p.B, p.FB, p.CR);
    cudaError_t err = cudaFuncSetAttribute(wrnn_loop_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    LoopParams pp = p;
    void* args[] = {&pp};
    // cooperative launch: guarantees the 128 CTAs are co-resident (they wait on one another)
    return cudaLaunchCooperativeKernel((const void*)wrnn_loop_f32_kernel, dim3(kCtasF32), dim3(NT), args, smem, stream);
}

}  // namespace wrnn
