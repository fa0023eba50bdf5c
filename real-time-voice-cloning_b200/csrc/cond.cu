// cond.cu -- conditioning front end (reference: UpsampleNetwork.forward / MelResNet / Stretch2d,
// vocoder/models/fatchord_version.py:9-85) restructured for the persistent loop:
//   * BatchNorm (eval) is folded into the 1x1 / k=5 convolutions at load time (engine.cu), so the
//     MelResNet is a chain of 22 dense [frames x K] x [K x 128] contractions with fused ReLU/residual.
//   * Stretch2d is never materialised: aux is constant over a 200-sample frame, and the three
//     stretch+box-filter layers are a fixed 5-tap interpolation of padded mel frames whose weights
//     depend only on the sample's phase (n mod 200) -- table `coef` built in engine.cu.
//   * the mel/aux columns of I, rnn2.W_ih, fc1, fc2 -- and rnn1.W_ih applied to the I layer's
//     output -- are pre-multiplied here into per-frame tables TA1/TA2/TA3/TQ (engine_internal.h), so
//     the loop never touches an 80- or 128-wide conditioning vector.
// This file holds the fp32 SIMT form (exact parity mode).
#include "engine_internal.h"
#include "cond_expand.cuh"

namespace wrnn {

namespace {

__device__ __forceinline__ int find_utt_by_ta(const UttDesc* utts, int n, int row) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (utts[mid].ta_row0 <= row) lo = mid; else hi = mid - 1;
    }
    return lo;
}
__device__ __forceinline__ int find_utt_by_tq(const UttDesc* utts, int n, int row) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (utts[mid].tq_row0 <= row) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// X0[row][c*5 + j] = melpad[c][f + j]  (conv_in, fatchord_version.py:31,39; pad_tensor :171,275-288)
// MP[row][c]       = melpad[c][fp]
__global__ void im2col_kernel(const float* __restrict__ mel, const UttDesc* __restrict__ utts, int n_utts,
                              int ta_rows, int tq_rows, float* __restrict__ X0, float* __restrict__ MP) {
    const long long total0 = (long long)ta_rows * (kFeat * 5);
    const long long total1 = (long long)tq_rows * kFeat;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total0 + total1;
         i += (long long)gridDim.x * blockDim.x) {
        if (i < total0) {
            const int row = (int)(i / (kFeat * 5)), col = (int)(i % (kFeat * 5));
            const int c = col / 5, j = col % 5;
            const UttDesc u = utts[find_utt_by_ta(utts, n_utts, row)];
            const int f = row - u.ta_row0;          // 0..T (row T is the bias-only row: any finite input)
            const int t = f + j - kPad;
            float v = 0.f;
            if (f < u.T && t >= 0 && t < u.T) v = mel[u.mel_off + (long long)c * u.T + t];
            X0[i] = v;
        } else {
            const long long k = i - total0;
            const int row = (int)(k / kFeat), c = (int)(k % kFeat);
            const UttDesc u = utts[find_utt_by_tq(utts, n_utts, row)];
            const int t = row - u.tq_row0 - kPad;
            float v = 0.f;
            if (t >= 0 && t < u.T) v = mel[u.mel_off + (long long)c * u.T + t];
            MP[k] = v;
        }
    }
}

__global__ void zero_rows_kernel(float* __restrict__ aux, const UttDesc* __restrict__ utts, int n_utts) {
    const int u = blockIdx.x;
    if (u < n_utts) aux[(size_t)(utts[u].ta_row0 + utts[u].T) * blockDim.x + threadIdx.x] = 0.f;     // (blockDim.x = channels per row)
}

// 64x64 tile, BK=16, 256 threads, 4x4 outputs per thread.
constexpr int TM = 64, TN = 64, TK = 16;
__global__ void __launch_bounds__(256) gemm_f32_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                       const float* __restrict__ bias, const float* __restrict__ R,
                                                       float* __restrict__ C, int M, int N, int K, int relu) {
    __shared__ float sA[TK][TM + 4];
    __shared__ float sW[TK][TN + 4];
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int lr = tid / 4, lk = (tid % 4) * 4;   // 64 rows x 4 float4 per k-tile
    for (int k0 = 0; k0 < K; k0 += TK) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m0 + lr < M) a = *reinterpret_cast<const float4*>(A + (size_t)(m0 + lr) * K + k0 + lk);
        const float4 w = *reinterpret_cast<const float4*>(W + (size_t)(n0 + lr) * K + k0 + lk);
        sA[lk + 0][lr] = a.x; sA[lk + 1][lr] = a.y; sA[lk + 2][lr] = a.z; sA[lk + 3][lr] = a.w;
        sW[lk + 0][lr] = w.x; sW[lk + 1][lr] = w.y; sW[lk + 2][lr] = w.z; sW[lk + 3][lr] = w.w;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            const float4 av = *reinterpret_cast<const float4*>(&sA[k][ty * 4]);
            const float4 wv = *reinterpret_cast<const float4*>(&sW[k][tx * 4]);
            const float aa[4] = {av.x, av.y, av.z, av.w}, ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], ww[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
        float4 o;
        float* op = &o.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            float v = acc[i][j] + (bias ? bias[n] : 0.f);
            if (relu) v = fmaxf(v, 0.f);
            if (R) v += R[(size_t)m * N + n];
            op[j] = v;
        }
        *reinterpret_cast<float4*>(C + (size_t)m * N + n0 + tx * 4) = o;
    }
}

// Per-sample conditioning for the tensor-core loop (record layout: cond_expand.cuh), expanded ahead of the loop.
__global__ void __launch_bounds__(256) expand_cond_kernel(const float4* __restrict__ TA1, const float4* __restrict__ TA2,
                                                          const float4* __restrict__ TQ1, const float4* __restrict__ TQ2,
                                                          const float* __restrict__ coef, const FoldDesc* __restrict__ folds,
                                                          int S, int cs_steps, int Mg, float4* __restrict__ CS) {
    const int b = blockIdx.x;
    const FoldDesc fd = folds[b];
    expand_cond_item(TA1, TA2, TQ1, TQ2, coef, fd, b, blockIdx.y * kExpandSteps, min(S, (int)(blockIdx.y + 1) * kExpandSteps), cs_steps, Mg, CS,
                     threadIdx.x);
}

// Same expansion for the cluster-local loop (loop_tc2.cu): CS[cluster][t][CTA 16][gate 4][unit 32][fold 32][2],
// values {gi1_g, gi2_g} for gates g = r,z,n and {fc1, fc2} for g = 3.  Block = (cluster, CTA, 8 steps); thread = (unit, fold).
__global__ void __launch_bounds__(1024) expand_cond2_kernel(const float4* __restrict__ TA1, const float4* __restrict__ TA2,
                                                            const float4* __restrict__ TQ1, const float4* __restrict__ TQ2,
                                                            const float* __restrict__ coef, const FoldDesc* __restrict__ folds,
                                                            int B, int Bc, int S, float* __restrict__ CS) {
    __shared__ float2 tile[4][32][33];                    // [gate][unit][fold], padded
    const int cl = blockIdx.x >> 4, crank = blockIdx.x & 15;
    const int u = threadIdx.x & 31, fold = threadIdx.x >> 5, j = crank * 32 + u;     // lanes along units: coalesced table reads
    const int b = cl * Bc + fold;
    const bool live = fold < Bc && b < B;
    FoldDesc fd;
    if (live) fd = folds[b];
    const int t1 = min(S, (int)(blockIdx.y + 1) * 8);
    for (int t = blockIdx.y * 8; t < t1; ++t) {
        float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), a2 = a1;
        if (live) {
            const int n = fd.n0 + t;
            const bool valid = n < fd.N;
            const int q0 = valid ? n / kHop : 0;
            const size_t ra = (size_t)(fd.ta_row0 + (valid ? q0 : fd.T)) * kRnn + j;
            a1 = __ldg(TA1 + ra); a2 = __ldg(TA2 + ra);
            if (valid) {
                const float* cf = coef + (n - q0 * kHop) * kTaps;
#pragma unroll
                for (int d = 0; d < kTaps; ++d) {
                    const float c = __ldg(cf + d);
                    if (c != 0.f) {
                        const size_t rq = (size_t)(fd.tq_row0 + q0 + d) * kRnn + j;
                        const float4 q1 = __ldg(TQ1 + rq), q2 = __ldg(TQ2 + rq);
                        a1.x = fmaf(c, q1.x, a1.x); a1.y = fmaf(c, q1.y, a1.y); a1.z = fmaf(c, q1.z, a1.z); a1.w = fmaf(c, q1.w, a1.w);
                        a2.x = fmaf(c, q2.x, a2.x); a2.y = fmaf(c, q2.y, a2.y); a2.z = fmaf(c, q2.z, a2.z);
                    }
                }
            }
        }
        __syncthreads();                                   // previous step's tile fully written out
        tile[0][u][fold] = make_float2(a1.x, a2.x);
        tile[1][u][fold] = make_float2(a1.y, a2.y);
        tile[2][u][fold] = make_float2(a1.z, a2.z);
        tile[3][u][fold] = make_float2(a1.w, a2.w);
        __syncthreads();
        float2* out = reinterpret_cast<float2*>(CS) + (((size_t)cl * S + t) * 16 + crank) * 4 * 32 * 32;
        const int of = threadIdx.x & 31, ou = threadIdx.x >> 5;                     // lanes along folds: coalesced stores
#pragma unroll
        for (int g = 0; g < 4; ++g) __stcs(out + (g * 32 + ou) * 32 + of, tile[g][ou][of]);
    }
}

}  // namespace

cudaError_t launch_expand_cond2(const float4* TA1, const float4* TA2, const float4* TQ1, const float4* TQ2, const float* coef,
                                const FoldDesc* folds, int B, int Bc, int n_clusters, int S, float* CS, cudaStream_t stream) {
    dim3 grid(n_clusters * 16, (S + 7) / 8);
    expand_cond2_kernel<<<grid, 1024, 0, stream>>>(TA1, TA2, TQ1, TQ2, coef, folds, B, Bc, S, CS);
    return cudaGetLastError();
}

cudaError_t launch_expand_cond(const float4* TA1, const float4* TA2, const float4* TQ1, const float4* TQ2, const float* coef,
                               const FoldDesc* folds, int B, int S, int Mg, float4* CS, cudaStream_t stream) {
    dim3 grid(B, (S + kExpandSteps - 1) / kExpandSteps);
    expand_cond_kernel<<<grid, 256, 0, stream>>>(TA1, TA2, TQ1, TQ2, coef, folds, S, (int)grid.y * kExpandSteps, Mg, CS);
    return cudaGetLastError();
}

cudaError_t launch_im2col(const float* mel, const UttDesc* utts, int n_utts, int ta_rows, int tq_rows, float* X0,
                          float* MP, cudaStream_t stream) {
    const long long total = (long long)ta_rows * 400 + (long long)tq_rows * kFeat;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    im2col_kernel<<<blocks, 256, 0, stream>>>(mel, utts, n_utts, ta_rows, tq_rows, X0, MP);
    return cudaGetLastError();
}

cudaError_t launch_gemm_f32(const float* A, const float* W, const float* bias, const float* R, float* C, int M, int N,
                            int K, int relu, cudaStream_t stream) {
    if (K % TK != 0 || N % TN != 0) return cudaErrorInvalidValue;
    dim3 grid(N / TN, (M + TM - 1) / TM);
    gemm_f32_kernel<<<grid, 256, 0, stream>>>(A, W, bias, R, C, M, N, K, relu);
    return cudaGetLastError();
}

cudaError_t launch_zero_rows(float* aux, const UttDesc* utts, int n_utts, cudaStream_t stream, int channels) {
    zero_rows_kernel<<<n_utts, channels, 0, stream>>>(aux, utts, n_utts);
    return cudaGetLastError();
}

}  // namespace wrnn
