// post.cu -- everything WaveRNN.generate does after the loop (fatchord_version.py:238-255), on the GPU in
// float64 like the reference's numpy code:
//   xfade_and_unfold (:342-404)  -> decode_mu_law (vocoder/audio.py:102-107, AFTER the crossfade, Q6)
//   -> de_emphasis (audio.py:92-93, lfilter([1],[1,-0.97]))  -> [:wave_len]  -> 4000-sample linear fade-out.
// Unfold indices are bit-exact integer arithmetic; each output sample has at most two addends, so the
// float64 sums are order-independent and bit-identical to numpy's.  The IIR is evaluated as a blocked
// parallel scan with a 2048-sample warm-up (0.97^2048 = 8e-28, far below one ulp).
#include "engine_internal.h"

namespace wrnn {

namespace {

constexpr double kPreemph = 0.97;     // sp.preemphasis, config/hparams.py:49
constexpr int kFadeLen = 20 * kHop;   // fatchord_version.py:253

__global__ void unfold_decode_kernel(const float* __restrict__ samples, const PostUtt* __restrict__ utts, int S,
                                     int batched, int target, int overlap, const double* __restrict__ fade_in,
                                     const double* __restrict__ fade_out, int mu_law, int n_classes,
                                     double* __restrict__ z) {
    const PostUtt u = utts[blockIdx.y];
    const float* y = samples + u.samp_off;
    const int stride = target + overlap;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < u.wave_len; p += gridDim.x * blockDim.x) {
        double v;
        if (!batched) {
            v = (double)y[p];
        } else {
            v = 0.0;
            int hi = p / stride;
            if (hi > u.F - 1) hi = u.F - 1;
            // folds are added in increasing order like the reference's loop (:400-403)
            for (int i = (hi > 0 ? hi - 1 : 0); i <= hi; ++i) {
                const int q = p - i * stride;
                if (q < 0 || q >= S) continue;
                double g = (double)y[(size_t)i * S + q];
                if (q < overlap) g = __dmul_rn(g, fade_in[q]);
                if (q >= S - overlap) g = __dmul_rn(g, fade_out[q - (S - overlap)]);
                v = __dadd_rn(v, g);
            }
        }
        if (mu_law) {
            const double mu = (double)(n_classes - 1);
            const double sgn = (v > 0.0) ? 1.0 : ((v < 0.0) ? -1.0 : 0.0);
            v = sgn / mu * (pow(1.0 + mu, fabs(v)) - 1.0);
        }
        z[u.wav_off + p] = v;
    }
}

// y[n] = z[n] + a*y[n-1].  Block: 384 threads x 16 samples = 2048 warm-up + 4096 outputs.
constexpr int kSeg = 16, kScanThreads = 384, kWarm = 2048, kOutPerBlock = kScanThreads * kSeg - kWarm;

__global__ void __launch_bounds__(kScanThreads) deemph_fade_kernel(const double* __restrict__ z,
                                                                   const PostUtt* __restrict__ utts, int preemph,
                                                                   double* __restrict__ wav) {
    __shared__ double carry[kScanThreads];
    const PostUtt u = utts[blockIdx.y];
    const int start = blockIdx.x * kOutPerBlock;
    if (start >= u.wave_len) return;
    const double* zi = z + u.wav_off;
    const int tid = threadIdx.x;
    const int s0 = start - kWarm + tid * kSeg;
    double yl[kSeg];
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < kSeg; ++k) {
        const int n = s0 + k;
        const double x = (n >= 0 && n < u.wave_len) ? zi[n] : 0.0;
        acc = preemph ? (x + kPreemph * acc) : x;
        yl[k] = acc;
    }
    if (preemph) {
        // carry[i] = filter state entering thread i's segment: c_i = a^16 c_{i-1} + yl_{i-1}[15]
        double apow = 1.0;
#pragma unroll
        for (int k = 0; k < kSeg; ++k) apow *= kPreemph;   // a^16
        carry[tid] = (tid + 1 < kScanThreads) ? yl[kSeg - 1] : 0.0;
        __syncthreads();
        // shift by one: value for thread i is the inclusive scan of ends up to thread i-1
        double mine = (tid > 0) ? carry[tid - 1] : 0.0;
        __syncthreads();
        carry[tid] = mine;
        __syncthreads();
        double f = apow;
        for (int off = 1; off < kScanThreads; off <<= 1) {
            double add = (tid >= off) ? f * carry[tid - off] : 0.0;
            __syncthreads();
            carry[tid] += add;
            __syncthreads();
            f *= f;
        }
        const double c = carry[tid];
        double ap = kPreemph;
#pragma unroll
        for (int k = 0; k < kSeg; ++k) {
            yl[k] += ap * c;
            ap *= kPreemph;
        }
    }
    const double step = (0.0 - 1.0) / (double)(kFadeLen - 1);   // np.linspace(1, 0, 4000)
#pragma unroll
    for (int k = 0; k < kSeg; ++k) {
        const int n = s0 + k;
        if (n >= start && n < u.wave_len && n < start + kOutPerBlock) {
            double v = yl[k];
            const int fi = n - (u.wave_len - kFadeLen);
            if (fi >= 0) {
                const double g = (fi == kFadeLen - 1) ? 0.0 : __dadd_rn(__dmul_rn((double)fi, step), 1.0);
                v = __dmul_rn(v, g);
            }
            wav[u.wav_off + n] = v;
        }
    }
}

// xfade_and_unfold alone on float64 input (the helper callers may use directly, fatchord_version.py:342).
__global__ void xfade_unfold_f64_kernel(const double* __restrict__ y, int F, int S, int overlap,
                                        const double* __restrict__ fade_in, const double* __restrict__ fade_out,
                                        long long total_len, double* __restrict__ out) {
    const int stride = S - overlap;   // target + overlap
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total_len; p += (long long)gridDim.x * blockDim.x) {
        double v = 0.0;
        int hi = (int)(p / stride);
        if (hi > F - 1) hi = F - 1;
        for (int i = (hi > 0 ? hi - 1 : 0); i <= hi; ++i) {
            const long long q = p - (long long)i * stride;
            if (q < 0 || q >= S) continue;
            double g = y[(size_t)i * S + q];
            if (q < overlap) g = __dmul_rn(g, fade_in[q]);
            if (q >= S - overlap) g = __dmul_rn(g, fade_out[q - (S - overlap)]);
            v = __dadd_rn(v, g);
        }
        out[p] = v;
    }
}

}  // namespace

cudaError_t launch_xfade_unfold_f64(const double* y, int F, int S, int overlap, const double* fade_in,
                                    const double* fade_out, long long total_len, double* out, cudaStream_t stream) {
    long long bx = (total_len + 255) / 256;
    if (bx > 148 * 8) bx = 148 * 8;
    xfade_unfold_f64_kernel<<<(int)bx, 256, 0, stream>>>(y, F, S, overlap, fade_in, fade_out, total_len, out);
    return cudaGetLastError();
}

cudaError_t launch_post(const float* samples, const PostUtt* utts, int n_utts, int max_wave_len, int S, int batched,
                        int target, int overlap, const double* fade_in, const double* fade_out, int mu_law, int n_classes,
                        int preemph, double* scratch, double* wav, cudaStream_t stream) {
    if (n_utts <= 0 || max_wave_len <= 0) return cudaSuccess;
    {
        int bx = (max_wave_len + 255) / 256;
        if (bx > 148 * 8) bx = 148 * 8;
        unfold_decode_kernel<<<dim3(bx, n_utts), 256, 0, stream>>>(samples, utts, S, batched, target, overlap, fade_in,
                                                                   fade_out, mu_law, n_classes, scratch);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    const int bx = (max_wave_len + kOutPerBlock - 1) / kOutPerBlock;
    deemph_fade_kernel<<<dim3(bx, n_utts), kScanThreads, 0, stream>>>(scratch, utts, preemph, wav);
    return cudaGetLastError();
}

}  // namespace wrnn
