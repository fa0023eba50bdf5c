// common.cuh -- device helpers shared by the WaveRNN loop kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace wrnn {

constexpr int kRnn = 512;          // rnn_dims = fc_dims (config/hparams.py:226-227)
constexpr int kHop = 200;          // sp.hop_size
constexpr int kAux = 32;           // res_out_dims / 4
constexpr int kFeat = 80;          // sp.num_mels
constexpr int kPad = 2;            // hparams.pad
constexpr int kTaps = 5;           // padded frames a sample's upsampled mel can touch (reach +-248 samples)
constexpr int kMaxClasses = 1024;

// One unit of work of the sample loop: one fold of one utterance.
struct FoldDesc {
    int ta_row0;   // first row of this utterance in the per-frame aux tables (T+1 rows, last = bias-only row)
    int tq_row0;   // first row of this utterance in the per-padded-frame mel tables (T+4 rows)
    int T;         // frames of the utterance
    int N;         // samples of the utterance = 200*T (positions >= N are fold tail padding, Q9)
    int n0;        // first sample of this fold: fold * (target + overlap)
    int utt;       // Philox utterance counter
    int fold;      // Philox fold counter (fold index inside the utterance)
    int pad_;
};

// ------------------------------------------------------------------------------------------------
// Flag-in-data exchange words: {fp32 value, 32-bit tag} stored/loaded as one aligned 8-byte access,
// which is single-copy atomic, so a consumer that sees the tag sees the value (the trick NCCL's LL
// protocol uses).  No fence / barrier / atomic is needed between producer and consumer.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ll_store(unsigned long long* p, float v, uint32_t tag) {
    unsigned long long w = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long ll_load(const unsigned long long* p) {
    unsigned long long w;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    return w;
}
// two adjacent words in one 16-byte request; each half is validated on its own tag
__device__ __forceinline__ void ll_load2(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ uint32_t ll_tag(unsigned long long w) { return (uint32_t)(w >> 32); }
__device__ __forceinline__ float ll_val(unsigned long long w) { return __uint_as_float((uint32_t)w); }

__device__ __forceinline__ int ld_volatile_i32(const int* p) {
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. SC'11); bit-identical to oracle/philox.py.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * 5.9604644775390625e-08f; }
__device__ __forceinline__ uint32_t word_of(uint4 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace wrnn
