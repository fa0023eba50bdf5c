// cond_tc.cu -- the conditioning front end's dense contractions on the tensor cores (tcgen05 + TMA).
//
// MelResNet (conv_in k=5, 20 1x1 convs, conv_out; BatchNorm folded) and the four table projections are all
// C[M x N] = act(A[M x K] W[N x K]^T + bias) (+ residual) with M = frames.  To keep fp32-grade accuracy on fp16
// tensor cores every operand is carried as a hi/lo pair of fp16 (x = hi + lo, 22 significant bits) and a product
// is three MMAs: Ahi*Whi + Ahi*Wlo + Alo*Whi (the lo*lo term is below fp32 rounding).  Weights are pre-scaled by
// 2^8 on the host so their lo parts stay normal; the epilogue multiplies by 2^-8.
//
// One CTA per 128 x 128 output tile: a TMA producer warp streams K in 64-column blocks (A hi/lo and W hi/lo tiles,
// 128B swizzle, 3-slot mbarrier ring), one thread issues 12 tcgen05.mma (M=128, N=128, fp32 accumulator in TMEM)
// per block, four epilogue warps read the accumulator with tcgen05.ld and write fp32 and/or the hi/lo pair that
// the next layer's TMA will read.  conv_in needs no im2col: tap j of the k=5 convolution is the same padded mel
// matrix loaded j rows lower (row_shift per K block).
#include "engine_internal.h"
#include "tc_common.cuh"

namespace wrnn {
namespace {
using namespace tc;

constexpr int kSlotsC = 3;
constexpr int kTile = 128 * 128;            // bytes of one [128 rows x 64 fp16] tile
constexpr int kSlotBytes = 4 * kTile;       // A hi, A lo, W hi, W lo
constexpr long long kDeadlineC = 2000000000LL;

struct CtlC {
    uint64_t full[kSlotsC];
    uint64_t empty[kSlotsC];
    uint64_t accfull;
    uint32_t tmem;
};

__device__ __forceinline__ bool wait_bar_c(uint64_t* bar, uint32_t parity) {
    long long t0 = 0;
    int spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (((++spins) & 1023) == 0) {
            if (t0 == 0) t0 = clock64();
            if (clock64() - t0 > kDeadlineC) return false;
        }
    }
    return true;
}

__global__ void __launch_bounds__(192, 1)
gemm_tc_split_kernel(const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo,
                     const __grid_constant__ CUtensorMap tmWhi, const __grid_constant__ CUtensorMap tmWlo, GemmTcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    CtlC* ctl = reinterpret_cast<CtlC*>(smem + kSlotsC * kSlotBytes);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.x * 128, m0 = blockIdx.y * 128;

    if (tid == 0) {
        for (int i = 0; i < kSlotsC; ++i) { mbar_init(&ctl->full[i], 1); mbar_init(&ctl->empty[i], 1); }
        mbar_init(&ctl->accfull, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&ctl->tmem, 128);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = ctl->tmem;

    if (warp == 4) {
        if (lane == 0) {
            tma_prefetch_desc(&tmAhi); tma_prefetch_desc(&tmAlo); tma_prefetch_desc(&tmWhi); tma_prefetch_desc(&tmWlo);
            for (int kb = 0; kb < a.nkb; ++kb) {
                const int slot = kb % kSlotsC, round = kb / kSlotsC;
                if (round > 0 && !wait_bar_c(&ctl->empty[slot], (round - 1) & 1)) { atomicExch(a.status, 2); break; }
                uint8_t* s = smem + slot * kSlotBytes;
                const int tap = kb / a.kb_per_tap, acol = (kb % a.kb_per_tap) * 64;
                mbar_arrive_expect_tx(&ctl->full[slot], kSlotBytes);
                tma_load_2d(s + 0 * kTile, &tmAhi, &ctl->full[slot], acol, m0 + tap * a.row_shift);
                tma_load_2d(s + 1 * kTile, &tmAlo, &ctl->full[slot], acol, m0 + tap * a.row_shift);
                tma_load_2d(s + 2 * kTile, &tmWhi, &ctl->full[slot], kb * 64, n0);
                tma_load_2d(s + 3 * kTile, &tmWlo, &ctl->full[slot], kb * 64, n0);
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16(128, 128);
            for (int kb = 0; kb < a.nkb; ++kb) {
                const int slot = kb % kSlotsC, round = kb / kSlotsC;
                if (!wait_bar_c(&ctl->full[slot], round & 1)) { atomicExch(a.status, 3); break; }
                tcgen05_fence_after();
                const uint32_t s = smem_u32(smem + slot * kSlotBytes);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint64_t ahi = umma_desc_sw128(s + 0 * kTile + j * 32), alo = umma_desc_sw128(s + 1 * kTile + j * 32);
                    const uint64_t whi = umma_desc_sw128(s + 2 * kTile + j * 32), wlo = umma_desc_sw128(s + 3 * kTile + j * 32);
                    umma_f16(tmem, ahi, whi, idesc, (kb | j) ? 1u : 0u);
                    umma_f16(tmem, ahi, wlo, idesc, 1u);
                    umma_f16(tmem, alo, whi, idesc, 1u);
                }
                umma_commit(&ctl->empty[slot]);
            }
            umma_commit(&ctl->accfull);
        }
    } else {
        const bool ok = wait_bar_c(&ctl->accfull, 0);
        tcgen05_fence_after();
        if (!ok) { if (lane == 0) atomicExch(a.status, 4); }
        else {
            const int m = m0 + warp * 32 + lane;
            const bool live = m < a.M;
            const float mask = (a.rowmask && live) ? a.rowmask[m] : 1.0f;
            for (int c0 = 0; c0 < 128; c0 += 8) {
                float v[8];
                tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
                tmem_ld_wait();
                if (live) {
                    const int n = n0 + c0;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float x = v[i] * a.scale + (a.bias ? __ldg(a.bias + n + i) : 0.f);
                        if (a.relu) x = fmaxf(x, 0.f);
                        if (a.R) x += a.R[(size_t)m * a.N + n + i];
                        v[i] = x * mask;
                    }
                    if (a.C) {
                        float4* dst = reinterpret_cast<float4*>(a.C + (size_t)m * a.N + n);
                        dst[0] = make_float4(v[0], v[1], v[2], v[3]);
                        dst[1] = make_float4(v[4], v[5], v[6], v[7]);
                    }
                    if (a.Chi) {
                        __half hi[8], lo[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) { hi[i] = __float2half_rn(v[i]); lo[i] = __float2half_rn(v[i] - __half2float(hi[i])); }
                        *reinterpret_cast<uint4*>(a.Chi + (size_t)m * a.N + n) = *reinterpret_cast<uint4*>(hi);
                        *reinterpret_cast<uint4*>(a.Clo + (size_t)m * a.N + n) = *reinterpret_cast<uint4*>(lo);
                    }
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

// padded mel, time-major, channels padded 80 -> 128, as hi/lo fp16:  MP[row][c] = melpad[c][row - r0]
__global__ void mel_split_kernel(const float* __restrict__ mel, const UttDesc* __restrict__ utts, int n_utts, int rows,
                                 __half* __restrict__ hi, __half* __restrict__ lo, float* __restrict__ rowmask) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)rows * 128; i += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(i >> 7), c = (int)(i & 127);
        int lo_u = 0, hi_u = n_utts - 1;
        while (lo_u < hi_u) { const int mid = (lo_u + hi_u + 1) >> 1; if (utts[mid].tq_row0 <= row) lo_u = mid; else hi_u = mid - 1; }
        const UttDesc u = utts[lo_u];
        const int t = row - u.tq_row0 - kPad;
        float v = 0.f;
        if (c < kFeat && t >= 0 && t < u.T) v = mel[u.mel_off + (long long)c * u.T + t];
        const __half h = __float2half_rn(v);
        hi[i] = h;
        lo[i] = __float2half_rn(v - __half2float(h));
        if (c == 0) rowmask[row] = (row - u.ta_row0 < u.T) ? 1.0f : 0.0f;     // frames are real, the 4 trailing rows are not
    }
}
}  // namespace

cudaError_t launch_mel_split(const float* mel, const UttDesc* utts, int n_utts, int rows, __half* hi, __half* lo, float* rowmask,
                             cudaStream_t stream) {
    int blocks = (int)(((long long)rows * 128 + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    mel_split_kernel<<<blocks, 256, 0, stream>>>(mel, utts, n_utts, rows, hi, lo, rowmask);
    return cudaGetLastError();
}

// A: [Arows][Acols] hi/lo fp16 row-major; W: [N][K] hi/lo fp16 row-major (K = nkb*64); see GemmTcArgs.
cudaError_t launch_gemm_tc_split(const __half* Ahi, const __half* Alo, int Arows, int Acols, const __half* Whi, const __half* Wlo,
                                 const GemmTcArgs& args, cudaStream_t stream) {
    if (args.N % 128 != 0 || args.nkb < 1) return cudaErrorInvalidValue;
    alignas(64) CUtensorMap maps[4];
    cudaError_t e;
    if ((e = make_tmap_f16_2d(&maps[0], Ahi, Arows, Acols, 128, 64)) != cudaSuccess) return e;
    if ((e = make_tmap_f16_2d(&maps[1], Alo, Arows, Acols, 128, 64)) != cudaSuccess) return e;
    if ((e = make_tmap_f16_2d(&maps[2], Whi, args.N, (uint64_t)args.nkb * 64, 128, 64)) != cudaSuccess) return e;
    if ((e = make_tmap_f16_2d(&maps[3], Wlo, args.N, (uint64_t)args.nkb * 64, 128, 64)) != cudaSuccess) return e;
    const int smem = kSlotsC * kSlotBytes + 256 + 1024;
    if ((e = cudaFuncSetAttribute(gemm_tc_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
    dim3 grid(args.N / 128, (args.M + 127) / 128);
    gemm_tc_split_kernel<<<grid, 192, smem, stream>>>(maps[0], maps[1], maps[2], maps[3], args);
    return cudaGetLastError();
}

}  // namespace wrnn
