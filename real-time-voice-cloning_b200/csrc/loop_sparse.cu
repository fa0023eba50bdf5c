// loop_sparse.cu -- the sample loop for PRUNED checkpoints (reference: vocoder/pruner.py 1x4 magnitude pruning;
// CPU counterpart vocoder/libwavernn CompMatrix, wavernn.cpp:162-184): block-sparse fp32 variant.
//
// At ~90 % sparsity the loop weights (values + indices) are ~1.5 MB: they fit in the shared memory of ONE thread-block
// cluster.  So the design flips from "all SMs cooperate through L2" to "one 16-CTA cluster owns a few folds
// completely": every exchange of a step (h1, h2, f1, f2, logits) is a DSMEM store into the peers' activation
// buffers followed by a hardware cluster barrier (~0.2 us) instead of a trip through L2 (~1-1.6 us), and clusters
// never talk to each other, so folds are simply partitioned across clusters (no co-residency requirement).
//  * CTA c of a cluster owns hidden units [U*c, U*c+U) of every layer (U = 512 / cluster size); its rows are stored
//    CSR-like: rowptr, one byte per kept 1x4 group (its group column) and the 4 weights.  Four lanes share a row.
//  * sampling is done redundantly by every CTA from identical logits, so the sample never needs to be exchanged.
//  * same algebra, tables, noise and post chain as the dense loops; fp32 throughout (parity mode).
#include <cooperative_groups.h>

#include "engine_internal.h"

namespace cg = cooperative_groups;

namespace wrnn {
namespace {

constexpr int NTS = 256;          // threads per CTA
constexpr int kLanesPerRow = 4;

struct SpView {
    const int* rowptr;            // [rows + 1]
    const unsigned char* col;     // [nnz] group column (0..127)
    const float4* w;              // [nnz]
};

// out[b][r] = sum over kept groups of row r:  w . act[b][4*col .. 4*col+3]      (b < nb <= 8)
__device__ __forceinline__ void spmv(const SpView& m, int rows, const float* __restrict__ act, int nb, float* __restrict__ out,
                                     int ldo) {
    const int sub = threadIdx.x & (kLanesPerRow - 1), slot = threadIdx.x / kLanesPerRow;
    for (int r = slot; r < rows; r += NTS / kLanesPerRow) {
        float acc[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[b] = 0.f;
        const int g1 = m.rowptr[r + 1];
        for (int g = m.rowptr[r] + sub; g < g1; g += kLanesPerRow) {
            const float4 w = m.w[g];
            const float* x = act + 4 * (int)m.col[g];
#pragma unroll
            for (int b = 0; b < 8; ++b)
                if (b < nb) {
                    const float4 a = *reinterpret_cast<const float4*>(x + b * kRnn);
                    acc[b] = fmaf(w.x, a.x, fmaf(w.y, a.y, fmaf(w.z, a.z, fmaf(w.w, a.w, acc[b]))));
                }
        }
#pragma unroll
        for (int b = 0; b < 8; ++b)
            if (b < nb) {
                float v = acc[b];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                if (sub == 0) out[b * ldo + r] = v;
            }
    }
}

template <int CL>
__global__ void __launch_bounds__(NTS, 1) wrnn_loop_sparse_kernel(SparseParams p) {
    constexpr int U = kRnn / CL;                 // hidden units per CTA
    constexpr int RB = 7 * U, RC = 4 * U;        // stage rows: [W_hh1 | W_ih2a | W_fc1a], [W_hh2 | W_fc1a]
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int crank = (int)cluster.block_rank(), cl = blockIdx.x / CL;
    const int b0 = cl * p.Bc, nb = min(p.Bc, p.B - b0);        // my cluster's folds (nb >= 1 by construction)
    const int j0 = crank * U;
    const int C = p.C, CRs = p.CRs;

    // ---- carve shared memory: weight image first (copied verbatim), then buffers ----------------------------------
    const int* hdr = reinterpret_cast<const int*>(p.wimg + (size_t)crank * p.img_stride);
    {
        const uint4* src = reinterpret_cast<const uint4*>(hdr);
        uint4* dst = reinterpret_cast<uint4*>(smem_raw);
        for (int i = tid; i < p.img_stride / 16; i += NTS) dst[i] = src[i];
    }
    __syncthreads();
    const int* sh = reinterpret_cast<const int*>(smem_raw);
    SpView st[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        st[s].rowptr = reinterpret_cast<const int*>(smem_raw + sh[4 + 3 * s + 0]);
        st[s].col = smem_raw + sh[4 + 3 * s + 1];
        st[s].w = reinterpret_cast<const float4*>(smem_raw + sh[4 + 3 * s + 2]);
    }
    float* fbase = reinterpret_cast<float*>(smem_raw + p.img_stride);
    float* buf0 = fbase;                          fbase += p.Bc * kRnn;       // exchange buffers (peers write into them)
    float* buf1 = fbase;                          fbase += p.Bc * kRnn;
    float* lg = fbase;                            fbase += p.Bc * p.Cpad;
    float* tmp = fbase;                           fbase += p.Bc * RB;
    float4* c1 = reinterpret_cast<float4*>(fbase); fbase += p.Bc * U * 4;
    float4* c2 = reinterpret_cast<float4*>(fbase); fbase += p.Bc * U * 4;
    float* h1 = fbase;                            fbase += p.Bc * U;
    float* h2 = fbase;                            fbase += p.Bc * U;
    float* p3 = fbase;                            fbase += p.Bc * U;
    float* gh1 = fbase;                           fbase += p.Bc * 3 * U;
    float* gh2 = fbase;                           fbase += p.Bc * 3 * U;
    float* xs = fbase;                            fbase += 8;
    float* v1 = fbase;                            fbase += 3 * U;
    float* v2 = fbase;                            fbase += 3 * U;
    float* v3 = fbase;                            fbase += U;
    float* bh1 = fbase;                           fbase += U;
    float* bh2 = fbase;                           fbase += U;
    float* b3 = fbase;                            fbase += ((CRs + 3) & ~3);
    float* coef = fbase;

    for (int i = tid; i < 3 * U; i += NTS) { v1[i] = p.v1[(i / U) * kRnn + j0 + i % U]; v2[i] = p.v2[(i / U) * kRnn + j0 + i % U]; }
    for (int i = tid; i < U; i += NTS) { v3[i] = p.v3[j0 + i]; bh1[i] = p.bhn1[j0 + i]; bh2[i] = p.bhn2[j0 + i]; }
    for (int i = tid; i < CRs; i += NTS) b3[i] = (crank * CRs + i < C) ? p.bfc3[crank * CRs + i] : 0.f;
    for (int i = tid; i < kHop * kTaps; i += NTS) coef[i] = p.coef[i];
    for (int i = tid; i < p.Bc * U; i += NTS) { h1[i] = 0.f; h2[i] = 0.f; p3[i] = 0.f; }
    for (int i = tid; i < p.Bc * 3 * U; i += NTS) { gh1[i] = 0.f; gh2[i] = 0.f; }
    if (tid < 8) xs[tid] = 0.f;
    const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32));
    cluster.sync();

    // publish v for (fold b, my unit u) into every CTA's `dst` buffer at column j0+u
    auto publish = [&](float* dst, int b, int col, float v, int ld) {
#pragma unroll
        for (int r = 0; r < CL; ++r) cluster.map_shared_rank(dst, r)[b * ld + col] = v;
    };

    for (int t = 0; t < p.S; ++t) {
        // ---- conditioning of this step for my units -----------------------------------------------------------------
        for (int e = tid; e < nb * U; e += NTS) {
            const int b = e / U, j = j0 + e % U;
            const FoldDesc fd = p.folds[b0 + b];
            const int n = fd.n0 + t;
            const bool valid = n < fd.N;
            const int q0 = valid ? n / kHop : 0;
            const size_t ra = (size_t)(fd.ta_row0 + (valid ? q0 : fd.T)) * kRnn + j;
            float4 a1 = __ldg(p.TA1 + ra), a2 = __ldg(p.TA2 + ra);
            if (valid) {
                const float* cf = coef + (n - q0 * kHop) * kTaps;
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
            c1[e] = a1;
            c2[e] = a2;
        }
        __syncthreads();
        // ---- A: GRU1 for my units, h1 -> everyone's buf0 ---------------------------------------------------------------
        for (int e = tid; e < nb * U; e += NTS) {
            const int b = e / U, u = e % U;
            const float x = xs[b];
            const float4 c = c1[e];
            const float* gh = gh1 + b * 3 * U;
            const float r = sigmoid_acc(fmaf(v1[u], x, c.x) + gh[u]);
            const float z = sigmoid_acc(fmaf(v1[U + u], x, c.y) + gh[U + u]);
            const float nn = tanhf(fmaf(v1[2 * U + u], x, c.z) + r * (gh[2 * U + u] + bh1[u]));
            const float h = (1.0f - z) * nn + z * h1[e];
            h1[e] = h;
            publish(buf0, b, j0 + u, h, kRnn);
        }
        cluster.sync();
        // ---- B: [W_hh1 | W_ih2a | W_fc1a] h1 ; GRU2 ; h2 -> buf1 --------------------------------------------------------
        spmv(st[0], RB, buf0, nb, tmp, RB);
        __syncthreads();
        for (int e = tid; e < nb * U; e += NTS) {
            const int b = e / U, u = e % U;
            const float* r_ = tmp + b * RB;
            const float x = xs[b];
            const float4 cc = c2[e];
            const float* gh = gh2 + b * 3 * U;
            const float r = sigmoid_acc(r_[3 * U + u] + fmaf(v2[u], x, cc.x) + gh[u]);
            const float z = sigmoid_acc(r_[4 * U + u] + fmaf(v2[U + u], x, cc.y) + gh[U + u]);
            const float nn = tanhf(r_[5 * U + u] + fmaf(v2[2 * U + u], x, cc.z) + r * (gh[2 * U + u] + bh2[u]));
            const float h = (1.0f - z) * nn + z * h2[e];
            h2[e] = h;
            p3[e] = r_[6 * U + u];
            publish(buf1, b, j0 + u, h, kRnn);
        }
        for (int e = tid; e < nb * 3 * U; e += NTS) gh1[e] = tmp[(e / (3 * U)) * RB + e % (3 * U)];
        cluster.sync();
        // ---- C: [W_hh2 | W_fc1a] h2 ; f1 -> buf0 -------------------------------------------------------------------------
        spmv(st[1], RC, buf1, nb, tmp, RB);
        __syncthreads();
        for (int e = tid; e < nb * U; e += NTS) {
            const int b = e / U, u = e % U;
            const float v = p3[e] + tmp[b * RB + 3 * U + u] + fmaf(v3[u], xs[b], c1[e].w);
            publish(buf0, b, j0 + u, fmaxf(v, 0.f), kRnn);
        }
        for (int e = tid; e < nb * 3 * U; e += NTS) gh2[e] = tmp[(e / (3 * U)) * RB + e % (3 * U)];
        cluster.sync();
        // ---- D: fc2 ; f2 -> buf1 ---------------------------------------------------------------------------------------
        spmv(st[2], U, buf0, nb, tmp, RB);
        __syncthreads();
        for (int e = tid; e < nb * U; e += NTS) {
            const int b = e / U, u = e % U;
            publish(buf1, b, j0 + u, fmaxf(tmp[b * RB + u] + c2[e].w, 0.f), kRnn);
        }
        cluster.sync();
        // ---- E: my classes of fc3 -> everyone's logits -----------------------------------------------------------------
        spmv(st[3], CRs, buf1, nb, tmp, RB);
        __syncthreads();
        for (int e = tid; e < nb * CRs; e += NTS) {
            const int b = e / CRs, r = e % CRs, cls = crank * CRs + r;
            if (cls < C) {
                const float v = tmp[b * RB + r] + b3[r];
                publish(lg, b, cls, v, p.Cpad);
                if (p.logits_out) p.logits_out[((size_t)(b0 + b) * p.S + t) * C + cls] = v;
            }
        }
        cluster.sync();
        // ---- F: every CTA draws the same sample from the same logits (nothing to exchange) ---------------------------------
        for (int b = warp; b < nb; b += NTS / 32) {
            const FoldDesc fd = p.folds[b0 + b];
            const float* row = lg + b * p.Cpad;
            float xsv;
            if (p.mode == 1) {
                float score = -INFINITY;
                const float lv = (lane < 30) ? row[lane] : 0.f;
                if (lane < 10) {
                    const uint4 r = philox4x32_10(make_uint4((uint32_t)t, (uint32_t)fd.fold, (uint32_t)fd.utt, (uint32_t)(lane >> 2)), key);
                    const float um = 1e-5f + u01(word_of(r, lane & 3)) * (1.0f - 2e-5f);
                    score = lv - logf(-logf(um));
                }
                int idx = lane;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float os = __shfl_xor_sync(0xffffffffu, score, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
                    if (os > score || (os == score && oi < idx)) { score = os; idx = oi; }
                }
                const float mean = __shfl_sync(0xffffffffu, lv, 10 + idx);
                const float lsc = fmaxf(__shfl_sync(0xffffffffu, lv, 20 + idx), -32.23619130191664f);
                const uint4 r2 = philox4x32_10(make_uint4((uint32_t)t, (uint32_t)fd.fold, (uint32_t)fd.utt, 2u), key);
                const float ul = 1e-5f + u01(r2.z) * (1.0f - 2e-5f);
                xsv = fminf(fmaxf(mean + expf(lsc) * (logf(ul) - logf(1.0f - ul)), -1.0f), 1.0f);
            } else {
                const uint4 r = philox4x32_10(make_uint4((uint32_t)t, (uint32_t)fd.fold, (uint32_t)fd.utt, 0u), key);
                const float u = u01(r.x);
                const int seg = C / 32;                                  // contiguous classes per lane (8, 16 or 32)
                float m = -INFINITY;
                for (int i = 0; i < seg; ++i) m = fmaxf(m, row[lane * seg + i]);
                m = warp_max(m);
                float s = 0.f;
                for (int i = 0; i < seg; ++i) s += expf(row[lane * seg + i] - m);
                const float total = warp_sum(s);
                float ls = 0.f;
                for (int i = 0; i < seg; ++i) ls += expf(row[lane * seg + i] - m) / total;
                float incl = ls;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float tt = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += tt;
                }
                const unsigned hit = __ballot_sync(0xffffffffu, incl >= u);
                int k = C - 1;
                if (hit) {
                    const int L = __ffs(hit) - 1;
                    int kk = seg - 1;
                    if (lane == L) {
                        float c = incl - ls;
                        for (int i = 0; i < seg; ++i) {
                            c += expf(row[lane * seg + i] - m) / total;
                            if (c >= u) { kk = i; break; }
                        }
                        kk += L * seg;
                    }
                    k = __shfl_sync(0xffffffffu, kk, L);
                }
                xsv = 2.0f * (float)k / ((float)C - 1.0f) - 1.0f;
            }
            if (lane == 0) {
                if (crank == 0) p.samples[(size_t)(b0 + b) * p.S + t] = xsv;
                xs[b] = p.forced ? p.forced[(size_t)(b0 + b) * p.S + t] : xsv;
            }
        }
        __syncthreads();
        if (blockIdx.x == 0 && tid == 0 && (t % 100) == 0 && p.progress) {
            *reinterpret_cast<volatile int*>(p.progress) = t;
            __threadfence_system();
        }
    }
    cluster.sync();      // no CTA may exit while peers can still write into its shared memory
}

template <int CL>
size_t sparse_buffers_bytes(int Bc, int Cpad, int CRs) {
    constexpr int U = kRnn / CL;
    size_t f = (size_t)2 * Bc * kRnn + (size_t)Bc * Cpad + (size_t)Bc * 7 * U + (size_t)2 * Bc * U * 4 + (size_t)3 * Bc * U +
               (size_t)2 * Bc * 3 * U + 8 + 6 * U + 3 * U + ((CRs + 3) & ~3) + kHop * kTaps;
    return f * sizeof(float);
}

template <int CL>
cudaError_t launch_sparse_t(const SparseParams& p, int n_clusters, cudaStream_t stream) {
    const size_t smem = (size_t)p.img_stride + sparse_buffers_bytes<CL>(p.Bc, p.Cpad, p.CRs);
    cudaError_t e = cudaFuncSetAttribute(wrnn_loop_sparse_kernel<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (CL > 8) {
        e = cudaFuncSetAttribute(wrnn_loop_sparse_kernel<CL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_clusters * CL);
    cfg.blockDim = dim3(NTS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, wrnn_loop_sparse_kernel<CL>, p);
}
}  // namespace

size_t loop_sparse_smem_bytes(int cluster, int img_stride, int Bc, int Cpad, int CRs) {
    return (size_t)img_stride + (cluster == 16 ? sparse_buffers_bytes<16>(Bc, Cpad, CRs) : sparse_buffers_bytes<8>(Bc, Cpad, CRs));
}

cudaError_t launch_loop_sparse(const SparseParams& p, int cluster, int n_clusters, cudaStream_t stream) {
    if (cluster == 16) return launch_sparse_t<16>(p, n_clusters, stream);
    if (cluster == 8) return launch_sparse_t<8>(p, n_clusters, stream);
    return cudaErrorInvalidValue;
}

}  // namespace wrnn
