// bench_floor.cu -- measures the floor of one inter-SM exchange of the persistent loop, the quantity the
// north_star asks the per-step latency to be compared with.  Two protocols over the same 128-CTA
// cooperative grid, no arithmetic:
//   LL      : every CTA publishes 4 {value,tag} words, every CTA spins until all 512 words carry the tag
//   counter : plain stores + __threadfence + atomicAdd on one counter + spin + plain loads
#include <cooperative_groups.h>

#include "engine_internal.h"

namespace wrnn {
namespace {
constexpr long long kDeadline = 1500000000LL;

__global__ void __launch_bounds__(512, 1) floor_ll_kernel(unsigned long long* buf, int rounds, int* abort_flag) {
    const int tid = threadIdx.x, cta = blockIdx.x;
    float sink = 0.f;
    for (int r = 0; r < rounds; ++r) {
        unsigned long long* b = buf + (r & 1) * kRnn;
        const uint32_t tag = (uint32_t)r + 1u;
        if (tid < kUnitsF32) ll_store(b + cta * kUnitsF32 + tid, (float)r, tag);
        int failed = 0, spins = 0;
        long long t0 = 0;
        while (true) {
            unsigned long long w = ll_load(b + tid);
            if (ll_tag(w) == tag) { sink += ll_val(w); break; }
            if (((++spins) & 63) == 0) {
                if (t0 == 0) t0 = clock64();
                if (clock64() - t0 > kDeadline || ld_volatile_i32(abort_flag)) { failed = 1; break; }
            }
        }
        if (__syncthreads_or(failed)) { if (tid == 0) atomicExch(abort_flag, 1); return; }
    }
    if (sink == -1.f) buf[2 * kRnn] = 0;
}

__global__ void __launch_bounds__(512, 1) floor_counter_kernel(unsigned int* counter, float* data, int rounds,
                                                               int* abort_flag) {
    const int tid = threadIdx.x, cta = blockIdx.x;
    float sink = 0.f;
    for (int r = 0; r < rounds; ++r) {
        float* d = data + (r & 1) * kRnn;
        if (tid < kUnitsF32) d[cta * kUnitsF32 + tid] = (float)r;
        __syncthreads();
        int failed = 0;
        if (tid == 0) {
            __threadfence();
            atomicAdd(counter, 1u);
            const unsigned int want = (unsigned int)(r + 1) * gridDim.x;
            int spins = 0;
            long long t0 = 0;
            while (true) {
                unsigned int v;
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
                if (v >= want) break;
                if (((++spins) & 63) == 0) {
                    if (t0 == 0) t0 = clock64();
                    if (clock64() - t0 > kDeadline || ld_volatile_i32(abort_flag)) { failed = 1; break; }
                }
            }
        }
        if (__syncthreads_or(failed)) { if (tid == 0) atomicExch(abort_flag, 1); return; }
        sink += __ldcg(d + tid);
    }
    if (sink == -1.f) data[2 * kRnn] = 0.f;
}
// cluster-local exchange: every CTA stores 32 floats into each of its 16 peers' shared memory, then cluster barrier
__global__ void __launch_bounds__(256, 1) floor_cluster_kernel(int rounds, float* sink_out) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ float buf[2][kRnn];
    const int tid = threadIdx.x, rank = (int)cluster.block_rank(), n = (int)cluster.num_blocks();
    float sink = 0.f;
    cluster.sync();
    for (int r = 0; r < rounds; ++r) {
        float* b = buf[r & 1];
        if (tid < 32)
            for (int q = 0; q < n; ++q) cluster.map_shared_rank(b, q)[rank * 32 + tid] = (float)r;
        cluster.sync();
        sink += b[tid] + b[tid + 256];
    }
    cluster.sync();
    if (sink == -1.f) *sink_out = sink;
}
}  // namespace

cudaError_t launch_floor_cluster(int cluster_size, int rounds, float* sink, cudaStream_t stream) {
    cudaError_t e = cudaSuccess;
    if (cluster_size > 8) e = cudaFuncSetAttribute(floor_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cluster_size);
    cfg.blockDim = dim3(256);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_size; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, floor_cluster_kernel, rounds, sink);
}

cudaError_t launch_floor_ll(unsigned long long* buf, int rounds, int* abort_flag, cudaStream_t stream) {
    void* args[] = {&buf, &rounds, &abort_flag};
    return cudaLaunchCooperativeKernel((const void*)floor_ll_kernel, dim3(kCtasF32), dim3(512), args, 0, stream);
}
cudaError_t launch_floor_counter(unsigned int* counter, float* data, int rounds, int* abort_flag, cudaStream_t stream) {
    void* args[] = {&counter, &data, &rounds, &abort_flag};
    return cudaLaunchCooperativeKernel((const void*)floor_counter_kernel, dim3(kCtasF32), dim3(512), args, 0, stream);
}
}  // namespace wrnn
