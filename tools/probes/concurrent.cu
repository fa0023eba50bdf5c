// Does a small kernel on a second stream run while a 130-CTA spinning kernel (1 CTA/SM, 200 KB smem) is resident?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void spinner(volatile int* flag, long long* out, long long limit) {
    extern __shared__ char sm[];
    sm[threadIdx.x] = 0;
    long long t0 = clock64(), t = t0;
    while (*flag == 0 && (t = clock64()) - t0 < limit) {}
    if (threadIdx.x == 0) out[blockIdx.x] = (*flag != 0) ? (t - t0) : -1;
}
__global__ void setter(int* flag) { if (threadIdx.x == 0 && blockIdx.x == 0) { atomicExch(flag, 1); } }
int main(int argc, char** argv) {
    const int coop = argc > 1 ? atoi(argv[1]) : 0, nblk = argc > 2 ? atoi(argv[2]) : 130, setter_blocks = argc > 3 ? atoi(argv[3]) : 1;
    int* flag; long long* out;
    cudaMalloc(&flag, 4); cudaMalloc(&out, 8 * 256);
    cudaMemset(flag, 0, 4);
    cudaStream_t s1, s2;
    int lo, hi; cudaDeviceGetStreamPriorityRange(&lo, &hi);
    cudaStreamCreateWithPriority(&s1, cudaStreamNonBlocking, hi);
    cudaStreamCreateWithPriority(&s2, cudaStreamNonBlocking, lo);
    cudaFuncSetAttribute(spinner, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    long long limit = 200000000LL;   // ~0.1 s
    volatile int* f = flag;
    void* args[] = {(void*)&f, (void*)&out, (void*)&limit};
    cudaError_t e;
    if (coop) e = cudaLaunchCooperativeKernel((const void*)spinner, dim3(nblk), dim3(608), args, 200 * 1024, s1);
    else { spinner<<<nblk, 608, 200 * 1024, s1>>>(flag, out, limit); e = cudaGetLastError(); }
    printf("launch spinner: %s\n", cudaGetErrorString(e));
    setter<<<setter_blocks, 256, 0, s2>>>(flag);
    printf("launch setter: %s\n", cudaGetErrorString(cudaGetLastError()));
    cudaDeviceSynchronize();
    long long h[256]; cudaMemcpy(h, out, 8 * nblk, cudaMemcpyDeviceToHost);
    int seen = 0; long long mx = 0;
    for (int i = 0; i < nblk; ++i) { if (h[i] >= 0) { ++seen; if (h[i] > mx) mx = h[i]; } }
    printf("coop=%d blocks=%d setter_blocks=%d: %d CTAs saw the flag while spinning (max wait %lld clk)\n", coop, nblk, setter_blocks, seen, mx);
    return 0;
}
