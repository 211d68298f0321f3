// Which SM sub-partition does warp w of a CTA run on?  Prints (smid, %warpid) of every warp of
// the CTAs resident on SM 0 for a launch shaped like lockstep_kernel (128 threads, ~240
// registers, two CTAs per SM), and times a DMMA loop with the heavy warp of the second CTA
// rotated or not.   nvcc -arch=sm_100a -o tools/bin/warp_map tools/warp_map.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(128, 1) probe(int *out, int regs_dummy) {
    unsigned smid, warpid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
    __shared__ char pad[10000];
    pad[threadIdx.x] = 0;
    if ((threadIdx.x & 31) == 0) {
        int *o = out + (blockIdx.x * 4 + (threadIdx.x >> 5)) * 2;
        o[0] = smid;
        o[1] = warpid;
    }
    // keep the CTA resident for a while so that a second wave does not replace it
    long long t0 = clock64();
    while (clock64() - t0 < 2000000) { }
}
int main() {
    int n = 296, *d, *h = new int[n * 8];
    cudaMalloc(&d, n * 8 * sizeof(int));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 0);
    probe<<<n, 128>>>(d, 0);
    cudaMemcpy(h, d, n * 8 * sizeof(int), cudaMemcpyDeviceToHost);
    for (int b = 0; b < n; ++b)
        if (h[b * 8] < 2) {
            printf("block %3d on SM %d: warpid", b, h[b * 8]);
            for (int w = 0; w < 4; ++w) printf(" %2d", h[(b * 4 + w) * 2 + 1]);
            printf("\n");
        }
    return 0;
}
