// ubench.cu — B200 micro-benchmarks that ground the kernel design:
// DFMA latency/throughput, DMMA m8n8k4 latency/throughput, broadcast LDS.128 cost,
// named-barrier latency for 2/4 warps.  Build: nvcc -arch=sm_100a -O3 tools/ubench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__global__ void dfma_lat(double *out, long long *cyc, int iters) {
    double a = out[0], b = 1.0000001;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 32; ++k) a = fma(a, b, 0.5);
    }
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int CH>
__global__ void dfma_tput(double *out, long long *cyc, int iters) {
    double a[CH];
    for (int k = 0; k < CH; ++k) a[k] = out[k];
    double b = 1.0000001;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int k = 0; k < CH; ++k) a[k] = fma(a[k], b, 0.5);
    }
    long long t1 = clock64();
    double s = 0;
    for (int k = 0; k < CH; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int CH>
__global__ void dmma_tput(double *out, long long *cyc, int iters) {
    double d0[CH], d1[CH];
    for (int k = 0; k < CH; ++k) { d0[k] = out[k]; d1[k] = out[k + 1]; }
    double a = 1.0000001, b = 0.999999;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int k = 0; k < CH; ++k) dmma(d0[k], d1[k], a, b);
    }
    long long t1 = clock64();
    double s = 0;
    for (int k = 0; k < CH; ++k) s += d0[k] + d1[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// broadcast LDS.128 x14 + dependent use, like the forward kernel's exchange
__global__ void lds_bcast(double *out, long long *cyc, int iters) {
    __shared__ __align__(16) double xs[64];
    xs[threadIdx.x] = out[threadIdx.x];
    xs[threadIdx.x + 32] = out[threadIdx.x + 32];
    __syncwarp();
    double acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        const double2 *x2 = reinterpret_cast<const double2 *>(xs + (i & 1) * 32);
#pragma unroll
        for (int k = 0; k < 14; ++k) { double2 p = x2[k]; acc += p.x + p.y; }
        xs[((i + 1) & 1) * 32 + threadIdx.x] = acc;     // STS -> next iteration's LDS
        __syncwarp();
    }
    long long t1 = clock64();
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// STS -> syncwarp -> single LDS round trip (dependent)
__global__ void sts_lds_lat(double *out, long long *cyc, int iters) {
    __shared__ double xs[64];
    double v = out[threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        xs[(i & 1) * 32 + threadIdx.x] = v;
        __syncwarp();
        v = xs[(i & 1) * 32 + ((threadIdx.x + 1) & 31)];
    }
    long long t1 = clock64();
    out[threadIdx.x] = v;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int NW>
__global__ void bar_lat(double *out, long long *cyc, int iters) {
    __shared__ double xs[2][128];
    double v = out[threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        xs[i & 1][threadIdx.x] = v;
        asm volatile("bar.sync 1, %0;" ::"n"(NW * 32));
        v = xs[i & 1][(threadIdx.x + 32) % (NW * 32)];
    }
    long long t1 = clock64();
    out[threadIdx.x] = v;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void shfl_lat(double *out, long long *cyc, int iters) {
    double v = out[threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) v = __shfl_xor_sync(0xffffffffu, v, 1) + 1.0;
    long long t1 = clock64();
    out[threadIdx.x] = v;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void dsetp_sel_lat(double *out, long long *cyc, int iters) {
    double best = out[threadIdx.x], x = out[threadIdx.x + 1];
    int arg = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            double m = x + (double)k;   // independent of best
            if (m > best) { best = m; arg = i + k; }
            x = -x;
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = best + arg;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    double *d; long long *c, h;
    CK(cudaMalloc(&d, 1 << 22)); CK(cudaMemset(d, 0, 1 << 22)); CK(cudaMalloc(&c, 8));
    const int it = 20000;
    auto rep = [&](const char *name, double per) {
        CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost));
        printf("%-34s %8.2f cycles per op\n", name, (double)h / per);
    };
    dfma_lat<<<1, 32>>>(d, c, it); rep("DFMA dependent latency", it * 32.0);
    dfma_tput<4><<<1, 32>>>(d, c, it); rep("DFMA 1 warp, 4 chains", it * 32.0);
    dfma_tput<8><<<1, 32>>>(d, c, it); rep("DFMA 1 warp, 8 chains", it * 64.0);
    dfma_tput<8><<<1, 128>>>(d, c, it); rep("DFMA 4 warps(4 SMSP), 8 ch /warp", it * 64.0);
    dfma_tput<8><<<1, 256>>>(d, c, it); rep("DFMA 8 warps, 8 ch (per warp-op)", it * 64.0);
    dmma_tput<1><<<1, 32>>>(d, c, it); rep("DMMA m8n8k4 dependent latency", it * 8.0);
    dmma_tput<4><<<1, 32>>>(d, c, it); rep("DMMA 1 warp, 4 chains", it * 32.0);
    dmma_tput<8><<<1, 32>>>(d, c, it); rep("DMMA 1 warp, 8 chains", it * 64.0);
    dmma_tput<8><<<1, 128>>>(d, c, it); rep("DMMA 4 warps, 8 ch (per warp-op)", it * 64.0);
    dmma_tput<8><<<1, 256>>>(d, c, it); rep("DMMA 8 warps, 8 ch (per warp-op)", it * 64.0);
    lds_bcast<<<1, 32>>>(d, c, it); rep("14x LDS.128 bcast + STS round", it * 1.0);
    sts_lds_lat<<<1, 32>>>(d, c, it); rep("STS->syncwarp->LDS round trip", it * 1.0);
    bar_lat<2><<<1, 64>>>(d, c, it); rep("STS->bar.sync(2 warps)->LDS", it * 1.0);
    bar_lat<4><<<1, 128>>>(d, c, it); rep("STS->bar.sync(4 warps)->LDS", it * 1.0);
    shfl_lat<<<1, 32>>>(d, c, it); rep("SHFL(64-bit)+DADD dependent", it * 1.0);
    dsetp_sel_lat<<<1, 32>>>(d, c, it); rep("DADD+DSETP+select chain (per i)", it * 16.0);
    // whole-GPU DFMA throughput
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w : {4, 8, 16}) {
        dfma_tput<8><<<148 * 2, w * 16>>>(d, c, 2000);
        CK(cudaEventRecord(e0));
        dfma_tput<8><<<148 * 2, w * 16>>>(d, c, it);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * 148 * 2 * w * 16 * (double)it * 64;
        printf("DFMA whole GPU, %2d warps/SM: %.2f TFLOP/s\n", w, flops / ms / 1e9);
        dmma_tput<8><<<148 * 2, w * 16>>>(d, c, 2000);
        CK(cudaEventRecord(e0));
        dmma_tput<8><<<148 * 2, w * 16>>>(d, c, it);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        flops = 2.0 * 256 * 148 * 2 * (w / 2) * (double)it * 64;
        printf("DMMA whole GPU, %2d warps/SM: %.2f TFLOP/s\n", w, flops / ms / 1e9);
    }
    return 0;
}
