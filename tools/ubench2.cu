// ubench2.cu — shared-memory broadcast delivery rate and compare/select latencies on B200
// (grounds the exchange design of the recursion kernels).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <int MODE>   // 0: LDS.128 broadcast, 1: LDS.64 broadcast, 2: LDS.32 broadcast, 3: LDS.64 per-lane distinct, 4: LDS.128 4 distinct addrs
__global__ void lds_tput(double *out, long long *cyc, int iters) {
    __shared__ __align__(16) double xs[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) xs[i] = out[i];
    __syncthreads();
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    const int lane = threadIdx.x & 31;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        const int base = (i & 7) * 64;
#pragma unroll
        for (int k = 0; k < 14; ++k) {
            if (MODE == 0) { double2 p = *reinterpret_cast<const double2 *>(xs + base + 2 * k); a0 += p.x; a1 += p.y; }
            if (MODE == 1) { double p = xs[base + 2 * k]; double q = xs[base + 2 * k + 1]; a0 += p; a1 += q; }
            if (MODE == 2) { float p = reinterpret_cast<const float *>(xs)[2 * base + 4 * k]; float q = reinterpret_cast<const float *>(xs)[2 * base + 4 * k + 1];
                             float r = reinterpret_cast<const float *>(xs)[2 * base + 4 * k + 2]; float s = reinterpret_cast<const float *>(xs)[2 * base + 4 * k + 3]; a0 += p; a1 += q; a2 += r; a3 += s; }
            if (MODE == 3) { double p = xs[base + 32 * (k & 1) + lane]; a0 += p; }
            if (MODE == 4) { double2 p = *reinterpret_cast<const double2 *>(xs + base + 8 * (lane & 3) + 2 * (k & 3)); a0 += p.x; a1 += p.y; }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void shfl_tput(double *out, long long *cyc, int iters) {
    double x = out[threadIdx.x], a0 = 0, a1 = 0;
    const int lane = threadIdx.x & 31;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 14; ++k) {
            a0 += __shfl_sync(0xffffffffu, x, (lane + 2 * k) & 31);
            a1 += __shfl_sync(0xffffffffu, x, (lane + 2 * k + 1) & 31);
        }
        x += 1.0;
    }
    long long t1 = clock64();
    out[threadIdx.x] = a0 + a1;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// dependent chains
__global__ void lat_chain(double *out, long long *cyc, int iters, int mode) {
    double a = out[threadIdx.x], b = out[threadIdx.x + 32], c = out[threadIdx.x + 64];
    long long t0 = clock64();
    if (mode == 0) for (int i = 0; i < iters; ++i) {           // DADD chain
#pragma unroll
        for (int k = 0; k < 16; ++k) a = __dadd_rn(a, b);
    }
    if (mode == 1) for (int i = 0; i < iters; ++i) {           // DSETP -> 2 SEL chain (max)
#pragma unroll
        for (int k = 0; k < 16; ++k) { a = (b > a) ? b : a; b = -b; }
    }
    if (mode == 2) for (int i = 0; i < iters; ++i) {           // DMUL chain
#pragma unroll
        for (int k = 0; k < 16; ++k) a = __dmul_rn(a, b);
    }
    if (mode == 3) for (int i = 0; i < iters; ++i) {           // 64-bit integer compare/select chain
        long long x = __double_as_longlong(a), y = __double_as_longlong(b);
#pragma unroll
        for (int k = 0; k < 16; ++k) { x = (y < x) ? y : x; y ^= 0x5555; }
        a = __longlong_as_double(x); b = __longlong_as_double(y);
    }
    long long t1 = clock64();
    out[threadIdx.x] = a + b + c;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// STS -> LDS.128 broadcast of 28 values -> 28 DFMA (4 acc) -> 2 DADD -> DMUL : the forward column
__global__ void fwd_col(double *out, long long *cyc, int iters) {
    __shared__ __align__(16) double xs[2][32];
    double col[28];
    for (int k = 0; k < 28; ++k) col[k] = out[threadIdx.x * 28 + k] * 1e-3 + 0.03;
    double x = out[threadIdx.x] + 1.0, e = 0.999;
    int buf = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        xs[buf][threadIdx.x] = x;
        __syncwarp();
        const double2 *x2 = reinterpret_cast<const double2 *>(xs[buf]);
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
        for (int k = 0; k < 28; k += 4) {
            double2 p = x2[k / 2], q = x2[k / 2 + 1];
            a0 = fma(p.x, col[k], a0); a1 = fma(p.y, col[k + 1], a1); a2 = fma(q.x, col[k + 2], a2); a3 = fma(q.y, col[k + 3], a3);
        }
        x = ((a0 + a1) + (a2 + a3)) * e;
        buf ^= 1;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    double *d; long long *c, h;
    CK(cudaMalloc(&d, 1 << 22)); CK(cudaMemset(d, 0, 1 << 22)); CK(cudaMalloc(&c, 8));
    const int it = 20000;
    auto rep = [&](const char *name, double per) {
        CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost));
        printf("%-52s %8.2f cycles per op\n", name, (double)h / per);
    };
    lds_tput<0><<<1, 32>>>(d, c, it); rep("LDS.128 broadcast, 1 warp (per instr)", it * 14.0);
    lds_tput<1><<<1, 32>>>(d, c, it); rep("LDS.64 broadcast x2, 1 warp (per pair)", it * 14.0);
    lds_tput<2><<<1, 32>>>(d, c, it); rep("LDS.32 broadcast x4, 1 warp (per quad)", it * 14.0);
    lds_tput<3><<<1, 32>>>(d, c, it); rep("LDS.64 distinct per lane, 1 warp (per instr)", it * 14.0);
    lds_tput<4><<<1, 32>>>(d, c, it); rep("LDS.128 4 distinct addrs, 1 warp (per instr)", it * 14.0);
    lds_tput<0><<<1, 128>>>(d, c, it); rep("LDS.128 broadcast, 4 warps (per instr per warp)", it * 14.0);
    lds_tput<0><<<1, 256>>>(d, c, it); rep("LDS.128 broadcast, 8 warps (per instr per warp)", it * 14.0);
    lds_tput<3><<<1, 128>>>(d, c, it); rep("LDS.64 distinct, 4 warps (per instr per warp)", it * 14.0);
    shfl_tput<<<1, 32>>>(d, c, it); rep("SHFL 64-bit rotate, 1 warp (per 64-bit shuffle)", it * 28.0);
    lat_chain<<<1, 32>>>(d, c, it, 0); rep("DADD dependent latency", it * 16.0);
    lat_chain<<<1, 32>>>(d, c, it, 1); rep("DSETP+SEL(max) dependent latency", it * 16.0);
    lat_chain<<<1, 32>>>(d, c, it, 2); rep("DMUL dependent latency", it * 16.0);
    lat_chain<<<1, 32>>>(d, c, it, 3); rep("int64 min (ISETPx2+SELx2) dependent latency", it * 16.0);
    fwd_col<<<1, 32>>>(d, c, it); rep("forward column (STS,14 LDS.128,28 DFMA,3 DADD/DMUL)", it * 1.0);
    return 0;
}
