// Issue throughput of the instructions of the FP32 screen (viterbi_check32_kernel) on one SM
// sub-partition: warp-instructions per cycle with 1, 2, 3, 4 warps per scheduler, each warp
// running NI independent chains of one instruction type.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/ubench3 tools/ubench3.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int NI = 16, ITER = 4096;

__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
    float r;
    asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float max2(float a, float b) {
    float r;
    asm volatile("max.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float fadd(float a, float b) {
    float r;
    asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ int imax(int a, int b) {
    int r;
    asm volatile("max.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ unsigned lop(unsigned a, unsigned b, unsigned c) {
    unsigned r;
    asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ double dadd_(double a, double b) {
    double r;
    asm volatile("add.rn.f64 %0, %1, %2;" : "=d"(r) : "d"(a), "d"(b));
    return r;
}

template <int MODE>
__global__ void k(float *out, long long *cyc, float seed) {
    __shared__ __align__(16) float sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = seed * i;
    __syncthreads();
    float f[NI];
    unsigned long long u[NI];
    double d[NI / 2];
#pragma unroll
    for (int i = 0; i < NI; ++i) { f[i] = seed + i + threadIdx.x; u[i] = (unsigned long long)__float_as_uint(f[i]) * 0x100000001ull; }
#pragma unroll
    for (int i = 0; i < NI / 2; ++i) d[i] = seed + i;
    const unsigned long long ub = (unsigned long long)__float_as_uint(seed) * 0x100000001ull;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < NI; ++i) f[i] = fadd(f[i], seed);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < NI; ++i) u[i] = add2(u[i], ub);
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < NI; ++i) f[i] = max2(f[i], seed);
        } else if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < NI; ++i) f[i] = max3(f[i], seed, f[(i + 1) % NI]);
        } else if (MODE == 4) {      // the screen's mix: FADD2 + FMNMX3 alternating
#pragma unroll
            for (int i = 0; i < NI; i += 2) {
                u[i] = add2(u[i], ub);
                f[i + 1] = max3(f[i + 1], __uint_as_float((unsigned)u[i]), __uint_as_float((unsigned)(u[i] >> 32)));
            }
        } else if (MODE == 5) {
#pragma unroll
            for (int i = 0; i < NI; ++i) f[i] = __int_as_float(imax(__float_as_int(f[i]), __float_as_int(seed)));
        } else if (MODE == 6) {
#pragma unroll
            for (int i = 0; i < NI; ++i) f[i] = __uint_as_float(lop(__float_as_uint(f[i]), 0x1234u, __float_as_uint(seed)));
        } else if (MODE == 7) {
#pragma unroll
            for (int i = 0; i < NI / 2; ++i) d[i] = dadd_(d[i], (double)seed);
        } else if (MODE == 8) {      // broadcast LDS.128
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((unsigned)__cvta_generic_to_shared(sm) + 16u * ((i + it) & 63)));
                f[i] += v.x;
            }
        } else if (MODE == 9) {      // scalar FADD + FMNMX alternating (the same work as mode 4 in 2x the instructions)
#pragma unroll
            for (int i = 0; i < NI; i += 2) {
                f[i] = fadd(f[i], seed);
                f[i + 1] = max2(f[i + 1], f[i]);
            }
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < NI; ++i) s += f[i] + __uint_as_float((unsigned)u[i]);
#pragma unroll
    for (int i = 0; i < NI / 2; ++i) s += (float)d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int ninst) {
    float *out;
    long long *cyc, h;
    cudaMalloc(&out, 1 << 20);
    cudaMalloc(&cyc, 8);
    printf("%-28s", name);
    for (int wps = 1; wps <= 4; ++wps) {
        k<MODE><<<1, 128 * wps>>>(out, cyc, 1.5f);
        k<MODE><<<1, 128 * wps>>>(out, cyc, 1.5f);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        const double per = (double)h / ((double)ITER * ninst);          // cycles per instruction of ONE warp
        printf("  %dw: %5.2f cyc/instr/warp = %4.2f instr/cyc/sched", wps, per, wps / per);
    }
    printf("\n");
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    run<0>("FADD", NI);
    run<1>("FADD2 (add.f32x2)", NI);
    run<2>("FMNMX", NI);
    run<3>("FMNMX3 (max.f32 3-input)", NI);
    run<4>("FADD2 + FMNMX3 mix", NI);
    run<5>("IMNMX", NI);
    run<6>("LOP3", NI);
    run<7>("DADD", NI / 2);
    run<8>("LDS.128 broadcast (+FADD)", 2 * NI);
    run<9>("FADD + FMNMX mix", NI);
    return 0;
}
