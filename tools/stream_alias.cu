// Which CUDA streams share a hardware work queue?  A long spin kernel goes into stream k,
// a tiny kernel + event into every other stream; events that are not complete while the spin
// kernel still runs belong to streams that alias with k.
// nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/stream_alias tools/stream_alias.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <unistd.h>
#include <cuda_runtime.h>
__global__ void spin(long long cycles) { long long t0 = clock64(); while (clock64() - t0 < cycles) {} }
__global__ void tiny() {}
int main(int argc, char **argv) {
    const int NS = argc > 1 ? atoi(argv[1]) : 24;
    std::vector<cudaStream_t> s(NS);
    std::vector<cudaEvent_t> e(NS);
    for (int i = 0; i < NS; ++i) { cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking); cudaEventCreateWithFlags(&e[i], cudaEventDisableTiming); }
    tiny<<<1, 1>>>(); cudaDeviceSynchronize();
    printf("CUDA_DEVICE_MAX_CONNECTIONS=%s, %d streams\n", getenv("CUDA_DEVICE_MAX_CONNECTIONS") ? getenv("CUDA_DEVICE_MAX_CONNECTIONS") : "(unset)", NS);
    const bool copy_mode = argc > 2;
    char *dbuf = nullptr, *hbuf = nullptr;
    const size_t CB = 64u << 20;
    if (copy_mode) { cudaMalloc(&dbuf, CB); cudaMallocHost(&hbuf, CB); printf("blocking work = 16 x 64 MB D2H copies\n"); }
    for (int k = 0; k < NS; ++k) {
        if (copy_mode) for (int r = 0; r < 16; ++r) cudaMemcpyAsync(hbuf, dbuf, CB, cudaMemcpyDeviceToHost, s[k]);
        else spin<<<1, 1, 0, s[k]>>>(20000000LL);       // ~10 ms
        for (int i = 0; i < NS; ++i) if (i != k) { tiny<<<1, 1, 0, s[i]>>>(); cudaEventRecord(e[i], s[i]); }
        usleep(3000);
        printf("stream %2d aliases with:", k);
        for (int i = 0; i < NS; ++i) if (i != k && cudaEventQuery(e[i]) != cudaSuccess) printf(" %d", i);
        printf("\n");
        cudaDeviceSynchronize();
    }
    return 0;
}
