// d2h_ceiling.cu — the box's device-to-host ceiling at N = 1, 2, 4, 8 GPUs.
//
// Bare cudaMemcpyAsync from device memory into page-locked host memory, one host thread
// and one stream per GPU, all GPUs at the same time; one cudaMemcpyAsync per piece (no
// batch-copy API).  Two destinations: a ring of 4 pinned slots (what itr_posterior_stream
// uses) and one large pinned buffer; plain and write-combined pinned memory.  This is
// the number bench.py's e2e is compared with (`e2e.frac_of_d2h_ceiling`).
//
//   nvcc -O2 -o tools/bin/d2h_ceiling tools/d2h_ceiling.cu -lpthread
//   tools/bin/d2h_ceiling [GB per GPU = 6] > gpurun_out/d2h_ceiling.json
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

struct Barrier {
    std::atomic<int> count{0}, gen{0};
    int n;
    explicit Barrier(int n_) : n(n_) {}
    void wait() {
        const int g = gen.load();
        if (count.fetch_add(1) + 1 == n) { count = 0; gen++; }
        else while (gen.load() == g) std::this_thread::yield();
    }
};

static double run(int n_gpus, size_t bytes_per_gpu, size_t piece, int n_slots, unsigned flags, bool ring) {
    Barrier bar(n_gpus + 1);
    std::vector<std::thread> th;
    std::vector<double> secs(n_gpus, 0.0);
    for (int d = 0; d < n_gpus; ++d)
        th.emplace_back([&, d]() {
            CK(cudaSetDevice(d));
            char *src = nullptr, *dst = nullptr;
            const size_t src_bytes = 1ull << 30;
            CK(cudaMalloc(&src, src_bytes));
            CK(cudaMemset(src, 1, src_bytes));
            const size_t dst_bytes = ring ? piece * n_slots : bytes_per_gpu;
            CK(cudaHostAlloc(&dst, dst_bytes, flags));
            cudaStream_t st;
            CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
            std::vector<cudaEvent_t> ev(n_slots);
            for (auto &e : ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            const size_t n_pieces = bytes_per_gpu / piece;
            for (int rep = 0; rep < 2; ++rep) {          // rep 0 warms up (first touch of the pinned pages)
                bar.wait();
                const auto t0 = std::chrono::steady_clock::now();
                for (size_t i = 0; i < n_pieces; ++i) {
                    const int s = (int)(i % n_slots);
                    if (i >= (size_t)n_slots) CK(cudaEventSynchronize(ev[s]));
                    char *to = ring ? dst + (size_t)s * piece : dst + i * piece;
                    CK(cudaMemcpyAsync(to, src + (i * piece) % (src_bytes - piece + 1) / 256 * 256, piece, cudaMemcpyDeviceToHost, st));
                    CK(cudaEventRecord(ev[s], st));
                }
                CK(cudaStreamSynchronize(st));
                secs[d] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                bar.wait();
            }
            CK(cudaFreeHost(dst));
            CK(cudaFree(src));
        });
    double worst = 0.0;
    for (int rep = 0; rep < 2; ++rep) {
        bar.wait();
        bar.wait();
        worst = 0.0;
        for (double s : secs) worst = s > worst ? s : worst;
    }
    for (auto &t : th) t.join();
    return (double)n_gpus * (double)(bytes_per_gpu / piece * piece) / worst / 1e9;
}

int main(int argc, char **argv) {
    const double gb = argc > 1 ? atof(argv[1]) : 6.0;
    int n_dev = 0;
    CK(cudaGetDeviceCount(&n_dev));
    const size_t bytes = (size_t)(gb * 1e9);
    printf("{\n \"what\": \"aggregate device-to-host GB/s into page-locked memory, all GPUs copying at once, %.1f GB per GPU\",\n", gb);
    printf(" \"devices\": %d,\n", n_dev);
    const char *sep = "";
    struct Cfg { const char *name; size_t piece; int slots; unsigned flags; bool ring; } cfgs[] = {
        {"ring_4x256MiB", 256ull << 20, 4, cudaHostAllocDefault, true},
        {"ring_4x256MiB_write_combined", 256ull << 20, 4, cudaHostAllocWriteCombined, true},
    };
    for (const Cfg &c : cfgs) {
        printf("%s \"%s\": {", sep, c.name);
        sep = ",\n";
        const char *s2 = "";
        for (int n = 1; n <= n_dev && n <= 8; n *= 2) {
            const double g = run(n, bytes, c.piece, c.slots, c.flags, c.ring);
            printf("%s\"%d\": %.2f", s2, n, g);
            s2 = ", ";
            fflush(stdout);
        }
        printf("}");
    }
    printf("\n}\n");
    return 0;
}
