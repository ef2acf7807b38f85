// Cold-start cost of page-locked host memory: cudaHostAlloc against mmap + parallel first touch + cudaHostRegister in pieces.
//   nvcc -O2 -o /tmp/pin_bench tools/pin_bench.cu -lpthread && /tmp/pin_bench [GiB] [threads]
#include <cuda_runtime.h>
#include <sys/mman.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
    const size_t gib = argc > 1 ? (size_t)atoi(argv[1]) : 8;
    const int nthr = argc > 2 ? atoi(argv[2]) : 16;
    const size_t n = gib << 30;
    cudaFree(0);
    double t0 = now();
    void* p = nullptr;
    if (cudaHostAlloc(&p, n, cudaHostAllocPortable) != cudaSuccess) { printf("cudaHostAlloc failed\n"); return 1; }
    double t1 = now();
    printf("{\"gib\": %zu, \"cudaHostAlloc_s\": %.3f", gib, t1 - t0);
    cudaFreeHost(p);
    double t2 = now();
    printf(", \"cudaFreeHost_s\": %.3f", t2 - t1);

    t0 = now();
    uint8_t* q = (uint8_t*)mmap(nullptr, n, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (q == MAP_FAILED) { printf("mmap failed\n"); return 1; }
    madvise(q, n, MADV_HUGEPAGE);
    std::vector<std::thread> th;
    for (int t = 0; t < nthr; t++)
        th.emplace_back([=] {
            const size_t a = n / nthr * t, b = (t + 1 == nthr) ? n : n / nthr * (t + 1);
            for (size_t i = a; i < b; i += 4096) q[i] = 0;
        });
    for (auto& x : th) x.join();
    t1 = now();
    printf(", \"threads\": %d, \"touch_s\": %.3f", nthr, t1 - t0);
    const size_t piece = 1ull << 30;
    for (size_t a = 0; a < n; a += piece)
        if (cudaHostRegister(q + a, (n - a < piece) ? n - a : piece, cudaHostRegisterPortable) != cudaSuccess) { printf("register failed\n"); return 1; }
    t2 = now();
    printf(", \"register_s\": %.3f", t2 - t1);
    // parallel registration of the pieces
    for (size_t a = 0; a < n; a += piece) cudaHostUnregister(q + a);
    double t3 = now();
    printf(", \"unregister_s\": %.3f", t3 - t2);
    th.clear();
    const size_t npieces = (n + piece - 1) / piece;
    for (int t = 0; t < nthr; t++)
        th.emplace_back([=] {
            for (size_t k = t; k < npieces; k += nthr) cudaHostRegister(q + k * piece, (n - k * piece < piece) ? n - k * piece : piece, cudaHostRegisterPortable);
        });
    for (auto& x : th) x.join();
    double t4 = now();
    printf(", \"register_parallel_s\": %.3f", t4 - t3);
    // a copy out of it works at full speed?
    void* d;
    cudaMalloc(&d, 1ull << 30);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < 4; i++) cudaMemcpyAsync(q + ((size_t)i % gib << 30), d, 1ull << 30, cudaMemcpyDeviceToHost);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf(", \"d2h_gb_per_s\": %.1f}\n", 4.0 * 1.073741824 / (ms * 1e-3));
    return 0;
}
