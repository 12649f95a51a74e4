// h2d_probe.cu -- what the box can move from host memory into N GPUs at once (SURVEY 7.1's "first gpurun call").
//
// One host thread per GPU, all started together.  Every thread copies `gb_per_gpu` GB to its device with plain
// cudaMemcpyAsync in pieces of `piece_mb`, from (a) cudaMallocHost memory, (b) a private file in /dev/shm that is
// mmap()ed and cudaHostRegister()ed (what msd_open does with a BAM in the page cache), (c) pageable memory through a
// pinned bounce buffer filled by memcpy on `copy_threads` threads (the library's fallback).  Device-timed with CUDA
// events per GPU; the aggregate is bytes of all GPUs / the slowest GPU's time.  Prints one JSON line.
//
// usage: h2d_probe [n_gpus=1] [gb_per_gpu=4] [piece_mb=256] [copy_threads=8]
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include <fcntl.h>
#include <string>
#include <sys/mman.h>
#include <thread>
#include <unistd.h>
#include <vector>

#define OK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Result { double pinned_gbs = 0, registered_gbs = 0, register_s = 0, bounce_gbs = 0, memcpy_gbs = 0, alloc_pinned_s = 0; };

static void spin_barrier(std::atomic<int>& c, int n) { c.fetch_add(1); while (c.load() % n) std::this_thread::yield(); }

static void par_memcpy(uint8_t* dst, const uint8_t* src, size_t n, int threads) {
    std::vector<std::thread> th; const size_t per = (n + threads - 1) / threads;
    for (int t = 0; t < threads; t++) { const size_t a = std::min(n, per * t), b = std::min(n, a + per); if (b > a) th.emplace_back([=] { memcpy(dst + a, src + a, b - a); }); }
    for (auto& x : th) x.join();
}

int main(int argc, char** argv) {
    int n_gpus = argc > 1 ? atoi(argv[1]) : 1; const double gb = argc > 2 ? atof(argv[2]) : 4.0; const size_t piece = (size_t)(argc > 3 ? atoi(argv[3]) : 256) << 20;
    const int copy_threads = argc > 4 ? atoi(argv[4]) : 8;
    int have = 0; OK(cudaGetDeviceCount(&have)); n_gpus = std::min(n_gpus, have);
    const size_t total = ((size_t)(gb * 1e9) / piece) * piece;
    std::vector<Result> res(n_gpus);
    std::atomic<int> b1{0}, b2{0}, b3{0}, b4{0};
    std::vector<std::thread> th;
    for (int g = 0; g < n_gpus; g++) th.emplace_back([&, g] {
        OK(cudaSetDevice(g));
        cudaStream_t s; OK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        uint8_t* d; OK(cudaMalloc(&d, total));
        cudaEvent_t e0, e1; OK(cudaEventCreate(&e0)); OK(cudaEventCreate(&e1));
        auto timed_copy = [&](const uint8_t* src, std::atomic<int>& bar) {
            OK(cudaMemcpyAsync(d, src, piece, cudaMemcpyHostToDevice, s)); OK(cudaStreamSynchronize(s));   // warm
            spin_barrier(bar, n_gpus);
            OK(cudaEventRecord(e0, s));
            for (size_t o = 0; o < total; o += piece) OK(cudaMemcpyAsync(d + o, src + o, piece, cudaMemcpyHostToDevice, s));
            OK(cudaEventRecord(e1, s)); OK(cudaStreamSynchronize(s));
            float ms; OK(cudaEventElapsedTime(&ms, e0, e1)); return total / 1e9 / (ms / 1e3);
        };
        // (a) cudaMallocHost
        uint8_t* h; double t0 = now(); OK(cudaMallocHost(&h, total)); memset(h, g + 1, total); res[g].alloc_pinned_s = now() - t0;
        res[g].pinned_gbs = timed_copy(h, b1);
        OK(cudaFreeHost(h));
        // (b) mmap of a /dev/shm file + cudaHostRegister (read-only mapping)
        const std::string path = "/dev/shm/h2d_probe_" + std::to_string(g) + ".bin";
        int fd = open(path.c_str(), O_CREAT | O_RDWR | O_TRUNC, 0600);
        if (fd >= 0 && ftruncate(fd, (off_t)total) == 0) {
            uint8_t* w = (uint8_t*)mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
            if (w != MAP_FAILED) { memset(w, g + 3, total); munmap(w, total); }
            uint8_t* m = (uint8_t*)mmap(nullptr, total, PROT_READ, MAP_SHARED, fd, 0);
            if (m != MAP_FAILED) {
                t0 = now();
                cudaError_t e = cudaHostRegister(m, total, cudaHostRegisterReadOnly);
                if (e != cudaSuccess) { cudaGetLastError(); e = cudaHostRegister(m, total, cudaHostRegisterDefault); }
                res[g].register_s = now() - t0;
                if (e == cudaSuccess) { res[g].registered_gbs = timed_copy(m, b2); cudaHostUnregister(m); }
                else { cudaGetLastError(); spin_barrier(b2, n_gpus); res[g].registered_gbs = -1; }
                // (c) pageable (the same mapping, unregistered) -> pinned bounce (memcpy on copy_threads threads) -> device, double buffered
                uint8_t* bb[2]; OK(cudaMallocHost(&bb[0], piece)); OK(cudaMallocHost(&bb[1], piece));
                cudaEvent_t fe[2]; OK(cudaEventCreate(&fe[0])); OK(cudaEventCreate(&fe[1]));
                spin_barrier(b3, n_gpus);
                t0 = now(); double tm = 0;
                for (size_t o = 0, k = 0; o < total; o += piece, k ^= 1) {
                    OK(cudaEventSynchronize(fe[k]));
                    const double c0 = now(); par_memcpy(bb[k], m + o, piece, copy_threads); tm += now() - c0;
                    OK(cudaMemcpyAsync(d + o, bb[k], piece, cudaMemcpyHostToDevice, s)); OK(cudaEventRecord(fe[k], s));
                }
                OK(cudaStreamSynchronize(s));
                res[g].bounce_gbs = total / 1e9 / (now() - t0); res[g].memcpy_gbs = total / 1e9 / tm;
                cudaFreeHost(bb[0]); cudaFreeHost(bb[1]);
                munmap(m, total);
            }
            close(fd);
        }
        unlink(path.c_str());
        spin_barrier(b4, n_gpus);
        cudaFree(d);
    });
    for (auto& t : th) t.join();
    auto agg = [&](double Result::*f) { double mn = 1e30; for (auto& r : res) mn = std::min(mn, r.*f); return mn * n_gpus; };
    printf("{\"n_gpus\": %d, \"gb_per_gpu\": %.2f, \"piece_mb\": %zu, \"copy_threads\": %d, \"pinned_gbs_aggregate\": %.1f, \"registered_mmap_gbs_aggregate\": %.1f, \"bounce_gbs_aggregate\": %.1f, \"per_gpu\": [",
           n_gpus, total / 1e9, piece >> 20, copy_threads, agg(&Result::pinned_gbs), agg(&Result::registered_gbs), agg(&Result::bounce_gbs));
    for (int g = 0; g < n_gpus; g++)
        printf("%s{\"pinned_gbs\": %.1f, \"registered_mmap_gbs\": %.1f, \"register_s\": %.3f, \"bounce_gbs\": %.1f, \"bounce_memcpy_gbs\": %.1f, \"alloc_pinned_s\": %.3f}", g ? ", " : "",
               res[g].pinned_gbs, res[g].registered_gbs, res[g].register_s, res[g].bounce_gbs, res[g].memcpy_gbs, res[g].alloc_pinned_s);
    printf("]}\n");
    return 0;
}
