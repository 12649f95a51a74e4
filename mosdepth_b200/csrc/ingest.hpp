// ingest.hpp -- host side of the compressed-byte ingest (the step in FRONT of the hot path: what htslib's
// reader thread pool does behind open(bam, threads) mosdepth.nim:968 and bam.query :185).
//
//   chase_blocks()   every BGZF block of a byte range, found by all host threads at once.  A BGZF file is a linked list (block
//                    start -> BSIZE -> next block start): walking it touches two cache lines per block, 1.5 M blocks for half a
//                    genome = ~0.3 s on one thread.  Each thread instead takes a slice of the range, finds the first block header
//                    in it by its 16 fixed bytes (1f 8b 08 04 .. 06 00 'B' 'C' 02 00), walks to the end of its slice, and the
//                    slices are stitched (a slice is only accepted if its predecessor's walk lands exactly on its first block).
//   IngestPump       when the BAM is a mapped file (msd_open) or any pageable buffer, the bytes have to cross pinned memory on their
//                    way to the device.  r1 did that on the thread that also plans chunks and launches kernels, through four
//                    512 MB pinned buffers (0.9 s of cudaMallocHost per run, 22 GB/s: profiles/README.md r2b).  The pump is a thread
//                    of its own with a pool of copy workers and a ring of small pinned pieces: memcpy of piece p+1 overlaps the DMA
//                    of piece p, and the launch thread never copies.
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace msd {

inline int ingest_threads_default() {
    const char* e = getenv("MSD_INGEST_THREADS");
    if (e && *e) { const int v = atoi(e); return v < 1 ? 1 : v > 64 ? 64 : v; }
    const unsigned hw = std::thread::hardware_concurrency();
    return (int)std::min<unsigned>(16u, std::max<unsigned>(1u, hw / 2));
}

inline void parallel_copy(void* dst, const void* src, size_t n, int threads) {
    constexpr size_t kMinPiece = 4u << 20;
    const size_t t = std::min<size_t>((size_t)std::max(threads, 1), n / kMinPiece);
    if (t <= 1) { memcpy(dst, src, n); return; }
    const size_t piece = ((n + t - 1) / t + 4095) & ~(size_t)4095;
    std::vector<std::thread> pool;
    for (size_t i = 1; i < t; i++) {
        const size_t a = i * piece;
        if (a >= n) break;
        const size_t len = std::min(piece, n - a);
        pool.emplace_back([=]() { memcpy(static_cast<char*>(dst) + a, static_cast<const char*>(src) + a, len); });
    }
    memcpy(dst, src, std::min(piece, n));
    for (auto& th : pool) th.join();
}

// ---- BGZF block list -----------------------------------------------------------------------------------------------------------
struct BlockInfo { uint64_t coff; uint32_t blen, xlen, isize; };

// one BGZF header at `pos` (RFC 1952 member with the BC extra subfield, SAM spec 4.1); returns the block length or 0
inline uint32_t bgzf_header_at(const uint8_t* f, uint64_t n, uint64_t pos, uint32_t* xlen_out, uint32_t* isize_out) {
    if (pos + 18 > n) return 0;
    const uint8_t* h = f + pos;
    if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return 0;
    uint32_t xlen = h[10] | (h[11] << 8), q = 12, bsize = 0; bool have = false;
    if (pos + 12 + xlen > n) return 0;
    while (q + 4 <= 12 + xlen) {
        uint32_t slen = h[q + 2] | (h[q + 3] << 8);
        if (h[q] == 66 && h[q + 1] == 67 && slen == 2 && q + 6 <= 12 + xlen) { bsize = h[q + 4] | (h[q + 5] << 8); have = true; }
        q += 4 + slen;
    }
    if (!have) return 0;
    uint32_t blen = bsize + 1;
    if (blen < xlen + 20 || pos + blen > n) return 0;
    memcpy(isize_out, f + pos + blen - 4, 4);
    *xlen_out = xlen;
    return blen;
}

// walk from `pos` (a block start) while pos < stop; appends to out.  Returns the landing position, or ~0 on an invalid header (*bad = where).
inline uint64_t chase_serial(const uint8_t* f, uint64_t n, uint64_t pos, uint64_t stop, std::vector<BlockInfo>& out, uint64_t* bad) {
    while (pos < stop && pos < n) {
        uint32_t xlen = 0, isize = 0; const uint32_t blen = bgzf_header_at(f, n, pos, &xlen, &isize);
        if (!blen) { *bad = pos; return ~0ull; }
        out.push_back(BlockInfo{pos, blen, xlen, isize});
        pos += blen;
    }
    return pos;
}

// Every block that starts in [beg, stop) (beg is a block start).  Returns false on a broken chain (*bad_at = the offending offset).
inline bool chase_blocks(const uint8_t* f, uint64_t n, uint64_t beg, uint64_t stop, int threads, std::vector<BlockInfo>& out, uint64_t* bad_at) {
    out.clear();
    stop = std::min(stop, n);
    if (beg >= stop) return true;
    const uint64_t span = stop - beg;
    const char* sk = getenv("MSD_CHASE_SLICE_KB");      // slices of at least 8 MB (tests shrink them)
    const uint64_t min_slice = (sk && *sk) ? std::max<uint64_t>(128, strtoull(sk, nullptr, 10)) << 10 : (8ull << 20);
    const int T = (int)std::min<uint64_t>((uint64_t)std::max(threads, 1), span / min_slice);
    if (T <= 1) { uint64_t bad = 0; if (chase_serial(f, n, beg, stop, out, &bad) == ~0ull) { *bad_at = bad; return false; } return true; }
    struct Slice { uint64_t first = ~0ull, land = ~0ull, bad = 0; bool broken = false; std::vector<BlockInfo> blocks; };
    std::vector<Slice> sl((size_t)T);
    std::vector<std::thread> pool;
    for (int t = 0; t < T; t++) pool.emplace_back([&, t] {
        Slice& S = sl[(size_t)t];
        const uint64_t a = beg + span * (uint64_t)t / (uint64_t)T, b = beg + span * (uint64_t)(t + 1) / (uint64_t)T;
        uint64_t p = a;
        if (t > 0) {   // the first header at or after a: the fixed bytes, then a header that parses
            for (; p < b; p++) {
                const uint8_t* h = f + p;
                if (p + 18 <= n && h[0] == 0x1f && h[1] == 0x8b && h[2] == 8 && h[3] == 4 && h[12] == 66 && h[13] == 67 && h[14] == 2 && h[15] == 0) {
                    uint32_t xl, is; if (bgzf_header_at(f, n, p, &xl, &is)) break;
                }
            }
            if (p >= b) return;   // no block starts in this slice (cannot happen for slices >= 64 KB of a valid file; the stitch copes)
        }
        S.first = p;
        S.blocks.reserve((size_t)((b - p) / 12000 + 16));
        S.land = chase_serial(f, n, p, b, S.blocks, &S.bad);
        S.broken = S.land == ~0ull;
    });
    for (auto& th : pool) th.join();
    // stitch: slice t is right iff the walk so far lands exactly on its first block; otherwise walk on by hand to the next slice that fits
    uint64_t pos = beg; size_t total = 0; for (auto& S : sl) total += S.blocks.size();
    out.reserve(total + 16);
    for (int t = 0; t < T; t++) {
        Slice& S = sl[(size_t)t];
        const uint64_t b = beg + span * (uint64_t)(t + 1) / (uint64_t)T;
        if (S.first == pos) {
            if (S.broken) { out.insert(out.end(), S.blocks.begin(), S.blocks.end()); *bad_at = S.bad; return false; }
            out.insert(out.end(), S.blocks.begin(), S.blocks.end()); pos = S.land;
        } else if (pos < b) {   // a block that straddles the slice start hid the real first header (or a false signature was found): redo this slice from pos
            uint64_t bad = 0; pos = chase_serial(f, n, pos, b, out, &bad);
            if (pos == ~0ull) { *bad_at = bad; return false; }
        }
    }
    return true;
}

// ---- pinned-ring copy pump -------------------------------------------------------------------------------------------------------
class CopyWorkers {      // a persistent pool: one memcpy cut into n pieces, the caller takes piece 0 and waits for the rest
  public:
    explicit CopyWorkers(int n) {
        for (int i = 1; i < std::max(n, 1); i++) th_.emplace_back([this, i] { run(i); });
        n_ = (int)th_.size() + 1;
    }
    ~CopyWorkers() { { std::lock_guard<std::mutex> g(mu_); stop_ = true; gen_++; } cv_.notify_all(); for (auto& t : th_) t.join(); }
    void copy(void* dst, const void* src, size_t len) {
        if (n_ == 1 || len < (size_t)n_ * (1u << 20)) { memcpy(dst, src, len); return; }
        { std::lock_guard<std::mutex> g(mu_); dst_ = (char*)dst; src_ = (const char*)src; len_ = len; pending_ = n_ - 1; gen_++; }
        cv_.notify_all();
        piece(0);
        std::unique_lock<std::mutex> g(mu_);
        done_.wait(g, [this] { return pending_ == 0; });
    }
  private:
    void piece(int i) {
        const size_t per = ((len_ + n_ - 1) / n_ + 4095) & ~(size_t)4095, a = std::min(len_, per * (size_t)i), b = std::min(len_, a + per);
        if (b > a) memcpy(dst_ + a, src_ + a, b - a);
    }
    void run(int i) {
        uint64_t seen = 0;
        for (;;) {
            { std::unique_lock<std::mutex> g(mu_); cv_.wait(g, [&] { return gen_ != seen; }); seen = gen_; if (stop_) return; }
            piece(i);
            { std::lock_guard<std::mutex> g(mu_); if (--pending_ == 0) done_.notify_one(); }
        }
    }
    std::vector<std::thread> th_; int n_ = 1;
    std::mutex mu_; std::condition_variable cv_, done_;
    char* dst_ = nullptr; const char* src_ = nullptr; size_t len_ = 0; int pending_ = 0; uint64_t gen_ = 0; bool stop_ = false;
};

struct CopyRequest {
    const uint8_t* src = nullptr; uint8_t* dst = nullptr; uint64_t bytes = 0;     // pageable host bytes -> device
    cudaEvent_t wait_before = nullptr;      // the stream waits for this before the first piece (the previous reader of dst is done)
    cudaEvent_t t0 = nullptr, t1 = nullptr; // timing events around the copies (may be null)
    cudaEvent_t done = nullptr;             // recorded after the last piece
};

class IngestPump {
  public:
    IngestPump(int device, cudaStream_t stream, int threads, size_t piece_bytes, int n_pieces)
        : device_(device), stream_(stream), piece_(piece_bytes), n_pieces_(n_pieces), threads_(threads) { th_ = std::thread([this] { run(); }); }
    ~IngestPump() {
        { std::lock_guard<std::mutex> g(mu_); stop_ = true; }
        cv_.notify_all(); th_.join();
        for (auto p : ring_) if (p) cudaFreeHost(p);
        for (auto e : free_) if (e) cudaEventDestroy(e);
    }
    uint64_t submit(const CopyRequest& r) { std::lock_guard<std::mutex> g(mu_); q_.push_back(r); cv_.notify_all(); return ++submitted_; }
    // blocks until request number `ticket` (1-based, in submission order) has been handed to the stream completely; false on a CUDA error
    bool wait_issued(uint64_t ticket) { std::unique_lock<std::mutex> g(mu_); cv_done_.wait(g, [&] { return issued_ >= ticket || failed_; }); return !failed_; }
    const char* error() const { return err_; }

  private:
    void fail(const char* what, cudaError_t e) { std::lock_guard<std::mutex> g(mu_); failed_ = true; snprintf(err_, sizeof err_, "%s: %s", what, cudaGetErrorString(e)); cv_done_.notify_all(); }
    void run() {
        cudaSetDevice(device_);
        // the ring is allocated here, while the launch thread is still allocating device buffers and planning chunks
        ring_.assign((size_t)n_pieces_, nullptr); free_.assign((size_t)n_pieces_, nullptr); used_.assign((size_t)n_pieces_, 0);
        for (int i = 0; i < n_pieces_; i++) {
            cudaError_t e = cudaMallocHost((void**)&ring_[(size_t)i], piece_);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&free_[(size_t)i], cudaEventDisableTiming);
            if (e != cudaSuccess) { fail("pinned ring", e); return; }
        }
        CopyWorkers workers(threads_);
        size_t slot = 0;
        for (;;) {
            CopyRequest r;
            { std::unique_lock<std::mutex> g(mu_); cv_.wait(g, [&] { return stop_ || !q_.empty(); }); if (q_.empty()) return; r = q_.front(); q_.pop_front(); }
            cudaError_t e = cudaSuccess;
            if (r.wait_before) e = cudaStreamWaitEvent(stream_, r.wait_before, 0);
            if (e == cudaSuccess && r.t0) e = cudaEventRecord(r.t0, stream_);
            for (uint64_t off = 0; e == cudaSuccess && off < r.bytes; off += piece_) {
                const size_t len = (size_t)std::min<uint64_t>(piece_, r.bytes - off);
                if (used_[slot]) e = cudaEventSynchronize(free_[slot]);
                if (e != cudaSuccess) break;
                workers.copy(ring_[slot], r.src + off, len);
                e = cudaMemcpyAsync(r.dst + off, ring_[slot], len, cudaMemcpyHostToDevice, stream_);
                if (e == cudaSuccess) e = cudaEventRecord(free_[slot], stream_);
                used_[slot] = 1; slot = (slot + 1) % (size_t)n_pieces_;
            }
            if (e == cudaSuccess && r.t1) e = cudaEventRecord(r.t1, stream_);
            if (e == cudaSuccess && r.done) e = cudaEventRecord(r.done, stream_);
            if (e != cudaSuccess) { fail("ingest copy", e); return; }
            { std::lock_guard<std::mutex> g(mu_); issued_++; }
            cv_done_.notify_all();
        }
    }
    int device_; cudaStream_t stream_; size_t piece_; int n_pieces_, threads_;
    std::thread th_; std::mutex mu_; std::condition_variable cv_, cv_done_;
    std::deque<CopyRequest> q_; uint64_t submitted_ = 0, issued_ = 0; bool stop_ = false, failed_ = false; char err_[256] = "";
    std::vector<uint8_t*> ring_; std::vector<cudaEvent_t> free_; std::vector<char> used_;
};

}  // namespace msd
