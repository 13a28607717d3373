// Per-kernel device timing with CUDA events on the library's launch stream (used by bench.py for the
// roofline numbers; disabled by default so the product path pays nothing).
// With B200_NVTX=1 every timed section is also an NVTX range (header-only NVTX v3: a no-op unless a profiler injects its library), so an
// nsys / ncu --nvtx timeline shows the stages of stark_gen ("tracing"; the reference logs stage times with `log::info!`).
#include "b200_internal.h"
#include <nvtx3/nvToolsExt.h>
#include <cstdlib>
#include <map>
#include <atomic>
#include <mutex>

namespace b200 {
struct Rec { std::string name; double bytes; cudaEvent_t a, b; };
static bool g_on = false;
static std::vector<Rec> g_recs;
static std::vector<size_t> g_open;
static std::mutex g_mu;                    // the recorder is process-wide (a bench tool); entries from concurrent devices interleave
static std::atomic<u64> g_launches{0};
static const bool g_nvtx = [] { const char* e = getenv("B200_NVTX"); return e && e[0] == '1'; }();

void timing_enable(bool on) { g_on = on; }
void timing_reset() {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& r : g_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_recs.clear(); g_open.clear();
}
void timing_begin(const char* name, double bytes) {
    if (g_nvtx) nvtxRangePushA(name);
    if (!g_on) return;
    std::lock_guard<std::mutex> lk(g_mu);
    Rec r; r.name = name; r.bytes = bytes;
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    cudaEventRecord(r.a, stream());
    g_recs.push_back(r); g_open.push_back(g_recs.size() - 1);
}
void timing_end() {
    if (g_nvtx) nvtxRangePop();
    if (!g_on) return;
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_open.empty()) return;
    cudaEventRecord(g_recs[g_open.back()].b, stream());
    g_open.pop_back();
}
std::vector<TimingRow> timing_collect() {
    std::vector<TimingRow> out;
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_recs.empty()) return out;
    cudaStreamSynchronize(stream());
    std::map<std::string, size_t> idx;
    for (auto& r : g_recs) {
        float ms = 0; cudaEventElapsedTime(&ms, r.a, r.b);
        auto it = idx.find(r.name);
        if (it == idx.end()) { idx[r.name] = out.size(); out.push_back(TimingRow{r.name, 1, ms, r.bytes}); }
        else { out[it->second].launches++; out[it->second].ms += ms; out[it->second].bytes += r.bytes; }
    }
    return out;
}
u64 launch_count() { return g_launches.load(); }
void launch_count_add(u64 n) { g_launches += n; }

// ---------------------------------------------------------------------------------------------- arena
void Arena::reserve(size_t bytes) {
    if (bytes <= cap) return;
    release();
    B200_CUDA_CHECK(cudaMalloc(&base, bytes));
    cap = bytes; off = 0;
}
u64* Arena::alloc_u64(size_t n) {
    size_t bytes = (n * 8 + 255) & ~(size_t)255;
    if (off + bytes > cap) throw std::runtime_error("device arena exhausted: need " + std::to_string(off + bytes) + " of " + std::to_string(cap));
    u64* p = reinterpret_cast<u64*>(base + off);
    off += bytes; if (off > high) high = off;
    return p;
}
void Arena::release() { if (base) { cudaFree(base); base = nullptr; } cap = off = 0; }
}  // namespace b200
