// common.cu -- error plumbing, device checks, version.
#include "common.cuh"
#include <algorithm>
#include <cstdlib>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace sedb200 {

char* err_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

namespace {
struct DevInfo { int checked = 0, major = 0, minor = 0, sms = 0; };
DevInfo g_dev[64];
std::mutex g_mu;

int probe(int dev, DevInfo** out) {
    if (dev < 0 || dev >= 64) return fail(SEDB200_EINVAL, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lk(g_mu);
    DevInfo& d = g_dev[dev];
    if (!d.checked) {
        SED_CUDA_OK(cudaDeviceGetAttribute(&d.major, cudaDevAttrComputeCapabilityMajor, dev));
        SED_CUDA_OK(cudaDeviceGetAttribute(&d.minor, cudaDevAttrComputeCapabilityMinor, dev));
        SED_CUDA_OK(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev));
        d.checked = 1;
    }
    *out = &d;
    return SEDB200_OK;
}
}  // namespace

int require_sm100() {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess)
        return fail(SEDB200_ECUDA, "cudaGetDevice -> %s (no CUDA device: this library has no CPU path)",
                    cudaGetErrorString(e));
    DevInfo* d = nullptr;
    int rc = probe(dev, &d);
    if (rc) return rc;
    if (d->major != 10)
        return fail(SEDB200_EARCH, "device %d is sm_%d%d; libsedb200 is built for sm_100a only", dev,
                    d->major, d->minor);
    return SEDB200_OK;
}

namespace {
std::mutex g_smem_mu;
std::map<std::pair<int, const void*>, int> g_smem_set;     // (device, kernel) -> bytes already granted
}  // namespace

int ensure_dyn_smem(const void* func, int bytes) {
    int dev = 0;
    SED_CUDA_OK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_smem_mu);
    auto it = g_smem_set.find({dev, func});
    if (it != g_smem_set.end() && it->second >= bytes) return SEDB200_OK;
    SED_CUDA_OK(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    g_smem_set[{dev, func}] = bytes;
    return SEDB200_OK;
}

namespace {
std::atomic<long> g_launches{0};
std::atomic<bool> g_prof{false};
struct ProfRec { std::string name; cudaEvent_t a, b; };
std::vector<ProfRec> g_recs;
std::vector<size_t> g_open;          // stack of scopes whose end event is not recorded yet
std::mutex g_prof_mu;
}  // namespace

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
namespace {
std::atomic<cudaStream_t> g_plain_streams[16];
std::atomic<int> g_n_plain{0};
}
void pdl_exclude_stream(cudaStream_t st) {
    if (pdl_excluded(st)) return;
    const int i = g_n_plain.fetch_add(1);
    if (i < 16) g_plain_streams[i].store(st);
}
bool pdl_excluded(cudaStream_t st) {
    const int n = std::min(16, g_n_plain.load(std::memory_order_acquire));
    for (int i = 0; i < n; ++i) if (g_plain_streams[i].load(std::memory_order_relaxed) == st) return true;
    return false;
}
bool& pdl_in_capture() {
    static thread_local bool on = false;
    return on;
}
bool pdl_on() {
    static const bool on = [] { const char* e = std::getenv("SEDB200_PDL"); return !e || std::atoi(e) != 0; }();
    return on;
}
bool prof_on() { return g_prof.load(std::memory_order_relaxed); }
void prof_begin(const char* name, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfRec r;
    r.name = name;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    g_open.push_back(g_recs.size());
    g_recs.push_back(r);
}
void prof_end(cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (g_open.empty()) return;
    cudaEventRecord(g_recs[g_open.back()].b, st);
    g_open.pop_back();
}

int sm_count() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    DevInfo* d = nullptr;
    if (probe(dev, &d)) return 148;
    return d->sms > 0 ? d->sms : 148;
}

}  // namespace sedb200

extern "C" {

int sedb200_version(void) { return SEDB200_VERSION; }

const char* sedb200_last_error(void) { return sedb200::err_buf(); }

long sedb200_launch_count(void) { return sedb200::g_launches.load(); }

int sedb200_prof_enable(int on) {
    std::lock_guard<std::mutex> lk(sedb200::g_prof_mu);
    for (auto& r : sedb200::g_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    sedb200::g_recs.clear();
    sedb200::g_open.clear();
    sedb200::g_prof.store(on != 0);
    return SEDB200_OK;
}

int sedb200_prof_report(char* buf, size_t n) {
    if (!buf || n == 0) return sedb200::fail(SEDB200_EINVAL, "prof_report: no buffer");
    std::lock_guard<std::mutex> lk(sedb200::g_prof_mu);
    std::map<std::string, std::pair<double, long>> acc;
    std::vector<std::string> order;
    for (auto& r : sedb200::g_recs) {
        if (cudaEventSynchronize(r.b) != cudaSuccess) continue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) continue;
        if (!acc.count(r.name)) order.push_back(r.name);
        acc[r.name].first += ms;
        acc[r.name].second += 1;
    }
    size_t o = 0;
    buf[0] = 0;
    for (auto& k : order) {
        int w = snprintf(buf + o, n - o, "%s %.6f %ld\n", k.c_str(), acc[k].first, acc[k].second);
        if (w < 0 || (size_t)w >= n - o) break;
        o += (size_t)w;
    }
    return SEDB200_OK;
}

int sedb200_device_check(int device) {
    if (device >= 0) {
        cudaError_t e = cudaSetDevice(device);
        if (e != cudaSuccess)
            return sedb200::fail(SEDB200_ECUDA, "cudaSetDevice(%d) -> %s", device, cudaGetErrorString(e));
    }
    return sedb200::require_sm100();
}

}  // extern "C"
