// pp_api.cu — library-level plumbing of the C ABI: error text, launch accounting.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "pp_common.cuh"

namespace pp {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();  // clear the (non-sticky) launch error
        set_error("%s: %s", what, cudaGetErrorString(e));
        return PP_ERR_CUDA;
    }
    return PP_OK;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- certification of the exact constant division (pp_common.cuh, Div<DM_FAST>) -------------
// For divisor s: check q == x/s for all 2^23 mantissas x in [1,2).  fmaf/division here are the
// host's IEEE operations (this file is compiled without contraction: -fmad only affects device
// code and the expressions below are explicit).  ~20 ms per new divisor, cached.
static std::mutex g_cert_mu;
static std::map<uint32_t, bool> g_cert;

bool div_certified(float s) {
    if (!(s >= 9.5367431640625e-07f && s <= 16777216.0f)) return false;  // 2^-20 .. 2^24
    uint32_t key;
    memcpy(&key, &s, 4);
    {
        std::lock_guard<std::mutex> lk(g_cert_mu);
        auto it = g_cert.find(key);
        if (it != g_cert.end()) return it->second;
    }
    const volatile float inv_v = 1.0f / s;
    const float inv = inv_v;
    const volatile float lo_v = recip_lo(s);
    const float inv_lo = lo_v;
    bool ok = (inv_lo == 0.0f) || (fabsf(inv_lo) >= 1.3552527156068805e-20f);  // 2^-66
    for (uint32_t m = 0; m < (1u << 23) && ok; m++) {
        uint32_t bits = 0x3f800000u | m;
        float x;
        memcpy(&x, &bits, 4);
        volatile float t0 = x * inv_lo;
        float q = fmaf(x, inv, t0);
        volatile float t = x / s;
        ok = (q == t);
    }
    std::lock_guard<std::mutex> lk(g_cert_mu);
    g_cert[key] = ok;
    return ok;
}

// ---- per-kernel device timing (tracing aid; off by default) --------------------------------
// When enabled, every launch site brackets its kernel with a cudaEvent pair on the launch
// stream.  pp_profile_get() synchronises the recorded events and aggregates by kernel name.
struct ProfRec {
    const char* name;
    cudaEvent_t e0, e1;
};
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::vector<std::string> g_prof_names;
static std::vector<int64_t> g_prof_count;
static std::vector<double> g_prof_ms;

ProfScope::ProfScope(const char* name, cudaStream_t st) : name_(name), st_(st), e1_(nullptr) {
    // Runs right before every launch of this library (PP_LAUNCH): drop a stale NON-sticky error an earlier call of another
    // library left in this thread's error slot, so that check_launch() after the launch reports this launch and nothing else
    // (a sticky error survives cudaGetLastError and is still reported).
    (void)cudaGetLastError();
    if (!g_prof_on) return;
    cudaEvent_t e0;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1_);
    cudaEventRecord(e0, st_);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(ProfRec{name_, e0, e1_});
}
ProfScope::~ProfScope() {
    if (e1_) cudaEventRecord(e1_, st_);
}

static void prof_clear() {
    for (auto& r : g_prof) {
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    g_prof.clear();
}

static void prof_aggregate() {
    g_prof_names.clear();
    g_prof_count.clear();
    g_prof_ms.clear();
    for (auto& r : g_prof) {
        cudaEventSynchronize(r.e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        size_t i = 0;
        for (; i < g_prof_names.size(); i++)
            if (g_prof_names[i] == r.name) break;
        if (i == g_prof_names.size()) {
            g_prof_names.push_back(r.name);
            g_prof_count.push_back(0);
            g_prof_ms.push_back(0.0);
        }
        g_prof_count[i] += 1;
        g_prof_ms[i] += ms;
    }
}

}  // namespace pp

extern "C" {

int pp_abi_version(void) { return 1; }
const char* pp_last_error(void) { return pp::g_err; }
int64_t pp_launch_count(void) { return pp::g_launches.load(std::memory_order_relaxed); }

int pp_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(pp::g_prof_mu);
    pp::prof_clear();
    pp::g_prof_on = on != 0;
    return PP_OK;
}

int pp_profile_num_kernels(void) {
    std::lock_guard<std::mutex> lk(pp::g_prof_mu);
    pp::prof_aggregate();
    return (int)pp::g_prof_names.size();
}

int pp_profile_get(int idx, char* name, int name_cap, int64_t* launches, double* total_ms) {
    std::lock_guard<std::mutex> lk(pp::g_prof_mu);
    if (idx < 0 || idx >= (int)pp::g_prof_names.size() || !name || name_cap <= 0) {
        pp::set_error("pp_profile_get: bad index %d", idx);
        return PP_ERR_INVALID;
    }
    strncpy(name, pp::g_prof_names[idx].c_str(), name_cap - 1);
    name[name_cap - 1] = 0;
    if (launches) *launches = pp::g_prof_count[idx];
    if (total_ms) *total_ms = pp::g_prof_ms[idx];
    return PP_OK;
}

}  // extern "C"
