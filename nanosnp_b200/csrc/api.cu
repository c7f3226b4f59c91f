// C-ABI plumbing: error reporting, defaults, device-side status decoding.
#include <stdarg.h>
#include <mutex>
#include <vector>
#include "common.cuh"

namespace nsnp {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

int cuda_status(const char* what) {
    const cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return NSNP_OK;
    return set_error(NSNP_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

// ---- per-kernel timing ----------------------------------------------------------------------------
struct ProfPending { int slot; cudaEvent_t a, b; };
static bool g_prof_on = false;
static std::vector<ProfPending> g_prof_pending;
static std::vector<cudaEvent_t> g_prof_pool;
static std::mutex g_prof_mu;

static cudaEvent_t prof_event() {
    if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
ProfScope::ProfScope(int slot_, cudaStream_t s) : slot(slot_), stream(s), pending(nullptr) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfPending* p = new ProfPending{slot, prof_event(), prof_event()};
    cudaEventRecord(p->a, stream);
    pending = p;
}
ProfScope::~ProfScope() {
    if (!pending) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfPending* p = (ProfPending*)pending;
    cudaEventRecord(p->b, stream);
    g_prof_pending.push_back(*p);
    delete p;
}

}  // namespace nsnp

extern "C" {

void nsnp_profile_enable(int on) { nsnp::g_prof_on = on != 0; }

int nsnp_profile_read(double* ms_out, int64_t* launches_out) {
    std::lock_guard<std::mutex> lk(nsnp::g_prof_mu);
    for (int i = 0; i < NSNP_PROF_SLOTS; ++i) { if (ms_out) ms_out[i] = 0.0; if (launches_out) launches_out[i] = 0; }
    for (auto& p : nsnp::g_prof_pending) {
        float ms = 0.f;
        if (cudaEventSynchronize(p.b) != cudaSuccess || cudaEventElapsedTime(&ms, p.a, p.b) != cudaSuccess)
            return nsnp::set_error(NSNP_E_CUDA, "nsnp_profile_read: %s", cudaGetErrorString(cudaGetLastError()));
        if (ms_out) ms_out[p.slot] += ms;
        if (launches_out) launches_out[p.slot] += 1;
        nsnp::g_prof_pool.push_back(p.a); nsnp::g_prof_pool.push_back(p.b);
    }
    nsnp::g_prof_pending.clear();
    return NSNP_OK;
}


int nsnp_abi_version(void) { return NSNP_ABI_VERSION; }
const char* nsnp_last_error(void) { return nsnp::g_err; }

void nsnp_default_params(nsnp_params_t* p) {
    if (!p) return;
    p->snp_min_af = 0.12;      // make_predict_data.sh:122
    p->indel_min_af = 0.12;    // make_predict_data.sh:123
    p->min_coverage = 6;       // make_predict_data.sh:124
    p->min_mapq = 20;          // make_predict_data.sh:117 --min-MQ 20
    p->excl_flags = 2316;      // make_predict_data.sh:117 --excl-flags 2316
    p->max_depth = 144;        // make_predict_data.sh:117 --max-depth 144
}

int nsnp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int nsnp_check_status(int32_t* status_dev, void* stream) {
    if (!status_dev) return nsnp::set_error(NSNP_E_INVALID, "nsnp_check_status: null status");
    int32_t h[4] = {0, 0, 0, 0};
    cudaError_t e = cudaMemcpyAsync(h, status_dev, sizeof h, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) return nsnp::set_error(NSNP_E_CUDA, "nsnp_check_status: %s", cudaGetErrorString(e));
    if (h[nsnp::ST_ERR] != nsnp::DEV_OK) {                      // report once: the next call starts from a clean status word
        cudaMemsetAsync(status_dev, 0, sizeof h, (cudaStream_t)stream);
        cudaStreamSynchronize((cudaStream_t)stream);
    }
    switch (h[nsnp::ST_ERR]) {
        case nsnp::DEV_OK: return NSNP_OK;
        case nsnp::DEV_E_DEPTH: return nsnp::set_error(NSNP_E_OVERFLOW, "more than 16383 reads overlap the 1024-bp tile at position %d", h[1]);
        case nsnp::DEV_E_INDEL_SLAB: return nsnp::set_error(NSNP_E_OVERFLOW, "indel event slab overflow in tile %d", h[1]);
        case nsnp::DEV_E_SEQ_SPAN: return nsnp::set_error(NSNP_E_OVERFLOW, "reads over one tile span >= 2^32 bases (tile %d)", h[1]);
        case nsnp::DEV_E_CAND_CAP: return nsnp::set_error(NSNP_E_OVERFLOW, "candidate buffer too small: %d sites", h[1]);
        case nsnp::DEV_E_CIGAR: return nsnp::set_error(NSNP_E_UNSUPPORTED, "read %d: CIGAR is not canonical (P op, or adjacent I I / D D not merged)", h[1]);
        case nsnp::DEV_E_UNSORTED: return nsnp::set_error(NSNP_E_INVALID, "reads are not sorted by position (read %d)", h[1]);
        default: return nsnp::set_error(NSNP_E_CUDA, "unknown device status %d", h[0]);
    }
}

}  // extern "C"
