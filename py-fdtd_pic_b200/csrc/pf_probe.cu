// pf_probe.cu -- measurement support: per-kernel launch timing with CUDA events (bench.py's roofline numerators) and a
// probe of what the FP64 pipe of this GPU delivers (the roofline denominator of the temporally blocked tile kernel, which
// is bound by that pipe and not by HBM).  Nothing here is on the hot path.
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include "pf_common.cuh"

namespace pf {

static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mutex;   // guards g_prof_events (launches may come from one host thread per GPU)
struct ProfEvent {
    const char *name;
    cudaEvent_t a, b;
};
static std::vector<ProfEvent> g_prof_events;

ProfScope::ProfScope(cudaStream_t s, const char *kernel_name) : st(s), name(kernel_name)
{
    if (g_prof_on.load(std::memory_order_relaxed) && cudaEventCreate(&a) == cudaSuccess && cudaEventCreate(&b) == cudaSuccess)
        cudaEventRecord(a, st);
}
ProfScope::~ProfScope()
{
    if (a && b) {
        cudaEventRecord(b, st);
        std::lock_guard<std::mutex> lk(g_prof_mutex);
        g_prof_events.push_back(ProfEvent{name, a, b});
    }
}

// independent DMUL + DADD streams, 8 per thread: the separately rounded fp64 instruction rate the exact-arithmetic kernels
// are limited by (FMA = true: the DFMA rate, for the contracted modes)
template <bool FMA>
__global__ void __launch_bounds__(1024) k_fp64_stream(double *out, int iters, double a, double b)
{
    double v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = FMA ? __fma_rn(v[i], a, b) : __dadd_rn(__dmul_rn(v[i], a), b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace pf

using namespace pf;

extern "C" {

int pf_profile_enable(int on)
{
    g_prof_on = on != 0;
    return PF_OK;
}

// aggregates and clears the event list; report (may be NULL): "name|launches|ms;..." per kernel name
static int profile_drain(double *ms_total, int *n_launches, char *report, size_t report_bytes)
{
    std::lock_guard<std::mutex> lk(g_prof_mutex);
    std::map<std::string, std::pair<int, double>> agg;
    double tot = 0.0;
    int n = 0;
    for (auto &ev : g_prof_events) {
        float ms = 0.f;
        PF_CUDA(cudaEventSynchronize(ev.b));
        PF_CUDA(cudaEventElapsedTime(&ms, ev.a, ev.b));
        tot += ms;
        ++n;
        auto &e = agg[ev.name ? ev.name : "?"];
        e.first += 1;
        e.second += ms;
        cudaEventDestroy(ev.a);
        cudaEventDestroy(ev.b);
    }
    g_prof_events.clear();
    if (ms_total) *ms_total = tot;
    if (n_launches) *n_launches = n;
    if (report && report_bytes) {
        std::string r;
        for (auto &kv : agg) {
            char line[160];
            snprintf(line, sizeof(line), "%s|%d|%.6f;", kv.first.c_str(), kv.second.first, kv.second.second);
            r += line;
        }
        snprintf(report, report_bytes, "%s", r.c_str());
    }
    return PF_OK;
}

int pf_profile_collect(double *ms_total, int *n_launches) { return profile_drain(ms_total, n_launches, nullptr, 0); }
int pf_profile_report(char *report, size_t report_bytes)
{
    if (!report || !report_bytes) return set_err(PF_E_ARG, "pf_profile_report: no buffer");
    return profile_drain(nullptr, nullptr, report, report_bytes);
}

int pf_probe_fp64(double *dmul_dadd_instr_per_s, double *dfma_per_s, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    PF_CUDA(cudaGetDevice(&dev));
    PF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int threads = 1024, blocks = sms * 2, iters = 4096;
    double *out = nullptr;
    PF_CUDA(cudaMalloc(&out, sizeof(double) * (size_t)blocks * threads));
    cudaEvent_t a, b;
    PF_CUDA(cudaEventCreate(&a));
    PF_CUDA(cudaEventCreate(&b));
    double res[2] = {0.0, 0.0};
    for (int fma = 0; fma < 2; ++fma) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {   // rep 0 warms up
            PF_CUDA(cudaEventRecord(a, st));
            if (fma) k_fp64_stream<true><<<blocks, threads, 0, st>>>(out, iters, 1.0000001, 1e-9);
            else k_fp64_stream<false><<<blocks, threads, 0, st>>>(out, iters, 1.0000001, 1e-9);
            PF_LAUNCH_CHECK("k_fp64_stream");
            PF_CUDA(cudaEventRecord(b, st));
            PF_CUDA(cudaEventSynchronize(b));
            float ms = 0.f;
            PF_CUDA(cudaEventElapsedTime(&ms, a, b));
            if (rep > 0 && ms < best) best = ms;
        }
        const double per_thread = (double)iters * 8 * (fma ? 1 : 2);   // instructions
        res[fma] = per_thread * blocks * threads / (best * 1e-3);
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(out);
    if (dmul_dadd_instr_per_s) *dmul_dadd_instr_per_s = res[0];
    if (dfma_per_s) *dfma_per_s = res[1];
    return PF_OK;
}

}  // extern "C"
