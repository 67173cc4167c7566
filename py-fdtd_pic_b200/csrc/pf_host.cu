// pf_host.cu -- host-side plumbing of libpyfdtd_b200: error reporting, device info, host-derived
// constants.  No kernels here.
#include <cmath>
#include <cstdarg>
#include <cstring>
#include "pf_common.cuh"

namespace pf {

thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

int set_err(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_cuda(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return 0;
    return set_err(PF_E_CUDA, "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
}

// CubicEquationSolver.findF / findG / findH (CubicEquationSolver.py:94-105) for the terms that
// depend on (a,b,c) only.  CPython evaluates `x ** 2.0` / `x ** 3.0` with libm pow(), so this does too.
CubicConsts cubic_consts_host(double a, double b, double c)
{
    CubicConsts k;
    memset(&k, 0, sizeof(k));
    k.a = a; k.b = b; k.c = c;
    if (a == 0.0) return k;
    k.inv_a = 1.0 / a;
    k.f = ((3.0 * c / a) - (std::pow(b, 2.0) / std::pow(a, 2.0))) / 3.0;
    k.g_ab = ((2.0 * std::pow(b, 3.0)) / std::pow(a, 3.0)) - ((9.0 * b * c) / std::pow(a, 2.0));
    k.f3_27 = std::pow(k.f, 3.0) / 27.0;
    k.b_3a = b / (3.0 * a);
    k.inv_c = (c != 0.0) ? 1.0 / c : 0.0;
    return k;
}

GridDev make_grid_dev(const PfGrid &g)
{
    GridDev d;
    d.g = g;
    d.k = cubic_consts_host(g.cub_a, g.cub_b, g.cub_c);
    d.k.newton = (g.flags & PF_F_NEWTON) && g.cub_a >= 0.0 && g.cub_b >= 0.0 && g.cub_c > 0.0;
    d.inv_eps0 = 1.0 / g.eps0;
    return d;
}

}  // namespace pf

extern "C" {

int pf_abi_version(void) { return PF_ABI_VERSION; }
const char *pf_last_error(void) { return pf::g_err; }
unsigned long long pf_launch_count(void) { return pf::g_launches.load(); }

int pf_host_exp(const double *x, double *y, long long n)
{
    if (n < 0 || (n > 0 && (!x || !y))) return pf::set_err(PF_E_ARG, "pf_host_exp: bad arguments");
    for (long long i = 0; i < n; ++i) y[i] = exp(x[i]);
    return PF_OK;
}

int pf_host_pow(const double *x, double e, double *y, long long n)
{
    if (n < 0 || (n > 0 && (!x || !y))) return pf::set_err(PF_E_ARG, "pf_host_pow: bad arguments");
    for (long long i = 0; i < n; ++i) y[i] = pow(x[i], e);
    return PF_OK;
}

int pf_device_info(int *sm_count, int *cc_major, int *cc_minor, size_t *free_bytes, size_t *total_bytes)
{
    int dev = 0;
    PF_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    PF_CUDA(cudaGetDeviceProperties(&p, dev));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    size_t f = 0, t = 0;
    PF_CUDA(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return PF_OK;
}

int pf_sync(void *stream)
{
    PF_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return PF_OK;
}

}  // extern "C"
