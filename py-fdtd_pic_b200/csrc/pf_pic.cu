// pf_pic.cu -- electron-beam PIC on the 1-D Yee grid: relativistic Boris push with linear field
// gather, cell sort, and deterministic cell-sorted current deposition into the reference's Jx slot.
//
// The reference contains NO particle code; its only PIC contract is the current
// slot V.Jx that ADE_ExUpdate subtracts (BaseFDTD11.py:667).  The model implemented here is the
// builder-defined spec of DESIGN.md section "PIC", restated on the CPU in oracle/pic_oracle.py:
//
//   particle state : z [m], ux = gamma*vx, uz = gamma*vz [m/s], weight w            (1D2V)
//   fields         : Ex[nz] at z = nz*dz ; By = mu0*Hy, Hy[nz] at z = (nz+1/2)*dz   (Yee staggering)
//   gather         : linear (CIC) interpolation of Ex and Hy to the particle
//   push           : Boris -- half electric kick, magnetic rotation, half electric kick, drift
//   walls          : specular reflection at z = 0 and z = (L-1)*dz
//   deposit        : Jx[nz] = jx_scale * sum_p w_p vx_p S(z_p/dz - nz), S = linear (CIC) shape
//
// Deposition is deterministic: particles are kept sorted by cell (stable radix sort), each cell is
// owned by one warp, lane l accumulates that cell's particles l, l+32, ... in order, the 32 partial
// sums are combined by a fixed xor-butterfly of warp shuffles, and node nz is written once as
// (own-cell left share) + (cell nz-1 right share).  No atomics, no dependence on scheduling.
#include <cub/device/device_radix_sort.cuh>
#include "pf_common.cuh"

namespace pf {

constexpr int PIC_THREADS = 256;

// ------------------------------------------------------------------------------------------------ push
__global__ void __launch_bounds__(PIC_THREADS) k_pic_push(PfPic p)
{
    long long i = (long long)blockIdx.x * PIC_THREADS + threadIdx.x;
    if (i >= p.n) return;
    const double inv_dz = 1.0 / p.dz;
    const double zmax = (double)(p.L - 1) * p.dz;
    double z = p.z[i], ux = p.ux[i], uz = p.uz[i];

    // gather Ex (integer nodes)
    double s = z * inv_dz;
    int c = (int)floor(s);
    c = max(0, min(c, p.L - 2));
    double f = s - (double)c;
    double Ex = (1.0 - f) * p.Ex[c] + f * p.Ex[c + 1];
    // gather Hy (half nodes): Hy[k] sits at (k+1/2) dz
    double sh = s - 0.5;
    int ch = (int)floor(sh);
    ch = max(0, min(ch, p.L - 2));
    double fh = sh - (double)ch;
    fh = fmin(fmax(fh, 0.0), 1.0);
    double By = p.mu0 * ((1.0 - fh) * p.Hy[ch] + fh * p.Hy[ch + 1]);

    const double qmdt2 = p.q_over_m * p.dt * 0.5;
    const double inv_c2 = 1.0 / (p.c * p.c);
    // half electric kick
    double uxm = ux + qmdt2 * Ex;
    double uzm = uz;
    // magnetic rotation about y
    double gm = sqrt(1.0 + (uxm * uxm + uzm * uzm) * inv_c2);
    double t = qmdt2 * By / gm;
    double sfac = 2.0 * t / (1.0 + t * t);
    double uxp = uxm - uzm * t;
    double uzp = uzm + uxm * t;
    double uxn = uxm - uzp * sfac;
    double uzn = uzm + uxp * sfac;
    // half electric kick
    uxn = uxn + qmdt2 * Ex;
    // drift
    double g = sqrt(1.0 + (uxn * uxn + uzn * uzn) * inv_c2);
    z = z + (uzn / g) * p.dt;
    // specular walls
    if (z < 0.0) { z = -z; uzn = -uzn; }
    if (z > zmax) { z = 2.0 * zmax - z; uzn = -uzn; }
    z = fmin(fmax(z, 0.0), zmax);

    p.z[i] = z;
    p.ux[i] = uxn;
    p.uz[i] = uzn;
    int cn = (int)floor(z * inv_dz);
    p.cell[i] = max(0, min(cn, p.L - 2));
}

// ------------------------------------------------------------------------------------------------ sort
__global__ void __launch_bounds__(PIC_THREADS) k_iota(int *idx, long long n)
{
    long long i = (long long)blockIdx.x * PIC_THREADS + threadIdx.x;
    if (i < n) idx[i] = (int)i;
}

__global__ void __launch_bounds__(PIC_THREADS) k_pic_permute(PfPic p, const int *__restrict__ order)
{
    long long i = (long long)blockIdx.x * PIC_THREADS + threadIdx.x;
    if (i >= p.n) return;
    int src = order[i];
    p.z_alt[i] = p.z[src];
    p.ux_alt[i] = p.ux[src];
    p.uz_alt[i] = p.uz[src];
    p.w_alt[i] = p.w[src];
}

static inline size_t al256(size_t x) { return (x + 255) / 256 * 256; }

static int key_bits(int L)
{
    int b = 1;
    while ((1 << b) < L) ++b;
    return b;
}

struct PicPlan {
    size_t off_idx_in, off_idx_out, off_cub, off_acc, cub_bytes, total;
};

static PicPlan pic_plan(const PfPic *p)
{
    PicPlan pl;
    size_t n = (size_t)p->n;
    pl.cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, pl.cub_bytes, (const int *)nullptr, (int *)nullptr, (const int *)nullptr,
                                    (int *)nullptr, (int)n, 0, key_bits(p->L));
    pl.off_idx_in = 0;
    pl.off_idx_out = al256(sizeof(int) * n);
    pl.off_cub = pl.off_idx_out + al256(sizeof(int) * n);
    pl.off_acc = pl.off_cub + al256(pl.cub_bytes);
    pl.total = pl.off_acc + al256(sizeof(double) * 2 * (size_t)p->L);
    return pl;
}

// ------------------------------------------------------------------------------------------------ deposit
__device__ __forceinline__ long long lower_bound_cell(const int *__restrict__ cell, long long n, int key)
{
    long long lo = 0, hi = n;
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        if (cell[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// one warp per cell: acc[2c] = sum w vx (1-f), acc[2c+1] = sum w vx f over the cell's particles
__global__ void __launch_bounds__(PIC_THREADS) k_pic_cell_sums(PfPic p, double *__restrict__ acc)
{
    const int warp = (blockIdx.x * PIC_THREADS + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= p.L) return;
    const int c = warp;
    long long start = 0, end = 0;
    if (lane == 0) {
        start = lower_bound_cell(p.cell, p.n, c);
        end = lower_bound_cell(p.cell, p.n, c + 1);
    }
    start = __shfl_sync(0xffffffffu, start, 0);
    end = __shfl_sync(0xffffffffu, end, 0);
    const double inv_dz = 1.0 / p.dz;
    const double inv_c2 = 1.0 / (p.c * p.c);
    double a0 = 0.0, a1 = 0.0;
    for (long long i = start + lane; i < end; i += 32) {
        double z = p.z[i], ux = p.ux[i], uz = p.uz[i], w = p.w[i];
        double g = sqrt(1.0 + (ux * ux + uz * uz) * inv_c2);
        double wv = w * (ux / g);
        double f = z * inv_dz - (double)c;
        a0 = a0 + wv * (1.0 - f);
        a1 = a1 + wv * f;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        a0 = a0 + __shfl_xor_sync(0xffffffffu, a0, off);
        a1 = a1 + __shfl_xor_sync(0xffffffffu, a1, off);
    }
    if (lane == 0) {
        acc[2 * c] = a0;
        acc[2 * c + 1] = a1;
    }
}

// deterministic flush: node nz = scale * (left share of cell nz + right share of cell nz-1)
__global__ void __launch_bounds__(PIC_THREADS) k_pic_flush(PfPic p, const double *__restrict__ acc)
{
    int nz = blockIdx.x * PIC_THREADS + threadIdx.x;
    if (nz >= p.L) return;
    double v = acc[2 * nz];
    if (nz > 0) v = v + acc[2 * (nz - 1) + 1];
    p.Jx[nz] = p.jx_scale * v;
}

static int validate_pic(const PfPic *p)
{
    if (!p) return set_err(PF_E_ARG, "null PfPic");
    if (p->n < 0 || p->n > 2000000000LL) return set_err(PF_E_ARG, "particle count out of range");
    if (p->L < 3) return set_err(PF_E_ARG, "PIC grid too small");
    if (p->n > 0 && (!p->z || !p->ux || !p->uz || !p->w || !p->cell)) return set_err(PF_E_ARG, "particle arrays missing");
    return 0;
}

}  // namespace pf

using namespace pf;

extern "C" {

size_t pf_pic_scratch_bytes(const PfPic *p)
{
    if (!p) return 256;
    return pic_plan(p).total;   // n = 0 still needs the per-cell accumulators of the deposit
}

int pf_pic_push(const PfPic *p, void *stream)
{
    int rc = validate_pic(p);
    if (rc) return rc;
    if (!p->Ex || !p->Hy) return set_err(PF_E_ARG, "pf_pic_push: field arrays missing");
    if (p->n == 0) return PF_OK;
    unsigned blocks = (unsigned)((p->n + PIC_THREADS - 1) / PIC_THREADS);
    k_pic_push<<<blocks, PIC_THREADS, 0, (cudaStream_t)stream>>>(*p);
    PF_LAUNCH_CHECK("k_pic_push");
    return PF_OK;
}

int pf_pic_sort(const PfPic *p, void *scratch, size_t scratch_bytes, void *stream)
{
    int rc = validate_pic(p);
    if (rc) return rc;
    if (p->n == 0) return PF_OK;
    if (!p->z_alt || !p->ux_alt || !p->uz_alt || !p->w_alt || !p->cell_alt)
        return set_err(PF_E_ARG, "pf_pic_sort: alternate (output) arrays missing");
    PicPlan pl = pic_plan(p);
    if (!scratch || scratch_bytes < pl.total) return set_err(PF_E_SCRATCH, "pf_pic_sort needs %zu bytes of scratch", pl.total);
    cudaStream_t st = (cudaStream_t)stream;
    char *s = (char *)scratch;
    int *idx_in = (int *)(s + pl.off_idx_in), *idx_out = (int *)(s + pl.off_idx_out);
    unsigned blocks = (unsigned)((p->n + PIC_THREADS - 1) / PIC_THREADS);
    k_iota<<<blocks, PIC_THREADS, 0, st>>>(idx_in, p->n);
    PF_LAUNCH_CHECK("k_iota");
    size_t cb = pl.cub_bytes;
    PF_CUDA(cub::DeviceRadixSort::SortPairs(s + pl.off_cub, cb, (const int *)p->cell, p->cell_alt, (const int *)idx_in,
                                            idx_out, (int)p->n, 0, key_bits(p->L), st));
    pf::g_launches += 3;  // CUB's histogram + onesweep passes
    k_pic_permute<<<blocks, PIC_THREADS, 0, st>>>(*p, idx_out);
    PF_LAUNCH_CHECK("k_pic_permute");
    return PF_OK;
}

int pf_pic_deposit(const PfPic *p, void *scratch, size_t scratch_bytes, void *stream)
{
    int rc = validate_pic(p);
    if (rc) return rc;
    if (!p->Jx) return set_err(PF_E_ARG, "pf_pic_deposit: Jx missing");
    PicPlan pl = pic_plan(p);
    if (!scratch || scratch_bytes < pl.total) return set_err(PF_E_SCRATCH, "pf_pic_deposit needs %zu bytes of scratch", pl.total);
    cudaStream_t st = (cudaStream_t)stream;
    double *acc = (double *)((char *)scratch + pl.off_acc);
    unsigned blocks = (unsigned)(((long long)p->L * 32 + PIC_THREADS - 1) / PIC_THREADS);
    k_pic_cell_sums<<<blocks, PIC_THREADS, 0, st>>>(*p, acc);
    PF_LAUNCH_CHECK("k_pic_cell_sums");
    k_pic_flush<<<(p->L + PIC_THREADS - 1) / PIC_THREADS, PIC_THREADS, 0, st>>>(*p, acc);
    PF_LAUNCH_CHECK("k_pic_flush");
    return PF_OK;
}

}  // extern "C"
