// pf_ops.cu -- ENGINE_OPS: one streaming kernel per reference leaf op, general per-cell coefficient
// arrays, state in HBM.  This is the drop-in for the individual BaseFDTD11 functions and the
// fallback integrator for grids whose coefficient arrays are not in the tile engine's canonical
// form.  One thread per cell, coalesced fp64 loads; every kernel is a pure streaming sweep.
#include "pf_common.cuh"

namespace pf {

constexpr int OPS_THREADS = 256;
static inline int ops_blocks(int n) { return (n + OPS_THREADS - 1) / OPS_THREADS; }

// local cell index -> global cell index of the undecomposed grid
#define PF_GZ(g, nz) ((long long)(g).z0 + (nz))

// BaseFDTD11.py:663-669  ADE_ExUpdate: cells 1..Lg-1
template <class A>
__global__ void __launch_bounds__(OPS_THREADS) k_ex_update(PfGrid g)
{
    int nz = blockIdx.x * OPS_THREADS + threadIdx.x;
    if (nz < 1 || nz >= g.L) return;
    long long gz = PF_GZ(g, nz);
    if (gz < 1 || gz > g.Lg - 1) return;
    double dH = A::sub(g.Hy[nz], g.Hy[nz - 1]);
    if (g.Jx) dH = A::sub(dH, g.Jx[nz]);
    g.Ex[nz] = A::add(g.Ex[nz], A::mul(A::mul(dH, g.UpExMat[nz]), g.denE[nz]));
}

// BaseFDTD11.py:640-656  ADE_HyUpdate: cells 1..Nz-1 = 1..Lg-2
template <class A>
__global__ void __launch_bounds__(OPS_THREADS) k_hy_update(PfGrid g)
{
    int nz = blockIdx.x * OPS_THREADS + threadIdx.x;
    if (nz >= g.L - 1) return;
    long long gz = PF_GZ(g, nz);
    if (gz < 1 || gz > g.Lg - 2) return;
    double dE = A::sub(g.Ex[nz + 1], g.Ex[nz]);
    g.Hy[nz] = A::add(A::mul(g.Hy[nz], g.UpHySelf[nz]), A::mul(A::mul(dE, g.UpHyMat[nz]), g.denH[nz]));
}

__device__ __forceinline__ bool in_pml_e(const PfGrid &g, long long gz)
{
    return ((g.flags & PF_F_CPML_M) && gz >= 1 && gz < g.pw) ||
           ((g.flags & PF_F_CPML_P) && gz >= g.Lg - g.pw && gz < g.Lg);
}
__device__ __forceinline__ bool in_pml_h(const PfGrid &g, long long gz)
{
    return ((g.flags & PF_F_CPML_M) && gz >= 1 && gz < g.pw) ||
           ((g.flags & PF_F_CPML_P) && gz >= g.Lg - g.pw && gz < g.Lg - 1);
}

// BaseFDTD11.py:364-376  CPML_Psi_e_Update
template <class A>
__global__ void __launch_bounds__(OPS_THREADS) k_psi_e(PfGrid g)
{
    int nz = blockIdx.x * OPS_THREADS + threadIdx.x;
    if (nz < 1 || nz >= g.L) return;
    if (!in_pml_e(g, PF_GZ(g, nz))) return;
    double dH = A::sub(g.Hy[nz], g.Hy[nz - 1]);
    double psi = A::add(A::mul(g.beX[nz], g.psiE[nz]), A::mul(g.ceX[nz], dH));
    g.psiE[nz] = psi;
    g.Ex[nz] = A::sub(g.Ex[nz], A::mul(g.Cb[nz], psi));
}

// BaseFDTD11.py:381-393  CPML_Psi_m_Update
template <class A>
__global__ void __launch_bounds__(OPS_THREADS) k_psi_m(PfGrid g)
{
    int nz = blockIdx.x * OPS_THREADS + threadIdx.x;
    if (nz >= g.L - 1) return;
    if (!in_pml_h(g, PF_GZ(g, nz))) return;
    double dE = A::sub(g.Ex[nz + 1], g.Ex[nz]);
    double psi = A::add(A::mul(g.bmY[nz], g.psiH[nz]), A::mul(g.cmY[nz], dE));
    g.psiH[nz] = psi;
    g.Hy[nz] = A::add(g.Hy[nz], A::mul(g.C2[nz], psi));
}

// Solver_Engine.py:307-310: Ex[nzsrc] += Exs[n]/courantNo ; Hy[nzsrc-1] -= Hys[n]/courantNo (TF/SF)
__global__ void k_source(PfGrid g, int n)
{
    long long ls = (long long)g.nzsrc - g.z0;  // local index of the source cell
    if (threadIdx.x == 0 && ls >= 0 && ls < g.L) g.Ex[ls] = __dadd_rn(g.Ex[ls], g.srcE[n]);
    if (threadIdx.x == 1 && (g.flags & PF_F_TFSF) && ls - 1 >= 0 && ls - 1 < g.L)
        g.Hy[ls - 1] = __dsub_rn(g.Hy[ls - 1], g.srcH[n]);
}

__device__ __forceinline__ bool in_slab(const PfGrid &g, long long gz) { return gz >= g.mf && gz < g.mr; }

// BaseFDTD11.py:750-760  ADE_DxUpdate
template <class A>
__global__ void __launch_bounds__(OPS_THREADS) k_dx_update(PfGrid g)
{
    int nz = blockIdx.x * OPS_THREADS + threadIdx.x;
    if (nz < 1 || nz >= g.L) return;
    if (!in_slab(g, PF_GZ(g, nz))) return;
    double dH = A::sub(g.Hy[nz], g.Hy[nz - 1]);
    // current slot in material cells (builder-defined PIC coupling, dD/dt = curl H - J: the same bracket as ADE_ExUpdate;
    // see include/pyfdtd_b200.h).  Jx == NULL in every reference run.
    if (g.Jx) dH = A::sub(dH, g.Jx[nz]);
    g.Dx[nz] = A::add(g.Dx[nz], A::mul(A::mul(dH, g.dt_over_dz), g.denE[nz]));
}

// BaseFDTD11.py:487-538 history shift fused with :609-633 ADE_PolarisationCurrent_Ex:
// P^{n+1} = A P^n + B P^{n-1} + C E^n ; the P^{n-1} slot is rotated in place (no list copies).
template <class A>
__global__ void __launch_bounds__(OPS_THREADS) k_pol_update(PfGrid g)
{
    int nz = blockIdx.x * OPS_THREADS + threadIdx.x;
    if (nz >= g.L) return;
    if (!in_slab(g, PF_GZ(g, nz))) return;
    double pn = g.P[nz];
    g.P[nz] = A::add(A::add(A::mul(g.polA, pn), A::mul(g.polB, g.Pprev[nz])), A::mul(g.polC, g.Ex[nz]));
    g.Pprev[nz] = pn;
}

// BaseFDTD11.py:712-725  ADE_ExCreate
template <class A>
__global__ void __launch_bounds__(OPS_THREADS) k_ex_create(PfGrid g, double inv_eps0)
{
    int nz = blockIdx.x * OPS_THREADS + threadIdx.x;
    if (nz >= g.L) return;
    if (!in_slab(g, PF_GZ(g, nz))) return;
    g.Ex[nz] = div_const(A::sub(g.Dx[nz], g.P[nz]), g.eps0, inv_eps0);
}

// BaseFDTD11.py:793-853  AcubicFinder (Nonlin_Eqn_Setup + Nonlin_Cubic_Solver)
__global__ void __launch_bounds__(OPS_THREADS) k_acubic(PfGrid g, CubicConsts k, double inv_eps0)
{
    int nz = blockIdx.x * OPS_THREADS + threadIdx.x;
    if (nz >= g.L) return;
    if (!in_slab(g, PF_GZ(g, nz))) return;
    g.Acubic[nz] = acubic_cell(k, g.Dx[nz], g.eps0, inv_eps0);
}

// PF_LORENTZ_NL: the same two leaf ops on Dn = Dx - P (see include/pyfdtd_b200.h)
__global__ void __launch_bounds__(OPS_THREADS) k_acubic_dn(PfGrid g, CubicConsts k, double inv_eps0)
{
    int nz = blockIdx.x * OPS_THREADS + threadIdx.x;
    if (nz >= g.L) return;
    if (!in_slab(g, PF_GZ(g, nz))) return;
    g.Acubic[nz] = acubic_cell(k, __dsub_rn(g.Dx[nz], g.P[nz]), g.eps0, inv_eps0);
}

template <class A>
__global__ void __launch_bounds__(OPS_THREADS) k_nl_ex_dn(PfGrid g)
{
    int nz = blockIdx.x * OPS_THREADS + threadIdx.x;
    if (nz >= g.L) return;
    if (!in_slab(g, PF_GZ(g, nz))) return;
    g.Ex[nz] = __ddiv_rn(A::sub(g.Dx[nz], g.P[nz]), A::add(g.nl_den0, A::mul(g.nl_den1, g.Acubic[nz])));
}

// BaseFDTD11.py:858-877  NonLinExUpdate
template <class A>
__global__ void __launch_bounds__(OPS_THREADS) k_nl_ex(PfGrid g)
{
    int nz = blockIdx.x * OPS_THREADS + threadIdx.x;
    if (nz >= g.L) return;
    if (!in_slab(g, PF_GZ(g, nz))) return;
    g.Ex[nz] = __ddiv_rn(g.Dx[nz], A::add(g.nl_den0, A::mul(g.nl_den1, g.Acubic[nz])));
}

// Solver_Engine.py:16-54  probeSim: probe_out[p][n] = Ex[probe_idx[p]]
__global__ void k_probe(PfGrid g, int n)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g.n_probes) return;
    long long lp = (long long)g.probe_idx[p] - g.z0;
    if (lp >= 0 && lp < g.L) g.probe_out[(size_t)p * g.probe_stride + n] = g.Ex[lp];
}

__global__ void __launch_bounds__(OPS_THREADS) k_cubic_root0(const double *__restrict__ co, double *__restrict__ out, int n,
                                                            int newton)
{
    int i = blockIdx.x * OPS_THREADS + threadIdx.x;
    if (i >= n) return;
    double a = co[4 * i], b = co[4 * i + 1], c = co[4 * i + 2], d = co[4 * i + 3];
    CubicConsts k = cubic_consts_dev(a, b, c);
    // PF_F_NEWTON only where its monotone-convergence conditions hold; the closed form otherwise
    if (newton && a >= 0.0 && b >= 0.0 && c > 0.0 && d < 0.0) out[i] = cubic_root0_newton(k, -d);
    else out[i] = cubic_root0(k, d);
}

// All roots of a x^3 + b x^2 + c x + d as CubicEquationSolver.solve returns them
// (CubicEquationSolver.py:29-90): roots[i][r] = (re, im), nroots[i] in {1,2,3}.
__global__ void __launch_bounds__(OPS_THREADS) k_cubic_solve(const double *__restrict__ co, double *__restrict__ roots,
                                                            int *__restrict__ nroots, int n)
{
    int i = blockIdx.x * OPS_THREADS + threadIdx.x;
    if (i >= n) return;
    const double a = co[4 * i], b = co[4 * i + 1], c = co[4 * i + 2], d = co[4 * i + 3];
    double re[3] = {0, 0, 0}, im[3] = {0, 0, 0};
    int nr = 3;
    if (a == 0.0 && b == 0.0) {                      // linear
        re[0] = __ddiv_rn(__dmul_rn(-d, 1.0), c);
        nr = 1;
    } else if (a == 0.0) {                           // quadratic
        double D = __dsub_rn(__dmul_rn(c, c), __dmul_rn(__dmul_rn(4.0, b), d));
        const double twob = __dmul_rn(2.0, b);
        if (D >= 0.0) {
            D = sqrt(D);
            re[0] = __ddiv_rn(__dadd_rn(-c, D), twob);
            re[1] = __ddiv_rn(__dsub_rn(-c, D), twob);
        } else {
            D = sqrt(-D);
            re[0] = __ddiv_rn(-c, twob); im[0] = __ddiv_rn(D, twob);
            re[1] = __ddiv_rn(-c, twob); im[1] = __ddiv_rn(-D, twob);
        }
        nr = 2;
    } else {
        const CubicConsts k = cubic_consts_dev(a, b, c);
        const double g = __ddiv_rn(__dadd_rn(k.g_ab, __ddiv_rn(__dmul_rn(27.0, d), a)), 27.0);
        const double gg4 = __dmul_rn(__dmul_rn(g, g), 0.25);
        const double h = __dadd_rn(gg4, k.f3_27);
        const double ghalf = __dmul_rn(g, 0.5);
        if (k.f == 0.0 && g == 0.0 && h == 0.0) {    // three equal real roots
            const double da = __ddiv_rn(d, a);
            const double x = (da >= 0.0) ? -pow_third(da) : pow_third(-da);
            re[0] = re[1] = re[2] = x;
        } else if (h <= 0.0) {                       // three real roots
            const double ii = sqrt(__dsub_rn(gg4, h));
            const double j = pow_third(ii);
            const double kk = acos(-__ddiv_rn(g, __dmul_rn(2.0, ii)));
            const double Lm = -j, M = cos(__ddiv_rn(kk, 3.0)), N = __dmul_rn(sqrt(3.0), sin(__ddiv_rn(kk, 3.0)));
            const double P = -k.b_3a;
            re[0] = __dsub_rn(__dmul_rn(__dmul_rn(2.0, j), M), k.b_3a);
            re[1] = __dadd_rn(__dmul_rn(Lm, __dadd_rn(M, N)), P);
            re[2] = __dadd_rn(__dmul_rn(Lm, __dsub_rn(M, N)), P);
        } else {                                     // one real root, two complex
            const double sh = sqrt(h);
            const double S = signed_cbrt_pow(__dadd_rn(-ghalf, sh));
            const double U = signed_cbrt_pow(__dsub_rn(-ghalf, sh));
            const double SU = __dadd_rn(S, U);
            re[0] = __dsub_rn(SU, k.b_3a);
            const double rr = __dsub_rn(__ddiv_rn(-SU, 2.0), k.b_3a);
            const double ii2 = __dmul_rn(__dmul_rn(__dsub_rn(S, U), sqrt(3.0)), 0.5);
            re[1] = rr; im[1] = ii2;
            re[2] = rr; im[2] = -ii2;
        }
    }
    for (int r = 0; r < 3; ++r) {
        roots[6 * i + 2 * r] = re[r];
        roots[6 * i + 2 * r + 1] = im[r];
    }
    nroots[i] = nr;
}

static int validate(const PfGrid *g)
{
    if (!g) return set_err(PF_E_ARG, "null grid");
    if (g->L < 3) return set_err(PF_E_ARG, "grid too small (L=%d)", g->L);
    if (!g->Ex || !g->Hy) return set_err(PF_E_ARG, "Ex/Hy missing");
    return 0;
}

#define PF_DISPATCH(kern, g, st, ...)                                                        \
    do {                                                                                     \
        if ((g)->flags & PF_F_FMA)                                                           \
            kern<Fused><<<ops_blocks((g)->L), OPS_THREADS, 0, st>>>(__VA_ARGS__);            \
        else                                                                                 \
            kern<Exact><<<ops_blocks((g)->L), OPS_THREADS, 0, st>>>(__VA_ARGS__);            \
        PF_LAUNCH_CHECK(#kern);                                                              \
    } while (0)

int ops_step(const PfGrid *g, const GridDev &gd, int mode, int do_pol, int n, cudaStream_t st)
{
    bool cpml = g->flags & (PF_F_CPML_M | PF_F_CPML_P);
    if ((mode == PF_LORENTZ || mode == PF_LORENTZ_NL) && do_pol) PF_DISPATCH(k_pol_update, g, st, *g);
    PF_DISPATCH(k_ex_update, g, st, *g);
    if (cpml) PF_DISPATCH(k_psi_e, g, st, *g);
    k_source<<<1, 32, 0, st>>>(*g, n);
    PF_LAUNCH_CHECK("k_source");
    if (mode == PF_LORENTZ) {
        PF_DISPATCH(k_dx_update, g, st, *g);
        PF_DISPATCH(k_ex_create, g, st, *g, gd.inv_eps0);
    } else if (mode == PF_NL) {
        PF_DISPATCH(k_dx_update, g, st, *g);
        k_acubic<<<ops_blocks(g->L), OPS_THREADS, 0, st>>>(*g, gd.k, gd.inv_eps0);
        PF_LAUNCH_CHECK("k_acubic");
        PF_DISPATCH(k_nl_ex, g, st, *g);
    } else if (mode == PF_LORENTZ_NL) {
        PF_DISPATCH(k_dx_update, g, st, *g);
        k_acubic_dn<<<ops_blocks(g->L), OPS_THREADS, 0, st>>>(*g, gd.k, gd.inv_eps0);
        PF_LAUNCH_CHECK("k_acubic_dn");
        PF_DISPATCH(k_nl_ex_dn, g, st, *g);
    }
    PF_DISPATCH(k_hy_update, g, st, *g);
    if (cpml) PF_DISPATCH(k_psi_m, g, st, *g);
    if (g->n_probes > 0) {
        k_probe<<<(g->n_probes + 31) / 32, 32, 0, st>>>(*g, n);
        PF_LAUNCH_CHECK("k_probe");
    }
    return 0;
}

int ops_run_pass(const PfGrid *g, int mode, int do_pol, int n0, int nsteps, double *snap_out,
                 int snap_interval, int snap_rows, cudaStream_t st)
{
    GridDev gd = make_grid_dev(*g);
    // PF_LORENTZ_NL is specified with the converged (Newton) root wherever the coefficients admit it
    if (mode == PF_LORENTZ_NL && g->cub_a >= 0.0 && g->cub_b >= 0.0 && g->cub_c > 0.0) gd.k.newton = 1;
    for (int n = n0; n < n0 + nsteps; ++n) {
        int rc = ops_step(g, gd, mode, do_pol, n, st);
        if (rc) return rc;
        if (snap_out && snap_interval > 0 && n > 0 && n % snap_interval == 0) {
            int row = n / snap_interval;
            if (row < snap_rows)
                PF_CUDA(cudaMemcpyAsync(snap_out + (size_t)row * g->L, g->Ex, sizeof(double) * g->L,
                                        cudaMemcpyDeviceToDevice, st));
        }
    }
    return 0;
}

}  // namespace pf

using namespace pf;

extern "C" {

#define PF_LEAF(name, kern)                                    \
    int name(const PfGrid *g, void *stream)                    \
    {                                                          \
        int rc = validate(g);                                  \
        if (rc) return rc;                                     \
        cudaStream_t st = (cudaStream_t)stream;                \
        PF_DISPATCH(kern, g, st, *g);                          \
        return PF_OK;                                          \
    }

PF_LEAF(pf_ade_ex_update, k_ex_update)
PF_LEAF(pf_ade_hy_update, k_hy_update)
PF_LEAF(pf_cpml_psi_e_update, k_psi_e)
PF_LEAF(pf_cpml_psi_m_update, k_psi_m)
PF_LEAF(pf_ade_dx_update, k_dx_update)
PF_LEAF(pf_ade_polarisation_update, k_pol_update)
PF_LEAF(pf_nonlin_ex_update, k_nl_ex)

int pf_ade_ex_create(const PfGrid *g, void *stream)
{
    int rc = validate(g);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    PF_DISPATCH(k_ex_create, g, st, *g, 1.0 / g->eps0);
    return PF_OK;
}

int pf_acubic_finder(const PfGrid *g, void *stream)
{
    int rc = validate(g);
    if (rc) return rc;
    GridDev gd = make_grid_dev(*g);
    k_acubic<<<ops_blocks(g->L), OPS_THREADS, 0, (cudaStream_t)stream>>>(*g, gd.k, gd.inv_eps0);
    PF_LAUNCH_CHECK("k_acubic");
    return PF_OK;
}

int pf_source_inject(const PfGrid *g, int n, void *stream)
{
    int rc = validate(g);
    if (rc) return rc;
    k_source<<<1, 32, 0, (cudaStream_t)stream>>>(*g, n);
    PF_LAUNCH_CHECK("k_source");
    return PF_OK;
}

int pf_probe_record(const PfGrid *g, int n, void *stream)
{
    int rc = validate(g);
    if (rc) return rc;
    if (g->n_probes <= 0) return PF_OK;
    k_probe<<<(g->n_probes + 31) / 32, 32, 0, (cudaStream_t)stream>>>(*g, n);
    PF_LAUNCH_CHECK("k_probe");
    return PF_OK;
}

int pf_cubic_solve(const double *coeffs, double *roots, int *nroots, int n, void *stream)
{
    if (!coeffs || !roots || !nroots || n < 0) return set_err(PF_E_ARG, "pf_cubic_solve: bad arguments");
    if (n == 0) return PF_OK;
    k_cubic_solve<<<ops_blocks(n), OPS_THREADS, 0, (cudaStream_t)stream>>>(coeffs, roots, nroots, n);
    PF_LAUNCH_CHECK("k_cubic_solve");
    return PF_OK;
}

int pf_cubic_root0(const double *coeffs, double *root0, int n, void *stream)
{
    if (!coeffs || !root0 || n < 0) return set_err(PF_E_ARG, "pf_cubic_root0: bad arguments");
    if (n == 0) return PF_OK;
    k_cubic_root0<<<ops_blocks(n), OPS_THREADS, 0, (cudaStream_t)stream>>>(coeffs, root0, n, 0);
    PF_LAUNCH_CHECK("k_cubic_root0");
    return PF_OK;
}

int pf_cubic_root0_newton(const double *coeffs, double *root0, int n, void *stream)
{
    if (!coeffs || !root0 || n < 0) return set_err(PF_E_ARG, "pf_cubic_root0_newton: bad arguments");
    if (n == 0) return PF_OK;
    k_cubic_root0<<<ops_blocks(n), OPS_THREADS, 0, (cudaStream_t)stream>>>(coeffs, root0, n, 1);
    PF_LAUNCH_CHECK("k_cubic_root0");
    return PF_OK;
}

}  // extern "C"
