// pf_dormant.cu -- the reference's DORMANT material / boundary models: leaf functions that exist in
// BaseFDTD11.py but that no integrator of the reference calls, and the Drude current-form loop of its scratch script.
// One streaming kernel per reference leaf function, one thread per cell, every multiply and add separately rounded in the
// reference's order (Exact policy), so results are bit-identical to the Python / numba originals.
//
//   ADE_NonLin_Pol_Ex_Pbar       BaseFDTD11.py:567-577   Pbar3 = eps0 (chi1 E + chi3 (alpha3 E^3 + (1-alpha3) Q E))     (Varin)
//   ADE_Lin_Curr_And_Pol_Varin   BaseFDTD11.py:580-594   J = (A/D) J + (B/D)(Pbar3 - P);  P = P + dt J                   (Varin)
//   ADE_Nonlin_Q_and_G           BaseFDTD11.py:596-609   G = (e/f) G + (h/f)(E^2 - Q);    Q = Q + dt G                   (Raman)
//   KerrNonlin                   BaseFDTD11.py:762-766   JxKerr = (alpha3 eps0 chi3 / dt)(|E|^2 E - |Eold|^2 Eold)
//   MUR1DEx                      BaseFDTD11.py:769-788   first-order Mur ABC on cells 1..4 and Nz-1..Nz-5
//   TESTBOXDIPSERSE.py:79-94     Drude medium in current (J) form, hard source, no PML -- the whole loop as written
#include "pf_common.cuh"

namespace pf {

constexpr int DT = 256;
static inline int dblocks(long long n) { return (int)((n + DT - 1) / DT); }

using A = Exact;

__global__ void __launch_bounds__(DT) k_varin_pbar(PfDormant d)
{
    const int nz = d.mf + blockIdx.x * DT + threadIdx.x;
    if (nz >= d.mr) return;
    const double e = d.Ex[nz];
    // P.permit_0 * (chi1*E + chi3*(alpha3*E*E*E + (1-alpha3)*Q*E)), left to right
    const double cub = A::mul(A::mul(A::mul(d.alpha3, e), e), e);
    const double ram = A::mul(A::mul(d.one_minus_alpha3, d.Qx3[nz]), e);
    d.Pbar3[nz] = A::mul(d.eps0, A::add(A::mul(d.chi1, e), A::mul(d.chi3, A::add(cub, ram))));
}

__global__ void __launch_bounds__(DT) k_varin_lin(PfDormant d)
{
    const int nz = d.mf + blockIdx.x * DT + threadIdx.x;
    if (nz >= d.mr) return;
    const double j = A::add(A::mul(d.lin_AoverD, d.Jx[nz]), A::mul(d.lin_BoverD, A::sub(d.Pbar3[nz], d.P[nz])));
    d.Jx[nz] = j;
    d.P[nz] = A::add(d.P[nz], A::mul(d.dt, j));
}

__global__ void __launch_bounds__(DT) k_varin_qg(PfDormant d)
{
    const int nz = d.mf + blockIdx.x * DT + threadIdx.x;
    if (nz >= d.mr) return;
    const double e = d.Ex[nz];
    const double g = A::add(A::mul(d.ram_eoverf, d.Gx3[nz]), A::mul(d.ram_hoverf, A::sub(A::mul(e, e), d.Qx3[nz])));
    d.Gx3[nz] = g;
    d.Qx3[nz] = A::add(d.Qx3[nz], A::mul(d.dt, g));
}

__global__ void __launch_bounds__(DT) k_kerr_nonlin(PfDormant d)
{
    const int nz = blockIdx.x * DT + threadIdx.x;
    if (nz >= d.L) return;
    const double e = d.Ex[nz], o = d.Eold[nz];
    const double ae = fabs(e), ao = fabs(o);
    // coef * (np.abs(E)**2*E - np.abs(Eold)**2*Eold): numba lowers **2 to a multiplication
    d.JxKerr[nz] = A::mul(d.kerr_coef, A::sub(A::mul(A::mul(ae, ae), e), A::mul(A::mul(ao, ao), o)));
}

// every target cell reads only cells the loops have not written yet (ascending loop reads nz+1, descending loop nz-1),
// so the ten assignments are independent; both ends are evaluated from a snapshot taken in registers first
__global__ void k_mur1d(PfDormant d)
{
    const int t = threadIdx.x;          // 0..3: cells 1..4 ; 4..8: cells Nz-1 .. Nz-5
    const int Nz = d.L - 1;
    double v = 0.0;
    int nz = -1;
    if (t < 4) {
        nz = 1 + t;
        if (nz + 1 < d.L) v = A::add(d.Eold[nz + 1], A::mul(d.mur_mult, A::sub(d.Ex[nz + 1], d.Eold[nz])));
        else nz = -1;
    } else if (t < 9) {
        nz = Nz - 1 - (t - 4);
        if (nz - 1 >= 0 && nz < d.L) v = A::add(d.Eold[nz - 1], A::mul(d.mur_mult, A::sub(d.Ex[nz - 1], d.Eold[nz])));
        else nz = -1;
    }
    __syncthreads();                    // all reads before any write (a short grid could make the two ends overlap)
    if (nz >= 0) d.Ex[nz] = v;
}

// ---- Drude J-form sandbox (TESTBOXDIPSERSE.py:79-94), one time step = two launches ----------------------------------
__global__ void __launch_bounds__(DT) k_drude_h(PfDrudeJ d)
{
    const int nz = blockIdx.x * DT + threadIdx.x;
    if (nz >= d.n - 1) return;
    d.Hy[nz] = A::add(d.Hy[nz], A::mul(A::sub(d.Ex[nz + 1], d.Ex[nz]), d.inv_cour));
}
__global__ void __launch_bounds__(DT) k_drude_je(PfDrudeJ d, int i)
{
    const int nz = blockIdx.x * DT + threadIdx.x;
    if (nz >= d.n) return;
    double e = d.Ex[nz], j = d.Jx[nz];
    if (nz >= d.mat_front && nz < d.mat_rear) {
        j = A::mul(A::add(A::mul(d.kapE, j), A::mul(d.betaE, A::add(e, d.tempEOld[nz]))), d.inv_cour);
        d.Jx[nz] = j;
    }
    d.tempEOld[nz] = d.tempE[nz];
    d.tempE[nz] = e;
    const double hl = d.Hy[nz == 0 ? d.n - 1 : nz - 1];          // Python's Hy[-1] for nz = 0
    const double curl = A::sub(A::sub(d.Hy[nz], hl), A::mul(d.half_one_plus_kap, j));
    e = A::add(A::mul(d.c_self, e), A::mul(A::mul(curl, d.c_curl), d.inv_cour));
    if (nz == d.src) e = d.Hys[i];                               // hard source, after the E loop
    d.Ex[nz] = e;
}

}  // namespace pf

using namespace pf;

static int dormant_ok(const PfDormant *d, const char *who)
{
    if (!d || d->L <= 0 || !d->Ex) return set_err(PF_E_ARG, "%s: bad descriptor", who);
    if (d->mf < 0 || d->mr > d->L || d->mf > d->mr) return set_err(PF_E_ARG, "%s: bad slab range", who);
    return 0;
}

extern "C" {

int pf_varin_pbar(const PfDormant *d, void *stream)
{
    int rc = dormant_ok(d, "pf_varin_pbar");
    if (rc) return rc;
    if (!d->Qx3 || !d->Pbar3) return set_err(PF_E_ARG, "pf_varin_pbar: Qx3 / Pbar3 missing");
    if (d->mr > d->mf) k_varin_pbar<<<dblocks(d->mr - d->mf), DT, 0, (cudaStream_t)stream>>>(*d);
    PF_LAUNCH_CHECK("k_varin_pbar");
    return PF_OK;
}
int pf_varin_lin_curr_pol(const PfDormant *d, void *stream)
{
    int rc = dormant_ok(d, "pf_varin_lin_curr_pol");
    if (rc) return rc;
    if (!d->Jx || !d->P || !d->Pbar3) return set_err(PF_E_ARG, "pf_varin_lin_curr_pol: Jx / P / Pbar3 missing");
    if (d->mr > d->mf) k_varin_lin<<<dblocks(d->mr - d->mf), DT, 0, (cudaStream_t)stream>>>(*d);
    PF_LAUNCH_CHECK("k_varin_lin");
    return PF_OK;
}
int pf_varin_q_and_g(const PfDormant *d, void *stream)
{
    int rc = dormant_ok(d, "pf_varin_q_and_g");
    if (rc) return rc;
    if (!d->Qx3 || !d->Gx3) return set_err(PF_E_ARG, "pf_varin_q_and_g: Qx3 / Gx3 missing");
    if (d->mr > d->mf) k_varin_qg<<<dblocks(d->mr - d->mf), DT, 0, (cudaStream_t)stream>>>(*d);
    PF_LAUNCH_CHECK("k_varin_qg");
    return PF_OK;
}
int pf_kerr_nonlin(const PfDormant *d, void *stream)
{
    int rc = dormant_ok(d, "pf_kerr_nonlin");
    if (rc) return rc;
    if (!d->Eold || !d->JxKerr) return set_err(PF_E_ARG, "pf_kerr_nonlin: Eold / JxKerr missing");
    k_kerr_nonlin<<<dblocks(d->L), DT, 0, (cudaStream_t)stream>>>(*d);
    PF_LAUNCH_CHECK("k_kerr_nonlin");
    return PF_OK;
}
int pf_mur1d_ex(const PfDormant *d, void *stream)
{
    int rc = dormant_ok(d, "pf_mur1d_ex");
    if (rc) return rc;
    if (!d->Eold) return set_err(PF_E_ARG, "pf_mur1d_ex: Eold missing");
    if (d->L < 12) return set_err(PF_E_ARG, "pf_mur1d_ex: grid too short for the two five-cell boundary loops");
    k_mur1d<<<1, 32, 0, (cudaStream_t)stream>>>(*d);
    PF_LAUNCH_CHECK("k_mur1d");
    return PF_OK;
}

int pf_drude_j_run(const PfDrudeJ *d, int i0, int nsteps, void *stream)
{
    if (!d || d->n < 2 || !d->Ex || !d->Hy || !d->Jx || !d->tempE || !d->tempEOld || !d->Hys)
        return set_err(PF_E_ARG, "pf_drude_j_run: bad descriptor");
    if (i0 < 0 || nsteps < 0 || (long long)i0 + nsteps > d->n_src) return set_err(PF_E_ARG, "pf_drude_j_run: steps run past the source table");
    if (d->src < 0 || d->src >= d->n || d->mat_front < 0 || d->mat_rear > d->n) return set_err(PF_E_ARG, "pf_drude_j_run: bad indices");
    cudaStream_t st = (cudaStream_t)stream;
    for (int i = i0; i < i0 + nsteps; ++i) {
        k_drude_h<<<dblocks(d->n - 1), DT, 0, st>>>(*d);
        PF_LAUNCH_CHECK("k_drude_h");
        k_drude_je<<<dblocks(d->n), DT, 0, st>>>(*d, i);
        PF_LAUNCH_CHECK("k_drude_je");
    }
    return PF_OK;
}

}  // extern "C"
