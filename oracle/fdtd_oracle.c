/*
 * fdtd_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Plain-C restatement of the reference's per-step field arithmetic, one function per
 * reference leaf op, called in the reference's integrator order.  Compiled with
 *   gcc -O2 -ffp-contract=off -fno-fast-math
 * so every expression is evaluated in IEEE fp64 exactly as written (the reference's numba/LLVM
 * loops do not contract a*b+c either).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product never does.
 *
 * Each function cites the reference file:line it follows (paths under /root/reference).
 */
#include <math.h>
#include <stddef.h>
#include <string.h>

typedef struct {
    /* geometry */
    int L;        /* Nz+1 = len(V.Ex)                         MasterController.py:149-159 */
    int pw;       /* P.pmlWidth                                                           */
    int mf, mr;   /* P.materialFrontEdge / P.materialRearEdge                             */
    int nzsrc;    /* P.nzsrc                                                              */
    int tfsf;     /* P.TFSF                                                               */
    int cpml_m, cpml_p; /* P.CPMLXm / P.CPMLXp                                            */
    /* scalars */
    double dt_over_dz;  /* P.delT/P.dz as evaluated in ADE_DxUpdate   BaseFDTD11.py:753   */
    double eps0;        /* P.permit_0                                                     */
    double polA, polB, polC; /* Lorentz ADE coefficients              BaseFDTD11.py:620-626 */
    double cub_a, cub_b, cub_c; /* cubic coefficients cub/qua/one     BaseFDTD11.py:808-810 */
    double nl_den0;     /* eps0*sqrt(1.2)                             BaseFDTD11.py:864-869 */
    double nl_den1;     /* eps0*chi3Stat                                                  */
    /* state, all length L (psi too) */
    double *Ex, *Hy, *Dx, *P, *Pprev, *psiE, *psiH, *Acubic;
    /* coefficients, all length L */
    const double *Jx, *UpExMat, *denE, *UpHySelf, *UpHyMat, *denH;
    const double *beX, *ceX, *Cb, *bmY, *cmY, *C2;
    /* per-step source terms, length T: srcE[n] = Exs[n]/courantNo, srcH[n] = Hys[n]/courantNo */
    const double *srcE, *srcH;
    /* probes: values of Ex[probe_idx[p]] after each step -> probe_out[p*T + n] */
    int n_probes;
    const int *probe_idx;
    double *probe_out;
    /* snapshots: every snap_interval steps (n>0), Ex -> snap_out[(n/interval)*L ...] if row < snap_rows */
    int snap_interval, snap_rows;
    double *snap_out;
} OrcGrid;

/* BaseFDTD11.py:663-669  ADE_ExUpdate */
void orc_ex_update(OrcGrid *g)
{
    for (int nz = 1; nz < g->L; ++nz)
        g->Ex[nz] = g->Ex[nz] + (g->Hy[nz] - g->Hy[nz - 1] - g->Jx[nz]) * g->UpExMat[nz] * g->denE[nz];
}

/* BaseFDTD11.py:640-656  ADE_HyUpdate  (range(1, P.Nz), Nz = L-1) */
void orc_hy_update(OrcGrid *g)
{
    for (int nz = 1; nz < g->L - 1; ++nz)
        g->Hy[nz] = g->Hy[nz] * g->UpHySelf[nz] + (g->Ex[nz + 1] - g->Ex[nz]) * g->UpHyMat[nz] * g->denH[nz];
}

/* BaseFDTD11.py:364-376  CPML_Psi_e_Update */
void orc_psi_e(OrcGrid *g)
{
    if (g->cpml_m)
        for (int nz = 1; nz < g->pw; ++nz) {
            g->psiE[nz] = g->beX[nz] * g->psiE[nz] + g->ceX[nz] * (g->Hy[nz] - g->Hy[nz - 1]);
            g->Ex[nz] = g->Ex[nz] - g->Cb[nz] * g->psiE[nz];
        }
    if (g->cpml_p)
        for (int nz = g->L - g->pw; nz < g->L; ++nz) {
            g->psiE[nz] = g->beX[nz] * g->psiE[nz] + g->ceX[nz] * (g->Hy[nz] - g->Hy[nz - 1]);
            g->Ex[nz] = g->Ex[nz] - g->Cb[nz] * g->psiE[nz];
        }
}

/* BaseFDTD11.py:381-393  CPML_Psi_m_Update */
void orc_psi_m(OrcGrid *g)
{
    if (g->cpml_m)
        for (int nz = 1; nz < g->pw; ++nz) {
            g->psiH[nz] = g->bmY[nz] * g->psiH[nz] + g->cmY[nz] * (g->Ex[nz + 1] - g->Ex[nz]);
            g->Hy[nz] = g->Hy[nz] + g->C2[nz] * g->psiH[nz];
        }
    if (g->cpml_p)
        for (int nz = g->L - g->pw; nz < g->L - 1; ++nz) {
            g->psiH[nz] = g->bmY[nz] * g->psiH[nz] + g->cmY[nz] * (g->Ex[nz + 1] - g->Ex[nz]);
            g->Hy[nz] = g->Hy[nz] + g->C2[nz] * g->psiH[nz];
        }
}

/* Solver_Engine.py:177-179 / 249-252 / 307-310  soft source + one-point TF/SF correction */
void orc_source(OrcGrid *g, int n)
{
    g->Ex[g->nzsrc] += g->srcE[n];
    if (g->tfsf)
        g->Hy[g->nzsrc - 1] -= g->srcH[n];
}

/* BaseFDTD11.py:750-760  ADE_DxUpdate */
/* Current slot in material cells -- BUILDER-DEFINED (parity unpinned by the reference): the reference's only PIC contract
 * is the Jx that ADE_ExUpdate subtracts (:667), but inside [mf, mr) ADE_ExCreate / NonLinExUpdate overwrite Ex from Dx, so
 * a beam current there would drive nothing.  Ampere's law for the flux density is dD/dt = curl H - J, which in the
 * reference's units (the slot holds J*dz, the bracket is multiplied by dt/dz) is the same bracket as in ADE_ExUpdate:
 *     Dx += (Hy[nz] - Hy[nz-1] - Jx[nz]) * dt/dz * denE.
 * With Jx = 0 (every reference run) this is ADE_DxUpdate bit for bit: x - 0.0 == x.                                   */
void orc_dx_update(OrcGrid *g)
{
    for (int nz = g->mf; nz < g->mr; ++nz)
        g->Dx[nz] = g->Dx[nz] + (g->Hy[nz] - g->Hy[nz - 1] - g->Jx[nz]) * g->dt_over_dz * g->denE[nz];
}

/* BaseFDTD11.py:487-538 (history rotation) + :609-633 ADE_PolarisationCurrent_Ex.
 * The reference copies P^n into tempVarPol and P^{n-1} into tempTempVarPol for every cell,
 * then P^{n+1} = A*P^n + B*P^{n-1} + C*E^n inside the slab. */
void orc_pol_update(OrcGrid *g)
{
    for (int nz = g->mf; nz < g->mr; ++nz) {
        double pn = g->P[nz];
        g->P[nz] = g->polA * pn + g->polB * g->Pprev[nz] + g->polC * g->Ex[nz];
        g->Pprev[nz] = pn;
    }
}

/* BaseFDTD11.py:712-725  ADE_ExCreate */
void orc_ex_create(OrcGrid *g)
{
    for (int nz = g->mf; nz < g->mr; ++nz)
        g->Ex[nz] = (g->Dx[nz] - g->P[nz]) / g->eps0;
}

/* CubicEquationSolver.py:29-105: first root of a x^3 + b x^2 + c x + d (a != 0 assumed unless
 * the degenerate branches below).  Only root[0] is consumed (BaseFDTD11.py:838). */
static double cbrt_like_ref(double v) { return pow(v, 1 / 3.0); }

double orc_cubic_root0(double a, double b, double c, double d)
{
    if (a == 0 && b == 0)
        return (-d * 1.0) / c;
    if (a == 0) {
        double D = c * c - 4.0 * b * d;
        if (D >= 0) {
            D = sqrt(D);
            return (-c + D) / (2.0 * b);
        }
        return (-c) / (2.0 * b); /* real part of the complex root */
    }
    /* findF / findG / findH, CubicEquationSolver.py:94-105 ("**" is libm pow in CPython) */
    double f = ((3.0 * c / a) - (pow(b, 2.0) / pow(a, 2.0))) / 3.0;
    double gg = (((2.0 * pow(b, 3.0)) / pow(a, 3.0)) - ((9.0 * b * c) / pow(a, 2.0)) + (27.0 * d / a)) / 27.0;
    double h = (pow(gg, 2.0) / 4.0 + pow(f, 3.0) / 27.0);
    if (f == 0 && gg == 0 && h == 0) {
        if ((d / a) >= 0)
            return pow(d / (1.0 * a), 1 / 3.0) * -1;
        return pow(-d / (1.0 * a), 1 / 3.0);
    }
    if (h <= 0) {
        double i = sqrt((pow(gg, 2.0) / 4.0) - h);
        double j = pow(i, 1 / 3.0);
        double k = acos(-(gg / (2 * i)));
        return 2 * j * cos(k / 3.0) - (b / (3.0 * a));
    }
    double R = -(gg / 2.0) + sqrt(h);
    double S = (R >= 0) ? cbrt_like_ref(R) : cbrt_like_ref(-R) * -1;
    double T = -(gg / 2.0) - sqrt(h);
    double U = (T >= 0) ? cbrt_like_ref(T) : cbrt_like_ref(-T) * -1;
    return (S + U) - (b / (3.0 * a));
}

/* BaseFDTD11.py:793-853: cubPoly = [cub, qua, one, -|Dx/eps0|^2]; Acubic = Re(root0) where |d|>1e-8 else 0 */
void orc_acubic(OrcGrid *g)
{
    for (int nz = g->mf; nz < g->mr; ++nz) {
        double q = fabs(g->Dx[nz] / g->eps0);
        double d = -pow(q, 2.0); /* -np.abs(x)**2 : unary minus binds looser than ** */
        double out = 0.0;
        if (fabs(d) > 1e-8)
            out = orc_cubic_root0(g->cub_a, g->cub_b, g->cub_c, d);
        g->Acubic[nz] = out;
    }
}

/* BaseFDTD11.py:858-877  NonLinExUpdate */
void orc_nl_ex(OrcGrid *g)
{
    for (int nz = g->mf; nz < g->mr; ++nz)
        g->Ex[nz] = g->Dx[nz] / (g->nl_den0 + g->nl_den1 * g->Acubic[nz]);
}

/* Mode LORENTZ_NL (config 5's "dispersive and nonlinear" material; NOT in the reference, whose nonlinear
 * integrator has no dispersion ADE -- builder-defined composition, parity unpinned): the Lorentz ADE
 * carries the dispersive polarisation P (orc_pol_update), and the instantaneous Kerr response is solved
 * on what is left of D, Dn = Dx - P (the numerator of ADE_ExCreate, BaseFDTD11.py:712-725), with the
 * reference's own cubic chain (BaseFDTD11.py:793-877):
 *   Acubic = positive root of [cub, qua, one, -|Dn/eps0|^2] if that |d| > 1e-8 else 0;  Ex = Dn/(den0 + den1*Acubic)
 * With cub = chi3^2, qua = 2*eps_inf*chi3, one = eps_inf^2, den0 = eps0*eps_inf, den1 = eps0*chi3 this is
 * Dn = eps0*(eps_inf + chi3*|E|^2)*E. */
/* The composition's Acubic is DEFINED as the positive root of the cubic, converged to rounding level.  The
 * reference's closed form is not used here: with Kerr coefficients (cub = chi3^2, qua = 2 chi3, one = 1) it runs
 * in its three-real-root branch, x = 2 j cos(acos(.)/3) - b/3a, which is ill-conditioned (acos near +-1
 * amplifies one ulp of its argument to ~1e-8 of the root), so two libm implementations disagree at 1e-8.
 * For a, b >= 0, c > 0, q2 > 0 the polynomial is increasing and convex on x > 0 and every single term bounds
 * the root from above, so Newton from x0 = min(q2/c, sqrt(q2/b), cbrt(q2/a)) descends monotonically. */
double orc_cubic_root_newton(double a, double b, double c, double q2)
{
    if (!(a >= 0 && b >= 0 && c > 0))
        return orc_cubic_root0(a, b, c, -q2);
    double x = q2 / c;
    if (b > 0 && sqrt(q2 / b) < x) x = sqrt(q2 / b);
    if (a > 0 && cbrt(q2 / a) < x) x = cbrt(q2 / a);
    for (int it = 0; it < 200; ++it) {
        double p = ((a * x + b) * x + c) * x - q2;
        double dp = (3.0 * a * x + 2.0 * b) * x + c;
        double step = p / dp;
        x -= step;
        if (fabs(step) <= 1e-13 * x)
            break;
    }
    return x;
}

void orc_acubic_dn(OrcGrid *g)
{
    for (int nz = g->mf; nz < g->mr; ++nz) {
        double q = fabs((g->Dx[nz] - g->P[nz]) / g->eps0);
        double d = -pow(q, 2.0);
        double out = 0.0;
        if (fabs(d) > 1e-8)
            out = orc_cubic_root_newton(g->cub_a, g->cub_b, g->cub_c, -d);
        g->Acubic[nz] = out;
    }
}

void orc_nl_ex_dn(OrcGrid *g)
{
    for (int nz = g->mf; nz < g->mr; ++nz)
        g->Ex[nz] = (g->Dx[nz] - g->P[nz]) / (g->nl_den0 + g->nl_den1 * g->Acubic[nz]);
}

static void orc_record(OrcGrid *g, int n, int T)
{
    for (int p = 0; p < g->n_probes; ++p)
        g->probe_out[(size_t)p * T + n] = g->Ex[g->probe_idx[p]];
    if (g->snap_out && g->snap_interval > 0 && n > 0 && n % g->snap_interval == 0) {
        int row = n / g->snap_interval; /* Solver_Engine.py:57-68 vidMake */
        if (row < g->snap_rows)
            memcpy(g->snap_out + (size_t)row * g->L, g->Ex, sizeof(double) * g->L);
    }
}

enum { ORC_FREE = 0, ORC_LORENTZ = 1, ORC_NL = 2, ORC_LORENTZ_NL = 3 };

/* One pass of T steps starting at step n0 (sources indexed by absolute step).
 * mode FREE    : Solver_Engine.py:167-183
 * mode LORENTZ : Solver_Engine.py:294-316 (do_pol = pass index == 1)
 * mode NL      : Solver_Engine.py:236-261
 * mode LORENTZ_NL : the Lorentz loop with ADE_ExCreate replaced by the cubic chain on Dx - P (see above) */
int orc_run(OrcGrid *g, int mode, int do_pol, int n0, int nsteps, int T_total)
{
    int cpml = g->cpml_m || g->cpml_p;
    for (int n = n0; n < n0 + nsteps; ++n) {
        if ((mode == ORC_LORENTZ || mode == ORC_LORENTZ_NL) && do_pol)
            orc_pol_update(g);
        orc_ex_update(g);
        if (cpml)
            orc_psi_e(g);
        orc_source(g, n);
        if (mode == ORC_LORENTZ) {
            orc_dx_update(g);
            orc_ex_create(g);
        } else if (mode == ORC_NL) {
            orc_dx_update(g);
            orc_acubic(g);
            orc_nl_ex(g);
        } else if (mode == ORC_LORENTZ_NL) {
            orc_dx_update(g);
            orc_acubic_dn(g);
            orc_nl_ex_dn(g);
        }
        orc_hy_update(g);
        if (cpml)
            orc_psi_m(g);
        orc_record(g, n, T_total);
    }
    return 0;
}

/* Batch driver for the CPU baseline: members are independent (MasterController.py:543-563),
 * so they are handed out to n_threads host threads through a shared counter (pthreads; this
 * image has no libgomp). */
#include <pthread.h>
typedef struct {
    OrcGrid *grids;
    int n_members, mode, do_pol, n0, nsteps;
    const int *T_total;
    int next;
    pthread_mutex_t mu;
} OrcBatch;

static void *orc_batch_worker(void *arg)
{
    OrcBatch *b = (OrcBatch *)arg;
    for (;;) {
        pthread_mutex_lock(&b->mu);
        int m = b->next++;
        pthread_mutex_unlock(&b->mu);
        if (m >= b->n_members)
            break;
        orc_run(&b->grids[m], b->mode, b->do_pol, b->n0, b->nsteps, b->T_total[m]);
    }
    return NULL;
}

int orc_run_batch(OrcGrid *grids, int n_members, int mode, int do_pol, int n0, int nsteps,
                  const int *T_total, int n_threads)
{
    OrcBatch b = {grids, n_members, mode, do_pol, n0, nsteps, T_total, 0, PTHREAD_MUTEX_INITIALIZER};
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    for (int t = 0; t < n_threads; ++t)
        pthread_create(&th[t], NULL, orc_batch_worker, &b);
    for (int t = 0; t < n_threads; ++t)
        pthread_join(th[t], NULL);
    return 0;
}

size_t orc_sizeof_grid(void) { return sizeof(OrcGrid); }

/* ---- dormant models: leaf functions of the reference that no integrator calls -------------------- */
/* BaseFDTD11.py:567-577  ADE_NonLin_Pol_Ex_Pbar (Varin: linear + instantaneous Kerr + Raman polarisation target) */
void orc_varin_pbar(int mf, int mr, double eps0, double chi1, double chi3, double alpha3, const double *Ex, const double *Qx3,
                    double *Pbar3)
{
    for (int nz = mf; nz < mr; ++nz)
        Pbar3[nz] = eps0 * (chi1 * Ex[nz] + chi3 * (alpha3 * Ex[nz] * Ex[nz] * Ex[nz] + (1 - alpha3) * Qx3[nz] * Ex[nz]));
}
/* BaseFDTD11.py:580-594  ADE_Lin_Curr_And_Pol_Varin */
void orc_varin_lin(int mf, int mr, double gammaE, double omega0, double dt, double *Jx, double *P, const double *Pbar3)
{
    double Gamma = (gammaE * dt) / 2, A = 1 - Gamma, D = 1 + Gamma, B = omega0 * omega0 * dt;
    for (int nz = mf; nz < mr; ++nz) {
        Jx[nz] = (A / D) * Jx[nz] + (B / D) * (Pbar3[nz] - P[nz]);
        P[nz] = P[nz] + dt * Jx[nz];
    }
}
/* BaseFDTD11.py:596-609  ADE_Nonlin_Q_and_G (Raman oscillator) */
void orc_varin_qg(int mf, int mr, double gamma3, double omega3, double dt, const double *Ex, double *Gx3, double *Qx3)
{
    double Gamma = (gamma3 * dt) / 2, e = 1 - Gamma, f = 1 + Gamma, h = omega3 * omega3 * dt;
    for (int nz = mf; nz < mr; ++nz) {
        Gx3[nz] = (e / f) * Gx3[nz] + (h / f) * (Ex[nz] * Ex[nz] - Qx3[nz]);
        Qx3[nz] = Qx3[nz] + dt * Gx3[nz];
    }
}
/* BaseFDTD11.py:762-766  KerrNonlin */
void orc_kerr_nonlin(int n, double alpha3, double eps0, double chi3, double dt, const double *Ex, const double *Eold, double *JxKerr)
{
    double coef = (alpha3 * eps0 * chi3) / dt;
    for (int nz = 0; nz < n; ++nz) {
        double ae = fabs(Ex[nz]), ao = fabs(Eold[nz]);
        JxKerr[nz] = coef * (ae * ae * Ex[nz] - ao * ao * Eold[nz]);
    }
}
/* BaseFDTD11.py:769-788  MUR1DEx */
void orc_mur1d(int Nz, double c0, double dt, double dz, double *Ex, const double *Eold)
{
    double m = (c0 * dt - dz) / (c0 * dt + dz);
    for (int nz = 1; nz < 5; ++nz)
        Ex[nz] = Eold[nz + 1] + m * (Ex[nz + 1] - Eold[nz]);
    for (int nz = Nz - 1; nz > Nz - 6; --nz)
        Ex[nz] = Eold[nz - 1] + m * (Ex[nz - 1] - Eold[nz]);
}
/* TESTBOXDIPSERSE.py:79-94: the Drude J-form loop as written (Hy[nz-1] of cell 0 is Python's Hy[-1]) */
void orc_drude_j(int n, int tim, int src, int mat_front, int mat_rear, double cour, double kapE, double betaE, double perm0,
                 double dt, const double *Hys, double *Ex, double *Hy, double *Jx, double *tempE, double *tempEOld)
{
    for (int i = 0; i < tim; ++i) {
        for (int nz = 0; nz < n - 1; ++nz)
            Hy[nz] = Hy[nz] + (Ex[nz + 1] - Ex[nz]) * (1 / cour);
        for (int nz = mat_front; nz < mat_rear; ++nz)
            Jx[nz] = (kapE * Jx[nz] + betaE * (Ex[nz] + tempEOld[nz])) * (1 / cour);
        for (int nz = 0; nz < n; ++nz) {
            tempEOld[nz] = tempE[nz];
            tempE[nz] = Ex[nz];
            double hl = Hy[nz == 0 ? n - 1 : nz - 1];
            Ex[nz] = ((2 * perm0 - betaE * dt) / (2 * perm0 + betaE * dt)) * Ex[nz]
                     + (Hy[nz] - hl - 0.5 * (1 + kapE) * Jx[nz]) * ((2 * dt) / (2 * perm0 + betaE * dt)) * (1 / cour);
        }
        Ex[src] = Hys[i];
    }
}
