/*
 * pyfdtd_b200.h -- C-ABI of libpyfdtd_b200.so, the sm_100a implementation of the Py-FDTD_PIC
 * time-stepping hot path.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types cross this boundary.
 *   - Every `double*` / `int*` inside PfGrid and the PIC structs is a DEVICE pointer (the Python
 *     host layer takes them from torch.Tensor.data_ptr(); torch is only the buffer carrier).
 *     PfGrid structs themselves and the `const int* T_total`-style arrays are HOST memory.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls enqueue
 *     work and return; pf_sync() or the caller's own stream sync waits for it.
 *   - Return value: 0 = ok, <0 = PF_E_* (pf_last_error() gives text).  The reference reports
 *     errors by print+sys.exit (Environment_Setup.py:59-155, BaseFDTD11.py:36-66); the Python
 *     layer turns these codes into exceptions.
 *   - The library never allocates caller-visible memory; scratch is provided by the caller and
 *     sized by the pf_*_scratch_bytes() queries.
 *   - Process-global state: the last-error string (per thread), the launch counter (atomic) and the
 *     optional profiling event list (mutex-guarded).  Nothing is remembered about earlier calls'
 *     buffers: every entry point works from its arguments alone, so calls from several host threads
 *     (one per GPU) do not interact.
 *
 * All arrays of one grid have length L = Nz+1 (MasterController.py:149-159), indices are the
 * reference's own cell indices.
 */
#ifndef PYFDTD_B200_H
#define PYFDTD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PF_ABI_VERSION 2

enum { PF_OK = 0, PF_E_ARG = -1, PF_E_CUDA = -2, PF_E_UNSUPPORTED = -3, PF_E_SCRATCH = -4 };

/* integrator selector: which Solver_Engine.Integrator*1D loop body is run */
enum {
    PF_FREE = 0,    /* Solver_Engine.py:167-183  IntegratorFreeSpace1D */
    PF_LORENTZ = 1, /* Solver_Engine.py:294-316  IntegratorLinLor1D    */
    PF_NL = 2,      /* Solver_Engine.py:236-261  IntegratorNL1D        */
    PF_LORENTZ_NL = 3 /* BASELINE config 5's "dispersive and nonlinear" material.  NOT a reference
                         integrator (its nonlinear loop has no dispersion ADE): the IntegratorLinLor1D
                         loop with ADE_ExCreate (BaseFDTD11.py:712-725) replaced by the reference's cubic
                         chain (BaseFDTD11.py:793-877) applied to Dn = Dx - P:
                           Acubic = positive root of [cub_a, cub_b, cub_c, -|Dn/eps0|^2] (0 where |d| <= 1e-8),
                           Ex = Dn / (nl_den0 + nl_den1*Acubic).
                         The root is the converged (Newton) one whenever cub_a, cub_b >= 0, cub_c > 0: with
                         Kerr coefficients the reference's closed form runs in its ill-conditioned
                         three-real-root branch (acos near +-1), where libm implementations differ at 1e-8.
                         Builder-defined composition; parity is against oracle/fdtd_oracle.c only.  */
};

/* engine selector for pf_run_pass / pf_run_batch */
enum {
    PF_ENGINE_OPS = 0, /* one kernel per reference leaf op, general per-cell coefficient arrays   */
    PF_ENGINE_TILE = 1 /* fused, register/shared-memory resident, k-step temporally blocked tiles;
                          requires PfGrid.canonical (see below)                                   */
};

/* flags for PfGrid.flags */
enum {
    PF_F_TFSF = 1,      /* P.TFSF                                              */
    PF_F_CPML_M = 2,    /* P.CPMLXm                                            */
    PF_F_CPML_P = 4,    /* P.CPMLXp                                            */
    PF_F_CANONICAL = 8, /* coefficient arrays have the piecewise form the tile engine assumes:
                           denE = denH = UpHySelf = 1, bmY == beX, Jx absent/0,
                           UpExMat/UpHyMat constant outside [mf,mr) and constant inside,
                           Cb[nz] == UpExMat[nz] and C2[nz] == c2_pml on the CPML correction
                           ranges and 0 at nz = L-pw (BaseFDTD11.py:306-330).  Set by the host
                           layer after checking the arrays bit for bit.                          */
    PF_F_FMA = 16,      /* allow fused multiply-add contraction (faster, not bit-identical to the
                           reference's un-contracted fp64 arithmetic; <= 1e-10 relative)         */
    PF_F_NEWTON = 64,   /* nonlinear mode: find the positive root of the per-cell cubic by Newton iteration
                           instead of the reference's closed form (CubicEquationSolver.py:29-105) when
                           cub, qua >= 0 and one > 0 (otherwise the closed form is used).  ~3x fewer
                           instructions; differs from the closed form by its own cancellation error
                           (<= 1e-10 absolute on Acubic, <= 1e-10 relative on the fields).          */
    PF_F_FP32 = 32      /* optional single-precision mode of the tile engine: the on-chip state is
                           advanced in fp32 (contracted; Lorentz ADE in difference form, cubic root by
                           Newton iteration); the arrays in device memory stay fp64 and are converted at
                           tile load / store.  Not a parity mode.  Stated tolerance: 1e-5 of the peak
                           (fields: of their own peak; probe traces: of the peak of the probe traces)
                           for runs of up to 8192 time steps, growing like sqrt(steps) beyond -- the
                           single-precision rounding of the running fields is a random walk -- i.e.
                           1.7e-5 at the reference's default 23997 steps (measured there: 5.6e-6 to
                           1.3e-5), 2e-5 at its maximum 2^15.  Carrying the fields as float pairs would
                           hold 1e-5 at any length but costs more than fp64 on this GPU (FP32 : FP64
                           issue rate is 2 : 1).  PF_ENGINE_OPS rejects the flag (PF_E_UNSUPPORTED).  */
};

/* One 1-D grid: a single simulation, one sweep member, or one rank's slab of a long grid.
 * Mirrors the arrays of the reference's Variables / CPML_Variables jitclasses
 * (MasterController.py:145-210, 401-434) that the hot loop touches. */
typedef struct PfGrid {
    /* geometry (Params, MasterController.py:287-336) */
    int32_t L;        /* Nz+1                                                                    */
    int32_t pw;       /* pmlWidth                                                                */
    int32_t mf, mr;   /* materialFrontEdge, materialRearEdge: slab = [mf, mr)                    */
    int32_t nzsrc;    /* source cell; TF/SF correction is applied to Hy[nzsrc-1]                 */
    int32_t flags;    /* PF_F_*                                                                  */
    int32_t n_probes; /* number of probe cells                                                   */
    int32_t probe_stride; /* row length of probe_out (= timeSteps)                               */
    int32_t n_src;    /* entries of srcE / srcH (= timeSteps).  pf_run_* reject n0 + nsteps > n_src
                         (and > probe_stride when n_probes > 0) with PF_E_ARG instead of reading /
                         writing past the tables; 0 = length not given, not checked                 */
    int32_t reserved0; /* must be 0                                                                 */
    /* global-index window for domain decomposition: this grid holds cells [z0, z0+L) of a global
     * grid of Lg cells (z0 = 0, Lg = L for an undecomposed grid).  Index-dependent rules (update
     * ranges, CPML ranges, slab, source, probes) are evaluated on global indices.
     * LIMIT: the kernels evaluate these rules in 32-bit arithmetic, so Lg must be < 2^31 - 2^12 cells
     * (a 1e9-cell grid is fine; pf_run_* return PF_E_UNSUPPORTED beyond); L itself is an int32.   */
    int64_t z0, Lg;
    /* scalars */
    double dt_over_dz;          /* P.delT/P.dz                         BaseFDTD11.py:753         */
    double eps0;                /* P.permit_0                                                    */
    double polA, polB, polC;    /* Lorentz ADE                         BaseFDTD11.py:620-626     */
    double cub_a, cub_b, cub_c; /* cubic coefficients cub, qua, one    BaseFDTD11.py:808-810     */
    double nl_den0, nl_den1;    /* eps0*sqrt(1.2), eps0*chi3Stat       BaseFDTD11.py:864-869     */
    /* canonical-form scalars (used by the tile engine only) */
    double cE0, cE1;            /* UpExMat outside / inside [mf,mr)                              */
    double cH0, cH1;            /* UpHyMat outside / inside [mf,mr)                              */
    double c2_pml;              /* C2 on the CPML ranges (delT/mu0)                              */
    /* state (read and written) */
    double *Ex, *Hy, *Dx, *P, *Pprev, *psiE, *psiH, *Acubic;
    /* coefficients (read only); Jx may be NULL (= 0): the PIC current slot, in units of J*dz.  ADE_ExUpdate subtracts it
     * (BaseFDTD11.py:667: Ex += (Hy[nz]-Hy[nz-1]-Jx[nz])*UpExMat*den).  Inside the slab [mf, mr) the reference's loops
     * overwrite Ex from Dx (ADE_ExCreate / NonLinExUpdate), so there the same bracket enters ADE_DxUpdate instead:
     * Dx += (Hy[nz]-Hy[nz-1]-Jx[nz])*dt/dz*den (dD/dt = curl H - J) -- builder-defined, the reference has no particles;
     * with Jx NULL or 0 both updates are the reference's bit for bit.                                                */
    const double *Jx, *UpExMat, *denE, *UpHySelf, *UpHyMat, *denH;
    const double *beX, *ceX, *Cb, *bmY, *cmY, *C2;
    /* per-step source terms, indexed by absolute step n:
     *   srcE[n] = Exs[n]/P.courantNo, srcH[n] = Hys[n]/P.courantNo   Solver_Engine.py:307-310   */
    const double *srcE, *srcH;
    /* probes: Ex[probe_idx[p]] after step n -> probe_out[p*probe_stride + n]
     * (Solver_Engine.probeSim :16-54; windowing of x1ColBe/x1ColAf is applied by the host layer) */
    const int32_t *probe_idx;
    double *probe_out;
} PfGrid;

/* ---- library / device ------------------------------------------------------------------- */
int pf_abi_version(void);
const char *pf_last_error(void);
/* fills sm_count, cc_major, cc_minor, free/total bytes; returns PF_E_CUDA without a device      */
int pf_device_info(int *sm_count, int *cc_major, int *cc_minor, size_t *free_bytes, size_t *total_bytes);
int pf_sync(void *stream);
/* number of kernels this library has launched since load (bench.py's gpu_launches)             */
unsigned long long pf_launch_count(void);

/* Host-side helpers of the setup chain (no GPU involved): y[i] = exp(x[i]) / pow(x[i], e) through the C
 * library, one element at a time.  The reference's numba-compiled setup loops (BaseFDTD11.py:237-240,
 * 278-294) call libm per element; NumPy's SIMD exp/pow differ from it in the last bit.               */
int pf_host_exp(const double *x, double *y, long long n);
int pf_host_pow(const double *x, double e, double *y, long long n);

int pf_host_sin(const double *x, double *y, long long n, int threads);   /* libm sin / cos, elementwise, on `threads`  */
int pf_host_cos(const double *x, double *y, long long n, int threads);   /* host threads (0 = all)                      */

/* Host-side setup of a whole sweep in one call: for every member the CPML recursive-convolution
 * profiles (BaseFDTD11.CPML_ScalingCalc :222-272, CPML_Ex_RC_Define :274-286, CPML_HY_RC_Define :288-296) and the
 * per-step source tables (Solver_Engine.SourceManager :89-124 over BaseFDTD11.SmoothTurnOn :104-120, divided by
 * courantNo as PfGrid.srcE / srcH want them), written straight into the caller's upload buffer `out` at the given
 * offsets (in doubles).  Same IEEE operations in the same order as the reference's loops (libm exp / pow / sin):
 * bit-identical to the per-member Python chain.  Members are independent; `threads` host threads (0 = all).        */
typedef struct PfSetupMember {
    int32_t L;        /* Nz+1: the three profile arrays are written over [0, L), zero outside the CPML cells        */
    int32_t pw;       /* pmlWidth                                                                                   */
    int32_t n_src;    /* source entries to generate (<= timeSteps)                                                  */
    int32_t src_kind; /* 0: the caller fills the source tables itself, 1: P.SineCont (SmoothTurnOn)                 */
    int32_t tfsf;     /* P.TFSF: Hys scaled by 1/CharImp                          Solver_Engine.py:103-104         */
    int32_t pump;     /* P.nonLinMed: pump at 0.8 f added, 0.1 / 0.01 weights     Solver_Engine.py:95-100          */
    double dz, delT, eps0;
    double kappaMax, r_scale, r_a_scale, sigmaOpt, alphaMax;   /* CPML_Params, MasterController.py:341-360          */
    double c0, freq, courantNo, period, periods, charImp;      /* Params members the source uses                    */
    double amp;       /* the member's amplitude: multiplies Exs and Hys (1.0 for a plain frequency sweep)           */
    int64_t off_beX, off_ceX, off_cmY; /* offsets into out[]; off_beX < 0: profiles not written (shared with another member) */
    int64_t off_srcE, off_srcH;        /* off_srcE < 0: sources not written                                          */
} PfSetupMember;
int pf_host_sweep_inputs(const PfSetupMember *members, int n_members, double *out, int threads);

/* ---- leaf ops: one call = one reference leaf function on one grid ---------------------- */
int pf_ade_ex_update(const PfGrid *g, void *stream);        /* BaseFDTD11.py:663-669  ADE_ExUpdate               */
int pf_ade_hy_update(const PfGrid *g, void *stream);        /* BaseFDTD11.py:640-656  ADE_HyUpdate               */
int pf_cpml_psi_e_update(const PfGrid *g, void *stream);    /* BaseFDTD11.py:364-376  CPML_Psi_e_Update          */
int pf_cpml_psi_m_update(const PfGrid *g, void *stream);    /* BaseFDTD11.py:381-393  CPML_Psi_m_Update          */
int pf_source_inject(const PfGrid *g, int n, void *stream); /* Solver_Engine.py:307-310                          */
int pf_ade_dx_update(const PfGrid *g, void *stream);        /* BaseFDTD11.py:750-760  ADE_DxUpdate               */
int pf_ade_polarisation_update(const PfGrid *g, void *stream); /* BaseFDTD11.py:487-538 + 609-633 (history shift + ADE) */
int pf_ade_ex_create(const PfGrid *g, void *stream);        /* BaseFDTD11.py:712-725  ADE_ExCreate               */
int pf_acubic_finder(const PfGrid *g, void *stream);        /* BaseFDTD11.py:793-853  AcubicFinder + cubic solve */
int pf_nonlin_ex_update(const PfGrid *g, void *stream);     /* BaseFDTD11.py:858-877  NonLinExUpdate             */
int pf_probe_record(const PfGrid *g, int n, void *stream);  /* Solver_Engine.py:16-54 probeSim                   */
/* CubicEquationSolver.solve root[0] (CubicEquationSolver.py:29-105) for n polynomials,
 * coeffs = [n][4] (a,b,c,d) device, root0 = [n] device                                          */
int pf_cubic_root0(const double *coeffs, double *root0, int n, void *stream);
/* same contract, PF_F_NEWTON arithmetic: Newton iteration where a, b >= 0, c > 0, d < 0 (one positive
 * root, monotone convergence from -d/c), the closed form elsewhere                                */
int pf_cubic_root0_newton(const double *coeffs, double *root0, int n, void *stream);
/* every root, as CubicEquationSolver.solve returns them: roots = [n][3][2] (re, im) device,
 * nroots = [n] device (1 linear, 2 quadratic, 3 cubic)                                          */
int pf_cubic_solve(const double *coeffs, double *roots, int *nroots, int n, void *stream);

/* ---- dormant models of the reference: leaf functions no reference integrator calls ------------
 * One call = one reference leaf function on one grid, bit-identical to the Python / numba original.  Pointers are device
 * arrays of length L (Nz+1 here; the reference allocates these members with Nz entries and only touches [mf, mr)).     */
typedef struct PfDormant {
    int32_t L, mf, mr, reserved0;
    double eps0, dt;
    double chi1, chi3, alpha3, one_minus_alpha3;  /* V.chi1Stat, V.chi3Stat, V.alpha3 (float32 member), 1 - V.alpha3        */
    double lin_AoverD, lin_BoverD;   /* (1-G)/(1+G), w0^2 dt/(1+G), G = gammaE dt/2          BaseFDTD11.py:583-586       */
    double ram_eoverf, ram_hoverf;   /* same with nonLin3gammaE / nonLin3Omega_0E            BaseFDTD11.py:598-601       */
    double kerr_coef;                /* (alpha3 eps0 chi3)/dt                                BaseFDTD11.py:765           */
    double mur_mult;                 /* (c0 dt - dz)/(c0 dt + dz)                            BaseFDTD11.py:771           */
    double *Ex;                      /* V.Ex            (MUR1DEx writes it)                                             */
    const double *Eold;              /* V.tempTempVarE  (KerrNonlin, MUR1DEx)                                           */
    double *Jx, *P, *Pbar3, *Qx3, *Gx3, *JxKerr;   /* V.Jx, V.polarisationCurr, V.Pbar3, V.Qx3, V.Gx3, V.JxKerr         */
} PfDormant;
int pf_varin_pbar(const PfDormant *d, void *stream);          /* BaseFDTD11.py:567-577  ADE_NonLin_Pol_Ex_Pbar       */
int pf_varin_lin_curr_pol(const PfDormant *d, void *stream);  /* BaseFDTD11.py:580-594  ADE_Lin_Curr_And_Pol_Varin   */
int pf_varin_q_and_g(const PfDormant *d, void *stream);       /* BaseFDTD11.py:596-609  ADE_Nonlin_Q_and_G           */
int pf_kerr_nonlin(const PfDormant *d, void *stream);         /* BaseFDTD11.py:762-766  KerrNonlin                   */
int pf_mur1d_ex(const PfDormant *d, void *stream);            /* BaseFDTD11.py:769-788  MUR1DEx                      */

/* The Drude medium in current (J) form as the reference steps it in its scratch script (TESTBOXDIPSERSE.py:79-94): vacuum
 * H update, J update inside [mat_front, mat_rear), E update of every cell with the J term (Hy[-1] wraps for cell 0, as the
 * Python does), hard source Ex[src] = Hys[i]; no PML.  Steps i0 .. i0+nsteps-1.                                        */
typedef struct PfDrudeJ {
    int32_t n, src, mat_front, mat_rear, n_src, reserved0;
    double inv_cour, kapE, betaE, c_self, c_curl, half_one_plus_kap;
    double *Ex, *Hy, *Jx, *tempE, *tempEOld;
    const double *Hys;
} PfDrudeJ;
int pf_drude_j_run(const PfDrudeJ *d, int i0, int nsteps, void *stream);

/* ---- integrator passes ------------------------------------------------------------------ */
/* One pass of `nsteps` steps starting at absolute step n0 on ONE grid: the body of the
 * `for counts in range(P.timeSteps)` loop of Solver_Engine.Integrator{FreeSpace,LinLor,NL}1D
 * (Solver_Engine.py:167 / 294 / 236).  do_pol = 1 runs the polarisation update (pass i==1 of the
 * Lorentz integrator).  snap_out (may be NULL): Ex is copied to row n/snap_interval after step
 * n whenever n>0 and n % snap_interval == 0 and row < snap_rows (Solver_Engine.vidMake :57-68);
 * ENGINE_TILE writes the rows from inside its kernel in modes PF_FREE / PF_LORENTZ and ends a launch at
 * every snapshot step in the cubic modes.
 * scratch: pf_run_scratch_bytes(g,1,engine) bytes of device memory (may be NULL for ENGINE_OPS). */
int pf_run_pass(const PfGrid *g, int mode, int do_pol, int n0, int nsteps, int engine,
                double *snap_out, int snap_interval, int snap_rows,
                void *scratch, size_t scratch_bytes, void *stream);

/* The same pass over n_grids independent grids (the members of a LoopedSim sweep,
 * MasterController.py:543-563), fused: every launch advances every member by k steps.
 * Members may differ in every field of PfGrid; nsteps[m] gives each member's step count
 * (members that finish early idle).  k_block = steps per launch; 0 = library default: 64, or 128 (the
 * maximum, pf_tile_config) when all tiles of the batch are resident at once (fewer tiles than SMs: such a
 * run is bound by step latency and launch count, not by the halo's redundant cells).  The results do not
 * depend on k_block.                                                                                */
int pf_run_batch(const PfGrid *grids, int n_grids, int mode, int do_pol, int n0, const int *nsteps,
                 int k_block, void *scratch, size_t scratch_bytes, void *stream);
size_t pf_run_scratch_bytes(const PfGrid *grids, int n_grids, int engine);

/* Caller-managed ping-pong variant for long / domain-decomposed grids: ONE launch advances every grid
 * by `ksteps` steps (absolute steps n0 .. n0+ksteps-1), reading the arrays of src[m] and writing the
 * arrays of dst[m]; the caller alternates the two buffer sets and exchanges ghost cells in between
 * (pf_halo_pack / pf_halo_unpack).  `halo` (>= ksteps) is the overlap the tiles are cut with.  Arrays
 * that exist only on CPML / slab cells may be NULL for a piece that holds no such cell.  Both buffer
 * sets must start out identical.  scratch: pf_run_block_scratch_bytes() bytes (tile tables only).
 * block_flags: PF_BLOCK_F_* (0 is always valid).                                                      */
enum {
    PF_BLOCK_F_TABLES_VALID = 1, /* promise by the caller: `scratch` still holds the tile tables the previous
                                    pf_run_block call on it wrote (nothing else has written to that memory since)
                                    and they were built from these same two descriptor arrays with the same `mode`
                                    and `halo` (the tables hold the tiling and the warp classes of every tile, which
                                    depend on both) -- the library then skips the host-side rebuild, the H2D copy of
                                    the tables and their classification kernel.  The library keeps
                                    NO record of earlier calls: without this flag the tables are always rebuilt.  */
    PF_BLOCK_F_SWAPPED = 2,      /* with TABLES_VALID: src / dst are exchanged relative to the call that built the
                                    tables (the steady state of a ping-pong run alternates this bit)               */
    PF_BLOCK_F_EDGE_TILES = 4,   /* advance only the tiles that read ghost cells (the first / last tiles of a piece
                                    that has a neighbour on that side: z0 > 0, z0 + L < Lg) ...                     */
    PF_BLOCK_F_INNER_TILES = 8   /* ... or only the others.  A block is complete after both calls (any order, same
                                    arguments otherwise); the inner tiles read no ghost cell, so they can run while
                                    the ghost exchange of the block is still in flight.  Neither flag: all tiles.   */
};
int pf_run_block(const PfGrid *src, const PfGrid *dst, int n_grids, int mode, int do_pol, int n0, int ksteps, int halo,
                 int block_flags, void *scratch, size_t scratch_bytes, void *stream);
size_t pf_run_block_scratch_bytes(const PfGrid *grids, int n_grids, int halo);

/* Per-launch timing of the dominant kernels (k_tile<...>, and the PIC step's k_pic_count / k_pic_move): while enabled,
 * every such launch is bracketed by CUDA events on its own stream.  pf_profile_collect() waits for them, returns the
 * summed kernel time [ms] and the number of launches, and clears the list; pf_profile_report() does the same but
 * returns the figures per kernel name as text, "name|launches|ms;name|launches|ms;...".            */
int pf_profile_enable(int on);
int pf_profile_collect(double *ms_total, int *n_launches);
int pf_profile_report(char *report, size_t report_bytes);
/* What the FP64 pipe of the current device delivers, measured now (a few ms of GPU time): separately rounded
 * DMUL + DADD instructions per second over all SMs (the roofline of the exact-arithmetic kernels) and DFMA per
 * second (the contracted modes).  Either pointer may be NULL.                                           */
int pf_probe_fp64(double *dmul_dadd_instr_per_s, double *dfma_per_s, void *stream);

/* tile-engine introspection (bench / tests): tile width in cells, max k, threads per CTA       */
int pf_tile_config(int *tile_cells, int *k_max, int *threads);

/* ---- halo exchange support for 1-D domain decomposition (config 5) --------------------- */
/* Pack / unpack the k-cell ghost zones of every state array of grid g into / from a contiguous
 * buffer (order: Ex,Hy,Dx,P,Pprev,psiE,psiH; arrays the mode does not use are skipped).
 * side: 0 = left edge, 1 = right edge.  `interior`=1 packs the k owned cells next to the edge
 * (to send), 0 addresses the k ghost cells (to receive).  Returns doubles moved, <0 on error.  */
long long pf_halo_pack(const PfGrid *g, int mode, int side, int k, double *buf, void *stream);
long long pf_halo_unpack(const PfGrid *g, int mode, int side, int k, const double *buf, void *stream);

/* ---- PIC (builder-defined spec; the reference only has the Jx slot) ------ */
enum {
    PF_PIC_F_OFFSETS_VALID = 1 /* promise by the caller: the particle set is exactly what the previous
                                  pf_pic_push_sorted / pf_pic_step_sorted call on this scratch buffer produced
                                  (alternate arrays swapped in, nothing modified since), so the per-cell offsets
                                  that call left in the scratch are reused instead of being searched again   */
};
typedef struct PfPic {
    int64_t n;            /* macro-particles                                                     */
    int32_t L;            /* grid cells (Nz+1)                                                   */
    int32_t flags;        /* PF_PIC_F_*; 0 is always valid                                         */
    double dz, dt;        /* grid spacing / time step                                            */
    double q_over_m;      /* charge/mass of the species [C/kg]                                   */
    double c;             /* speed of light                                                      */
    double mu0;           /* By = mu0*Hy                                                         */
    double jx_scale;      /* Jx_slot contribution per particle = jx_scale * w * vx * shape       */
    /* particle SoA (device): position z [m], momenta ux = gamma*vx, uz = gamma*vz [m/s], weight */
    double *z, *ux, *uz, *w;
    int32_t *cell;        /* cell index floor(z/dz) clamped to [0, L-2], maintained by push      */
    /* output arrays of pf_pic_sort (same sizes); the caller swaps the two sets after a sort       */
    double *z_alt, *ux_alt, *uz_alt, *w_alt;
    int32_t *cell_alt;
    /* fields (device, length L): Ex at integer nodes, Hy at half nodes nz+1/2                   */
    const double *Ex, *Hy;
    double *Jx;           /* output current slot, length L (overwritten by deposit)             */
} PfPic;

/* relativistic Boris push + linear (CIC) field gather; updates z, ux, uz, cell                  */
int pf_pic_push(const PfPic *p, void *stream);
/* stable sort of the particles by cell: reads z,ux,uz,w,cell, writes the *_alt arrays;
 * scratch sized by pf_pic_scratch_bytes                                                          */
int pf_pic_sort(const PfPic *p, void *scratch, size_t scratch_bytes, void *stream);
/* push + stable re-sort fused, for a set that IS sorted by cell on entry: exploits |v| dt < dz (a
 * particle changes cell by at most one per step) -- count pass, cell scan, move pass; reads the primary
 * arrays, writes the pushed AND sorted particles to the *_alt arrays (caller swaps; the primary z/ux/uz
 * are used as staging and hold the pushed particles in their old order afterwards).  Bit-identical to
 * pf_pic_push followed by pf_pic_sort.  pf_pic_check returns 1 if a particle crossed more than one cell. */
int pf_pic_push_sorted(const PfPic *p, void *scratch, size_t scratch_bytes, void *stream);
int pf_pic_check(const PfPic *p, void *scratch, size_t scratch_bytes, void *stream);
/* deterministic cell-sorted deposition of Jx (requires particles sorted by cell)               */
int pf_pic_deposit(const PfPic *p, void *scratch, size_t scratch_bytes, void *stream);
/* One whole particle step for a cell-sorted set: pf_pic_push_sorted with the deposition of the NEW state fused
 * into its move pass (the particles are deposited while they are in registers; saves re-reading them).  Jx is
 * deterministic (fixed summation tree: per old cell and sub-warp, lane-strided in chunk order, xor butterfly, then
 * a fixed-order sum of the partial sums per node) but the tree differs from pf_pic_deposit's, so the two agree to
 * rounding (~1e-15 relative), not bit for bit; oracle/pic_oracle.py: deposit_fused().  pf_pic_sub_warps() is the
 * number of sub-warps per cell the tree is built with (a function of n / L).                        */
int pf_pic_step_sorted(const PfPic *p, void *scratch, size_t scratch_bytes, void *stream);
int pf_pic_sub_warps(const PfPic *p);
size_t pf_pic_scratch_bytes(const PfPic *p);

#ifdef __cplusplus
}
#endif
#endif /* PYFDTD_B200_H */
