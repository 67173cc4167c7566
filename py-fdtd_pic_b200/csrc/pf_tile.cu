// pf_tile.cu -- ENGINE_TILE: the fused, on-chip, temporally blocked integrator.
//
// Work decomposition
//   A launch advances every grid of a batch (sweep members, or one long grid) by up to k steps.
//   Each CTA owns one TILE of TILE_CELLS consecutive cells of one grid: TILE_CELLS - 2k interior
//   cells plus a k-cell halo on either side (overlapped / trapezoidal time blocking: after s steps
//   the s outermost cells on each side are stale, so after k steps exactly the interior is valid).
//   State is read once from HBM, advanced k steps on chip, interior written once to the other
//   ping-pong buffer.  HBM traffic per cell-update is (algorithmic bytes)/k * (1 + 2k/W).
//
// On-chip layout
//   Thread t owns the C consecutive cells [t*C, t*C+C) of the tile.  Ex and Hy of those cells
//   and all material (Dx, P, P^{n-1}) and CPML (psi_E, psi_H) state live in registers for the whole
//   launch.  Per-cell coefficients (CPML profiles b, c_e, c_m; masked update coefficients of mixed
//   warps) live in shared memory in [array][j][thread] order, so each thread only ever touches its
//   own slots (no barrier needed, lane-consecutive 8-byte words are bank-conflict free).  Warps are
//   specialised by cell class (vacuum | slab | CPML | slab+CPML | mixed).  The only inter-thread traffic is
//   one Hy value to the right neighbour before the E half-step and one Ex value to the left
//   neighbour before the H half-step, through a 2*NT-double shared edge buffer: two
//   __syncthreads per time step.
//
// Arithmetic
//   Identical, operation for operation, to the reference loop bodies (see pf_ops.cu for the
//   one-kernel-per-leaf-op statement): class A = Exact keeps every multiply and add separately
//   rounded, so results are independent of the tiling and bit-identical to ENGINE_OPS.
//
// Requirements (PF_F_CANONICAL, checked bit-for-bit by the host layer): denE = denH = UpHySelf = 1,
// bmY == beX, no Jx, UpExMat/UpHyMat two-valued (outside/inside the slab), Cb == UpExMat and
// C2 == c2_pml on the CPML correction ranges, both 0 at the single cell Lg-pw
// (the reference's exclusive range end, BaseFDTD11.py:312,326).
#include <type_traits>
#include <algorithm>
#include <atomic>
#include <cstring>
#include <mutex>
#include <vector>
#include "pf_common.cuh"

namespace pf {

#ifndef PF_TILE_CELLS
#define PF_TILE_CELLS 1024
#endif
#ifndef PF_TILE_C
#define PF_TILE_C 2
#endif
#ifndef PF_TILE_MINBLOCKS
#define PF_TILE_MINBLOCKS 2
#endif
#ifndef PF_TILE_KDEF
#define PF_TILE_KDEF 64
#endif
constexpr int TILE_CELLS = PF_TILE_CELLS;   // cells per tile (interior + 2 halos)
constexpr int TILE_KMAX = TILE_CELLS / 8;   // halo <= TILE_KMAX per side
constexpr int TILE_KDEF = PF_TILE_KDEF;     // default steps per launch

struct TileGrid {
    GridDev d;
    double *buf[2][7];   // ping-pong state: [which][Ex,Hy,psiE,psiH,Dx,P,Pprev]
    int nsteps;          // total steps this grid runs in the current pf_run_* call
    int snap_interval;   // > 0 (linear modes, single grid): the kernel itself writes Ex of the grid into snap_out[row] after every
    double *snap_out;    //      absolute step n > 0 with n % snap_interval == 0, row = n / snap_interval < snap_rows (vidMake)
    int snap_rows;
    int pad;
};
enum { S_EX = 0, S_HY, S_PSIE, S_PSIH, S_DX, S_P, S_PP, S_COUNT };

constexpr int TILE_MAX_WARPS = 32;
struct TileDesc {
    int grid;
    int base;        // local index of the tile's first cell (interior starts at base + halo)
    unsigned flags;  // TILE_F_*: written by k_tile_classify
    int cls_c;       // cells per thread the classes below were computed for
    // warp class of every warp of the tile (k_tile_classify): the per-tile prologue of k_tile reads one byte instead of
    // re-deriving the region masks of its cells on every launch (the geometry of a tile never changes)
    unsigned char cls[TILE_MAX_WARPS];
};
enum { TILE_F_SPECIAL = 1 };   // the tile holds a source cell or a probe

// shared memory carve-up (doubles): 7 per-cell coefficient arrays + edge exchange + source tables
template <int MODE, int C, class R>
struct TileSmem {
    static constexpr int NT = TILE_CELLS / C;
    static constexpr int N_ARR = 7;
    // + warp-exchange area (PF_WARP_XCHG): two mbarriers (8 B each) per warp
    static constexpr size_t bytes = sizeof(R) * ((size_t)N_ARR * TILE_CELLS + 2 * NT + 2 + 2 * TILE_KMAX) + 16 + 16 * (NT / 32 + 1);
};

// Shared memory is addressed as pf_smem[offset + index] with plain integer offsets: going through
// generic double* members made the compiler re-derive the shared window (S2R SR_CgaCtaId + LEA) after
// every barrier, on the critical path (ncu r1_final: ~10 % of the hot loop's stall samples).
extern __shared__ double pf_smem[];

// One shared-memory slot of the working precision R addressed in the shared state space (ld.shared /
// st.shared with a 32-bit address): reads and writes look like array accesses at the call sites.
template <class R>
struct SmemSlot;
template <>
struct SmemSlot<double> {
    unsigned addr;
    __device__ __forceinline__ operator double() const
    {
        double v;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
        return v;
    }
    __device__ __forceinline__ void operator=(double v) const
    {
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
    }
};
template <>
struct SmemSlot<float> {
    unsigned addr;
    __device__ __forceinline__ operator float() const
    {
        float v;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
        return v;
    }
    __device__ __forceinline__ void operator=(float v) const
    {
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
    }
};
template <class R>
struct SmemArray {
    unsigned base;   // byte address in the shared window
    __device__ __forceinline__ SmemSlot<R> operator[](int i) const { return SmemSlot<R>{base + (unsigned)sizeof(R) * (unsigned)i}; }
};

template <class R>
struct TileShared {
    // per-cell coefficient slots, [j][thread] order (each thread touches only its own slots)
    SmemArray<R> be, ce, cm;      // CPML recursive-convolution profiles (0 outside the CPML)
    SmemArray<R> cEu, cHu;        // update coefficient of the cell, 0 where the field is never updated
    SmemArray<R> cb, c2u;         // CPML field-correction coefficients, 0 outside / at the quirk cell
    SmemArray<R> edgeH, edgeE, srcE, srcH;
    unsigned xw;                  // PF_WARP_XCHG: byte address of the per-warp exchange area (16-byte aligned)
};

// ---- neighbour-warp exchange (PF_WARP_XCHG) ---------------------------------------------------------
// The one Hy / Ex value a thread needs from its neighbour thread travels through the shared edge arrays exactly as
// in the barrier scheme; what changes is who waits for whom.  Inside a warp a __syncwarp orders the store and the
// neighbour lane's load.  Between adjacent warps the value is handed over with an mbarrier of arrival count 1
// (producer lane: store, __syncwarp, arrive(release); consumer warp: try_wait(acquire) on the phase parity, load).
// The two directions of a warp boundary alternate strictly (H_init, E_0, H_0, E_1, ...), so neither warp can run
// more than one phase ahead of the other: the single edge slot is never overwritten before it was read, one parity
// bit per direction tracks the phase, and no CTA-wide barrier is left in the time loop.  Warps whose neighbour
// holds no grid cell (dead warp / tile end) neither wait nor arrive on that side; the slot they read stays 0.
// Exchange area: per warp w 16 bytes = [mbarH u64][mbarE u64].
__device__ __forceinline__ unsigned xw_barH(unsigned xw, int w) { return xw + 16u * (unsigned)w; }
__device__ __forceinline__ unsigned xw_barE(unsigned xw, int w) { return xw + 16u * (unsigned)w + 8u; }
__device__ __forceinline__ void xw_init(unsigned bar)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void xw_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// orders a lane's edge store before the neighbour lane's load (and before the publishing lane's arrive)
__device__ __forceinline__ void xw_syncwarp()
{
#ifndef PF_WX_NOSYNCWARP   // timing experiment: cost of the convergence check the compiler wraps around bar.warp.sync
    __syncwarp();
#endif
}
__device__ __forceinline__ void xw_wait(unsigned bar, unsigned parity)
{
    // every lane of the warp tests the same barrier in the same instruction: the loop branch is uniform
    asm volatile("{\n.reg .pred p;\nXW_WAIT:\nmbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n@!p bra.uni XW_WAIT;\n}"
                 ::"r"(bar), "r"(parity) : "memory");
}
// what a warp knows about its two neighbours (warp-uniform)
struct WarpLink {
    bool sendH, sendE;     // this LANE publishes the warp's Hy edge to the right (lane 31) / Ex edge to the left (lane 0)
    unsigned pH, pE;       // phase parity of the next H / E hand-over this warp WAITS for
    unsigned fH, fE;       // 1: the parity alternates (live neighbour); 0: no neighbour on that side -- the wait is on the
                           // CTA's dummy barrier with parity 1 (the phase before the first: always complete), no branch
    unsigned barL, barR;   // mbarrier this warp waits on: the left neighbour's H barrier, the right neighbour's E barrier
    unsigned barS;         // this warp's own H barrier (its E barrier is at +8)
};

// CTA-wide barrier usable from warp-uniform divergent code: every warp executes exactly two of
// these per time step (plus one before the loop), whichever body it runs.
#ifdef PF_EXPERIMENT_NO_SYNC   // timing experiment only (results are wrong): upper bound of a barrier-free design
__device__ __forceinline__ void cta_sync() { __syncwarp(); }
#else
__device__ __forceinline__ void cta_sync() { asm volatile("bar.sync 0;" ::: "memory"); }
#endif

// per-thread description of its C cells
struct CellMasks {
    unsigned valid, updE, updH, pmlE, pmlH, slab, store, quirk;
    int jsrc, jtfsf, pj0, pj1;
    size_t po0, po1;
};

// Masks of the C cells [lz0, lz0+C) of one thread: the index rules of the reference loops (update ranges, CPML ranges with
// the exclusive-end quirk cell, slab, source cells) plus which cells this tile stores (its interior).
template <int C>
__device__ __forceinline__ void tile_cell_masks(const PfGrid &g, int tile_base, int lz0, int halo, CellMasks &M)
{
    const int L = g.L;
    const int z0 = (int)g.z0, Lg = (int)g.Lg;
    const int pw = g.pw, mf = g.mf, mr = g.mr;
    const int flags = g.flags;
    M.valid = M.updE = M.updH = M.pmlE = M.pmlH = M.slab = M.store = M.quirk = 0;
    M.jsrc = M.jtfsf = M.pj0 = M.pj1 = -1;
    M.po0 = M.po1 = 0;
    const bool cpml_m = flags & PF_F_CPML_M, cpml_p = flags & PF_F_CPML_P;
#pragma unroll
    for (int j = 0; j < C; ++j) {
        int lz = lz0 + j, gz = z0 + lz;
        bool valid = lz >= 0 && lz < L;
        bool updE = valid && lz >= 1 && gz >= 1 && gz <= Lg - 1;
        bool updH = valid && lz <= L - 2 && gz >= 1 && gz <= Lg - 2;
        bool inL = cpml_m && gz < pw, inR = cpml_p && gz >= Lg - pw;
        bool slab = valid && lz >= 1 && gz >= mf && gz < mr;
        bool interior = (lz - tile_base) >= halo && (lz - tile_base) < TILE_CELLS - halo;
        M.valid |= (unsigned)valid << j;
        M.updE |= (unsigned)updE << j;
        M.updH |= (unsigned)updH << j;
        M.pmlE |= (unsigned)(updE && (inL || inR)) << j;
        M.pmlH |= (unsigned)(updH && (inL || inR)) << j;
        M.slab |= (unsigned)slab << j;
        M.store |= (unsigned)(valid && interior) << j;
        M.quirk |= (unsigned)(gz == Lg - pw) << j;
    }
}

// masks of a thread whose warp is entirely inside one region (warp classes 0..3): only the store mask is ever looked at
template <int C>
__device__ __forceinline__ void tile_plain_masks(int tid, int halo, int cls, CellMasks &M)
{
    constexpr unsigned ALL = (1u << C) - 1u;
    M.valid = M.updE = M.updH = (cls == 5) ? 0u : ALL;
    M.pmlE = M.pmlH = (cls & 2) && cls != 5 ? ALL : 0u;
    M.slab = (cls & 1) && cls != 5 ? ALL : 0u;
    M.quirk = 0;
    M.store = 0;
    M.jsrc = M.jtfsf = M.pj0 = M.pj1 = -1;
    M.po0 = M.po1 = 0;
#pragma unroll
    for (int j = 0; j < C; ++j) M.store |= (unsigned)(tid * C + j >= halo && tid * C + j < TILE_CELLS - halo) << j;
}

// source cells and probes among the thread's cells (at most 2 probes per thread; guaranteed by the host layer);
// returns true if the thread owns any
template <int C>
__device__ __forceinline__ bool tile_special_cells(const PfGrid &g, int lz0, CellMasks &M)
{
    const int z0 = (int)g.z0;
    {
        const int j = g.nzsrc - z0 - lz0;
        if (j >= 0 && j < C && lz0 + j >= 0 && lz0 + j < g.L) M.jsrc = j;
        if ((g.flags & PF_F_TFSF) && j - 1 >= 0 && j - 1 < C && lz0 + j - 1 >= 0 && lz0 + j - 1 < g.L) M.jtfsf = j - 1;
    }
    for (int p = 0; p < g.n_probes; ++p) {
        int j = g.probe_idx[p] - z0 - lz0;
        if (j >= 0 && j < C && ((M.store >> j) & 1)) {
            if (M.pj0 < 0) { M.pj0 = j; M.po0 = (size_t)p * g.probe_stride; }
            else { M.pj1 = j; M.po1 = (size_t)p * g.probe_stride; }
        }
    }
    return M.jsrc >= 0 || M.jtfsf >= 0 || M.pj0 >= 0;
}

// warp class: 0 vacuum, 1 slab, 2 CPML, 3 slab+CPML, 4 mixed, 5 dead (warp-uniform result)
// (a source cell inside a material-law cell would be overwritten anyway; it is sent to the
//  mixed body only to keep the fast slab body free of the test)
template <int C>
__device__ __forceinline__ int tile_warp_class(const CellMasks &M)
{
    constexpr unsigned ALL = (1u << C) - 1u;
    int cls = 4;
    const bool plain = M.valid == ALL && M.updE == ALL && M.updH == ALL && M.quirk == 0;
    const bool pmlAll = M.pmlE == ALL && M.pmlH == ALL, pmlNone = (M.pmlE | M.pmlH) == 0;
    const bool slabAll = M.slab == ALL, slabNone = M.slab == 0;
    if (plain && (pmlAll || pmlNone) && (slabAll || slabNone)) cls = (slabAll ? 1 : 0) + (pmlAll ? 2 : 0);
    if (M.valid == 0) cls = 5;                    // cells beyond the end of the grid
    const int cls0 = __shfl_sync(0xffffffffu, cls, 0);
    return __all_sync(0xffffffffu, cls == cls0) ? cls0 : 4;
}

// -------------------------------------------------------------------------------------------------
// One body for every warp class.  All field / material / CPML state of the thread's C cells lives in
// registers for the whole launch.
//   GEN = false : every cell of every thread of the warp is a plain interior cell of one class
//                 (SLAB? x PML?) -> straight-line code, no masks, scalar coefficients.
//   GEN = true  : mixed warp (region boundary, grid end, CPML quirk cell).  Same arithmetic with
//                 every term switched on, per-cell coefficients from shared memory: a coefficient of
//                 exactly 0 turns its term into an exact no-op (x + y*0 == x), so no per-cell branch
//                 is needed; only the material law is selected per cell.
// -------------------------------------------------------------------------------------------------
// Where a warp keeps the CPML profiles b, c_e, c_m of its cells: mixed warps in their shared-memory slots, every uniform
// class in registers.  -DPF_PML_COEF_SMEM=1 moves them to shared memory for the slab+CPML body too (7 state arrays, 3
// profiles and the step constants are more than the 64-register budget holds comfortably): measured 546 against 597
// Gcell-updates/s on the sweep, the loads land on the step's dependency chain (profiles/r2_tile_experiments.md).
#ifndef PF_PML_COEF_SMEM
#define PF_PML_COEF_SMEM 0
#endif
template <int MODE, int C, bool GEN, bool SLAB, bool PML, class R>
__device__ __forceinline__ constexpr bool pml_coef_in_smem()
{
    return GEN || (PF_PML_COEF_SMEM && PML && SLAB && MODE != PF_FREE && C == 2 && std::is_same<R, double>::value);
}

// Everything a time step needs besides the per-cell register arrays.
template <class R>
struct StepConsts {
    R cEs, cHs, c2s, dtdz, eps0, inv_eps0, pA, pB, pC, den0, den1;
    // fp32 mode works in scaled variables (H/cH0, psi_E/cH0, D/eps0, P/eps0: everything O(E), the H update
    // coefficient exactly 1) and carries the two coefficients that set the numerical wave speed as
    // hi + lo pairs, so that no coefficient rounding accumulates as a phase drift.
    R cEs_lo, dtdz_lo;
    R pG, pK;            // fp32 mode: Lorentz ADE in difference form, G = 1 + B, K = 1 - A - B
    R ca, cb, cc, inv_cc;   // fp32 / Newton modes: cubic coefficients (and 1/c) in registers
    NlFastConsts kf;        // closed-form cubic law, fast path (fp64 only)
    bool kf_ok;             // the fast path applies to this grid (cub != 0: a genuine cubic)
    int jsrc, jtfsf;
    bool wSrc;
    unsigned mSlab;
};

// PF_F_FP32 material law of one nonlinear cell.  The closed form the reference uses
// (CubicEquationSolver.py:29-105) cancels ~4 digits in (S+U) - b/3a, which single precision cannot
// afford, so the positive root of a x^3 + b x^2 + c x - q^2 (a, b >= 0, c > 0: increasing and convex
// for x > 0) is found by Newton iteration from x0 = q^2/c >= root, which converges monotonically.
struct NlResultF {
    float a, e;
};
__device__ __forceinline__ NlResultF nl_material_law_f32(float ca, float cb, float cc, float dx, float inv_eps0,
                                                        float den0, float den1)
{
    NlResultF r;
    const float q = dx * inv_eps0;   // inv_eps0 = 1 in the scaled variables of the fp32 mode
    const float d = q * q;
    float x = 0.f;
    if (d > 1e-8f) {
        x = __fdividef(d, cc);
        float step;
        int it = 0;
        do {   // two iterations in the weakly nonlinear regime of the sweeps; ~log_1.5(x0/root) more when x0 is far above
            const float p = fmaf(fmaf(fmaf(ca, x, cb), x, cc), x, -d);
            const float dp = fmaf(fmaf(3.f * ca, x, 2.f * cb), x, cc);
            step = __fdividef(p, dp);
            x -= step;
        } while (++it < 48 && fabsf(step) > 1e-6f * x);
    }
    r.a = x;
    r.e = __fdividef(dx, fmaf(den1, x, den0));
    return r;
}

// One full time step (E half-step, barrier, H half-step) on the thread's C cells.
//   pc = P^n (current polarisation), pq = P^{n-1}: the new P^{n+1} is written over pq, so the
//   caller alternates (pc,pq) <-> (pq,pc) instead of shifting the history (no register moves).
template <int MODE, bool POL, int C, class A, bool GEN, bool SLAB, bool PML, bool SP, bool JX, class R = typename A::real>
__device__ __forceinline__ void tile_step(const TileShared<R> &S, const StepConsts<R> &K, const CubicConsts *kcp, WarpLink &W, int tid, int s, bool more,
                                          R (&ex)[C], R (&hy)[C], R (&dx)[C], R (&pc)[C],
                                          R (&pq)[C], R (&pe)[C], R (&ph)[C], R (&acub)[C],
                                          R (&rbe)[C], R (&rce)[C], R (&rcm)[C], const R (&jx)[C])
{
    constexpr int NT = TILE_CELLS / C;
    constexpr bool F32 = std::is_same<R, float>::value;
    constexpr bool LOR = MODE == PF_LORENTZ || MODE == PF_LORENTZ_NL;   // Lorentz ADE polarisation
    constexpr bool HAS_MAT = (GEN || SLAB) && MODE != PF_FREE;
    constexpr bool ALL_MAT = !GEN && SLAB && MODE != PF_FREE;
    constexpr bool HAS_PML = GEN || PML;
    constexpr bool PS = pml_coef_in_smem<MODE, C, GEN, SLAB, PML, R>();
    // ===== E half-step: history shift + polarisation, ADE_ExUpdate, CPML_Psi_e, source,
    //                    ADE_DxUpdate, ADE_ExCreate | AcubicFinder + NonLinExUpdate =====
#ifdef PF_WARP_XCHG
    xw_wait(W.barL, W.pH);   // the left neighbour warp's Hy edge of the previous step (or H_init)
    W.pH ^= W.fH;
#endif
    R hl = S.edgeH[tid - 1];
    unsigned divkey = 0;
    unsigned nlbad = 0;   // closed-form cubic law: cells whose fast path has to be redone by the general law
#pragma unroll
    for (int j = 0; j < C; ++j) {
        const R dH = A::sub(hy[j], hl);
        // JX: the current slot (PIC) enters the bracket of ADE_ExUpdate / ADE_DxUpdate, (Hy[nz] - Hy[nz-1] - Jx[nz]); the
        // CPML convolution below keeps the plain difference (CPML_Psi_e_Update has no current term)
        const R dHJ = JX ? A::sub(dH, jx[j]) : dH;
        hl = hy[j];
        R e = ex[j];
        R pnow = R(0);
        R vnew = R(0);   // fp32 mode: P^{n+1} - P^n of this step
        if (LOR && HAS_MAT) {
            if (POL) {
                if constexpr (F32) {
                    // difference form: pq holds v = P^n - P^{n-1};  v' = v - G v - K P + C E,  P' = P + v'
                    R v = pq[j];
                    v = A::nmad(K.pG, v, v);
                    v = A::nmad(K.pK, pc[j], v);
                    v = A::mad(K.pC, e, v);
                    pq[j] = v;
                    pc[j] = A::add(pc[j], v);
                    pnow = pc[j];
                    vnew = v;
                } else {
#ifndef PF_PHASE_BALANCE
                    pq[j] = A::add(A::add(A::mul(K.pA, pc[j]), A::mul(K.pB, pq[j])), A::mul(K.pC, e));
#endif
                    // PF_PHASE_BALANCE: P^{n+1} was already written to pq behind the previous step's H half-step (below)
                    pnow = pq[j];
                }
            } else {
                pnow = pc[j];
            }
        }
        if (!ALL_MAT) {
            if constexpr (F32 && !GEN) e = A::add(e, A::mad(dH, K.cEs, A::mul(dH, K.cEs_lo)));
            else e = A::mad(dHJ, GEN ? (R)S.cEu[j * NT + tid] : K.cEs, e);
        }
        if (HAS_PML) {
            const R b = PS ? (R)S.be[j * NT + tid] : rbe[j];
            const R c = PS ? (R)S.ce[j * NT + tid] : rce[j];
            const R psi = A::mad(b, pe[j], A::mul(c, dH));
            pe[j] = psi;
            if (!ALL_MAT) e = A::nmad(GEN ? (R)S.cb[j * NT + tid] : K.cEs, psi, e);
        }
        if (HAS_MAT) {
            if (MODE == PF_LORENTZ) {
                R em;
                if constexpr (F32) {
                    // fp32 mode carries Dn = D - P instead of D: Dn' = Dn + dD - (P' - P).  D and P are each a few
                    // times E, so forming D - P every step would amplify their accumulated rounding (x4.8 at 9 GHz)
                    dx[j] = A::add(dx[j], A::sub(A::mad(dH, K.dtdz, A::mul(dH, K.dtdz_lo)), vnew));
                    em = dx[j];
                } else {
                    dx[j] = A::mad(dHJ, K.dtdz, dx[j]);
                    em = div_const_fast(A::sub(dx[j], pnow), K.eps0, K.inv_eps0, divkey);
                }
                e = (!GEN || ((K.mSlab >> j) & 1)) ? em : e;
            } else if constexpr (F32) {   // cubic law on Dx (PF_NL) or on Dx - P (PF_LORENTZ_NL)
                if (!GEN || ((K.mSlab >> j) & 1)) {
                    dx[j] = A::add(dx[j], A::sub(A::mad(dH, K.dtdz, A::mul(dH, K.dtdz_lo)), vnew));   // Dn (LOR) or D
                    const R dn = dx[j];
                    const NlResultF nl = nl_material_law_f32(K.ca, K.cb, K.cc, dn, K.inv_eps0, K.den0, K.den1);
                    acub[j] = nl.a;
                    e = nl.e;
                }
            } else if constexpr (A::newton) {   // PF_F_NEWTON: small enough to be inlined per cell
                if (!GEN || ((K.mSlab >> j) & 1)) {
                    dx[j] = A::mad(dHJ, K.dtdz, dx[j]);
                    const R dn = LOR ? A::sub(dx[j], pnow) : dx[j];
                    nl_material_law_newton(NlNewtonConsts{K.ca, K.cb, K.cc, K.inv_cc}, dn, K.inv_eps0, K.den0, K.den1, acub[j], e);
                }
            } else {
#ifdef PF_NL_OUT_OF_LINE     // round-1 arrangement: the closed-form law as an out-of-line call (kept for A/B timing)
                if (!GEN) {
                    dx[j] = A::mad(dHJ, K.dtdz, dx[j]);      // the material law of all C cells follows the loop
                } else if ((K.mSlab >> j) & 1) {
                    dx[j] = A::mad(dHJ, K.dtdz, dx[j]);
                    const NlResult nl = nl_material_law(kcp, LOR ? A::sub(dx[j], pnow) : dx[j], K.eps0, K.inv_eps0, K.den0, K.den1);
                    acub[j] = nl.a;
                    e = nl.e;
                }
#else
                if (!GEN || ((K.mSlab >> j) & 1)) {          // closed form, inlined fast path (pf_common.cuh)
                    dx[j] = A::mad(dHJ, K.dtdz, dx[j]);
                    const R dn = LOR ? A::sub(dx[j], pnow) : dx[j];
                    const bool ok = K.kf_ok && nl_material_law_fast(K.kf, dn, K.eps0, K.inv_eps0, K.den0, K.den1, acub[j], e);
                    nlbad |= (ok ? 0u : 1u) << j;
                }
#endif
            }
        }
        ex[j] = e;
    }
#ifdef PF_NL_OUT_OF_LINE
    if constexpr ((MODE == PF_NL || MODE == PF_LORENTZ_NL) && ALL_MAT && !F32 && !A::newton) {
        NlVec<C> dv;
#pragma unroll
        for (int j = 0; j < C; ++j) dv.v[j] = LOR ? A::sub(dx[j], POL ? pq[j] : pc[j]) : dx[j];
        const NlResultVec<C> nl = nl_material_law_vec<C>(kcp, dv, K.eps0, K.inv_eps0, K.den0, K.den1);
#pragma unroll
        for (int j = 0; j < C; ++j) { acub[j] = nl.a[j]; ex[j] = nl.e[j]; }
    }
#else
    if constexpr ((MODE == PF_NL || MODE == PF_LORENTZ_NL) && HAS_MAT && !F32 && !A::newton) {
        // cells the fast path declined (other branch of the cubic, out-of-range operand): the general out-of-line law
        if (__any_sync(0xffffffffu, nlbad != 0u)) {
#pragma unroll
            for (int j = 0; j < C; ++j) {
                if ((nlbad >> j) & 1) {
                    const NlResult nl = nl_material_law(kcp, LOR ? A::sub(dx[j], POL ? pq[j] : pc[j]) : dx[j], K.eps0, K.inv_eps0, K.den0, K.den1);
                    acub[j] = nl.a;
                    ex[j] = nl.e;
                }
            }
        }
    }
#endif
    // warp-uniform test: a per-lane branch here opens a divergent region, which costs the uniform
    // registers holding the shared-memory window (re-read with S2UR after every step)
#ifdef PF_EXPERIMENT_NO_GUARD   // timing experiment only: cost of the division-range guard
    if (false) {
#else
    if (!F32 && MODE == PF_LORENTZ && HAS_MAT && __any_sync(0xffffffffu, !div_const_in_range(divkey))) {
#endif
        // some cell's (Dx - P) was zero, in the denormal range or non-finite: redo those divisions
        // exactly (rare once the wave has arrived; integer tests only for the zero case)
#pragma unroll
        for (int j = 0; j < C; ++j) {
            if constexpr (!F32) {
                if (!GEN || ((K.mSlab >> j) & 1)) {
                    const double x = A::sub(dx[j], POL ? pq[j] : pc[j]);
                    ex[j] = div_const_fix(x, K.eps0, K.inv_eps0, ex[j]);
                }
            }
        }
    }
    if (!ALL_MAT && SP && K.wSrc) {
        // soft source (Solver_Engine.py:307): the last E operation of a non-material cell; a source
        // inside a material-law cell is overwritten by ADE_ExCreate in the reference too
#pragma unroll
        for (int j = 0; j < C; ++j)
            if (j == K.jsrc && !(HAS_MAT && ((K.mSlab >> j) & 1))) ex[j] = A::add(ex[j], S.srcE[s]);
    }
    S.edgeE[tid] = ex[0];
#ifdef PF_WARP_XCHG
    xw_syncwarp();
    if (W.sendE) xw_arrive(W.barS + 8u);   // lane 0 of a warp with a live left neighbour
#else
    cta_sync();
#endif
    if (SP && K.wSrc) {   // one-point TF/SF correction (Solver_Engine.py:309-310), before ADE_HyUpdate
#pragma unroll
        for (int j = 0; j < C; ++j)
            if (j == K.jtfsf) hy[j] = A::sub(hy[j], S.srcH[s]);
    }

    // ===== H half-step: TF/SF correction, ADE_HyUpdate, CPML_Psi_m =====
    // the neighbour's Ex is requested first and consumed last (cell C-1), behind the cells that only
    // need the thread's own Ex, so the shared-memory latency is covered
#ifdef PF_WARP_XCHG
    xw_wait(W.barR, W.pE);   // the right neighbour warp's Ex edge of this step
    W.pE ^= W.fE;
#endif
    const R exr = S.edgeE[tid + 1];
#pragma unroll
    for (int j = 0; j < C; ++j) {
        R h = hy[j];
        const R dE = A::sub((j == C - 1) ? exr : ex[(j + 1) % C], ex[j]);
        h = A::mad(dE, GEN ? (R)S.cHu[j * NT + tid] : K.cHs, h);
        if (HAS_PML) {
            const R b = PS ? (R)S.be[j * NT + tid] : rbe[j];
            const R c = PS ? (R)S.cm[j * NT + tid] : rcm[j];
            const R psi = A::mad(b, ph[j], A::mul(c, dE));
            ph[j] = psi;
            h = A::mad(GEN ? (R)S.c2u[j * NT + tid] : K.c2s, psi, h);
        }
        hy[j] = h;
    }
    S.edgeH[tid] = hy[C - 1];
#ifdef PF_WARP_XCHG
    xw_syncwarp();
    if (W.sendH) xw_arrive(W.barS);        // lane 31 of a warp with a live right neighbour
#endif
#ifdef PF_PHASE_BALANCE
    // The polarisation update of the NEXT step (ADE_PolarisationCurrent_Ex: P^{n+2} from P^{n+1}, P^n and the Ex this step
    // just produced) is issued here, behind the H half-step, instead of at the top of the next E half-step: the same
    // operations on the same values, but the fp64 work is now split 14 : 16 between the two barrier-separated phases of a
    // step instead of 24 : 6, so two co-resident CTAs load the FP64 pipe evenly whatever their relative phase.
    // Roles are those of the next step: its current P is this step's pq, its previous P (overwritten) this step's pc.
    if constexpr (LOR && HAS_MAT && POL && !F32) {
        if (more) {
#pragma unroll
            for (int j = 0; j < C; ++j)
                pc[j] = A::add(A::add(A::mul(K.pA, pq[j]), A::mul(K.pB, pc[j])), A::mul(K.pC, ex[j]));
        }
    }
#endif
}

// -------------------------------------------------------------------------------------------------
// One body for every warp class.  All field / material / CPML state of the thread's C cells lives in
// registers for the whole launch.
//   GEN = false : every cell of every thread of the warp is a plain interior cell of one class
//                 (SLAB? x PML?) -> straight-line code, no masks, scalar coefficients, CPML
//                 profiles in registers.
//   GEN = true  : mixed warp (region boundary, grid end, CPML quirk cell).  Same arithmetic with
//                 every term switched on, per-cell coefficients from shared memory: a coefficient of
//                 exactly 0 turns its term into an exact no-op (x + y*0 == x), so no per-cell branch
//                 is needed; only the material law is selected per cell.
// -------------------------------------------------------------------------------------------------
template <int MODE, bool POL, int C, class A, bool GEN, bool SLAB, bool PML, bool JX, class R = typename A::real>
__device__ __forceinline__ void tile_body(const TileGrid &TG, const TileShared<R> &S, const CellMasks &M, WarpLink W,
                                          int tid, int lz0, int ks, int src, int nabs0,
                                          const TileGrid *__restrict__ grids, const TileDesc *__restrict__ tiles, int halo)
{
    constexpr int NT = TILE_CELLS / C;
    constexpr bool F32 = std::is_same<R, float>::value;
    constexpr bool LOR = MODE == PF_LORENTZ || MODE == PF_LORENTZ_NL;
    constexpr bool CUB = MODE == PF_NL || MODE == PF_LORENTZ_NL;
    constexpr bool HAS_MAT = (GEN || SLAB) && MODE != PF_FREE;   // material arrays present
    constexpr bool HAS_PML = GEN || PML;
    const PfGrid &g = TG.d.g;
    const unsigned mSlab = M.slab;
    // scaled variables of the fp32 mode (identity scales otherwise): H' = H sH, psi_E' = psi_E sH, D' = D sD, P' = P sD
    const double sH = F32 ? 1.0 / g.cH0 : 1.0, uH = F32 ? g.cH0 : 1.0;
    const double sD = F32 ? TG.d.inv_eps0 : 1.0, uD = F32 ? g.eps0 : 1.0;
    R ex[C], hy[C], dx[C], pa[C], pb[C], pe[C], ph[C], acub[C], rbe[C], rce[C], rcm[C], jx[C];

    // ---- load ---------------------------------------------------------------------------------
    {
        const double *__restrict__ inEx = TG.buf[src][S_EX];
        const double *__restrict__ inHy = TG.buf[src][S_HY];
#pragma unroll
        for (int j = 0; j < C; ++j) {
            bool v = !GEN || ((M.valid >> j) & 1);
            ex[j] = v ? inEx[lz0 + j] : 0.0;
            hy[j] = F32 ? (v ? inHy[lz0 + j] * sH : 0.0) : (v ? inHy[lz0 + j] : 0.0);
        }
        if (JX) {   // current slot: constant over the steps of a launch (a PIC-coupled run launches one step at a time)
#pragma unroll
            for (int j = 0; j < C; ++j) {
                const bool v = (!GEN || ((M.valid >> j) & 1)) && g.Jx != nullptr;
                jx[j] = v ? (R)g.Jx[lz0 + j] : R(0);
            }
        }
        if (HAS_PML) {
            const double *__restrict__ inPe = TG.buf[src][S_PSIE];
            const double *__restrict__ inPh = TG.buf[src][S_PSIH];
#pragma unroll
            for (int j = 0; j < C; ++j) {
                bool e_ = !GEN || ((M.pmlE >> j) & 1), h_ = !GEN || ((M.pmlH >> j) & 1);
                int lz = lz0 + j;
                pe[j] = F32 ? (e_ ? inPe[lz] * sH : 0.0) : (e_ ? inPe[lz] : 0.0);
                ph[j] = h_ ? inPh[lz] : 0.0;
                const double b = (e_ || h_) ? g.beX[lz] : 0.0, c1 = e_ ? g.ceX[lz] : 0.0, c2 = h_ ? g.cmY[lz] : 0.0;
                if (pml_coef_in_smem<MODE, C, GEN, SLAB, PML, R>()) {
                    S.be[j * NT + tid] = (R)b;
                    S.ce[j * NT + tid] = (R)c1;
                    S.cm[j * NT + tid] = (R)c2;
                } else {
                    rbe[j] = b; rce[j] = c1; rcm[j] = c2;
                }
            }
        }
        if (HAS_MAT) {
            const double *__restrict__ inDx = TG.buf[src][S_DX];
#pragma unroll
            for (int j = 0; j < C; ++j) {
                const double d0 = (!GEN || ((mSlab >> j) & 1)) ? inDx[lz0 + j] : 0.0;
                dx[j] = F32 ? d0 * sD : d0;
            }
            if (LOR) {
                const double *__restrict__ inP = TG.buf[src][S_P];
                const double *__restrict__ inPp = TG.buf[src][S_PP];
#pragma unroll
                for (int j = 0; j < C; ++j) {
                    bool sl = !GEN || ((mSlab >> j) & 1);
                    const double p0 = sl ? inP[lz0 + j] : 0.0, p1 = sl ? inPp[lz0 + j] : 0.0;
                    pa[j] = F32 ? p0 * sD : p0;
                    pb[j] = F32 ? (p0 - p1) * sD : p1;   // fp32 mode carries P^n - P^{n-1} (difference form)
                    if (F32) dx[j] = sl ? (inDx[lz0 + j] - p0) * sD : 0.0;   // ... and Dn = D - P in place of D
                }
            }
        }
        if (GEN) {
#pragma unroll
            for (int j = 0; j < C; ++j) {
                const bool sl = (mSlab >> j) & 1, q = (M.quirk >> j) & 1;
                const double cE = (sl ? g.cE1 : g.cE0) * uH, cH = (sl ? g.cH1 : g.cH0) * sH;
                S.cEu[j * NT + tid] = (R)(((M.updE >> j) & 1) ? cE : 0.0);
                S.cHu[j * NT + tid] = (R)(((M.updH >> j) & 1) ? cH : 0.0);
                S.cb[j * NT + tid] = (R)((((M.pmlE >> j) & 1) && !q) ? cE : 0.0);
                S.c2u[j * NT + tid] = (R)((((M.pmlH >> j) & 1) && !q) ? g.c2_pml * sH : 0.0);
            }
        }
        if (CUB) {
#pragma unroll
            for (int j = 0; j < C; ++j) acub[j] = R(0);
        }
    }
    StepConsts<R> K;
    {
        const double cE = (SLAB ? g.cE1 : g.cE0) * uH, dd = g.dt_over_dz * uH * sD;
        K.cEs = cE; K.cHs = (SLAB ? g.cH1 : g.cH0) * sH; K.c2s = g.c2_pml * sH;
        K.dtdz = dd; K.eps0 = g.eps0; K.inv_eps0 = F32 ? 1.0 : TG.d.inv_eps0;
        K.cEs_lo = (R)(cE - (double)K.cEs); K.dtdz_lo = (R)(dd - (double)K.dtdz);
        K.pA = g.polA; K.pB = g.polB; K.pC = g.polC * sD; K.den0 = g.nl_den0 * sD; K.den1 = g.nl_den1 * sD;
    }
    K.pG = (R)(1.0 + g.polB); K.pK = (R)((1.0 - g.polA) - g.polB);
    K.ca = (R)g.cub_a; K.cb = (R)g.cub_b; K.cc = (R)g.cub_c; K.inv_cc = (R)TG.d.k.inv_c;
    K.kf_ok = TG.d.k.a != 0.0;
    K.kf.a = TG.d.k.a; K.kf.inv_a = TG.d.k.inv_a; K.kf.g_ab = TG.d.k.g_ab; K.kf.f3_27 = TG.d.k.f3_27; K.kf.b_3a = TG.d.k.b_3a;
    K.jsrc = M.jsrc; K.jtfsf = M.jtfsf; K.mSlab = mSlab;
    K.wSrc = __any_sync(0xffffffffu, M.jsrc >= 0 || M.jtfsf >= 0);
    const CubicConsts *kc = &TG.d.k;
    const int pj0 = M.pj0, pj1 = M.pj1;
    const bool wProbe = __any_sync(0xffffffffu, pj0 >= 0);
    // snapshot rows (vidMake), linear modes: every warp of a launch that holds a snapshot step runs the loop copy with the
    // per-step tests.  (The cubic modes have one loop copy only and keep ending a launch at each snapshot step instead: the
    // test costs their 150-instruction material law 13 %.)
    int snapS = -1;              // launch-relative step whose result is the next snapshot row
    bool wSnap = false;
    if constexpr (!CUB) {
        const int snapI = TG.snap_interval;
        if (snapI > 0) {
            snapS = (snapI - nabs0 % snapI) % snapI;
            if (nabs0 + snapS == 0) snapS += snapI;
            wSnap = snapS < ks;
        }
    }

    auto probes = [&](int s) {   // Solver_Engine.probeSim: Ex after the step
        if constexpr (!CUB) {
            if (wSnap && s == snapS) {
                const int row = (nabs0 + s) / TG.snap_interval;
                if (row < TG.snap_rows) {
                    double *__restrict__ out = TG.snap_out + (size_t)row * g.L;
#pragma unroll
                    for (int j = 0; j < C; ++j)
                        if ((M.store >> j) & 1) out[lz0 + j] = ex[j];
                }
                snapS += TG.snap_interval;
            }
        }
        if (wProbe && pj0 >= 0) {
            R v = R(0);
#pragma unroll
            for (int j = 0; j < C; ++j) v = (j == pj0) ? ex[j] : v;
            g.probe_out[M.po0 + nabs0 + s] = v;
            if (pj1 >= 0) {
#pragma unroll
                for (int j = 0; j < C; ++j) v = (j == pj1) ? ex[j] : v;
                g.probe_out[M.po1 + nabs0 + s] = v;
            }
        }
    };

    S.edgeH[tid] = hy[C - 1];
#ifdef PF_WARP_XCHG
    xw_syncwarp();
    if (W.sendH) xw_arrive(W.barS);   // H_init: the first E half-step of the right neighbour warp needs it
#else
    cta_sync();
#endif
    constexpr bool SWAP = LOR && HAS_MAT && POL && !F32;   // P history alternates between pa and pb
    bool swapped = false;   // true: current P is in pb, previous in pa
#ifdef PF_PHASE_BALANCE
    if constexpr (SWAP) {   // step 0's polarisation update (every later one rides behind the previous step's H half-step)
#pragma unroll
        for (int j = 0; j < C; ++j)
            pb[j] = A::add(A::add(A::mul(g.polA, pa[j]), A::mul(g.polB, pb[j])), A::mul(g.polC, ex[j]));
    }
#endif
    // The time loop exists twice: warps that own a source cell or a probe run the version with those
    // (warp-uniform) tests, every other warp a loop with nothing in it but the update itself.
    auto time_loop = [&](auto sp) {
        constexpr bool SP = decltype(sp)::value;
        int s = 0;
#ifdef PF_WARP_XCHG
#define PF_STEP_SYNC()
#else
#define PF_STEP_SYNC() cta_sync()
#endif
        for (; s + 1 < ks; s += 2) {
            tile_step<MODE, POL, C, A, GEN, SLAB, PML, SP, JX>(S, K, kc, W, tid, s, true, ex, hy, dx, pa, pb, pe, ph, acub, rbe, rce, rcm, jx);
            if (SP) probes(s);
            PF_STEP_SYNC();
            if (SWAP) tile_step<MODE, POL, C, A, GEN, SLAB, PML, SP, JX>(S, K, kc, W, tid, s + 1, s + 2 < ks, ex, hy, dx, pb, pa, pe, ph, acub, rbe, rce, rcm, jx);
            else tile_step<MODE, POL, C, A, GEN, SLAB, PML, SP, JX>(S, K, kc, W, tid, s + 1, s + 2 < ks, ex, hy, dx, pa, pb, pe, ph, acub, rbe, rce, rcm, jx);
            if (SP) probes(s + 1);
            PF_STEP_SYNC();
        }
        if (s < ks) {
            tile_step<MODE, POL, C, A, GEN, SLAB, PML, SP, JX>(S, K, kc, W, tid, s, false, ex, hy, dx, pa, pb, pe, ph, acub, rbe, rce, rcm, jx);
            if (SP) probes(s);
            PF_STEP_SYNC();
            swapped = SWAP;
        }
    };
    // (the cubic material law dwarfs those tests and is large: one copy of its loop only)
    if (CUB || K.wSrc || wProbe || wSnap) time_loop(std::true_type{});
    else time_loop(std::false_type{});

    // ---- store interior ---------------------------------------------------------------------------
    // Uniform-class warps re-derive where their cells go (tile base, grid, store mask) from the block / thread index and the
    // kernel's parameters instead of carrying those values through the time loop: the loop of the Lorentz slab body sits
    // exactly at the 64-register budget of two 512-thread CTAs per SM, and four spilled words in it cost 8 % of the sweep.
    const TileGrid *TGo = &TG;
    int lz0o = lz0;
    unsigned st = M.store;
    if constexpr (!GEN) {
        unsigned bid, tid2;
        asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(bid));
        asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid2));
        const TileDesc *tdp = tiles + bid;
        TGo = grids + tdp->grid;
        lz0o = tdp->base + (int)tid2 * C;
        st = 0;
#pragma unroll
        for (int j = 0; j < C; ++j) st |= (unsigned)((int)tid2 * C + j >= halo && (int)tid2 * C + j < TILE_CELLS - halo) << j;
    }
    const int dst = src ^ 1;
    double *__restrict__ outEx = TGo->buf[dst][S_EX];
    double *__restrict__ outHy = TGo->buf[dst][S_HY];
#pragma unroll
    for (int j = 0; j < C; ++j)
        if ((st >> j) & 1) { outEx[lz0o + j] = ex[j]; outHy[lz0o + j] = F32 ? hy[j] * uH : hy[j]; }
    if (HAS_PML) {
        double *__restrict__ outPe = TGo->buf[dst][S_PSIE];
        double *__restrict__ outPh = TGo->buf[dst][S_PSIH];
        const unsigned se = GEN ? (st & M.pmlE) : st, sh = GEN ? (st & M.pmlH) : st;
#pragma unroll
        for (int j = 0; j < C; ++j) {
            if ((se >> j) & 1) outPe[lz0o + j] = F32 ? pe[j] * uH : pe[j];
            if ((sh >> j) & 1) outPh[lz0o + j] = ph[j];
        }
    }
    if (HAS_MAT) {
        const unsigned sm = GEN ? (st & mSlab) : st;
        double *__restrict__ outDx = TGo->buf[dst][S_DX];
#pragma unroll
        for (int j = 0; j < C; ++j)
            if ((sm >> j) & 1) outDx[lz0o + j] = F32 ? (LOR ? ((double)dx[j] + (double)pa[j]) * uD : dx[j] * uD) : dx[j];
        if (LOR) {
            double *__restrict__ outP = TGo->buf[dst][S_P];
            double *__restrict__ outPp = TGo->buf[dst][S_PP];
#pragma unroll
            for (int j = 0; j < C; ++j)
                if ((sm >> j) & 1) {
                    if (F32) {
                        outP[lz0o + j] = pa[j] * uD;
                        outPp[lz0o + j] = ((double)pa[j] - (double)pb[j]) * uD;
                    } else {
                        outP[lz0o + j] = swapped ? pb[j] : pa[j];
                        outPp[lz0o + j] = swapped ? pa[j] : pb[j];
                    }
                }
        }
        if (CUB && TGo->d.g.Acubic) {
#pragma unroll
            for (int j = 0; j < C; ++j)
                if ((sm >> j) & 1) TGo->d.g.Acubic[lz0o + j] = acub[j];
        }
    }
}

#ifndef PF_TILE_MINBLOCKS_F32
#define PF_TILE_MINBLOCKS_F32 4   // measured on the Lorentz sweep: C=2 x 4 CTAs/SM 1023, C=2 x 3 978, C=4 x 3 869, C=4 x 2 762, C=1 x 2 778 Gcell-updates/s
#endif
#ifndef PF_TILE_MINBLOCKS_NEWTON
#define PF_TILE_MINBLOCKS_NEWTON PF_TILE_MINBLOCKS
#endif
#ifndef PF_TILE_MINBLOCKS_FREE
#define PF_TILE_MINBLOCKS_FREE 4   // 128 instead of 164 registers, 4 CTAs of 128 threads per SM: 1670 vs 1645 Gcell-updates/s (5e7 cells)
#endif
#ifndef PF_TILE_MINBLOCKS_CUBIC
#define PF_TILE_MINBLOCKS_CUBIC PF_TILE_MINBLOCKS   // 1 CTA/SM (128 registers, no spills): closed form 122 vs 148, Newton 186 vs 214 Gcell-updates/s
#endif
template <int MODE, class A>
constexpr int tile_minblocks()
{
    return std::is_same<typename A::real, float>::value ? PF_TILE_MINBLOCKS_F32
           : (A::newton ? PF_TILE_MINBLOCKS_NEWTON
              : (MODE == PF_FREE ? PF_TILE_MINBLOCKS_FREE : ((MODE == PF_NL || MODE == PF_LORENTZ_NL) ? PF_TILE_MINBLOCKS_CUBIC : PF_TILE_MINBLOCKS)));
}

template <int MODE, bool POL, int C, class A, bool JX = false>
__global__ void __launch_bounds__(TILE_CELLS / C, tile_minblocks<MODE, A>())
k_tile(const TileGrid *__restrict__ grids, const TileDesc *__restrict__ tiles, int src, int n_done,
       int n0, int ksteps, int halo)
{
    using R = typename A::real;
    constexpr unsigned RB = (unsigned)sizeof(R);
    constexpr int NT = TILE_CELLS / C;
    TileShared<R> S;
    unsigned sbase = (unsigned)__cvta_generic_to_shared(pf_smem);
    // keep the window address in an ordinary (per-thread) register: as a uniform value it is
    // re-derived from SR_CgaCtaId after every potentially divergent region of the time loop
    asm volatile("xor.b32 %0, %0, %1;" : "+r"(sbase) : "r"(threadIdx.x & 0u));
    S.be.base = sbase;
    S.ce.base = S.be.base + RB * TILE_CELLS;
    S.cm.base = S.ce.base + RB * TILE_CELLS;
    S.cEu.base = S.cm.base + RB * TILE_CELLS;
    S.cHu.base = S.cEu.base + RB * TILE_CELLS;
    S.cb.base = S.cHu.base + RB * TILE_CELLS;
    S.c2u.base = S.cb.base + RB * TILE_CELLS;
    // edgeH[-1] and edgeE[NT] are zero pads: the first / last thread reads its missing neighbour
    // without a per-step test
    S.edgeH.base = S.c2u.base + RB * TILE_CELLS + RB;
    S.edgeE.base = S.edgeH.base + RB * NT;
    S.srcE.base = S.edgeE.base + RB * NT + RB;
    S.srcH.base = S.srcE.base + RB * TILE_KMAX;
    S.xw = (S.srcH.base + RB * TILE_KMAX + 15u) & ~15u;

    const TileDesc *__restrict__ tdp = tiles + blockIdx.x;
    const int tbase = tdp->base;
    const unsigned tflags = tdp->flags;
    const TileGrid &TG = grids[tdp->grid];
    const int ks = min(ksteps, TG.nsteps - n_done);
    if (ks <= 0) return;

    const PfGrid &g = TG.d.g;
    const int tid = threadIdx.x;
    const int flags = g.flags;
    const int lz0 = tbase + tid * C;
#ifdef PF_WARP_XCHG
    const int L = g.L;
#endif

    // ---- warp class (precomputed per tile) and per-cell masks (bit j <-> cell lz0+j) ------
    const int cls_t = tdp->cls[tid >> 5];
    CellMasks M;
    if (cls_t == 4) tile_cell_masks<C>(g, tbase, lz0, halo, M);
    else tile_plain_masks<C>(tid, halo, cls_t, M);
    if (tflags & TILE_F_SPECIAL) tile_special_cells<C>(g, lz0, M);
    const int cls = tile_warp_class<C>(M);
    if (tid == 0) { S.edgeH[-1] = R(0); S.edgeE[NT] = R(0); }
    // source tables of this launch's steps (CTA-uniform load)
    const int nabs0 = n0 + n_done;
    for (int s = tid; s < ks; s += NT) {
        S.srcE[s] = (R)g.srcE[nabs0 + s];
        S.srcH[s] = (R)((flags & PF_F_TFSF) ? g.srcH[nabs0 + s] * (std::is_same<R, float>::value ? 1.0 / g.cH0 : 1.0) : 0.0);
    }

    WarpLink W;
    W.pH = W.pE = W.fH = W.fE = 0u;
    W.sendH = W.sendE = false;
    W.barL = W.barR = W.barS = 0u;
#ifdef PF_WARP_XCHG
    {
        // a neighbour warp is live iff at least one of its cells lies inside the grid (same test as cls == 5 above)
        const int w = tid >> 5, lane = tid & 31;
        const int wfirst = tbase + w * 32 * C;          // first cell of this warp
        const bool live = cls != 5;
        const bool left = live && w > 0 && wfirst - 1 >= 0 && wfirst - 32 * C < L;
        const bool right = live && w < NT / 32 - 1 && wfirst + 32 * C < L && wfirst + 64 * C - 1 >= 0;
        const unsigned dummy = xw_barH(S.xw, NT / 32);   // never arrived on
        W.sendH = right && lane == 31;
        W.sendE = left && lane == 0;
        W.barS = xw_barH(S.xw, w);
        W.barL = left ? xw_barH(S.xw, w - 1) : dummy;
        W.barR = right ? xw_barE(S.xw, w + 1) : dummy;
        W.fH = left ? 1u : 0u;  W.pH = left ? 0u : 1u;
        W.fE = right ? 1u : 0u; W.pE = right ? 0u : 1u;
        if (lane == 0) {
            xw_init(xw_barH(S.xw, w));
            xw_init(xw_barE(S.xw, w));
            if (w == 0) xw_init(dummy);
        }
        S.edgeH[tid] = R(0);   // slots of dead warps are read as 0 by their live neighbours
        S.edgeE[tid] = R(0);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncthreads();   // the only CTA-wide barrier of the launch: mbarriers initialised, source tables loaded
        if (cls == 5) return;
    }
#endif

    switch (cls) {
    case 0: tile_body<MODE, POL, C, A, false, false, false, JX>(TG, S, M, W, tid, lz0, ks, src, nabs0, grids, tiles, halo); break;
    case 1: tile_body<MODE, POL, C, A, false, true, false, JX>(TG, S, M, W, tid, lz0, ks, src, nabs0, grids, tiles, halo); break;
    case 2: tile_body<MODE, POL, C, A, false, false, true, JX>(TG, S, M, W, tid, lz0, ks, src, nabs0, grids, tiles, halo); break;
    case 3: tile_body<MODE, POL, C, A, false, true, true, JX>(TG, S, M, W, tid, lz0, ks, src, nabs0, grids, tiles, halo); break;
    case 5:
        // nothing to compute: publish zero edges once, then only keep the CTA's barrier count
        S.edgeH[tid] = R(0);
        S.edgeE[tid] = R(0);
        cta_sync();
        for (int s = 0; s < ks; ++s) { cta_sync(); cta_sync(); }
        break;
    default: tile_body<MODE, POL, C, A, true, true, true, JX>(TG, S, M, W, tid, lz0, ks, src, nabs0, grids, tiles, halo); break;
    }
}

// Writes the warp classes and flags of every tile (TileDesc::cls / flags): the same mask functions the kernels use, evaluated
// once per tile table instead of once per launch.  Launched with the thread geometry of the kernel that will read them.
template <int C>
__global__ void __launch_bounds__(TILE_CELLS / C) k_tile_classify(const TileGrid *__restrict__ grids, TileDesc *__restrict__ tiles, int halo)
{
    TileDesc &td = tiles[blockIdx.x];
    const PfGrid &g = grids[td.grid].d.g;
    const int tid = threadIdx.x;
    const int tbase = td.base;
    const int lz0 = tbase + tid * C;
    CellMasks M;
    tile_cell_masks<C>(g, tbase, lz0, halo, M);
    const bool special = tile_special_cells<C>(g, lz0, M);
    const int cls = tile_warp_class<C>(M);
    if ((tid & 31) == 0) td.cls[tid >> 5] = (unsigned char)cls;
    const int any = __syncthreads_or(special);
    if (tid == 0) { td.flags = any ? TILE_F_SPECIAL : 0u; td.cls_c = C; }
}

// Copy the interior of every tile from buffer `from` to buffer `to` -- only the cells a launch stores: Ex / Hy everywhere,
// psi_E / psi_H on their CPML update ranges, Dx / P / Pprev inside the slab (the same index rules as k_tile's store masks).
// Every other cell of the scratch buffer is never written and never read, so it needs no seeding, and the cells the
// caller's arrays hold there are left untouched.  only_odd = 1 restricts the copy to grids whose result ended in
// buffer 1 (an odd number of launches), for the final copy-back.
__global__ void __launch_bounds__(256) k_tile_copy(const TileGrid *__restrict__ grids,
                                                  const TileDesc *__restrict__ tiles, int halo,
                                                  int k_block, int mode, int from, int to, int only_odd)
{
    const TileDesc td = tiles[blockIdx.x];
    const TileGrid &TG = grids[td.grid];
    if (only_odd) {
        int launches = (TG.nsteps + k_block - 1) / k_block;
        if ((launches & 1) == 0) return;
    }
    const PfGrid &g = TG.d.g;
    const int W = TILE_CELLS - 2 * halo;
    const int narr = (mode == PF_LORENTZ || mode == PF_LORENTZ_NL) ? 7 : (mode == PF_NL ? 5 : 4);
    const int L = g.L, z0 = (int)g.z0, Lg = (int)g.Lg;
    const bool cpml_m = g.flags & PF_F_CPML_M, cpml_p = g.flags & PF_F_CPML_P;
    for (int a = 0; a < narr; ++a) {
        const double *__restrict__ s = TG.buf[from][a];
        double *__restrict__ d = TG.buf[to][a];
        if (!s || !d) continue;
        for (int i = threadIdx.x; i < W; i += blockDim.x) {
            const int lz = td.base + halo + i;
            if (lz < 0 || lz >= L) continue;
            const int gz = z0 + lz;
            bool ok = true;
            if (a == S_PSIE || a == S_PSIH) {
                const bool in_pml = (cpml_m && gz < g.pw) || (cpml_p && gz >= Lg - g.pw);
                const bool upd = (a == S_PSIE) ? (lz >= 1 && gz >= 1 && gz <= Lg - 1) : (lz <= L - 2 && gz >= 1 && gz <= Lg - 2);
                ok = in_pml && upd;
            } else if (a >= S_DX) {
                ok = lz >= 1 && gz >= g.mf && gz < g.mr;
            }
            if (ok) d[lz] = s[lz];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int n_state_arrays(int mode) { return (mode == PF_LORENTZ || mode == PF_LORENTZ_NL) ? 7 : (mode == PF_NL ? 5 : 4); }

// does the piece [z0, z0+L) of the global grid contain CPML / slab cells?
static inline bool piece_has_pml(const PfGrid &g)
{
    long long a = g.z0, b = g.z0 + g.L;   // [a, b)
    bool left = (g.flags & PF_F_CPML_M) && a < g.pw;
    bool right = (g.flags & PF_F_CPML_P) && b > g.Lg - g.pw;
    return g.pw > 0 && (left || right);
}
static inline bool piece_has_slab(const PfGrid &g) { return g.z0 < g.mr && g.z0 + g.L > g.mf; }

// n0 .. n0+nsteps-1 must lie inside the caller's source tables and probe rows (PfGrid.n_src / probe_stride)
int check_step_range(const PfGrid &g, int n0, int nsteps, const char *who)
{
    if (n0 < 0 || nsteps < 0) return set_err(PF_E_ARG, "%s: negative step range", who);
    if (nsteps == 0) return 0;
    const long long end = (long long)n0 + nsteps;
    if (g.n_src > 0 && end > g.n_src)
        return set_err(PF_E_ARG, "%s: steps %d..%lld run past the source tables (n_src = %d)", who, n0, end - 1, g.n_src);
    if (g.n_probes > 0 && end > g.probe_stride)
        return set_err(PF_E_ARG, "%s: steps %d..%lld run past the probe rows (probe_stride = %d)", who, n0, end - 1, g.probe_stride);
    return 0;
}

static int tile_supported(const PfGrid &g, int mode)
{
    if (g.Lg >= (1LL << 31) - (1LL << 12) || g.z0 < -(1LL << 30) || g.z0 + g.L > g.Lg + (1LL << 12))
        return set_err(PF_E_UNSUPPORTED, "tile engine: global indices are evaluated in 32 bits (Lg = %lld)", (long long)g.Lg);
    if (!(g.flags & PF_F_CANONICAL)) return set_err(PF_E_UNSUPPORTED, "tile engine needs PF_F_CANONICAL coefficients");
    if (mode != PF_FREE && (g.flags & PF_F_TFSF) && g.nzsrc - 1 >= g.mf - 1 && g.nzsrc - 1 < g.mr)
        return set_err(PF_E_UNSUPPORTED, "tile engine: TF/SF point inside the slab");
    if (g.n_probes > 64) return set_err(PF_E_UNSUPPORTED, "tile engine: more than 64 probes");
    if (!g.Ex || !g.Hy) return set_err(PF_E_ARG, "Ex/Hy missing");
    // arrays that only exist on CPML / slab cells may be NULL for a piece of a decomposed grid that
    // holds no such cell (the kernel never dereferences them there)
    if (piece_has_pml(g) && (!g.psiE || !g.psiH || !g.beX || !g.ceX || !g.cmY)) return set_err(PF_E_ARG, "CPML arrays missing");
    if (mode != PF_FREE && piece_has_slab(g)) {
        if (!g.Dx) return set_err(PF_E_ARG, "Dx missing");
        if ((mode == PF_LORENTZ || mode == PF_LORENTZ_NL) && (!g.P || !g.Pprev)) return set_err(PF_E_ARG, "P/Pprev missing");
    }
    return 0;
}

struct TilePlan {
    size_t off_grids, off_tiles, off_state, total;
    int n_tiles;
};

static TilePlan tile_plan(const PfGrid *grids, int n, int mode, int halo)
{
    TilePlan p;
    int W = TILE_CELLS - 2 * halo;
    long long nt = 0;
    size_t state = 0;
    for (int m = 0; m < n; ++m) {
        nt += (grids[m].L + W - 1) / W;
        const PfGrid &g = grids[m];
        const double *prim[7] = {g.Ex, g.Hy, g.psiE, g.psiH, g.Dx, g.P, g.Pprev};
        for (int a = 0; a < n_state_arrays(mode); ++a)
            if (prim[a]) state += align_up(sizeof(double) * g.L, 256);
    }
    p.n_tiles = (int)nt;
    p.off_grids = 0;
    p.off_tiles = align_up(sizeof(TileGrid) * (size_t)n, 256);
    p.off_state = p.off_tiles + align_up(sizeof(TileDesc) * (size_t)nt, 256);
    p.total = p.off_state + state;
    return p;
}

enum { ARITH_EXACT = 0, ARITH_FUSED = 1, ARITH_FP32 = 2, ARITH_NEWTON = 3, ARITH_EXACT_JX = 4 };

template <int MODE, bool POL, int C, class A>
static const char *tile_kernel_name()
{
    static char name[64] = "";
    if (!name[0]) {
        const char *mode = MODE == PF_FREE ? "FREE" : MODE == PF_LORENTZ ? "LORENTZ" : MODE == PF_NL ? "NL" : "LORENTZ_NL";
        const char *ar = std::is_same<A, Exact>::value ? "Exact" : std::is_same<A, Fused>::value ? "Fused"
                         : std::is_same<A, Fast32>::value ? "Fast32" : "ExactNewton";
        snprintf(name, sizeof(name), "k_tile<%s,POL=%d,C=%d,%s>", mode, (int)POL, C, ar);
    }
    return name;
}

// cls_c: cells per thread the tile table's warp classes were computed for (k_tile_classify) -- must be this kernel's C
template <int MODE, bool POL, int C, class A, bool JX = false>
static int launch_tile_a(int n_tiles, const TileGrid *dg, const TileDesc *dt, int src, int n_done,
                         int n0, int ks, int halo, cudaStream_t st, int cls_c)
{
    if (cls_c != C) return set_err(PF_E_ARG, "tile engine: the tile table was classified for %d cells per thread, the kernel has %d", cls_c, C);
    const size_t sm = TileSmem<MODE, C, typename A::real>::bytes;
    ProfScope prof(st, tile_kernel_name<MODE, POL, C, A>());
    static std::atomic<unsigned long long> attr_set{0};   // the attribute is per device; bit = device ordinal (idempotent, race-free)
    int dev = 0;
    PF_CUDA(cudaGetDevice(&dev));
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(attr_set.load(std::memory_order_acquire) & bit)) {
        PF_CUDA(cudaFuncSetAttribute(k_tile<MODE, POL, C, A, JX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        attr_set.fetch_or(bit, std::memory_order_release);
    }
    k_tile<MODE, POL, C, A, JX><<<n_tiles, TILE_CELLS / C, sm, st>>>(dg, dt, src, n_done, n0, ks, halo);
    PF_LAUNCH_CHECK("k_tile");
    return 0;
}

template <int MODE, bool POL, int C>
static int launch_tile(int arith, int n_tiles, const TileGrid *dg, const TileDesc *dt, int src, int n_done,
                       int n0, int ks, int halo, cudaStream_t st, int cls_c)
{
    if (arith == ARITH_FUSED) return launch_tile_a<MODE, POL, C, Fused>(n_tiles, dg, dt, src, n_done, n0, ks, halo, st, cls_c);
    if constexpr (MODE == PF_NL || MODE == PF_LORENTZ_NL) {
        if (arith == ARITH_NEWTON) return launch_tile_a<MODE, POL, C, ExactNewton>(n_tiles, dg, dt, src, n_done, n0, ks, halo, st, cls_c);
    }
    return launch_tile_a<MODE, POL, C, Exact>(n_tiles, dg, dt, src, n_done, n0, ks, halo, st, cls_c);
}

constexpr int TILE_C = PF_TILE_C;            // cells per thread, Lorentz and nonlinear modes
#ifndef PF_TILE_C_FREE
#define PF_TILE_C_FREE 8
#endif
// The vacuum / dielectric mode carries little state per cell, so more cells per thread (more ILP, fewer
// edge exchanges per cell) win: measured on a 1.7e7-cell grid 1038 (C=2), 1105 (C=4), 1353 (C=8)
// Gcell-updates/s, while the Lorentz sweep prefers C=2 (491 vs 399 vs 219) and the cubic path too.
constexpr int TILE_C_FREE = PF_TILE_C_FREE;

// wide = true: grids with (almost) no CPML cells, e.g. the pieces of a long grid -- 4 cells per thread
// (measured 733 vs 684 Gcell-updates/s on a 1e8-cell Lorentz grid; the CPML-heavy sweep members prefer 2).
#ifndef PF_TILE_C_F32
#define PF_TILE_C_F32 2
#endif
#ifndef PF_TILE_C_F32_FREE
#define PF_TILE_C_F32_FREE 8
#endif
// cells per thread of the kernel launch_tile_mode picks (the tile table's warp classes depend on it)
static int tile_c_for(int mode, int fma, bool wide)
{
    if (mode == PF_FREE) return fma == ARITH_FP32 ? PF_TILE_C_F32_FREE : TILE_C_FREE;
    if (fma == ARITH_FP32) return PF_TILE_C_F32;
    if (fma == ARITH_EXACT_JX) return TILE_C;
    return (mode == PF_LORENTZ && wide) ? 2 * TILE_C : TILE_C;
}

static int tile_classify(int c, int n_tiles, const TileGrid *dg, TileDesc *dt, int halo, cudaStream_t st)
{
    ProfScope prof(st, "k_tile_classify");
    switch (c) {
    case 1: k_tile_classify<1><<<n_tiles, TILE_CELLS / 1, 0, st>>>(dg, dt, halo); break;
    case 2: k_tile_classify<2><<<n_tiles, TILE_CELLS / 2, 0, st>>>(dg, dt, halo); break;
    case 4: k_tile_classify<4><<<n_tiles, TILE_CELLS / 4, 0, st>>>(dg, dt, halo); break;
    case 8: k_tile_classify<8><<<n_tiles, TILE_CELLS / 8, 0, st>>>(dg, dt, halo); break;
    default: return set_err(PF_E_ARG, "tile engine: %d cells per thread", c);
    }
    PF_LAUNCH_CHECK("k_tile_classify");
    return 0;
}

static int launch_tile_mode(int mode, int do_pol, int fma, bool wide, int n_tiles, const TileGrid *dg, const TileDesc *dt,
                            int src, int n_done, int n0, int ks, int halo, cudaStream_t st, int cls_c)
{
#ifdef PF_EXPERIMENT   // compile-time experiments on one kernel (tools/sass_loops.py): instantiate the headline kernel only
    return launch_tile_a<PF_LORENTZ, true, TILE_C, Exact>(n_tiles, dg, dt, src, n_done, n0, ks, halo, st, cls_c);
#else
    if (fma == ARITH_EXACT_JX) {   // a current slot (PIC coupling) is present: exact arithmetic, one geometry per mode
        if (mode == PF_FREE) return launch_tile_a<PF_FREE, false, TILE_C_FREE, Exact, true>(n_tiles, dg, dt, src, n_done, n0, ks, halo, st, cls_c);
        if (mode == PF_LORENTZ)
            return do_pol ? launch_tile_a<PF_LORENTZ, true, TILE_C, Exact, true>(n_tiles, dg, dt, src, n_done, n0, ks, halo, st, cls_c)
                          : launch_tile_a<PF_LORENTZ, false, TILE_C, Exact, true>(n_tiles, dg, dt, src, n_done, n0, ks, halo, st, cls_c);
        if (mode == PF_NL) return launch_tile_a<PF_NL, false, TILE_C, Exact, true>(n_tiles, dg, dt, src, n_done, n0, ks, halo, st, cls_c);
        return set_err(PF_E_UNSUPPORTED, "tile engine: a current slot (Jx) is supported in modes FREE, LORENTZ and NL");
    }
    if (fma == ARITH_FP32) {   // PF_F_FP32: one geometry per mode
        if (mode == PF_FREE) return launch_tile_a<PF_FREE, false, PF_TILE_C_F32_FREE, Fast32>(n_tiles, dg, dt, src, n_done, n0, ks, halo, st, cls_c);
        if (mode == PF_LORENTZ)
            return do_pol ? launch_tile_a<PF_LORENTZ, true, PF_TILE_C_F32, Fast32>(n_tiles, dg, dt, src, n_done, n0, ks, halo, st, cls_c)
                          : launch_tile_a<PF_LORENTZ, false, PF_TILE_C_F32, Fast32>(n_tiles, dg, dt, src, n_done, n0, ks, halo, st, cls_c);
        if (mode == PF_NL) return launch_tile_a<PF_NL, false, PF_TILE_C_F32, Fast32>(n_tiles, dg, dt, src, n_done, n0, ks, halo, st, cls_c);
        if (mode == PF_LORENTZ_NL)
            return do_pol ? launch_tile_a<PF_LORENTZ_NL, true, PF_TILE_C_F32, Fast32>(n_tiles, dg, dt, src, n_done, n0, ks, halo, st, cls_c)
                          : launch_tile_a<PF_LORENTZ_NL, false, PF_TILE_C_F32, Fast32>(n_tiles, dg, dt, src, n_done, n0, ks, halo, st, cls_c);
        return set_err(PF_E_ARG, "bad mode %d", mode);
    }
    if (mode == PF_FREE) return launch_tile<PF_FREE, false, TILE_C_FREE>(fma, n_tiles, dg, dt, src, n_done, n0, ks, halo, st, cls_c);
    if (mode == PF_LORENTZ) {
        if (wide)
            return do_pol ? launch_tile<PF_LORENTZ, true, 2 * TILE_C>(fma, n_tiles, dg, dt, src, n_done, n0, ks, halo, st, cls_c)
                          : launch_tile<PF_LORENTZ, false, 2 * TILE_C>(fma, n_tiles, dg, dt, src, n_done, n0, ks, halo, st, cls_c);
        return do_pol ? launch_tile<PF_LORENTZ, true, TILE_C>(fma, n_tiles, dg, dt, src, n_done, n0, ks, halo, st, cls_c)
                      : launch_tile<PF_LORENTZ, false, TILE_C>(fma, n_tiles, dg, dt, src, n_done, n0, ks, halo, st, cls_c);
    }
    if (mode == PF_NL) return launch_tile<PF_NL, false, TILE_C>(fma, n_tiles, dg, dt, src, n_done, n0, ks, halo, st, cls_c);
    if (mode == PF_LORENTZ_NL)
        return do_pol ? launch_tile<PF_LORENTZ_NL, true, TILE_C>(fma, n_tiles, dg, dt, src, n_done, n0, ks, halo, st, cls_c)
                      : launch_tile<PF_LORENTZ_NL, false, TILE_C>(fma, n_tiles, dg, dt, src, n_done, n0, ks, halo, st, cls_c);
    return set_err(PF_E_ARG, "bad mode %d", mode);
#endif
}

// arithmetic class of a launch: PF_F_FP32 on any grid selects single precision, else PF_F_FMA contraction
static int arith_of(const PfGrid *grids, int n, int mode, int *arith)
{
    int a = ARITH_EXACT;
    const bool cub = mode == PF_NL || mode == PF_LORENTZ_NL;
    bool newton = cub;   // the inlined Newton law needs every grid of the launch to ask for it and to be admissible
    for (int m = 0; m < n; ++m) {
        const PfGrid &g = grids[m];
        // (PF_LORENTZ_NL is specified with the converged root: Newton whenever the coefficients admit it)
        newton = newton && ((g.flags & PF_F_NEWTON) || mode == PF_LORENTZ_NL) && g.cub_a >= 0.0 && g.cub_b >= 0.0 && g.cub_c > 0.0;
        if (g.flags & PF_F_FP32) {
            // the fp32 cubic root is a Newton iteration that needs an increasing, convex polynomial
            if ((mode == PF_NL || mode == PF_LORENTZ_NL) && !(g.cub_a >= 0.0 && g.cub_b >= 0.0 && g.cub_c > 0.0))
                return set_err(PF_E_UNSUPPORTED, "PF_F_FP32: nonlinear mode needs cub >= 0, qua >= 0, one > 0");
            a = ARITH_FP32;
        } else if ((g.flags & PF_F_FMA) && a == ARITH_EXACT) {
            a = ARITH_FUSED;
        }
    }
    bool jx = false;
    for (int m = 0; m < n; ++m) jx = jx || grids[m].Jx != nullptr;
    if (jx) {
        if (a != ARITH_EXACT) return set_err(PF_E_UNSUPPORTED, "tile engine: a current slot (Jx) needs exact arithmetic (no PF_F_FMA / PF_F_FP32)");
        *arith = ARITH_EXACT_JX;      // (PF_F_NEWTON is ignored: the closed-form law runs)
        return 0;
    }
    if (a == ARITH_EXACT && newton) a = ARITH_NEWTON;
    *arith = a;
    return 0;
}

// fraction of CPML cells over a set of grids < 1/8 ?
static bool mostly_interior(const PfGrid *grids, int n)
{
    long long cells = 0, pml = 0;
    for (int m = 0; m < n; ++m) {
        const PfGrid &g = grids[m];
        cells += g.L;
        long long a = g.z0, b = g.z0 + g.L;
        if (g.flags & PF_F_CPML_M) pml += std::max(0LL, std::min<long long>(b, g.pw) - a);
        if (g.flags & PF_F_CPML_P) pml += std::max(0LL, b - std::max<long long>(a, g.Lg - g.pw));
    }
    return pml * 8 < cells;
}

// Runs the tile engine over n grids.  snap_* only with n == 1.
int tile_run(const PfGrid *grids, int n, int mode, int do_pol, int n0, const int *nsteps, int k_block,
             double *snap_out, int snap_interval, int snap_rows, void *scratch, size_t scratch_bytes,
             cudaStream_t st)
{
    if (n <= 0) return PF_OK;
    if (k_block <= 0) {
        // default: TILE_KDEF; but a batch whose tiles all fit the machine at once (a single run: 12-15 tiles) is bound by step
        // latency and by the ~9 us a launch costs whatever it holds, not by the halo's redundant cells: longest blocks then
        k_block = TILE_KDEF;
        int dev = 0, sms = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess) {
            long long nt = 0;
            for (int m = 0; m < n; ++m) nt += (grids[m].L + (TILE_CELLS - 2 * TILE_KMAX) - 1) / (TILE_CELLS - 2 * TILE_KMAX);
            if (nt <= sms) k_block = TILE_KMAX;
        }
    }
    if (k_block > TILE_KMAX) k_block = TILE_KMAX;
    for (int m = 0; m < n; ++m) {
        int rc = tile_supported(grids[m], mode);
        if (rc) return rc;
        rc = check_step_range(grids[m], n0, nsteps[m], "pf_run_batch");
        if (rc) return rc;
    }
    const int halo = k_block;
    const int W = TILE_CELLS - 2 * halo;
    TilePlan plan = tile_plan(grids, n, mode, halo);
    if (!scratch || scratch_bytes < plan.total)
        return set_err(PF_E_SCRATCH, "tile engine needs %zu bytes of scratch, got %zu", plan.total, scratch_bytes);

    // host images of the device tables
    std::vector<TileGrid> hg(n);
    std::vector<TileDesc> ht;
    ht.reserve(plan.n_tiles);
    char *sbase = (char *)scratch;
    size_t off = plan.off_state;
    int max_steps = 0;
    int fma = ARITH_EXACT;
    {
        int rc = arith_of(grids, n, mode, &fma);
        if (rc) return rc;
    }
    const int na = n_state_arrays(mode);
    for (int m = 0; m < n; ++m) {
        const PfGrid &g = grids[m];
        TileGrid &t = hg[m];
        t.d = make_grid_dev(g);
        double *prim[7] = {g.Ex, g.Hy, g.psiE, g.psiH, g.Dx, g.P, g.Pprev};
        for (int a = 0; a < 7; ++a) {
            t.buf[0][a] = prim[a];
            t.buf[1][a] = nullptr;
            if (a < na && prim[a]) {
                t.buf[1][a] = (double *)(sbase + off);
                off += align_up(sizeof(double) * g.L, 256);
            }
        }
        t.nsteps = nsteps[m];
        t.snap_interval = 0;
        t.snap_out = nullptr;
        t.snap_rows = 0;
        t.pad = 0;
        max_steps = std::max(max_steps, nsteps[m]);
        int ntile = (g.L + W - 1) / W;
        for (int i = 0; i < ntile; ++i) ht.push_back(TileDesc{m, i * W - halo});
    }
    // snapshot rows (n == 1 only).  Linear modes: written by the kernel at the step they belong to, so the run is cut into the
    // same k-step launches as one without; cubic modes: a launch ends at every snapshot step and the row is copied out.
    const bool snaps_any = snap_out && snap_interval > 0 && n == 1;
    const bool snaps_in_kernel = snaps_any && (mode == PF_FREE || mode == PF_LORENTZ);
    const bool snaps = snaps_any && !snaps_in_kernel;
    if (snaps_in_kernel) {
        hg[0].snap_out = snap_out;
        hg[0].snap_interval = snap_interval;
        hg[0].snap_rows = snap_rows;
    }
    TileGrid *dg = (TileGrid *)(sbase + plan.off_grids);
    TileDesc *dt = (TileDesc *)(sbase + plan.off_tiles);
    PF_CUDA(cudaMemcpyAsync(dg, hg.data(), sizeof(TileGrid) * n, cudaMemcpyHostToDevice, st));
    PF_CUDA(cudaMemcpyAsync(dt, ht.data(), sizeof(TileDesc) * ht.size(), cudaMemcpyHostToDevice, st));
    // A launch stores only the cells a state array is defined on (psi inside the CPML, Dx/P inside the slab) and
    // reads only those; the copy-back below moves exactly those cells, so the scratch buffer needs no seeding.

    const bool wide = mostly_interior(grids, n);
    const int cls_c = tile_c_for(mode, fma, wide);
    {
        int rc = tile_classify(cls_c, (int)ht.size(), dg, dt, halo, st);
        if (rc) return rc;
    }
    int n_done = 0, src = 0;
    while (n_done < max_steps) {
        int ks = std::min(k_block, max_steps - n_done);
        if (snaps) {
            // end this launch right after the next snapshot step (n>0, n % interval == 0)
            int nabs = n0 + n_done;
            int next_snap = ((nabs / snap_interval) + 1) * snap_interval;  // first multiple > nabs
            if (nabs % snap_interval == 0 && nabs > 0) next_snap = nabs;   // step nabs itself is a snapshot step
            ks = std::min(ks, next_snap - nabs + 1);
        }
        int rc = launch_tile_mode(mode, do_pol, fma, wide, (int)ht.size(), dg, dt, src, n_done, n0, ks, halo, st, cls_c);
        if (rc) return rc;
        n_done += ks;
        src ^= 1;
        if (snaps) {
            int nlast = n0 + n_done - 1;
            if (nlast > 0 && nlast % snap_interval == 0) {
                int row = nlast / snap_interval;
                if (row < snap_rows)
                    PF_CUDA(cudaMemcpyAsync(snap_out + (size_t)row * grids[0].L, hg[0].buf[src][S_EX],
                                            sizeof(double) * grids[0].L, cudaMemcpyDeviceToDevice, st));
            }
        }
    }
    if (snaps) {
        // variable-length launches: parity is not a function of nsteps/k_block; copy back explicitly
        if (src == 1) {
            k_tile_copy<<<(int)ht.size(), 256, 0, st>>>(dg, dt, halo, k_block, mode, 1, 0, 0);
            PF_LAUNCH_CHECK("k_tile_copy");
        }
    } else {
        bool any_odd = false;      // grids whose result ended in the scratch buffer
        for (int m = 0; m < n; ++m) any_odd = any_odd || ((((nsteps[m] + k_block - 1) / k_block) & 1) != 0);
        if (any_odd) {
            k_tile_copy<<<(int)ht.size(), 256, 0, st>>>(dg, dt, halo, k_block, mode, 1, 0, 1);
            PF_LAUNCH_CHECK("k_tile_copy");
        }
    }
    // host vectors go out of scope: the async H2D copies above were issued from pageable memory,
    // which cudaMemcpyAsync stages before returning.
    return PF_OK;
}

// A tile of a piece of a decomposed grid is an EDGE tile if it reads ghost cells: cells [0, halo) of a piece that has a left
// neighbour (z0 > 0), cells [L - halo, L) of one that has a right neighbour (z0 + L < Lg).  pf_run_block keeps the edge tiles
// in front of the table so that a caller can advance the inner tiles while the ghost exchange is still in flight.
static inline bool tile_is_edge(const PfGrid &g, int base, int halo)
{
    const bool left = g.z0 > 0 && base < halo;
    const bool right = g.z0 + g.L < g.Lg && (long long)base + TILE_CELLS > (long long)g.L - halo;
    return left || right;
}
// number of edge tiles of one piece (only its first and last few tiles can be: halo <= TILE_KMAX, W >= 3/4 TILE_CELLS)
static int tile_count_edge(const PfGrid &g, int halo)
{
    const int W = TILE_CELLS - 2 * halo;
    const int ntile = (g.L + W - 1) / W;
    int c = 0;
    for (int i = 0; i < ntile; ++i) {
        if (i == 4 && ntile > 8) i = ntile - 4;
        c += tile_is_edge(g, i * W - halo, halo) ? 1 : 0;
    }
    return c;
}

// One launch: advance n grids by `ks` steps (absolute steps n0 .. n0+ks-1) reading the arrays of
// src[m] and writing the arrays of dst[m] (the caller owns both buffers and alternates them).
// halo >= ks is the overlap the tiles are cut with.  scratch holds only the tile tables: they carry both
// buffer sets (the kernel's `src` argument selects the direction), so the steady state of a ping-pong run
// -- the caller promises PF_BLOCK_F_TABLES_VALID -- costs no host-side rebuild and no H2D copy.  Nothing about
// earlier calls is remembered here.
int tile_block(const PfGrid *src, const PfGrid *dst, int n, int mode, int do_pol, int n0, int ks, int halo, int block_flags,
               void *scratch, size_t scratch_bytes, cudaStream_t st)
{
    if (n <= 0 || ks <= 0) return PF_OK;
    if (halo < ks) halo = ks;
    if (halo > TILE_KMAX) return set_err(PF_E_ARG, "pf_run_block: ksteps/halo %d exceeds the tile engine limit %d", halo, TILE_KMAX);
    const size_t off_tiles = align_up(sizeof(TileGrid) * (size_t)n, 256);
    TileGrid *dg = (TileGrid *)scratch;
    TileDesc *dt = (TileDesc *)((char *)scratch + off_tiles);
    int fma = ARITH_EXACT;
    {
        int rc = arith_of(src, n, mode, &fma);
        if (rc) return rc;
    }
    const int W = TILE_CELLS - 2 * halo;
    long long n_tiles = 0;
    for (int m = 0; m < n; ++m) {
        int rc = tile_supported(src[m], mode);
        if (rc) return rc;
        if (dst[m].L != src[m].L) return set_err(PF_E_ARG, "pf_run_block: src/dst length mismatch in grid %d", m);
        rc = check_step_range(src[m], n0, ks, "pf_run_block");
        if (rc) return rc;
        n_tiles += (src[m].L + W - 1) / W;
    }
    const size_t need = off_tiles + align_up(sizeof(TileDesc) * (size_t)n_tiles, 256);
    if (!scratch || scratch_bytes < need) return set_err(PF_E_SCRATCH, "pf_run_block needs %zu bytes of scratch, got %zu", need, scratch_bytes);
    const bool wide = mostly_interior(src, n);
    const int cls_c = tile_c_for(mode, fma, wide);
    // the part of the table this call launches: [edge tiles | inner tiles]
    if ((block_flags & PF_BLOCK_F_EDGE_TILES) && (block_flags & PF_BLOCK_F_INNER_TILES))
        return set_err(PF_E_ARG, "pf_run_block: PF_BLOCK_F_EDGE_TILES and PF_BLOCK_F_INNER_TILES exclude each other (neither = all tiles)");
    long long n_edge = 0;
    for (int m = 0; m < n; ++m) n_edge += tile_count_edge(src[m], halo);
    const long long part_off = (block_flags & PF_BLOCK_F_INNER_TILES) ? n_edge : 0;
    const long long part_n = (block_flags & PF_BLOCK_F_EDGE_TILES) ? n_edge : ((block_flags & PF_BLOCK_F_INNER_TILES) ? n_tiles - n_edge : n_tiles);
    if (block_flags & PF_BLOCK_F_TABLES_VALID) {
        if (part_n == 0) return PF_OK;
        return launch_tile_mode(mode, do_pol, fma, wide, (int)part_n, dg, dt + part_off, (block_flags & PF_BLOCK_F_SWAPPED) ? 1 : 0, 0, n0, ks, halo, st, cls_c);
    }
    std::vector<TileGrid> hg(n);
    std::vector<TileDesc> ht, inner;
    ht.reserve((size_t)n_tiles);
    inner.reserve((size_t)n_tiles);
    for (int m = 0; m < n; ++m) {
        TileGrid &t = hg[m];
        t.d = make_grid_dev(src[m]);
        double *a0[7] = {src[m].Ex, src[m].Hy, src[m].psiE, src[m].psiH, src[m].Dx, src[m].P, src[m].Pprev};
        double *a1[7] = {dst[m].Ex, dst[m].Hy, dst[m].psiE, dst[m].psiH, dst[m].Dx, dst[m].P, dst[m].Pprev};
        for (int a = 0; a < 7; ++a) {
            if ((a0[a] == nullptr) != (a1[a] == nullptr)) return set_err(PF_E_ARG, "pf_run_block: array %d present in only one buffer", a);
            t.buf[0][a] = a0[a];
            t.buf[1][a] = a1[a];
        }
        t.nsteps = 1 << 30;   // the step count of a block launch is the kernel's ksteps argument
        t.snap_interval = 0;
        t.snap_out = nullptr;
        t.snap_rows = 0;
        t.pad = 0;
        int ntile = (src[m].L + W - 1) / W;
        for (int i = 0; i < ntile; ++i) (tile_is_edge(src[m], i * W - halo, halo) ? ht : inner).push_back(TileDesc{m, i * W - halo});
    }
    if ((long long)ht.size() != n_edge) return set_err(PF_E_ARG, "pf_run_block: internal: edge tile count %zu != %lld", ht.size(), n_edge);
    ht.insert(ht.end(), inner.begin(), inner.end());
    PF_CUDA(cudaMemcpyAsync(dg, hg.data(), sizeof(TileGrid) * n, cudaMemcpyHostToDevice, st));
    PF_CUDA(cudaMemcpyAsync(dt, ht.data(), sizeof(TileDesc) * ht.size(), cudaMemcpyHostToDevice, st));
    {
        int rc = tile_classify(cls_c, (int)ht.size(), dg, dt, halo, st);
        if (rc) return rc;
    }
    if (part_n == 0) return PF_OK;
    return launch_tile_mode(mode, do_pol, fma, wide, (int)part_n, dg, dt + part_off, 0, 0, n0, ks, halo, st, cls_c);
}

int ops_run_pass(const PfGrid *g, int mode, int do_pol, int n0, int nsteps, double *snap_out,
                 int snap_interval, int snap_rows, cudaStream_t st);

}  // namespace pf

using namespace pf;

extern "C" {

size_t pf_run_scratch_bytes(const PfGrid *grids, int n_grids, int engine)
{
    if (engine != PF_ENGINE_TILE || !grids || n_grids <= 0) return 0;
    // sized for the largest state set (Lorentz) and the smallest tile interior (largest tile count)
    return tile_plan(grids, n_grids, PF_LORENTZ, TILE_KMAX).total;
}

int pf_run_block(const PfGrid *src, const PfGrid *dst, int n_grids, int mode, int do_pol, int n0, int ksteps, int halo,
                 int block_flags, void *scratch, size_t scratch_bytes, void *stream)
{
    if (!src || !dst || n_grids < 0) return set_err(PF_E_ARG, "pf_run_block: bad arguments");
    if (mode < PF_FREE || mode > PF_LORENTZ_NL) return set_err(PF_E_ARG, "pf_run_block: bad mode %d", mode);
    return tile_block(src, dst, n_grids, mode, do_pol, n0, ksteps, halo, block_flags, scratch, scratch_bytes, (cudaStream_t)stream);
}

size_t pf_run_block_scratch_bytes(const PfGrid *grids, int n_grids, int halo)
{
    if (!grids || n_grids <= 0) return 0;
    if (halo <= 0 || halo > TILE_KMAX) halo = TILE_KMAX;
    const int W = TILE_CELLS - 2 * halo;
    size_t nt = 0;
    for (int m = 0; m < n_grids; ++m) nt += (grids[m].L + W - 1) / W;
    return align_up(sizeof(TileGrid) * (size_t)n_grids, 256) + align_up(sizeof(TileDesc) * nt, 256);
}

int pf_tile_config(int *tile_cells, int *k_max, int *threads)
{
    if (tile_cells) *tile_cells = TILE_CELLS;
    if (k_max) *k_max = TILE_KMAX;
    if (threads) *threads = TILE_CELLS / TILE_C;
    return PF_OK;
}

int pf_run_pass(const PfGrid *g, int mode, int do_pol, int n0, int nsteps, int engine, double *snap_out,
                int snap_interval, int snap_rows, void *scratch, size_t scratch_bytes, void *stream)
{
    if (!g || nsteps < 0) return set_err(PF_E_ARG, "pf_run_pass: bad arguments");
    if (mode < PF_FREE || mode > PF_LORENTZ_NL) return set_err(PF_E_ARG, "pf_run_pass: bad mode %d", mode);
    cudaStream_t st = (cudaStream_t)stream;
    if (engine == PF_ENGINE_OPS && (g->flags & PF_F_FP32))
        return set_err(PF_E_UNSUPPORTED, "PF_F_FP32 is a mode of the tile engine only");
    {
        int rc = check_step_range(*g, n0, nsteps, "pf_run_pass");
        if (rc) return rc;
    }
    if (engine == PF_ENGINE_OPS) return ops_run_pass(g, mode, do_pol, n0, nsteps, snap_out, snap_interval, snap_rows, st);
    if (engine == PF_ENGINE_TILE)
        return tile_run(g, 1, mode, do_pol, n0, &nsteps, 0, snap_out, snap_interval, snap_rows, scratch, scratch_bytes, st);
    return set_err(PF_E_ARG, "pf_run_pass: bad engine %d", engine);
}

int pf_run_batch(const PfGrid *grids, int n_grids, int mode, int do_pol, int n0, const int *nsteps, int k_block,
                 void *scratch, size_t scratch_bytes, void *stream)
{
    if (!grids || n_grids < 0 || !nsteps) return set_err(PF_E_ARG, "pf_run_batch: bad arguments");
    if (mode < PF_FREE || mode > PF_LORENTZ_NL) return set_err(PF_E_ARG, "pf_run_batch: bad mode %d", mode);
    return tile_run(grids, n_grids, mode, do_pol, n0, nsteps, k_block, nullptr, 0, 0, scratch, scratch_bytes,
                    (cudaStream_t)stream);
}

}  // extern "C"
