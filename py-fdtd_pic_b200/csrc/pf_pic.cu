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
struct Pushed {
    double z, ux, uz;
    int cell;
};

// Loop-invariant scalars of the push, evaluated once on the host with the same IEEE expressions the
// oracle uses (1/dz, (L-1)*dz, q/m*dt*0.5, 1/(c*c)): saves two fp64 divisions per particle.
struct PicDerived {
    double inv_dz, zmax, qmdt2, inv_c2;
};
static PicDerived pic_derived(const PfPic *p)
{
    PicDerived d;
    d.inv_dz = 1.0 / p->dz;
    d.zmax = (double)(p->L - 1) * p->dz;
    d.qmdt2 = p->q_over_m * p->dt * 0.5;
    d.inv_c2 = 1.0 / (p->c * p->c);
    return d;
}

// Boris push of one particle (the single definition used by every kernel, so that the counting pass
// and the moving pass of the fused re-sort see bit-identical results).
__device__ __forceinline__ Pushed pic_push_one(const PfPic &p, const PicDerived &D, double z, double ux, double uz)
{
    const double inv_dz = D.inv_dz;
    const double zmax = D.zmax;
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

    const double qmdt2 = D.qmdt2;
    const double inv_c2 = D.inv_c2;
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
    Pushed r;
    r.z = z; r.ux = uxn; r.uz = uzn;
    int cn = (int)floor(z * inv_dz);
    r.cell = max(0, min(cn, p.L - 2));
    return r;
}

__global__ void __launch_bounds__(PIC_THREADS) k_pic_push(PfPic p, PicDerived D)
{
    long long i = (long long)blockIdx.x * PIC_THREADS + threadIdx.x;
    if (i >= p.n) return;
    Pushed r = pic_push_one(p, D, p.z[i], p.ux[i], p.uz[i]);
    p.z[i] = r.z;
    p.ux[i] = r.ux;
    p.uz[i] = r.uz;
    p.cell[i] = r.cell;
}

// ------------------------------------------------------------------------------------------------ fused push + re-sort
// A particle moves less than one cell per step (|v| dt < c dt = 0.95 dz), so a cell-sorted set stays sorted up to
// exchanges between neighbouring cells.  The stable sort by new cell is then a LOCAL problem: the new population of cell c
// is [right-movers of c-1][stayers of c][left-movers of c+1], each group in its old order, and with L(c), R(c) the numbers
// of left- / right-movers of old cell c the new offsets follow from the old ones by boundary flows alone,
//     new_start[c] = start[c] - R(c-1) + L(c),
// so the first slot of each group needs only the counts of the cells c-1, c, c+1 -- no global scan:
//     stayers of c      : start[c] + L(c)                   left-movers of c : start[c] - R(c-1)
//     right-movers of c : start[c+1] - R(c) + L(c+1)        (plus, inside a cell, the counts of its earlier pieces).
//
// ONE pass over the particles (k_pic_step1).  A cell is cut into S contiguous pieces (sub_range), one warp per piece:
//   1. push the piece's particles and keep them ON CHIP (32 PIC_K = 128 per piece: z, ux, uz, w in the warp's own
//      4 KB of shared memory -- registers would cost the occupancy the latency-bound push lives on: a first version with
//      8 particles per lane in registers ran at 128 registers / 16 warps per SM and 1.33 ms per 2e7-particle step),
//      counting left / stay / right by ballot;
//   2. publish the three counts as one 64-bit word (valid bit + 3 x 20 bits; the words are zeroed before the launch);
//   3. wait for the words of every piece of the cells c-1, c, c+1 (lanes poll in parallel), derive the offsets;
//   4. write every particle straight to its final slot of the alternate arrays -- and, DEP = true, accumulate its CIC
//      shares (pf_pic_step_sorted: deposition of the new state while it is in registers).
// HBM traffic: 32 B read + 36 B written per particle-step (the two-pass version of round 1 stored the pushed state between
// its passes: 116 B).  Pieces longer than 32 PIC_K particles (a cell far above the average population) keep the old
// behaviour inside the same kernel: pushed state stored in place, re-read for placement.
// The grid is persistent (every CTA resident, pieces dealt round-robin to warps): a warp waits only for publications, and
// every warp publishes before it waits, so the wait graph has no cycle (a piece's neighbours are at most 2 S pieces away,
// far less than the number of warps in flight, so no warp ever waits for a piece of its own next round).
__global__ void __launch_bounds__(PIC_THREADS) k_pic_cell_start(const int *__restrict__ cell, long long n, int L,
                                                                long long *__restrict__ start)
{
    int c = blockIdx.x * PIC_THREADS + threadIdx.x;
    if (c > L) return;
    long long lo = 0, hi = n;            // lower bound of c in the sorted cell array
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        if (cell[mid] < c) lo = mid + 1; else hi = mid;
    }
    start[c] = lo;
}

// particle range of sub-warp s (of S) of a cell holding [a, b): contiguous pieces, multiples of 32
__device__ __forceinline__ void sub_range(long long a, long long b, int S, int s, long long &lo, long long &hi)
{
    const long long q = (((b - a) + S - 1) / S + 31) / 32 * 32;
    lo = min(b, a + (long long)s * q);
    hi = min(b, lo + q);
}

constexpr int PIC_K = 4;                      // particles per lane a piece keeps on chip (32 PIC_K per piece, in shared memory)
#ifndef PF_PIC_STEP_MINBLOCKS
#define PF_PIC_STEP_MINBLOCKS 5
#endif
constexpr int PIC_STEP_MINBLOCKS = PF_PIC_STEP_MINBLOCKS;   // 40 warps / SM: the push is fp64-latency bound and lives on thread-level parallelism
constexpr unsigned long long PUB_VALID = 1ull << 63;
constexpr int PUB_BITS = 20, PUB_MASK = (1 << PUB_BITS) - 1;

__device__ __forceinline__ unsigned long long pub_load(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void pub_store(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// DEP = true additionally deposits while the particles are in registers (pf_pic_step_sorted): lane l of piece (c, s)
// accumulates, in chunk order, the CIC shares of its particles at the four nodes c-1 .. c+2 a particle of old cell c can
// touch; the 32 lanes are combined by the fixed xor butterfly and the four sums go to part[(c*S+s)*4 + k].  k_pic_flush4
// then adds the partial sums of every node in a fixed order.  Deterministic (no atomics), but a different summation tree
// from pf_pic_deposit's -- oracle/pic_oracle.py: deposit_fused().
#ifndef PF_PIC_COUNT_MINBLOCKS
#define PF_PIC_COUNT_MINBLOCKS 5   // 48 registers, 40 warps per SM: 0.558 against 0.582 ms per 2e7-particle step with 64 registers / 32 warps (no spills; 6 spills, 0.555)
#endif
#ifndef PF_PIC_PLACE_MINBLOCKS
#define PF_PIC_PLACE_MINBLOCKS 4   // round 1: 62 instead of 78 registers for the depositing placement, 0.589 vs 0.616 ms per 2e7-particle step
#endif
template <bool DEP, int PASS>
__global__ void __launch_bounds__(PIC_THREADS, PASS == 0 ? PIC_STEP_MINBLOCKS : (PASS == 1 ? PF_PIC_COUNT_MINBLOCKS : PF_PIC_PLACE_MINBLOCKS)) k_pic_step1(PfPic p, PicDerived D, const long long *__restrict__ start,
                                                             long long *__restrict__ new_start, unsigned long long *pub,
                                                             int *__restrict__ err, int S, long long n_pieces,
                                                             double *__restrict__ part)
{
    constexpr int pass = PASS;
    // pass 0: the whole step in one launch (pieces wait for their neighbours' publications);
    // pass 1 / pass 2: the same step as two launches -- 1 = push + count + publish with the pushed state stored in place,
    // 2 = offsets + placement from the stored state (every word is published by then: no waiting)
    __shared__ double stage[PASS == 0 ? PIC_THREADS / 32 : 1][4][PASS == 0 ? 32 * PIC_K : 1];   // [warp][z, ux, uz, w][particle of the piece]
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    double(*my)[PASS == 0 ? 32 * PIC_K : 1] = stage[PASS == 0 ? (threadIdx.x >> 5) : 0];
    const long long warp0 = ((long long)blockIdx.x * PIC_THREADS + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * PIC_THREADS) >> 5;
    for (long long piece = warp0; piece < n_pieces; piece += n_warps) {
        const int c = (int)(piece / S), sub = (int)(piece % S);
        long long a, b;
        sub_range(start[c], start[c + 1], S, sub, a, b);
        const long long n = b - a;
        const bool fast = pass == 0 && n <= 32 * PIC_K;
        if (n > PUB_MASK) atomicExch(err, 2);            // a piece of more than 2^20 particles: counts would not fit their field
        unsigned dbits = 0;                               // 2 bits per held particle: d + 1 (0..2), 3 = no particle
        int nl = 0, ns = 0, nr = 0;
        // ---- 1. push + count ------------------------------------------------------------------------------------
        if (pass == 2) {
            // already pushed, counted and published by the first launch
        } else if (fast) {
#pragma unroll
            for (int k = 0; k < PIC_K; ++k) {
                const long long i = a + 32 * k + lane;
                int d = 2;
                if (32 * k < n) {                         // warp-uniform
                    if (i < b) {
                        Pushed r = pic_push_one(p, D, p.z[i], p.ux[i], p.uz[i]);
                        my[0][32 * k + lane] = r.z; my[1][32 * k + lane] = r.ux; my[2][32 * k + lane] = r.uz;
                        my[3][32 * k + lane] = p.w[i];
                        d = r.cell - c;
                        if (d < -1 || d > 1) { atomicExch(err, 1); d = max(-1, min(1, d)); }
                    }
                    nl += __popc(__ballot_sync(0xffffffffu, d == -1));
                    ns += __popc(__ballot_sync(0xffffffffu, d == 0));
                    nr += __popc(__ballot_sync(0xffffffffu, d == 1));
                }
                dbits |= (unsigned)(d + 1) << (2 * k);
            }
        } else {
            for (long long i0 = a; i0 < b; i0 += 32) {
                const long long i = i0 + lane;
                int d = 2;
                if (i < b) {
                    Pushed r = pic_push_one(p, D, p.z[i], p.ux[i], p.uz[i]);
                    p.z[i] = r.z;                         // pushed state, still in the old order: re-read for placement below
                    p.ux[i] = r.ux;
                    p.uz[i] = r.uz;
                    d = r.cell - c;
                    if (d < -1 || d > 1) { atomicExch(err, 1); d = max(-1, min(1, d)); }
                }
                nl += __popc(__ballot_sync(0xffffffffu, d == -1));
                ns += __popc(__ballot_sync(0xffffffffu, d == 0));
                nr += __popc(__ballot_sync(0xffffffffu, d == 1));
            }
        }
        // ---- 2. publish --------------------------------------------------------------------------------------------
        if (lane == 0 && pass != 2)
            pub_store(pub + piece, PUB_VALID | (unsigned long long)(nl & PUB_MASK) | ((unsigned long long)(ns & PUB_MASK) << PUB_BITS) |
                                       ((unsigned long long)(nr & PUB_MASK) << (2 * PUB_BITS)));
        if (pass == 1) continue;
        // ---- 3. counts of the cells c-1, c, c+1 -> first slots of the three groups -------------------------------
        int Lc = 0, Rc = 0, preL = 0, preS = 0, preR = 0, Rm = 0, Lp = 0;
        {
            const long long jlo = (long long)max(0, c - 1) * S, jhi = (long long)min(p.L, c + 2) * S;
            for (long long j0 = jlo; j0 < jhi; j0 += 32) {
                const long long j = j0 + lane;
                if (j < jhi) {
                    unsigned long long v = 0;
                    while (!((v = pub_load(pub + j)) & PUB_VALID)) __nanosleep(200);
                    const int l = (int)(v & PUB_MASK), st = (int)((v >> PUB_BITS) & PUB_MASK), r = (int)((v >> (2 * PUB_BITS)) & PUB_MASK);
                    const int cj = (int)(j / S), sj = (int)(j % S);
                    if (cj == c) {
                        Lc += l; Rc += r;
                        if (sj < sub) { preL += l; preS += st; preR += r; }
                    } else if (cj < c) {
                        Rm += r;
                    } else {
                        Lp += l;
                    }
                }
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                Lc += __shfl_xor_sync(0xffffffffu, Lc, off);
                Rc += __shfl_xor_sync(0xffffffffu, Rc, off);
                preL += __shfl_xor_sync(0xffffffffu, preL, off);
                preS += __shfl_xor_sync(0xffffffffu, preS, off);
                preR += __shfl_xor_sync(0xffffffffu, preR, off);
                Rm += __shfl_xor_sync(0xffffffffu, Rm, off);
                Lp += __shfl_xor_sync(0xffffffffu, Lp, off);
            }
        }
        const long long sc = start[c], sc1 = start[c + 1];
        long long posS = sc + Lc + preS, posL = sc - Rm + preL, posR = sc1 - Rc + Lp + preR;
        if (sub == 0 && lane == 0) {                      // the offsets of the NEW order (the next step's start[])
            new_start[c] = sc - Rm + Lc;
            if (c == p.L - 1) new_start[p.L] = sc1;
        }
        // ---- 4. place (+ deposit) ----------------------------------------------------------------------------------
        double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
        auto place = [&](bool have, int d, double z, double ux, double uz, double w) {
            const unsigned bl = __ballot_sync(0xffffffffu, d == -1);
            const unsigned bs = __ballot_sync(0xffffffffu, d == 0);
            const unsigned br = __ballot_sync(0xffffffffu, d == 1);
            if (have) {
                const long long dst = (d == -1) ? posL + __popc(bl & lt) : (d == 0) ? posS + __popc(bs & lt) : posR + __popc(br & lt);
                p.z_alt[dst] = z;
                p.ux_alt[dst] = ux;
                p.uz_alt[dst] = uz;
                p.w_alt[dst] = w;
                p.cell_alt[dst] = c + d;
            }
            if (DEP) {
                double t0 = 0.0, t1 = 0.0;
                if (have) {
                    const double g = sqrt(1.0 + (ux * ux + uz * uz) * D.inv_c2);
                    const double wv = w * (ux / g);
                    const double f = z * D.inv_dz - (double)(c + d);
                    t0 = wv * (1.0 - f);
                    t1 = wv * f;
                }
                // shares at nodes c-1, c, c+1, c+2 (adding 0.0 is exact, so absent lanes / other nodes do not perturb the sums)
                acc0 = acc0 + (d == -1 ? t0 : 0.0);
                acc1 = acc1 + (d == -1 ? t1 : (d == 0 ? t0 : 0.0));
                acc2 = acc2 + (d == 0 ? t1 : (d == 1 ? t0 : 0.0));
                acc3 = acc3 + (d == 1 ? t1 : 0.0);
            }
            posL += __popc(bl);
            posS += __popc(bs);
            posR += __popc(br);
        };
        if (fast) {
#pragma unroll
            for (int k = 0; k < PIC_K; ++k) {
                if (32 * k < n) {                         // warp-uniform
                    const int d = (int)((dbits >> (2 * k)) & 3u) - 1;     // 2 = no particle in this lane
                    const bool have = d != 2;                             // (each lane reads back only what it wrote itself)
                    place(have, d, have ? my[0][32 * k + lane] : 0.0, have ? my[1][32 * k + lane] : 0.0,
                          have ? my[2][32 * k + lane] : 0.0, have ? my[3][32 * k + lane] : 0.0);
                }
            }
        } else {
            for (long long i0 = a; i0 < b; i0 += 32) {
                const long long i = i0 + lane;
                int d = 2;
                double z = 0.0, ux = 0.0, uz = 0.0, w = 0.0;
                if (i < b) {
                    z = p.z[i]; ux = p.ux[i]; uz = p.uz[i]; w = p.w[i];
                    const int cn = (int)floor(z * D.inv_dz);     // same expression as pic_push_one's cell
                    d = max(-1, min(1, max(0, min(cn, p.L - 2)) - c));
                }
                place(i < b, d, z, ux, uz, w);
            }
        }
        if (DEP) {
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                acc0 = acc0 + __shfl_xor_sync(0xffffffffu, acc0, off);
                acc1 = acc1 + __shfl_xor_sync(0xffffffffu, acc1, off);
                acc2 = acc2 + __shfl_xor_sync(0xffffffffu, acc2, off);
                acc3 = acc3 + __shfl_xor_sync(0xffffffffu, acc3, off);
            }
            if (lane == 0) {
                double *o = part + ((size_t)c * S + sub) * 4;
                o[0] = acc0; o[1] = acc1; o[2] = acc2; o[3] = acc3;
            }
        }
    }
}

// node nz = scale * sum over old cells c = nz-2 .. nz+1 (ascending), sub-warps ascending, of part[c][s][nz - c + 1]
__global__ void __launch_bounds__(PIC_THREADS) k_pic_flush4(PfPic p, const double *__restrict__ part, int S)
{
    const int nz = blockIdx.x * PIC_THREADS + threadIdx.x;
    if (nz >= p.L) return;
    double v = 0.0;
    for (int c = nz - 2; c <= nz + 1; ++c) {
        if (c < 0 || c >= p.L) continue;
        const int k = nz - c + 1;
        for (int s2 = 0; s2 < S; ++s2) v = v + part[((size_t)c * S + s2) * 4 + k];
    }
    p.Jx[nz] = p.jx_scale * v;
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

// the fused push + re-sort runs up to this many warps per cell (each on a contiguous piece of the cell's particles)
#ifdef PF_PIC_SINGLE_PASS
// (sized so that a piece of an average cell is 3/4 of what a warp keeps on chip, 128 particles: cells up to a third above
//  the average population still take the on-chip path)
constexpr int PIC_SUB_MAX = 128;
static inline int pic_sub_warps(const PfPic *p)
{
    long long per_cell = p->n / std::max(1, p->L);
    return (int)std::min<long long>(PIC_SUB_MAX, std::max<long long>(1, (per_cell + 95) / 96));
}
#else
constexpr int PIC_SUB_MAX = 8;
static inline int pic_sub_warps(const PfPic *p)
{
    long long per_cell = p->n / std::max(1, p->L);
    return (int)std::min<long long>(PIC_SUB_MAX, std::max<long long>(1, (per_cell + 255) / 256));
}
#endif

static inline size_t al256(size_t x) { return (x + 255) / 256 * 256; }

static int key_bits(int L)
{
    int b = 1;
    while ((1 << b) < L) ++b;
    return b;
}

struct PicPlan {
    size_t off_idx_in, off_idx_out, off_cub, off_acc, off_start, off_new_start, off_counts, off_tot, off_part, off_err, cub_bytes, total;
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
    pl.off_start = pl.off_acc + al256(sizeof(double) * 2 * (size_t)p->L);
    pl.off_new_start = pl.off_start + al256(sizeof(long long) * ((size_t)p->L + 1));
    pl.off_counts = pl.off_new_start + al256(sizeof(long long) * ((size_t)p->L + 1));
    pl.off_tot = pl.off_counts + al256(sizeof(unsigned long long) * (size_t)p->L * PIC_SUB_MAX);   // publication words
    pl.off_part = pl.off_tot + al256(sizeof(int) * 3 * (size_t)p->L);
    pl.off_err = pl.off_part + al256(sizeof(double) * 4 * (size_t)p->L * PIC_SUB_MAX);
    pl.total = pl.off_err + 256;
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
__global__ void __launch_bounds__(PIC_THREADS) k_pic_cell_sums(PfPic p, PicDerived D, double *__restrict__ acc)
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
    const double inv_dz = D.inv_dz;
    const double inv_c2 = D.inv_c2;
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
    k_pic_push<<<blocks, PIC_THREADS, 0, (cudaStream_t)stream>>>(*p, pic_derived(p));
    PF_LAUNCH_CHECK("k_pic_push");
    return PF_OK;
}

static int pic_push_sorted(const PfPic *p, void *scratch, size_t scratch_bytes, void *stream, bool deposit)
{
    int rc = validate_pic(p);
    if (rc) return rc;
    if (!p->Ex || !p->Hy) return set_err(PF_E_ARG, "pf_pic_push_sorted: field arrays missing");
    if (deposit && !p->Jx) return set_err(PF_E_ARG, "pf_pic_step_sorted: Jx missing");
    if (p->n == 0) {
        if (deposit) PF_CUDA(cudaMemsetAsync(p->Jx, 0, sizeof(double) * (size_t)p->L, (cudaStream_t)stream));
        return PF_OK;
    }
    if (!p->z_alt || !p->ux_alt || !p->uz_alt || !p->w_alt || !p->cell_alt)
        return set_err(PF_E_ARG, "pf_pic_push_sorted: alternate (output) arrays missing");
    PicPlan pl = pic_plan(p);
    if (!scratch || scratch_bytes < pl.total) return set_err(PF_E_SCRATCH, "pf_pic_push_sorted needs %zu bytes of scratch", pl.total);
    cudaStream_t st = (cudaStream_t)stream;
    char *s = (char *)scratch;
    long long *start = (long long *)(s + pl.off_start), *new_start = (long long *)(s + pl.off_new_start);
    int *err = (int *)(s + pl.off_err);
    PF_CUDA(cudaMemsetAsync(err, 0, sizeof(int), st));
    if (p->flags & PF_PIC_F_OFFSETS_VALID) {
        // the previous fused call's new offsets ARE this call's offsets (caller's promise): 8 (L+1) bytes instead of a
        // binary search per cell
        PF_CUDA(cudaMemcpyAsync(start, new_start, sizeof(long long) * ((size_t)p->L + 1), cudaMemcpyDeviceToDevice, st));
    } else {
        k_pic_cell_start<<<(p->L + 1 + PIC_THREADS - 1) / PIC_THREADS, PIC_THREADS, 0, st>>>(p->cell, p->n, p->L, start);
        PF_LAUNCH_CHECK("k_pic_cell_start");
    }
    const int S = pic_sub_warps(p);
    unsigned long long *pub = (unsigned long long *)(s + pl.off_counts);
    const long long n_pieces = (long long)p->L * S;
    PF_CUDA(cudaMemsetAsync(pub, 0, sizeof(unsigned long long) * (size_t)n_pieces, st));
    const long long need = (n_pieces * 32 + PIC_THREADS - 1) / PIC_THREADS;
#ifdef PF_PIC_SINGLE_PASS
    // persistent grid: every CTA resident (the kernel's warps wait for one another's publications)
    int dev = 0, sms = 0, per_sm = 0;
    PF_CUDA(cudaGetDevice(&dev));
    PF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (deposit) PF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pic_step1<true, 0>, PIC_THREADS, 0));
    else PF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pic_step1<false, 0>, PIC_THREADS, 0));
    if (per_sm < 1) return set_err(PF_E_CUDA, "k_pic_step1 does not fit on an SM");
    const unsigned blocks = (unsigned)std::min<long long>(need, (long long)sms * per_sm);
    if ((long long)blocks < need && (long long)blocks * (PIC_THREADS / 32) <= 4LL * S)
        return set_err(PF_E_UNSUPPORTED, "pf_pic_step_sorted: too few resident warps for %d pieces per cell", S);
#else
    const unsigned blocks = (unsigned)need;
#endif
    double *part = (double *)(s + pl.off_part);
    // PF_PIC_SINGLE_PASS: one launch (pass 0).  Default: two launches of the same kernel (pass 1, pass 2) -- measured on a B200
    // at 2e7 particles the single launch is slower (1.28 ms against ~0.6 ms): its warps spend their time polling for the
    // neighbours' publications instead of hiding the push's fp64 latency (profiles/r2_pic.md)
#ifdef PF_PIC_SINGLE_PASS
    const int passes[2] = {0, -1};
#else
    const int passes[2] = {1, 2};
#endif
    for (int k = 0; k < 2 && passes[k] >= 0; ++k) {
        const int pass = passes[k];
        const unsigned grid = pass == 0 ? blocks : (unsigned)need;
        {
            ProfScope prof(st, pass == 1 ? "k_pic_step1<count>" : (deposit ? "k_pic_step1<place+deposit>" : "k_pic_step1<place>"));
#define PF_PIC_LAUNCH(DEPV, PASSV) k_pic_step1<DEPV, PASSV><<<grid, PIC_THREADS, 0, st>>>(*p, pic_derived(p), start, new_start, pub, err, S, n_pieces, DEPV ? part : nullptr)
#ifdef PF_PIC_SINGLE_PASS
            if (pass == 0) { if (deposit) PF_PIC_LAUNCH(true, 0); else PF_PIC_LAUNCH(false, 0); } else
#endif
            if (pass == 1) PF_PIC_LAUNCH(false, 1);          // (the count pass does not deposit)
            else { if (deposit) PF_PIC_LAUNCH(true, 2); else PF_PIC_LAUNCH(false, 2); }
#undef PF_PIC_LAUNCH
        }
        PF_LAUNCH_CHECK("k_pic_step1");
    }
    if (deposit) {
        k_pic_flush4<<<(p->L + PIC_THREADS - 1) / PIC_THREADS, PIC_THREADS, 0, st>>>(*p, part, S);
        PF_LAUNCH_CHECK("k_pic_flush4");
    }
    return PF_OK;
}

int pf_pic_push_sorted(const PfPic *p, void *scratch, size_t scratch_bytes, void *stream)
{
    return pic_push_sorted(p, scratch, scratch_bytes, stream, false);
}

int pf_pic_step_sorted(const PfPic *p, void *scratch, size_t scratch_bytes, void *stream)
{
    return pic_push_sorted(p, scratch, scratch_bytes, stream, true);
}

int pf_pic_sub_warps(const PfPic *p) { return p ? pic_sub_warps(p) : 1; }

int pf_pic_check(const PfPic *p, void *scratch, size_t scratch_bytes, void *stream)
{
    // 1 if the last pf_pic_push_sorted saw a particle cross more than one cell (CFL violation), else 0
    PicPlan pl = pic_plan(p);
    if (!scratch || scratch_bytes < pl.total) return set_err(PF_E_SCRATCH, "pf_pic_check: scratch too small");
    int h = 0;
    PF_CUDA(cudaMemcpyAsync(&h, (char *)scratch + pl.off_err, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    PF_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return h;
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
    k_pic_cell_sums<<<blocks, PIC_THREADS, 0, st>>>(*p, pic_derived(p), acc);
    PF_LAUNCH_CHECK("k_pic_cell_sums");
    k_pic_flush<<<(p->L + PIC_THREADS - 1) / PIC_THREADS, PIC_THREADS, 0, st>>>(*p, acc);
    PF_LAUNCH_CHECK("k_pic_flush");
    return PF_OK;
}

}  // extern "C"
