// pf_halo.cu -- ghost-zone pack / unpack for the 1-D domain decomposition of a long grid (config 5).
// A rank's local arrays hold k ghost cells next to every interior boundary.  Every k steps the k
// owned cells adjacent to a boundary are packed into one contiguous buffer (all state arrays of the
// mode back to back), shipped to the neighbour (NCCL send/recv or a peer copy, done by the caller)
// and unpacked into the neighbour's ghost cells.  Tiny messages: 7*k doubles per side at most.
#include "pf_common.cuh"

namespace pf {

struct HaloPtrs {
    double *a[7];
    int n;
};

static HaloPtrs halo_arrays(const PfGrid *g, int mode)
{
    HaloPtrs h;
    double *all[7] = {g->Ex, g->Hy, g->psiE, g->psiH, g->Dx, g->P, g->Pprev};
    h.n = (mode == PF_LORENTZ || mode == PF_LORENTZ_NL) ? 7 : (mode == PF_NL ? 5 : 4);
    for (int i = 0; i < 7; ++i) h.a[i] = all[i];
    return h;
}

__global__ void k_halo_copy(HaloPtrs h, int start, int k, double *buf, int to_buf)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= h.n * k) return;
    int a = i / k, j = i - a * k;
    if (to_buf) buf[i] = h.a[a] ? h.a[a][start + j] : 0.0;
    else if (h.a[a]) h.a[a][start + j] = buf[i];
}

static long long halo_move(const PfGrid *g, int mode, int side, int k, double *buf, int to_buf, cudaStream_t st)
{
    if (!g || !buf || k <= 0 || 2 * k > g->L) return set_err(PF_E_ARG, "halo: bad arguments (k=%d, L=%d)", k, g ? g->L : 0);
    HaloPtrs h = halo_arrays(g, mode);
    // arrays a piece does not carry (no CPML / slab cell in it) travel as zeros and are not unpacked
    // pack (to_buf): the k owned cells next to the edge; unpack: the k ghost cells at the edge
    int start;
    if (side == 0) start = to_buf ? k : 0;
    else start = to_buf ? g->L - 2 * k : g->L - k;
    int n = h.n * k;
    k_halo_copy<<<(n + 255) / 256, 256, 0, st>>>(h, start, k, buf, to_buf);
    ++g_launches;
    int rc = check_cuda(cudaGetLastError(), "k_halo_copy");
    if (rc) return rc;
    return n;
}

}  // namespace pf

extern "C" {

long long pf_halo_pack(const PfGrid *g, int mode, int side, int k, double *buf, void *stream)
{
    return pf::halo_move(g, mode, side, k, buf, 1, (cudaStream_t)stream);
}

long long pf_halo_unpack(const PfGrid *g, int mode, int side, int k, const double *buf, void *stream)
{
    return pf::halo_move(g, mode, side, k, const_cast<double *>(buf), 0, (cudaStream_t)stream);
}

}  // extern "C"
