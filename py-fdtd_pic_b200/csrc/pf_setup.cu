// pf_setup.cu -- host-side setup of a whole sweep in one native call (no kernels here).
//
// The reference rebuilds, per sweep member and per pass, the CPML profiles (BaseFDTD11.CPML_ScalingCalc :222-272,
// CPML_Ex_RC_Define :274-286, CPML_HY_RC_Define :288-296) and the source tables (Solver_Engine.SourceManager :89-124
// over BaseFDTD11.SmoothTurnOn :104-120) in Python / numba loops.  Once the stepping runs on the GPU that chain is
// what a sweep waits for (14 ms per member in the Python mirror against ~1 ms of GPU time), so this file restates it
// as one loop over members that writes straight into the caller's (pinned) upload buffer, on several host threads.
//
// Parity: every value is produced by the same IEEE operations in the same order as the Python mirror
// (py-fdtd_pic_b200/BaseFDTD11.py, Solver_Engine.py), which is pinned bit for bit to the reference goldens;
// exp / pow / sin are the C library's, exactly what the numba-compiled reference loops and numpy's float64 sin call
// (tests/test_sweep_setup.py holds this builder to bit-equality with the per-member chain).  This translation unit is
// compiled like the rest of the library with --fmad=false and the host compiler's default -ffp-contract=off for
// nvcc-generated host code, so no multiply-add is fused.
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>
#include "pf_common.cuh"

namespace pf {

// CPML profiles of one member: b_e (= b_m), c_e, c_m on the 2*pw CPML cells, 0 elsewhere.
static void setup_profiles(const PfSetupMember &m, double *out, std::vector<double> &tmp)
{
    const int L = m.L, pw = m.pw;
    double *be = out + m.off_beX, *ce = out + m.off_ceX, *cm = out + m.off_cmY;
    memset(be, 0, sizeof(double) * (size_t)L);
    memset(ce, 0, sizeof(double) * (size_t)L);
    memset(cm, 0, sizeof(double) * (size_t)L);
    if (pw <= 0 || 2 * pw > L) return;
    tmp.resize(3 * (size_t)pw);
    double *pb = tmp.data(), *pce = pb + pw, *pcm = pce + pw;
    for (int n = 0; n < pw; ++n) {
        // CPML_ScalingCalc: polynomial grading (libm pow, as numba calls it)
        const double depth = (double)(pw - n) / (double)pw;
        const double graded = pow(depth, m.r_scale);
        const double ramp = pow((double)(n + 1) / (double)pw, m.r_a_scale);
        const double kap = 1.0 + (m.kappaMax - 1.0) * graded;
        const double sig = m.sigmaOpt * graded;
        const double alp = m.alphaMax * ramp;
        // CPML_Ex_RC_Define / CPML_HY_RC_Define
        const double arg = -((sig * m.delT / (kap * m.eps0)) + ((alp * m.delT) / m.eps0));
        const double b = exp(arg);
        const double den = sig * kap + alp * kap * kap;
        pb[n] = b;
        pce[n] = (b - 1.0) * sig / den;
        pcm[n] = (b - 1.0) * sig / (den * m.dz);
    }
    for (int n = 0; n < pw; ++n) {   // left: cell n; right: mirrored (CPML_ScalingCalc: dst[L-pw:L] = prof[::-1])
        be[n] = pb[n]; ce[n] = pce[n]; cm[n] = pcm[n];
        const int r = L - pw + n;
        be[r] = pb[pw - 1 - n]; ce[r] = pce[pw - 1 - n]; cm[r] = pcm[pw - 1 - n];
    }
}

// Source tables of one member: srcE[n] = Exs[n]*amp/courantNo, srcH[n] = Hys[n]*amp/courantNo with Exs, Hys as
// SourceManager builds them for P.SineCont (TF/SF scaling and the nonlinear run's pump at 0.8 f included).
static void setup_sources(const PfSetupMember &m, double *out)
{
    double *sE = out + m.off_srcE, *sH = out + m.off_srcH;
    const int n_src = m.n_src;
    if (m.src_kind != 1) return;
    const double cN = m.courantNo;
    const double ppw = m.c0 / (m.freq * m.dz);
    const double w = 2.0 * M_PI / ppw;
    const double fp = m.freq * 0.8;
    const double ppw_p = m.c0 / (fp * m.dz);
    const double w_p = 2.0 * M_PI / ppw_p;
    const double t_on = m.period * m.periods;
    const double inv_imp = 1.0 / m.charImp;
    // Hys[n] = sin(w (cN (n+1))) is the same expression as Exs[n+1]: one sine per step, carried over (and none at all once
    // the source is switched off -- `on` is monotone in n)
    double s_cur = 0.0, p_cur = 0.0;
    bool have = false;
    for (int n = 0; n < n_src; ++n) {
        const bool on = (double)n * m.delT < t_on;
        double e = 0.0, h = 0.0, ep = 0.0, hp = 0.0;
        if (on) {
            if (!have) {
                s_cur = sin(w * (cN * (double)n));
                if (m.pump) p_cur = sin(w_p * (cN * (double)n));
                have = true;
            }
            const double s_nxt = sin(w * (cN * (double)(n + 1)));
            e = s_cur;
            h = s_nxt;
            s_cur = s_nxt;
            if (m.pump) {
                const double p_nxt = sin(w_p * (cN * (double)(n + 1)));
                ep = p_cur;
                hp = p_nxt;
                p_cur = p_nxt;
            }
        }
        if (m.pump) {
            ep = ep * cN;
            hp = hp * cN;
            ep = ep * 0.1;
            hp = hp * 0.01;
        }
        e = e * cN + ep;
        h = h * cN + hp;
        if (m.tfsf) h = h * inv_imp;
        e = e * m.amp;
        h = h * m.amp;
        sE[n] = e / cN;
        sH[n] = h / cN;
    }
}

}  // namespace pf

using namespace pf;

extern "C" {

int pf_host_sweep_inputs(const PfSetupMember *members, int n_members, double *out, int threads)
{
    if (n_members < 0 || (n_members > 0 && (!members || !out))) return set_err(PF_E_ARG, "pf_host_sweep_inputs: bad arguments");
    for (int i = 0; i < n_members; ++i) {
        const PfSetupMember &m = members[i];
        if (m.L <= 0 || m.pw < 0 || m.n_src < 0 || (m.src_kind != 0 && m.src_kind != 1))
            return set_err(PF_E_ARG, "pf_host_sweep_inputs: bad member %d", i);
    }
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads <= 0) threads = 1;
    if (threads > n_members) threads = n_members > 0 ? n_members : 1;
    auto work = [&](int t) {
        std::vector<double> tmp;
        for (int i = t; i < n_members; i += threads) {
            const PfSetupMember &m = members[i];
            if (m.off_beX >= 0) setup_profiles(m, out, tmp);
            if (m.off_srcE >= 0) setup_sources(m, out);
        }
    };
    if (threads == 1) {
        work(0);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) pool.emplace_back(work, t);
        for (auto &th : pool) th.join();
    }
    return PF_OK;
}

// elementwise libm maps on several host threads (the numpy-vectorised scalar chain of sweep_setup.py uses them where
// the reference evaluates Python-float `**` / math functions one value at a time)
static int host_map(const double *x, double *y, long long n, int threads, double (*fn)(double, double), double e)
{
    if (n < 0 || (n > 0 && (!x || !y))) return set_err(PF_E_ARG, "pf_host_map: bad arguments");
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads <= 0 || n < 4096) threads = 1;
    auto work = [&](int t) {
        const long long lo = n * t / threads, hi = n * (t + 1) / threads;
        for (long long i = lo; i < hi; ++i) y[i] = fn(x[i], e);
    };
    if (threads == 1) { work(0); return PF_OK; }
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(work, t);
    for (auto &th : pool) th.join();
    return PF_OK;
}

int pf_host_sin(const double *x, double *y, long long n, int threads)
{
    return host_map(x, y, n, threads, [](double v, double) { return sin(v); }, 0.0);
}
int pf_host_cos(const double *x, double *y, long long n, int threads)
{
    return host_map(x, y, n, threads, [](double v, double) { return cos(v); }, 0.0);
}

}  // extern "C"
