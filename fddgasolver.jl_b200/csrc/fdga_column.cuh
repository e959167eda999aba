// fdga_column.cuh -- "column" kernels: the optimised form of the full-momentum gather contractions
//   BSE_K2! (fd / mfRG)   src/nonlocal_2/BSEa/BSEa_K2.jl:72-125
//   BSE_L_K2!             src/nonlocal_2/BSEa/BSEa_K2.jl:17-43
//   SDE_channel_L_pp!/ph! src/nonlocal_2/SDE.jl:16-33, 54-73
//
// All four have the shape  out[W,nu,P,k] = scale * sum_{w,q} L(nu,k; w,q) * R[w,q | W,P]  with a left factor L made of
// vertex evaluations.  Work decomposition:
//   * one CTA per output COLUMN (W, P, k) holding up to NV class representatives (different nu);
//   * one thread per inner momentum q (and a contiguous chunk of inner frequencies w): every momentum conversion /
//     Brillouin-zone fold / table base offset is computed ONCE per (column, q) and kept in registers, so the innermost
//     loops over (w, nu) only do Matsubara box tests and gathers at fixed momenta;
//   * momentum-independent levels of the F0 chain (local Vertex, RefVertex core) are pre-tabulated per (W, nu, w)
//     by loc_table_kernel with the generic evaluator and enter as one cached load per term;
//   * the own-channel nu -> infinity differences are taken analytically (K2[W,nu,P,k] + K3[W,nu,w,P] inside the boxes)
//     instead of evaluating F twice (the reference's (F(nu) - F(inf)) differs from this by rounding only).
#pragma once
#include "fdga_kernels.cuh"

namespace fdga {

#define FDGA_NV 8    // representatives (nu values) per column chunk

enum { JOB_K2 = 0, JOB_K2_MF = 1, JOB_LK2 = 2, JOB_SDE_PP = 3, JOB_SDE_PH = 4,
       JOB_LK2_LOC = 5 /* BSE_L_K2! of the local solver, src/BSEa/BSEa_K2.jl:17-34: left vertex at uncrossed arguments */ };

struct ColDev {                 // columns of class representatives, built on the host (fdga_lib.cu: build_columns)
    int ncol;
    const int* iW;              // position of W in the OUTPUT (K2) bosonic mesh
    const int* iP;
    const int* ik;
    const int* start;           // ncol + 1
    const int* rep_inu;         // position of nu in the K2 fermionic mesh
    const int* rep_cls;         // class id (slot in repvals)
    int ngrp;                   // groups of <= FDGA_WGROUP consecutive columns sharing (P, k): one CTA each
    const int* grp_start;       // ngrp + 1 (column index)
};

struct ColJob {
    int lev_first;              // K2 / LK2: first level of the left chain; SDE: the level l of the recursion
    int n_nl2;                  // number of leading NL2 levels in the chain
    int own_only;               // SDE: SURVEY-E2 toggle
    int k1_direct;              // 1: cross-channel K1 terms summed inside the column kernel; 0: by slab_conv_kernel (momentum convolution)
    int nw, Ninner;             // inner frequency mesh (count, N)
    int slabW_N;                // bosonic mesh N used to index the R slab ([w + nw*(q + NP*(posB(W) + nB*iP))])
    double scale_re, scale_im;  // complex prefactor applied at the end
    const int* slabmap;         // (position of W in the slab mesh, P) -> slab number of R (compact storage; null: natural order)
};

struct MomOff { int oK1, oK2A, oK2B, oK3; };

FDGA_HD int fold1(int a, int L) {   // a in (-3L, 3L)
    a += (a < 0) ? L : 0; a += (a < 0) ? L : 0; a += (a < 0) ? L : 0;
    a -= (a >= L) ? L : 0; a -= (a >= L) ? L : 0; a -= (a >= L) ? L : 0;
    return a;
}
FDGA_HD int foldidx(int x, int y, int L) { return fold1(x, L) + L * fold1(y, L); }

// offsets of the frequency sub-arrays of channel r of level lv at momenta converted from `form` to r
FDGA_HD MomOff mom_offsets(const DevLevel& lv, int form, int r, int L, int NP,
                                              int Px, int Py, int kx, int ky, int qx, int qy) {
    Arg a; a.W = 0; a.v = 0; a.w = 0; a.Px = Px; a.Py = Py; a.kx = kx; a.ky = ky; a.qx = qx; a.qy = qy;
    Arg b = convert(a, form, r);
    int iP = foldidx(b.Px, b.Py, L), ik = foldidx(b.kx, b.ky, L), iq = foldidx(b.qx, b.qy, L);
    int nB2 = 2 * lv.nK2b - 1, nF2 = 2 * lv.nK2f, nB3 = 2 * lv.nK3b - 1, nF3 = 2 * lv.nK3f;
    MomOff m;
    m.oK1 = (2 * lv.nK1 - 1) * iP;
    m.oK2A = nB2 * nF2 * (iP + NP * ik);
    m.oK2B = nB2 * nF2 * (iP + NP * iq);
    m.oK3 = nB3 * nF3 * nF3 * iP;
    return m;
}
// gamma_r at fixed momenta, all K switches on; v, w finite (same box logic as nl2_chan)
FDGA_HD C chan_off(const DevLevel& lv, int r, const MomOff& m, int W, int v, int w) {
    C val = zeroC();
    if (!inB(W, lv.nK1)) return val;
    const DevChan& c = lv.ch[r];
    val += ldg(c.K1 + m.oK1 + posB(W, lv.nK1));
    if (!inB(W, lv.nK2b)) return val;
    const bool a = inF(v, lv.nK2f), b = inF(w, lv.nK2f);
    const int nB = 2 * lv.nK2b - 1;
    const int pW = posB(W, lv.nK2b);
    if (a) val += ldg(c.K2 + m.oK2A + pW + nB * posF(v, lv.nK2f));
    if (b) val += ldg(c.K2 + m.oK2B + pW + nB * posF(w, lv.nK2f));
    if (a && b && inB(W, lv.nK3b) && inF(v, lv.nK3f) && inF(w, lv.nK3f)) {
        const int nB3 = 2 * lv.nK3b - 1, nF3 = 2 * lv.nK3f;
        val += ldg(c.K3 + m.oK3 + posB(W, lv.nK3b) + nB3 * (posF(v, lv.nK3f) + nF3 * posF(w, lv.nK3f)));
    }
    return val;
}
// gamma_r(W, v, w) - gamma_r(W, inf, w) at fixed momenta: K2[W,v,P,k] + K3[W,v,w,P] inside the boxes
FDGA_HD C chan_off_diff_v(const DevLevel& lv, int r, const MomOff& m, int W, int v, int w) {
    C val = zeroC();
    if (!inB(W, lv.nK2b) || !inF(v, lv.nK2f)) return val;     // K2 Omega-box is inside the K1 box
    const DevChan& c = lv.ch[r];
    const int nB = 2 * lv.nK2b - 1;
    val += ldg(c.K2 + m.oK2A + posB(W, lv.nK2b) + nB * posF(v, lv.nK2f));
    if (inF(w, lv.nK2f) && inB(W, lv.nK3b) && inF(v, lv.nK3f) && inF(w, lv.nK3f)) {
        const int nB3 = 2 * lv.nK3b - 1, nF3 = 2 * lv.nK3f;
        val += ldg(c.K3 + m.oK3 + posB(W, lv.nK3b) + nB3 * (posF(v, lv.nK3f) + nF3 * posF(w, lv.nK3f)));
    }
    return val;
}

// frequency part of _convert_channel (no momenta)
FDGA_HD void convert_freq(int W, int v, int w, int from, int to, int& W2, int& v2, int& w2) {
    Arg a; a.W = W; a.v = v; a.w = w; a.Px = a.Py = a.kx = a.ky = a.qx = a.qy = 0;
    Arg b = convert(a, from, to);
    W2 = b.W; v2 = b.v; w2 = b.w;
}

// ---- interval machinery: after channel conversion every Matsubara argument is LINEAR in the inner index `win`
// with slope in {-1, 0, +1}, so "argument inside a mesh box" is a contiguous interval of win.
struct Lin { int x0, s; };      // x(win) = x0 + s * win
FDGA_HD void clip_interval(const Lin& x, int lo, int hi, int& a, int& b) {   // intersect [a, b] with {win : lo <= x(win) <= hi}
    if (x.s > 0) { a = max(a, lo - x.x0); b = min(b, hi - x.x0); }
    else if (x.s < 0) { a = max(a, x.x0 - hi); b = min(b, x.x0 - lo); }
    else if (x.x0 < lo || x.x0 > hi) { b = a - 1; }
}
// streaming correlation  sum_{win in [a, b]} tab[o + st * win] * Rq[(win + Nin) * rst]  with two independent accumulators
// (rst = NP: the slab is stored with the inner momentum fastest, slab_at);
// a constant table entry (st == 0) over the full chunk uses the pre-summed R of the chunk instead of a loop
FDGA_HD C stream_sum(const C* __restrict__ tab, int o, int st, int a, int b, const C* __restrict__ Rq, int Nin, int rst,
                     int a0, int b0, C rs_chunk) {
    if (a > b) return zeroC();
    if (st == 0 && a == a0 && b == b0) return ldg(tab + o) * rs_chunk;
    C p0 = zeroC(), p1 = zeroC();
    int win = a;
#pragma unroll 4
    for (; win + 1 <= b; win += 2) {
        p0 += ldg(tab + (o + st * win)) * Rq[(size_t)(win + Nin) * rst];
        p1 += ldg(tab + (o + st * (win + 1))) * Rq[(size_t)(win + 1 + Nin) * rst];
    }
    if (win <= b) p0 += ldg(tab + (o + st * win)) * Rq[(size_t)(win + Nin) * rst];
    return p0 + p1;
}
// sum over iw in [w_lo, w_hi) of gamma_r(W2, v2, w2) * Rq[iw] at fixed momenta (all K switches on), W2/v2/w2 linear in win = iw - Nin.
// Same box logic as chan_off: every term is a streaming correlation over its analytically clipped in-box interval; only the
// (rare) K3 term is evaluated term by term inside the K2 band.  rs_chunk = sum of Rq over the chunk.
FDGA_HD C chan_lin_sum(const DevLevel& lv, int r, const MomOff& m, Lin W2, Lin v2, Lin w2,
                       const C* __restrict__ Rq, int Nin, int rst, int w_lo, int w_hi, C rs_chunk, bool withK1) {
    const DevChan& c = lv.ch[r];
    const int a0 = w_lo - Nin, b0 = w_hi - 1 - Nin;      // inclusive win range of this chunk
    C part = zeroC();
    if (withK1) {
        int a = a0, b = b0; clip_interval(W2, -(lv.nK1 - 1), lv.nK1 - 1, a, b);
        part = stream_sum(c.K1, m.oK1 + posB(W2.x0, lv.nK1), W2.s, a, b, Rq, Nin, rst, a0, b0, rs_chunk);
    }
    int a = a0, b = b0; clip_interval(W2, -(lv.nK2b - 1), lv.nK2b - 1, a, b);
    if (a > b) return part;
    const int nB = 2 * lv.nK2b - 1, nB3 = 2 * lv.nK3b - 1, nF3 = 2 * lv.nK3f;
    {
        int aa = a, bb = b; clip_interval(v2, -lv.nK2f, lv.nK2f - 1, aa, bb);
        part += stream_sum(c.K2, m.oK2A + posB(W2.x0, lv.nK2b) + nB * posF(v2.x0, lv.nK2f), W2.s + nB * v2.s, aa, bb, Rq, Nin, rst, a0, b0, rs_chunk);
    }
    {
        int aa = a, bb = b; clip_interval(w2, -lv.nK2f, lv.nK2f - 1, aa, bb);
        part += stream_sum(c.K2, m.oK2B + posB(W2.x0, lv.nK2b) + nB * posF(w2.x0, lv.nK2f), W2.s + nB * w2.s, aa, bb, Rq, Nin, rst, a0, b0, rs_chunk);
    }
    {   // K3: short loop inside the K3 Omega-box band, explicit box tests (see DESIGN.md "toolchain pitfall")
        int aa = a, bb = b; clip_interval(W2, -(lv.nK3b - 1), lv.nK3b - 1, aa, bb);
        for (int win = aa; win <= bb; ++win) {
            const int Wc = W2.x0 + W2.s * win, vc = v2.x0 + v2.s * win, wc = w2.x0 + w2.s * win;
            if (inF(vc, lv.nK2f) && inF(wc, lv.nK2f) && inF(vc, lv.nK3f) && inF(wc, lv.nK3f))
                part += ldg(c.K3 + (m.oK3 + posB(Wc, lv.nK3b) + nB3 * (posF(vc, lv.nK3f) + nF3 * posF(wc, lv.nK3f)))) * Rq[(size_t)(win + Nin) * rst];
        }
    }
    return part;
}
// same for gamma_r(W, v, w) - gamma_r(W, inf, w): K2[W, v | P, k] + K3[W, v, w | P] inside the boxes
FDGA_HD C chan_lin_sum_diff_v(const DevLevel& lv, int r, const MomOff& m, Lin W2, Lin v2, Lin w2,
                              const C* __restrict__ Rq, int Nin, int rst, int w_lo, int w_hi, C rs_chunk) {
    const DevChan& c = lv.ch[r];
    const int a0 = w_lo - Nin, b0 = w_hi - 1 - Nin;
    int a = a0, b = b0;
    clip_interval(W2, -(lv.nK2b - 1), lv.nK2b - 1, a, b);
    clip_interval(v2, -lv.nK2f, lv.nK2f - 1, a, b);
    if (a > b) return zeroC();
    const int nB = 2 * lv.nK2b - 1, nB3 = 2 * lv.nK3b - 1, nF3 = 2 * lv.nK3f;
    C part = stream_sum(c.K2, m.oK2A + posB(W2.x0, lv.nK2b) + nB * posF(v2.x0, lv.nK2f), W2.s + nB * v2.s, a, b, Rq, Nin, rst, a0, b0, rs_chunk);
    int aa = a, bb = b; clip_interval(W2, -(lv.nK3b - 1), lv.nK3b - 1, aa, bb);
    for (int win = aa; win <= bb; ++win) {
        const int Wc = W2.x0 + W2.s * win, vc = v2.x0 + v2.s * win, wc = w2.x0 + w2.s * win;
        if (inF(wc, lv.nK2f) && inF(vc, lv.nK3f) && inF(wc, lv.nK3f))
            part += ldg(c.K3 + (m.oK3 + posB(Wc, lv.nK3b) + nB3 * (posF(vc, lv.nK3f) + nF3 * posF(wc, lv.nK3f)))) * Rq[(size_t)(win + Nin) * rst];
    }
    return part;
}

// forms (channel parametrisations evaluated in parallel spin) and their weights:
//   K2 / LK2 jobs: p -> {(p,1)}, a -> {(a,1)}, t (dSp = 2 pSp + xSp, xSp(t) = -a-form) -> {(t,2),(a,-1)}
//   SDE pp -> {(p,1)} ; SDE ph -> {(a,1),(t,1)}
template <int KIND, int CH> struct Forms;
template <int KIND> struct Forms<KIND, CH_P> { static constexpr int n = 1; __host__ __device__ static int ch(int) { return CH_P; } __host__ __device__ static double coef(int) { return 1.0; } };
template <int KIND> struct Forms<KIND, CH_A> { static constexpr int n = 1; __host__ __device__ static int ch(int) { return CH_A; } __host__ __device__ static double coef(int) { return 1.0; } };
template <int KIND> struct Forms<KIND, CH_T> { static constexpr int n = 2; __host__ __device__ static int ch(int i) { return i == 0 ? CH_T : CH_A; } __host__ __device__ static double coef(int i) { return i == 0 ? 2.0 : -1.0; } };
template <> struct Forms<JOB_SDE_PH, CH_A> { static constexpr int n = 2; __host__ __device__ static int ch(int i) { return i == 0 ? CH_A : CH_T; } __host__ __device__ static double coef(int) { return 1.0; } };

// map (output nu, inner w) -> vertex frequency arguments (v, w) of the job
template <int KIND, int CH>
FDGA_HD void job_freq_args(int W, int nu, int win, int& v, int& w) {
    if (KIND == JOB_K2 || KIND == JOB_SDE_PH || KIND == JOB_LK2_LOC) { v = nu; w = win; }
    else if (KIND == JOB_K2_MF || KIND == JOB_LK2) { v = nu; w = (CH == CH_P) ? W - win - 1 : win; }
    else { v = W - win - 1; w = nu; }                                   // JOB_SDE_PP: F(W, W - w, nu, ...)
}

// momentum arguments (k, q) of the vertex in the form-channel parametrisation, from the column momentum k and the inner q
template <int KIND, int CH>
FDGA_HD void job_mom_args(int Px, int Py, int kx, int ky, int qx, int qy, int& akx, int& aky, int& aqx, int& aqy) {
    if (KIND == JOB_K2 || KIND == JOB_SDE_PH || KIND == JOB_LK2_LOC) { akx = kx; aky = ky; aqx = qx; aqy = qy; }
    else if (KIND == JOB_K2_MF || KIND == JOB_LK2) { akx = kx; aky = ky; aqx = (CH == CH_P) ? Px - qx : qx; aqy = (CH == CH_P) ? Py - qy : qy; }
    else { akx = Px - qx; aky = Py - qy; aqx = kx; aqy = ky; }              // SDE pp: (P, P - q, k)
}

// ---- momentum-independent part of the left factor, tabulated per (W, nu, w) -----------------------------------
// T[iw + nw*(inu + nF2*iWo)]  (W on the output K2 bosonic mesh, nu on the output K2 fermionic mesh)
template <int KIND, int CH>
FDGA_HD C loc_table_entry(const DevChain& V, const ColJob& job, const Grid& g, long long i) {
    typedef Forms<KIND, CH> FM;
    const int nF2 = 2 * g.nK2f;
    int iw = i % job.nw; int inu = (i / job.nw) % nF2; int iWo = i / ((long long)job.nw * nF2);
    int W = iWo - (g.nK2b - 1), nu = inu - g.nK2f, win = iw - job.Ninner;
    Arg a; a.W = W; a.Px = a.Py = a.kx = a.ky = a.qx = a.qy = 0;
    job_freq_args<KIND, CH>(W, nu, win, a.v, a.w);
    C val = zeroC();
    const int lloc = job.n_nl2;                     // first momentum-independent level of the chain
#pragma unroll
    for (int f = 0; f < FM::n; ++f) {
        const int form = FM::ch(f);
        C x = zeroC();
        if (KIND == JOB_K2 || KIND == JOB_K2_MF) {
            if (lloc < V.nlev) {
                Arg ai = a; ai.v = FDGA_INF;
                x = eval_vertex<false>(V, max(lloc, job.lev_first), form, SP_P, a, FL_ALL) - eval_vertex<false>(V, max(lloc, job.lev_first), form, SP_P, ai, FL_ALL);
            }
        } else if (KIND == JOB_SDE_PP || KIND == JOB_SDE_PH) {
            // all momentum-independent pieces of the fused recursion: own gamma of every LOCAL level m >= l0, the cross
            // channels of every LOCAL level m > l0 (E2), and (core - U) / 3 for the terminating RefVertex
            for (int m = job.lev_first; m < V.nlev; ++m) {
                const DevLevel& lv = V.lev[m];
                if (lv.type == LV_CORE) { x += (core_eval(lv, form, SP_P, a.W, a.v, a.w) - lv.U) * (1.0 / 3.0); }
                else if (lv.type == LV_LOCAL) {
                    x += loc_chan(lv, form, a.W, a.v, a.w);
                    if (!job.own_only && m > job.lev_first) {
#pragma unroll
                        for (int r = 0; r < 3; ++r) if (r != form) { Arg b = convert(a, form, r); x += loc_chan(lv, r, b.W, b.v, b.w); }
                    }
                }
            }
        }
        val += x * FM::coef(f);
    }
    return val;
}
template <int KIND, int CH>
__global__ void loc_table_kernel(const __grid_constant__ DevChain V, ColJob job, Grid g, C* __restrict__ T) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long n = (long long)job.nw * (2 * g.nK2f) * (2 * g.nK2b - 1);
    if (i < n) T[i] = loc_table_entry<KIND, CH>(V, job, g, i);
}

// ---- own-channel pieces, hoisted out of the column kernel ------------------------------------------------------
// For every job the own-channel gamma of a level, summed against R over (win, q), splits exactly into
//     sum_{win,q} A'(win,q) R[win,q]  +  B(nu,k) * Rtot  +  sum_win K3(nu,win) Rq[win]
// (Rq[win] = sum_q R[win,q], Rtot = sum_win Rq[win]) with the same box logic as chan_off / chan_off_diff_v:
//   K2 jobs (nu -> inf difference):  A' = 0,  B = K2[W,nu|P,k],  K3 = K3[W,nu,w(win)|P]
//   SDE pp  gamma_p(W, W-w, nu; P, P-q, k):  A' = K1[W|P] + K2[W,W-w|P,P-q],  B = K2[W,nu|P,k],  K3 = K3[W,W-w,nu|P]
//   SDE ph  gamma_f(W, nu, w; P, k, q):      A' = K1[W|P] + K2[W,w|P,q],      B = K2[W,nu|P,k],  K3 = K3[W,nu,w|P]
// Only B depends on the column momentum k; everything else is a per-slab quantity computed once by slab_own_kernel.
template <int KIND, int CH>
FDGA_HD C own_A_term(const DevChain& V, const ColJob& job, const Grid& g, int W, int iP, int win, int iq) {
    typedef Forms<KIND, CH> FM;
    C val = zeroC();
    if (!(KIND == JOB_SDE_PP || KIND == JOB_SDE_PH)) return val;
    const int L = g.L, NP = g.NP;
#pragma unroll
    for (int f = 0; f < FM::n; ++f) {
        const int form = FM::ch(f);
        C x = zeroC();
        for (int l = job.lev_first; l < job.n_nl2; ++l) {
            const DevLevel& lv = V.lev[l];
            const DevChan& c = lv.ch[form];
            if (!inB(W, lv.nK1)) continue;
            x += ldg(c.K1 + (posB(W, lv.nK1) + (2 * lv.nK1 - 1) * iP));
            if (!inB(W, lv.nK2b)) continue;
            const int nB = 2 * lv.nK2b - 1, nF = 2 * lv.nK2f;
            int vv, kq;
            if (KIND == JOB_SDE_PP) { vv = W - win - 1; kq = foldidx(iP % L - iq % L, iP / L - iq / L, L); }   // K2[W, W-w | P, P-q]
            else { vv = win; kq = iq; }                                                                      // K2[W, w | P, q]
            if (inF(vv, lv.nK2f)) x += ldg(c.K2 + (posB(W, lv.nK2b) + nB * posF(vv, lv.nK2f) + nB * nF * (iP + NP * kq)));
        }
        val += x * FM::coef(f);
    }
    return val;
}
template <int KIND, int CH>
FDGA_HD C own_K3_term(const DevChain& V, const ColJob& job, const Grid& g, int W, int iP, int nu, int win) {
    typedef Forms<KIND, CH> FM;
    C val = zeroC();
    if (KIND == JOB_LK2 || KIND == JOB_LK2_LOC) return val;
    int v, w; job_freq_args<KIND, CH>(W, nu, win, v, w);
#pragma unroll
    for (int f = 0; f < FM::n; ++f) {
        const int form = FM::ch(f);
        C x = zeroC();
        for (int l = job.lev_first; l < job.n_nl2; ++l) {
            const DevLevel& lv = V.lev[l];
            if (inB(W, lv.nK2b) && inF(v, lv.nK2f) && inF(w, lv.nK2f) && inB(W, lv.nK3b) && inF(v, lv.nK3f) && inF(w, lv.nK3f)) {
                const int nB3 = 2 * lv.nK3b - 1, nF3 = 2 * lv.nK3f;
                x += ldg(lv.ch[form].K3 + (posB(W, lv.nK3b) + nB3 * (posF(v, lv.nK3f) + nF3 * (posF(w, lv.nK3f) + nF3 * iP))));
            }
        }
        val += x * FM::coef(f);
    }
    return val;
}
template <int KIND, int CH>
FDGA_HD C own_B_term(const DevChain& V, const ColJob& job, const Grid& g, int W, int iP, int ik, int nu) {
    typedef Forms<KIND, CH> FM;
    C val = zeroC();
    if (KIND == JOB_LK2 || KIND == JOB_LK2_LOC) return val;
#pragma unroll
    for (int f = 0; f < FM::n; ++f) {
        const int form = FM::ch(f);
        C x = zeroC();
        for (int l = job.lev_first; l < job.n_nl2; ++l) {
            const DevLevel& lv = V.lev[l];
            if (inB(W, lv.nK2b) && inF(nu, lv.nK2f)) {
                const int nB = 2 * lv.nK2b - 1, nF = 2 * lv.nK2f;
                x += ldg(lv.ch[form].K2 + (posB(W, lv.nK2b) + nB * posF(nu, lv.nK2f) + nB * nF * (iP + g.NP * ik)));
            }
        }
        val += x * FM::coef(f);
    }
    return val;
}
// k-independent part for one (W, nu, P): own A' and K3 pieces plus the pre-tabulated local/core levels (straightforward
// form, used by the host unit test; slab_own_kernel computes the same numbers with Rq staged in shared memory)
template <int KIND, int CH>
FDGA_HD C slab_own_entry(const DevChain& V, const ColJob& job, const Grid& g, const C* Rs, const C* T, int iW, int iP, int inu) {
    const int W = iW - (g.nK2b - 1), nu = inu - g.nK2f, nw = job.nw, nF2 = 2 * g.nK2f;
    C o = zeroC();
    for (int iq = 0; iq < g.NP; ++iq) for (int iw = 0; iw < nw; ++iw) {
        C t = own_A_term<KIND, CH>(V, job, g, W, iP, iw - job.Ninner, iq) + own_K3_term<KIND, CH>(V, job, g, W, iP, nu, iw - job.Ninner);
        if (T != nullptr) t += T[iw + nw * (inu + nF2 * iW)];
        o += t * Rs[slab_at(iw, iq, g.NP)];
    }
    return o;
}
// ---- TMA (bulk asynchronous copy) staging of one contiguous R slab into shared memory, completion on an mbarrier ------------
// The slab [q, w | W, P] is ONE contiguous run (nw * NP * 16 bytes, 16-byte aligned), so a single cp.async.bulk issued by one
// thread brings it in while the CTA builds its piece table; the tiles are then cut out of shared memory.
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ unsigned fdga_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fdga_mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(fdga_smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fdga_tma_load_1d(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(fdga_smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(fdga_smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(fdga_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fdga_mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(fdga_smem_u32(bar)), "r"(parity) : "memory");
}
#endif

// one CTA per active slab (W on the K2 mesh, P): OwnTab[nu | W, P] and Rtot[W, P]
template <int KIND, int CH>
__global__ void __launch_bounds__(256)
slab_own_kernel(const __grid_constant__ DevChain V, ColJob job, const int4* __restrict__ slabs, const C* __restrict__ R,
                const C* __restrict__ T, C* __restrict__ OwnTab, C* __restrict__ Rtot, Grid g, int PC, int use_tma) {
    extern __shared__ __align__(128) double sm_raw[];
    __shared__ __align__(8) unsigned long long slab_bar;
    __shared__ C s_sa;
    const int iW = slabs[blockIdx.x].x, iP = slabs[blockIdx.x].y;
    const int W = iW - (g.nK2b - 1), nw = job.nw, NP = g.NP, nF2 = 2 * g.nK2f, nB2 = 2 * g.nK2b - 1;
    C* Rsm = reinterpret_cast<C*>(sm_raw);                // [nw][NP] TMA-staged slab (absent when !use_tma)
    C* Rq = Rsm + (use_tma ? (size_t)nw * NP : 0);        // [nw], then part[PC * nw]
    const C* Rs = R + (size_t)nw * NP * slab_index(job.slabmap, posB(W, job.slabW_N), 2 * job.slabW_N - 1, iP);
#if defined(__CUDA_ARCH__)
    if (use_tma) {       // the slab is read twice (momentum sums, own-channel A' terms): one bulk copy, both passes from shared memory
        if (threadIdx.x == 0) fdga_mbar_init(&slab_bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) fdga_tma_load_1d(Rsm, Rs, (unsigned)((size_t)nw * NP * sizeof(C)), &slab_bar);
        fdga_mbar_wait(&slab_bar, 0);
        Rs = Rsm;
    }
#endif
    {   // Rq[iw] = sum_q Rs[q, iw]: one warp per inner frequency at a time, lanes over the (contiguous) inner momentum
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
        for (int iw = wid; iw < nw; iw += nwarp) {
            C s0 = zeroC(), s1 = zeroC();
            const C* row = Rs + (size_t)NP * iw;
            int iq = lane;
            for (; iq + 32 < NP; iq += 64) { s0 += row[iq]; s1 += row[iq + 32]; }
            if (iq < NP) s0 += row[iq];
            s0 += s1;
            for (int o = 16; o > 0; o >>= 1) { s0.x += __shfl_xor_sync(0xffffffffu, s0.x, o); s0.y += __shfl_xor_sync(0xffffffffu, s0.y, o); }
            if (lane == 0) Rq[iw] = s0;
        }
        __syncthreads();
    }
    C sa = zeroC();
    if (KIND == JOB_SDE_PP || KIND == JOB_SDE_PH)
#pragma unroll 4
        for (int t = threadIdx.x; t < nw * NP; t += blockDim.x) sa += own_A_term<KIND, CH>(V, job, g, W, iP, t / NP - job.Ninner, t % NP) * Rs[t];
    sa = block_reduce(sa);
    if (threadIdx.x == 0) s_sa = sa;
    __syncthreads();
    C rt = zeroC();
    for (int iw = threadIdx.x; iw < nw; iw += blockDim.x) rt += Rq[iw];
    rt = block_reduce(rt);
    if (threadIdx.x == 0) Rtot[iW + nB2 * iP] = rt;
    C* part = Rq + nw;                                    // [PC * nw]: PC values of nu at a time
    for (int nu0 = 0; nu0 < nF2; nu0 += PC) {
        const int nc = min(PC, nF2 - nu0);
        __syncthreads();
        for (int t = threadIdx.x; t < nc * nw; t += blockDim.x) {
            const int j = t / nw, iw = t - j * nw, inu = nu0 + j;
            C x = own_K3_term<KIND, CH>(V, job, g, W, iP, inu - g.nK2f, iw - job.Ninner);
            if (T != nullptr) x += ldg(T + iw + nw * (inu + nF2 * iW));
            part[t] = x * Rq[iw];
        }
        __syncthreads();
        for (int j = threadIdx.x; j < nc; j += blockDim.x) {
            C o0 = s_sa, o1 = zeroC();
            int iw = 0;
            for (; iw + 1 < nw; iw += 2) { o0 += part[j * nw + iw]; o1 += part[j * nw + iw + 1]; }
            if (iw < nw) o0 += part[j * nw + iw];
            OwnTab[nu0 + j + nF2 * (iW + nB2 * iP)] = o0 + o1;
        }
    }
}

// ---- cross-channel K1 pieces as a momentum convolution ---------------------------------------------------------
// After channel conversion the K1 argument of a cross channel r is
//     K1_r[ W'(nu, win) | P'(k, q) ],   W' = W0(W, nu) + sW * win,   P' = c(P) + sk * k + sq * q   (sk, sq = +-1),
// so its contribution  X[nu, k] = sum_{win, q} K1_r[W' | P'] R[win, q]  is a correlation in the frequency AND in the
// momentum.  With hat f[kappa] = sum_x f[x] exp(-2 pi i kappa.x / L):
//     X[nu, k] = 1/NP sum_kappa exp(+2 pi i kappa.k / L) Z[nu, kappa],
//     Z[nu, kappa] = sum_win hatK1_r[W' | sk kappa] * exp(2 pi i sk kappa.c / L) * hatR[win | -sq sk kappa],
// i.e. O(NP) instead of O(NP^2) work per (slab, nu, win).  One CTA per active slab (W, P) transforms its R slab in shared
// memory, accumulates Z over all (form, level, r) pieces and transforms back: ConvTab[k, nu | W, P].
struct ConvPiece { int W0, sW, cx, cy, sk, sq; };
template <int KIND, int CH>
FDGA_HD ConvPiece conv_piece(int form, int r, int W, int nu, int Px, int Py) {
    ConvPiece pc;
    int v_a, w_a, v_b, w_b, W0, v0, w0, W1, v1, w1;
    job_freq_args<KIND, CH>(W, nu, 0, v_a, w_a); job_freq_args<KIND, CH>(W, nu, 1, v_b, w_b);
    convert_freq(W, v_a, w_a, form, r, W0, v0, w0); convert_freq(W, v_b, w_b, form, r, W1, v1, w1);
    pc.W0 = W0; pc.sW = W1 - W0;
    Arg a; a.W = a.v = a.w = 0; a.Px = Px; a.Py = Py;
    job_mom_args<KIND, CH>(Px, Py, 0, 0, 0, 0, a.kx, a.ky, a.qx, a.qy);
    Arg b0 = convert(a, form, r);
    job_mom_args<KIND, CH>(Px, Py, 1, 0, 0, 0, a.kx, a.ky, a.qx, a.qy);
    Arg bk = convert(a, form, r);
    job_mom_args<KIND, CH>(Px, Py, 0, 0, 1, 0, a.kx, a.ky, a.qx, a.qy);
    Arg bq = convert(a, form, r);
    pc.cx = b0.Px; pc.cy = b0.Py; pc.sk = bk.Px - b0.Px; pc.sq = bq.Px - b0.Px;
    return pc;
}
// levels whose cross channels enter the column sum (same selection as column_thread)
template <int KIND>
FDGA_HD bool conv_level_on(const ColJob& job, int l) {
    if ((KIND == JOB_SDE_PP || KIND == JOB_SDE_PH) && (job.own_only || l == job.lev_first)) return false;
    return true;
}
template <int KIND>
FDGA_HD int conv_level_end(const ColJob& job) { return (KIND == JOB_LK2 || KIND == JOB_LK2_LOC) ? job.lev_first + 1 : job.n_nl2; }

// straightforward form of the same numbers (host unit test / A-B reference)
template <int KIND, int CH>
FDGA_HD C k1_cross_direct(const DevChain& V, const ColJob& job, const Grid& g, const C* Rs, int W, int iP, int ik, int nu) {
    typedef Forms<KIND, CH> FM;
    const int L = g.L, NP = g.NP, nw = job.nw;
    const int Px = iP % L, Py = iP / L, kx = ik % L, ky = ik / L;
    C acc = zeroC();
    for (int f = 0; f < FM::n; ++f) {
        const int form = FM::ch(f);
        for (int l = job.lev_first; l < conv_level_end<KIND>(job); ++l) {
            if (!conv_level_on<KIND>(job, l)) continue;
            const DevLevel& lv = V.lev[l];
            for (int r = 0; r < 3; ++r) {
                if (r == form) continue;
                const ConvPiece pc = conv_piece<KIND, CH>(form, r, W, nu, Px, Py);
                C part = zeroC();
                for (int iq = 0; iq < NP; ++iq) {
                    const int qx = iq % L, qy = iq / L;
                    const int iPp = foldidx(pc.cx + pc.sk * kx + pc.sq * qx, pc.cy + pc.sk * ky + pc.sq * qy, L);
                    for (int iw = 0; iw < nw; ++iw) {
                        const int Wc = pc.W0 + pc.sW * (iw - job.Ninner);
                        if (inB(Wc, lv.nK1)) part += ldg(lv.ch[r].K1 + (posB(Wc, lv.nK1) + (2 * lv.nK1 - 1) * iPp)) * Rs[slab_at(iw, iq, NP)];
                    }
                }
                acc += part * FM::coef(f);
            }
        }
    }
    return acc;
}

// hatK1[kappa + NP * iW] = sum_P K1[iW + nB1 * P] exp(-2 pi i kappa.P / L) for the three channels of one level
struct K1hOut { C* p[3]; };
__global__ void k1_dft_kernel(DevLevel lv, int L, int NP, const C* __restrict__ tw, K1hOut out) {
    extern __shared__ double sm_raw[];
    C* row = reinterpret_cast<C*>(sm_raw);                // [NP] K1[iW | .]
    C* stw = row + NP;                                    // [L]
    const int nB1 = 2 * lv.nK1 - 1, iW = blockIdx.x;
    const C* K1 = lv.ch[blockIdx.y].K1;
    for (int iP = threadIdx.x; iP < NP; iP += blockDim.x) row[iP] = K1[iW + nB1 * iP];
    for (int j = threadIdx.x; j < L; j += blockDim.x) stw[j] = tw[j];
    __syncthreads();
    for (int kap = threadIdx.x; kap < NP; kap += blockDim.x) {
        const int kx = kap % L, ky = kap / L;
        C s0 = zeroC(), s1 = zeroC();
        for (int y = 0; y < L; ++y) {
            C t = zeroC();
            for (int x = 0; x < L; ++x) t += row[x + L * y] * conjC(stw[(kx * x) % L]);
            if (y & 1) s1 += t * conjC(stw[(ky * y) % L]); else s0 += t * conjC(stw[(ky * y) % L]);
        }
        out.p[blockIdx.y][kap + (size_t)NP * iW] = s0 + s1;
    }
}

// per-CTA piece table: conv_piece at nu = 0 plus its slope in nu (W0 is linear in nu), level / channel / weight
struct ConvPieceS { int W0, dW0, sW, cx, cy, sk, sq, lev, r, nK1; double cf; };
#define FDGA_CONV_MAXP 24     // 2 forms x 6 levels x 2 cross channels
template <int KIND, int CH>
__global__ void __launch_bounds__(512)
slab_conv_kernel(const __grid_constant__ DevChain V, ColJob job, const int4* __restrict__ slabs, const C* __restrict__ R,
                 const C* __restrict__ tw, C* __restrict__ ConvTab, Grid g, int TW, int use_tma) {
    typedef Forms<KIND, CH> FM;
    extern __shared__ __align__(128) double sm_raw[];
    __shared__ __align__(8) unsigned long long slab_bar;
    __shared__ ConvPieceS pcs[FDGA_CONV_MAXP];
    __shared__ int s_npc, s_cnt;
    __shared__ unsigned char nuList[64];
    const int L = g.L, NP = g.NP, nF2 = 2 * g.nK2f, nB2 = 2 * g.nK2b - 1, nw = job.nw, Nin = job.Ninner;
    const int TWp = TW | 1;                               // odd row stride: conflict-free column gathers
    C* Rsm = reinterpret_cast<C*>(sm_raw);                // [nw][NP] the whole R slab (TMA staging; absent when !use_tma)
    C* A = Rsm + (use_tma ? (size_t)nw * NP : 0);         // [NP][TWp]
    C* B = A + (size_t)NP * TWp;                          // [NP][TWp]
    C* Z = B + (size_t)NP * TWp;                          // [nF2][NP]
    C* stw = Z + (size_t)nF2 * NP;                        // [L]  exp(+2 pi i j / L)
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int4 sl = slabs[blockIdx.x];
    const int iW = sl.x, iP = sl.y;
    const int W = iW - (g.nK2b - 1), Px = iP % L, Py = iP / L;
    const C* Rs = R + (size_t)nw * NP * slab_index(job.slabmap, posB(W, job.slabW_N), 2 * job.slabW_N - 1, iP);
#if defined(__CUDA_ARCH__)
    if (use_tma) {
        if (tid == 0) fdga_mbar_init(&slab_bar, 1);
        __syncthreads();
        if (tid == 0) fdga_tma_load_1d(Rsm, Rs, (unsigned)((size_t)nw * NP * sizeof(C)), &slab_bar);
    }
#endif
    for (int j = tid; j < L; j += nthr) stw[j] = tw[j];
    if (tid < FDGA_CONV_MAXP) {        // piece table, one thread per (form, level, cross channel) slot
        const int nl = conv_level_end<KIND>(job) - job.lev_first;
        const int rr = tid & 1, fl = tid >> 1, f = nl > 0 ? fl / nl : FM::n, l = job.lev_first + (nl > 0 ? fl % nl : 0);
        ConvPieceS q; q.cf = 0.0; q.lev = -1;
        if (f < FM::n && conv_level_on<KIND>(job, l)) {
            const int form = FM::ch(f);
            const int r = (form == 0) ? 1 + rr : (form == 1 ? 2 * rr : rr);      // the two channels != form
            const ConvPiece p0 = conv_piece<KIND, CH>(form, r, W, 0, Px, Py), p1 = conv_piece<KIND, CH>(form, r, W, 1, Px, Py);
            q.W0 = p0.W0; q.dW0 = p1.W0 - p0.W0; q.sW = p0.sW; q.cx = fold1(p0.cx, L); q.cy = fold1(p0.cy, L);
            q.sk = p0.sk; q.sq = p0.sq; q.lev = l; q.r = r; q.nK1 = V.lev[l].nK1; q.cf = FM::coef(f);
        }
        pcs[tid] = q;
    }
    if (tid == 32) {
        // nu values some column of this slab needs (bit mask built on the host; more than 64 values: all of them)
        int c = 0;
        const unsigned long long m = ((unsigned long long)(unsigned)sl.w << 32) | (unsigned)sl.z;
        if (nF2 <= 64) { for (int i = 0; i < nF2; ++i) if ((m >> i) & 1ULL) nuList[c++] = (unsigned char)i; }
        s_cnt = (nF2 <= 64) ? c : nF2;
        s_npc = min(FDGA_CONV_MAXP, 2 * FM::n * max(0, conv_level_end<KIND>(job) - job.lev_first));
    }
    __syncthreads();
    const int npc = s_npc, cnt = s_cnt;
    for (int o = tid; o < cnt * NP; o += nthr) Z[o] = zeroC();
    const C* Rsrc = Rs;
#if defined(__CUDA_ARCH__)
    if (use_tma) { fdga_mbar_wait(&slab_bar, 0); Rsrc = Rsm; }      // the slab has landed in shared memory
#endif

    for (int t0 = 0; t0 < nw; t0 += TW) {
        const int tn = min(TW, nw - t0);
        __syncthreads();
        for (int e = tid; e < tn * NP; e += nthr) { const int j = e / NP, q = e - j * NP; A[q * TWp + j] = Rsrc[slab_at(t0 + j, q, NP)]; }
        __syncthreads();
        for (int e = tid; e < tn * NP; e += nthr) {       // x axis
            const int kq = e / tn, j = e - kq * tn, qy = kq / L, kx = kq - qy * L;
            const C* a = A + (size_t)(L * qy) * TWp + j;
            C s0 = zeroC(), s1 = zeroC();
            int ph = 0;
            for (int qx = 0; qx < L; qx += 2) {
                s0 += a[qx * TWp] * conjC(stw[ph]); ph += kx; ph -= (ph >= L) ? L : 0;
                if (qx + 1 < L) { s1 += a[(qx + 1) * TWp] * conjC(stw[ph]); ph += kx; ph -= (ph >= L) ? L : 0; }
            }
            B[kq * TWp + j] = s0 + s1;
        }
        __syncthreads();
        for (int e = tid; e < tn * NP; e += nthr) {       // y axis
            const int kq = e / tn, j = e - kq * tn, ky = kq / L, kx = kq - ky * L;
            const C* b = B + (size_t)kx * TWp + j;
            C s0 = zeroC(), s1 = zeroC();
            int ph = 0;
            for (int qy = 0; qy < L; qy += 2) {
                s0 += b[(L * qy) * TWp] * conjC(stw[ph]); ph += ky; ph -= (ph >= L) ? L : 0;
                if (qy + 1 < L) { s1 += b[(L * (qy + 1)) * TWp] * conjC(stw[ph]); ph += ky; ph -= (ph >= L) ? L : 0; }
            }
            A[kq * TWp + j] = s0 + s1;
        }
        __syncthreads();
        for (int o = tid; o < cnt * NP; o += nthr) {
            const int in = o / NP, ko = o - in * NP, koy = ko / L, kox = ko - koy * L;
            const int inu = (nF2 <= 64) ? nuList[in] : in, nu = inu - g.nK2f;
            C z = zeroC();
            for (int p = 0; p < npc; ++p) {
                const ConvPieceS pc = pcs[p];
                if (pc.lev < 0) continue;
                const int kx = (pc.sk > 0 || kox == 0) ? kox : L - kox, ky = (pc.sk > 0 || koy == 0) ? koy : L - koy;
                const int rx = (pc.sq < 0 || kx == 0) ? kx : L - kx, ry = (pc.sq < 0 || ky == 0) ? ky : L - ky;
                const int W0 = pc.W0 + pc.dW0 * nu;
                Lin lW = {W0, pc.sW};
                int a = t0 - Nin, b = t0 + tn - 1 - Nin;
                clip_interval(lW, -(pc.nK1 - 1), pc.nK1 - 1, a, b);
                const C* Kh = V.lev[pc.lev].ch[pc.r].K1h + (kx + L * ky) + (size_t)NP * posB(W0, pc.nK1);
                const C* Ar = A + (size_t)(rx + L * ry) * TWp + (Nin - t0);
                const int stepK = NP * pc.sW;
                C s0 = zeroC(), s1 = zeroC();
#pragma unroll 4
                for (int win = a; win <= b; win += 2) {
                    const bool two = win + 1 <= b;
                    const C k0 = ldg(Kh + (ptrdiff_t)stepK * win);
                    const C k1 = two ? ldg(Kh + (ptrdiff_t)stepK * (win + 1)) : zeroC();
                    s0 += k0 * Ar[win];
                    s1 += k1 * Ar[two ? win + 1 : win];
                }
                const int ph = (kx * pc.cx + ky * pc.cy) % L;
                z += (s0 + s1) * stw[ph] * pc.cf;
            }
            Z[o] += z;
        }
    }
    __syncthreads();
    for (int e = tid; e < cnt * NP; e += nthr) {          // back transform, x axis
        const int in = e / NP, kq = e - in * NP, ky = kq / L, kx = kq - ky * L;
        const C* zr = Z + (size_t)in * NP + L * ky;
        C s0 = zeroC(), s1 = zeroC();
        int ph = 0;
        for (int x = 0; x < L; x += 2) {
            s0 += zr[x] * stw[ph]; ph += kx; ph -= (ph >= L) ? L : 0;
            if (x + 1 < L) { s1 += zr[x + 1] * stw[ph]; ph += kx; ph -= (ph >= L) ? L : 0; }
        }
        A[e] = s0 + s1;
    }
    __syncthreads();
    const double inv = 1.0 / (double)NP;
    for (int e = tid; e < cnt * NP; e += nthr) {          // y axis
        const int in = e / NP, kq = e - in * NP, ky = kq / L, kx = kq - ky * L;
        const int inu = (nF2 <= 64) ? nuList[in] : in;
        const C* ar = A + (size_t)in * NP + kx;
        C s0 = zeroC(), s1 = zeroC();
        int ph = 0;
        for (int y = 0; y < L; y += 2) {
            s0 += ar[L * y] * stw[ph]; ph += ky; ph -= (ph >= L) ? L : 0;
            if (y + 1 < L) { s1 += ar[L * (y + 1)] * stw[ph]; ph += ky; ph -= (ph >= L) ? L : 0; }
        }
        ConvTab[kq + (size_t)NP * (inu + nF2 * (iW + (size_t)nB2 * iP))] = (s0 + s1) * inv;
    }
}

// ---- the column kernel -----------------------------------------------------------------------------------------
// One CTA per GROUP of up to FDGA_WGROUP columns that share (P, k) and differ in W.  thread <-> (representative nu, inner
// momentum slot): the momentum conversions / table base offsets of a (k, q) pair are computed once and reused for every
// W of the group, and the table blocks fetched for one W are still in L1 when the next W gathers from them.  Lanes with
// consecutive nu gather neighbouring table entries; the R slab element is a broadcast inside a q-slot.
// per-thread part (host-callable for CPU unit tests): contribution of thread `tid` of `nthreads` to the columns of group `grp`
#ifndef FDGA_WGROUP
#define FDGA_WGROUP 1
#endif
template <int KIND, int CH>
FDGA_HD void column_thread(const DevChain& V, const ColJob& job, const ColDev& cols, const C* __restrict__ R,
                           const Grid& g, int grp, int tid, int nthreads, C* __restrict__ acc /* [FDGA_WGROUP] */) {
    typedef Forms<KIND, CH> FM;
    const int c0 = cols.grp_start[grp], ng = cols.grp_start[grp + 1] - c0;
    const int iP = cols.iP[c0], ik = cols.ik[c0];
    const int L = g.L, NP = g.NP;
    const int Px = iP % L, Py = iP / L, kx = ik % L, ky = ik / L;
    const int nw = job.nw;
    int maxrep = 1;
    for (int i = 0; i < ng; ++i) maxrep = max(maxrep, cols.start[c0 + i + 1] - cols.start[c0 + i]);
    int NVc = 1;
    while (NVc < maxrep) NVc <<= 1;                       // 1, 2, 4, 8 (<= FDGA_NV)
    const int n = tid & (NVc - 1);
    const int qs = tid / NVc, nqs = nthreads / NVc;
    int Wg[FDGA_WGROUP], nug[FDGA_WGROUP]; bool actg[FDGA_WGROUP]; bool any = false;
#pragma unroll
    for (int i = 0; i < FDGA_WGROUP; ++i) {
        acc[i] = zeroC(); Wg[i] = 0; nug[i] = 0; actg[i] = false;
        if (i < ng) {
            const int r0 = cols.start[c0 + i];
            Wg[i] = cols.iW[c0 + i] - (g.nK2b - 1);
            actg[i] = n < cols.start[c0 + i + 1] - r0;
            nug[i] = actg[i] ? cols.rep_inu[r0 + n] - g.nK2f : 0;
            any = any || actg[i];
        }
    }
    const size_t slab_stride = (size_t)nw * NP;

    // w is split in WS chunks so that small momentum meshes still fill the CTA
    int WS = 1;
    while (NP * WS < nqs && WS * 2 <= nw) WS *= 2;
    const int wchunk = (nw + WS - 1) / WS;
    const int l0 = job.lev_first;
    const int l_end = (KIND == JOB_LK2 || KIND == JOB_LK2_LOC) ? l0 + 1 : job.n_nl2;
    const bool withK1 = (KIND == JOB_LK2_LOC) || job.k1_direct;   // otherwise the K1 pieces come from slab_conv_kernel

    if (any)
    for (int item = qs; item < NP * WS; item += nqs) {
        const int iq = item / WS, ws = item - iq * WS;
        const int qx = iq % L, qy = iq / L;
        const int w_lo = ws * wchunk, w_hi = min(nw, w_lo + wchunk);
        // momentum arguments of the vertex for this (k, q)
        int akx, aky, aqx, aqy;
        job_mom_args<KIND, CH>(Px, Py, kx, ky, qx, qy, akx, aky, aqx, aqy);
        const C rs_chunk = zeroC();                             // cross-channel arguments are never constant in win

#pragma unroll
        for (int f = 0; f < FM::n; ++f) {
            const int form = FM::ch(f);
            const double cf = FM::coef(f);
            // cross-channel pieces (own-channel pieces and local / core levels live in slab_own_kernel + the epilogue):
            //   K2 jobs : every leading NL2 level of the left chain
            //   L_K2    : level l0 only (F0 = false)
            //   SDE     : ("as coded") every level l > l0 of the fused recursion SDE!(..., F.F0)
            for (int l = l0; l < l_end; ++l) {
                if ((KIND == JOB_SDE_PP || KIND == JOB_SDE_PH) && (job.own_only || l == l0)) continue;
                const DevLevel& lv = V.lev[l];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    if (r == form) continue;
                    const MomOff mo = mom_offsets(lv, form, r, L, NP, Px, Py, akx, aky, aqx, aqy);
#pragma unroll
                    for (int i = 0; i < FDGA_WGROUP; ++i) {
                        if (!actg[i]) continue;
                        // vertex frequency arguments are linear in win: v(win), w(win)
                        int v_a, w_a, v_b, w_b, W0, v0, w0, W1, v1, w1;
                        job_freq_args<KIND, CH>(Wg[i], nug[i], 0, v_a, w_a); job_freq_args<KIND, CH>(Wg[i], nug[i], 1, v_b, w_b);
                        convert_freq(Wg[i], v_a, w_a, form, r, W0, v0, w0); convert_freq(Wg[i], v_b, w_b, form, r, W1, v1, w1);
                        Lin lW = {W0, W1 - W0}, lv2 = {v0, v1 - v0}, lw2 = {w0, w1 - w0};
                        const C* Rq = R + slab_stride * slab_index(job.slabmap, posB(Wg[i], job.slabW_N), 2 * job.slabW_N - 1, iP) + iq;      // element win of this q: Rq[(win + Nin) * NP]
                        acc[i] += chan_lin_sum(lv, r, mo, lW, lv2, lw2, Rq, job.Ninner, NP, w_lo, w_hi, rs_chunk, withK1) * cf;
                    }
                }
            }
        }
    }
}

#ifndef FDGA_COL_MINB
#define FDGA_COL_MINB 6
#endif
template <int KIND, int CH>
__global__ void __launch_bounds__(128, FDGA_COL_MINB)
column_kernel(const __grid_constant__ DevChain V, ColJob job, ColDev cols, const C* __restrict__ R,
              const C* __restrict__ OwnTab, const C* __restrict__ Rtot, const C* __restrict__ ConvTab,
              C* __restrict__ repvals, Grid g) {
    const int grp = blockIdx.x;
    const int c0 = cols.grp_start[grp], ng = cols.grp_start[grp + 1] - c0;
    int maxrep = 1;
    for (int i = 0; i < ng; ++i) maxrep = max(maxrep, cols.start[c0 + i + 1] - cols.start[c0 + i]);
    int NVc = 1;
    while (NVc < maxrep) NVc <<= 1;
    C acc[FDGA_WGROUP];
    column_thread<KIND, CH>(V, job, cols, R, g, grp, threadIdx.x, blockDim.x, acc);
    // reduction over the momentum slots of each representative, one column of the group at a time
    __shared__ double redx[FDGA_WGROUP][128], redy[FDGA_WGROUP][128];
#pragma unroll
    for (int i = 0; i < FDGA_WGROUP; ++i) { redx[i][threadIdx.x] = acc[i].x; redy[i][threadIdx.x] = acc[i].y; }
    __syncthreads();
    for (int t = threadIdx.x; t < ng * NVc; t += blockDim.x) {
        const int i = t / NVc, nn = t - i * NVc, col = c0 + i;
        const int r0 = cols.start[col], nrep = cols.start[col + 1] - r0;
        if (nn >= nrep) continue;
        double x = 0.0, y = 0.0;
        for (int j = nn; j < (int)blockDim.x; j += NVc) { x += redx[i][j]; y += redy[i][j]; }
        C val = mkC(x, y);
        const int iW = cols.iW[col], iP = cols.iP[col], ik = cols.ik[col], inu = cols.rep_inu[r0 + nn];
        const int nF2 = 2 * g.nK2f, nB2 = 2 * g.nK2b - 1;
        if (ConvTab != nullptr)       // cross-channel K1 pieces (momentum convolution per slab)
            val += ConvTab[ik + (size_t)g.NP * (inu + nF2 * (iW + (size_t)nB2 * iP))];
        if (OwnTab != nullptr) {      // hoisted own-channel / local-level pieces
            val += OwnTab[inu + nF2 * (iW + nB2 * iP)]
                 + own_B_term<KIND, CH>(V, job, g, iW - (g.nK2b - 1), iP, ik, inu - g.nK2f) * Rtot[iW + nB2 * iP];
        }
        C s = mkC(job.scale_re, job.scale_im);
        repvals[cols.rep_cls[r0 + nn]] = val * s;
    }
}

}  // namespace fdga
