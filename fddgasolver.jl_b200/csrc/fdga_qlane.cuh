// fdga_qlane.cuh -- "q-lane" contraction kernel: the cross-channel K2 / K3 pieces of
//   BSE_K2! (fd / mfRG)   src/nonlocal_2/BSEa/BSEa_K2.jl:72-125
//   BSE_L_K2!             src/nonlocal_2/BSEa/BSEa_K2.jl:17-43
//   SDE_channel_L_pp!/ph! src/nonlocal_2/SDE.jl:16-33, 54-73
// with ONE WARP per class representative (W, nu, P, k) and the LANES running over the inner momentum q.
//
// Why: after _convert_channel every momentum argument of a cross-channel vertex is affine in q with slope 0 or +-1,
//     P' = P0 + sP q,  k' = k0 + sk q,  q' = q0 + sq q,
// so if the vertex tables are ALSO stored with a momentum index fastest (four "momentum layouts" of every K2 table, one of
// K3 and K1, see MomLay), the 32 lanes of a warp gather 32 consecutive table elements: every load of the hot loop is a fully
// coalesced 512-byte request (4 L1 wavefronts) instead of 32 scattered 16-byte gathers, and all Matsubara box logic is
// warp-uniform.  The right factor R is stored [q, w | W, P] (slab_at), so its loads coalesce the same way, and one R element
// feeds the K2(v'), K2(w') and K3 terms of a piece.  The sum over q ends in a warp-shuffle reduction.
#pragma once
#include "fdga_column.cuh"

namespace fdga {

// ---- momentum layouts ------------------------------------------------------------------------------------------------
// K2 tables, element (pW, pv, P, k):   m + NP * (pW + nB * (pv + nF * blk))
//   ML_P : m = P,  blk = k          (k' constant along q)
//   ML_K : m = k,  blk = P          (P' constant along q)
//   ML_S : m = k,  blk = P + k      (P' + k' constant along q: opposite slopes)
//   ML_D : m = k,  blk = P - k      (P' - k' constant along q: equal slopes)
// K3 tables, element (pW, pv, pw, P):  P + NP * (pW + nB3 * (pv + nF3 * pw));   K1: P + NP * pW
enum { ML_P = 0, ML_K = 1, ML_S = 2, ML_D = 3, ML_K3 = 4, ML_K1 = 5, ML_COUNT = 6 };

// source momenta (iP, ik) of the element (m, blk) of K2 layout `lay`
FDGA_HD void mom_layout_source(int lay, int m, int blk, int L, int& iP, int& ik) {
    const int mx = m % L, my = m / L, bx = blk % L, by = blk / L;
    if (lay == ML_P) { iP = m; ik = blk; }
    else if (lay == ML_K) { ik = m; iP = blk; }
    else if (lay == ML_S) { ik = m; iP = fold1(bx - mx, L) + L * fold1(by - my, L); }      // P = S - k
    else { ik = m; iP = fold1(bx + mx, L) + L * fold1(by + my, L); }                       // P = D + k
}
FDGA_HD int mom_layout_of(int sP, int s2) { return s2 == 0 ? ML_P : (sP == 0 ? ML_K : (sP == s2 ? ML_D : ML_S)); }

struct MomOut { C* p[3][ML_COUNT]; };       // [channel][layout]; null = not requested
// one thread per output element of one (channel, layout) target: blockIdx.y = channel * ML_COUNT + layout
__global__ void mom_layout_kernel(DevLevel lv, int L, int NP, MomOut out) {
    const int ch = blockIdx.y / ML_COUNT, lay = blockIdx.y % ML_COUNT;
    C* dst = out.p[ch][lay];
    if (dst == nullptr) return;
    const DevChan& c = lv.ch[ch];
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (lay == ML_K1) {
        const int nB1 = 2 * lv.nK1 - 1;
        if (i >= (long long)nB1 * NP) return;
        const int iP = (int)(i % NP), pW = (int)(i / NP);
        dst[i] = c.K1[pW + (size_t)nB1 * iP];
    } else if (lay == ML_K3) {
        const long long n3 = (long long)(2 * lv.nK3b - 1) * (2 * lv.nK3f) * (2 * lv.nK3f);
        if (i >= n3 * NP) return;
        const int iP = (int)(i % NP); const long long row = i / NP;
        dst[i] = c.K3[row + (size_t)n3 * iP];
    } else {
        const long long n2 = (long long)(2 * lv.nK2b - 1) * (2 * lv.nK2f);
        if (i >= n2 * NP * NP) return;
        const int m = (int)(i % NP); const long long t = i / NP; const long long row = t % n2; const int blk = (int)(t / n2);
        int iP, ik; mom_layout_source(lay, m, blk, L, iP, ik);
        dst[i] = c.K2[row + (size_t)n2 * (iP + (size_t)NP * ik)];
    }
}

// ---- per-piece setup (warp-uniform) ----------------------------------------------------------------------------------
struct MomSel { int lay, bx, by, s, blk; };     // lane momentum = fold(b + s * q), block index blk, layout lay
// (P', X') = (P0 + sP q, X0 + sX q): pick the layout in which the element run over q is contiguous
FDGA_HD MomSel mom_select(int P0x, int P0y, int sP, int X0x, int X0y, int sX, int L) {
    MomSel m; m.lay = mom_layout_of(sP, sX);
    const int px = modL(P0x, L), py = modL(P0y, L), xx = modL(X0x, L), xy = modL(X0y, L);
    if (m.lay == ML_P) { m.bx = px; m.by = py; m.s = sP; m.blk = xx + L * xy; }
    else {
        m.bx = xx; m.by = xy; m.s = sX;
        if (m.lay == ML_K) m.blk = px + L * py;
        else if (m.lay == ML_S) m.blk = modL(px + xx, L) + L * modL(py + xy, L);
        else m.blk = modL(px - xx, L) + L * modL(py - xy, L);
    }
    return m;
}
FDGA_HD int mom_lane(const MomSel& m, int qx, int qy, int L) {      // b in [0, L), s q in (-L, L)
    int x = m.bx + m.s * qx, y = m.by + m.s * qy;
    x += (x < 0) ? L : 0; x -= (x >= L) ? L : 0;
    y += (y < 0) ? L : 0; y -= (y >= L) ? L : 0;
    return x + L * y;
}

// layouts a job kind reads, per table channel r (bit = layout id): host side, to build only what is needed
template <int KIND, int CH>
inline void qlane_needed_layouts(unsigned mask[3], bool withK1) {
    typedef Forms<KIND, CH> FM;
    for (int f = 0; f < FM::n; ++f) for (int r = 0; r < 3; ++r) {
        const int form = FM::ch(f);
        if (r == form) continue;
        Arg a; a.W = a.v = a.w = 0; a.Px = 5; a.Py = 3;
        job_mom_args<KIND, CH>(5, 3, 2, 1, 0, 0, a.kx, a.ky, a.qx, a.qy);
        const Arg b0 = convert(a, form, r);
        job_mom_args<KIND, CH>(5, 3, 2, 1, 1, 0, a.kx, a.ky, a.qx, a.qy);
        const Arg b1 = convert(a, form, r);
        const int sP = b1.Px - b0.Px, sk = b1.kx - b0.kx, sq = b1.qx - b0.qx;
        mask[r] |= 1u << mom_layout_of(sP, sk);
        mask[r] |= 1u << mom_layout_of(sP, sq);
        mask[r] |= 1u << ML_K3;
        if (withK1) mask[r] |= 1u << ML_K1;
    }
}

// ---- the contraction: contribution of one lane to the representative (iW, inu, iP, ik) -----------------------------------
// Host-callable (tests/host_column_test.cu emulates the warp by summing the 32 lanes).
template <int KIND, int CH>
FDGA_HD C qlane_lane(const DevChain& V, const ColJob& job, const Grid& g, const C* __restrict__ R,
                     int iW, int inu, int iP, int ik, int lane) {
    typedef Forms<KIND, CH> FM;
    const int L = g.L, NP = g.NP, nw = job.nw, Nin = job.Ninner;
    const int W = iW - (g.nK2b - 1), nu = inu - g.nK2f;
    const int Px = iP % L, Py = iP / L, kx = ik % L, ky = ik / L;
    const C* __restrict__ Rs = R + (size_t)nw * NP * (posB(W, job.slabW_N) + (size_t)(2 * job.slabW_N - 1) * iP);
    const int l0 = job.lev_first;
    const int l_end = (KIND == JOB_LK2 || KIND == JOB_LK2_LOC) ? l0 + 1 : job.n_nl2;
    const bool withK1 = job.k1_direct != 0;
    C acc = zeroC();
    for (int q0 = lane; q0 < NP; q0 += 64) {          // two momentum slots per pass: q0 and q0 + 32
        const bool has1 = q0 + 32 < NP;
        const int q1 = has1 ? q0 + 32 : q0;
        const int q0x = q0 % L, q0y = q0 / L, q1x = q1 % L, q1y = q1 / L;
#pragma unroll
        for (int f = 0; f < FM::n; ++f) {
            const int form = FM::ch(f);
            const double cf = FM::coef(f);
            for (int l = l0; l < l_end; ++l) {
                if ((KIND == JOB_SDE_PP || KIND == JOB_SDE_PH) && (job.own_only || l == l0)) continue;
                const DevLevel& lv = V.lev[l];
                const int nB = 2 * lv.nK2b - 1, nF = 2 * lv.nK2f, nB3 = 2 * lv.nK3b - 1, nF3 = 2 * lv.nK3f;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    if (r == form) continue;
                    const DevChan& c = lv.ch[r];
                    // frequency arguments, linear in win
                    int v_a, w_a, v_b, w_b, W0, v0, w0, W1, v1, w1;
                    job_freq_args<KIND, CH>(W, nu, 0, v_a, w_a); job_freq_args<KIND, CH>(W, nu, 1, v_b, w_b);
                    convert_freq(W, v_a, w_a, form, r, W0, v0, w0); convert_freq(W, v_b, w_b, form, r, W1, v1, w1);
                    const Lin lW = {W0, W1 - W0}, lv2 = {v0, v1 - v0}, lw2 = {w0, w1 - w0};
                    // momentum arguments, affine in q
                    Arg a; a.W = a.v = a.w = 0; a.Px = Px; a.Py = Py;
                    job_mom_args<KIND, CH>(Px, Py, kx, ky, 0, 0, a.kx, a.ky, a.qx, a.qy);
                    const Arg b0 = convert(a, form, r);
                    job_mom_args<KIND, CH>(Px, Py, kx, ky, 1, 0, a.kx, a.ky, a.qx, a.qy);
                    const Arg b1 = convert(a, form, r);
                    const int sP = b1.Px - b0.Px, sk = b1.kx - b0.kx, sq = b1.qx - b0.qx;
                    const MomSel mA = mom_select(b0.Px, b0.Py, sP, b0.kx, b0.ky, sk, L);
                    const MomSel mB = mom_select(b0.Px, b0.Py, sP, b0.qx, b0.qy, sq, L);
                    MomSel m3; m3.lay = ML_P; m3.bx = modL(b0.Px, L); m3.by = modL(b0.Py, L); m3.s = sP; m3.blk = 0;
                    const C* __restrict__ tA = c.K2m[mA.lay] + (size_t)NP * nB * nF * mA.blk;
                    const C* __restrict__ tB = c.K2m[mB.lay] + (size_t)NP * nB * nF * mB.blk;
                    const C* __restrict__ t3 = c.K3m;
                    const int oA0 = mom_lane(mA, q0x, q0y, L), oA1 = mom_lane(mA, q1x, q1y, L);
                    const int oB0 = mom_lane(mB, q0x, q0y, L), oB1 = mom_lane(mB, q1x, q1y, L);
                    const int o30 = mom_lane(m3, q0x, q0y, L), o31 = mom_lane(m3, q1x, q1y, L);
                    C p0 = zeroC(), p1 = zeroC();
                    if (withK1) {       // cross-channel K1 term inside the kernel (otherwise: slab_conv_kernel)
                        int a1 = -Nin, b1w = nw - 1 - Nin; clip_interval(lW, -(lv.nK1 - 1), lv.nK1 - 1, a1, b1w);
                        const C* __restrict__ t1 = c.K1m;
                        for (int win = a1; win <= b1w; ++win) {
                            const size_t row = (size_t)NP * posB(lW.x0 + lW.s * win, lv.nK1);
                            const size_t rr = (size_t)NP * (win + Nin);
                            p0 += ldg(t1 + row + o30) * Rs[rr + q0];
                            p1 += ldg(t1 + row + o31) * Rs[rr + q1];
                        }
                    }
                    int wa = -Nin, wb = nw - 1 - Nin; clip_interval(lW, -(lv.nK2b - 1), lv.nK2b - 1, wa, wb);
                    for (int win = wa; win <= wb; ++win) {
                        const int Wc = lW.x0 + lW.s * win, vc = lv2.x0 + lv2.s * win, wc = lw2.x0 + lw2.s * win;
                        const bool inA = inF(vc, lv.nK2f), inBt = inF(wc, lv.nK2f);
                        const bool in3 = inA && inBt && inB(Wc, lv.nK3b) && inF(vc, lv.nK3f) && inF(wc, lv.nK3f);
                        if (!(inA || inBt)) continue;
                        const int pW = posB(Wc, lv.nK2b);
                        const size_t rr = (size_t)NP * (win + Nin);
                        const C r0 = Rs[rr + q0], r1 = Rs[rr + q1];
                        C t0 = zeroC(), t1 = zeroC();
                        if (inA) {
                            const size_t row = (size_t)NP * (pW + nB * posF(vc, lv.nK2f));
                            t0 += ldg(tA + row + oA0); t1 += ldg(tA + row + oA1);
                        }
                        if (inBt) {
                            const size_t row = (size_t)NP * (pW + nB * posF(wc, lv.nK2f));
                            t0 += ldg(tB + row + oB0); t1 += ldg(tB + row + oB1);
                        }
                        if (in3) {
                            const size_t row = (size_t)NP * (posB(Wc, lv.nK3b) + nB3 * (posF(vc, lv.nK3f) + nF3 * posF(wc, lv.nK3f)));
                            t0 += ldg(t3 + row + o30); t1 += ldg(t3 + row + o31);
                        }
                        p0 += t0 * r0; p1 += t1 * r1;
                    }
                    if (!has1) p1 = zeroC();
                    acc += (p0 + p1) * cf;
                }
            }
        }
    }
    return acc;
}

// class representatives of this rank, sorted by slab: (iW | inu << 16, iP, ik, class slot)
struct RepDev { int nrep; const int4* rep; };

#ifndef FDGA_QL_WARPS
#define FDGA_QL_WARPS 4
#endif
template <int KIND, int CH>
__global__ void __launch_bounds__(32 * FDGA_QL_WARPS)
qlane_kernel(const __grid_constant__ DevChain V, ColJob job, RepDev reps, const C* __restrict__ R,
             const C* __restrict__ OwnTab, const C* __restrict__ Rtot, const C* __restrict__ ConvTab,
             C* __restrict__ repvals, Grid g) {
    const int w = blockIdx.x * FDGA_QL_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (w >= reps.nrep) return;
    const int4 rp = reps.rep[w];
    const int iW = rp.x & 0xffff, inu = rp.x >> 16, iP = rp.y, ik = rp.z;
    C acc = qlane_lane<KIND, CH>(V, job, g, R, iW, inu, iP, ik, lane);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o); }
    if (lane == 0) {
        const int nF2 = 2 * g.nK2f, nB2 = 2 * g.nK2b - 1;
        C val = acc;
        if (ConvTab != nullptr)       // cross-channel K1 pieces (momentum convolution per slab)
            val += ConvTab[ik + (size_t)g.NP * (inu + nF2 * (iW + (size_t)nB2 * iP))];
        if (OwnTab != nullptr)        // hoisted own-channel / local-level pieces
            val += OwnTab[inu + nF2 * (iW + nB2 * iP)]
                 + own_B_term<KIND, CH>(V, job, g, iW - (g.nK2b - 1), iP, ik, inu - g.nK2f) * Rtot[iW + nB2 * iP];
        repvals[rp.w] = val * mkC(job.scale_re, job.scale_im);
    }
}

}  // namespace fdga
