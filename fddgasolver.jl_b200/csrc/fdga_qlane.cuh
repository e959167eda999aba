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

#ifndef FDGA_QL_UNROLL
#define FDGA_QL_UNROLL 1
#endif

namespace fdga {

constexpr int QL_UNROLL = FDGA_QL_UNROLL;      // unroll factor of the win loop of qlane_lane

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
// (P', X') = (P0 + sP q, X0 + sX q), components of P0, X0 in (-3L, 3L) (sums of at most three mesh momenta): pick the layout in
// which the element run over q is contiguous.  No integer division: fold1 is branch-free.
FDGA_HD MomSel mom_select(int P0x, int P0y, int sP, int X0x, int X0y, int sX, int L) {
    MomSel m; m.lay = mom_layout_of(sP, sX);
    const int px = fold1(P0x, L), py = fold1(P0y, L), xx = fold1(X0x, L), xy = fold1(X0y, L);
    if (m.lay == ML_P) { m.bx = px; m.by = py; m.s = sP; m.blk = xx + L * xy; }
    else {
        m.bx = xx; m.by = xy; m.s = sX;
        if (m.lay == ML_K) m.blk = px + L * py;
        else if (m.lay == ML_S) m.blk = fold1(px + xx, L) + L * fold1(py + xy, L);
        else m.blk = fold1(px - xx, L) + L * fold1(py - xy, L);
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

// ---- the contraction ---------------------------------------------------------------------------------------------------
// A PIECE is one (spin form f, chain level l, cross channel r) term of the left factor of one representative.  Its setup is
// warp-uniform; the box logic of the inner frequency is evaluated ONCE per (piece, win) into a 16-byte ENTRY (element offsets
// of the three table rows and of the R row; terms outside their Matsubara box point at the NP zeros that pad every table copy),
// one entry per lane, kept in shared memory.  The hot loop then only reads an entry (broadcast), issues its eight coalesced
// 512-byte loads (two momentum slots) and does 2 complex adds + 1 complex multiply-add per slot.
struct QPiece {
    const C* tA; const C* tB; const C* t3; const C* t1;     // table copies at this piece's block (lane offset not included)
    int zA, zB, z3;                                         // element offset of the zero padding from tA / tB / t3
    MomSel mA, mB, m3;
    Lin lW, lv2, lw2;
    int wa, wb;                                             // win range with W' inside the K2 bosonic box
};
// levels whose cross channels enter (same selection as column_thread / conv_level_on)
template <int KIND>
FDGA_HD bool qlane_level_on(const ColJob& job, int l) {
    return !((KIND == JOB_SDE_PP || KIND == JOB_SDE_PH) && (job.own_only || l == job.lev_first));
}
template <int KIND>
FDGA_HD int qlane_level_end(const ColJob& job) { return (KIND == JOB_LK2 || KIND == JOB_LK2_LOC) ? job.lev_first + 1 : job.n_nl2; }

template <int KIND, int CH>
FDGA_HD void qlane_piece(const DevLevel& lv, const ColJob& job, const Grid& g, int form, int r, int W, int nu,
                         int Px, int Py, int kx, int ky, QPiece& pc) {
    const int L = g.L, NP = g.NP, nw = job.nw, Nin = job.Ninner;
    const DevChan& c = lv.ch[r];
    const int nB = 2 * lv.nK2b - 1, nF = 2 * lv.nK2f, nB3 = 2 * lv.nK3b - 1, nF3 = 2 * lv.nK3f;
    // frequency arguments, linear in win
    int v_a, w_a, v_b, w_b, W0, v0, w0, W1, v1, w1;
    job_freq_args<KIND, CH>(W, nu, 0, v_a, w_a); job_freq_args<KIND, CH>(W, nu, 1, v_b, w_b);
    convert_freq(W, v_a, w_a, form, r, W0, v0, w0); convert_freq(W, v_b, w_b, form, r, W1, v1, w1);
    pc.lW.x0 = W0; pc.lW.s = W1 - W0; pc.lv2.x0 = v0; pc.lv2.s = v1 - v0; pc.lw2.x0 = w0; pc.lw2.s = w1 - w0;
    // momentum arguments, affine in q
    Arg a; a.W = a.v = a.w = 0; a.Px = Px; a.Py = Py;
    job_mom_args<KIND, CH>(Px, Py, kx, ky, 0, 0, a.kx, a.ky, a.qx, a.qy);
    const Arg b0 = convert(a, form, r);
    job_mom_args<KIND, CH>(Px, Py, kx, ky, 1, 0, a.kx, a.ky, a.qx, a.qy);
    const Arg b1 = convert(a, form, r);
    const int sP = b1.Px - b0.Px, sk = b1.kx - b0.kx, sq = b1.qx - b0.qx;
    pc.mA = mom_select(b0.Px, b0.Py, sP, b0.kx, b0.ky, sk, L);
    pc.mB = mom_select(b0.Px, b0.Py, sP, b0.qx, b0.qy, sq, L);
    pc.m3.lay = ML_P; pc.m3.bx = fold1(b0.Px, L); pc.m3.by = fold1(b0.Py, L); pc.m3.s = sP; pc.m3.blk = 0;
    const int blk2 = NP * nB * nF, len2 = blk2 * NP;
    pc.tA = c.K2m[pc.mA.lay] + (size_t)blk2 * pc.mA.blk; pc.zA = len2 - blk2 * pc.mA.blk;
    pc.tB = c.K2m[pc.mB.lay] + (size_t)blk2 * pc.mB.blk; pc.zB = len2 - blk2 * pc.mB.blk;
    pc.t3 = c.K3m; pc.z3 = NP * nB3 * nF3 * nF3;
    pc.t1 = c.K1m;
    pc.wa = -Nin; pc.wb = nw - 1 - Nin; clip_interval(pc.lW, -(lv.nK2b - 1), lv.nK2b - 1, pc.wa, pc.wb);
}
// entry of inner frequency win: element offsets (rowA, rowB, row3, R row)
FDGA_HD uint4 qlane_entry(const QPiece& pc, const DevLevel& lv, int NP, int Nin, int win) {
    const int nB = 2 * lv.nK2b - 1, nB3 = 2 * lv.nK3b - 1, nF3 = 2 * lv.nK3f;
    const int Wc = pc.lW.x0 + pc.lW.s * win, vc = pc.lv2.x0 + pc.lv2.s * win, wc = pc.lw2.x0 + pc.lw2.s * win;
    const bool inA = inF(vc, lv.nK2f), inBt = inF(wc, lv.nK2f);
    const bool in3 = inA && inBt && inB(Wc, lv.nK3b) && inF(vc, lv.nK3f) && inF(wc, lv.nK3f);
    const int pW = posB(Wc, lv.nK2b);
    uint4 e;
    e.x = inA ? NP * (pW + nB * posF(vc, lv.nK2f)) : pc.zA;
    e.y = inBt ? NP * (pW + nB * posF(wc, lv.nK2f)) : pc.zB;
    e.z = in3 ? NP * (posB(Wc, lv.nK3b) + nB3 * (posF(vc, lv.nK3f) + nF3 * posF(wc, lv.nK3f))) : pc.z3;
    e.w = NP * (win + Nin);
    return e;
}
// one lane's share of a piece: momentum slots q0 = (q0x, q0y) and q1 (q1 == q0 with has1 == false: second slot idle)
FDGA_HD C qlane_consume(const QPiece& pc, const uint4* __restrict__ ent, int n, const C* __restrict__ r0, const C* __restrict__ r1,
                        int q0x, int q0y, int q1x, int q1y, bool has1, int L) {
    const C* __restrict__ a0 = pc.tA + mom_lane(pc.mA, q0x, q0y, L); const C* __restrict__ a1 = pc.tA + mom_lane(pc.mA, q1x, q1y, L);
    const C* __restrict__ b0 = pc.tB + mom_lane(pc.mB, q0x, q0y, L); const C* __restrict__ b1 = pc.tB + mom_lane(pc.mB, q1x, q1y, L);
    const C* __restrict__ c0 = pc.t3 + mom_lane(pc.m3, q0x, q0y, L); const C* __restrict__ c1 = pc.t3 + mom_lane(pc.m3, q1x, q1y, L);
    C p0 = zeroC(), p1 = zeroC();
#pragma unroll QL_UNROLL
    for (int i = 0; i < n; ++i) {
        const uint4 e = ent[i];
        const C va0 = ldg(a0 + e.x), va1 = ldg(a1 + e.x), vb0 = ldg(b0 + e.y), vb1 = ldg(b1 + e.y), vc0 = ldg(c0 + e.z), vc1 = ldg(c1 + e.z);
        const C x0 = r0[e.w], x1 = r1[e.w];
        p0 += ((va0 + vb0) + vc0) * x0; p1 += ((va1 + vb1) + vc1) * x1;
    }
    if (!has1) p1 = zeroC();
    return p0 + p1;
}
#ifndef FDGA_QL_DEPTH
#define FDGA_QL_DEPTH 0        // stages of the cp.async ring of qlane_consume_async (0: plain loads, qlane_consume)
#endif
#if defined(__CUDA_ARCH__)
// The same sum with the loads software-pipelined through shared memory: every lane copies its own eight 16-byte elements of
// entry i + DEPTH - 1 with cp.async (L1-allocating, so the R rows and zero rows shared with neighbouring warps still hit) into a
// per-warp ring of DEPTH stages while it consumes entry i.  Loads in flight do not hold registers, and there is no barrier: a
// lane only ever reads back what it copied itself.  Stage layout: [8 loads][32 lanes] x 16 B = 4 KB.
template <int DEPTH>
__device__ __forceinline__ C qlane_consume_async(const QPiece& pc, const uint4* __restrict__ ent, int n, const C* __restrict__ r0, const C* __restrict__ r1,
                                                 int q0x, int q0y, int q1x, int q1y, bool has1, int L, C* ring_lane) {
    const C* a0 = pc.tA + mom_lane(pc.mA, q0x, q0y, L); const C* a1 = pc.tA + mom_lane(pc.mA, q1x, q1y, L);
    const C* b0 = pc.tB + mom_lane(pc.mB, q0x, q0y, L); const C* b1 = pc.tB + mom_lane(pc.mB, q1x, q1y, L);
    const C* c0 = pc.t3 + mom_lane(pc.m3, q0x, q0y, L); const C* c1 = pc.t3 + mom_lane(pc.m3, q1x, q1y, L);
    const unsigned sb = (unsigned)__cvta_generic_to_shared(ring_lane);
    auto cp16 = [](unsigned dst, const C* src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory"); };
    auto issue = [&](int i) {
        const uint4 e = ent[i];
        const unsigned d = sb + (unsigned)(i % DEPTH) * 4096u;
        cp16(d, a0 + e.x); cp16(d + 512u, a1 + e.x); cp16(d + 1024u, b0 + e.y); cp16(d + 1536u, b1 + e.y);
        cp16(d + 2048u, c0 + e.z); cp16(d + 2560u, c1 + e.z); cp16(d + 3072u, r0 + e.w); cp16(d + 3584u, r1 + e.w);
    };
#pragma unroll
    for (int i = 0; i < DEPTH - 1; ++i) { if (i < n) issue(i); asm volatile("cp.async.commit_group;" ::: "memory"); }
    C p0 = zeroC(), p1 = zeroC();
    for (int i = 0; i < n; ++i) {
        asm volatile("cp.async.wait_group %0;" :: "n"(DEPTH - 2) : "memory");      // entry i has landed
        const C* st = ring_lane + (size_t)(i % DEPTH) * 256;                         // 4096 B = 256 elements per stage
        const C va0 = st[0], va1 = st[32], vb0 = st[64], vb1 = st[96], vc0 = st[128], vc1 = st[160], x0 = st[192], x1 = st[224];
        if (i + DEPTH - 1 < n) issue(i + DEPTH - 1);                                 // into the stage consumed one iteration ago
        asm volatile("cp.async.commit_group;" ::: "memory");
        p0 += ((va0 + vb0) + vc0) * x0; p1 += ((va1 + vb1) + vc1) * x1;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (!has1) p1 = zeroC();
    return p0 + p1;
}
#endif
// cross-channel K1 term summed inside the kernel (FDGA_OPT_DIRECT_K1, or slabs too large for slab_conv_kernel)
FDGA_HD C qlane_k1_direct(const QPiece& pc, const DevLevel& lv, const C* __restrict__ r0, const C* __restrict__ r1, int NP, int nw, int Nin,
                          int q0x, int q0y, int q1x, int q1y, bool has1, int L) {
    const int o0 = mom_lane(pc.m3, q0x, q0y, L), o1 = mom_lane(pc.m3, q1x, q1y, L);
    int a1 = -Nin, b1 = nw - 1 - Nin; clip_interval(pc.lW, -(lv.nK1 - 1), lv.nK1 - 1, a1, b1);
    C p0 = zeroC(), p1 = zeroC();
    for (int win = a1; win <= b1; ++win) {
        const size_t row = (size_t)NP * posB(pc.lW.x0 + pc.lW.s * win, lv.nK1), rr = (size_t)NP * (win + Nin);
        p0 += ldg(pc.t1 + row + o0) * r0[rr];
        p1 += ldg(pc.t1 + row + o1) * r1[rr];
    }
    if (!has1) p1 = zeroC();
    return p0 + p1;
}
#define FDGA_QL_ENT 32     // entries (inner frequencies) staged per pass: one per lane

// The warp's work on one representative.  SYNC = true: device, entries staged through `ent` (32 slots of shared memory per warp)
// with __syncwarp; SYNC = false: the host restatement of ONE lane with a private entry array (tests/host_column_test.cu sums it
// over the 32 lanes).  Same building blocks either way.
template <int KIND, int CH, bool SYNC>
FDGA_HD C qlane_rep(const DevChain& V, const ColJob& job, const Grid& g, const C* __restrict__ R, uint4* ent, C* ring_lane,
                    int iW, int inu, int iP, int Px, int Py, int kx, int ky, int lane) {
    typedef Forms<KIND, CH> FM;
    const int L = g.L, NP = g.NP, nw = job.nw, Nin = job.Ninner;
    const int W = iW - (g.nK2b - 1), nu = inu - g.nK2f;
    const C* __restrict__ Rs = R + (size_t)nw * NP * slab_index(job.slabmap, posB(W, job.slabW_N), 2 * job.slabW_N - 1, iP);
    const int l0 = job.lev_first, l_end = qlane_level_end<KIND>(job);
    C acc = zeroC();
    for (int q0 = lane; q0 < NP || SYNC; q0 += 64) {       // SYNC: every lane runs the loop (warp-wide barriers inside), idle lanes masked
        if (SYNC && (q0 - lane) >= NP) break;
        const bool has0 = q0 < NP, has1 = q0 + 32 < NP;
        const int qa = has0 ? q0 : 0, qb = has1 ? q0 + 32 : qa;
        const int q0y = qa / L, q0x = qa - q0y * L, q1y = qb / L, q1x = qb - q1y * L;
        const C* __restrict__ r0 = Rs + qa; const C* __restrict__ r1 = Rs + qb;
#pragma unroll
        for (int f = 0; f < FM::n; ++f) {
            const int form = FM::ch(f);
            C pf = zeroC();
            for (int l = l0; l < l_end; ++l) {
                if (!qlane_level_on<KIND>(job, l)) continue;
                const DevLevel& lv = V.lev[l];
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int r = (form == 0) ? 1 + rr : (form == 1 ? 2 * rr : rr);      // the two channels != form
                    QPiece pc;
                    qlane_piece<KIND, CH>(lv, job, g, form, r, W, nu, Px, Py, kx, ky, pc);
                    for (int e0 = pc.wa; e0 <= pc.wb; e0 += FDGA_QL_ENT) {
                        const int n = min(FDGA_QL_ENT, pc.wb - e0 + 1);
#if defined(__CUDA_ARCH__)
                        if (SYNC) {
                            __syncwarp();
                            if (lane < n) ent[lane] = qlane_entry(pc, lv, NP, Nin, e0 + lane);
                            __syncwarp();
                        } else
#endif
                        for (int i = 0; i < n; ++i) ent[i] = qlane_entry(pc, lv, NP, Nin, e0 + i);
#if defined(__CUDA_ARCH__) && FDGA_QL_DEPTH >= 2
                        const C part = SYNC ? qlane_consume_async<FDGA_QL_DEPTH>(pc, ent, n, r0, r1, q0x, q0y, q1x, q1y, has1, L, ring_lane)
                                            : qlane_consume(pc, ent, n, r0, r1, q0x, q0y, q1x, q1y, has1, L);
#else
                        const C part = qlane_consume(pc, ent, n, r0, r1, q0x, q0y, q1x, q1y, has1, L);
#endif
                        if (has0) pf += part;
                    }
                    if (job.k1_direct) {
                        const C part = qlane_k1_direct(pc, lv, r0, r1, NP, nw, Nin, q0x, q0y, q1x, q1y, has1, L);
                        if (has0) pf += part;
                    }
                }
            }
            acc += pf * FM::coef(f);
        }
    }
    return acc;
}
template <int KIND, int CH>
FDGA_HD C qlane_lane(const DevChain& V, const ColJob& job, const Grid& g, const C* __restrict__ R,
                     int iW, int inu, int iP, int ik, int lane) {
    uint4 ent[FDGA_QL_ENT];
    return qlane_rep<KIND, CH, false>(V, job, g, R, ent, nullptr, iW, inu, iP, iP % g.L, iP / g.L, ik % g.L, ik / g.L, lane);
}

// class representatives of this rank, sorted by slab: (iW | inu << 16, iP | ik << 16, Px | Py << 8 | kx << 16 | ky << 24, class slot)
struct RepDev { int nrep; const int4* rep; };

#ifndef FDGA_QL_WARPS
#define FDGA_QL_WARPS 4
#endif
#ifndef FDGA_QL_MINB
#define FDGA_QL_MINB 6
#endif
template <int KIND, int CH>
__global__ void __launch_bounds__(32 * FDGA_QL_WARPS, FDGA_QL_MINB)
qlane_kernel(const __grid_constant__ DevChain V, ColJob job, RepDev reps, const C* __restrict__ R,
             const C* __restrict__ OwnTab, const C* __restrict__ Rtot, const C* __restrict__ ConvTab,
             C* __restrict__ repvals, Grid g) {
    __shared__ uint4 s_ent[FDGA_QL_WARPS][FDGA_QL_ENT];
    extern __shared__ __align__(128) double ql_ring[];      // FDGA_QL_DEPTH x 4 KB per warp (cp.async ring), empty when FDGA_QL_DEPTH == 0
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = blockIdx.x * FDGA_QL_WARPS + wid;
    if (w >= reps.nrep) return;                     // whole warps leave together; only __syncwarp below
    const int4 rp = reps.rep[w];
    const int iW = rp.x & 0xffff, inu = rp.x >> 16, iP = rp.y & 0xffff, ik = (rp.y >> 16) & 0xffff;
    const int Px = rp.z & 0xff, Py = (rp.z >> 8) & 0xff, kx = (rp.z >> 16) & 0xff, ky = (rp.z >> 24) & 0xff;
    C* ring_lane = reinterpret_cast<C*>(ql_ring) + (size_t)wid * (FDGA_QL_DEPTH > 0 ? FDGA_QL_DEPTH : 1) * 256 + lane;
    C acc = qlane_rep<KIND, CH, true>(V, job, g, R, s_ent[wid], ring_lane, iW, inu, iP, Px, Py, kx, ky, lane);
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
