// fdga_swave.cuh -- kernels of the s-wave solver (NL_ParquetSolver, src/nonlocal/): vertices with bosonic momentum dependence
// only, K1[W,P], K2[W,v,P], K3[W,v,v',P], bubbles Pi[W,w,P]; every fermionic momentum is the s-wave point kSW.
//
// A context is in s-wave mode when dims.lev[0].type == FDGA_LV_NL.  On the device an NL level is an NL2 level whose table of
// momentum means K2swk[W,v,P] IS its K2 array (fdga_lib.cu: dev_level), so eval_vertex<true> (fdga_device.cuh) evaluates it as
// src/nonlocal/vertex.jl:213-377 prescribes: own channel at P, cross channels averaged over their transfer momentum.  The K3
// kernels and the caches (fdga_kernels.cuh) are shared with the NL2 solver: src/nonlocal/BSEa/BSEa_K3.jl and
// src/nonlocal/build_K3_cache.jl:18-94 are the NL2 files with Pi[W,w,P,kSW] -> Pi[W,w,P].
//
// The contractions are one-dimensional sums over the inner frequency (no momentum sum, no 1/N_q).  There are few class
// representatives (config 3: 240 for K1, ~500 for K2) and up to 2 N_Pi_nu = 1024 inner frequencies each, so the parallelism is
// taken from the inner sum: CTA = true gives one CTA of FDGA_SW_THREADS threads per representative (long sums over the bubble
// mesh), CTA = false one warp per representative (the 2 N_K2_nu terms of BSE_L_K2!).
#pragma once
#include "fdga_kernels.cuh"

namespace fdga {

#ifndef FDGA_SW_WARPS
#define FDGA_SW_WARPS 16     // warps per CTA (CTA-per-representative mode: 1024 inner frequencies -> 2 per thread)
#endif
#define FDGA_SW_THREADS (32 * FDGA_SW_WARPS)

// work split of the contraction kernels: representative handled by this thread, its first inner index and the stride
template <bool CTA>
struct SwSplit {
    long long cls; int first, stride; bool active;
    __device__ __forceinline__ SwSplit(long long c0, long long c1) {
        if (CTA) { cls = c0 + blockIdx.x; first = threadIdx.x; stride = FDGA_SW_THREADS; }
        else { cls = c0 + blockIdx.x * (long long)FDGA_SW_WARPS + (threadIdx.x >> 5); first = threadIdx.x & 31; stride = 32; }
        active = cls < c1;
    }
    // sum over the threads of the representative; the result is valid where `leader()` holds
    __device__ __forceinline__ C reduce(C v) const { return CTA ? block_reduce(v) : warp_reduce_sw(v); }
    __device__ __forceinline__ bool leader() const { return CTA ? threadIdx.x == 0 : (threadIdx.x & 31) == 0; }
    static __device__ __forceinline__ C warp_reduce_sw(C v) {
        for (int o = 16; o > 0; o >>= 1) { v.x += __shfl_down_sync(0xffffffffu, v.x, o); v.y += __shfl_down_sync(0xffffffffu, v.y, o); }
        return v;
    }
};

__device__ __forceinline__ Arg sw_arg(int W, int v, int w, int iP, int L) {
    Arg a; a.W = W; a.v = v; a.w = w; a.Px = iP % L; a.Py = iP / L; a.kx = a.ky = a.qx = a.qy = 0;
    return a;
}
template <int CH> __device__ __forceinline__ int crossing(int W, int w) { return CH == CH_P ? W - w - 1 : w; }      // src/convention.jl:39-47

// ---- BZ means of an NL level: K1sw[W], K2sww[W,v], K3sw[W,v,w] (src/nonlocal/swave.jl:32-75) ----
// one warp per output element, lanes over the momentum
__global__ void swave_tables_nl_kernel(DevLevel lv, int NP, SwOut out) {
    const int r = blockIdx.y;
    const DevChan& c = lv.ch[r];
    const long long n1 = 2 * lv.nK1 - 1, n2 = (long long)(2 * lv.nK2b - 1) * (2 * lv.nK2f), n3 = (long long)(2 * lv.nK3b - 1) * (2 * lv.nK3f) * (2 * lv.nK3f);
    long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const C* src; C* dst; long long n;
    if (i < n1) { src = c.K1; dst = out.p[r][0]; n = n1; }
    else if (i < n1 + n2) { i -= n1; src = c.K2; dst = out.p[r][2]; n = n2; }
    else if (i < n1 + n2 + n3) { i -= n1 + n2; src = c.K3; dst = out.p[r][3]; n = n3; }
    else return;
    C s = zeroC();
    for (int p = lane; p < NP; p += 32) s += src[i + n * p];
    s = SwSplit<false>::warp_reduce_sw(s);
    if (lane == 0) dst[i] = s / (double)NP;
}

// ---- BSE_K1!: src/nonlocal/BSEa/BSEa_K1.jl:2-58.  One warp per class representative (W, P) ----
template <int CH, bool MF, bool CTA>
__global__ void __launch_bounds__(FDGA_SW_THREADS)
sw_bse_k1_kernel(const __grid_constant__ DevChain F0, const __grid_constant__ DevChain F, const __grid_constant__ DevChain FL,
                 const C* __restrict__ Pi0, const C* __restrict__ Pi, C* __restrict__ repvals, SymDev sg, long long c0, long long c1, Grid g, double scale) {
    constexpr int SP = (CH == CH_T) ? SP_D : SP_P;
    const SwSplit<CTA> sp(c0, c1);
    if (!sp.active) return;           // uniform per warp (CTA = false) resp. per CTA
    const long long cls = sp.cls;
    const long long idx = sg.index[sg.offsets[cls]];
    const int nB1 = 2 * g.nK1 - 1;
    const int W = (int)(idx % nB1) - (g.nK1 - 1), iP = (int)(idx / nB1);
    C acc = zeroC();
    for (int w = -g.nPiF + sp.first; w < g.nPiF; w += sp.stride) {
        const size_t pi = piswat(g, W, w, iP);
        const int wc = crossing<CH>(W, w);
        const C FLr = eval_vertex<true>(FL, 0, CH, SP, sw_arg(W, wc, FDGA_INF, iP, g.L), FL_ALL);
        if (MF) {
            const C Fl = eval_vertex<true>(F0, 0, CH, SP, sw_arg(W, FDGA_INF, w, iP, g.L), FL_ALL);
            acc += Fl * Pi0[pi] * FLr;
        } else {
            const C Fl = eval_vertex<true>(F, 0, CH, SP, sw_arg(W, FDGA_INF, w, iP, g.L), FL_ALL);
            const C F0r = eval_vertex<true>(F0, 0, CH, SP, sw_arg(W, wc, FDGA_INF, iP, g.L), FL_ALL);
            const C p1 = Pi[pi], p0 = Pi0[pi];
            acc += Fl * ((p1 - p0) * F0r + p1 * FLr);
        }
    }
    acc = sp.reduce(acc);
    if (sp.leader()) repvals[cls] = acc * scale;
}

// K2-shaped representatives (W, v, P)
__device__ __forceinline__ void decode_k2sw(const Grid& g, long long idx, int& W, int& v, int& iP) {
    const int nB2 = 2 * g.nK2b - 1, nF2 = 2 * g.nK2f;
    const int iW = (int)(idx % nB2); idx /= nB2; const int iv = (int)(idx % nF2); iP = (int)(idx / nF2);
    W = iW - (g.nK2b - 1); v = iv - g.nK2f;
}

// ---- BSE_L_K2!: src/nonlocal/BSEa/BSEa_K2.jl:1-41 (w over the K2 fermionic mesh) ----
template <int CH, bool CTA>
__global__ void __launch_bounds__(FDGA_SW_THREADS)
sw_bse_lk2_kernel(const __grid_constant__ DevChain F0, const __grid_constant__ DevChain F, const C* __restrict__ Pi0,
                  C* __restrict__ repvals, SymDev sg, long long c0, long long c1, Grid g, double scale) {
    constexpr int SP = (CH == CH_T) ? SP_D : SP_P;
    constexpr unsigned FLG = (CH == CH_P ? 0u : FL_GP) | (CH == CH_T ? 0u : FL_GT) | (CH == CH_A ? 0u : FL_GA);
    const SwSplit<CTA> sp(c0, c1);
    if (!sp.active) return;
    const long long cls = sp.cls;
    int W, v, iP; decode_k2sw(g, sg.index[sg.offsets[cls]], W, v, iP);
    C acc = zeroC();
    for (int w = -g.nK2f + sp.first; w < g.nK2f; w += sp.stride) {
        const C Gl = eval_vertex<true>(F, 0, CH, SP, sw_arg(W, v, crossing<CH>(W, w), iP, g.L), FLG);
        const C F0r = eval_vertex<true>(F0, 0, CH, SP, sw_arg(W, w, FDGA_INF, iP, g.L), FL_ALL);
        acc += Gl * Pi0[piswat(g, W, w, iP)] * F0r;
    }
    acc = sp.reduce(acc);
    if (sp.leader()) repvals[cls] = acc * scale;
}

// ---- BSE_K2!: src/nonlocal/BSEa/BSEa_K2.jl:44-106 (the FL.K2 add is the caller's post-fix) ----
template <int CH, bool MF, bool CTA>
__global__ void __launch_bounds__(FDGA_SW_THREADS)
sw_bse_k2_kernel(const __grid_constant__ DevChain F0, const __grid_constant__ DevChain F, const __grid_constant__ DevChain FL,
                 const C* __restrict__ Pi0, const C* __restrict__ Pi, C* __restrict__ repvals, SymDev sg, long long c0, long long c1, Grid g, double scale) {
    constexpr int SP = (CH == CH_T) ? SP_D : SP_P;
    const SwSplit<CTA> sp(c0, c1);
    if (!sp.active) return;
    const long long cls = sp.cls;
    int W, v, iP; decode_k2sw(g, sg.index[sg.offsets[cls]], W, v, iP);
    C acc = zeroC();
    for (int w = -g.nPiF + sp.first; w < g.nPiF; w += sp.stride) {
        const size_t pi = piswat(g, W, w, iP);
        const int wc = crossing<CH>(W, w);
        if (MF) {
            const C Fl = eval_vertex<true>(F0, 0, CH, SP, sw_arg(W, v, wc, iP, g.L), FL_ALL) - eval_vertex<true>(F0, 0, CH, SP, sw_arg(W, FDGA_INF, wc, iP, g.L), FL_ALL);
            const C FLr = eval_vertex<true>(FL, 0, CH, SP, sw_arg(W, w, FDGA_INF, iP, g.L), FL_ALL);
            acc += Fl * Pi0[pi] * FLr;
        } else {
            const C Fl = eval_vertex<true>(F, 0, CH, SP, sw_arg(W, v, w, iP, g.L), FL_ALL) - eval_vertex<true>(F, 0, CH, SP, sw_arg(W, FDGA_INF, w, iP, g.L), FL_ALL);
            const C F0r = eval_vertex<true>(F0, 0, CH, SP, sw_arg(W, wc, FDGA_INF, iP, g.L), FL_ALL);
            const C FLr = eval_vertex<true>(FL, 0, CH, SP, sw_arg(W, wc, FDGA_INF, iP, g.L), FL_ALL);
            const C p1 = Pi[pi], p0 = Pi0[pi];
            acc += Fl * ((p1 - p0) * F0r + p1 * FLr);
        }
    }
    acc = sp.reduce(acc);
    if (sp.leader()) repvals[cls] = acc * scale;
}

// ---- SDE_channel_L_pp! / ph!: src/nonlocal/SDE.jl:3-146.  The levels `from..nlev-1` of the chain are summed in one pass
// (the rest of SDE_compute! is linear in L): own reducible vertex of a Vertex / NL_Vertex level, (core - bare) / 3 of the
// RefVertex level (SDE.jl:179-183) ----
template <bool PP, bool CTA>
__global__ void __launch_bounds__(FDGA_SW_THREADS)
sw_sde_L_kernel(const __grid_constant__ DevChain V, int from, const C* __restrict__ Pi, C* __restrict__ repvals, SymDev sg,
                long long c0, long long c1, Grid g, C U, double scale) {
    const SwSplit<CTA> sp(c0, c1);
    if (!sp.active) return;
    const long long cls = sp.cls;
    int W, v, iP; decode_k2sw(g, sg.index[sg.offsets[cls]], W, v, iP);
    C acc = zeroC();
    for (int w = -g.nPiF + sp.first; w < g.nPiF; w += sp.stride) {
        C d = zeroC();
        for (int l = from; l < V.nlev; ++l) {
            const DevLevel& lv = V.lev[l];
            if (lv.type == LV_CORE) {
                C c = PP ? core_eval(lv, CH_P, SP_P, W, W - w - 1, v) - U
                         : core_eval(lv, CH_T, SP_P, W, v, w) + core_eval(lv, CH_A, SP_P, W, v, w) - U - U;
                d += c * (1.0 / 3.0);
            } else if (lv.type == LV_LOCAL) {
                d += PP ? loc_chan(lv, CH_P, W, W - w - 1, v) : loc_chan(lv, CH_T, W, v, w) + loc_chan(lv, CH_A, W, v, w);
            } else {
                d += PP ? nl2_chan_sw_own(lv, CH_P, V.NP, W, W - w - 1, v, iP)
                        : nl2_chan_sw_own(lv, CH_T, V.NP, W, v, w, iP) + nl2_chan_sw_own(lv, CH_A, V.NP, W, v, w, iP);
            }
        }
        acc += U * Pi[piswat(g, W, w, iP)] * d;
    }
    acc = sp.reduce(acc);
    if (sp.leader()) repvals[cls] = acc * scale;
}

// the elements R of the window [-h, h] with R = t (mod n), 0 <= t < n (n >= 2h): at most two
__device__ __forceinline__ int window_images(int t, int n, int h, int* out) {
    int k = 0;
    if (t <= h) out[k++] = t;
    if (t - n >= -h) out[k++] = t - n;
    return k;
}

// ---- bubbles_real_space!(::NL_MF_Pi): src/nonlocal/bubble.jl:87-158.  GR = fft(G) / LG^2 [nu, R].  Real-space fill in gather
// form: one thread per (W, w, R mod L); the back transform over the two momentum axes follows (dft_axis_kernel).
//   Pipp(R) = G(W - w, R) G(w, R) wt(R),  Piph(R) = G(W + w, R) G(w, -R) wt(R),  R in [-L/2, L/2]^2,
//   wt = 1/2 per component with |R_c| = LG/2 (even LG); at R = 0 the Green function continues as 1/nu outside its mesh. ----
__global__ void sw_bubbles_rs_kernel(const C* __restrict__ GR, C* __restrict__ PippR, C* __restrict__ PiphR, Grid g, int use_tail) {
    const int nBP = 2 * g.nPiB - 1, nFP = 2 * g.nPiF, L = g.L, LG = g.LG, nG = g.nG, h = L / 2;
    const long long n = (long long)nBP * nFP * g.NP;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int iW = (int)(i % nBP); long long r = i / nBP; const int iw = (int)(r % nFP); const int iR = (int)(r / nFP);
    const int W = iW - (g.nPiB - 1), w = iw - g.nPiF, tx = iR % L, ty = iR / L;
    const double pi = 3.14159265358979323846;
    C pp = zeroC(), ph = zeroC();
    // the R in [-h, h] with R = t (mod L): t itself and t - L (both at the zone edge of an even mesh)
    int c1[2], c2[2]; const int n1 = window_images(tx, L, h, c1), n2 = window_images(ty, L, h, c2);
    for (int i2 = 0; i2 < n2; ++i2) {
        const int R2 = c2[i2];
        for (int i1 = 0; i1 < n1; ++i1) {
            const int R1 = c1[i1];
            double wt = 1.0;
            if (LG % 2 == 0) { if (abs(R1) == LG / 2) wt *= 0.5; if (abs(R2) == LG / 2) wt *= 0.5; }
            const bool tail = use_tail && R1 == 0 && R2 == 0;
            const size_t pR = (size_t)2 * nG * (modL(R1, LG) + (size_t)LG * modL(R2, LG)), mR = (size_t)2 * nG * (modL(-R1, LG) + (size_t)LG * modL(-R2, LG));
            auto gat = [&](int nn, size_t off) -> C {
                if (inF(nn, nG)) return GR[posF(nn, nG) + off];
                return tail ? mkC(1.0 / ((2 * nn + 1) * pi * g.T), 0.0) : zeroC();
            };
            const C gw = gat(w, pR);
            pp += gat(W - w - 1, pR) * gw * wt;
            ph += gat(W + w, pR) * gat(w, mR) * wt;
        }
    }
    PippR[i] = pp; PiphR[i] = ph;
}

// The same bubbles without the bubble-sized transforms (default route; the kernel above + two DFT passes is the literal one,
// FDGA_BUBBLES_RS=1).  Pi(R)[W,w] vanishes unless BOTH frequencies lie on the G mesh -- 2 N_G of the 2 N_Pi_nu inner frequencies
// (32 of 1024 at config 3) -- except for the tail, which only lives at R = 0 and is therefore P-independent.  So
//   both on the mesh:  Pi[W,w,P] = sum_{R in window} wt(R) G(a,R) G(w,+-R) exp(+2 pi i P.R / L)   (the back transform as a direct sum)
//   otherwise:         Pi[W,w,P] = g~(a) g~(w)  with the local G (R = 0) continued by 1/nu   (0 without the tail).
__global__ void sw_bubbles_direct_kernel(const C* __restrict__ GR, C* __restrict__ Pipp, C* __restrict__ Piph, Grid g, int use_tail,
                                         const C* __restrict__ twL, long long nblk_fill) {
    const int nBP = 2 * g.nPiB - 1, nFP = 2 * g.nPiF, L = g.L, LG = g.LG, NP = g.NP, nG = g.nG, h = L / 2;
    const double pi = 3.14159265358979323846;
    auto tail = [&](int n) { return mkC(1.0 / ((2 * n + 1) * pi * g.T), 0.0); };
    if ((long long)blockIdx.x < nblk_fill) {
        // role 1, one thread per element whose inner frequency w lies OUTSIDE the G mesh: the P-independent tail product (or 0)
        const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
        if (i >= (long long)nBP * nFP * NP) return;
        long long t = i;
        const int iW = t % nBP; t /= nBP; const int iw = t % nFP;
        const int W = iW - (g.nPiB - 1), w = iw - g.nPiF;
        if (inF(w, nG)) return;                              // role 2 writes these
        C pp = zeroC(), ph = zeroC();
        if (use_tail) {
            const C gb = tail(w);
            const int app = W - w - 1, aph = W + w;
            pp = (inF(app, nG) ? GR[posF(app, nG)] : tail(app)) * gb;
            ph = (inF(aph, nG) ? GR[posF(aph, nG)] : tail(aph)) * gb;
        }
        Pipp[i] = pp; Piph[i] = ph;
        return;
    }
    // role 2, one warp per element (W, w on the G mesh, P): lanes over the (L+1)^2 window of R, shuffle reduction
    const long long e = ((blockIdx.x - nblk_fill) * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long nE = (long long)nBP * (2 * nG) * NP;
    if (e >= nE) return;
    long long t = e;
    const int iW = t % nBP; t /= nBP; const int w = (int)(t % (2 * nG)) - nG; const int iP = (int)(t / (2 * nG));
    const int W = iW - (g.nPiB - 1), Px = iP % L, Py = iP / L;
    if (!inF(w, g.nPiF)) return;                             // G mesh wider than the bubble's inner mesh
    const int app = W - w - 1, aph = W + w;
    const bool pin = inF(app, nG), hin = inF(aph, nG);
    C pp = zeroC(), ph = zeroC();
    if (pin || hin) {
        const int win = 2 * h + 1;
        for (int j = lane; j < win * win; j += 32) {
            const int R1 = j % win - h, R2 = j / win - h;
            double wt = 1.0;
            if (LG % 2 == 0) { if (abs(R1) == LG / 2) wt *= 0.5; if (abs(R2) == LG / 2) wt *= 0.5; }
            const size_t pR = (size_t)2 * nG * (modL(R1, LG) + (size_t)LG * modL(R2, LG)), mR = (size_t)2 * nG * (modL(-R1, LG) + (size_t)LG * modL(-R2, LG));
            const C ph_ = twL[modL(Px * R1 + Py * R2, L)] * wt;
            if (pin) pp += GR[posF(app, nG) + pR] * GR[posF(w, nG) + pR] * ph_;
            if (hin) ph += GR[posF(aph, nG) + pR] * GR[posF(w, nG) + mR] * ph_;
        }
        pp = SwSplit<false>::warp_reduce_sw(pp); ph = SwSplit<false>::warp_reduce_sw(ph);
    }
    if (lane != 0) return;
    const C gb = GR[posF(w, nG)];
    if (!pin) pp = use_tail ? tail(app) * gb : zeroC();
    if (!hin) ph = use_tail ? tail(aph) * gb : zeroC();
    const size_t o = (size_t)iW + (size_t)nBP * (posF(w, g.nPiF) + (size_t)nFP * iP);
    Pipp[o] = pp; Piph[o] = ph;
}

// ---- SDE_compute_inner! (use_real_space = true): src/nonlocal/SDE.jl:191-275.  LppR / LphR = fft(L) / L^2 over the momentum
// axis [W, v, R]; GR = fft(G) / LG^2.  Gather form, one thread per (nu, target R of Sigma):
//   Sigma( R) += G(W - nu, -R) Lpp(W, nu, R) wt,   Sigma(-R) += G(W + nu, -R) Lph(W, nu, R) wt,   R in [-L/2, L/2]^2 ----
__global__ void sw_sde_rs_kernel(const C* __restrict__ GR, const C* __restrict__ LppR, const C* __restrict__ LphR, C* __restrict__ SigR, Grid g) {
    const int nB = 2 * g.nK2b - 1, nF = 2 * g.nK2f, L = g.L, LG = g.LG, nG = g.nG, h = L / 2;
    const long long n = (long long)2 * nG * LG * LG;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int v = (int)(i % (2 * nG)) - nG; const int t = (int)(i / (2 * nG)), tx = t % LG, ty = t / LG;
    C acc = zeroC();
    if (inF(v, g.nK2f)) {
        const size_t pre = (size_t)nB * nF;
        // sources R with R = t (pp term) or -R = t (ph term) modulo LG, both inside the window: the images of t and of -t
        int cand1[4], cand2[4];
        int n1 = window_images(tx, LG, h, cand1), n2 = window_images(ty, LG, h, cand2);
        n1 += window_images(modL(-tx, LG), LG, h, cand1 + n1); n2 += window_images(modL(-ty, LG), LG, h, cand2 + n2);
        for (int i2 = 0; i2 < n2; ++i2) for (int i1 = 0; i1 < n1; ++i1) {
            const int R1 = cand1[i1], R2 = cand2[i2];
            bool dup = false;       // t = -t (mod LG) lists the same R twice
            for (int j = 0; j < i1; ++j) dup = dup || cand1[j] == R1;
            for (int j = 0; j < i2; ++j) dup = dup || cand2[j] == R2;
            if (dup) continue;
            const bool hit_p = modL(R1, LG) == tx && modL(R2, LG) == ty, hit_m = modL(-R1, LG) == tx && modL(-R2, LG) == ty;
            if (!hit_p && !hit_m) continue;
            double wt = 1.0;
            if (L % 2 == 0) { if (abs(R1) == L / 2) wt *= 0.5; if (abs(R2) == L / 2) wt *= 0.5; }
            const size_t iRL = pre * (modL(R1, L) + (size_t)L * modL(R2, L)) + (size_t)nB * posF(v, g.nK2f);
            const size_t mG = (size_t)2 * nG * (modL(-R1, LG) + (size_t)LG * modL(-R2, LG));
            for (int iW = 0; iW < nB; ++iW) {
                const int W = iW - (g.nK2b - 1);
                if (hit_p && inF(W - v - 1, nG)) acc += GR[posF(W - v - 1, nG) + mG] * LppR[iW + iRL] * wt;
                if (hit_m && inF(W + v, nG))     acc += GR[posF(W + v, nG) + mG] * LphR[iW + iRL] * wt;
            }
        }
    }
    SigR[i] = acc * g.T;
}

// mix / copy helpers on bubble-shaped arrays are the generic axpby kernels of fdga_kernels.cuh

}  // namespace fdga
