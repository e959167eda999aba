// fdga_device.cuh -- device-side data model and vertex evaluators (sm_100a).
//
// Device restatement of the reference's callable vertex structs as a FLATTENED chain of
// levels (no recursion, no dynamic dispatch):
//   NL2_Channel evaluator      src/nonlocal_2/channel.jl:58-194
//   local Channel evaluator    src/channel.jl:220-339
//   RefVertex                  src/refvertex.jl:90-216
//   Vertex / NL2_Vertex        src/vertex.jl:209-336, src/nonlocal/vertex.jl:69-211,
//                              src/nonlocal_2/vertex.jl:207-259 (double s-wave)
//   channel conversions        src/convention.jl:4-37
// The s-wave point kSW (src/nonlocal/swave.jl) never triggers a momentum sum here: the BZ means
// are pre-tabulated per level (K1sw, K2swk, K2sww, K3sw) by swave_tables_kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define FDGA_MAXLEV 6
// evaluators are host-callable too so that the index machinery can be unit-tested without a GPU (tests/host_eval_test.cu)
#define FDGA_HD __host__ __device__ __forceinline__

namespace fdga {

struct __align__(16) C {
    double x, y;
};
FDGA_HD C mkC(double x, double y) { C r; r.x = x; r.y = y; return r; }
FDGA_HD C operator+(C a, C b) { return mkC(a.x + b.x, a.y + b.y); }
FDGA_HD C operator-(C a, C b) { return mkC(a.x - b.x, a.y - b.y); }
FDGA_HD C operator-(C a) { return mkC(-a.x, -a.y); }
FDGA_HD C operator*(C a, C b) { return mkC(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
FDGA_HD C operator*(C a, double s) { return mkC(a.x * s, a.y * s); }
FDGA_HD C operator*(double s, C a) { return mkC(a.x * s, a.y * s); }
FDGA_HD C operator/(C a, double s) { return mkC(a.x / s, a.y / s); }
FDGA_HD C& operator+=(C& a, C b) { a.x += b.x; a.y += b.y; return a; }
FDGA_HD C conjC(C a) { return mkC(a.x, -a.y); }
FDGA_HD C zeroC() { return mkC(0.0, 0.0); }

FDGA_HD C ldg(const C* p) {
#if defined(__CUDA_ARCH__) && !defined(FDGA_PLAIN_LDG)
    double2 v = __ldg(reinterpret_cast<const double2*>(p));
    return mkC(v.x, v.y);
#else
    return *p;
#endif
}

// ---- Matsubara index arithmetic (fermion n <-> (2n+1) pi T, boson m <-> 2 m pi T) ----------
#define FDGA_INF (1 << 28)
FDGA_HD bool isinfF(int a) { return a >= (1 << 27); }
FDGA_HD bool inB(int m, int N) { return m >= -(N - 1) && m <= N - 1; }
FDGA_HD bool inF(int n, int N) { return n >= -N && n <= N - 1; }   // false for FDGA_INF
FDGA_HD int posB(int m, int N) { return m + N - 1; }
FDGA_HD int posF(int n, int N) { return n + N; }
FDGA_HD int modL(int a, int L) { int r = a % L; return r < 0 ? r + L : r; }
FDGA_HD int kidx(int x, int y, int L) { return modL(x, L) + L * modL(y, L); }

enum { CH_P = 0, CH_T = 1, CH_A = 2 };
enum { SP_P = 0, SP_X = 1, SP_D = 2 };
enum { LV_NL2 = 0, LV_LOCAL = 1, LV_CORE = 2 };
enum { FL_F0 = 1, FL_GP = 2, FL_GT = 4, FL_GA = 8, FL_ALL = 15 };

struct DevChan {
    const C *K1, *K2, *K3;
    const C *K1sw;    // [W]        mean_P K1[W,P]
    const C *K2swk;   // [W,v,P]    mean_k K2[W,v,P,k]
    const C *K2sww;   // [W,v]      mean_{P,k} K2
    const C *K3sw;    // [W,v,w]    mean_P K3[W,v,w,P]
    const C *K1h;     // [kappa,W]  sum_P K1[W,P] exp(-2 pi i kappa.P / L)  (kappa fast; fdga_column.cuh: slab_conv_kernel)
    const C *K2m[4];  // K2 in the four momentum-fastest layouts ML_P, ML_K, ML_S, ML_D (fdga_qlane.cuh); null until built
    const C *K3m;     // [P,W,v,w]  K3 with the transfer momentum fastest
    const C *K1m;     // [P,W]      K1 with the transfer momentum fastest
};
struct DevLevel {
    int type, nK1, nK2b, nK2f, nK3b, nK3f, mbe /* 1: MBEVertex / NL2_MBEVertex evaluation (src/boson_exchange.jl) */, pad1;
    C U;
    DevChan ch[3];
    const C* core[4];
};
struct DevChain {
    int nlev, L, NP, pad;
    DevLevel lev[FDGA_MAXLEV];
};

// ---- channel evaluators (all K switches on) -------------------------------------------------
// v or w may be FDGA_INF; the unified form below reproduces the four reference methods
// (src/nonlocal_2/channel.jl:58-194) for K1 = K2 = K3 = true.
FDGA_HD C nl2_chan(const DevLevel& lv, int r, int NP, int W, int v, int w, int iP, int ik, int iq) {
    C val = zeroC();
    if (!inB(W, lv.nK1)) return val;
    const DevChan& c = lv.ch[r];
    val += ldg(c.K1 + posB(W, lv.nK1) + (size_t)(2 * lv.nK1 - 1) * iP);
    if (!inB(W, lv.nK2b)) return val;
    bool a = inF(v, lv.nK2f), b = inF(w, lv.nK2f);
    int nB = 2 * lv.nK2b - 1, nF = 2 * lv.nK2f;
    size_t sP = (size_t)nB * nF;
    if (a) val += ldg(c.K2 + posB(W, lv.nK2b) + (size_t)nB * posF(v, lv.nK2f) + sP * (iP + (size_t)NP * ik));
    if (b) val += ldg(c.K2 + posB(W, lv.nK2b) + (size_t)nB * posF(w, lv.nK2f) + sP * (iP + (size_t)NP * iq));
    if (a && b && inB(W, lv.nK3b) && inF(v, lv.nK3f) && inF(w, lv.nK3f)) {
        int nB3 = 2 * lv.nK3b - 1, nF3 = 2 * lv.nK3f;
        val += ldg(c.K3 + posB(W, lv.nK3b) + (size_t)nB3 * (posF(v, lv.nK3f) + (size_t)nF3 * (posF(w, lv.nK3f) + (size_t)nF3 * iP)));
    }
    return val;
}
// own channel with k = q = kSW: K1[W,P] + mean_k K2[W,v,P,k] + mean_k K2[W,w,P,k] + K3[W,v,w,P]
FDGA_HD C nl2_chan_sw_own(const DevLevel& lv, int r, int NP, int W, int v, int w, int iP) {
    C val = zeroC();
    if (!inB(W, lv.nK1)) return val;
    const DevChan& c = lv.ch[r];
    val += ldg(c.K1 + posB(W, lv.nK1) + (size_t)(2 * lv.nK1 - 1) * iP);
    if (!inB(W, lv.nK2b)) return val;
    bool a = inF(v, lv.nK2f), b = inF(w, lv.nK2f);
    int nB = 2 * lv.nK2b - 1, nF = 2 * lv.nK2f;
    if (a) val += ldg(c.K2swk + posB(W, lv.nK2b) + (size_t)nB * (posF(v, lv.nK2f) + (size_t)nF * iP));
    if (b) val += ldg(c.K2swk + posB(W, lv.nK2b) + (size_t)nB * (posF(w, lv.nK2f) + (size_t)nF * iP));
    if (a && b && inB(W, lv.nK3b) && inF(v, lv.nK3f) && inF(w, lv.nK3f)) {
        int nB3 = 2 * lv.nK3b - 1, nF3 = 2 * lv.nK3f;
        val += ldg(c.K3 + posB(W, lv.nK3b) + (size_t)nB3 * (posF(v, lv.nK3f) + (size_t)nF3 * (posF(w, lv.nK3f) + (size_t)nF3 * iP)));
    }
    return val;
}
// cross channel with (kSW, kSW, kSW): everything BZ-averaged
FDGA_HD C nl2_chan_sw_cross(const DevLevel& lv, int r, int W, int v, int w) {
    C val = zeroC();
    if (!inB(W, lv.nK1)) return val;
    const DevChan& c = lv.ch[r];
    val += ldg(c.K1sw + posB(W, lv.nK1));
    if (!inB(W, lv.nK2b)) return val;
    bool a = inF(v, lv.nK2f), b = inF(w, lv.nK2f);
    int nB = 2 * lv.nK2b - 1;
    if (a) val += ldg(c.K2sww + posB(W, lv.nK2b) + (size_t)nB * posF(v, lv.nK2f));
    if (b) val += ldg(c.K2sww + posB(W, lv.nK2b) + (size_t)nB * posF(w, lv.nK2f));
    if (a && b && inB(W, lv.nK3b) && inF(v, lv.nK3f) && inF(w, lv.nK3f)) {
        int nB3 = 2 * lv.nK3b - 1, nF3 = 2 * lv.nK3f;
        val += ldg(c.K3sw + posB(W, lv.nK3b) + (size_t)nB3 * (posF(v, lv.nK3f) + (size_t)nF3 * posF(w, lv.nK3f)));
    }
    return val;
}
FDGA_HD C loc_chan(const DevLevel& lv, int r, int W, int v, int w) {
    C val = zeroC();
    if (!inB(W, lv.nK1)) return val;
    const DevChan& c = lv.ch[r];
    val += ldg(c.K1 + posB(W, lv.nK1));
    if (!inB(W, lv.nK2b)) return val;
    bool a = inF(v, lv.nK2f), b = inF(w, lv.nK2f);
    int nB = 2 * lv.nK2b - 1;
    if (a) val += ldg(c.K2 + posB(W, lv.nK2b) + (size_t)nB * posF(v, lv.nK2f));
    if (b) val += ldg(c.K2 + posB(W, lv.nK2b) + (size_t)nB * posF(w, lv.nK2f));
    if (a && b && inB(W, lv.nK3b) && inF(v, lv.nK3f) && inF(w, lv.nK3f)) {
        int nB3 = 2 * lv.nK3b - 1, nF3 = 2 * lv.nK3f;
        val += ldg(c.K3 + posB(W, lv.nK3b) + (size_t)nB3 * (posF(v, lv.nK3f) + (size_t)nF3 * posF(w, lv.nK3f)));
    }
    return val;
}

// ---- RefVertex (src/refvertex.jl:90-216) ---------------------------------------------------
FDGA_HD C core_call(const DevLevel& lv, int which, int W, int v, int w) {
    if (!(inB(W, lv.nK3b) && inF(v, lv.nK3f) && inF(w, lv.nK3f))) return zeroC();
    int nB = 2 * lv.nK3b - 1, nF = 2 * lv.nK3f;
    return ldg(lv.core[which] + posB(W, lv.nK3b) + (size_t)nB * (posF(v, lv.nK3f) + (size_t)nF * posF(w, lv.nK3f)));
}
FDGA_HD C core_eval_px(const DevLevel& lv, int Ch, int Sp, int W, int v, int w) {
    if (isinfF(v) || isinfF(w)) return (Sp == SP_X) ? -lv.U : lv.U;
    if (Sp == SP_P) {
        if (Ch == CH_P) return core_call(lv, 0, W, v, w) + lv.U;
        if (Ch == CH_T) return core_call(lv, 2, W, v, w) + lv.U;
        return -core_call(lv, 3, W, w, v) + lv.U;
    }
    if (Ch == CH_P) return core_call(lv, 1, W, v, w) - lv.U;
    if (Ch == CH_T) return core_call(lv, 3, W, v, w) - lv.U;
    return -core_call(lv, 2, W, w, v) - lv.U;
}
FDGA_HD C core_eval(const DevLevel& lv, int Ch, int Sp, int W, int v, int w) {
    if (Sp != SP_D) return core_eval_px(lv, Ch, Sp, W, v, w);
    if (isinfF(v) || isinfF(w)) return lv.U;
    return 2.0 * core_eval_px(lv, Ch, SP_P, W, v, w) + core_eval_px(lv, Ch, SP_X, W, v, w);
}

// ---- _convert_channel (src/convention.jl:12-35); frequency offsets from the index algebra:
//   B(m)-F(n) = F(m-n-1), B(m)+F(n) = F(m+n), F(a)-F(b) = B(a-b), F(a)+F(b) = B(a+b+1)
struct Arg {
    int W, v, w;          // Matsubara indices (finite)
    int Px, Py, kx, ky, qx, qy;
};
FDGA_HD Arg convert(const Arg& a, int from, int to) {
    Arg b = a;
    if (from == to) return b;
    if (from == CH_P && to == CH_T) {
        b.W = a.W - a.v - a.w - 1; b.v = a.w; b.w = a.v;
        b.Px = a.Px - a.kx - a.qx; b.Py = a.Py - a.ky - a.qy; b.kx = a.qx; b.ky = a.qy; b.qx = a.kx; b.qy = a.ky;
    } else if (from == CH_P && to == CH_A) {
        b.W = a.v - a.w; b.v = a.W - a.v - 1; b.w = a.w;
        b.Px = a.kx - a.qx; b.Py = a.ky - a.qy; b.kx = a.Px - a.kx; b.ky = a.Py - a.ky; b.qx = a.qx; b.qy = a.qy;
    } else if (from == CH_T && to == CH_P) {
        b.W = a.W + a.v + a.w + 1; b.v = a.w; b.w = a.v;
        b.Px = a.Px + a.kx + a.qx; b.Py = a.Py + a.ky + a.qy; b.kx = a.qx; b.ky = a.qy; b.qx = a.kx; b.qy = a.ky;
    } else if (from == CH_T && to == CH_A) {
        b.W = a.w - a.v; b.v = a.W + a.v; b.w = a.v;
        b.Px = a.qx - a.kx; b.Py = a.qy - a.ky; b.kx = a.Px + a.kx; b.ky = a.Py + a.ky; b.qx = a.kx; b.qy = a.ky;
    } else if (from == CH_A && to == CH_P) {
        b.W = a.W + a.w + a.v + 1; b.v = a.W + a.w; b.w = a.w;
        b.Px = a.Px + a.qx + a.kx; b.Py = a.Py + a.qy + a.ky; b.kx = a.Px + a.qx; b.ky = a.Py + a.qy; b.qx = a.qx; b.qy = a.qy;
    } else {  // a -> t
        b.W = a.v - a.w; b.v = a.w; b.w = a.W + a.w;
        b.Px = a.kx - a.qx; b.Py = a.ky - a.qy; b.kx = a.qx; b.ky = a.qy; b.qx = a.Px + a.qx; b.qy = a.Py + a.qy;
    }
    return b;
}

// ---- multi-boson-exchange vertices (src/boson_exchange.jl) ---------------------------------------------------------------------
// class Cl of channel r summed down the chain from level l0 (:270-330): the MeshFunction CALL of every level's array (0 outside
// its own mesh and for infinite frequencies), Lambda from the RefVertex only (:171-232).  Momenta are Brillouin points.
enum { CL_K1 = 0, CL_K2 = 1, CL_K2P = 2, CL_K3 = 3, CL_LAMBDA = 4 };
struct Classes { C K1, K2, K2p, K3; };
__host__ __device__ inline Classes mbe_classes(const DevChain& c, int l0, int r, const Arg& b) {
    Classes o; o.K1 = o.K2 = o.K2p = o.K3 = zeroC();
    const int L = c.L, NP = c.NP;
    for (int l = l0; l < c.nlev; ++l) {
        const DevLevel& lv = c.lev[l];
        if (lv.type == LV_CORE) break;
        const DevChan& ch = lv.ch[r];
        const bool nl2 = lv.type == LV_NL2;
        const size_t iP = nl2 ? kidx(b.Px, b.Py, L) : 0, ik = nl2 ? kidx(b.kx, b.ky, L) : 0, iq = nl2 ? kidx(b.qx, b.qy, L) : 0;
        if (inB(b.W, lv.nK1)) o.K1 += ldg(ch.K1 + posB(b.W, lv.nK1) + (size_t)(2 * lv.nK1 - 1) * iP);
        if (inB(b.W, lv.nK2b)) {
            const int nB = 2 * lv.nK2b - 1, nF = 2 * lv.nK2f;
            const size_t sP = (size_t)nB * nF;
            if (inF(b.v, lv.nK2f)) o.K2 += ldg(ch.K2 + posB(b.W, lv.nK2b) + (size_t)nB * posF(b.v, lv.nK2f) + sP * (iP + (size_t)NP * ik));
            if (inF(b.w, lv.nK2f)) o.K2p += ldg(ch.K2 + posB(b.W, lv.nK2b) + (size_t)nB * posF(b.w, lv.nK2f) + sP * (iP + (size_t)NP * iq));
        }
        if (inB(b.W, lv.nK3b) && inF(b.v, lv.nK3f) && inF(b.w, lv.nK3f)) {
            const int nB3 = 2 * lv.nK3b - 1, nF3 = 2 * lv.nK3f;
            o.K3 += ldg(ch.K3 + posB(b.W, lv.nK3b) + (size_t)nB3 * (posF(b.v, lv.nK3f) + (size_t)nF3 * (posF(b.w, lv.nK3f) + (size_t)nF3 * iP)));
        }
    }
    return o;
}
__host__ __device__ inline C mbe_lambda(const DevChain& c, int Ch, const Arg& a) {
    const DevLevel& lv = c.lev[c.nlev - 1];
    if (Ch == CH_P) return core_call(lv, 0, a.W, a.v, a.w);
    if (Ch == CH_T) return core_call(lv, 2, a.W, a.v, a.w);
    return -core_call(lv, 3, a.W, a.w, a.v);
}
__host__ __device__ inline Arg convert_inf(const Arg& a, int from, int to) {      // convert() with infinite frequencies propagated
    Arg b = convert(a, from, to);
    if (from != to && (isinfF(a.v) || isinfF(a.w))) {
        // every converted frequency that involves an infinite one is infinite: the class arrays then evaluate to 0 (:349-422 with
        // the MeshFunction call of an InfiniteMatsubaraFrequency)
        b.W = FDGA_INF; b.v = isinfF(a.v) || isinfF(a.w) ? FDGA_INF : b.v; b.w = FDGA_INF;
    }
    return b;
}
FDGA_HD C mbe_sbe(const Classes& k, C u) {      // K1 + K2 + K2' + K2 K2' / (u + K1) + K3
    const C d = u + k.K1;
    const double n = d.x * d.x + d.y * d.y;
    const C num = k.K2 * k.K2p;
    const C q = mkC((num.x * d.x + num.y * d.y) / n, (num.y * d.x - num.x * d.y) / n);
    return k.K1 + k.K2 + k.K2p + q + k.K3;
}
// F(W, v, w, P, k, q, Ch, pSp; F0 = true, gamma switches) of the MBE chain starting at level l, Brillouin-point momenta (:349-416)
__host__ __device__ inline C mbe_total(const DevChain& c, int l, int Ch, const Arg& a, unsigned f) {
    const C U = c.lev[c.nlev - 1].U;
    C val = U;
    if (f & FL_GP) val += mbe_sbe(mbe_classes(c, l, CH_P, convert_inf(a, Ch, CH_P)), U);
    if (f & FL_GT) {      // (tCh, pSp) = (D - M) / 2,  M = -(aCh, pSp),  D = 2 (tCh, pSp) - (aCh, pSp)
        const Arg b = convert_inf(a, Ch, CH_T);
        const Classes ka = mbe_classes(c, l, CH_A, b), kt = mbe_classes(c, l, CH_T, b);
        Classes m; m.K1 = -ka.K1; m.K2 = -ka.K2; m.K2p = -ka.K2p; m.K3 = -ka.K3;
        Classes d; d.K1 = 2.0 * kt.K1 + m.K1; d.K2 = 2.0 * kt.K2 + m.K2; d.K2p = 2.0 * kt.K2p + m.K2p; d.K3 = 2.0 * kt.K3 + m.K3;
        val += (mbe_sbe(d, U) - mbe_sbe(m, -U)) * 0.5;
    }
    if (f & FL_GA) val += mbe_sbe(mbe_classes(c, l, CH_A, convert_inf(a, Ch, CH_A)), U);
    return val + mbe_lambda(c, Ch, a);
}

// ---- full vertex, parallel spin, chain from level lev0 ---------------------------------------
// SW = false: P, k, q are Brillouin points.  SW = true: k = q = kSW (P a Brillouin point).
// a.v or a.w may be FDGA_INF (then only the own channel contributes, src/nonlocal/vertex.jl:137-150).
// `flags` (F0 / gamma switches) act on level lev0 only: the reference calls F.F0(...) without
// forwarding them (src/nonlocal/vertex.jl:87-89).
// MBE (compile time): the kernel may meet MBE levels.  Only the generic kernels of MBE contexts are instantiated with MBE = true,
// so that the evaluator of every other kernel carries no trace of the nonlinear path (a run-time branch here cost the hot kernels
// 20-25 %: 920 -> 734 it/s at config 3).
template <bool SW, bool MBE = false> FDGA_HD C eval_p(const DevChain& c, int lev0, int Ch, const Arg& a, unsigned flags);
// parallel spin component of an MBE level; SW: k = q = kSW are explicit averages of the whole (nonlinear) expression over the
// momentum mesh (:481-560); F0 off subtracts the full evaluation of the chain below with the same gamma switches (:418-420)
template <bool SW>
__host__ __device__ inline C eval_p_mbe(const DevChain& c, int l, int Ch, const Arg& a, unsigned flags) {
    const bool below_mbe = c.lev[l + 1].mbe != 0;
    const unsigned g = flags | FL_F0;
    auto one = [&](const Arg& x) -> C {
        C v = mbe_total(c, l, Ch, x, flags);
        if (!(flags & FL_F0)) v = v - (below_mbe ? mbe_total(c, l + 1, Ch, x, flags) : eval_p<false, false>(c, l + 1, Ch, x, g));
        return v;
    };
    if (!SW || c.lev[l].type != LV_NL2) return one(a);
    C s = zeroC();
    for (int iq = 0; iq < c.NP; ++iq) for (int ik = 0; ik < c.NP; ++ik) {
        Arg x = a; x.kx = ik % c.L; x.ky = ik / c.L; x.qx = iq % c.L; x.qy = iq / c.L;
        s += one(x);
    }
    return s / ((double)c.NP * (double)c.NP);
}
template <bool SW, bool MBE>
FDGA_HD C eval_p(const DevChain& c, int lev0, int Ch, const Arg& a, unsigned flags) {
    if (MBE) { if (c.lev[lev0].mbe) return eval_p_mbe<SW>(c, lev0, Ch, a, flags); }
    C val = zeroC();
    const bool anyinf = isinfF(a.v) || isinfF(a.w);
    const int L = c.L, NP = c.NP;
    unsigned f = flags;
    for (int l = lev0; l < c.nlev; ++l) {
        const DevLevel& lv = c.lev[l];
        if (lv.type == LV_CORE) { val += core_eval_px(lv, Ch, SP_P, a.W, a.v, a.w); break; }
        if (anyinf) {
            if (f & (2u << Ch)) {
                if (lv.type == LV_LOCAL) val += loc_chan(lv, Ch, a.W, a.v, a.w);
                else if (SW) val += nl2_chan_sw_own(lv, Ch, NP, a.W, a.v, a.w, kidx(a.Px, a.Py, L));
                else val += nl2_chan(lv, Ch, NP, a.W, a.v, a.w, kidx(a.Px, a.Py, L), kidx(a.kx, a.ky, L), kidx(a.qx, a.qy, L));
            }
        } else {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                if (!(f & (2u << r))) continue;
                if (lv.type == LV_LOCAL) {
                    Arg b = convert(a, Ch, r);
                    val += loc_chan(lv, r, b.W, b.v, b.w);
                } else if (SW) {
                    if (r == Ch) val += nl2_chan_sw_own(lv, r, NP, a.W, a.v, a.w, kidx(a.Px, a.Py, L));
                    else { Arg b = convert(a, Ch, r); val += nl2_chan_sw_cross(lv, r, b.W, b.v, b.w); }
                } else {
                    Arg b = convert(a, Ch, r);
                    val += nl2_chan(lv, r, NP, b.W, b.v, b.w, kidx(b.Px, b.Py, L), kidx(b.kx, b.ky, L), kidx(b.qx, b.qy, L));
                }
            }
        }
        if (!(f & FL_F0)) break;
        f = FL_ALL;
    }
    return val;
}

// crossed spin: src/nonlocal/vertex.jl:157-185 (gamma_t / gamma_a switches swapped)
FDGA_HD unsigned swap_ta(unsigned f) {
    return (f & (FL_F0 | FL_GP)) | ((f & FL_GT) ? FL_GA : 0u) | ((f & FL_GA) ? FL_GT : 0u);
}

// generic entry: chain from lev0, any spin.  Sp, Ch are compile-time constants at all call sites.
template <bool SW, bool MBE = false>
FDGA_HD C eval_vertex(const DevChain& c, int lev0, int Ch, int Sp, const Arg& a, unsigned flags) {
    if (c.lev[lev0].type == LV_CORE) return core_eval(c.lev[lev0], Ch, Sp, a.W, a.v, a.w);
    C val = zeroC();
    if (Sp == SP_P || Sp == SP_D) {
        C p = eval_p<SW, MBE>(c, lev0, Ch, a, flags);
        val = (Sp == SP_D) ? p * 2.0 : p;
        if (Sp == SP_P) return val;
    }
    // xSp
    Arg b = a;
    int Ch2;
    if (Ch == CH_P) {
        Ch2 = CH_P;
        b.w = isinfF(a.w) ? FDGA_INF : a.W - a.w - 1;          // W - w'
        b.qx = a.Px - a.qx; b.qy = a.Py - a.qy;               // P - q  (kSW stays kSW)
    } else if (Ch == CH_T) Ch2 = CH_A;
    else Ch2 = CH_T;
    C x = -eval_p<SW, MBE>(c, lev0, Ch2, b, swap_ta(flags));
    return val + x;
}

}  // namespace fdga
