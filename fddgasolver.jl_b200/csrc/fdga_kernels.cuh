// fdga_kernels.cuh -- CUDA kernels of the BSE / cache / bubble / SDE hot path (sm_100a).
// Every kernel cites the reference function whose arithmetic it reproduces.
#pragma once
#include "fdga_device.cuh"

namespace fdga {

// Grid shape shared by all kernels
struct Grid {
    double T;
    int L, NP;            // vertex / bubble momentum mesh
    int nPiB, nPiF;       // N of the bubble meshes
    int nK1, nK2b, nK2f, nK3b, nK3f;   // level-0 (S.F) vertex meshes = output meshes
    int LG, nG;           // G / Sigma mesh
};

struct SymDev {            // symmetry classes (CSR); representative = first member
    long long ncls, nmem;
    const long long* offsets;
    const long long* index;
    const unsigned char* ops;
    const int* member_class;
};

__device__ __forceinline__ C apply_op(unsigned char op, C v) {
    if (op & 2) v = conjC(v);
    if (op & 1) v = -v;
    return v;
}

// ---- block reduction of a complex value (deterministic order) ---------------------------------
__device__ __forceinline__ C block_reduce(C v) {
    __shared__ double sx[32], sy[32];
    for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_down_sync(0xffffffffu, v.x, o);
        v.y += __shfl_down_sync(0xffffffffu, v.y, o);
    }
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) { sx[wid] = v.x; sy[wid] = v.y; }
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    if (wid == 0) {
        v.x = lane < nw ? sx[lane] : 0.0;
        v.y = lane < nw ? sy[lane] : 0.0;
        for (int o = 16; o > 0; o >>= 1) {
            v.x += __shfl_down_sync(0xffffffffu, v.x, o);
            v.y += __shfl_down_sync(0xffffffffu, v.y, o);
        }
    }
    return v;   // valid in thread 0
}

// ---- elementwise helpers -----------------------------------------------------------------------
// out = a*x + b*y  (y may be null -> a*x);  used for the t-channel post-fix (BSE_templates.jl:35-38)
__global__ void axpby_kernel(C* __restrict__ out, const C* x, double a, const C* y, double b, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    C v = x[i] * a;
    if (y) v += y[i] * b;
    out[i] = v;
}
// out += a*x + b*y
__global__ void add_axpby_kernel(C* __restrict__ out, const C* x, double a, const C* y, double b, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    C v = out[i] + x[i] * a;
    if (y) v += y[i] * b;
    out[i] = v;
}
__global__ void scale_copy_kernel(C* __restrict__ out, const C* __restrict__ x, double s, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = x[i] * s;
}
// y = x - F/factor (mfRG linear map, src/mfRG.jl:85-86)
__global__ void mfrg_residual_kernel(C* __restrict__ y, const C* __restrict__ x, const C* __restrict__ F, double factor, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) y[i] = x[i] - F[i] / factor;
}

// SG expansion: out[index[j]] = op_j(repvals[class(j)])   (MatsubaraFunctions SymmetryGroup call)
__global__ void expand_kernel(C* __restrict__ out, const C* __restrict__ repvals, SymDev sg) {
    long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= sg.nmem) return;
    out[sg.index[j]] = apply_op(sg.ops[j], repvals[sg.member_class[j]]);
}
// accumulate variant: out[index[j]] += w * op_j(repvals[class(j)])
__global__ void expand_add_kernel(C* __restrict__ out, const C* __restrict__ repvals, SymDev sg, double w) {
    long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= sg.nmem) return;
    out[sg.index[j]] += apply_op(sg.ops[j], repvals[sg.member_class[j]]) * w;
}
// ---- batched SG finish of one BSE stage -----------------------------------------------------------------------------------
// All post-processing that follows the SG(...) fills of a stage (BSE_templates.jl:35-38, 72-73, 107-108, 142-143, 176-177;
// BSEa_K2.jl:128-135) is linear and commutes with the class expansion (every array involved is symmetric under the SAME group
// with the same operations), so it is applied to the class REPRESENTATIVES -- a few thousand numbers instead of the full arrays:
//     a, p :  rep_r += FL_r[index of the representative]                 (K2 only: the FL add)
//     t    :  rep_t <- ( rep_t + [2 FL_t - FL_a][rep index] + rep_a ) / 2  (spin d -> p with the FINAL a-channel value)
// One launch handles every kernel class of the stage (blockIdx.y); then ONE launch expands all arrays (expand_multi_kernel).
struct RepFixJob { C* rep[3]; const C* fl[3]; const long long* offsets[3]; const long long* index[3]; long long ncls[3]; };
#define FDGA_MAXFIX 4
struct RepFixJobs { RepFixJob j[FDGA_MAXFIX]; };
__global__ void repfix_kernel(const __grid_constant__ RepFixJobs jobs) {
    const RepFixJob& J = jobs.j[blockIdx.y];
    const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (c < J.ncls[CH_P] && J.fl[CH_P])                              // the p channel lives in its own (pp) group
        J.rep[CH_P][c] += J.fl[CH_P][J.index[CH_P][J.offsets[CH_P][c]]];
    if (c >= J.ncls[CH_A]) return;                                   // a and t share the ph group
    const long long idx = J.fl[CH_A] ? J.index[CH_A][J.offsets[CH_A][c]] : 0;
    C a = J.rep[CH_A][c];
    if (J.fl[CH_A]) { a += J.fl[CH_A][idx]; J.rep[CH_A][c] = a; }
    C t = J.rep[CH_T][c];
    if (J.fl[CH_T]) t += J.fl[CH_T][idx] * 2.0 - J.fl[CH_A][idx];
    J.rep[CH_T][c] = (t + a) * 0.5;
}
struct ExpandJob { C* out; const C* rep; SymDev sg; };
#define FDGA_MAXEXP 12
struct ExpandJobs { ExpandJob j[FDGA_MAXEXP]; };
__global__ void expand_multi_kernel(const __grid_constant__ ExpandJobs jobs) {
    const ExpandJob& J = jobs.j[blockIdx.y];
    for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < J.sg.nmem; m += (long long)gridDim.x * blockDim.x)
        J.out[J.sg.index[m]] = apply_op(J.sg.ops[m], J.rep[J.sg.member_class[m]]);
}

// SG(f): symmetrise in place from the representatives
__global__ void symmetrize_kernel(C* __restrict__ f, SymDev sg) {
    long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= sg.nmem) return;
    long long c = sg.member_class[j];
    long long rep = sg.offsets[c];
    if (j == rep) return;
    f[sg.index[j]] = apply_op(sg.ops[j], f[sg.index[rep]]);
}

// ---- s-wave tables (src/nonlocal/swave.jl:32-136): BZ means of the three channels of one NL2 level --------------
struct SwOut { C* p[3][4]; };    // [channel][K1sw, K2swk, K2sww, K3sw]
__global__ void swave_tables_kernel(DevLevel lv, int NP, SwOut out) {
    const int r = blockIdx.y;
    const DevChan& c = lv.ch[r];
    int nB1 = 2 * lv.nK1 - 1, nB2 = 2 * lv.nK2b - 1, nF2 = 2 * lv.nK2f, nB3 = 2 * lv.nK3b - 1, nF3 = 2 * lv.nK3f;
    long long n1 = nB1, n2k = (long long)nB2 * nF2 * NP, n3 = (long long)nB3 * nF3 * nF3;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n1) {
        C s = zeroC();
        for (int p = 0; p < NP; ++p) s += c.K1[i + (size_t)nB1 * p];
        out.p[r][0][i] = s / (double)NP;
        return;
    }
    i -= n1;
    if (i < n2k) {      // mean over the 4th axis: index (W,v,P) contiguous
        C s = zeroC();
        size_t sk = (size_t)nB2 * nF2 * NP;
        for (int k = 0; k < NP; ++k) s += c.K2[i + sk * k];
        out.p[r][1][i] = s / (double)NP;
        return;
    }
    i -= n2k;
    if (i < n3) {
        C s = zeroC();
        for (int p = 0; p < NP; ++p) s += c.K3[i + (size_t)n3 * p];
        out.p[r][3][i] = s / (double)NP;
    }
}
// K2[W, v, kSW, kSW] = sum(view(f, i1, i2, :, :)) / N3 / N4, taken as the P-mean of the k-means (differs by rounding only)
__global__ void swave_tables2_kernel(DevLevel lv, int NP, SwOut out) {
    const int r = blockIdx.y;
    int nB2 = 2 * lv.nK2b - 1, nF2 = 2 * lv.nK2f;
    long long n2w = (long long)nB2 * nF2;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n2w) return;
    C s = zeroC();
    for (int p = 0; p < NP; ++p) s += out.p[r][1][i + n2w * p];
    out.p[r][2][i] = s / (double)NP;
}

// ---- bubble auxiliaries: PiT[q,w,W,P] (slab-contiguous copy, inner momentum fastest) and Pisw[W,w,P] = mean_k Pi[W,w,P,k] ----
// Slab layout used by every contraction kernel: element (w, q) of the slab (W, P) sits at  q + NP * w  (q fastest), so that a
// warp whose lanes run over the inner momentum q reads one contiguous run per inner frequency.
FDGA_HD size_t slab_at(int iw, int iq, int NP) { return (size_t)iq + (size_t)NP * iw; }
// Slab storage is COMPACT: a rank only holds the (W, P) slabs that carry one of its class representatives, in the order of its
// slab list; `map` translates (position of W in the bosonic mesh, P) into the position in that list (null: all slabs, natural order).
FDGA_HD size_t slab_index(const int* __restrict__ map, int iWo, int nBo, int iP) {
    const int j = iWo + nBo * iP;
    return map ? (size_t)map[j] : (size_t)j;
}
// slabs (list entries (iW, iP)) of a bubble given in the reference's layout [W, v, P, k]
__global__ void pi_gather_slabs_kernel(const C* __restrict__ Pi, C* __restrict__ PiT, int nB, int nF, int NP,
                                       const int4* __restrict__ slabs, int nslabs) {
    // one thread per output element, output index contiguous
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long n = (long long)nF * NP * nslabs;
    if (i >= n) return;
    long long t = i;
    int iq = t % NP; t /= NP; int iw = t % nF; int sl = (int)(t / nF);
    const int iW = slabs[sl].x, iP = slabs[sl].y;
    PiT[i] = Pi[iW + (size_t)nB * (iw + (size_t)nF * (iP + (size_t)NP * iq))];
}
__global__ void pi_swave_kernel(const C* __restrict__ Pi, C* __restrict__ Pisw, int nB, int nF, int NP) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long n = (long long)nB * nF * NP;
    if (i >= n) return;
    C s = zeroC();
    for (int k = 0; k < NP; ++k) s += Pi[i + (size_t)n * k];
    Pisw[i] = s / (double)NP;
}

// ---- right factor, hoisted out of the (nu, k) loops ---------------------------------
//  RK_FD    : (Pi - Pi0) * F0(W, w~, inf; P, q~, k0) + Pi * FL(W, w~, inf; P, q~, k0)   BSEa_K1.jl:41-47, BSEa_K2.jl:112-118
//  RK_MF_K1 : Pi0 * FL(W, w~, inf; P, q~, k0)                                           BSEa_K1.jl:33-37
//  RK_MF_K2 : Pi0 * FL(W, w, inf; P, q, k0)                                             BSEa_K2.jl:100-104
//  RK_LK2   : Pi0 * F0(W, w, inf; P, q, k0), w on the K2 nu-mesh                        BSEa_K2.jl:38-41
// Rt layout: [iq + NP*(iw + nw*slab)] (slab_at), slab = position of (W, P) in the list; W on the OUTPUT bosonic mesh (N = No).
//  RK_LK2_LOC : Pi0 * F0(W, w~, inf), w on the bubble nu-mesh (local solver)                src/BSEa/BSEa_K2.jl:27-30
//  RK_1L    : (Pi - Pi0) * F0(W, w~, inf; P, q~, k0)  (fd branch of the 1-loop variants)   BSE_1loop.jl:41-46,104-109
enum { RK_FD = 0, RK_MF_K1 = 1, RK_MF_K2 = 2, RK_LK2 = 3, RK_LK2_LOC = 4, RK_1L = 5 };

// Only the slabs (W, P) that hold class representatives of this rank are filled: `slabs` lists (iWo, iP) pairs.
template <int CH, int KIND, bool MBE = false>
__global__ void right_factor_kernel(const __grid_constant__ DevChain F0, const __grid_constant__ DevChain FL,
                                    const C* __restrict__ Pi0T, const C* __restrict__ PiT, C* __restrict__ Rt,
                                    Grid g, int No, int Ninner, const int4* __restrict__ slabs, int nslabs, const int* __restrict__ pimap) {
    constexpr int SP = (CH == CH_T) ? SP_D : SP_P;
    const int nw = 2 * Ninner, nBo = 2 * No - 1, nFP = 2 * g.nPiF, nBP = 2 * g.nPiB - 1;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long n = (long long)nw * g.NP * nslabs;
    if (i >= n) return;
    long long t = i;
    int iq = t % g.NP; t /= g.NP; int iw = t % nw; int sl = (int)(t / nw);
    const int4 s2 = slabs[sl];
    const int iWo = s2.x, iP = s2.y;
    int W = iWo - (No - 1), w = iw - Ninner;
    int Px = iP % g.L, Py = iP / g.L, qx = iq % g.L, qy = iq / g.L;
    size_t pidx = slab_at(posF(w, g.nPiF), iq, g.NP) + (size_t)nFP * g.NP * slab_index(pimap, posB(W, g.nPiB), nBP, iP);
    Arg a;
    a.W = W; a.w = FDGA_INF; a.Px = Px; a.Py = Py; a.qx = 0; a.qy = 0;
    if (KIND == RK_FD || KIND == RK_MF_K1 || KIND == RK_LK2_LOC || KIND == RK_1L) {   // crossed arguments (_crossing, BSE_templates.jl:4-6)
        a.v = (CH == CH_P) ? W - w - 1 : w;
        a.kx = (CH == CH_P) ? Px - qx : qx; a.ky = (CH == CH_P) ? Py - qy : qy;
    } else {
        a.v = w; a.kx = qx; a.ky = qy;
    }
    C r;
    if (KIND == RK_FD) {
        C F0r = eval_vertex<false, MBE>(F0, 0, CH, SP, a, FL_ALL);
        C FLr = eval_vertex<false, MBE>(FL, 0, CH, SP, a, FL_ALL);
        C p = PiT[pidx], p0 = Pi0T[pidx];
        r = (p - p0) * F0r + p * FLr;
    } else if (KIND == RK_1L) {
        r = (PiT[pidx] - Pi0T[pidx]) * eval_vertex<false, MBE>(F0, 0, CH, SP, a, FL_ALL);
    } else if (KIND == RK_LK2 || KIND == RK_LK2_LOC) {
        r = Pi0T[pidx] * eval_vertex<false, MBE>(F0, 0, CH, SP, a, FL_ALL);
    } else {
        r = Pi0T[pidx] * eval_vertex<false, MBE>(FL, 0, CH, SP, a, FL_ALL);
    }
    (void)nBo;
    Rt[slab_at(iw, iq, g.NP) + (size_t)nw * g.NP * sl] = r;      // compact: slab number = position in the list
}

// ---- BSE_K1!: src/nonlocal_2/BSEa/BSEa_K1.jl:19-52.  One CTA per class representative (W, P) ------
template <int CH, bool MBE = false>
__global__ void bse_k1_kernel(const __grid_constant__ DevChain Fleft, const C* __restrict__ Rt, C* __restrict__ repvals,
                              SymDev sg, long long c0, Grid g, double scale, const int* __restrict__ map) {
    constexpr int SP = (CH == CH_T) ? SP_D : SP_P;
    long long cls = c0 + blockIdx.x;
    long long idx = sg.index[sg.offsets[cls]];
    const int nB1 = 2 * g.nK1 - 1, nw = 2 * g.nPiF;
    int iW = idx % nB1, iP = idx / nB1;
    int W = iW - (g.nK1 - 1), Px = iP % g.L, Py = iP / g.L;
    const C* slab = Rt + (size_t)nw * g.NP * slab_index(map, iW, nB1, iP);
    C acc = zeroC();
    for (int t = threadIdx.x; t < nw * g.NP; t += blockDim.x) {
        int iq = t % g.NP, iw = t / g.NP;
        Arg a; a.W = W; a.v = FDGA_INF; a.w = iw - g.nPiF; a.Px = Px; a.Py = Py; a.kx = 0; a.ky = 0; a.qx = iq % g.L; a.qy = iq / g.L;
        C Fl = eval_vertex<false, MBE>(Fleft, 0, CH, SP, a, FL_ALL);
        acc += Fl * slab[t];
    }
    acc = block_reduce(acc);
    if (threadIdx.x == 0) repvals[cls] = acc * scale;
}

// ---- BSE_L_K2!: src/nonlocal_2/BSEa/BSEa_K2.jl:17-43.  One CTA per representative (W, v, P, k) -----
template <int CH, bool MBE = false>
__global__ void bse_lk2_kernel(const __grid_constant__ DevChain F, const C* __restrict__ Rt, C* __restrict__ repvals,
                               SymDev sg, long long c0, Grid g, double scale, const int* __restrict__ map) {
    constexpr int SP = (CH == CH_T) ? SP_D : SP_P;
    constexpr unsigned FLG = (CH == CH_P ? 0u : FL_GP) | (CH == CH_T ? 0u : FL_GT) | (CH == CH_A ? 0u : FL_GA);
    long long cls = c0 + blockIdx.x;
    long long idx = sg.index[sg.offsets[cls]];
    const int nB2 = 2 * g.nK2b - 1, nF2 = 2 * g.nK2f, nw = nF2;
    long long t0 = idx;
    int iW = t0 % nB2; t0 /= nB2; int iv = t0 % nF2; t0 /= nF2; int iP = t0 % g.NP; int ik = t0 / g.NP;
    int W = iW - (g.nK2b - 1), v = iv - g.nK2f, Px = iP % g.L, Py = iP / g.L, kx = ik % g.L, ky = ik / g.L;
    const C* slab = Rt + (size_t)nw * g.NP * slab_index(map, iW, nB2, iP);
    C acc = zeroC();
    for (int t = threadIdx.x; t < nw * g.NP; t += blockDim.x) {
        int iq = t % g.NP, iw = t / g.NP;
        int w = iw - g.nK2f, qx = iq % g.L, qy = iq / g.L;
        Arg a; a.W = W; a.v = v; a.Px = Px; a.Py = Py; a.kx = kx; a.ky = ky;
        a.w = (CH == CH_P) ? W - w - 1 : w;
        a.qx = (CH == CH_P) ? Px - qx : qx; a.qy = (CH == CH_P) ? Py - qy : qy;
        C Gl = eval_vertex<false, MBE>(F, 0, CH, SP, a, FLG);
        acc += Gl * slab[t];
    }
    acc = block_reduce(acc);
    if (threadIdx.x == 0) repvals[cls] = acc * scale;
}

// ---- BSE_K2!: src/nonlocal_2/BSEa/BSEa_K2.jl:72-125.  One CTA per representative (W, v, P, k) ------
//  fd   : [F(W,v,w;P,k,q) - F(W,inf,w;P,k,q)] * Rt                 (Fleft = S.F)
//  mfRG : [F0(W,v,w~;P,k,q~) - F0(W,inf,w~;P,k,q~)] * Rt           (Fleft = S.F0)
template <int CH, bool MF, bool MBE = false>
__global__ void bse_k2_kernel(const __grid_constant__ DevChain Fleft, const C* __restrict__ Rt, C* __restrict__ repvals,
                              SymDev sg, long long c0, Grid g, double scale, const int* __restrict__ map) {
    constexpr int SP = (CH == CH_T) ? SP_D : SP_P;
    long long cls = c0 + blockIdx.x;
    long long idx = sg.index[sg.offsets[cls]];
    const int nB2 = 2 * g.nK2b - 1, nF2 = 2 * g.nK2f, nw = 2 * g.nPiF;
    long long t0 = idx;
    int iW = t0 % nB2; t0 /= nB2; int iv = t0 % nF2; t0 /= nF2; int iP = t0 % g.NP; int ik = t0 / g.NP;
    int W = iW - (g.nK2b - 1), v = iv - g.nK2f, Px = iP % g.L, Py = iP / g.L, kx = ik % g.L, ky = ik / g.L;
    const C* slab = Rt + (size_t)nw * g.NP * slab_index(map, iW, nB2, iP);
    C acc = zeroC();
    for (int t = threadIdx.x; t < nw * g.NP; t += blockDim.x) {
        int iq = t % g.NP, iw = t / g.NP;
        int w = iw - g.nPiF, qx = iq % g.L, qy = iq / g.L;
        Arg a; a.W = W; a.v = v; a.Px = Px; a.Py = Py; a.kx = kx; a.ky = ky;
        if (MF) {
            a.w = (CH == CH_P) ? W - w - 1 : w;
            a.qx = (CH == CH_P) ? Px - qx : qx; a.qy = (CH == CH_P) ? Py - qy : qy;
        } else { a.w = w; a.qx = qx; a.qy = qy; }
        C f1 = eval_vertex<false, MBE>(Fleft, 0, CH, SP, a, FL_ALL);
        a.v = FDGA_INF;
        C f2 = eval_vertex<false, MBE>(Fleft, 0, CH, SP, a, FL_ALL);
        acc += (f1 - f2) * slab[t];
    }
    acc = block_reduce(acc);
    if (threadIdx.x == 0) repvals[cls] = acc * scale;
}

// ---- BSE_K1_new!: src/nonlocal_2/BSEa/BSEa_K1.jl:62-113, K1 = (U + K1 + K2') Pi U.  One CTA per representative (W, P).
//  fd   : [F(W,inf,w;P,k0,q) Pi - F0(W,inf,w;P,k0,q) Pi0] U          mfRG : [F - F0] Pi U
//  PiT / Pi0T are the transposed bubbles [q, w | W, P] (W on the bubble mesh, which is the K1 mesh).
template <int CH>
__global__ void bse_k1_new_kernel(const __grid_constant__ DevChain F, const __grid_constant__ DevChain F0,
                                  const C* __restrict__ Pi0T, const C* __restrict__ PiT, C* __restrict__ repvals,
                                  SymDev sg, long long c0, Grid g, C scaleU, int mfrg, const int* __restrict__ map) {
    constexpr int SP = (CH == CH_T) ? SP_D : SP_P;
    long long cls = c0 + blockIdx.x;
    long long idx = sg.index[sg.offsets[cls]];
    const int nB1 = 2 * g.nK1 - 1, nw = 2 * g.nPiF, nBP = 2 * g.nPiB - 1;
    int iW = idx % nB1, iP = idx / nB1;
    int W = iW - (g.nK1 - 1), Px = iP % g.L, Py = iP / g.L;
    const size_t off = (size_t)nw * g.NP * slab_index(map, posB(W, g.nPiB), nBP, iP);
    C acc = zeroC();
    for (int t = threadIdx.x; t < nw * g.NP; t += blockDim.x) {
        int iq = t % g.NP, iw = t / g.NP;
        Arg a; a.W = W; a.v = FDGA_INF; a.w = iw - g.nPiF; a.Px = Px; a.Py = Py; a.kx = 0; a.ky = 0; a.qx = iq % g.L; a.qy = iq / g.L;
        C Fl = eval_vertex<false>(F, 0, CH, SP, a, FL_ALL);
        C F0l = eval_vertex<false>(F0, 0, CH, SP, a, FL_ALL);
        if (mfrg) acc += (Fl - F0l) * PiT[off + t];
        else      acc += Fl * PiT[off + t] - F0l * Pi0T[off + t];
    }
    acc = block_reduce(acc);
    if (threadIdx.x == 0) repvals[cls] = acc * scaleU;
}

// ---- BSE_K2_new!: src/nonlocal_2/BSEa/BSEa_K2.jl:142-216.  One CTA per representative (W, v, P, k); the inner
// frequency runs over the K2 nu-mesh only (BSEa_K2.jl:189-193).
template <int CH>
__global__ void bse_k2_new_kernel(const __grid_constant__ DevChain F, const __grid_constant__ DevChain F0,
                                  const C* __restrict__ Pi0T, const C* __restrict__ PiT, C* __restrict__ repvals,
                                  SymDev sg, long long c0, Grid g, C scaleU, int mfrg, const int* __restrict__ map) {
    constexpr int SP = (CH == CH_T) ? SP_D : SP_P;
    long long cls = c0 + blockIdx.x;
    long long idx = sg.index[sg.offsets[cls]];
    const int nB2 = 2 * g.nK2b - 1, nF2 = 2 * g.nK2f, nFP = 2 * g.nPiF, nBP = 2 * g.nPiB - 1;
    long long t0 = idx;
    int iW = t0 % nB2; t0 /= nB2; int iv = t0 % nF2; t0 /= nF2; int iP = t0 % g.NP; int ik = t0 / g.NP;
    int W = iW - (g.nK2b - 1), v = iv - g.nK2f, Px = iP % g.L, Py = iP / g.L, kx = ik % g.L, ky = ik / g.L;
    const size_t off = (size_t)nFP * g.NP * slab_index(map, posB(W, g.nPiB), nBP, iP);
    C acc = zeroC();
    for (int t = threadIdx.x; t < nF2 * g.NP; t += blockDim.x) {
        int iq = t % g.NP, iw = t / g.NP;
        int w = iw - g.nK2f;
        Arg a; a.W = W; a.v = v; a.w = w; a.Px = Px; a.Py = Py; a.kx = kx; a.ky = ky; a.qx = iq % g.L; a.qy = iq / g.L;
        C Fl = eval_vertex<false>(F, 0, CH, SP, a, FL_ALL), F0l = eval_vertex<false>(F0, 0, CH, SP, a, FL_ALL);
        a.v = FDGA_INF;
        Fl = Fl - eval_vertex<false>(F, 0, CH, SP, a, FL_ALL); F0l = F0l - eval_vertex<false>(F0, 0, CH, SP, a, FL_ALL);
        size_t pidx = off + slab_at(posF(w, g.nPiF), iq, g.NP);
        if (mfrg) acc += (Fl - F0l) * PiT[pidx];
        else      acc += Fl * PiT[pidx] - F0l * Pi0T[pidx];
    }
    acc = block_reduce(acc);
    if (threadIdx.x == 0) repvals[cls] = acc * scaleU;
}

// ---- K3 index helper -----------------------------------------------------------------------------
__device__ __forceinline__ size_t k3at(const Grid& g, int W, int v, int w, int iP) {
    int nB = 2 * g.nK3b - 1, nF = 2 * g.nK3f;
    return posB(W, g.nK3b) + (size_t)nB * (posF(v, g.nK3f) + (size_t)nF * (posF(w, g.nK3f) + (size_t)nF * iP));
}
__device__ __forceinline__ size_t piswat(const Grid& g, int W, int w, int iP) {
    return posB(W, g.nPiB) + (size_t)(2 * g.nPiB - 1) * (posF(w, g.nPiF) + (size_t)(2 * g.nPiF) * iP);
}
__device__ __forceinline__ void decode_k3(const Grid& g, long long idx, int& W, int& v, int& w, int& iP) {
    int nB = 2 * g.nK3b - 1, nF = 2 * g.nK3f;
    int iW = idx % nB; idx /= nB; int iv = idx % nF; idx /= nF; int iw = idx % nF; iP = (int)(idx / nF);
    W = iW - (g.nK3b - 1); v = iv - g.nK3f; w = iw - g.nK3f;
}

// ---- BSE_L_K3!: src/nonlocal_2/BSEa/BSEa_K3.jl:19-34.  One thread per representative ------------------
__global__ void bse_lk3_kernel(const C* __restrict__ cache_G, const C* __restrict__ cache_F0, const C* __restrict__ Pi0sw,
                               C* __restrict__ repvals, SymDev sg, long long c0, long long c1, Grid g, double scale) {
    long long cls = c0 + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (cls >= c1) return;
    long long idx = sg.index[sg.offsets[cls]];
    int W, v, vp, iP; decode_k3(g, idx, W, v, vp, iP);
    C val = zeroC();
    for (int w = -g.nK3f; w < g.nK3f; ++w)
        val += cache_G[k3at(g, W, v, w, iP)] * Pi0sw[piswat(g, W, w, iP)] * cache_F0[k3at(g, W, w, vp, iP)];
    repvals[cls] = val * scale;
}

// ---- BSE_K3!: src/nonlocal_2/BSEa/BSEa_K3.jl:62-121.  One thread per representative -----------------
// ONELOOP: BSE_K3_1loop! (src/nonlocal_2/BSEa/BSE_1loop.jl:123-199): fd branch keeps the (Pi - Pi0) F0 term only;
// neither branch adds FL.K3 to the result.
template <int CH, bool MF, bool ONELOOP = false>
__global__ void bse_k3_kernel(const C* __restrict__ FLown, const C* __restrict__ FLt, const C* __restrict__ FLa,
                              const C* __restrict__ cache_G, const C* __restrict__ cache_F, const C* __restrict__ cache_F0,
                              const C* __restrict__ Pi0sw, const C* __restrict__ Pisw, C* __restrict__ repvals,
                              SymDev sg, long long c0, long long c1, Grid g, double sign1, double sign2) {
    long long cls = c0 + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (cls >= c1) return;
    long long idx = sg.index[sg.offsets[cls]];
    int W, v, vp, iP; decode_k3(g, idx, W, v, vp, iP);
    C val = zeroC();
    for (int w = -g.nK3f; w < g.nK3f; ++w) {
        C Gs = (CH == CH_P) ? cache_G[k3at(g, W, w, vp, iP)] : cache_G[k3at(g, W, vp, w, iP)];
        C Fs = cache_F[k3at(g, W, v, w, iP)];
        C P0 = Pi0sw[piswat(g, W, w, iP)];
        int wc = (CH == CH_P) ? W - w - 1 : w;
        bool cin = inF(wc, g.nK3f);
        C cen = zeroC();
        if (cin) cen = (CH == CH_T) ? (2.0 * FLt[k3at(g, W, wc, vp, iP)] - FLa[k3at(g, W, wc, vp, iP)]) : FLown[k3at(g, W, wc, vp, iP)];
        if (MF) {
            val += Fs * P0 * Gs * sign1;
            if (cin) val += Fs * P0 * cen * sign2;
        } else if (ONELOOP) {
            C P1 = Pisw[piswat(g, W, w, iP)];
            C F0s = cache_F0[k3at(g, W, w, vp, iP)];
            val += Fs * (P1 - P0) * F0s * sign1;
        } else {
            C P1 = Pisw[piswat(g, W, w, iP)];
            C F0s = cache_F0[k3at(g, W, w, vp, iP)];
            val += Fs * ((P1 - P0) * F0s + P1 * Gs) * sign1;
            if (cin) val += Fs * P1 * cen * sign2;
        }
    }
    if (ONELOOP) { repvals[cls] = val * g.T; return; }
    C add = (CH == CH_T) ? (2.0 * FLt[k3at(g, W, v, vp, iP)] - FLa[k3at(g, W, v, vp, iP)]) : FLown[k3at(g, W, v, vp, iP)];
    repvals[cls] = val * g.T + add;
}

// ---- build_K3_cache!: src/nonlocal_2/build_K3_cache.jl:35-84.  One thread per K3 grid point ---------
struct CachePtrs { C* c[10]; };
__global__ void build_cache_kernel(const __grid_constant__ DevChain F0, const __grid_constant__ DevChain F, CachePtrs out,
                                   Grid g, long long i0, long long i1) {
    long long i = i0 + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= i1) return;
    int W, a1, b1, iP; decode_k3(g, i, W, a1, b1, iP);
    Arg a; a.W = W; a.v = a1; a.w = b1; a.Px = iP % g.L; a.Py = iP / g.L; a.kx = a.ky = a.qx = a.qy = 0;
    Arg ai = a; ai.w = FDGA_INF;      // second frequency -> infinity
    Arg av = a; av.v = FDGA_INF;      // first frequency -> infinity
    // (W, w, v') block
    out.c[0][i] = eval_vertex<true>(F, 0, CH_P, SP_X, a, FL_GT | FL_GA);
    out.c[1][i] = eval_vertex<true>(F0, 0, CH_P, SP_X, a, FL_ALL) - eval_vertex<true>(F0, 0, CH_P, SP_X, ai, FL_ALL);
    C f0a = eval_vertex<true>(F0, 0, CH_A, SP_P, a, FL_ALL) - eval_vertex<true>(F0, 0, CH_A, SP_P, ai, FL_ALL);
    C f0t = eval_vertex<true>(F0, 0, CH_T, SP_P, a, FL_ALL) - eval_vertex<true>(F0, 0, CH_T, SP_P, ai, FL_ALL);
    out.c[2][i] = f0a;
    out.c[3][i] = 2.0 * f0t - f0a;
    // (W, v, w) block
    C gpp = eval_vertex<true>(F, 0, CH_P, SP_P, a, FL_GT | FL_GA);
    C ga  = eval_vertex<true>(F, 0, CH_A, SP_P, a, FL_GP | FL_GT);
    C gt  = eval_vertex<true>(F, 0, CH_T, SP_P, a, FL_GP | FL_GA);
    C fp = (eval_vertex<true>(F, 0, CH_P, SP_P, a, FL_F0 | FL_GP) - eval_vertex<true>(F, 0, CH_P, SP_P, av, FL_F0 | FL_GP)) + gpp;
    C fa = (eval_vertex<true>(F, 0, CH_A, SP_P, a, FL_F0 | FL_GA) - eval_vertex<true>(F, 0, CH_A, SP_P, av, FL_F0 | FL_GA)) + ga;
    C ft = (eval_vertex<true>(F, 0, CH_T, SP_P, a, FL_F0 | FL_GT) - eval_vertex<true>(F, 0, CH_T, SP_P, av, FL_F0 | FL_GT)) + gt;
    out.c[4][i] = gpp;
    out.c[5][i] = ga;
    out.c[6][i] = gt * 2.0 - ga;
    out.c[7][i] = fp;
    out.c[8][i] = fa;
    out.c[9][i] = ft * 2.0 - fa;
}

// ---- build_K3_cache! for MBE vertices: src/boson_exchange.jl:947-1017 (NL2_MBEVertex), :599-668 (MBEVertex = the same on a 1 x 1
// mesh).  cache_F0* / cache_F*: channel-U-irreducible vertex T_r = (I_r - U) + M_r; cache_G*: F - F.F0.  The kSW evaluations of a
// nonlinear vertex are explicit N_P^2 averages (eval_p_mbe<true>), as in the reference: a small-mesh path. ----
// One CTA per K3 grid point: the threads share the (k, q) pairs of the s-wave averages, 11 partial sums are block-reduced.
__global__ void build_cache_mbe_kernel(const __grid_constant__ DevChain F0, const __grid_constant__ DevChain F, CachePtrs out,
                                       Grid g, long long i0, long long i1) {
    const long long i = i0 + blockIdx.x;
    if (i >= i1) return;
    int W, a1, b1, iP; decode_k3(g, i, W, a1, b1, iP);
    Arg a; a.W = W; a.v = a1; a.w = b1; a.Px = iP % g.L; a.Py = iP / g.L; a.kx = a.ky = a.qx = a.qy = 0;
    const C U = F.lev[F.nlev - 1].U;
    constexpr unsigned NP_ = FL_F0 | FL_GT | FL_GA, NA_ = FL_F0 | FL_GP | FL_GT, NT_ = FL_F0 | FL_GP | FL_GA;      // gamma_r = false
    // the reference chain is momentum independent (a RefVertex, or local levels only): one evaluation instead of N_P^2
    bool f0_flat = true;
    for (int l = 0; l < F0.nlev; ++l) f0_flat = f0_flat && F0.lev[l].type != LV_NL2;
    C s[11];
#pragma unroll
    for (int j = 0; j < 11; ++j) s[j] = zeroC();
    const int npair = g.NP * g.NP;
    for (int t = threadIdx.x; t < npair; t += blockDim.x) {
        Arg x = a; const int ik = t % g.NP, iq = t / g.NP;
        x.kx = ik % g.L; x.ky = ik / g.L; x.qx = iq % g.L; x.qy = iq / g.L;
        s[0] += eval_vertex<false, true>(F, 0, CH_P, SP_X, x, NP_) - eval_vertex<false, true>(F, 1, CH_P, SP_X, x, NP_);
        if (!f0_flat) {
            s[1] += eval_vertex<false, true>(F0, 0, CH_P, SP_X, x, NP_);
            s[2] += eval_vertex<false, true>(F0, 0, CH_A, SP_P, x, NA_);
            s[3] += eval_vertex<false, true>(F0, 0, CH_T, SP_P, x, NT_);
        }
        s[4] += eval_vertex<false, true>(F, 0, CH_P, SP_P, x, NP_);
        s[5] += eval_vertex<false, true>(F, 0, CH_A, SP_P, x, NA_);
        s[6] += eval_vertex<false, true>(F, 0, CH_T, SP_P, x, NT_);
        s[7] += eval_vertex<false, true>(F, 1, CH_P, SP_P, x, NP_);
        s[8] += eval_vertex<false, true>(F, 1, CH_A, SP_P, x, NA_);
        s[9] += eval_vertex<false, true>(F, 1, CH_T, SP_P, x, NT_);
    }
    const double inv = 1.0 / (double)npair;
#pragma unroll
    for (int j = 0; j < 10; ++j) s[j] = block_reduce(s[j]) * inv;
    if (threadIdx.x != 0) return;
    if (f0_flat) {
        s[1] = eval_vertex<false, true>(F0, 0, CH_P, SP_X, a, NP_);
        s[2] = eval_vertex<false, true>(F0, 0, CH_A, SP_P, a, NA_);
        s[3] = eval_vertex<false, true>(F0, 0, CH_T, SP_P, a, NT_);
    }
    // class K3 of a chain at (W, v, w, P): independent of the fermionic momenta
    auto k3 = [&](const DevChain& V, int l, int r, int v, int w) { Arg x = a; x.v = v; x.w = w; return mbe_classes(V, l, r, x).K3; };
    // (W, w, v', P) block: vertices multiplied by bubbles to the left
    out.c[0][i] = s[0];
    out.c[1][i] = s[1] + U - k3(F0, 0, CH_P, a1, W - b1 - 1);
    const C f0a = s[2] - U + k3(F0, 0, CH_A, a1, b1);
    const C f0t = s[3] - U + k3(F0, 0, CH_T, a1, b1);
    out.c[2][i] = f0a;
    out.c[3][i] = 2.0 * f0t - f0a;
    // (W, v, w, P) block: vertices multiplied by bubbles from the right
    const C gpp = s[4] - s[7], ga = s[5] - s[8], gt = s[6] - s[9];
    const C fp = s[4] - U + k3(F, 0, CH_P, a1, b1), fa = s[5] - U + k3(F, 0, CH_A, a1, b1), ft = s[6] - U + k3(F, 0, CH_T, a1, b1);
    out.c[4][i] = gpp;
    out.c[5][i] = ga;
    out.c[6][i] = gt * 2.0 - ga;
    out.c[7][i] = fp;
    out.c[8][i] = fa;
    out.c[9][i] = ft * 2.0 - fa;
}

// ---- the vertex as a callable: F(W, v, w, P, k, q, Ch, Sp; switches) of the chain from `level` at a list of points (what the
// reference's tests and conversions do on the host, e.g. test/test_boson_exchange_local.jl:76, src/boson_exchange.jl:672-735).
// Frequencies may be FDGA_INF; ip / ik / iq are linear momentum indices; sw != 0: k = q = kSW. ----
template <bool MBE>
__global__ void eval_points_kernel(const __grid_constant__ DevChain V, int level, int Ch, int Sp, unsigned flags, int sw, long long n,
                                   const int* __restrict__ W, const int* __restrict__ v, const int* __restrict__ w,
                                   const int* __restrict__ ip, const int* __restrict__ ik, const int* __restrict__ iq, C* __restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Arg a; a.W = W[i]; a.v = v[i]; a.w = w[i];
    a.Px = ip[i] % V.L; a.Py = ip[i] / V.L; a.kx = ik[i] % V.L; a.ky = ik[i] / V.L; a.qx = iq[i] % V.L; a.qy = iq[i] / V.L;
    out[i] = sw ? eval_vertex<true, MBE>(V, level, Ch, Sp, a, flags) : eval_vertex<false, MBE>(V, level, Ch, Sp, a, flags);
}

// ---- BSE_L_K2! of the local solver, generic form (one CTA per representative (W, v)): src/BSEa/BSEa_K2.jl:1-40, w over the bubble
// mesh, crossing on the right vertex (which is inside Rt = RK_LK2_LOC).  Used for MBE vertices, whose left factor does not split
// into the per-level pieces of the column kernels. ----
template <int CH>
__global__ void bse_lk2_loc_kernel(const __grid_constant__ DevChain F, const C* __restrict__ Rt, C* __restrict__ repvals,
                                   SymDev sg, long long c0, Grid g, double scale, const int* __restrict__ map) {
    constexpr int SP = (CH == CH_T) ? SP_D : SP_P;
    constexpr unsigned FLG = (CH == CH_P ? 0u : FL_GP) | (CH == CH_T ? 0u : FL_GT) | (CH == CH_A ? 0u : FL_GA);
    long long cls = c0 + blockIdx.x;
    long long idx = sg.index[sg.offsets[cls]];
    const int nB2 = 2 * g.nK2b - 1, nF2 = 2 * g.nK2f, nw = 2 * g.nPiF;
    long long t0 = idx;
    int iW = t0 % nB2; t0 /= nB2; int iv = t0 % nF2; t0 /= nF2; int iP = t0 % g.NP; int ik = t0 / g.NP;
    int W = iW - (g.nK2b - 1), v = iv - g.nK2f, Px = iP % g.L, Py = iP / g.L, kx = ik % g.L, ky = ik / g.L;
    const C* slab = Rt + (size_t)nw * g.NP * slab_index(map, iW, nB2, iP);
    C acc = zeroC();
    for (int t = threadIdx.x; t < nw * g.NP; t += blockDim.x) {
        int iq = t % g.NP, iw = t / g.NP;
        Arg a; a.W = W; a.v = v; a.w = iw - g.nPiF; a.Px = Px; a.Py = Py; a.kx = kx; a.ky = ky; a.qx = iq % g.L; a.qy = iq / g.L;
        acc += eval_vertex<false, true>(F, 0, CH, SP, a, FLG) * slab[t];
    }
    acc = block_reduce(acc);
    if (threadIdx.x == 0) repvals[cls] = acc * scale;
}

// ---- build_K3_cache_mfRG!: src/nonlocal_2/build_K3_cache.jl:108-161.  One thread per class rep --------
//  kind 0: Gpx (pCh,xSp)  1: Gpp (pCh,pSp)  2: Ga  3: Gt   = S.F(...; g_r=false) - S.F.F0(...; g_r=false)
//  kind 4: Fp  5: Fa  6: Ft                                = S.F0(W,v,w) - S.F0(W,inf,w)
template <bool MBE = false>
__global__ void cache_mfrg_kernel(const __grid_constant__ DevChain F0, const __grid_constant__ DevChain F, int kind,
                                  C* __restrict__ repvals, SymDev sg, long long c0, long long c1, Grid g) {
    long long cls = c0 + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (cls >= c1) return;
    long long idx = sg.index[sg.offsets[cls]];
    int W, a1, b1, iP; decode_k3(g, idx, W, a1, b1, iP);
    Arg a; a.W = W; a.v = a1; a.w = b1; a.Px = iP % g.L; a.Py = iP / g.L; a.kx = a.ky = a.qx = a.qy = 0;
    C r;
    if (kind < 4) {
        int Ch = (kind <= 1) ? CH_P : (kind == 2 ? CH_A : CH_T);
        int Sp = (kind == 0) ? SP_X : SP_P;
        unsigned f = FL_ALL & ~(2u << Ch);
        r = eval_vertex<true, MBE>(F, 0, Ch, Sp, a, f) - eval_vertex<true, MBE>(F, 1, Ch, Sp, a, f);
    } else {
        int Ch = (kind == 4) ? CH_P : (kind == 5 ? CH_A : CH_T);
        Arg av = a; av.v = FDGA_INF;
        r = eval_vertex<true, MBE>(F0, 0, Ch, SP_P, a, FL_ALL) - eval_vertex<true, MBE>(F0, 0, Ch, SP_P, av, FL_ALL);
    }
    repvals[cls] = r;
}

// ---- SDE_channel_L_pp!/ph!: src/nonlocal_2/SDE.jl:16-33, 54-73, 96-111, 130-145.  CTA per rep --------
template <bool PP, bool MBE = false>
__global__ void sde_L_kernel(const __grid_constant__ DevChain V, int level, const C* __restrict__ PiT,
                             C* __restrict__ repvals, SymDev sg, long long c0, Grid g, C U, double scale, int own_only, const int* __restrict__ map) {
    long long cls = c0 + blockIdx.x;
    long long idx = sg.index[sg.offsets[cls]];
    const int nB2 = 2 * g.nK2b - 1, nF2 = 2 * g.nK2f, nw = 2 * g.nPiF, nBP = 2 * g.nPiB - 1;
    long long t0 = idx;
    int iW = t0 % nB2; t0 /= nB2; int iv = t0 % nF2; t0 /= nF2; int iP = t0 % g.NP; int ik = t0 / g.NP;
    int W = iW - (g.nK2b - 1), v = iv - g.nK2f, Px = iP % g.L, Py = iP / g.L, kx = ik % g.L, ky = ik / g.L;
    const C* slab = PiT + (size_t)nw * g.NP * slab_index(map, posB(W, g.nPiB), nBP, iP);
    const bool is_core = V.lev[level].type == LV_CORE;
    C acc = zeroC();
    for (int t = threadIdx.x; t < nw * g.NP; t += blockDim.x) {
        int iq = t % g.NP, iw = t / g.NP;
        int w = iw - g.nPiF, qx = iq % g.L, qy = iq / g.L;
        C d;
        Arg a; a.W = W; a.Px = Px; a.Py = Py;
        if (PP) {
            a.v = W - w - 1; a.w = v; a.kx = Px - qx; a.ky = Py - qy; a.qx = kx; a.qy = ky;
            if (is_core) d = core_eval(V.lev[level], CH_P, SP_P, a.W, a.v, a.w) - U;
            else if (own_only) d = eval_vertex<false, MBE>(V, level, CH_P, SP_P, a, FL_GP);
            else d = eval_vertex<false, MBE>(V, level, CH_P, SP_P, a, FL_F0 | FL_GP) - eval_vertex<false, MBE>(V, level + 1, CH_P, SP_P, a, FL_F0 | FL_GP);
        } else {
            a.v = v; a.w = w; a.kx = kx; a.ky = ky; a.qx = qx; a.qy = qy;
            if (is_core) d = core_eval(V.lev[level], CH_A, SP_P, W, v, w) + core_eval(V.lev[level], CH_T, SP_P, W, v, w) - U - U;
            else if (own_only) d = eval_vertex<false, MBE>(V, level, CH_A, SP_P, a, FL_GA) + eval_vertex<false, MBE>(V, level, CH_T, SP_P, a, FL_GT);
            else d = eval_vertex<false, MBE>(V, level, CH_A, SP_P, a, FL_F0 | FL_GA) + eval_vertex<false, MBE>(V, level, CH_T, SP_P, a, FL_F0 | FL_GT)
                   - eval_vertex<false, MBE>(V, level + 1, CH_A, SP_P, a, FL_F0 | FL_GA) - eval_vertex<false, MBE>(V, level + 1, CH_T, SP_P, a, FL_F0 | FL_GT);
        }
        acc += U * slab[t] * d;
    }
    acc = block_reduce(acc);
    if (threadIdx.x == 0) repvals[cls] = acc * scale;
}

// ---- small DFT along one axis (FFTW conventions: sgn=-1 forward, +1 backward, unnormalised) -----------
// data viewed as [pre][n][post] column-major; out-of-place; result multiplied by `scale`.
__global__ void dft_axis_kernel(const C* __restrict__ in, C* __restrict__ out, long long pre, int n, long long post, int sgn, double scale,
                                const C* __restrict__ tw /* tw[j] = exp(+2 pi i j / n) */) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long tot = pre * n * post;
    if (i >= tot) return;
    long long a = i % pre; long long t = i / pre; int k = t % n; long long b = t / n;
    const C* p = in + a + pre * n * b;
    C s = zeroC();
    int jk = 0;
    for (int j = 0; j < n; ++j) {
        C w = tw[jk];
        if (sgn < 0) w.y = -w.y;
        s += p[pre * j] * w;
        jk += k; if (jk >= n) jk -= n;
    }
    out[i] = s * scale;
}
// 2-d DFT over two ADJACENT axes of size n in one launch: data viewed as [pre][n][n][post], one CTA per (a, b).  The n x n tile
// is gathered into shared memory, transformed along x, then along y, and written back; in-place use is safe (a CTA reads its whole
// tile before it writes, tiles are disjoint).  Replaces two dft_axis_kernel launches (and their global round trip) per 2-d transform.
__global__ void dft2_tile_kernel(const C* __restrict__ in, C* __restrict__ out, long long pre, int n, long long post, int sgn, double scale,
                                 const C* __restrict__ tw /* tw[j] = exp(+2 pi i j / n) */) {
    extern __shared__ __align__(16) double dft_sm[];
    C* A = reinterpret_cast<C*>(dft_sm); C* B = A + n * n; C* w = B + n * n;
    const long long a = blockIdx.x % pre, b = blockIdx.x / pre;
    (void)post;
    const long long base = a + pre * (long long)n * n * b;
    const int nn = n * n;
    for (int j = threadIdx.x; j < n; j += blockDim.x) { C t = tw[j]; if (sgn < 0) t.y = -t.y; w[j] = t; }
    for (int t = threadIdx.x; t < nn; t += blockDim.x) A[t] = in[base + pre * t];
    __syncthreads();
    for (int t = threadIdx.x; t < nn; t += blockDim.x) {
        const int kx = t % n, y = t / n;
        C s = zeroC(); int jk = 0;
        for (int x = 0; x < n; ++x) { s += A[x + n * y] * w[jk]; jk += kx; if (jk >= n) jk -= n; }
        B[t] = s;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nn; t += blockDim.x) {
        const int kx = t % n, ky = t / n;
        C s = zeroC(); int jk = 0;
        for (int y = 0; y < n; ++y) { s += B[kx + n * y] * w[jk]; jk += ky; if (jk >= n) jk -= n; }
        out[base + pre * t] = s * scale;
    }
}
__global__ void twiddle_kernel(C* tw, int n) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double sn, cs; sincospi(2.0 * (double)j / (double)n, &sn, &cs);
    tw[j] = mkC(cs, sn);
}

// G_R(n) call semantics: 0 outside the fermionic mesh
__device__ __forceinline__ C gr_call(const C* GR, int nG, int LG, int n, int ix, int iy) {
    if (!inF(n, nG)) return zeroC();
    return GR[posF(n, nG) + (size_t)(2 * nG) * (ix + (size_t)LG * iy)];
}

// ---- bubbles_real_space!: src/nonlocal_2/bubble.jl:68-116.  One thread per Pi_R element ---------------
__global__ void bubbles_rs_kernel(const C* __restrict__ GR, C* __restrict__ PippR, C* __restrict__ PiphR, Grid g) {
    const int nBP = 2 * g.nPiB - 1, nFP = 2 * g.nPiF, L = g.L, LG = g.LG, h = L / 2;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long n = (long long)nBP * nFP * g.NP * g.NP;
    if (i >= n) return;
    long long t = i;
    int iW = t % nBP; t /= nBP; int iv = t % nFP; t /= nFP; int ia = t % g.NP; int ib = t / g.NP;
    int W = iW - (g.nPiB - 1), v = iv - g.nPiF;
    int ax = ia % L, ay = ia / L, bx = ib % L, by = ib / L;
    C spp = zeroC(), sph = zeroC();
    const bool even = (LG % 2 == 0);
    for (int R2 = -h; R2 <= h; ++R2) { if (modL(R2, L) != ay) continue;
    for (int R1 = -h; R1 <= h; ++R1) { if (modL(R1, L) != ax) continue;
        const C g1pp = gr_call(GR, g.nG, LG, W - v - 1, modL(R1, LG), modL(R2, LG));
        const C g1ph = gr_call(GR, g.nG, LG, W + v, modL(R1, LG), modL(R2, LG));
        double wR = 1.0;
        if (even) { if (abs(R1) == LG / 2) wR *= 0.5; if (abs(R2) == LG / 2) wR *= 0.5; }
#pragma unroll
        for (int ph = 0; ph < 2; ++ph) {
            // pp: R' - R == b (mod L)  ->  R' == b + R ; ph: R' + R == b  ->  R' == b - R ; R' restricted to [-h, h]
            int c1 = modL(ph ? bx - R1 : bx + R1, L), c2 = modL(ph ? by - R2 : by + R2, L);
            while (c1 > h) c1 -= L;
            while (c2 > h) c2 -= L;
            for (int Rp2 = c2; Rp2 >= -h; Rp2 -= L) for (int Rp1 = c1; Rp1 >= -h; Rp1 -= L) {
                double wgt = wR;
                if (even) { if (abs(Rp1) == LG / 2) wgt *= 0.5; if (abs(Rp2) == LG / 2) wgt *= 0.5; }
                const C g2 = gr_call(GR, g.nG, LG, v, modL(Rp1, LG), modL(Rp2, LG));
                if (ph) sph += g1ph * g2 * wgt; else spp += g1pp * g2 * wgt;
            }
        }
    }}
    PippR[i] = spp; PiphR[i] = sph;
}

// ---- bubbles_real_space! in product form --------------------------------------------------------------------------------
// The reference fills Pi_R[W, v, R, R' -+ R] += G_R(W -+ v) G_R'(v) w(R) w(R') for R, R' in [-h, h]^2 (h = L / 2) and transforms
// the four momentum axes back (src/nonlocal_2/bubble.jl:68-119).  Both the fill and the weights factorise, so
//     Pipp[W, v, P, k] = Ghat(W - v - 1, P - k) * Ghat(v, k),      Piph[W, v, P, k] = Ghat(W + v, P + k) * Ghat(v, k),
//     Ghat(n, p) = sum_{R in [-h, h]^2} w(R) G_R(n; R mod LG) exp(+2 pi i p.R / L)
// with the COARSE-GRAINED Green function Ghat on the vertex mesh (0 outside the fermionic mesh, use_G_tail = false).  One
// pointwise product per element, for any subset of the bubble: no 4-d transform, no bubble-sized intermediate.
// w(R) = 1/2 per component with |R_c| = LG / 2 (LG even): as coded, the test is against the G mesh.
__global__ void coarse_green_kernel(const C* __restrict__ GR, C* __restrict__ Ghat, int nG, int LG, int L, const C* __restrict__ twL) {
    const int nGf = 2 * nG, NP = L * L, h = L / 2;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)nGf * NP) return;
    const int n = (int)(i % nGf), p = (int)(i / nGf), px = p % L, py = p / L;
    const bool even = (LG % 2 == 0);
    C s = zeroC();
    for (int R2 = -h; R2 <= h; ++R2) for (int R1 = -h; R1 <= h; ++R1) {
        double w = 1.0;
        if (even) { if (abs(R1) == LG / 2) w *= 0.5; if (abs(R2) == LG / 2) w *= 0.5; }
        const C gr = GR[n + (size_t)nGf * (modL(R1, LG) + (size_t)LG * modL(R2, LG))];
        s += gr * twL[modL(px * R1 + py * R2, L)] * w;
    }
    Ghat[i] = s;
}
FDGA_HD C ghat_call(const C* __restrict__ Ghat, int nG, int n, int ip) {
    return inF(n, nG) ? Ghat[posF(n, nG) + (size_t)(2 * nG) * ip] : zeroC();
}
// one bubble on the listed slabs, slab layout [q, w | slab]
__global__ void bubble_slabs_kernel(const C* __restrict__ Ghat, C* __restrict__ PiT, Grid g, int pp, const int4* __restrict__ slabs, int nslabs) {
    const int nFP = 2 * g.nPiF, L = g.L, NP = g.NP;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)nFP * NP * nslabs) return;
    long long t = i;
    const int iq = t % NP; t /= NP; const int iw = t % nFP; const int sl = (int)(t / nFP);
    const int W = slabs[sl].x - (g.nPiB - 1), iP = slabs[sl].y, v = iw - g.nPiF;
    const int Px = iP % L, Py = iP / L, kx = iq % L, ky = iq / L;
    const C gk = ghat_call(Ghat, g.nG, v, iq);
    PiT[i] = (pp ? ghat_call(Ghat, g.nG, W - v - 1, kidx(Px - kx, Py - ky, L)) : ghat_call(Ghat, g.nG, W + v, kidx(Px + kx, Py + ky, L))) * gk;
}
// Pisw[W, v, P] = mean_k Pi[W, v, P, k] of one bubble
__global__ void bubble_swave_kernel(const C* __restrict__ Ghat, C* __restrict__ Pisw, Grid g, int pp) {
    const int nBP = 2 * g.nPiB - 1, nFP = 2 * g.nPiF, L = g.L, NP = g.NP;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)nBP * nFP * NP) return;
    long long t = i;
    const int iW = t % nBP; t /= nBP; const int iv = t % nFP; const int iP = (int)(t / nFP);
    const int W = iW - (g.nPiB - 1), v = iv - g.nPiF, Px = iP % L, Py = iP / L;
    C s = zeroC();
    for (int ik = 0; ik < NP; ++ik) {
        const int kx = ik % L, ky = ik / L;
        s += (pp ? ghat_call(Ghat, g.nG, W - v - 1, kidx(Px - kx, Py - ky, L)) : ghat_call(Ghat, g.nG, W + v, kidx(Px + kx, Py + ky, L))) * ghat_call(Ghat, g.nG, v, ik);
    }
    Pisw[i] = s / (double)NP;
}
// the whole bubbles in the reference's layout [W, v, P, k]
__global__ void bubbles_product_kernel(const C* __restrict__ Ghat, C* __restrict__ Pipp, C* __restrict__ Piph, Grid g) {
    const int nBP = 2 * g.nPiB - 1, nFP = 2 * g.nPiF, L = g.L;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)nBP * nFP * g.NP * g.NP) return;
    long long t = i;
    const int iW = t % nBP; t /= nBP; const int iv = t % nFP; t /= nFP; const int iP = t % g.NP; const int ik = (int)(t / g.NP);
    const int W = iW - (g.nPiB - 1), v = iv - g.nPiF;
    const int Px = iP % L, Py = iP / L, kx = ik % L, ky = ik / L;
    const C gk = ghat_call(Ghat, g.nG, v, ik);
    Pipp[i] = ghat_call(Ghat, g.nG, W - v - 1, kidx(Px - kx, Py - ky, L)) * gk;
    Piph[i] = ghat_call(Ghat, g.nG, W + v, kidx(Px + kx, Py + ky, L)) * gk;
}

// ---- bubbles_momentum_space!: src/nonlocal_2/bubble.jl:1-37 -----------------------------------------
__global__ void bubbles_ms_kernel(const C* __restrict__ G, C* __restrict__ Pipp, C* __restrict__ Piph, Grid g) {
    const int nBP = 2 * g.nPiB - 1, nFP = 2 * g.nPiF, L = g.L, LG = g.LG, ratio = LG / L, nGf = 2 * g.nG;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long n = (long long)nBP * nFP * g.NP * g.NP;
    if (i >= n) return;
    long long t = i;
    int iW = t % nBP; t /= nBP; int iv = t % nFP; t /= nFP; int iP = t % g.NP; int ik = t / g.NP;
    int W = iW - (g.nPiB - 1), v = iv - g.nPiF;
    int Px = (iP % L) * ratio, Py = (iP / L) * ratio, kx = (ik % L) * ratio, ky = (ik / L) * ratio;
    C pp = zeroC(), ph = zeroC();
    if (inF(v, g.nG)) {
        C gk = G[posF(v, g.nG) + (size_t)nGf * kidx(kx, ky, LG)];
        int a = W - v - 1, b = W + v;
        if (inF(a, g.nG)) pp = gk * G[posF(a, g.nG) + (size_t)nGf * kidx(Px - kx, Py - ky, LG)];
        if (inF(b, g.nG)) ph = gk * G[posF(b, g.nG) + (size_t)nGf * kidx(Px + kx, Py + ky, LG)];
    }
    Pipp[i] = pp; Piph[i] = ph;
}

// ---- bubbles! of the local solver: src/bubble.jl:9-36 (use_G_tail = true: 1/nu outside the G mesh) -------------
__global__ void bubbles_local_kernel(const C* __restrict__ G, C* __restrict__ Pipp, C* __restrict__ Piph, Grid g) {
    const int nBP = 2 * g.nPiB - 1, nFP = 2 * g.nPiF;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)nBP * nFP) return;
    int W = (int)(i % nBP) - (g.nPiB - 1), v = (int)(i / nBP) - g.nPiF;
    auto Gt = [&](int n) -> C { return inF(n, g.nG) ? G[posF(n, g.nG)] : mkC(1.0 / ((2 * n + 1) * 3.141592653589793 * g.T), 0.0); };
    C Gv = Gt(v);
    Pipp[i] = Gv * Gt(W - v - 1);
    Piph[i] = Gv * Gt(W + v);
}

// ---- Dyson!: src/dyson.jl:20-31 --------------------------------------------------------------------
__global__ void dyson_kernel(C* __restrict__ G, const C* __restrict__ Sigma, const C* __restrict__ Gbare, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    C gb = Gbare[i];
    double d = gb.x * gb.x + gb.y * gb.y;
    C inv = mkC(gb.x / d, -gb.y / d) + Sigma[i];
    double e = inv.x * inv.x + inv.y * inv.y;
    G[i] = mkC(inv.x / e, -inv.y / e);
}

// ---- compute_occupation: src/dyson.jl:39-41 (single CTA, deterministic) ---------------------------------
__global__ void occupation_kernel(const C* __restrict__ G, long long n, double T, double Nk, double* occ) {
    C acc = zeroC();
    for (long long i = threadIdx.x; i < n; i += blockDim.x) acc += G[i];
    acc = block_reduce(acc);
    if (threadIdx.x == 0) *occ = 0.5 + acc.y * T / Nk;
}
// ---- FP64 FMA micro-benchmark (roofline denominator of the contraction kernels) ---------------------------------
// 8 independent DFMA chains per thread; the result is written so that the loop cannot be removed.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// ---- Fourier interpolation between momentum meshes: src/interpolate.jl:1-165 -------------------------------------------
// The reference transforms to real space (fft / Li^d), copies the coefficients R in [-Li/2, Li/2]^d to the output mesh (half
// weights at |R_c| = Li/2 for even Li) and transforms back (bfft).  All three steps are linear and factorise over the momentum
// components, so per component it is ONE small matrix  M[xo, xi] = 1/Li sum_R w(R) exp(2 pi i R (xo/Lo - xi/Li))  applied along
// that axis (built on the host in double precision, fdga_lib.cu: interp_matrix).
// Step 1: re-box the frequency axes (values outside the input meshes: 0, or clamped to the edge for the self-energy,
// src/interpolate.jl:176-186).  Frequency index maps are shifts: i_in = i_out + (N_in - N_out) for bosonic and fermionic meshes.
struct InterpBox { int nd; int no[3]; int ni[3]; int shift[3]; };
__global__ void interp_rebox_kernel(const C* __restrict__ in, C* __restrict__ out, InterpBox b, long long M, int clamp) {
    long long Fo = (long long)b.no[0] * b.no[1] * b.no[2], Fi = (long long)b.ni[0] * b.ni[1] * b.ni[2];
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= Fo * M) return;
    long long fo = i % Fo, m = i / Fo;
    int io[3]; io[0] = (int)(fo % b.no[0]); io[1] = (int)((fo / b.no[0]) % b.no[1]); io[2] = (int)(fo / ((long long)b.no[0] * b.no[1]));
    long long fi = 0, stride = 1; bool ok = true;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        int j = io[d] + b.shift[d];
        if (j < 0 || j >= b.ni[d]) { if (clamp) j = j < 0 ? 0 : b.ni[d] - 1; else ok = false; }
        fi += stride * j; stride *= b.ni[d];
    }
    out[i] = ok ? in[fi + Fi * m] : zeroC();
}
// Step 2 (once per momentum component): out[pre, xo, post] = sum_xi M[xo + Lo * xi] in[pre, xi, post]
__global__ void interp_axis_kernel(const C* __restrict__ in, C* __restrict__ out, long long pre, int Li, int Lo, long long post,
                                   const C* __restrict__ M) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= pre * Lo * post) return;
    const long long ip = i % pre, r = i / pre; const int xo = (int)(r % Lo); const long long q = r / Lo;
    C s0 = zeroC(), s1 = zeroC();
    const C* src = in + ip + pre * (size_t)Li * q;
    for (int xi = 0; xi < Li; xi += 2) {
        s0 += M[xo + Lo * xi] * src[pre * xi];
        if (xi + 1 < Li) s1 += M[xo + Lo * (xi + 1)] * src[pre * (xi + 1)];
    }
    out[i] = s0 + s1;
}

// ---- hubbard_bare_Green: src/models/hubbard.jl:8-44.  Stored quantity is i*G0 = i / (i nu + mu - eps_k) ----------------
__device__ __forceinline__ C hubbard_bare_entry(int n, int ik, int LG, double T, double mu, double t1, double t2, double t3) {
    const int ix = ik % LG, iy = ik / LG;
    const double k1 = 2.0 * M_PI * ix / LG, k2 = 2.0 * M_PI * iy / LG;
    double ek = -2.0 * t1 * (cos(k1) + cos(k2)); ek += -4.0 * t2 * cos(k1) * cos(k2); ek += -2.0 * t3 * (cos(2.0 * k1) + cos(2.0 * k2));
    const double nu = (2 * n + 1) * M_PI * T, re = mu - ek;        // 1 / (re + i nu) * i = (nu + i re) / (re^2 + nu^2)
    const double d = re * re + nu * nu;
    return mkC(nu / d, re / d);
}
__global__ void hubbard_bare_green_kernel(C* __restrict__ G, int nG, int LG, double T, double mu, double t1, double t2, double t3) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long n = (long long)2 * nG * LG * LG;
    if (i >= n) return;
    G[i] = hubbard_bare_entry((int)(i % (2 * nG)) - nG, (int)(i / (2 * nG)), LG, T, mu, t1, t2, t3);
}
// occupation(mu) of compute_hubbard_chemical_potential (src/dyson.jl:45-57): Gbare(mu) -> Dyson -> compute_occupation, fused;
// single CTA, deterministic
__global__ void occupation_mu_kernel(const C* __restrict__ Sigma, int nG, int LG, double T, double mu, double t1, double t2, double t3,
                                     double* occ) {
    const long long n = (long long)2 * nG * LG * LG;
    C acc = zeroC();
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const C gb = hubbard_bare_entry((int)(i % (2 * nG)) - nG, (int)(i / (2 * nG)), LG, T, mu, t1, t2, t3);
        const double d = gb.x * gb.x + gb.y * gb.y;
        const C inv = mkC(gb.x / d, -gb.y / d) + Sigma[i];
        const double e = inv.x * inv.x + inv.y * inv.y;
        acc += mkC(inv.x / e, -inv.y / e);
    }
    acc = block_reduce(acc);
    if (threadIdx.x == 0) *occ = 0.5 + acc.y * T / ((double)LG * LG);
}
// Sigma += sgn * i (n - 1/2) U    (Hartree, src/nonlocal_2/SDE.jl:317-321, src/SDE.jl:19-23)
__global__ void hartree_kernel(C* __restrict__ Sigma, const double* occ, C U, double sgn, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    C h = (U * (*occ - 0.5)) * mkC(0.0, 1.0);
    Sigma[i] += h * sgn;
}

// ---- real-space contraction of SDE_compute!: src/nonlocal_2/SDE.jl:200-250 -----------------------------
// SigR[v, tx, ty] (nSf x LS x LS, pre-zeroed) = T * sum_{R,Rp in [-h,h]^2} w(R,Rp) *
//     ( [R+Rp == t mod LS] G_R(W-v; -R) Lpp_R[W,v,R,Rp] + [-R+Rp == t mod LS] G_R(W+v; -R) Lph_R[W,v,R,Rp] )
// One WARP per (v on the K2 mesh, t): lanes split the R sum, Rp is enumerated by congruence instead of scanning all
// (R, Rp) pairs.  Only t inside the window [-2h, 2h]^2 (mod LS) can be reached by R + Rp / -R + Rp, so only those
// warps are launched (SigR is pre-zeroed): grid = nF * tw * tw warps with tw = min(LS, 4h + 1).
__global__ void sde_rs_kernel(const C* __restrict__ GR, const C* __restrict__ LppR, const C* __restrict__ LphR,
                              C* __restrict__ SigR, Grid g, int nSig, int LS, int tw) {
    const int L = g.L, h = L / 2, LG = g.LG, nB = 2 * g.nK2b - 1, nF = 2 * g.nK2f, nSf = 2 * nSig;
    const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= (long long)nF * tw * tw) return;
    const int iv = wid % nF, wx = (wid / nF) % tw, wy = (wid / nF) / tw;
    const int tx = (tw == LS) ? wx : modL(wx - 2 * h, LS), ty = (tw == LS) ? wy : modL(wy - 2 * h, LS);
    const int v = iv - g.nK2f;
    if (!inF(v, nSig)) return;
    C acc = zeroC();
    const size_t pre = (size_t)nB * nF;
    const bool even = (L % 2 == 0);
    const int nR = 2 * h + 1;
    for (int r = lane; r < nR * nR; r += 32) {
        const int R1 = r % nR - h, R2 = r / nR - h;
        const int gx = modL(-R1, LG), gy = modL(-R2, LG);
        const int iRL = modL(R1, L) + L * modL(R2, L);
        double wR = 1.0;
        if (even) { if (abs(R1) == h) wR *= 0.5; if (abs(R2) == h) wR *= 0.5; }
#pragma unroll
        for (int ph = 0; ph < 2; ++ph) {
            // pp: Rp = t - R ; ph: Rp = t + R  (mod LS), restricted to [-h, h]
            int c1 = modL(ph ? tx + R1 : tx - R1, LS), c2 = modL(ph ? ty + R2 : ty - R2, LS);
            while (c1 > h) c1 -= LS;
            while (c2 > h) c2 -= LS;
            while (c1 + LS <= h) c1 += LS;      // start from the largest representative <= h (only matters for LS <= h)
            while (c2 + LS <= h) c2 += LS;
            for (int Rp2 = c2; Rp2 >= -h; Rp2 -= LS) for (int Rp1 = c1; Rp1 >= -h; Rp1 -= LS) {
                double wgt = wR;
                if (even) { if (abs(Rp1) == h) wgt *= 0.5; if (abs(Rp2) == h) wgt *= 0.5; }
                const int iRpL = modL(Rp1, L) + L * modL(Rp2, L);
                const size_t lbase = pre * (iRL + (size_t)g.NP * iRpL) + (size_t)nB * iv;
                const C* Lr = ph ? LphR : LppR;
                C part = zeroC();
                for (int iW = 0; iW < nB; ++iW) {
                    const int W = iW - (g.nK2b - 1);
                    part += gr_call(GR, g.nG, LG, ph ? W + v : W - v - 1, gx, gy) * Lr[lbase + iW];
                }
                acc += part * wgt;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o); }
    if (lane == 0) SigR[posF(v, nSig) + (size_t)nSf * (tx + (size_t)LS * ty)] = acc * g.T;
}

// ---- SDE_U2_using_G: src/nonlocal/SDE.jl:421-438 ---------------------------------------------------------
// SR[v, R] = fac * sum_{n1, n2} Gm[n1, R] Gp[n2, R] Gp[n1 - n2 + v, R].  One CTA per R: the correlation
// c[d] = sum_{n1} Gm[n1] Gp[n1 + d] is formed once in shared memory, then SR[v] = fac * sum_{n2} Gp[n2] c[v - n2].
__global__ void sde_u2_kernel(const C* __restrict__ Gp, const C* __restrict__ Gm, C* __restrict__ SR, int nG, int LG, C fac) {
    extern __shared__ double sm_raw[];
    const int nGf = 2 * nG, nd = 2 * nGf - 1;
    C* gp = reinterpret_cast<C*>(sm_raw);     // [nGf]
    C* gm = gp + nGf;                         // [nGf]
    C* c = gm + nGf;                          // [nd]   d = -(nGf-1) .. nGf-1
    const size_t iR = blockIdx.x;
    for (int i = threadIdx.x; i < nGf; i += blockDim.x) { gp[i] = Gp[(size_t)nGf * iR + i]; gm[i] = Gm[(size_t)nGf * iR + i]; }
    __syncthreads();
    for (int j = threadIdx.x; j < nd; j += blockDim.x) {
        const int d = j - (nGf - 1);
        const int lo = max(0, -d), hi = min(nGf - 1, nGf - 1 - d);
        C s = zeroC();
        for (int i1 = lo; i1 <= hi; ++i1) s += gm[i1] * gp[i1 + d];
        c[j] = s;
    }
    __syncthreads();
    for (int iv = threadIdx.x; iv < nGf; iv += blockDim.x) {
        C acc = zeroC();
        for (int i2 = 0; i2 < nGf; ++i2) acc += gp[i2] * c[iv - i2 + (nGf - 1)];
        SR[(size_t)nGf * iR + iv] = acc * fac;
    }
}

}  // namespace fdga
