// =====================================================================================
// fdga_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
// A plain C++ restatement of the hot path of jaemolihm/fdDGAsolver.jl (Bethe-Salpeter step
// of the fd-parquet / mfRG iteration, K3-cache build, bubbles, SDE) in the REFERENCE'S OWN
// LOOP STRUCTURE: no hoisting of the right factor, no pre-tabulated s-wave means, the
// nested F0 chain evaluated recursively exactly as the Julia callable structs do.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library.  The product (libfdga.so) never links or calls it.
//
// Parity pinning: the reference itself (Julia + MatsubaraFunctions.jl@mesh_generalization_v3,
// neither available here) cannot be executed in this container; the oracle is pinned against
// the reference's own tests restated in tests/test_oracle_*.py (golden occupations
// test/test_hubbard.jl:84-88, evaluator identities test/test_nonlocal_2_vertex.jl:27-36,91-97,
// s-wave == explicit BZ average :114-222, bubble identities test/test_hubbard.jl:73-78,
// fdPA == scPA at zero reference test/test_nonlocal_2_fdPA.jl:39-40, converged fdPA vs scPA
// :59-81).  SymmetryGroup class order / BZ linear order / mfRG branches are "parity unpinned"
// (SURVEY.md section 8c).
//
// Conventions: all arrays column-major, first index fastest,
// complex double.  Matsubara frequencies are carried as integer indices:
//   fermion n <-> (2n+1) pi T on mesh n = -N..N-1;  boson m <-> 2 m pi T on mesh m = -(N-1)..N-1.
// nu = infinity (reference src/types.jl:134-145) is the sentinel INF.
// Channels: 0 = pCh, 1 = tCh, 2 = aCh (flatten order, src/vertex.jl:153-167).
// Spins:    0 = pSp, 1 = xSp, 2 = dSp.
// =====================================================================================
#include <complex>
#include <vector>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <climits>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef std::complex<double> cplx;
static const int INF = INT_MAX / 4;

enum { pCh = 0, tCh = 1, aCh = 2 };
enum { pSp = 0, xSp = 1, dSp = 2 };
enum { LV_NL2 = 0, LV_LOCAL = 1, LV_CORE = 2, LV_NL = 3 /* NL_Vertex: bosonic momentum only, src/nonlocal/vertex.jl */,
       LV_NL2_MBE = 4, LV_LOCAL_MBE = 5 /* NL2_MBEVertex / MBEVertex: same arrays, multi-boson-exchange evaluation, src/boson_exchange.jl */ };
enum { CL_K1 = 0, CL_K2 = 1, CL_K2P = 2, CL_K3 = 3, CL_LAMBDA = 4 };      // ClassTag, src/boson_exchange.jl:1-48
static inline bool is_mbe(int t) { return t == LV_NL2_MBE || t == LV_LOCAL_MBE; }
static inline bool is_local_lv(int t) { return t == LV_LOCAL || t == LV_LOCAL_MBE; }

// ---------------------------------------------------------------------------------
// Matsubara index arithmetic (derived from value(nu) = (2n+1) pi T etc.)
static inline bool isinf_(int a) { return a >= INF / 2; }
static inline int B_minus_F(int m, int n) { return (isinf_(m) || isinf_(n)) ? INF : m - n - 1; }  // -> F
static inline int B_plus_F (int m, int n) { return (isinf_(m) || isinf_(n)) ? INF : m + n; }      // -> F
static inline int F_minus_F(int a, int b) { return (isinf_(a) || isinf_(b)) ? INF : a - b; }      // -> B
static inline int F_plus_F (int a, int b) { return (isinf_(a) || isinf_(b)) ? INF : a + b + 1; }  // -> B
static inline bool inB(int m, int N) { return !isinf_(m) && m >= -(N - 1) && m <= N - 1; }
static inline bool inF(int n, int N) { return !isinf_(n) && n >= -N && n <= N - 1; }
static inline int posB(int m, int N) { return m + N - 1; }
static inline int posF(int n, int N) { return n + N; }

// Brillouin point (unfolded integer pair) or the s-wave point kSW (src/nonlocal/swave.jl:17-28)
struct Mom { int x, y; bool sw; };
static inline Mom mk(int x, int y) { Mom m; m.x = x; m.y = y; m.sw = false; return m; }
static inline Mom SW() { Mom m; m.x = 0; m.y = 0; m.sw = true; return m; }
static inline Mom operator+(Mom a, Mom b) { if (a.sw || b.sw) return SW(); return mk(a.x + b.x, a.y + b.y); }
static inline Mom operator-(Mom a, Mom b) { if (a.sw || b.sw) return SW(); return mk(a.x - b.x, a.y - b.y); }
static inline int mod_(int a, int L) { int r = a % L; return r < 0 ? r + L : r; }
// mesh_index_bc: fold and linearise, x fastest
static inline int kidx(Mom k, int L) { return mod_(k.x, L) + L * mod_(k.y, L); }

// ---------------------------------------------------------------------------------
// C-layout descriptors filled by the Python adapter (oracle/oracle.py)
extern "C" {
typedef struct {
    int type;                       // LV_NL2 / LV_LOCAL / LV_CORE / LV_NL
    int nK1, nK2b, nK2f, nK3b, nK3f; // mesh N's (CORE: nK3b, nK3f = box of the 4 core arrays)
    double U_re, U_im;              // CORE only
    const cplx* K1[3];              // [p, t, a]
    const cplx* K2[3];
    const cplx* K3[3];
    const cplx* core[4];            // Fp_p, Fp_x, Ft_p, Ft_x
} orc_level;

typedef struct {
    int nlev;                       // chain F -> F.F0 -> ... ; last level is CORE
    orc_level lev[8];
} orc_vertex;

typedef struct {
    double T;
    int L;                          // linear size of the vertex / bubble momentum mesh
    int nPiB, nPiF;                 // N of the bubble's bosonic / fermionic Matsubara meshes
    int swave;                      // 1: s-wave NL_ParquetSolver (bubbles Pi[W,w,P], K2[W,v,P]); 0: NL2 / local
} orc_grid;

typedef struct {                    // symmetry classes, CSR; representative = first member
    int64_t nclasses;
    const int64_t* offsets;         // nclasses + 1
    const int64_t* index;           // 0-based linear index into the target array
    const uint8_t* op;              // bit0 = sgn, bit1 = con
} orc_sg;
}

static inline cplx apply_op(uint8_t op, cplx v) {
    if (op & 2) v = std::conj(v);
    if (op & 1) v = -v;
    return v;
}

// ---------------------------------------------------------------------------------
// flags for the vertex evaluators' keyword switches
struct Flags { bool F0, gp, gt, ga; };
static inline Flags ALLF() { Flags f = {true, true, true, true}; return f; }

// ---- NL2_Channel evaluator, src/nonlocal_2/channel.jl:58-194 ------------------------
struct NL2Chan {
    int nK1, nK2b, nK2f, nK3b, nK3f, L, NP;
    const cplx *K1, *K2, *K3;

    // getindex with possible kSW arguments, src/nonlocal/swave.jl:32-136
    cplx k1(int W, Mom P) const {
        int iW = posB(W, nK1), nB = 2 * nK1 - 1;
        if (!P.sw) return K1[iW + (size_t)nB * kidx(P, L)];
        cplx s = 0; for (int p = 0; p < NP; p++) s += K1[iW + (size_t)nB * p];
        return s / (double)NP;
    }
    cplx k2(int W, int v, Mom P, Mom k) const {
        int nB = 2 * nK2b - 1, nF = 2 * nK2f;
        size_t base = posB(W, nK2b) + (size_t)nB * posF(v, nK2f);
        size_t sP = (size_t)nB * nF, sk = sP * NP;
        if (!P.sw && !k.sw) return K2[base + sP * kidx(P, L) + sk * kidx(k, L)];
        if (!P.sw && k.sw) {
            cplx s = 0; size_t b = base + sP * kidx(P, L);
            for (int i = 0; i < NP; i++) s += K2[b + sk * i];
            return s / (double)NP;
        }
        if (P.sw && !k.sw) {
            cplx s = 0; size_t b = base + sk * kidx(k, L);
            for (int i = 0; i < NP; i++) s += K2[b + sP * i];
            return s / (double)NP;
        }
        cplx s = 0;   // sum(view(f, i1, i2, :, :)) / N3 / N4 (column-major traversal)
        for (int j = 0; j < NP; j++) for (int i = 0; i < NP; i++) s += K2[base + sP * i + sk * j];
        return s / (double)NP / (double)NP;
    }
    cplx k3(int W, int v, int w, Mom P) const {
        int nB = 2 * nK3b - 1, nF = 2 * nK3f;
        size_t base = posB(W, nK3b) + (size_t)nB * (posF(v, nK3f) + (size_t)nF * posF(w, nK3f));
        size_t sP = (size_t)nB * nF * nF;
        if (!P.sw) return K3[base + sP * kidx(P, L)];
        cplx s = 0; for (int p = 0; p < NP; p++) s += K3[base + sP * p];
        return s / (double)NP;
    }

    cplx eval(int W, int v, int w, Mom P, Mom k, Mom q, bool fK1, bool fK2, bool fK3) const {
        cplx val = 0;
        bool vi = isinf_(v), wi = isinf_(w);
        if (!vi && !wi) {                                   // :58-102
            if (inB(W, nK1)) {
                if (fK1) val += k1(W, P);
                if (inB(W, nK2b)) {
                    bool a = inF(v, nK2f), b = inF(w, nK2f);
                    if (a && b) {
                        if (fK2) val += k2(W, v, P, k) + k2(W, w, P, q);
                        if (fK3 && inB(W, nK3b) && inF(v, nK3f) && inF(w, nK3f)) val += k3(W, v, w, P);
                    } else if (a) {
                        if (fK2) val += k2(W, v, P, k);
                    } else if (b) {
                        if (fK2) val += k2(W, w, P, q);
                    }
                }
            }
        } else if (vi && !wi) {                             // :105-137
            if (fK1) {
                if (inB(W, nK1)) {
                    val += k1(W, P);
                    if (fK2 && inB(W, nK2b) && inF(w, nK2f)) val += k2(W, w, P, q);
                }
            } else {
                if (fK2 && inB(W, nK2b) && inF(w, nK2f)) val += k2(W, w, P, q);
            }
        } else if (!vi && wi) {                             // :139-171
            if (fK1) {
                if (inB(W, nK1)) {
                    val += k1(W, P);
                    if (fK2 && inB(W, nK2b) && inF(v, nK2f)) val += k2(W, v, P, k);
                }
            } else {
                if (fK2 && inB(W, nK2b) && inF(v, nK2f)) val += k2(W, v, P, k);
            }
        } else {                                            // :173-194
            if (fK1 && inB(W, nK1)) val += k1(W, P);
        }
        return val;
    }
};

// ---- NL_Channel evaluator (K1[W,P], K2[W,v,P], K3[W,v,w,P]), src/nonlocal/channel.jl:68-226; P may be the s-wave point:
// getindex then averages over the momentum axis (src/nonlocal/swave.jl:32-74)
struct NLChan {
    int nK1, nK2b, nK2f, nK3b, nK3f, L, NP;
    const cplx *K1, *K2, *K3;
    cplx k1(int W, Mom P) const {
        int iW = posB(W, nK1), nB = 2 * nK1 - 1;
        if (!P.sw) return K1[iW + (size_t)nB * kidx(P, L)];
        cplx s = 0; for (int p = 0; p < NP; p++) s += K1[iW + (size_t)nB * p];
        return s / (double)NP;
    }
    cplx k2(int W, int v, Mom P) const {
        int nB = 2 * nK2b - 1, nF = 2 * nK2f;
        size_t base = posB(W, nK2b) + (size_t)nB * posF(v, nK2f), sP = (size_t)nB * nF;
        if (!P.sw) return K2[base + sP * kidx(P, L)];
        cplx s = 0; for (int p = 0; p < NP; p++) s += K2[base + sP * p];
        return s / (double)NP;
    }
    cplx k3(int W, int v, int w, Mom P) const {
        int nB = 2 * nK3b - 1, nF = 2 * nK3f;
        size_t base = posB(W, nK3b) + (size_t)nB * (posF(v, nK3f) + (size_t)nF * posF(w, nK3f)), sP = (size_t)nB * nF * nF;
        if (!P.sw) return K3[base + sP * kidx(P, L)];
        cplx s = 0; for (int p = 0; p < NP; p++) s += K3[base + sP * p];
        return s / (double)NP;
    }
    cplx eval(int W, int v, int w, Mom P, bool fK1 = true, bool fK2 = true, bool fK3 = true) const {
        cplx val = 0;
        bool vi = isinf_(v), wi = isinf_(w);
        if (!vi && !wi) {                                   // :90-130
            if (inB(W, nK1)) {
                if (fK1) val += k1(W, P);
                if (inB(W, nK2b)) {
                    bool a = inF(v, nK2f), b = inF(w, nK2f);
                    if (a && b) {
                        if (fK2) val += k2(W, v, P) + k2(W, w, P);
                        if (fK3 && inB(W, nK3b) && inF(v, nK3f) && inF(w, nK3f)) val += k3(W, v, w, P);
                    } else if (a) { if (fK2) val += k2(W, v, P); }
                    else if (b) { if (fK2) val += k2(W, w, P); }
                }
            }
        } else if (vi != wi) {                              // :133-200 (nu = inf: K2 at nu'; nu' = inf: K2 at nu)
            int x = vi ? w : v;
            if (fK1) {
                if (inB(W, nK1)) { val += k1(W, P); if (fK2 && inB(W, nK2b) && inF(x, nK2f)) val += k2(W, x, P); }
            } else if (fK2 && inB(W, nK2b) && inF(x, nK2f)) val += k2(W, x, P);
        } else {                                            // :203-225
            if (fK1 && inB(W, nK1)) val += k1(W, P);
        }
        return val;
    }
};

// ---- local Channel evaluator, src/channel.jl:220-339 --------------------------------
struct LocChan {
    int nK1, nK2b, nK2f, nK3b, nK3f;
    const cplx *K1, *K2, *K3;
    cplx k1(int W) const { return K1[posB(W, nK1)]; }
    cplx k2(int W, int v) const { return K2[posB(W, nK2b) + (size_t)(2 * nK2b - 1) * posF(v, nK2f)]; }
    cplx k3call(int W, int v, int w) const {  // call operator: 0 outside the box
        if (!(inB(W, nK3b) && inF(v, nK3f) && inF(w, nK3f))) return 0;
        int nB = 2 * nK3b - 1, nF = 2 * nK3f;
        return K3[posB(W, nK3b) + (size_t)nB * (posF(v, nK3f) + (size_t)nF * posF(w, nK3f))];
    }
    cplx eval(int W, int v, int w, bool fK1, bool fK2, bool fK3) const {
        cplx val = 0;
        bool vi = isinf_(v), wi = isinf_(w);
        if (!vi && !wi) {
            if (inB(W, nK1)) {
                if (fK1) val += k1(W);
                if (inB(W, nK2b)) {
                    bool a = inF(v, nK2f), b = inF(w, nK2f);
                    if (a && b) {
                        if (fK2) val += k2(W, v) + k2(W, w);
                        if (fK3) val += k3call(W, v, w);
                    } else if (a) { if (fK2) val += k2(W, v); }
                    else if (b)   { if (fK2) val += k2(W, w); }
                }
            }
        } else if (vi && !wi) {
            if (fK1) {
                if (inB(W, nK1)) {
                    val += k1(W);
                    if (fK2 && inB(W, nK2b) && inF(w, nK2f)) val += k2(W, w);
                }
            } else if (fK2 && inB(W, nK2b) && inF(w, nK2f)) val += k2(W, w);
        } else if (!vi && wi) {
            if (fK1) {
                if (inB(W, nK1)) {
                    val += k1(W);
                    if (fK2 && inB(W, nK2b) && inF(v, nK2f)) val += k2(W, v);
                }
            } else if (fK2 && inB(W, nK2b) && inF(v, nK2f)) val += k2(W, v);
        } else {
            if (fK1 && inB(W, nK1)) val += k1(W);
        }
        return val;
    }
};

// ---- _convert_channel, src/convention.jl:4-37 (frequencies typed, momenta untyped) -----
static inline void conv_freq(int W, int v, int w, int from, int to, int& W2, int& v2, int& w2) {
    if (from == to) { W2 = W; v2 = v; w2 = w; }
    else if (from == pCh && to == tCh) { W2 = F_minus_F(B_minus_F(W, v), w); v2 = w; w2 = v; }
    else if (from == pCh && to == aCh) { W2 = F_minus_F(v, w); v2 = B_minus_F(W, v); w2 = w; }
    else if (from == tCh && to == pCh) { W2 = F_plus_F(B_plus_F(W, v), w); v2 = w; w2 = v; }
    else if (from == tCh && to == aCh) { W2 = F_minus_F(w, v); v2 = B_plus_F(W, v); w2 = v; }
    else if (from == aCh && to == pCh) { W2 = F_plus_F(B_plus_F(W, w), v); v2 = B_plus_F(W, w); w2 = w; }
    else /* a -> t */                  { W2 = F_minus_F(v, w); v2 = w; w2 = B_plus_F(W, w); }
}
static inline void conv_mom(Mom P, Mom k, Mom q, int from, int to, Mom& P2, Mom& k2, Mom& q2) {
    if (from == to) { P2 = P; k2 = k; q2 = q; }
    else if (from == pCh && to == tCh) { P2 = P - k - q; k2 = q; q2 = k; }
    else if (from == pCh && to == aCh) { P2 = k - q; k2 = P - k; q2 = q; }
    else if (from == tCh && to == pCh) { P2 = P + k + q; k2 = q; q2 = k; }
    else if (from == tCh && to == aCh) { P2 = q - k; k2 = P + k; q2 = k; }
    else if (from == aCh && to == pCh) { P2 = P + q + k; k2 = P + q; q2 = q; }
    else /* a -> t */                  { P2 = k - q; k2 = q; q2 = P + q; }
}

// ---- full vertex evaluator ------------------------------------------------------------
struct VertexEval {
    const orc_vertex* V;
    int L, NP;

    NL2Chan nl2(const orc_level& lv, int r) const {
        NL2Chan c; c.nK1 = lv.nK1; c.nK2b = lv.nK2b; c.nK2f = lv.nK2f; c.nK3b = lv.nK3b; c.nK3f = lv.nK3f;
        c.L = L; c.NP = NP; c.K1 = lv.K1[r]; c.K2 = lv.K2[r]; c.K3 = lv.K3[r]; return c;
    }
    NLChan nl(const orc_level& lv, int r) const {
        NLChan c; c.nK1 = lv.nK1; c.nK2b = lv.nK2b; c.nK2f = lv.nK2f; c.nK3b = lv.nK3b; c.nK3f = lv.nK3f;
        c.L = L; c.NP = NP; c.K1 = lv.K1[r]; c.K2 = lv.K2[r]; c.K3 = lv.K3[r]; return c;
    }
    LocChan loc(const orc_level& lv, int r) const {
        LocChan c; c.nK1 = lv.nK1; c.nK2b = lv.nK2b; c.nK2f = lv.nK2f; c.nK3b = lv.nK3b; c.nK3f = lv.nK3f;
        c.K1 = lv.K1[r]; c.K2 = lv.K2[r]; c.K3 = lv.K3[r]; return c;
    }

    // RefVertex, src/refvertex.jl:90-216
    cplx core_call(const orc_level& lv, int which, int W, int v, int w) const {
        if (!(inB(W, lv.nK3b) && inF(v, lv.nK3f) && inF(w, lv.nK3f))) return 0;
        int nB = 2 * lv.nK3b - 1, nF = 2 * lv.nK3f;
        return lv.core[which][posB(W, lv.nK3b) + (size_t)nB * (posF(v, lv.nK3f) + (size_t)nF * posF(w, lv.nK3f))];
    }
    cplx core_eval(const orc_level& lv, int W, int v, int w, int Ch, int Sp) const {
        cplx U(lv.U_re, lv.U_im);
        if (isinf_(v) || isinf_(w)) return (Sp == xSp) ? -U : U;     // bare_vertex(F, Sp) :212-216
        if (Sp == dSp) return 2.0 * core_eval(lv, W, v, w, Ch, pSp) + core_eval(lv, W, v, w, Ch, xSp);
        if (Sp == pSp) {
            if (Ch == pCh) return core_call(lv, 0, W, v, w) + U;
            if (Ch == tCh) return core_call(lv, 2, W, v, w) + U;
            return -core_call(lv, 3, W, w, v) + U;
        } else {
            if (Ch == pCh) return core_call(lv, 1, W, v, w) - U;
            if (Ch == tCh) return core_call(lv, 3, W, v, w) - U;
            return -core_call(lv, 2, W, w, v) - U;
        }
    }

    // generic entry: vertex = chain starting at level `l`
    cplx eval(int l, int W, int v, int w, Mom P, Mom k, Mom q, int Ch, int Sp, Flags f) const {
        const orc_level& lv = V->lev[l];
        if (lv.type == LV_CORE) return core_eval(lv, W, v, w, Ch, Sp);   // kwargs ignored
        if (Sp == dSp) {    // src/vertex.jl:316-336, src/nonlocal/vertex.jl:188-211
            cplx val = 0;
            val += eval(l, W, v, w, P, k, q, Ch, pSp, f) * 2.0;
            val += eval(l, W, v, w, P, k, q, Ch, xSp, f);
            return val;
        }
        if (Sp == xSp) {    // src/vertex.jl:288-313, src/nonlocal/vertex.jl:157-185
            Flags g = f; g.gt = f.ga; g.ga = f.gt;
            if (Ch == pCh) return -eval(l, W, v, B_minus_F(W, w), P, k, P - q, pCh, pSp, g);
            if (Ch == tCh) return -eval(l, W, v, w, P, k, q, aCh, pSp, g);
            return -eval(l, W, v, w, P, k, q, tCh, pSp, g);
        }
        // ---- pSp ----
        if (is_mbe(lv.type)) return eval_mbe(l, W, v, w, P, k, q, Ch, f);
        if (lv.type == LV_LOCAL) return eval_local(l, W, v, w, Ch, f);
        if (lv.type == LV_NL) return eval_nl(l, W, v, w, P, k, q, Ch, f);
        return eval_nl2(l, W, v, w, P, k, q, Ch, f);
    }

    // ---- multi-boson-exchange vertices, src/boson_exchange.jl ------------------------------------------------------------
    // F(W, v, w, P, k, q, Ch, Cl): class Cl of channel Ch summed down the chain from level l (:270-330): the MeshFunction CALL of
    // every level's array (0 outside its mesh and for infinite frequencies; momenta folded; kSW = momentum mean), Lambda from the
    // RefVertex only (:171-232)
    cplx eval_class(int l, int W, int v, int w, Mom P, Mom k, Mom q, int Ch, int Cl) const {
        cplx val = 0;
        for (int j = l; j < V->nlev; j++) {
            const orc_level& lv = V->lev[j];
            if (lv.type == LV_CORE) {
                if (Cl == CL_LAMBDA) {
                    if (Ch == pCh) val += core_call(lv, 0, W, v, w);
                    else if (Ch == tCh) val += core_call(lv, 2, W, v, w);
                    else val -= core_call(lv, 3, W, w, v);
                }
                break;
            }
            if (Cl == CL_LAMBDA) continue;
            if (is_local_lv(lv.type)) {
                LocChan c = loc(lv, Ch);
                if (Cl == CL_K1) { if (inB(W, c.nK1)) val += c.k1(W); }
                else if (Cl == CL_K2) { if (inB(W, c.nK2b) && inF(v, c.nK2f)) val += c.k2(W, v); }
                else if (Cl == CL_K2P) { if (inB(W, c.nK2b) && inF(w, c.nK2f)) val += c.k2(W, w); }
                else val += c.k3call(W, v, w);
            } else {
                NL2Chan c = nl2(lv, Ch);
                if (Cl == CL_K1) { if (inB(W, c.nK1)) val += c.k1(W, P); }
                else if (Cl == CL_K2) { if (inB(W, c.nK2b) && inF(v, c.nK2f)) val += c.k2(W, v, P, k); }
                else if (Cl == CL_K2P) { if (inB(W, c.nK2b) && inF(w, c.nK2f)) val += c.k2(W, w, P, q); }
                else if (inB(W, c.nK3b) && inF(v, c.nK3f) && inF(w, c.nK3f)) val += c.k3(W, v, w, P);
            }
        }
        return val;
    }
    cplx bare_U() const { const orc_level& c = V->lev[V->nlev - 1]; return cplx(c.U_re, c.U_im); }
    // parallel spin component of an MBE vertex, :349-422; kSW arguments are explicit mesh averages of the whole expression (:481-560)
    cplx eval_mbe(int l, int W, int v, int w, Mom P, Mom k, Mom q, int Ch, Flags f) const {
        if (!is_local_lv(V->lev[l].type) && (k.sw || q.sw)) {
            cplx s = 0; int n = 0;
            for (int iq = 0; iq < (q.sw ? NP : 1); iq++) for (int ik = 0; ik < (k.sw ? NP : 1); ik++) {
                Mom kk = k.sw ? mk(ik % L, ik / L) : k, qq = q.sw ? mk(iq % L, iq / L) : q;
                s += eval_mbe(l, W, v, w, P, kk, qq, Ch, f); n++;
            }
            return s / (double)n;
        }
        const cplx U = bare_U();
        cplx val = U;
        auto sbe = [&](cplx K1, cplx K2, cplx K2p, cplx K3, cplx u) { return K1 + K2 + K2p + K2 * K2p / (u + K1) + K3; };
        if (f.gp) {
            int W2, v2, w2; Mom P2, k2, q2; conv_freq(W, v, w, Ch, pCh, W2, v2, w2); conv_mom(P, k, q, Ch, pCh, P2, k2, q2);
            val += sbe(eval_class(l, W2, v2, w2, P2, k2, q2, pCh, CL_K1), eval_class(l, W2, v2, w2, P2, k2, q2, pCh, CL_K2),
                       eval_class(l, W2, v2, w2, P2, k2, q2, pCh, CL_K2P), eval_class(l, W2, v2, w2, P2, k2, q2, pCh, CL_K3), U);
        }
        if (f.gt) {       // (tCh, pSp) = (D - M) / 2 with M = -(aCh, pSp), D = 2 (tCh, pSp) - (aCh, pSp), :382-400
            int W2, v2, w2; Mom P2, k2, q2; conv_freq(W, v, w, Ch, tCh, W2, v2, w2); conv_mom(P, k, q, Ch, tCh, P2, k2, q2);
            cplx K1 = -eval_class(l, W2, v2, w2, P2, k2, q2, aCh, CL_K1), K2 = -eval_class(l, W2, v2, w2, P2, k2, q2, aCh, CL_K2);
            cplx K2p = -eval_class(l, W2, v2, w2, P2, k2, q2, aCh, CL_K2P), K3 = -eval_class(l, W2, v2, w2, P2, k2, q2, aCh, CL_K3);
            val -= sbe(K1, K2, K2p, K3, -U) / 2.0;
            K1 = 2.0 * eval_class(l, W2, v2, w2, P2, k2, q2, tCh, CL_K1) + K1; K2 = 2.0 * eval_class(l, W2, v2, w2, P2, k2, q2, tCh, CL_K2) + K2;
            K2p = 2.0 * eval_class(l, W2, v2, w2, P2, k2, q2, tCh, CL_K2P) + K2p; K3 = 2.0 * eval_class(l, W2, v2, w2, P2, k2, q2, tCh, CL_K3) + K3;
            val += sbe(K1, K2, K2p, K3, U) / 2.0;
        }
        if (f.ga) {
            int W2, v2, w2; Mom P2, k2, q2; conv_freq(W, v, w, Ch, aCh, W2, v2, w2); conv_mom(P, k, q, Ch, aCh, P2, k2, q2);
            val += sbe(eval_class(l, W2, v2, w2, P2, k2, q2, aCh, CL_K1), eval_class(l, W2, v2, w2, P2, k2, q2, aCh, CL_K2),
                       eval_class(l, W2, v2, w2, P2, k2, q2, aCh, CL_K2P), eval_class(l, W2, v2, w2, P2, k2, q2, aCh, CL_K3), U);
        }
        val += eval_class(l, W, v, w, P, k, q, Ch, CL_LAMBDA);
        if (!f.F0) { Flags g = f; g.F0 = true; val -= eval(l + 1, W, v, w, P, k, q, Ch, pSp, g); }
        return val;
    }

    // NL_Vertex (bosonic momentum dependence only), src/nonlocal/vertex.jl:69-153 (Brillouin points), :213-377 (s-wave points)
    cplx eval_nl(int l, int W, int v, int w, Mom P, Mom k, Mom q, int Ch, Flags f) const {
        const orc_level& lv = V->lev[l];
        cplx val = 0;
        if (f.F0) val += eval(l + 1, W, v, w, P, k, q, Ch, pSp, ALLF());
        bool fl[3] = {f.gp, f.gt, f.ga};
        if (isinf_(v) || isinf_(w)) {                     // :112-153: only the own channel, at P
            if (fl[Ch]) val += nl(lv, Ch).eval(W, v, w, P);
            return val;
        }
        for (int r = 0; r < 3; r++) if (fl[r]) {
            int W2, v2, w2; conv_freq(W, v, w, Ch, r, W2, v2, w2);
            if (!k.sw && !q.sw) {                         // :69-107: converted transfer momentum
                Mom P2, k2, q2; conv_mom(P, k, q, Ch, r, P2, k2, q2);
                val += nl(lv, r).eval(W2, v2, w2, P2);
            } else {                                      // :213-377: own channel at P, other channels averaged over their momentum
                val += nl(lv, r).eval(W2, v2, w2, r == Ch ? P : SW());
            }
        }
        return val;
    }

    // local Vertex, src/vertex.jl:209-284 (momentum args dropped, src/nonlocal/channel.jl:230-256)
    cplx eval_local(int l, int W, int v, int w, int Ch, Flags f) const {
        const orc_level& lv = V->lev[l];
        cplx val = 0;
        Mom z = mk(0, 0);
        if (f.F0) val += eval(l + 1, W, v, w, z, z, z, Ch, pSp, ALLF());
        bool fl[3] = {f.gp, f.gt, f.ga};
        if (!isinf_(v) && !isinf_(w)) {
            for (int r = 0; r < 3; r++) if (fl[r]) {
                int W2, v2, w2; conv_freq(W, v, w, Ch, r, W2, v2, w2);
                val += loc(lv, r).eval(W2, v2, w2, true, true, true);
            }
        } else {
            if (fl[Ch]) val += loc(lv, Ch).eval(W, v, w, true, true, true);
        }
        return val;
    }

    // NL2_Vertex, src/nonlocal/vertex.jl:69-153 and src/nonlocal_2/vertex.jl:55-259
    cplx eval_nl2(int l, int W, int v, int w, Mom P, Mom k, Mom q, int Ch, Flags f) const {
        const orc_level& lv = V->lev[l];
        cplx val = 0;
        if (f.F0) val += eval(l + 1, W, v, w, P, k, q, Ch, pSp, ALLF());
        bool fl[3] = {f.gp, f.gt, f.ga};
        if (isinf_(v) || isinf_(w)) {                     // nonlocal/vertex.jl:112-153
            if (fl[Ch]) val += nl2(lv, Ch).eval(W, v, w, P, k, q, true, true, true);
            return val;
        }
        if (!k.sw && !q.sw) {                             // nonlocal/vertex.jl:69-107
            for (int r = 0; r < 3; r++) if (fl[r]) {
                int W2, v2, w2; conv_freq(W, v, w, Ch, r, W2, v2, w2);
                Mom P2, k2, q2; conv_mom(P, k, q, Ch, r, P2, k2, q2);
                val += nl2(lv, r).eval(W2, v2, w2, P2, k2, q2, true, true, true);
            }
            return val;
        }
        if (k.sw && q.sw) {                               // nonlocal_2/vertex.jl:207-259
            for (int r = 0; r < 3; r++) if (fl[r]) {
                if (r == Ch) val += nl2(lv, r).eval(W, v, w, P, k, q, true, true, true);
                else {
                    int W2, v2, w2; conv_freq(W, v, w, Ch, r, W2, v2, w2);
                    val += nl2(lv, r).eval(W2, v2, w2, SW(), SW(), SW(), true, true, true);
                }
            }
            return val;
        }
        // exactly one of k, q is kSW                      // nonlocal_2/vertex.jl:55-205
        for (int r = 0; r < 3; r++) if (fl[r]) {
            if (r == Ch) val += nl2(lv, r).eval(W, v, w, P, k, q, true, true, true);
            else {
                int W2, v2, w2; conv_freq(W, v, w, Ch, r, W2, v2, w2);
                NL2Chan c = nl2(lv, r);
                val += c.eval(W2, v2, w2, SW(), SW(), SW(), true, false, true);
                for (int j = 0; j < NP; j++) {              // x-fastest traversal of the P mesh
                    Mom kint = mk(j % L, j / L);
                    Mom P2, k2, q2;
                    if (k.sw) conv_mom(P, kint, q, Ch, r, P2, k2, q2);
                    else      conv_mom(P, k, kint, Ch, r, P2, k2, q2);
                    val += c.eval(W2, v2, w2, P2, k2, q2, false, true, false) / (double)NP;
                }
            }
        }
        return val;
    }
};

static inline int crossingF(int W, int w, int Ch) { return Ch == pCh ? B_minus_F(W, w) : w; }   // BSE_templates.jl:4-6
static inline Mom crossingK(Mom P, Mom q, int Ch) { return Ch == pCh ? P - q : q; }

// ---------------------------------------------------------------------------------
// index decoding helpers
struct K2Shape { int nb, nf, NP; int nB() const { return 2 * nb - 1; } int nF() const { return 2 * nf; }
                 size_t len() const { return (size_t)nB() * nF() * NP * NP; } };
static inline void decodeK2(int64_t idx, const K2Shape& s, int L, int& W, int& v, Mom& P, Mom& k) {
    int nB = s.nB(), nF = s.nF();
    int iW = idx % nB; idx /= nB; int iv = idx % nF; idx /= nF; int iP = idx % s.NP; int ik = idx / s.NP;
    W = iW - (s.nb - 1); v = iv - s.nf; P = mk(iP % L, iP / L); k = mk(ik % L, ik / L);
}

template <class Fn>
static void sg_apply(cplx* out, const orc_sg* SG, Fn diagram, int64_t c0 = 0, int64_t c1 = -1) {
    // SG(f, InitFunction(diagram)): representative evaluated, members = op(value)  
    if (c1 < 0) c1 = SG->nclasses;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t c = c0; c < c1; c++) {
        int64_t b = SG->offsets[c], e = SG->offsets[c + 1];
        cplx val = diagram(SG->index[b]);
        out[SG->index[b]] = val;
        for (int64_t j = b + 1; j < e; j++) out[SG->index[j]] = apply_op(SG->op[j], val);
    }
}

extern "C" {

int orc_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#endif
}

// single evaluation (used by the data-model tests)
void orc_eval_vertex(const orc_vertex* V, int L, int level, int W, int v, int w,
                     const int* P, const int* k, const int* q, int ksw, int qsw, int Ch, int Sp,
                     int fF0, int fgp, int fgt, int fga, cplx* out) {
    VertexEval E; E.V = V; E.L = L; E.NP = L * L;
    Flags f = {fF0 != 0, fgp != 0, fgt != 0, fga != 0};
    Mom kk = ksw ? SW() : mk(k[0], k[1]);
    Mom qq = qsw ? SW() : mk(q[0], q[1]);
    *out = E.eval(level, W, v, w, mk(P[0], P[1]), kk, qq, Ch, Sp, f);
}

// single NL2_Channel evaluation with K1/K2/K3 switches; sw bits: 1=P, 2=k, 4=q
void orc_eval_channel(const orc_vertex* V, int L, int level, int r, int W, int v, int w,
                      const int* P, const int* k, const int* q, int swbits, int fK1, int fK2, int fK3, cplx* out) {
    VertexEval E; E.V = V; E.L = L; E.NP = L * L;
    Mom PP = (swbits & 1) ? SW() : mk(P[0], P[1]);
    Mom kk = (swbits & 2) ? SW() : mk(k[0], k[1]);
    Mom qq = (swbits & 4) ? SW() : mk(q[0], q[1]);
    if (V->lev[level].type == LV_NL) { *out = E.nl(V->lev[level], r).eval(W, v, w, PP, fK1 != 0, fK2 != 0, fK3 != 0); return; }
    *out = E.nl2(V->lev[level], r).eval(W, v, w, PP, kk, qq, fK1 != 0, fK2 != 0, fK3 != 0);
}

// ---- BSE_K1!, src/nonlocal_2/BSEa/BSEa_K1.jl:2-58 ------------------------------------
// F0l: level of the chain `F` at which the solver's F0 starts (S.F0 === S.F.F0 -> 1); F0 may
// also be an independent vertex, hence passed separately.
void orc_bse_K1(cplx* K1, int nK1, const orc_vertex* F0, const orc_vertex* F, const orc_vertex* FL,
                const cplx* Pi0, const cplx* Pi, const orc_sg* SG, int sign, int Ch, int Sp, int is_mfRG,
                const orc_grid* g, int64_t c0, int64_t c1) {
    int L = g->L, NP = L * L;
    VertexEval EF0 = {F0, L, NP}, EF = {F, L, NP}, EFL = {FL, L, NP};
    int nBP = 2 * g->nPiB - 1, nFP = 2 * g->nPiF, nB1 = 2 * nK1 - 1;
    double T = g->T;
    Mom k0 = mk(0, 0);
    auto diagram = [&](int64_t idx) -> cplx {
        int W = (int)(idx % nB1) - (nK1 - 1); int iP = (int)(idx / nB1); Mom P = mk(iP % L, iP / L);
        int iW = posB(W, g->nPiB);
        cplx val = 0;
        for (int iq = 0; iq < NP; iq++) for (int iw = 0; iw < nFP; iw++) {   // eachindex(Pi0slice): w fastest
            int w = iw - g->nPiF; Mom q = mk(iq % L, iq / L);
            size_t pidx = iW + (size_t)nBP * (iw + (size_t)nFP * (iP + (size_t)NP * iq));
            if (is_mfRG) {
                cplx Fl  = EF0.eval(0, W, INF, w, P, k0, q, Ch, Sp, ALLF());
                cplx FLr = EFL.eval(0, W, crossingF(W, w, Ch), INF, P, crossingK(P, q, Ch), k0, Ch, Sp, ALLF());
                val += Fl * Pi0[pidx] * FLr;
            } else {
                cplx Fl  = EF.eval(0, W, INF, w, P, k0, q, Ch, Sp, ALLF());
                cplx F0r = EF0.eval(0, W, crossingF(W, w, Ch), INF, P, crossingK(P, q, Ch), k0, Ch, Sp, ALLF());
                cplx FLr = EFL.eval(0, W, crossingF(W, w, Ch), INF, P, crossingK(P, q, Ch), k0, Ch, Sp, ALLF());
                val += Fl * ((Pi[pidx] - Pi0[pidx]) * F0r + Pi[pidx] * FLr);
            }
        }
        return T * val / (double)NP * (double)sign;
    };
    sg_apply(K1, SG, diagram, c0, c1);
}

// ---- BSE_L_K2!, src/nonlocal_2/BSEa/BSEa_K2.jl:1-49 ----------------------------------
void orc_bse_L_K2(cplx* K2, int nK2b, int nK2f, const orc_vertex* F0, const orc_vertex* F,
                  const cplx* Pi0, const orc_sg* SG, int sign, int Ch, int Sp,
                  const orc_grid* g, int64_t c0, int64_t c1) {
    int L = g->L, NP = L * L;
    VertexEval EF0 = {F0, L, NP}, EF = {F, L, NP};
    int nBP = 2 * g->nPiB - 1, nFP = 2 * g->nPiF;
    K2Shape s = {nK2b, nK2f, NP};
    double T = g->T;
    Mom k0 = mk(0, 0);
    Flags fl = {false, Ch != pCh, Ch != tCh, Ch != aCh};
    auto diagram = [&](int64_t idx) -> cplx {
        int W, v; Mom P, k; decodeK2(idx, s, L, W, v, P, k);
        int iW = posB(W, g->nPiB), iP = kidx(P, L);
        cplx val = 0;
        for (int iq = 0; iq < NP; iq++) {
            Mom q = mk(iq % L, iq / L);
            for (int iw = 0; iw < 2 * nK2f; iw++) {           // omega over the K2 nu-mesh
                int w = iw - nK2f;
                cplx Gl  = EF.eval(0, W, v, crossingF(W, w, Ch), P, k, crossingK(P, q, Ch), Ch, Sp, fl);
                cplx F0r = EF0.eval(0, W, w, INF, P, q, k0, Ch, Sp, ALLF());
                size_t pidx = iW + (size_t)nBP * (posF(w, g->nPiF) + (size_t)nFP * (iP + (size_t)NP * iq));
                val += Gl * Pi0[pidx] * F0r;
            }
        }
        return T * val / (double)NP * (double)sign;
    };
    sg_apply(K2, SG, diagram, c0, c1);
}

// ---- BSE_K2!, src/nonlocal_2/BSEa/BSEa_K2.jl:55-138 (SG part only; FL add done by orc_axpy) --
void orc_bse_K2(cplx* K2, int nK2b, int nK2f, const orc_vertex* F0, const orc_vertex* F, const orc_vertex* FL,
                const cplx* Pi0, const cplx* Pi, const orc_sg* SG, int sign, int Ch, int Sp, int is_mfRG,
                const orc_grid* g, int64_t c0, int64_t c1) {
    int L = g->L, NP = L * L;
    VertexEval EF0 = {F0, L, NP}, EF = {F, L, NP}, EFL = {FL, L, NP};
    int nBP = 2 * g->nPiB - 1, nFP = 2 * g->nPiF;
    K2Shape s = {nK2b, nK2f, NP};
    double T = g->T;
    Mom k0 = mk(0, 0);
    auto diagram = [&](int64_t idx) -> cplx {
        int W, v; Mom P, k; decodeK2(idx, s, L, W, v, P, k);
        int iW = posB(W, g->nPiB), iP = kidx(P, L);
        cplx val = 0;
        for (int iq = 0; iq < NP; iq++) {
            Mom q = mk(iq % L, iq / L);
            for (int iw = 0; iw < nFP; iw++) {                 // omega over the bubble nu-mesh
                int w = iw - g->nPiF;
                size_t pidx = iW + (size_t)nBP * (iw + (size_t)nFP * (iP + (size_t)NP * iq));
                if (is_mfRG) {
                    int wc = crossingF(W, w, Ch); Mom qc = crossingK(P, q, Ch);
                    cplx Fl  = EF0.eval(0, W, v, wc, P, k, qc, Ch, Sp, ALLF()) - EF0.eval(0, W, INF, wc, P, k, qc, Ch, Sp, ALLF());
                    cplx FLr = EFL.eval(0, W, w, INF, P, q, k0, Ch, Sp, ALLF());
                    val += Fl * Pi0[pidx] * FLr;
                } else {
                    cplx Fl  = EF.eval(0, W, v, w, P, k, q, Ch, Sp, ALLF()) - EF.eval(0, W, INF, w, P, k, q, Ch, Sp, ALLF());
                    cplx F0r = EF0.eval(0, W, crossingF(W, w, Ch), INF, P, crossingK(P, q, Ch), k0, Ch, Sp, ALLF());
                    cplx FLr = EFL.eval(0, W, crossingF(W, w, Ch), INF, P, crossingK(P, q, Ch), k0, Ch, Sp, ALLF());
                    val += Fl * ((Pi[pidx] - Pi0[pidx]) * F0r + Pi[pidx] * FLr);
                }
            }
        }
        return T * val / (double)NP * (double)sign;
    };
    sg_apply(K2, SG, diagram, c0, c1);
}

// ---- BSE_K1_new!, src/nonlocal_2/BSEa/BSEa_K1.jl:62-113: K1 = (U + K1 + K2') Pi U ------
// U = bare_vertex(F, Sp) = +U for pSp and dSp (src/refvertex.jl:213-215, src/vertex.jl:340)
void orc_bse_K1_new(cplx* K1, int nK1, const orc_vertex* F0, const orc_vertex* F,
                    const cplx* Pi0, const cplx* Pi, const orc_sg* SG, int sign, int Ch, int Sp, int is_mfRG,
                    const orc_grid* g, int64_t c0, int64_t c1) {
    int L = g->L, NP = L * L;
    VertexEval EF0 = {F0, L, NP}, EF = {F, L, NP};
    int nBP = 2 * g->nPiB - 1, nFP = 2 * g->nPiF, nB1 = 2 * nK1 - 1;
    double T = g->T;
    const orc_level& core = F->lev[F->nlev - 1];
    cplx U = cplx(core.U_re, core.U_im) * (Sp == xSp ? -1.0 : 1.0);
    Mom k0 = mk(0, 0);
    auto diagram = [&](int64_t idx) -> cplx {
        int W = (int)(idx % nB1) - (nK1 - 1); int iP = (int)(idx / nB1); Mom P = mk(iP % L, iP / L);
        int iW = posB(W, g->nPiB);
        cplx val = 0;
        for (int iq = 0; iq < NP; iq++) for (int iw = 0; iw < nFP; iw++) {
            int w = iw - g->nPiF; Mom q = mk(iq % L, iq / L);
            size_t pidx = iW + (size_t)nBP * (iw + (size_t)nFP * (iP + (size_t)NP * iq));
            cplx Fl  = EF.eval(0, W, INF, w, P, k0, q, Ch, Sp, ALLF());
            cplx F0l = EF0.eval(0, W, INF, w, P, k0, q, Ch, Sp, ALLF());
            if (is_mfRG) val += (Fl - F0l) * Pi[pidx] * U;
            else         val += (Fl * Pi[pidx] - F0l * Pi0[pidx]) * U;
        }
        return T * val / (double)NP * (double)sign;
    };
    sg_apply(K1, SG, diagram, c0, c1);
}

// ---- BSE_K2_new!, src/nonlocal_2/BSEa/BSEa_K2.jl:142-216: K2 = (K2 + K3 + ...) Pi U; omega runs over the K2 nu-mesh
void orc_bse_K2_new(cplx* K2, int nK2b, int nK2f, const orc_vertex* F0, const orc_vertex* F,
                    const cplx* Pi0, const cplx* Pi, const orc_sg* SG, int sign, int Ch, int Sp, int is_mfRG,
                    const orc_grid* g, int64_t c0, int64_t c1) {
    int L = g->L, NP = L * L;
    VertexEval EF0 = {F0, L, NP}, EF = {F, L, NP};
    int nBP = 2 * g->nPiB - 1, nFP = 2 * g->nPiF;
    K2Shape s = {nK2b, nK2f, NP};
    double T = g->T;
    const orc_level& core = F->lev[F->nlev - 1];
    cplx U = cplx(core.U_re, core.U_im) * (Sp == xSp ? -1.0 : 1.0);
    auto diagram = [&](int64_t idx) -> cplx {
        int W, v; Mom P, k; decodeK2(idx, s, L, W, v, P, k);
        int iW = posB(W, g->nPiB), iP = kidx(P, L);
        cplx val = 0;
        for (int iq = 0; iq < NP; iq++) {
            Mom q = mk(iq % L, iq / L);
            for (int iw = 0; iw < 2 * nK2f; iw++) {
                int w = iw - nK2f;
                size_t pidx = iW + (size_t)nBP * (posF(w, g->nPiF) + (size_t)nFP * (iP + (size_t)NP * iq));
                cplx Fl  = EF.eval(0, W, v, w, P, k, q, Ch, Sp, ALLF())  - EF.eval(0, W, INF, w, P, k, q, Ch, Sp, ALLF());
                cplx F0l = EF0.eval(0, W, v, w, P, k, q, Ch, Sp, ALLF()) - EF0.eval(0, W, INF, w, P, k, q, Ch, Sp, ALLF());
                if (is_mfRG) val += (Fl - F0l) * Pi[pidx] * U;
                else         val += (Fl * Pi[pidx] - F0l * Pi0[pidx]) * U;
            }
        }
        return T * val / (double)NP * (double)sign;
    };
    sg_apply(K2, SG, diagram, c0, c1);
}

// ---- BSE_K1_1loop!, src/nonlocal_2/BSEa/BSE_1loop.jl:2-56 -------------------------------
void orc_bse_K1_1loop(cplx* K1, int nK1, const orc_vertex* F0, const orc_vertex* F, const orc_vertex* FL,
                      const cplx* Pi0, const cplx* Pi, const orc_sg* SG, int sign, int Ch, int Sp, int is_mfRG,
                      const orc_grid* g, int64_t c0, int64_t c1) {
    int L = g->L, NP = L * L;
    VertexEval EF0 = {F0, L, NP}, EF = {F, L, NP}, EFL = {FL, L, NP};
    int nBP = 2 * g->nPiB - 1, nFP = 2 * g->nPiF, nB1 = 2 * nK1 - 1;
    double T = g->T;
    Mom k0 = mk(0, 0);
    auto diagram = [&](int64_t idx) -> cplx {
        int W = (int)(idx % nB1) - (nK1 - 1); int iP = (int)(idx / nB1); Mom P = mk(iP % L, iP / L);
        int iW = posB(W, g->nPiB);
        cplx val = 0;
        for (int iq = 0; iq < NP; iq++) for (int iw = 0; iw < nFP; iw++) {
            int w = iw - g->nPiF; Mom q = mk(iq % L, iq / L);
            size_t pidx = iW + (size_t)nBP * (iw + (size_t)nFP * (iP + (size_t)NP * iq));
            if (is_mfRG) {
                cplx Fl  = EF0.eval(0, W, INF, w, P, k0, q, Ch, Sp, ALLF());
                cplx FLr = EFL.eval(0, W, crossingF(W, w, Ch), INF, P, crossingK(P, q, Ch), k0, Ch, Sp, ALLF());
                val += Fl * Pi0[pidx] * FLr;
            } else {
                cplx Fl  = EF.eval(0, W, INF, w, P, k0, q, Ch, Sp, ALLF());
                cplx F0r = EF0.eval(0, W, crossingF(W, w, Ch), INF, P, crossingK(P, q, Ch), k0, Ch, Sp, ALLF());
                val += Fl * (Pi[pidx] - Pi0[pidx]) * F0r;
            }
        }
        return T * val / (double)NP * (double)sign;
    };
    sg_apply(K1, SG, diagram, c0, c1);
}

// ---- BSE_K2_1loop!, src/nonlocal_2/BSEa/BSE_1loop.jl:59-124 (SG part only; FL add done by the caller) ----
void orc_bse_K2_1loop(cplx* K2, int nK2b, int nK2f, const orc_vertex* F0, const orc_vertex* F, const orc_vertex* FL,
                      const cplx* Pi0, const cplx* Pi, const orc_sg* SG, int sign, int Ch, int Sp, int is_mfRG,
                      const orc_grid* g, int64_t c0, int64_t c1) {
    int L = g->L, NP = L * L;
    VertexEval EF0 = {F0, L, NP}, EF = {F, L, NP}, EFL = {FL, L, NP};
    int nBP = 2 * g->nPiB - 1, nFP = 2 * g->nPiF;
    K2Shape s = {nK2b, nK2f, NP};
    double T = g->T;
    Mom k0 = mk(0, 0);
    auto diagram = [&](int64_t idx) -> cplx {
        int W, v; Mom P, k; decodeK2(idx, s, L, W, v, P, k);
        int iW = posB(W, g->nPiB), iP = kidx(P, L);
        cplx val = 0;
        for (int iq = 0; iq < NP; iq++) {
            Mom q = mk(iq % L, iq / L);
            for (int iw = 0; iw < nFP; iw++) {
                int w = iw - g->nPiF;
                size_t pidx = iW + (size_t)nBP * (iw + (size_t)nFP * (iP + (size_t)NP * iq));
                if (is_mfRG) {
                    int wc = crossingF(W, w, Ch); Mom qc = crossingK(P, q, Ch);
                    cplx Fl  = EF0.eval(0, W, v, wc, P, k, qc, Ch, Sp, ALLF()) - EF0.eval(0, W, INF, wc, P, k, qc, Ch, Sp, ALLF());
                    cplx FLr = EFL.eval(0, W, w, INF, P, q, k0, Ch, Sp, ALLF());
                    val += Fl * Pi0[pidx] * FLr;
                } else {
                    cplx Fl  = EF.eval(0, W, v, w, P, k, q, Ch, Sp, ALLF()) - EF.eval(0, W, INF, w, P, k, q, Ch, Sp, ALLF());
                    cplx F0r = EF0.eval(0, W, crossingF(W, w, Ch), INF, P, crossingK(P, q, Ch), k0, Ch, Sp, ALLF());
                    val += Fl * (Pi[pidx] - Pi0[pidx]) * F0r;
                }
            }
        }
        return T * val / (double)NP * (double)sign;
    };
    sg_apply(K2, SG, diagram, c0, c1);
}

// Pi[W, w, P, kSW] -> mean over the 4th axis (src/nonlocal/swave.jl:91-105)
static inline cplx pi_sw(const cplx* Pi, const orc_grid* g, int W, int w, int iP) {
    int NP = g->L * g->L, nBP = 2 * g->nPiB - 1, nFP = 2 * g->nPiF;
    size_t base = posB(W, g->nPiB) + (size_t)nBP * (posF(w, g->nPiF) + (size_t)nFP * iP);
    if (g->swave) return Pi[base];          // NL_MF_Pi[W, w, P]: src/nonlocal/BSEa/BSEa_K3.jl:28,80-81
    size_t sk = (size_t)nBP * nFP * NP;
    cplx s = 0; for (int i = 0; i < NP; i++) s += Pi[base + sk * i];
    return s / (double)NP;
}

struct K3Shape { int nb, nf, NP; int nB() const { return 2 * nb - 1; } int nF() const { return 2 * nf; }
    size_t at(int W, int v, int w, int iP) const { return posB(W, nb) + (size_t)nB() * (posF(v, nf) + (size_t)nF() * (posF(w, nf) + (size_t)nF() * iP)); } };
static inline void decodeK3(int64_t idx, const K3Shape& s, int& W, int& v, int& w, int& iP) {
    int nB = s.nB(), nF = s.nF();
    int iW = idx % nB; idx /= nB; int iv = idx % nF; idx /= nF; int iw = idx % nF; iP = (int)(idx / nF);
    W = iW - (s.nb - 1); v = iv - s.nf; w = iw - s.nf;
}

// ---- BSE_L_K3!, src/nonlocal_2/BSEa/BSEa_K3.jl:1-40 -----------------------------------
void orc_bse_L_K3(cplx* K3, int nK3b, int nK3f, const cplx* cache_G, const cplx* cache_F0, const cplx* Pi0,
                  const orc_sg* SG, int sign, const orc_grid* g) {
    int NP = g->L * g->L;
    K3Shape s = {nK3b, nK3f, NP};
    double T = g->T;
    auto diagram = [&](int64_t idx) -> cplx {
        int W, v, vp, iP; decodeK3(idx, s, W, v, vp, iP);
        cplx val = 0;
        for (int iw = 0; iw < s.nF(); iw++) {
            int w = iw - nK3f;
            cplx P0 = pi_sw(Pi0, g, W, w, iP);
            val += cache_G[s.at(W, v, w, iP)] * P0 * cache_F0[s.at(W, w, vp, iP)];
        }
        return T * val * (double)sign;
    };
    sg_apply(K3, SG, diagram);
}

// ---- BSE_K3!, src/nonlocal_2/BSEa/BSEa_K3.jl:43-128 -----------------------------------
// FLt / FLa: K3 arrays of FL.gamma_t and FL.gamma_a (for Ch = t), FLown: K3 of FL.gamma_Ch
void orc_bse_K3(cplx* K3, int nK3b, int nK3f, const cplx* FLown, const cplx* FLt, const cplx* FLa,
                const cplx* cache_G, const cplx* cache_F, const cplx* cache_F0,
                const cplx* Pi0, const cplx* Pi, const orc_sg* SG, int sign1, int sign2, int Ch, int is_mfRG,
                const orc_grid* g) {
    int NP = g->L * g->L;
    K3Shape s = {nK3b, nK3f, NP};
    double T = g->T;
    auto diagram = [&](int64_t idx) -> cplx {
        int W, v, vp, iP; decodeK3(idx, s, W, v, vp, iP);
        cplx val = 0;
        for (int iw = 0; iw < s.nF(); iw++) {
            int w = iw - nK3f;
            cplx Gs = (Ch == pCh) ? cache_G[s.at(W, w, vp, iP)] : cache_G[s.at(W, vp, w, iP)];
            cplx Fs = cache_F[s.at(W, v, w, iP)];
            cplx F0s = cache_F0[s.at(W, w, vp, iP)];
            cplx P0 = pi_sw(Pi0, g, W, w, iP);
            cplx P1 = pi_sw(Pi, g, W, w, iP);
            int wc = crossingF(W, w, Ch);
            if (is_mfRG) {
                val += Fs * P0 * Gs * (double)sign1;
                if (inF(wc, nK3f)) {
                    if (Ch == aCh || Ch == pCh) val += Fs * P0 * FLown[s.at(W, wc, vp, iP)] * (double)sign2;
                    else val += Fs * P0 * (2.0 * FLt[s.at(W, wc, vp, iP)] - FLa[s.at(W, wc, vp, iP)]) * (double)sign2;
                }
            } else {
                val += Fs * ((P1 - P0) * F0s + P1 * Gs) * (double)sign1;
                if (inF(wc, nK3f)) {
                    if (Ch == aCh || Ch == pCh) val += Fs * P1 * FLown[s.at(W, wc, vp, iP)] * (double)sign2;
                    else val += Fs * P1 * (2.0 * FLt[s.at(W, wc, vp, iP)] - FLa[s.at(W, wc, vp, iP)]) * (double)sign2;
                }
            }
        }
        if (Ch == aCh || Ch == pCh) return T * val + FLown[s.at(W, v, vp, iP)];
        return T * val + 2.0 * FLt[s.at(W, v, vp, iP)] - FLa[s.at(W, v, vp, iP)];
    };
    sg_apply(K3, SG, diagram);
}

// ---- BSE_K3_1loop!, src/nonlocal_2/BSEa/BSE_1loop.jl:123-199: no FL.K3 added to the result -----
void orc_bse_K3_1loop(cplx* K3, int nK3b, int nK3f, const cplx* FLown, const cplx* FLt, const cplx* FLa,
                      const cplx* cache_G, const cplx* cache_F, const cplx* cache_F0,
                      const cplx* Pi0, const cplx* Pi, const orc_sg* SG, int sign1, int sign2, int Ch, int is_mfRG,
                      const orc_grid* g) {
    int NP = g->L * g->L;
    K3Shape s = {nK3b, nK3f, NP};
    double T = g->T;
    auto diagram = [&](int64_t idx) -> cplx {
        int W, v, vp, iP; decodeK3(idx, s, W, v, vp, iP);
        cplx val = 0;
        for (int iw = 0; iw < s.nF(); iw++) {
            int w = iw - nK3f;
            cplx Gs = (Ch == pCh) ? cache_G[s.at(W, w, vp, iP)] : cache_G[s.at(W, vp, w, iP)];
            cplx Fs = cache_F[s.at(W, v, w, iP)];
            cplx F0s = cache_F0[s.at(W, w, vp, iP)];
            cplx P0 = pi_sw(Pi0, g, W, w, iP);
            cplx P1 = pi_sw(Pi, g, W, w, iP);
            if (is_mfRG) {
                val += Fs * P0 * Gs * (double)sign1;
                int wc = crossingF(W, w, Ch);
                if (inF(wc, nK3f)) {
                    if (Ch == aCh || Ch == pCh) val += Fs * P0 * FLown[s.at(W, wc, vp, iP)] * (double)sign2;
                    else val += Fs * P0 * (2.0 * FLt[s.at(W, wc, vp, iP)] - FLa[s.at(W, wc, vp, iP)]) * (double)sign2;
                }
            } else {
                val += Fs * (P1 - P0) * F0s * (double)sign1;
            }
        }
        return T * val;
    };
    sg_apply(K3, SG, diagram);
}

// ---- build_K3_cache!, src/nonlocal_2/build_K3_cache.jl:18-94 ---------------------------
// caches: 0 Gpx, 1 F0p, 2 F0a, 3 F0t, 4 Gpp, 5 Ga, 6 Gt, 7 Fp, 8 Fa, 9 Ft
void orc_build_K3_cache(cplx* const* cache, int nK3b, int nK3f, const orc_vertex* F0, const orc_vertex* F,
                        const orc_grid* g, int64_t i0, int64_t i1) {
    int L = g->L, NP = L * L;
    VertexEval EF0 = {F0, L, NP}, EF = {F, L, NP};
    K3Shape s = {nK3b, nK3f, NP};
    int64_t len = (int64_t)s.nB() * s.nF() * s.nF() * NP;
    if (i1 < 0) i1 = len;
    Mom sw = SW();
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = i0; i < i1; i++) {
        int W, a, b, iP; decodeK3(i, s, W, a, b, iP); Mom P = mk(iP % L, iP / L);
        {   // (W, w, vp, P) = (W, a, b, P)
            Flags f = {false, false, true, true};
            cache[0][i] = EF.eval(0, W, a, b, P, sw, sw, pCh, xSp, f);
            cache[1][i] = EF0.eval(0, W, a, b, P, sw, sw, pCh, xSp, ALLF()) - EF0.eval(0, W, a, INF, P, sw, sw, pCh, xSp, ALLF());
            cache[2][i] = EF0.eval(0, W, a, b, P, sw, sw, aCh, pSp, ALLF()) - EF0.eval(0, W, a, INF, P, sw, sw, aCh, pSp, ALLF());
            cache[3][i] = EF0.eval(0, W, a, b, P, sw, sw, tCh, pSp, ALLF()) - EF0.eval(0, W, a, INF, P, sw, sw, tCh, pSp, ALLF());
            cache[3][i] = 2.0 * cache[3][i] - cache[2][i];
        }
        {   // (W, v, w, P) = (W, a, b, P)
            Flags fp = {false, false, true, true}, fa = {false, true, true, false}, ft = {false, true, false, true};
            cache[4][i] = EF.eval(0, W, a, b, P, sw, sw, pCh, pSp, fp);
            cache[5][i] = EF.eval(0, W, a, b, P, sw, sw, aCh, pSp, fa);
            cache[6][i] = EF.eval(0, W, a, b, P, sw, sw, tCh, pSp, ft);
            Flags op = {true, true, false, false}, oa = {true, false, false, true}, ot = {true, false, true, false};
            cache[7][i] = (EF.eval(0, W, a, b, P, sw, sw, pCh, pSp, op) - EF.eval(0, W, INF, b, P, sw, sw, pCh, pSp, op)) + cache[4][i];
            cache[8][i] = (EF.eval(0, W, a, b, P, sw, sw, aCh, pSp, oa) - EF.eval(0, W, INF, b, P, sw, sw, aCh, pSp, oa)) + cache[5][i];
            cache[9][i] = (EF.eval(0, W, a, b, P, sw, sw, tCh, pSp, ot) - EF.eval(0, W, INF, b, P, sw, sw, tCh, pSp, ot)) + cache[6][i];
            cache[6][i] = cache[6][i] * 2.0 - cache[5][i];
            cache[9][i] = cache[9][i] * 2.0 - cache[8][i];
        }
    }
}

// ---- build_K3_cache! for MBE vertices: src/boson_exchange.jl:947-1017 (NL2_MBEVertex) and :599-668 (MBEVertex, = the same on a
// 1 x 1 mesh).  cache_F0* / cache_F* hold the channel-U-irreducible vertex T_r = (I_r - U) + M_r, cache_G* the irreducible
// finite-difference vertex F - F.F0 ----
void orc_build_K3_cache_mbe(cplx* const* cache, int nK3b, int nK3f, const orc_vertex* F0, const orc_vertex* F,
                            const orc_grid* g, int64_t i0, int64_t i1) {
    int L = g->L, NP = L * L;
    VertexEval EF0 = {F0, L, NP}, EF = {F, L, NP};
    K3Shape s = {nK3b, nK3f, NP};
    int64_t len = (int64_t)s.nB() * s.nF() * s.nF() * NP;
    if (i1 < 0) i1 = len;
    Mom sw = SW();
    const cplx U = EF.bare_U();
    const Flags np_ = {true, false, true, true}, na = {true, true, true, false}, nt = {true, true, false, true};      // gamma_r = false
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t i = i0; i < i1; i++) {
        int W, a, b, iP; decodeK3(i, s, W, a, b, iP); Mom P = mk(iP % L, iP / L);
        // (W, w, vp, P) = (W, a, b, P): vertices multiplied by bubbles to the left
        cache[0][i] = EF.eval(0, W, a, b, P, sw, sw, pCh, xSp, np_) - EF.eval(1, W, a, b, P, sw, sw, pCh, xSp, np_);
        cache[1][i] = EF0.eval(0, W, a, b, P, sw, sw, pCh, xSp, np_) + U - EF0.eval_class(0, W, a, B_minus_F(W, b), P, sw, sw, pCh, CL_K3);
        cache[2][i] = EF0.eval(0, W, a, b, P, sw, sw, aCh, pSp, na) - U + EF0.eval_class(0, W, a, b, P, sw, sw, aCh, CL_K3);
        cache[3][i] = EF0.eval(0, W, a, b, P, sw, sw, tCh, pSp, nt) - U + EF0.eval_class(0, W, a, b, P, sw, sw, tCh, CL_K3);
        cache[3][i] = 2.0 * cache[3][i] - cache[2][i];
        // (W, v, w, P) = (W, a, b, P): vertices multiplied by bubbles from the right
        cplx Fp = EF.eval(0, W, a, b, P, sw, sw, pCh, pSp, np_), Fa = EF.eval(0, W, a, b, P, sw, sw, aCh, pSp, na), Ft = EF.eval(0, W, a, b, P, sw, sw, tCh, pSp, nt);
        cache[4][i] = Fp - EF.eval(1, W, a, b, P, sw, sw, pCh, pSp, np_);
        cache[5][i] = Fa - EF.eval(1, W, a, b, P, sw, sw, aCh, pSp, na);
        cache[6][i] = Ft - EF.eval(1, W, a, b, P, sw, sw, tCh, pSp, nt);
        cache[7][i] = Fp - U + EF.eval_class(0, W, a, b, P, sw, sw, pCh, CL_K3);
        cache[8][i] = Fa - U + EF.eval_class(0, W, a, b, P, sw, sw, aCh, CL_K3);
        cache[9][i] = Ft - U + EF.eval_class(0, W, a, b, P, sw, sw, tCh, CL_K3);
        cache[6][i] = cache[6][i] * 2.0 - cache[5][i];
        cache[9][i] = cache[9][i] * 2.0 - cache[8][i];
    }
}

// evaluator of one asymptotic class for the tests: F(W, v, w, P, k, q, Ch, Cl) from level `level` of the chain
void orc_eval_class(const orc_vertex* V, int L, int level, int W, int v, int w, const int* P, const int* k, const int* q, int ksw, int qsw,
                    int Ch, int Cl, cplx* out) {
    VertexEval E; E.V = V; E.L = L; E.NP = L * L;
    *out = E.eval_class(level, W, v, w, mk(P[0], P[1]), ksw ? SW() : mk(k[0], k[1]), qsw ? SW() : mk(q[0], q[1]), Ch, Cl);
}

// ---- build_K3_cache_mfRG!, src/nonlocal_2/build_K3_cache.jl:97-164 ---------------------
// F: chain of S.F (level 0) ; S.F.F0 = chain from level 1 ; F0: S.F0 (passed separately)
void orc_build_K3_cache_mfRG(cplx* const* cache, int nK3b, int nK3f, const orc_vertex* F0, const orc_vertex* F,
                             const orc_sg* SGpp3, const orc_sg* SGph3, int is_first, const orc_grid* g) {
    int L = g->L, NP = L * L;
    VertexEval EF0 = {F0, L, NP}, EF = {F, L, NP};
    K3Shape s = {nK3b, nK3f, NP};
    int64_t len = (int64_t)s.nB() * s.nF() * s.nF() * NP;
    Mom sw = SW();
    auto mk_diag = [&](int Ch, int Sp, Flags f) {
        return [&, Ch, Sp, f](int64_t i) -> cplx {
            int W, a, b, iP; decodeK3(i, s, W, a, b, iP); Mom P = mk(iP % L, iP / L);
            return EF.eval(0, W, a, b, P, sw, sw, Ch, Sp, f) - EF.eval(1, W, a, b, P, sw, sw, Ch, Sp, f);
        };
    };
    Flags fp = {true, false, true, true}, fa = {true, true, true, false}, ft = {true, true, false, true};
    sg_apply(cache[0], SGpp3, mk_diag(pCh, xSp, fp));
    sg_apply(cache[4], SGpp3, mk_diag(pCh, pSp, fp));
    sg_apply(cache[5], SGph3, mk_diag(aCh, pSp, fa));
    sg_apply(cache[6], SGph3, mk_diag(tCh, pSp, ft));
    for (int64_t i = 0; i < len; i++) cache[6][i] = cache[6][i] * 2.0 - cache[5][i];
    if (is_first) {
        auto mk_diag0 = [&](int Ch) {
            return [&, Ch](int64_t i) -> cplx {
                int W, a, b, iP; decodeK3(i, s, W, a, b, iP); Mom P = mk(iP % L, iP / L);
                return EF0.eval(0, W, a, b, P, sw, sw, Ch, pSp, ALLF()) - EF0.eval(0, W, INF, b, P, sw, sw, Ch, pSp, ALLF());
            };
        };
        sg_apply(cache[7], SGpp3, mk_diag0(pCh));
        sg_apply(cache[8], SGph3, mk_diag0(aCh));
        sg_apply(cache[9], SGph3, mk_diag0(tCh));
        for (int64_t i = 0; i < len; i++) cache[9][i] = cache[9][i] * 2.0 - cache[8][i];
    }
}

// Quirk toggle.  1 (default) = as coded: F(...; own gamma, F0 = true) - F.F0(...; own gamma),
// which for a nested nonlocal F0 also picks up F0's cross channels.  0 = as commented: own gamma of F only.
static int g_quirk_E2 = 1;
void orc_set_quirk_E2(int on) { g_quirk_E2 = on; }

// ---- SDE_channel_L_pp!/ph!, src/nonlocal_2/SDE.jl:3-151 --------------------------------
// level: the vertex F of this SDE recursion step = chain V from `level` (F.F0 = level + 1).
void orc_sde_L(cplx* Lout, int nK2b, int nK2f, const orc_vertex* V, int level, const cplx* Pi,
               const orc_sg* SG, int is_pp, const orc_grid* g, int64_t c0, int64_t c1) {
    int L = g->L, NP = L * L;
    VertexEval E = {V, L, NP};
    int nBP = 2 * g->nPiB - 1, nFP = 2 * g->nPiF;
    K2Shape s = {nK2b, nK2f, NP};
    double T = g->T;
    // bare_vertex(F) = U of the terminating core
    const orc_level& core = V->lev[V->nlev - 1];
    cplx U(core.U_re, core.U_im);
    bool is_core = V->lev[level].type == LV_CORE;
    auto diagram = [&](int64_t idx) -> cplx {
        int W, v; Mom P, k; decodeK2(idx, s, L, W, v, P, k);
        int iW = posB(W, g->nPiB), iP = kidx(P, L);
        cplx val = 0;
        for (int iq = 0; iq < NP; iq++) for (int iw = 0; iw < nFP; iw++) {
            int w = iw - g->nPiF; Mom q = mk(iq % L, iq / L);
            size_t pidx = iW + (size_t)nBP * (iw + (size_t)nFP * (iP + (size_t)NP * iq));
            if (is_core) {
                if (is_pp) val += U * Pi[pidx] * (E.core_eval(V->lev[level], W, B_minus_F(W, w), v, pCh, pSp) - U);
                else val += U * Pi[pidx] * (E.core_eval(V->lev[level], W, v, w, aCh, pSp) + E.core_eval(V->lev[level], W, v, w, tCh, pSp) - U - U);
            } else if (!g_quirk_E2) {
                if (is_pp) {
                    Flags f = {false, true, false, false};
                    val += U * Pi[pidx] * E.eval(level, W, B_minus_F(W, w), v, P, P - q, k, pCh, pSp, f);
                } else {
                    Flags fa = {false, false, false, true}, ft = {false, false, true, false};
                    val += U * Pi[pidx] * (E.eval(level, W, v, w, P, k, q, aCh, pSp, fa) + E.eval(level, W, v, w, P, k, q, tCh, pSp, ft));
                }
            } else if (is_pp) {
                Flags f = {true, true, false, false};
                int wc = B_minus_F(W, w); Mom qc = P - q;
                val += U * Pi[pidx] * ( E.eval(level, W, wc, v, P, qc, k, pCh, pSp, f)
                                      - E.eval(level + 1, W, wc, v, P, qc, k, pCh, pSp, f));
            } else {
                Flags fa = {true, false, false, true}, ft = {true, false, true, false};
                val += U * Pi[pidx] * ( E.eval(level, W, v, w, P, k, q, aCh, pSp, fa)
                                      + E.eval(level, W, v, w, P, k, q, tCh, pSp, ft)
                                      - E.eval(level + 1, W, v, w, P, k, q, aCh, pSp, fa)
                                      - E.eval(level + 1, W, v, w, P, k, q, tCh, pSp, ft));
            }
        }
        return T * val / (double)NP;
    };
    sg_apply(Lout, SG, diagram, c0, c1);
}

// ---- small DFT helpers (FFTW conventions: forward e^{-i}, backward e^{+i}, unnormalised) --
static void dft_axis(cplx* data, size_t pre, int n, size_t post, int sgn) {
    // data viewed as [pre][n][post] column-major => index = a + pre*(j + n*b)
    std::vector<cplx> tw(n);
    for (int j = 0; j < n; j++) tw[j] = std::polar(1.0, sgn * 2.0 * M_PI * j / n);
#pragma omp parallel
    {
        std::vector<cplx> tmp(n);
#pragma omp for collapse(2)
        for (size_t b = 0; b < post; b++) for (size_t a = 0; a < pre; a++) {
            cplx* p = data + a + pre * n * b;
            for (int k = 0; k < n; k++) {
                cplx s = 0;
                for (int j = 0; j < n; j++) s += p[pre * j] * tw[(int)(((int64_t)j * k) % n)];
                tmp[k] = s;
            }
            for (int k = 0; k < n; k++) p[pre * k] = tmp[k];
        }
    }
}

// ---- bubbles_real_space!, src/nonlocal_2/bubble.jl:42-122 ------------------------------
void orc_bubbles_real_space(cplx* Pipp, cplx* Piph, const cplx* G, int nG, int LG, const orc_grid* g) {
    int L = g->L, NP = L * L;
    int nBP = 2 * g->nPiB - 1, nFP = 2 * g->nPiF, nGf = 2 * nG;
    size_t len = (size_t)nBP * nFP * NP * NP;
    std::fill(Pipp, Pipp + len, cplx(0)); std::fill(Piph, Piph + len, cplx(0));
    std::vector<cplx> GR(G, G + (size_t)nGf * LG * LG);
    dft_axis(GR.data(), nGf, LG, LG, -1); dft_axis(GR.data(), (size_t)nGf * LG, LG, 1, -1);
    for (auto& x : GR) x /= (double)(LG * LG);
    int h = L / 2;
    auto Gcall = [&](int n, int ix, int iy) -> cplx { return inF(n, nG) ? GR[posF(n, nG) + (size_t)nGf * (ix + (size_t)LG * iy)] : cplx(0); };
    for (int Rp2 = -h; Rp2 <= h; Rp2++) for (int Rp1 = -h; Rp1 <= h; Rp1++)
    for (int R2 = -h; R2 <= h; R2++) for (int R1 = -h; R1 <= h; R1++) {
        double wgt = 1.0;
        if (LG % 2 == 0) {
            if (std::abs(Rp1) == LG / 2) wgt /= 2; if (std::abs(Rp2) == LG / 2) wgt /= 2;
            if (std::abs(R1) == LG / 2) wgt /= 2;  if (std::abs(R2) == LG / 2) wgt /= 2;
        }
        int gx = mod_(R1, LG), gy = mod_(R2, LG), gpx = mod_(Rp1, LG), gpy = mod_(Rp2, LG);
        int iR = mod_(R1, L) + L * mod_(R2, L);
        int iRm = mod_(Rp1 - R1, L) + L * mod_(Rp2 - R2, L);
        int iRp = mod_(Rp1 + R1, L) + L * mod_(Rp2 + R2, L);
        for (int iv = 0; iv < nFP; iv++) for (int iW = 0; iW < nBP; iW++) {
            int W = iW - (g->nPiB - 1), v = iv - g->nPiF;
            cplx g2 = Gcall(v, gpx, gpy);
            size_t base = iW + (size_t)nBP * iv;
            Pipp[base + (size_t)nBP * nFP * (iR + (size_t)NP * iRm)] += Gcall(B_minus_F(W, v), gx, gy) * g2 * wgt;
            Piph[base + (size_t)nBP * nFP * (iR + (size_t)NP * iRp)] += Gcall(B_plus_F(W, v), gx, gy) * g2 * wgt;
        }
    }
    size_t pre = (size_t)nBP * nFP;
    for (cplx* A : {Pipp, Piph}) {
        dft_axis(A, pre, L, (size_t)L * L * L, +1); dft_axis(A, pre * L, L, (size_t)L * L, +1);
        dft_axis(A, pre * L * L, L, L, +1);         dft_axis(A, pre * L * L * L, L, 1, +1);
    }
}

// ---- bubbles_momentum_space!, src/nonlocal_2/bubble.jl:1-37 (cross-check) --------------
void orc_bubbles_momentum_space(cplx* Pipp, cplx* Piph, const cplx* G, int nG, int LG, const orc_grid* g) {
    int L = g->L, NP = L * L;
    int nBP = 2 * g->nPiB - 1, nFP = 2 * g->nPiF, nGf = 2 * nG;
    size_t len = (size_t)nBP * nFP * NP * NP;
    std::fill(Pipp, Pipp + len, cplx(0)); std::fill(Piph, Piph + len, cplx(0));
    int ratio = LG / L;   // euclidean(k) on the coarse mesh mapped onto the G mesh (requires LG % L == 0)
    for (int iP = 0; iP < NP; iP++) for (int ik = 0; ik < NP; ik++) {
        int Px = (iP % L) * ratio, Py = (iP / L) * ratio, kx = (ik % L) * ratio, ky = (ik / L) * ratio;
        int kG = mod_(kx, LG) + LG * mod_(ky, LG);
        int PmkG = mod_(Px - kx, LG) + LG * mod_(Py - ky, LG);
        int PpkG = mod_(Px + kx, LG) + LG * mod_(Py + ky, LG);
        for (int iW = 0; iW < nBP; iW++) for (int iv = 0; iv < nFP; iv++) {
            int W = iW - (g->nPiB - 1), v = iv - g->nPiF;
            size_t idx = iW + (size_t)nBP * (iv + (size_t)nFP * (iP + (size_t)NP * ik));
            if (inF(v, nG)) {
                int a = B_minus_F(W, v), b = B_plus_F(W, v);
                if (inF(a, nG)) Pipp[idx] = G[posF(v, nG) + (size_t)nGf * kG] * G[posF(a, nG) + (size_t)nGf * PmkG];
                if (inF(b, nG)) Piph[idx] = G[posF(v, nG) + (size_t)nGf * kG] * G[posF(b, nG) + (size_t)nGf * PpkG];
            }
        }
    }
}

// ---- Dyson!, compute_occupation, hubbard_bare_Green (src/dyson.jl, src/models/hubbard.jl) --
void orc_dyson(cplx* G, const cplx* Sigma, const cplx* Gbare, int64_t n) {
    for (int64_t i = 0; i < n; i++) G[i] = 1.0 / (1.0 / Gbare[i] + Sigma[i]);
}
double orc_occupation(const cplx* G, int nG, int LG, double T) {
    cplx s = 0; size_t n = (size_t)2 * nG * LG * LG;
    for (size_t i = 0; i < n; i++) s += G[i];
    return 0.5 + s.imag() * T / (double)(LG * LG);
}
void orc_hubbard_bare_green(cplx* G, int nG, int LG, double T, double mu, double t1, double t2, double t3) {
    for (int iy = 0; iy < LG; iy++) for (int ix = 0; ix < LG; ix++) {
        double k1 = 2 * M_PI * ix / LG, k2 = 2 * M_PI * iy / LG;
        double ek = -2 * t1 * (std::cos(k1) + std::cos(k2)); ek += -4 * t2 * std::cos(k1) * std::cos(k2);
        ek += -2 * t3 * (std::cos(2 * k1) + std::cos(2 * k2));
        for (int n = -nG; n < nG; n++) {
            double nu = (2 * n + 1) * M_PI * T;
            G[posF(n, nG) + (size_t)2 * nG * (ix + (size_t)LG * iy)] = 1.0 / (cplx(0, nu) + mu - ek) * cplx(0, 1);
        }
    }
}

// SG(f): symmetrise in place from the representatives
void orc_symmetrize(cplx* f, const orc_sg* SG) {
    for (int64_t c = 0; c < SG->nclasses; c++) {
        int64_t b = SG->offsets[c], e = SG->offsets[c + 1];
        cplx val = f[SG->index[b]];
        for (int64_t j = b + 1; j < e; j++) f[SG->index[j]] = apply_op(SG->op[j], val);
    }
}

// ---- real-space part of SDE_compute!, src/nonlocal_2/SDE.jl:182-256 --------------------
// Lpp, Lph are clobbered (fft! in place). Sigma (nSig fermion N, LSig x LSig) is overwritten.
void orc_sde_real_space(cplx* Sigma, int nSig, int LSig, const cplx* G, int nG, int LG,
                        cplx* Lpp, cplx* Lph, int nK2b, int nK2f, const orc_sg* SGS, const orc_grid* g) {
    int L = g->L, NP = L * L;
    int nB = 2 * nK2b - 1, nF = 2 * nK2f, nGf = 2 * nG, nSf = 2 * nSig;
    double T = g->T;
    std::vector<cplx> GR(G, G + (size_t)nGf * LG * LG);
    dft_axis(GR.data(), nGf, LG, LG, -1); dft_axis(GR.data(), (size_t)nGf * LG, LG, 1, -1);
    for (auto& x : GR) x /= (double)(LG * LG);
    size_t pre = (size_t)nB * nF, len = pre * NP * NP;
    for (cplx* A : {Lpp, Lph}) {
        dft_axis(A, pre, L, (size_t)L * L * L, -1); dft_axis(A, pre * L, L, (size_t)L * L, -1);
        dft_axis(A, pre * L * L, L, L, -1);         dft_axis(A, pre * L * L * L, L, 1, -1);
    }
    double nrm = (double)L * L * L * L;
    std::vector<cplx> LppR(len), LphR(len);
    for (size_t i = 0; i < len; i++) { LppR[i] = Lpp[i] / nrm; LphR[i] = Lph[i] / nrm; }
    std::vector<cplx> SR((size_t)nSf * LSig * LSig, cplx(0));
    int h = L / 2;
    auto Gcall = [&](int n, int ix, int iy) -> cplx { return inF(n, nG) ? GR[posF(n, nG) + (size_t)nGf * (ix + (size_t)LG * iy)] : cplx(0); };
    for (int Rp2 = -h; Rp2 <= h; Rp2++) for (int Rp1 = -h; Rp1 <= h; Rp1++) {
        int iRpL = mod_(Rp1, L) + L * mod_(Rp2, L);
        for (int R2 = -h; R2 <= h; R2++) for (int R1 = -h; R1 <= h; R1++) {
            double wgt = 1.0;
            if (L % 2 == 0) {
                if (std::abs(Rp1) == L / 2) wgt /= 2; if (std::abs(Rp2) == L / 2) wgt /= 2;
                if (std::abs(R1) == L / 2) wgt /= 2;  if (std::abs(R2) == L / 2) wgt /= 2;
            }
            int iRL = mod_(R1, L) + L * mod_(R2, L);
            int gx = mod_(-R1, LG), gy = mod_(-R2, LG);
            size_t sPP = (size_t)nSf * (mod_(R1 + Rp1, LSig) + (size_t)LSig * mod_(R2 + Rp2, LSig));
            size_t sPH = (size_t)nSf * (mod_(-R1 + Rp1, LSig) + (size_t)LSig * mod_(-R2 + Rp2, LSig));
            size_t lbase = pre * (iRL + (size_t)NP * iRpL);
            for (int iv = 0; iv < nF; iv++) {
                int v = iv - nK2f;
                if (!inF(v, nSig)) continue;
                for (int iW = 0; iW < nB; iW++) {
                    int W = iW - (nK2b - 1);
                    SR[sPP + posF(v, nSig)] += Gcall(B_minus_F(W, v), gx, gy) * LppR[lbase + iW + (size_t)nB * iv] * wgt;
                    SR[sPH + posF(v, nSig)] += Gcall(B_plus_F(W, v), gx, gy) * LphR[lbase + iW + (size_t)nB * iv] * wgt;
                }
            }
        }
    }
    for (auto& x : SR) x *= T;
    dft_axis(SR.data(), nSf, LSig, LSig, +1); dft_axis(SR.data(), (size_t)nSf * LSig, LSig, 1, +1);
    std::copy(SR.begin(), SR.end(), Sigma);
    orc_symmetrize(Sigma, SGS);
}

// ---- SDE_U2_using_G, src/nonlocal/SDE.jl:400-447 --------------------------------------
void orc_sde_U2(cplx* SigU2, const cplx* G, int nG, int LG, double U_re, double U_im, double T, const orc_sg* SGS) {
    int nGf = 2 * nG; size_t n = (size_t)nGf * LG * LG;
    std::vector<cplx> Gp(G, G + n), Gm(G, G + n);
    dft_axis(Gp.data(), nGf, LG, LG, -1); dft_axis(Gp.data(), (size_t)nGf * LG, LG, 1, -1);
    dft_axis(Gm.data(), nGf, LG, LG, +1); dft_axis(Gm.data(), (size_t)nGf * LG, LG, 1, +1);
    for (size_t i = 0; i < n; i++) { Gp[i] /= (double)(LG * LG); Gm[i] /= (double)(LG * LG); }
    std::vector<cplx> SR(n, cplx(0));
#pragma omp parallel for
    for (int iR = 0; iR < LG * LG; iR++) {
        const cplx* gp = Gp.data() + (size_t)nGf * iR; const cplx* gm = Gm.data() + (size_t)nGf * iR;
        for (int i2 = 0; i2 < nGf; i2++) for (int i1 = 0; i1 < nGf; i1++) {
            int w1 = i1 - nG, w2 = i2 - nG;
            cplx gg = gm[i1] * gp[i2];
            for (int iv = 0; iv < nGf; iv++) {
                int v = iv - nG;
                int n3 = B_plus_F(F_minus_F(w1, w2), v);
                if (inF(n3, nG)) SR[(size_t)nGf * iR + iv] += gg * gp[posF(n3, nG)];
            }
        }
    }
    cplx U(U_re, U_im); cplx fac = U * U * (T * T);    // U^2 * T^2
    for (auto& x : SR) x *= fac;
    dft_axis(SR.data(), nGf, LG, LG, +1); dft_axis(SR.data(), (size_t)nGf * LG, LG, 1, +1);
    std::copy(SR.begin(), SR.end(), SigU2);
    orc_symmetrize(SigU2, SGS);
}

// =====================================================================================================
// s-wave NL_ParquetSolver (nl_method = 1 of script/run_Wu_point.jl): vertices with bosonic momentum dependence only,
// K1[W,P], K2[W,v,P], K3[W,v,w,P], bubbles Pi[W,w,P]; every fermionic momentum is the s-wave point kSW.
// The K3 kernels and build_K3_cache! are the NL2 ones above (same formulas, src/nonlocal/BSEa/BSEa_K3.jl,
// src/nonlocal/build_K3_cache.jl:18-94) with g->swave = 1 selecting the three-index bubble.
static inline void decodeK2nl(int64_t idx, int nb, int nf, int L, int& W, int& v, int& iP, Mom& P) {
    int nB = 2 * nb - 1, nF = 2 * nf;
    int iW = idx % nB; idx /= nB; int iv = idx % nF; iP = (int)(idx / nF);
    W = iW - (nb - 1); v = iv - nf; P = mk(iP % L, iP / L);
}
// ---- BSE_K1!, src/nonlocal/BSEa/BSEa_K1.jl:2-58 ----
void orc_nl_bse_K1(cplx* K1, int nK1, const orc_vertex* F0, const orc_vertex* F, const orc_vertex* FL,
                   const cplx* Pi0, const cplx* Pi, const orc_sg* SG, int sign, int Ch, int Sp, int is_mfRG,
                   const orc_grid* g, int64_t c0, int64_t c1) {
    int L = g->L, NP = L * L;
    VertexEval EF0 = {F0, L, NP}, EF = {F, L, NP}, EFL = {FL, L, NP};
    int nBP = 2 * g->nPiB - 1, nFP = 2 * g->nPiF, nB1 = 2 * nK1 - 1;
    double T = g->T;
    auto diagram = [&](int64_t idx) -> cplx {
        int W = (int)(idx % nB1) - (nK1 - 1); int iP = (int)(idx / nB1); Mom P = mk(iP % L, iP / L);
        int iW = posB(W, g->nPiB);
        cplx val = 0;
        for (int iw = 0; iw < nFP; iw++) {
            int w = iw - g->nPiF;
            size_t pidx = iW + (size_t)nBP * (iw + (size_t)nFP * iP);
            if (is_mfRG) {
                cplx Fl  = EF0.eval(0, W, INF, w, P, SW(), SW(), Ch, Sp, ALLF());
                cplx FLr = EFL.eval(0, W, crossingF(W, w, Ch), INF, P, SW(), SW(), Ch, Sp, ALLF());
                val += Fl * Pi0[pidx] * FLr;
            } else {
                cplx Fl  = EF.eval(0, W, INF, w, P, SW(), SW(), Ch, Sp, ALLF());
                cplx F0r = EF0.eval(0, W, crossingF(W, w, Ch), INF, P, SW(), SW(), Ch, Sp, ALLF());
                cplx FLr = EFL.eval(0, W, crossingF(W, w, Ch), INF, P, SW(), SW(), Ch, Sp, ALLF());
                val += Fl * ((Pi[pidx] - Pi0[pidx]) * F0r + Pi[pidx] * FLr);
            }
        }
        return T * val * (double)sign;
    };
    sg_apply(K1, SG, diagram, c0, c1);
}
// ---- BSE_L_K2!, src/nonlocal/BSEa/BSEa_K2.jl:1-41 (omega over the K2 nu-mesh) ----
void orc_nl_bse_L_K2(cplx* K2, int nK2b, int nK2f, const orc_vertex* F0, const orc_vertex* F,
                     const cplx* Pi0, const orc_sg* SG, int sign, int Ch, int Sp,
                     const orc_grid* g, int64_t c0, int64_t c1) {
    int L = g->L, NP = L * L;
    VertexEval EF0 = {F0, L, NP}, EF = {F, L, NP};
    int nBP = 2 * g->nPiB - 1, nFP = 2 * g->nPiF;
    double T = g->T;
    Flags fl = {false, Ch != pCh, Ch != tCh, Ch != aCh};
    auto diagram = [&](int64_t idx) -> cplx {
        int W, v, iP; Mom P; decodeK2nl(idx, nK2b, nK2f, L, W, v, iP, P);
        int iW = posB(W, g->nPiB);
        cplx val = 0;
        for (int iw = 0; iw < 2 * nK2f; iw++) {
            int w = iw - nK2f;
            cplx Gl  = EF.eval(0, W, v, crossingF(W, w, Ch), P, SW(), SW(), Ch, Sp, fl);
            cplx F0r = EF0.eval(0, W, w, INF, P, SW(), SW(), Ch, Sp, ALLF());
            val += Gl * Pi0[iW + (size_t)nBP * (posF(w, g->nPiF) + (size_t)nFP * iP)] * F0r;
        }
        return T * val * (double)sign;
    };
    sg_apply(K2, SG, diagram, c0, c1);
}
// ---- BSE_K2!, src/nonlocal/BSEa/BSEa_K2.jl:44-106 (SG part; the FL add is done by the caller) ----
void orc_nl_bse_K2(cplx* K2, int nK2b, int nK2f, const orc_vertex* F0, const orc_vertex* F, const orc_vertex* FL,
                   const cplx* Pi0, const cplx* Pi, const orc_sg* SG, int sign, int Ch, int Sp, int is_mfRG,
                   const orc_grid* g, int64_t c0, int64_t c1) {
    int L = g->L, NP = L * L;
    VertexEval EF0 = {F0, L, NP}, EF = {F, L, NP}, EFL = {FL, L, NP};
    int nBP = 2 * g->nPiB - 1, nFP = 2 * g->nPiF;
    double T = g->T;
    auto diagram = [&](int64_t idx) -> cplx {
        int W, v, iP; Mom P; decodeK2nl(idx, nK2b, nK2f, L, W, v, iP, P);
        int iW = posB(W, g->nPiB);
        cplx val = 0;
        for (int iw = 0; iw < nFP; iw++) {
            int w = iw - g->nPiF;
            size_t pidx = iW + (size_t)nBP * (iw + (size_t)nFP * iP);
            if (is_mfRG) {
                int wc = crossingF(W, w, Ch);
                cplx Fl  = EF0.eval(0, W, v, wc, P, SW(), SW(), Ch, Sp, ALLF()) - EF0.eval(0, W, INF, wc, P, SW(), SW(), Ch, Sp, ALLF());
                cplx FLr = EFL.eval(0, W, w, INF, P, SW(), SW(), Ch, Sp, ALLF());
                val += Fl * Pi0[pidx] * FLr;
            } else {
                cplx Fl  = EF.eval(0, W, v, w, P, SW(), SW(), Ch, Sp, ALLF()) - EF.eval(0, W, INF, w, P, SW(), SW(), Ch, Sp, ALLF());
                cplx F0r = EF0.eval(0, W, crossingF(W, w, Ch), INF, P, SW(), SW(), Ch, Sp, ALLF());
                cplx FLr = EFL.eval(0, W, crossingF(W, w, Ch), INF, P, SW(), SW(), Ch, Sp, ALLF());
                val += Fl * ((Pi[pidx] - Pi0[pidx]) * F0r + Pi[pidx] * FLr);
            }
        }
        return T * val * (double)sign;
    };
    sg_apply(K2, SG, diagram, c0, c1);
}
// ---- bubbles_real_space! for NL_MF_Pi, src/nonlocal/bubble.jl:86-158:  Pipp(R) = G(R) G(R), Piph(R) = G(R) G(-R), R in
// [-L/2, L/2]^2 weighted 1/2 per component with |R_c| = LG/2 (LG even), R = 0 with the 1/nu tail outside the G mesh
// (use_G_tail = true by default), backward transform over the momentum axes ----
void orc_nl_bubbles_real_space(cplx* Pipp, cplx* Piph, const cplx* G, int nG, int LG, const orc_grid* g, int use_G_tail) {
    const int L = g->L, NP = L * L, nBP = 2 * g->nPiB - 1, nFP = 2 * g->nPiF, nGf = 2 * nG, h = L / 2;
    const double pi = 3.14159265358979323846;
    std::vector<cplx> GR((size_t)nGf * LG * LG);
    std::copy(G, G + GR.size(), GR.begin());
    dft_axis(GR.data(), nGf, LG, LG, -1); dft_axis(GR.data(), (size_t)nGf * LG, LG, 1, -1);
    for (auto& x : GR) x /= (double)(LG * LG);
    const size_t pre = (size_t)nBP * nFP;
    std::fill(Pipp, Pipp + pre * NP, cplx(0)); std::fill(Piph, Piph + pre * NP, cplx(0));
    auto gat = [&](int n, int rx, int ry, bool tail) -> cplx {
        if (inF(n, nG)) return GR[posF(n, nG) + (size_t)nGf * (mod_(rx, LG) + (size_t)LG * mod_(ry, LG))];
        return tail ? cplx(1.0 / ((2 * n + 1) * pi * g->T), 0.0) : cplx(0);
    };
    for (int R2 = -h; R2 <= h; R2++) for (int R1 = -h; R1 <= h; R1++) {
        double weight = 1.0;
        if (LG % 2 == 0) { if (std::abs(R1) == LG / 2) weight /= 2; if (std::abs(R2) == LG / 2) weight /= 2; }
        const size_t iR = mod_(R1, L) + (size_t)L * mod_(R2, L);
        const bool tail = use_G_tail && R1 == 0 && R2 == 0;
        for (int iw = 0; iw < nFP; iw++) for (int iW = 0; iW < nBP; iW++) {
            const int W = iW - (g->nPiB - 1), w = iw - g->nPiF;
            Pipp[iW + (size_t)nBP * iw + pre * iR] += gat(W - w - 1, R1, R2, tail) * gat(w, R1, R2, tail) * weight;
            Piph[iW + (size_t)nBP * iw + pre * iR] += gat(W + w, R1, R2, tail) * gat(w, -R1, -R2, tail) * weight;
        }
    }
    dft_axis(Pipp, pre, L, L, +1); dft_axis(Pipp, pre * L, L, 1, +1);
    dft_axis(Piph, pre, L, L, +1); dft_axis(Piph, pre * L, L, 1, +1);
}
// ---- SDE_channel_L_pp! / ph!, src/nonlocal/SDE.jl:3-146: own gamma of the level only (NL_Vertex / Vertex), core - bare for the
// RefVertex level; `level` indexes the chain V ----
void orc_nl_sde_L(cplx* Lout, int nK2b, int nK2f, const orc_vertex* V, int level, const cplx* Pi, const orc_sg* SG, int is_pp,
                  const orc_grid* g, int64_t c0, int64_t c1) {
    int L = g->L, NP = L * L;
    VertexEval E = {V, L, NP};
    int nBP = 2 * g->nPiB - 1, nFP = 2 * g->nPiF;
    double T = g->T;
    const orc_level& lv = V->lev[level];
    const orc_level& core = V->lev[V->nlev - 1];
    cplx U(core.U_re, core.U_im);
    auto own = [&](int r, int W, int v, int w, Mom P) -> cplx {
        if (lv.type == LV_NL) return E.nl(lv, r).eval(W, v, w, P);
        if (lv.type == LV_LOCAL) return E.loc(lv, r).eval(W, v, w, true, true, true);
        return 0;
    };
    auto diagram = [&](int64_t idx) -> cplx {
        int W, v, iP; Mom P; decodeK2nl(idx, nK2b, nK2f, L, W, v, iP, P);
        int iW = posB(W, g->nPiB);
        cplx val = 0;
        for (int iw = 0; iw < nFP; iw++) {
            int w = iw - g->nPiF;
            cplx pi = Pi[iW + (size_t)nBP * (iw + (size_t)nFP * iP)];
            if (lv.type == LV_CORE) {
                if (is_pp) val += U * pi * (E.core_eval(lv, W, B_minus_F(W, w), v, pCh, pSp) - U);
                else val += U * pi * (E.core_eval(lv, W, v, w, tCh, pSp) + E.core_eval(lv, W, v, w, aCh, pSp) - 2.0 * U);
            } else {
                if (is_pp) val += U * pi * own(pCh, W, B_minus_F(W, w), v, P);
                else val += U * pi * (own(tCh, W, v, w, P) + own(aCh, W, v, w, P));
            }
        }
        return T * val;
    };
    sg_apply(Lout, SG, diagram, c0, c1);
}
// ---- SDE_compute_inner! (use_real_space = true), src/nonlocal/SDE.jl:191-275.  Lpp / Lph are read only here. ----
void orc_nl_sde_inner(cplx* Sigma, int nSig, int LSig, const cplx* G, int nG, int LG, const cplx* Lpp, const cplx* Lph,
                      int nK2b, int nK2f, const orc_sg* SGS, const orc_grid* g) {
    const int L = g->L, nB = 2 * nK2b - 1, nF = 2 * nK2f, nGf = 2 * nG, nSf = 2 * nSig, h = L / 2;
    const size_t pre = (size_t)nB * nF;
    std::vector<cplx> GR((size_t)nGf * LG * LG), A(Lpp, Lpp + pre * L * L), B(Lph, Lph + pre * L * L), SR((size_t)nSf * LSig * LSig, cplx(0));
    std::copy(G, G + GR.size(), GR.begin());
    dft_axis(GR.data(), nGf, LG, LG, -1); dft_axis(GR.data(), (size_t)nGf * LG, LG, 1, -1);
    for (auto& x : GR) x /= (double)(LG * LG);
    dft_axis(A.data(), pre, L, L, -1); dft_axis(A.data(), pre * L, L, 1, -1);
    dft_axis(B.data(), pre, L, L, -1); dft_axis(B.data(), pre * L, L, 1, -1);
    for (auto& x : A) x /= (double)(L * L);
    for (auto& x : B) x /= (double)(L * L);
    for (int R2 = -h; R2 <= h; R2++) for (int R1 = -h; R1 <= h; R1++) {
        double weight = 1.0;
        if (L % 2 == 0) { if (std::abs(R1) == L / 2) weight /= 2; if (std::abs(R2) == L / 2) weight /= 2; }
        const size_t iRL = mod_(R1, L) + (size_t)L * mod_(R2, L);
        const size_t imRG = mod_(-R1, LG) + (size_t)LG * mod_(-R2, LG);
        const size_t ipRS = mod_(R1, LSig) + (size_t)LSig * mod_(R2, LSig), imRS = mod_(-R1, LSig) + (size_t)LSig * mod_(-R2, LSig);
        for (int iv = 0; iv < nF; iv++) {
            const int v = iv - nK2f;
            if (!inF(v, nSig)) continue;
            for (int iW = 0; iW < nB; iW++) {
                const int W = iW - (nK2b - 1);
                if (inF(W - v - 1, nG)) SR[posF(v, nSig) + (size_t)nSf * ipRS] += GR[posF(W - v - 1, nG) + (size_t)nGf * imRG] * A[iW + (size_t)nB * iv + pre * iRL] * weight;
                if (inF(W + v, nG))     SR[posF(v, nSig) + (size_t)nSf * imRS] += GR[posF(W + v, nG) + (size_t)nGf * imRG] * B[iW + (size_t)nB * iv + pre * iRL] * weight;
            }
        }
    }
    for (auto& x : SR) x *= g->T;
    dft_axis(SR.data(), nSf, LSig, LSig, +1); dft_axis(SR.data(), (size_t)nSf * LSig, LSig, 1, +1);
    std::copy(SR.begin(), SR.end(), Sigma);
    orc_symmetrize(Sigma, SGS);
}

// ---- MatsubaraFunctions.SymmetryGroup(symmetries, f) restated ------
// kind: 0 Sigma(nu,k) 1 K1(W,P) 2 K2pp 3 K2ph 4 K3pp 5 K3ph 6 K3ppL 7 K3phL 8 K2pp[W,v,P] 9 K2ph[W,v,P] (s-wave NL solver)
// generators in the order of src/nonlocal_2/ParquetSolver.jl:200-291.
struct SymPt { int f[3]; Mom m[2]; };
static bool apply_gen(int kind, int gi, const SymPt& a, int L, SymPt& b, uint8_t& op) {
    auto neg = [&](Mom k) { return mk(mod_(-k.x, L), mod_(-k.y, L)); };
    auto ref = [&](Mom k) { return mk(k.y, k.x); };
    auto rot = [&](Mom k) { return mk(mod_(k.y, L), mod_(-k.x, L)); };
    auto fold = [&](Mom k) { return mk(mod_(k.x, L), mod_(k.y, L)); };
    b = a; op = 0;
    int ngen = 0;
    switch (kind) {
    case 0: case 1: ngen = 3;
        if (gi == 0) { b.f[0] = (kind == 0) ? -a.f[0] - 1 : -a.f[0]; b.m[0] = neg(a.m[0]); op = (kind == 0) ? 3 : 2; }
        else if (gi == 1) b.m[0] = ref(a.m[0]);
        else b.m[0] = rot(a.m[0]);
        break;
    case 2: case 3: ngen = 4;   // K2 NL2: src/nonlocal_2/symmetries.jl:4-46
        if (gi == 0) { b.f[0] = -a.f[0]; b.f[1] = -a.f[1] - 1; b.m[0] = neg(a.m[0]); b.m[1] = neg(a.m[1]); op = 2; }
        else if (gi == 1) {
            if (kind == 2) { b.f[1] = B_minus_F(a.f[0], a.f[1]); b.m[1] = fold(a.m[0] - a.m[1]); }
            else { b.f[0] = -a.f[0]; b.f[1] = B_plus_F(a.f[0], a.f[1]); b.m[0] = neg(a.m[0]); b.m[1] = fold(a.m[0] + a.m[1]); }
        }
        else if (gi == 2) { b.m[0] = ref(a.m[0]); b.m[1] = ref(a.m[1]); }
        else { b.m[0] = rot(a.m[0]); b.m[1] = rot(a.m[1]); }
        break;
    case 4: case 5: ngen = 5;   // K3: src/nonlocal/symmetries.jl:70-139
        if (gi == 0) { b.f[0] = -a.f[0]; b.f[1] = -a.f[1] - 1; b.f[2] = -a.f[2] - 1; b.m[0] = neg(a.m[0]); op = 2; }
        else if (gi == 1) { b.f[1] = a.f[2]; b.f[2] = a.f[1]; }
        else if (gi == 2) {
            if (kind == 4) { b.f[1] = B_minus_F(a.f[0], a.f[1]); b.f[2] = B_minus_F(a.f[0], a.f[2]); }
            else { b.f[0] = -a.f[0]; b.f[1] = B_plus_F(a.f[0], a.f[1]); b.f[2] = B_plus_F(a.f[0], a.f[2]); b.m[0] = neg(a.m[0]); }
        }
        else if (gi == 3) b.m[0] = ref(a.m[0]);
        else b.m[0] = rot(a.m[0]);
        break;
    case 8: case 9: ngen = 4;   // K2 with s-wave truncation: src/nonlocal/symmetries.jl:47-72, order of src/nonlocal/ParquetSolver.jl:215-250
        if (gi == 0) { b.f[0] = -a.f[0]; b.f[1] = -a.f[1] - 1; b.m[0] = neg(a.m[0]); op = 2; }
        else if (gi == 1) {
            if (kind == 8) b.f[1] = B_minus_F(a.f[0], a.f[1]);
            else { b.f[0] = -a.f[0]; b.f[1] = B_plus_F(a.f[0], a.f[1]); b.m[0] = neg(a.m[0]); }
        }
        else if (gi == 2) b.m[0] = ref(a.m[0]);
        else b.m[0] = rot(a.m[0]);
        break;
    case 6: case 7: ngen = 4;   // K3 left: generators 1, 3, ref, rot
        if (gi == 0) { b.f[0] = -a.f[0]; b.f[1] = -a.f[1] - 1; b.f[2] = -a.f[2] - 1; b.m[0] = neg(a.m[0]); op = 2; }
        else if (gi == 1) {
            if (kind == 6) { b.f[1] = B_minus_F(a.f[0], a.f[1]); b.f[2] = B_minus_F(a.f[0], a.f[2]); }
            else { b.f[0] = -a.f[0]; b.f[1] = B_plus_F(a.f[0], a.f[1]); b.f[2] = B_plus_F(a.f[0], a.f[2]); b.m[0] = neg(a.m[0]); }
        }
        else if (gi == 2) b.m[0] = ref(a.m[0]);
        else b.m[0] = rot(a.m[0]);
        break;
    }
    return gi < ngen;
}

// n0: N of the first (bosonic; fermionic for kind 0) mesh, n1: N of the fermionic meshes.
// Outputs sized to the array length; returns the number of classes.
int64_t orc_build_symmetry_group(int kind, int n0, int n1, int L, int64_t* offsets, int64_t* index, uint8_t* ops) {
    int NP = L * L;
    int nfreq = (kind <= 1) ? 1 : ((kind <= 3 || kind >= 8) ? 2 : 3);
    int nmom = (kind == 2 || kind == 3) ? 2 : 1;
    int len0 = (kind == 0) ? 2 * n0 : 2 * n0 - 1, len1 = 2 * n1;
    int dims[5]; int nd = 0;
    dims[nd++] = len0; for (int i = 1; i < nfreq; i++) dims[nd++] = len1; for (int i = 0; i < nmom; i++) dims[nd++] = NP;
    int64_t total = 1; for (int i = 0; i < nd; i++) total *= dims[i];
    auto decode = [&](int64_t idx, SymPt& p) {
        int i0 = idx % len0; idx /= len0; p.f[0] = (kind == 0) ? i0 - n0 : i0 - (n0 - 1);
        for (int i = 1; i < nfreq; i++) { p.f[i] = (int)(idx % len1) - n1; idx /= len1; }
        for (int i = 0; i < nmom; i++) { int ik = idx % NP; idx /= NP; p.m[i] = mk(ik % L, ik / L); }
    };
    auto inb = [&](const SymPt& p) {
        if (kind == 0) { if (!inF(p.f[0], n0)) return false; } else if (!inB(p.f[0], n0)) return false;
        for (int i = 1; i < nfreq; i++) if (!inF(p.f[i], n1)) return false;
        return true;
    };
    auto encode = [&](const SymPt& p) {
        int64_t idx = 0, stride = 1;
        idx += (kind == 0 ? posF(p.f[0], n0) : posB(p.f[0], n0)); stride *= len0;
        for (int i = 1; i < nfreq; i++) { idx += stride * posF(p.f[i], n1); stride *= len1; }
        for (int i = 0; i < nmom; i++) { idx += stride * kidx(p.m[i], L); stride *= NP; }
        return idx;
    };
    std::vector<uint8_t> checked(total, 0);
    int64_t ncls = 0, nmem = 0;
    struct Frame { SymPt p; uint8_t op; int gi; };
    std::vector<Frame> stack;
    for (int64_t idx = 0; idx < total; idx++) {
        if (checked[idx]) continue;
        checked[idx] = 1;
        offsets[ncls++] = nmem;
        index[nmem] = idx; ops[nmem] = 0; nmem++;
        Frame f0; decode(idx, f0.p); f0.op = 0; f0.gi = 0;
        stack.clear(); stack.push_back(f0);
        while (!stack.empty()) {                 // depth-first, generators in order
            Frame& fr = stack.back();
            SymPt b; uint8_t gop;
            if (!apply_gen(kind, fr.gi, fr.p, L, b, gop)) { stack.pop_back(); continue; }
            fr.gi++;
            uint8_t nop = gop ^ fr.op;
            if (!inb(b)) continue;
            int64_t j = encode(b);
            if (checked[j]) continue;
            checked[j] = 1;
            index[nmem] = j; ops[nmem] = nop; nmem++;
            Frame nf; nf.p = b; nf.op = nop; nf.gi = 0;
            stack.push_back(nf);
        }
    }
    offsets[ncls] = nmem;
    return ncls;
}


// =====================================================================================================
// Local (ParquetSolver, impurity) variants.  A local Vertex is carried as an NL2 vertex on a 1 x 1 momentum mesh
// (K1[W,1], K2[W,v,1,1], K3[W,v,w,1]); with NP = 1 the NL2 restatements above coincide term by term with
// src/BSEa/BSEa_K1.jl:2-54, BSEa_K2.jl:43-103, BSEa_K3.jl:1-128, src/build_K3_cache.jl:20-95 and src/SDE.jl:60-290
// (diffed against the NL2 files).  The functions that DIFFER are restated here.

// ---- BSE_L_K2! local, src/BSEa/BSEa_K2.jl:1-40: omega over the BUBBLE mesh, crossing on the right vertex ----
void orc_bse_L_K2_local(cplx* K2, int nK2b, int nK2f, const orc_vertex* F0, const orc_vertex* F,
                        const cplx* Pi0, const orc_sg* SG, int sign, int Ch, int Sp, const orc_grid* g, int64_t c0, int64_t c1) {
    VertexEval EF0 = {F0, 1, 1}, EF = {F, 1, 1};
    int nBP = 2 * g->nPiB - 1, nFP = 2 * g->nPiF;
    K2Shape s = {nK2b, nK2f, 1};
    double T = g->T;
    Mom z = mk(0, 0);
    Flags fl = {false, Ch != pCh, Ch != tCh, Ch != aCh};
    auto diagram = [&](int64_t idx) -> cplx {
        int W, v; Mom P, k; decodeK2(idx, s, 1, W, v, P, k);
        int iW = posB(W, g->nPiB);
        cplx val = 0;
        for (int iw = 0; iw < nFP; iw++) {
            int w = iw - g->nPiF;
            cplx Gp  = EF.eval(0, W, v, w, z, z, z, Ch, Sp, fl);
            cplx F0p = EF0.eval(0, W, crossingF(W, w, Ch), INF, z, z, z, Ch, Sp, ALLF());
            val += Gp * Pi0[iW + (size_t)nBP * iw] * F0p;
        }
        return T * val * (double)sign;
    };
    sg_apply(K2, SG, diagram, c0, c1);
}

// ---- bubbles! local, src/bubble.jl:9-36 (use_G_tail = true: 1/nu outside the G mesh) ------------------------
void orc_bubbles_local(cplx* Pipp, cplx* Piph, const cplx* G, int nG, const orc_grid* g) {
    int nBP = 2 * g->nPiB - 1, nFP = 2 * g->nPiF;
    double T = g->T;
    auto Gt = [&](int n) -> cplx { return inF(n, nG) ? G[posF(n, nG)] : cplx(1.0 / ((2 * n + 1) * M_PI * T), 0.0); };
    for (int iW = 0; iW < nBP; iW++) for (int iv = 0; iv < nFP; iv++) {
        int W = iW - (g->nPiB - 1), v = iv - g->nPiF;
        cplx Gv = Gt(v), Gm = Gt(B_minus_F(W, v)), Gp = Gt(B_plus_F(W, v));
        Pipp[iW + (size_t)nBP * iv] = Gv * Gm;
        Piph[iW + (size_t)nBP * iv] = Gv * Gp;
    }
}

// ---- siam_bare_Green, src/models/siam.jl:9-28 (stores i*G) --------------------------------------------------
void orc_siam_bare_green(cplx* G, int nG, double T, double e, double Delta, double D) {
    for (int n = -nG; n < nG; n++) {
        double nu = (2 * n + 1) * M_PI * T;
        if (std::isinf(D)) G[posF(n, nG)] = 1.0 / (cplx(nu, e) + Delta * (nu > 0 ? 1.0 : -1.0));
        else G[posF(n, nG)] = 1.0 / (cplx(nu, e) + 2 * Delta / M_PI * std::atan(D / nu));
    }
}

}  // extern "C"
