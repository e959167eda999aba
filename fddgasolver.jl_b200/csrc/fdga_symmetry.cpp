// fdga_symmetry.cpp -- host-side (integer only) construction of symmetry-class tables.
//
// Restates MatsubaraFunctions.jl's SymmetryGroup(symmetries, f) (not vendored in the reference;
// behaviour per SURVEY.md Appendix B: ascending scan of linear indices, depth-first closure under the
// generator list, out-of-box images dropped, accumulated Operation = xor of (sgn, con) flags) for the
// generator lists of init_sym_grp!(::NL2_ParquetSolver), src/nonlocal_2/ParquetSolver.jl:200-291:
//   SGSigma : sS_conj, sS_ref, sS_rot                      src/nonlocal/symmetries.jl:17-27
//   SGpp[1] = SGph[1] : sK1_conj, sK1_ref, sK1_rot         src/nonlocal/symmetries.jl:32-42
//   SGpp[2] : sK2_NL2_pp1, pp2, ref, rot                   src/nonlocal_2/symmetries.jl:4-32
//   SGph[2] : sK2_NL2_ph1, ph2, ref, rot                   src/nonlocal_2/symmetries.jl:35-46
//   SGpp[3] : sK3pp1, pp2, pp3, ref, rot   SGph[3] : sK3ph1, ph2, ph3, ref, rot    src/nonlocal/symmetries.jl:70-139
//   SGppL[3]: sK3pp1, pp3, ref, rot        SGphL[3]: sK3ph1, ph3, ref, rot
// and of init_sym_grp!(::NL_ParquetSolver), src/nonlocal/ParquetSolver.jl:196-290, whose K2[W,v,P] groups differ:
//   SGpp[2] : sK2pp1, pp2, ref, rot        SGph[2] : sK2ph1, ph2, ref, rot        src/nonlocal/symmetries.jl:47-72
// In drop-in use Julia passes SG.classes through fdga_set_symmetry_classes and this builder is not needed.
#include <cstdint>
#include <vector>
#include <functional>
#include "../../include/fdga.h"

namespace {

struct Point {            // up to three Matsubara indices and two folded momenta
    int f0, f1, f2;
    int p0x, p0y, p1x, p1y;
};
struct Image { Point p; uint8_t op; };

inline int fold(int a, int L) { int r = a % L; return r < 0 ? r + L : r; }

struct Builder {
    int which, n0, n1, L, NP;
    int nfreq, nmom;
    int len0, len1;

    Builder(int which_, int n0_, int n1_, int L_) : which(which_), n0(n0_), n1(n1_), L(L_), NP(L_ * L_) {
        nfreq = (which <= FDGA_SG_K1) ? 1 : ((which == FDGA_SG_PP2 || which == FDGA_SG_PH2 || which == FDGA_SG_NL_PP2 || which == FDGA_SG_NL_PH2) ? 2 : 3);
        nmom = (which == FDGA_SG_PP2 || which == FDGA_SG_PH2) ? 2 : 1;
        len0 = (which == FDGA_SG_SIGMA) ? 2 * n0 : 2 * n0 - 1;
        len1 = 2 * n1;
    }
    int64_t total() const {
        int64_t t = len0;
        for (int i = 1; i < nfreq; i++) t *= len1;
        for (int i = 0; i < nmom; i++) t *= NP;
        return t;
    }
    bool first_is_fermion() const { return which == FDGA_SG_SIGMA; }
    Point decode(int64_t idx) const {
        Point p = {0, 0, 0, 0, 0, 0, 0};
        int i0 = (int)(idx % len0); idx /= len0;
        p.f0 = first_is_fermion() ? i0 - n0 : i0 - (n0 - 1);
        if (nfreq > 1) { p.f1 = (int)(idx % len1) - n1; idx /= len1; }
        if (nfreq > 2) { p.f2 = (int)(idx % len1) - n1; idx /= len1; }
        int k = (int)(idx % NP); idx /= NP; p.p0x = k % L; p.p0y = k / L;
        if (nmom > 1) { k = (int)(idx % NP); p.p1x = k % L; p.p1y = k / L; }
        return p;
    }
    bool inbounds(const Point& p) const {
        if (first_is_fermion()) { if (p.f0 < -n0 || p.f0 > n0 - 1) return false; }
        else if (p.f0 < -(n0 - 1) || p.f0 > n0 - 1) return false;
        if (nfreq > 1 && (p.f1 < -n1 || p.f1 > n1 - 1)) return false;
        if (nfreq > 2 && (p.f2 < -n1 || p.f2 > n1 - 1)) return false;
        return true;
    }
    int64_t encode(const Point& p) const {
        int64_t idx = first_is_fermion() ? p.f0 + n0 : p.f0 + n0 - 1;
        int64_t stride = len0;
        if (nfreq > 1) { idx += stride * (p.f1 + n1); stride *= len1; }
        if (nfreq > 2) { idx += stride * (p.f2 + n1); stride *= len1; }
        idx += stride * (fold(p.p0x, L) + (int64_t)L * fold(p.p0y, L)); stride *= NP;
        if (nmom > 1) idx += stride * (fold(p.p1x, L) + (int64_t)L * fold(p.p1y, L));
        return idx;
    }

    // lattice operations, src/nonlocal/symmetries.jl:4-12
    void ref_all(Point& p) const { std::swap(p.p0x, p.p0y); if (nmom > 1) std::swap(p.p1x, p.p1y); }
    void rot_all(Point& p) const {
        int x = p.p0x, y = p.p0y; p.p0x = fold(y, L); p.p0y = fold(-x, L);
        if (nmom > 1) { x = p.p1x; y = p.p1y; p.p1x = fold(y, L); p.p1y = fold(-x, L); }
    }
    void neg_mom(int& x, int& y) const { x = fold(-x, L); y = fold(-y, L); }

    // the generator list; fermionic sign flip -F(n) = F(-n-1), bosonic -B(m) = B(-m),
    // B(m) - F(n) = F(m-n-1), B(m) + F(n) = F(m+n)
    std::vector<Image> images(const Point& a) const {
        std::vector<Image> out;
        auto push = [&](Point p, uint8_t op) { Image im; im.p = p; im.op = op; out.push_back(im); };
        Point b;
        const bool pp = (which == FDGA_SG_PP2 || which == FDGA_SG_PP3 || which == FDGA_SG_PPL3);
        switch (which) {
        case FDGA_SG_SIGMA:
            b = a; b.f0 = -a.f0 - 1; neg_mom(b.p0x, b.p0y); push(b, 3);     // sgn and con
            b = a; ref_all(b); push(b, 0);
            b = a; rot_all(b); push(b, 0);
            break;
        case FDGA_SG_K1:
            b = a; b.f0 = -a.f0; neg_mom(b.p0x, b.p0y); push(b, 2);
            b = a; ref_all(b); push(b, 0);
            b = a; rot_all(b); push(b, 0);
            break;
        case FDGA_SG_PP2: case FDGA_SG_PH2:
            b = a; b.f0 = -a.f0; b.f1 = -a.f1 - 1; neg_mom(b.p0x, b.p0y); neg_mom(b.p1x, b.p1y); push(b, 2);
            b = a;
            if (pp) { b.f1 = a.f0 - a.f1 - 1; b.p1x = fold(a.p0x - a.p1x, L); b.p1y = fold(a.p0y - a.p1y, L); }
            else    { b.f0 = -a.f0; b.f1 = a.f0 + a.f1; neg_mom(b.p0x, b.p0y); b.p1x = fold(a.p0x + a.p1x, L); b.p1y = fold(a.p0y + a.p1y, L); }
            push(b, 0);
            b = a; ref_all(b); push(b, 0);
            b = a; rot_all(b); push(b, 0);
            break;
        case FDGA_SG_NL_PP2: case FDGA_SG_NL_PH2:
            b = a; b.f0 = -a.f0; b.f1 = -a.f1 - 1; neg_mom(b.p0x, b.p0y); push(b, 2);
            b = a;
            if (which == FDGA_SG_NL_PP2) b.f1 = a.f0 - a.f1 - 1;
            else { b.f0 = -a.f0; b.f1 = a.f0 + a.f1; neg_mom(b.p0x, b.p0y); }
            push(b, 0);
            b = a; ref_all(b); push(b, 0);
            b = a; rot_all(b); push(b, 0);
            break;
        default: {   // K3 groups
            b = a; b.f0 = -a.f0; b.f1 = -a.f1 - 1; b.f2 = -a.f2 - 1; neg_mom(b.p0x, b.p0y); push(b, 2);
            const bool has_swap = (which == FDGA_SG_PP3 || which == FDGA_SG_PH3);
            if (has_swap) { b = a; b.f1 = a.f2; b.f2 = a.f1; push(b, 0); }
            b = a;
            if (pp) { b.f1 = a.f0 - a.f1 - 1; b.f2 = a.f0 - a.f2 - 1; }
            else    { b.f0 = -a.f0; b.f1 = a.f0 + a.f1; b.f2 = a.f0 + a.f2; neg_mom(b.p0x, b.p0y); }
            push(b, 0);
            b = a; ref_all(b); push(b, 0);
            b = a; rot_all(b); push(b, 0);
        } }
        return out;
    }
};

}  // namespace

extern "C" int64_t fdga_symgroup_build_host(int which, int n0, int n1, int nq, int64_t* offsets, int64_t* index, uint8_t* ops) {
    Builder B(which, n0, n1, nq);
    const int64_t total = B.total();
    std::vector<uint8_t> checked((size_t)total, 0);
    int64_t ncls = 0, nmem = 0;
    // depth-first closure: recursion mirrors the reference package's `reduce`
    std::function<void(const Point&, uint8_t)> visit = [&](const Point& w, uint8_t op) {
        for (const Image& im : B.images(w)) {
            uint8_t nop = im.op ^ op;
            if (!B.inbounds(im.p)) continue;
            int64_t j = B.encode(im.p);
            if (checked[(size_t)j]) continue;
            checked[(size_t)j] = 1;
            index[nmem] = j; ops[nmem] = nop; nmem++;
            visit(im.p, nop);
        }
    };
    for (int64_t idx = 0; idx < total; idx++) {
        if (checked[(size_t)idx]) continue;
        checked[(size_t)idx] = 1;
        offsets[ncls++] = nmem;
        index[nmem] = idx; ops[nmem] = 0; nmem++;
        visit(B.decode(idx), 0);
    }
    offsets[ncls] = nmem;
    return ncls;
}
