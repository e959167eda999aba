// Host-side unit test of the column-kernel index machinery (no GPU needed): the interval-based sums
// chan_lin_sum / chan_lin_sum_diff_v must equal the term-by-term evaluation with chan_off / chan_off_diff_v
// for every job kind, channel, form and cross channel, on random tables with ragged boxes.
// Built and run by tests/test_host_eval.py:  nvcc -std=c++17 -o host_eval_test tests/host_eval_test.cu
#include "../fddgasolver.jl_b200/csrc/fdga_column.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
using namespace fdga;

static double rnd() { return rand() / (double)RAND_MAX - 0.5; }

template <int KIND, int CH>
static double check(const DevLevel& lv, int L, int Nin, int nw, int nK2b_out, int nK2f_out) {
    typedef Forms<KIND, CH> FM;
    const int NP = L * L;
    std::vector<C> R(nw);
    for (auto& x : R) x = mkC(rnd(), rnd());
    double maxerr = 0.0;
    for (int W = -(nK2b_out - 1); W <= nK2b_out - 1; ++W)
    for (int nu = -nK2f_out; nu < nK2f_out; ++nu)
    for (int trial = 0; trial < 6; ++trial) {
        int Px = rand() % L, Py = rand() % L, kx = rand() % L, ky = rand() % L, qx = rand() % L, qy = rand() % L;
        int akx, aky, aqx, aqy;
        if (KIND == JOB_K2 || KIND == JOB_SDE_PH) { akx = kx; aky = ky; aqx = qx; aqy = qy; }
        else if (KIND == JOB_K2_MF || KIND == JOB_LK2) { akx = kx; aky = ky; aqx = (CH == CH_P) ? Px - qx : qx; aqy = (CH == CH_P) ? Py - qy : qy; }
        else { akx = Px - qx; aky = Py - qy; aqx = kx; aqy = ky; }
        for (int f = 0; f < FM::n; ++f) {
            const int form = FM::ch(f);
            MomOff mo[3];
            for (int r = 0; r < 3; ++r) mo[r] = mom_offsets(lv, form, r, L, NP, Px, Py, akx, aky, aqx, aqy);
            int v_a, w_a, v_b, w_b;
            job_freq_args<KIND, CH>(W, nu, 0, v_a, w_a); job_freq_args<KIND, CH>(W, nu, 1, v_b, w_b);
            for (int w_lo = 0; w_lo < nw; w_lo += (nw + 1) / 2) {
                int w_hi = std::min(nw, w_lo + (nw + 1) / 2);
                // brute force
                C b_cross = zeroC(), b_diff = zeroC(), b_full = zeroC();
                for (int iw = w_lo; iw < w_hi; ++iw) {
                    int v, w; job_freq_args<KIND, CH>(W, nu, iw - Nin, v, w);
                    for (int r = 0; r < 3; ++r) if (r != form) {
                        int W2, v2, w2; convert_freq(W, v, w, form, r, W2, v2, w2);
                        b_cross += chan_off(lv, r, mo[r], W2, v2, w2) * R[iw];
                    }
                    b_diff += chan_off_diff_v(lv, form, mo[form], W, v, w) * R[iw];
                    b_full += chan_off(lv, form, mo[form], W, v, w) * R[iw];
                }
                C rs = zeroC(); for (int iw = w_lo; iw < w_hi; ++iw) rs += R[iw];
                C l_cross = zeroC();
                for (int r = 0; r < 3; ++r) if (r != form) {
                    int W0, v0, w0, W1, v1, w1;
                    convert_freq(W, v_a, w_a, form, r, W0, v0, w0); convert_freq(W, v_b, w_b, form, r, W1, v1, w1);
                    Lin lW = {W0, W1 - W0}, lv2 = {v0, v1 - v0}, lw2 = {w0, w1 - w0};
                    l_cross += chan_lin_sum(lv, r, mo[r], lW, lv2, lw2, R.data(), Nin, 1, w_lo, w_hi, rs, true);
                }
                Lin lW = {W, 0}, lv2 = {v_a, v_b - v_a}, lw2 = {w_a, w_b - w_a};
                C l_diff = chan_lin_sum_diff_v(lv, form, mo[form], lW, lv2, lw2, R.data(), Nin, 1, w_lo, w_hi, rs);
                C l_full = chan_lin_sum(lv, form, mo[form], lW, lv2, lw2, R.data(), Nin, 1, w_lo, w_hi, rs, true);
                auto err = [](C a, C b) { return std::max(std::fabs(a.x - b.x), std::fabs(a.y - b.y)); };
                maxerr = std::max(maxerr, std::max(err(b_cross, l_cross), std::max(err(b_diff, l_diff), err(b_full, l_full))));
            }
        }
    }
    return maxerr;
}

int main() {
    setvbuf(stdout, NULL, _IONBF, 0); printf("start\n"); srand(12345);
    double worst = 0.0;
    // (nK1, nK2b, nK2f, nK3b, nK3f, L, output K2 mesh, inner mesh N) -- ragged boxes, level grid != output grid
    int cfgs[][10] = { {5, 3, 2, 2, 2, 3, 3, 2, 5, 0}, {6, 3, 4, 2, 2, 4, 3, 4, 6, 0}, {9, 4, 5, 2, 1, 2, 3, 2, 4, 0}, {4, 2, 2, 2, 2, 3, 2, 2, 2, 0} };
    for (auto& c : cfgs) {
        DevLevel lv; memset(&lv, 0, sizeof(lv));
        lv.type = LV_NL2; lv.nK1 = c[0]; lv.nK2b = c[1]; lv.nK2f = c[2]; lv.nK3b = c[3]; lv.nK3f = c[4];
        int L = c[5], NP = L * L;
        size_t n1 = (2 * lv.nK1 - 1) * NP, n2 = (size_t)(2 * lv.nK2b - 1) * (2 * lv.nK2f) * NP * NP, n3 = (size_t)(2 * lv.nK3b - 1) * (2 * lv.nK3f) * (2 * lv.nK3f) * NP;
        std::vector<std::vector<C>> store;
        for (int r = 0; r < 3; ++r) {
            for (size_t n : {n1, n2, n3}) { store.emplace_back(n); for (auto& x : store.back()) x = mkC(rnd(), rnd()); }
            lv.ch[r].K1 = store[store.size() - 3].data(); lv.ch[r].K2 = store[store.size() - 2].data(); lv.ch[r].K3 = store[store.size() - 1].data();
        }
        int nb = c[6], nf = c[7], Nin = c[8], nw = 2 * Nin;
        double e = 0; printf("cfg built\n");
        e = std::max(e, check<JOB_K2, CH_P>(lv, L, Nin, nw, nb, nf)); printf("k2p %.3e\n", e);
        e = std::max(e, check<JOB_K2, CH_T>(lv, L, Nin, nw, nb, nf));
        e = std::max(e, check<JOB_K2, CH_A>(lv, L, Nin, nw, nb, nf));
        e = std::max(e, check<JOB_K2_MF, CH_P>(lv, L, Nin, nw, nb, nf));
        e = std::max(e, check<JOB_K2_MF, CH_T>(lv, L, Nin, nw, nb, nf));
        e = std::max(e, check<JOB_LK2, CH_P>(lv, L, nf, 2 * nf, nb, nf));
        e = std::max(e, check<JOB_LK2, CH_A>(lv, L, nf, 2 * nf, nb, nf));
        e = std::max(e, check<JOB_SDE_PP, CH_P>(lv, L, Nin, nw, nb, nf));
        e = std::max(e, check<JOB_SDE_PH, CH_A>(lv, L, Nin, nw, nb, nf));
        printf("config nK1=%d nK2=(%d,%d) nK3=(%d,%d) L=%d : max abs deviation %.3e\n", c[0], c[1], c[2], c[3], c[4], L, e);
        worst = std::max(worst, e);
    }
    printf("WORST %.3e\n", worst);
    return worst < 1e-12 ? 0 : 1;
}
