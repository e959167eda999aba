// Host-side (CPU) test of the complete column-kernel arithmetic against the straightforward per-term evaluator
// (eval_vertex) for every job kind / channel on a random nested chain NL2 -> NL2 -> LOCAL -> CORE with ragged boxes.
// Built and run by tests/test_host_eval.py (nvcc, no GPU needed).
#include <cstdio>
#include "../fddgasolver.jl_b200/csrc/fdga_qlane.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
using namespace fdga;

static double rnd() { return rand() / (double)RAND_MAX - 0.5; }
static std::vector<std::vector<C>> g_store;
#ifdef DEVICE_CHECK
static const C* rnd_array(size_t n, double scale = 1.0) {
    C* p; cudaMallocManaged(&p, n * sizeof(C));
    for (size_t i = 0; i < n; ++i) p[i] = mkC(scale * rnd(), scale * rnd());
    return p;
}
template <int KIND, int CH>
__global__ void dev_run(DevChain V, ColJob job, ColDev cols, const C* R, const C* T, Grid g, C* out) {
    C acc[FDGA_WGROUP];
    column_thread<KIND, CH>(V, job, cols, R, g, 0, threadIdx.x, blockDim.x, acc);
    for (int i = 0; i < FDGA_WGROUP; ++i) out[threadIdx.x + blockDim.x * i] = acc[i];
}
// the q-lane contraction on the device: one warp per representative, lane partial sums written out
template <int KIND, int CH>
__global__ void dev_run_q(DevChain V, ColJob job, const int* rep4, int nrep, const C* R, Grid g, C* out) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (w < nrep) out[threadIdx.x] = qlane_lane<KIND, CH>(V, job, g, R, rep4[4 * w], rep4[4 * w + 1], rep4[4 * w + 2], rep4[4 * w + 3], lane);
}
#else
static const C* rnd_array(size_t n, double scale = 1.0) {
    g_store.emplace_back(n);
    for (auto& x : g_store.back()) x = mkC(scale * rnd(), scale * rnd());
    return g_store.back().data();
}
#endif

static C* raw_array(size_t n) {
#ifdef DEVICE_CHECK
    C* p; cudaMallocManaged(&p, n * sizeof(C)); return p;
#else
    g_store.emplace_back(n); return g_store.back().data();
#endif
}
// momentum-fastest layouts of an NL2 level (what mom_layout_kernel builds on the device), via the same index map
static void build_mom_layouts(DevLevel& lv, int L, int NP) {
    const int nB1 = 2 * lv.nK1 - 1; const size_t n2 = (size_t)(2 * lv.nK2b - 1) * (2 * lv.nK2f), n3 = (size_t)(2 * lv.nK3b - 1) * (2 * lv.nK3f) * (2 * lv.nK3f);
    for (int r = 0; r < 3; ++r) {
        DevChan& c = lv.ch[r];
        for (int lay = 0; lay < 4; ++lay) {
            C* d = raw_array(n2 * NP * NP + NP); for (int z = 0; z < NP; ++z) d[n2 * NP * NP + z] = zeroC();     // zero padding read by out-of-box terms
            for (int blk = 0; blk < NP; ++blk) for (size_t row = 0; row < n2; ++row) for (int m = 0; m < NP; ++m) {
                int iP, ik; mom_layout_source(lay, m, blk, L, iP, ik);
                d[m + (size_t)NP * (row + n2 * blk)] = c.K2[row + n2 * (iP + (size_t)NP * ik)];
            }
            c.K2m[lay] = d;
        }
        C* d3 = raw_array(n3 * NP + NP); for (int z = 0; z < NP; ++z) d3[n3 * NP + z] = zeroC();
        for (size_t row = 0; row < n3; ++row) for (int iP = 0; iP < NP; ++iP) d3[iP + (size_t)NP * row] = c.K3[row + n3 * iP];
        c.K3m = d3;
        C* d1 = raw_array((size_t)nB1 * NP);
        for (int pW = 0; pW < nB1; ++pW) for (int iP = 0; iP < NP; ++iP) d1[iP + (size_t)NP * pW] = c.K1[pW + (size_t)nB1 * iP];
        c.K1m = d1;
    }
}

// host restatement of slab_conv_kernel (same formulas, plain loops): X[k + NP * inu] for one slab
template <int KIND, int CH>
static std::vector<C> conv_host(const DevChain& V, const ColJob& job, const Grid& g, const C* Rs, int W, int iP) {
    typedef Forms<KIND, CH> FM;
    const int L = g.L, NP = g.NP, nF2 = 2 * g.nK2f, nw = job.nw, Nin = job.Ninner;
    const double tau = 6.283185307179586476925286766559;
    auto ph = [&](int j) { j = ((j % L) + L) % L; return mkC(std::cos(tau * j / L), std::sin(tau * j / L)); };
    std::vector<C> Rh((size_t)nw * NP), X((size_t)nF2 * NP, zeroC()), Z((size_t)nF2 * NP, zeroC());
    for (int iw = 0; iw < nw; ++iw) for (int kap = 0; kap < NP; ++kap) {
        C s = zeroC();
        for (int q = 0; q < NP; ++q) s += Rs[slab_at(iw, q, NP)] * ph(-((kap % L) * (q % L) + (kap / L) * (q / L)));
        Rh[iw + (size_t)nw * kap] = s;
    }
    const int Px = iP % L, Py = iP / L;
    for (int f = 0; f < FM::n; ++f) for (int l = job.lev_first; l < conv_level_end<KIND>(job); ++l) {
        if (!conv_level_on<KIND>(job, l)) continue;
        const DevLevel& lv = V.lev[l];
        const int nB1 = 2 * lv.nK1 - 1;
        for (int r = 0; r < 3; ++r) {
            if (r == FM::ch(f)) continue;
            std::vector<C> Kh((size_t)nB1 * NP);
            for (int i = 0; i < nB1; ++i) for (int kap = 0; kap < NP; ++kap) {
                C s = zeroC();
                for (int P = 0; P < NP; ++P) s += lv.ch[r].K1[i + nB1 * P] * ph(-((kap % L) * (P % L) + (kap / L) * (P / L)));
                Kh[kap + (size_t)NP * i] = s;
            }
            for (int inu = 0; inu < nF2; ++inu) for (int ko = 0; ko < NP; ++ko) {
                const ConvPiece pc = conv_piece<KIND, CH>(FM::ch(f), r, W, inu - g.nK2f, Px, Py);
                const int kx = fold1(pc.sk * (ko % L), L), ky = fold1(pc.sk * (ko / L), L);
                const int rx = fold1(-pc.sq * kx, L), ry = fold1(-pc.sq * ky, L);
                C s = zeroC();
                for (int iw = 0; iw < nw; ++iw) {
                    const int Wc = pc.W0 + pc.sW * (iw - Nin);
                    if (inB(Wc, lv.nK1)) s += Kh[kx + L * ky + (size_t)NP * posB(Wc, lv.nK1)] * Rh[iw + (size_t)nw * (rx + L * ry)];
                }
                Z[ko + (size_t)NP * inu] += s * ph(kx * pc.cx + ky * pc.cy) * FM::coef(f);
            }
        }
    }
    for (int inu = 0; inu < nF2; ++inu) for (int k = 0; k < NP; ++k) {
        C s = zeroC();
        for (int ko = 0; ko < NP; ++ko) s += Z[ko + (size_t)NP * inu] * ph((ko % L) * (k % L) + (ko / L) * (k / L));
        X[k + (size_t)NP * inu] = s * (1.0 / NP);
    }
    return X;
}

template <int KIND, int CH>
static double run_job(const DevChain& V, const Grid& g, int lev_first, int own_only, int Nin) {
    constexpr int SP = (CH == CH_T) ? SP_D : SP_P;
    const int NP = g.NP, L = g.L, nw = 2 * Nin, nB2 = 2 * g.nK2b - 1, nF2 = 2 * g.nK2f;
    const bool is_sde = (KIND == JOB_SDE_PP || KIND == JOB_SDE_PH);
    const int slabN = is_sde ? g.nPiB : g.nK2b, nBs = 2 * slabN - 1;     // (the K2 jobs are tested with W on the K2 mesh)
    ColJob job; job.lev_first = lev_first; job.n_nl2 = 0; while (job.n_nl2 < V.nlev && V.lev[job.n_nl2].type == LV_NL2) job.n_nl2++;
    job.own_only = own_only; job.nw = nw; job.Ninner = Nin; job.slabW_N = slabN; job.scale_re = 1.0; job.scale_im = 0.0; job.k1_direct = 0; job.slabmap = nullptr;
    const C* R = rnd_array((size_t)nw * NP * nBs * NP);
#ifdef DEVICE_CHECK
    C* Tm; cudaMallocManaged(&Tm, (size_t)nw * nF2 * nB2 * sizeof(C));
    struct { C* p; size_t n; C* data() { return p; } size_t size() { return n; } C& operator[](size_t i) { return p[i]; } } T = {Tm, (size_t)nw * nF2 * nB2};
    int* mi; cudaMallocManaged(&mi, 128 * sizeof(int)); C* dout; cudaMallocManaged(&dout, 128 * FDGA_WGROUP * sizeof(C));
    double maxdev = 0.0;
#else
    std::vector<C> T((size_t)nw * nF2 * nB2);
#endif
    for (size_t i = 0; i < T.size(); ++i) T[i] = (KIND == JOB_LK2 || KIND == JOB_LK2_LOC) ? zeroC() : loc_table_entry<KIND, CH>(V, job, g, (long long)i);
    double maxerr = 0.0, maxval = 0.0;
    for (int trial = 0; trial < 12; ++trial) {
        // one group: ng columns sharing (P, k), distinct W, each with its own representatives
        const int iP = rand() % NP, ik = rand() % NP;
        const int ng = 1 + rand() % std::min(FDGA_WGROUP, nB2);
        std::vector<int> iWv(ng), iPv(ng, iP), ikv(ng, ik), start(ng + 1, 0), inu, cls;
        {
            std::vector<int> perm(nB2); for (int i = 0; i < nB2; ++i) perm[i] = i;
            for (int i = 0; i < ng; ++i) { int j = i + rand() % (nB2 - i); std::swap(perm[i], perm[j]); iWv[i] = perm[i]; }
        }
        for (int i = 0; i < ng; ++i) {
            const int nrep = 1 + rand() % std::min(FDGA_NV, nF2);
            for (int n = 0; n < nrep; ++n) { inu.push_back((n * 3 + trial + i) % nF2); cls.push_back((int)cls.size()); }
            start[i + 1] = start[i] + nrep;
        }
        int gs[2] = {0, ng};
        ColDev cols; cols.ncol = ng; cols.iW = iWv.data(); cols.iP = iPv.data(); cols.ik = ikv.data(); cols.start = start.data();
        cols.rep_inu = inu.data(); cols.rep_cls = cls.data(); cols.ngrp = 1; cols.grp_start = gs;
        int maxrep = 1; for (int i = 0; i < ng; ++i) maxrep = std::max(maxrep, start[i + 1] - start[i]);
        int NVc = 1; while (NVc < maxrep) NVc <<= 1;
        const int ntot = start[ng];
        std::vector<C> got(ntot, zeroC());
        const int nthreads = 128;
        for (int tid = 0; tid < nthreads; ++tid) {
            C a[FDGA_WGROUP];
            column_thread<KIND, CH>(V, job, cols, R, g, 0, tid, nthreads, a);
            int n = tid & (NVc - 1);
            for (int i = 0; i < ng; ++i) if (n < start[i + 1] - start[i]) got[start[i] + n] += a[i];
        }
        if (KIND != JOB_LK2_LOC) for (int pass = 0; pass < 2; ++pass) {   // q-lane contraction (one warp per representative) against column_thread
            ColJob jq = job; jq.k1_direct = pass;
            std::vector<C> ref(ntot, zeroC());
            if (pass == 0) ref = got;
            else for (int tid = 0; tid < nthreads; ++tid) {
                C a[FDGA_WGROUP];
                column_thread<KIND, CH>(V, jq, cols, R, g, 0, tid, nthreads, a);
                int n = tid & (NVc - 1);
                for (int i = 0; i < ng; ++i) if (n < start[i + 1] - start[i]) ref[start[i] + n] += a[i];
            }
            for (int i = 0; i < ng; ++i) for (int n = start[i]; n < start[i + 1]; ++n) {
                C sq = zeroC();
                for (int lane = 0; lane < 32; ++lane) sq += qlane_lane<KIND, CH>(V, jq, g, R, iWv[i], inu[n], iP, ik, lane);
                const double d = std::max(std::fabs(sq.x - ref[n].x), std::fabs(sq.y - ref[n].y));
                if (d > 1e-11) printf("    QLANE != COLUMN pass %d trial %d rep %d: qlane (%g,%g) column (%g,%g)\n", pass, trial, n, sq.x, sq.y, ref[n].x, ref[n].y);
                maxerr = std::max(maxerr, d);
#ifdef DEVICE_CHECK
                if (pass == 0) {
                    int* r4; cudaMallocManaged(&r4, 4 * sizeof(int)); r4[0] = iWv[i]; r4[1] = inu[n]; r4[2] = iP; r4[3] = ik;
                    C* qo; cudaMallocManaged(&qo, 32 * sizeof(C));
                    dev_run_q<KIND, CH><<<1, 32>>>(V, jq, r4, 1, R, g, qo);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("CUDA error (qlane) %s\n", cudaGetErrorString(e)); exit(2); }
                    C dq = zeroC(); for (int lane = 0; lane < 32; ++lane) dq += qo[lane];
                    const double dd = std::max(std::fabs(dq.x - sq.x), std::fabs(dq.y - sq.y));
                    if (dd > 1e-10) printf("    DEVICE QLANE != HOST trial %d rep %d: dev (%g,%g) host (%g,%g)\n", trial, n, dq.x, dq.y, sq.x, sq.y);
                    maxdev = std::max(maxdev, dd);
                    cudaFree(r4); cudaFree(qo);
                }
#endif
            }
        }
#ifdef DEVICE_CHECK
        {   // the same column_thread on the device
            for (int i = 0; i < ng; ++i) { mi[i] = iWv[i]; mi[8 + i] = iP; mi[16 + i] = ik; }
            for (int i = 0; i <= ng; ++i) mi[24 + i] = start[i];
            for (int n = 0; n < ntot; ++n) { mi[32 + n] = inu[n]; mi[64 + n] = cls[n]; }
            mi[100] = 0; mi[101] = ng;
            ColDev dc; dc.ncol = ng; dc.iW = mi; dc.iP = mi + 8; dc.ik = mi + 16; dc.start = mi + 24; dc.rep_inu = mi + 32; dc.rep_cls = mi + 64; dc.ngrp = 1; dc.grp_start = mi + 100;
            dev_run<KIND, CH><<<1, nthreads>>>(V, job, dc, R, (KIND == JOB_LK2) ? nullptr : T.data(), g, dout);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(2); }
            std::vector<C> dgot(ntot, zeroC());
            for (int tid = 0; tid < nthreads; ++tid) { int n = tid & (NVc - 1); for (int i = 0; i < ng; ++i) if (n < start[i + 1] - start[i]) dgot[start[i] + n] += dout[tid + nthreads * i]; }
            for (int n = 0; n < ntot; ++n) {
                double d = std::max(std::fabs(dgot[n].x - got[n].x), std::fabs(dgot[n].y - got[n].y));
                if (d > 1e-10) printf("    DEVICE != HOST trial %d rep %d (iP %d ik %d ng %d): dev (%g,%g) host (%g,%g)\n", trial, n, iP, ik, ng, dgot[n].x, dgot[n].y, got[n].x, got[n].y);
                maxdev = std::max(maxdev, d);
            }
        }
#endif
        for (int gi = 0; gi < ng; ++gi) {
        const int iW = iWv[gi], nrep = start[gi + 1] - start[gi], off = start[gi];
        {   // hoisted own-channel / local-level pieces (slab_own_kernel + column epilogue)
            const int Wv = iW - (g.nK2b - 1);
            const C* Rs = R + (size_t)nw * NP * (posB(Wv, slabN) + (size_t)nBs * iP);
            if (KIND != JOB_LK2_LOC) {   // cross-channel K1 pieces: term by term, and through the momentum convolution
                std::vector<C> X = conv_host<KIND, CH>(V, job, g, Rs, Wv, iP);
#ifdef DEVICE_CHECK
                if (gi == 0) {   // slab_conv_kernel (and k1_dft_kernel) on the device against the host restatement, two tile sizes
                    DevChain Vd = V;
                    C* tw; cudaMallocManaged(&tw, L * sizeof(C));
                    for (int j = 0; j < L; ++j) tw[j] = mkC(std::cos(6.283185307179586 * j / L), std::sin(6.283185307179586 * j / L));
                    for (int l = 0; l < job.n_nl2; ++l) {
                        K1hOut o; const int nB1 = 2 * V.lev[l].nK1 - 1;
                        for (int r = 0; r < 3; ++r) { cudaMallocManaged(&o.p[r], (size_t)nB1 * NP * sizeof(C)); Vd.lev[l].ch[r].K1h = o.p[r]; }
                        k1_dft_kernel<<<dim3(nB1, 3), std::min(NP, 256), (size_t)(NP + L) * sizeof(C)>>>(V.lev[l], L, NP, tw, o);
                    }
                    int4* sl; cudaMallocManaged(&sl, sizeof(int4)); sl[0].x = iW; sl[0].y = iP; sl[0].z = sl[0].w = -1;
                    C* tab; cudaMallocManaged(&tab, (size_t)NP * nF2 * nB2 * NP * sizeof(C));
                    for (int pass = 0; pass < 2; ++pass) {
                        const int TW = pass == 0 ? std::max(nw, nF2) : nF2;
                        const size_t bytes = ((size_t)2 * NP * (TW | 1) + (size_t)nF2 * NP + L) * sizeof(C);
                        cudaFuncSetAttribute(slab_conv_kernel<KIND, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
                        slab_conv_kernel<KIND, CH><<<1, 256, bytes>>>(Vd, job, sl, R, tw, tab, g, TW, 0);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("CUDA error (slab_conv) %s\n", cudaGetErrorString(e)); exit(2); }
                        for (int inu2 = 0; inu2 < nF2; ++inu2) for (int k2 = 0; k2 < NP; ++k2) {
                            C dv = tab[k2 + (size_t)NP * (inu2 + nF2 * (iW + (size_t)nB2 * iP))], hv = X[k2 + (size_t)NP * inu2];
                            double d = std::max(std::fabs(dv.x - hv.x), std::fabs(dv.y - hv.y));
                            if (d > 1e-10) printf("    DEVICE CONV != HOST pass %d inu %d k %d: dev (%g,%g) host (%g,%g)\n", pass, inu2, k2, dv.x, dv.y, hv.x, hv.y);
                            maxdev = std::max(maxdev, d);
                        }
                    }
                    cudaFree(tab); cudaFree(sl); cudaFree(tw);
                }
#endif
                for (int n = 0; n < nrep; ++n) {
                    C d = k1_cross_direct<KIND, CH>(V, job, g, Rs, Wv, iP, ik, inu[off + n] - g.nK2f);
                    C x = X[ik + (size_t)NP * inu[off + n]];
                    double e = std::max(std::fabs(d.x - x.x), std::fabs(d.y - x.y));
                    if (e > 1e-11) { printf("    CONV != DIRECT trial %d rep %d: conv (%g,%g) direct (%g,%g)\n", trial, n, x.x, x.y, d.x, d.y); maxerr = std::max(maxerr, e); }
                    got[off + n] += x;
                }
            }
            C rtot = zeroC(); for (int i = 0; i < nw * NP; ++i) rtot += Rs[i];
            for (int n = 0; n < nrep; ++n)
                if (KIND != JOB_LK2 && KIND != JOB_LK2_LOC)
                    got[off + n] += slab_own_entry<KIND, CH>(V, job, g, Rs, T.data(), iW, iP, inu[off + n])
                            + own_B_term<KIND, CH>(V, job, g, Wv, iP, ik, inu[off + n] - g.nK2f) * rtot;
        }
        // brute force with the per-term evaluator (the arithmetic of bse_k2_kernel / bse_lk2_kernel / sde_L_kernel)
        const int W = iW - (g.nK2b - 1), Px = iP % L, Py = iP / L, kx = ik % L, ky = ik / L;
        const C* slab = R + (size_t)nw * NP * (posB(W, slabN) + (size_t)nBs * iP);
        for (int n = 0; n < nrep; ++n) {
            const int nu = inu[off + n] - g.nK2f;
            C ref = zeroC();
            for (int iq = 0; iq < NP; ++iq) for (int iw = 0; iw < nw; ++iw) {
                const int w = iw - Nin, qx = iq % L, qy = iq / L;
                Arg a; a.W = W; a.Px = Px; a.Py = Py;
                C d;
                if (KIND == JOB_K2 || KIND == JOB_K2_MF) {
                    a.v = nu; a.kx = kx; a.ky = ky;
                    if (KIND == JOB_K2_MF) { a.w = (CH == CH_P) ? W - w - 1 : w; a.qx = (CH == CH_P) ? Px - qx : qx; a.qy = (CH == CH_P) ? Py - qy : qy; }
                    else { a.w = w; a.qx = qx; a.qy = qy; }
                    C f1 = eval_vertex<false>(V, lev_first, CH, SP, a, FL_ALL); a.v = FDGA_INF;
                    d = f1 - eval_vertex<false>(V, lev_first, CH, SP, a, FL_ALL);
                } else if (KIND == JOB_LK2 || KIND == JOB_LK2_LOC) {
                    const unsigned FLG = (CH == CH_P ? 0u : FL_GP) | (CH == CH_T ? 0u : FL_GT) | (CH == CH_A ? 0u : FL_GA);
                    const bool cross = (KIND == JOB_LK2) && (CH == CH_P);
                    a.v = nu; a.kx = kx; a.ky = ky; a.w = cross ? W - w - 1 : w; a.qx = cross ? Px - qx : qx; a.qy = cross ? Py - qy : qy;
                    d = eval_vertex<false>(V, 0, CH, SP, a, FLG);
                } else {
                    // fused recursion: sum over levels l >= lev_first of the per-level integrand of sde_L_kernel (core level / 3)
                    d = zeroC();
                    for (int l = lev_first; l < V.nlev; ++l) {
                        const DevLevel& lv = V.lev[l];
                        C dl;
                        if (KIND == JOB_SDE_PP) {
                            a.v = W - w - 1; a.w = nu; a.kx = Px - qx; a.ky = Py - qy; a.qx = kx; a.qy = ky;
                            if (lv.type == LV_CORE) dl = (core_eval(lv, CH_P, SP_P, a.W, a.v, a.w) - lv.U) * (1.0 / 3.0);
                            else if (own_only) dl = eval_vertex<false>(V, l, CH_P, SP_P, a, FL_GP);
                            else dl = eval_vertex<false>(V, l, CH_P, SP_P, a, FL_F0 | FL_GP) - eval_vertex<false>(V, l + 1, CH_P, SP_P, a, FL_F0 | FL_GP);
                        } else {
                            a.v = nu; a.w = w; a.kx = kx; a.ky = ky; a.qx = qx; a.qy = qy;
                            if (lv.type == LV_CORE) dl = (core_eval(lv, CH_A, SP_P, W, nu, w) + core_eval(lv, CH_T, SP_P, W, nu, w) - lv.U - lv.U) * (1.0 / 3.0);
                            else if (own_only) dl = eval_vertex<false>(V, l, CH_A, SP_P, a, FL_GA) + eval_vertex<false>(V, l, CH_T, SP_P, a, FL_GT);
                            else dl = eval_vertex<false>(V, l, CH_A, SP_P, a, FL_F0 | FL_GA) + eval_vertex<false>(V, l, CH_T, SP_P, a, FL_F0 | FL_GT)
                                    - eval_vertex<false>(V, l + 1, CH_A, SP_P, a, FL_F0 | FL_GA) - eval_vertex<false>(V, l + 1, CH_T, SP_P, a, FL_F0 | FL_GT);
                        }
                        d += dl;
                    }
                }
                ref += d * slab[slab_at(iw, iq, NP)];
            }
            maxerr = std::max(maxerr, std::max(std::fabs(ref.x - got[off + n].x), std::fabs(ref.y - got[off + n].y)));
            maxval = std::max(maxval, std::max(std::fabs(ref.x), std::fabs(ref.y)));
        }
        }
    }
#ifdef DEVICE_CHECK
    maxerr = std::max(maxerr, maxdev);
#endif
    return maxerr / std::max(maxval, 1e-300);
}

static DevLevel make_level(int type, int nK1, int nK2b, int nK2f, int nK3b, int nK3f, int NP) {
    DevLevel lv; memset(&lv, 0, sizeof(lv));
    lv.type = type; lv.nK1 = nK1; lv.nK2b = nK2b; lv.nK2f = nK2f; lv.nK3b = nK3b; lv.nK3f = nK3f;
    int np = (type == LV_NL2) ? NP : 1;
    for (int r = 0; r < 3; ++r) {
        lv.ch[r].K1 = rnd_array((size_t)(2 * nK1 - 1) * np);
        lv.ch[r].K2 = rnd_array((size_t)(2 * nK2b - 1) * (2 * nK2f) * np * np);
        lv.ch[r].K3 = rnd_array((size_t)(2 * nK3b - 1) * (2 * nK3f) * (2 * nK3f) * np);
    }
    if (type == LV_NL2) { int L = 1; while (L * L < NP) ++L; build_mom_layouts(lv, L, NP); }
    return lv;
}

int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    srand(4242);
    double worst = 0.0;
    for (int cfg = 0; cfg < 5; ++cfg) {
        g_store.clear(); g_store.reserve(4096);
        const int L = (cfg == 0) ? 3 : (cfg == 1 ? 4 : (cfg == 2 ? 2 : (cfg == 3 ? 1 : 9))), NP = L * L;      // cfg 3: the local solver's 1 x 1 mesh; cfg 4: NP = 81, two momentum passes of the q-lane kernel
        Grid g; memset(&g, 0, sizeof(g));
        g.T = 0.3; g.L = L; g.NP = NP; g.nK1 = (cfg == 1) ? 6 : (cfg == 4 ? 8 : 5); g.nPiB = g.nK1; g.nPiF = g.nK1;
        g.nK2b = (cfg == 4) ? 6 : 3;      // cfg 4: 11-wide K2 bosonic box = two entry rounds of the q-lane kernel
        g.nK2f = (cfg == 1) ? 4 : (cfg == 3 ? 5 : 2); g.nK3b = 2; g.nK3f = 2;
        DevChain V; memset(&V, 0, sizeof(V)); V.L = L; V.NP = NP; V.nlev = 4;
        V.lev[0] = make_level(LV_NL2, g.nK1, g.nK2b, g.nK2f, g.nK3b, g.nK3f, NP);
        V.lev[1] = (cfg == 2) ? make_level(LV_NL2, 7, 4, 3, 2, 1, NP) : make_level(LV_NL2, g.nK1, g.nK2b, g.nK2f, g.nK3b, g.nK3f, NP);
        V.lev[2] = make_level(LV_LOCAL, 9, 5, 4, 1, 1, NP);
        DevLevel core; memset(&core, 0, sizeof(core)); core.type = LV_CORE; core.nK3b = 3; core.nK3f = 2; core.U = mkC(1.7, 0.0);
        for (int i = 0; i < 4; ++i) core.core[i] = rnd_array((size_t)5 * 4 * 4);
        V.lev[3] = core;
        double e = 0;
#define RUN(K, CHT, lf, oo, NN) { double x = run_job<K, CHT>(V, g, lf, oo, NN); printf("  cfg %d %-10s ch %d lev %d own %d : %.3e\n", cfg, #K, CHT, lf, oo, x); e = std::max(e, x); }
        RUN(JOB_K2, CH_P, 0, 0, g.nPiF) RUN(JOB_K2, CH_T, 0, 0, g.nPiF) RUN(JOB_K2, CH_A, 0, 0, g.nPiF)
        RUN(JOB_K2_MF, CH_P, 1, 0, g.nPiF) RUN(JOB_K2_MF, CH_T, 1, 0, g.nPiF) RUN(JOB_K2_MF, CH_A, 1, 0, g.nPiF)
        RUN(JOB_LK2, CH_P, 0, 0, g.nK2f) RUN(JOB_LK2, CH_T, 0, 0, g.nK2f) RUN(JOB_LK2, CH_A, 0, 0, g.nK2f)
        RUN(JOB_LK2_LOC, CH_P, 0, 0, g.nPiF) RUN(JOB_LK2_LOC, CH_T, 0, 0, g.nPiF) RUN(JOB_LK2_LOC, CH_A, 0, 0, g.nPiF)
        for (int lf = 0; lf < 4; ++lf) for (int oo = 0; oo < 2; ++oo) { RUN(JOB_SDE_PP, CH_P, lf, oo, g.nPiF) RUN(JOB_SDE_PH, CH_A, lf, oo, g.nPiF) }
        worst = std::max(worst, e);
    }
    printf("WORST %.3e\n", worst);
    return worst < 1e-11 ? 0 : 1;
}
