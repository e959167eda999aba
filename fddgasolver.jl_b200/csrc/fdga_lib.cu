// fdga_lib.cu -- libfdga.so: context, data movement, kernel orchestration and the C-ABI of include/fdga.h.
// Reference call structure being replaced: iterate_solver! (src/solve.jl:4-116), BSE_templates.jl:12-180,
// SDE! (src/SDE.jl:3-48), mfRGLinearMap (src/mfRG.jl:34-89).
#include "../../include/fdga.h"
#include "fdga_qlane.cuh"
#include "fdga_swave.cuh"
#include "fdga_krylov.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <dlfcn.h>

using namespace fdga;

extern "C" int64_t fdga_symgroup_build_host(int which_sg, int n0, int n1, int nq, int64_t* offsets, int64_t* index, uint8_t* ops);

static thread_local std::string g_create_error;

// ------------------------------------------------------------------------------------------------
struct LevelBuf {
    fdga_level_desc d;
    C* block;            // one allocation per level, laid out [p: K1,K2,K3][t: ...][a: ...] = the order of flatten(F)
    size_t blocklen;
    C* K[3][3];          // [channel][class] (pointers into block)
    size_t len[3];
    C* sw[3][4];         // [channel][K1sw, K2swk, K2sww, K3sw]
    C* K1h[3];           // [channel] momentum DFT of K1 (slab_conv_kernel)
    bool k1h_dirty;
    C* mom[3][ML_COUNT];  // [channel][layout] momentum-fastest copies of K2 / K3 / K1 (fdga_qlane.cuh), allocated by alloc_mom
    unsigned mom_valid[3]; // bit = layout whose copy is current
    C* core[4];
    size_t corelen;
    bool sw_dirty;
    bool mbe;             // MBEVertex / NL2_MBEVertex: same arrays as d.type (the BASE type), multi-boson-exchange evaluation
};
struct SymGroup {
    bool set;
    long long ncls, nmem;
    long long *d_offsets, *d_index;
    unsigned char* d_ops;
    int* d_member_class;
    C* d_repvals;         // current slot (= d_rep[channel]); padded to chunk * nranks
    C* d_rep[3];          // one slot per channel so that the all-gathers of p, a, t can be issued as one NCCL group
    long long chunk;
    std::vector<long long> h_offsets, h_index;            // internal class order (see rebuild_sg)
    std::vector<long long> o_offsets, o_index; std::vector<unsigned char> o_ops;   // as registered by the caller
    // columns (W, P, k) of this rank's class representatives (K2-shaped groups only)
    int ncol; int *d_col_iW, *d_col_iP, *d_col_ik, *d_col_start, *d_rep_inu, *d_rep_cls; int ngrp; int* d_grp_start;
    int nrep; int4* d_reps;   // the same representatives, slab-major, one per warp of qlane_kernel
};
struct TimedEvent { cudaEvent_t a, b; int cat; };
// a recorded sequence of calls (fdga_graph_begin / _end): the instantiated graph, the kernel launches it holds, and the lazy-state
// signature / allocation epoch it was recorded under (a replay is only valid from the same state)
struct GraphRec { cudaGraphExec_t exec; cudaGraph_t graph; long long launches; long long n_launch[FDGA_T_COUNT]; std::vector<long long> sig; long long epoch; bool live; };
struct Pending { SymGroup* s; C* rep; C* out; int kind, ch; bool expanded; };
enum { PI_NONE = 0, PI_GHAT = 1, PI_FULL = 2 };
enum { PK_K1 = 0, PK_LK2 = 1, PK_K2 = 2, PK_LK3 = 3, PK_K3 = 4, PK_K2_NOFL = 5 };

// NCCL through dlopen (no link-time dependency; the process may already hold torch's libnccl)
struct NcclApi {
    void* h;
    int (*GetUniqueId)(void*);
    void* CommInitRank;                                  // bound with the by-value ncclUniqueId signature at the call site
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t);
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
    int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t);
    int (*CommDestroy)(void*);
    int (*GroupStart)(); int (*GroupEnd)();
    const char* (*GetErrorString)(int);
};
struct UniqueId { char b[128]; };

struct fdga_ctx {
    fdga_dims dims;
    int device;
    cudaStream_t stream;
    Grid g;
    int nlev;
    LevelBuf lev[FDGA_MAX_LEVELS];
    LevelBuf FL, Fbuff;
    C* G[5]; size_t lenG;
    // bubbles.  Pi[i]: the reference's layout [W, v, P, k], allocated only when a caller reads / writes a bubble in that layout
    // or mixes bubbles; PiT[i]: COMPACT slab storage [q, w | slab] of the slabs this rank reads; Pisw[i] = mean_k.
    // pi_src: where the truth of bubble i lives (PI_NONE zeros, PI_GHAT product of the coarse-grained G, PI_FULL Pi[i])
    C* Pi[4]; C* PiT[4]; C* Pisw[4]; size_t lenPi, lenPisw; bool pi_dirty[4]; int pi_src[4]; bool pi_full_valid[4];
    C* cache[10]; size_t lenK3;
    C* L[2];              // Lpp, Lph (K2-shaped)
    C* Rt;                // hoisted right factor, bubble-sized (= RtL[0])
    // concurrency lanes: the three channels of a BSE stage (and the pp / ph / U^2 parts of the SDE) are independent, so
    // they are issued on three streams forked from / joined into the main stream; every lane owns its scratch tables
    cudaStream_t main_stream; cudaStream_t lane[3]; cudaEvent_t ev_fork, ev_join[3];
    cudaEvent_t ev_cache; bool cache_on_lane;      // fused driver: build_K3_cache! runs on a lane beside BSE_L_K2! / BSE_K1! / BSE_K2!
    int cur_lane; bool forked; int opt_serial;
    C* RtL[3]; C* TtabL[3]; C* OwnTabL[3]; C* RtotL[3]; C* ConvTabL[3];
    cudaStream_t copy_stream; cudaEvent_t ev_copy_ready, ev_copy_done; bool copy_pending;   // fdga_flatten_F_async
    C* SigR2;             // scratch of the U^2 term (its lane runs beside the real-space contraction)
    C* scratchA; C* scratchB; size_t lenScratch;   // bubble-sized ping-pong (DFTs)
    C* GR; C* GRm; C* SigR; C* SigTmp; C* SigAcc;  // G-sized scratch
    C* Ghat[2];           // coarse-grained G / G0 on the vertex momentum mesh [nu, p] (bubbles in product form)
    C* flat; size_t lenFlat;                       // flatten staging (device)
    C* flat2;
    C* stash;                                      // fdga_stash_F / fdga_unstash_F
    double* d_occ;
    SymGroup sg[FDGA_SG_COUNT];
    // multi-GPU
    int nranks, rank; void* comm; NcclApi nccl;
    // profiling
    bool profile; std::vector<TimedEvent> events; std::vector<cudaEvent_t> pool;
    double t_ms[FDGA_T_COUNT]; long long n_launch[FDGA_T_COUNT]; long long total_launches;
    int cur_cat; cudaEvent_t cur_a;
    int opt_sde_own_gamma;   // FDGA_OPT_SDE_OWN_GAMMA
    int opt_generic;         // FDGA_OPT_GENERIC_KERNELS
    int opt_hartree_once;    // FDGA_OPT_FD_HARTREE_ONCE
    int opt_local;           // FDGA_OPT_LOCAL_SOLVER
    bool defer; std::vector<struct Pending> pending;   // batched SG finishes (one NCCL group per BSE stage)
    int n_nl2;               // leading NL2 levels of the F chain
    int opt_direct_k1;       // FDGA_OPT_DIRECT_K1
    int opt_qlane;           // FDGA_OPT_QLANE
    unsigned mom_mask[3];    // momentum layouts (per table channel) any q-lane job reads
    // per lane: TtabL = momentum-independent left-factor table [nw, nF2, nB2]; OwnTabL[nu | W, P], RtotL[W, P] = per-slab hoisted
    // pieces (W on the K2 mesh); ConvTabL[k, nu | W, P] = cross-channel K1 pieces (slab_conv_kernel)
    C* twL; C* twLG;         // DFT twiddles exp(2 pi i j / n) for n = L, LG
    LevelBuf Fsum; bool has_fsum, fsum_dirty;   // K tables of lev[0] + lev[1] when both are NL2 on identical meshes
    int4* d_slabs[4]; int n_slabs[4]; bool slabs_dirty;   // active (W,P) slabs: [pp|ph] x [bubble mesh | K2 mesh]
    int* d_slabmap[4];    // (iW + nB * iP) -> position in d_slabs[kind] (-1: not held by this rank)
    size_t lenRtL;        // elements of RtL[i]
    C* scratchBig;        // bubble-sized scratch of the legacy bubble route (FDGA_BUBBLES_RS=1), allocated on demand
    // device-resident DQGMRES workspace (fdga_mfrg_dqgmres): rings of `kry_mem` basis / direction vectors, work vector, iterate
    C* kryV; C* kryP; C* kryW; C* kryX; C* kryH; C* kryPart; unsigned int* kryTicket; C* kryHhost; int kry_mem;
    C* itpA; C* itpB; size_t lenItp;   // ping-pong of fdga_interpolate_* when the bubble-sized scratch is too small (coarsening)
    C* PiMixed[2];           // Pipp_mixed, Piph_mixed of solve_using_mfRG! (fdga_mix_bubbles / fdga_update_reference)
    // s-wave solver (NL_ParquetSolver): lev[0] is an FDGA_LV_NL level; bubbles are Pi[W,w,P] and live in Pisw[] (fdga_swave.cuh)
    bool swave; C* swScratch[2];
    bool mbe;             // some level of the S.F chain is an MBE vertex: generic per-term kernels only (the evaluator is nonlinear)
    // CUDA graphs
    bool capturing; long long epoch; std::vector<GraphRec> graphs; std::vector<long long> cap_sig; long long cap_launches0; long long cap_n0[FDGA_T_COUNT];
    C* Rt3[3]; int rt_kind[3]; // per-channel right factors (W on the bubble mesh) reused between BSE_K1! and BSE_K2!
    std::string err, launch_err;
};

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_) + (ctx->launch_err.empty() ? "" : " [first failed launch: " + ctx->launch_err + "]"); \
    ctx->launch_err.clear(); return 1; } } while (0)
// remember the first kernel whose launch was rejected (bad configuration / shared memory request), for the error text
#define NOTE_LAUNCH(name) do { if (cudaPeekAtLastError() != cudaSuccess && ctx->launch_err.empty()) \
    ctx->launch_err = std::string(name) + ": " + cudaGetErrorString(cudaPeekAtLastError()); } while (0)
#define FAIL(msg) do { ctx->err = (msg); return 1; } while (0)

static void auto_forget(fdga_ctx* ctx);
static void invalidate_rt(fdga_ctx* ctx) { ctx->rt_kind[0] = ctx->rt_kind[1] = ctx->rt_kind[2] = -1; }
static inline unsigned nblk(long long n, int b) { return (unsigned)((n + b - 1) / b); }

// profiling scope: CUDA events on the launching stream around a group of launches
struct Scope {
    fdga_ctx* c; int cat; bool on; size_t idx;
    Scope(fdga_ctx* c_, int cat_) : c(c_), cat(cat_), on(false), idx(0) {
        if (c->profile && c->cur_cat < 0) {
            on = true; c->cur_cat = cat;
            TimedEvent ev; ev.cat = cat;
            cudaEventCreate(&ev.a); cudaEventCreate(&ev.b);
            cudaEventRecord(ev.a, c->stream);
            idx = c->events.size();
            c->events.push_back(ev);
        }
    }
    ~Scope() {
        if (on) { cudaEventRecord(c->events[idx].b, c->stream); c->cur_cat = -1; }
    }
};
#define LAUNCH(cat, kernel, grid, block, ...) do { \
    kernel<<<grid, block, 0, ctx->stream>>>(__VA_ARGS__); \
    NOTE_LAUNCH(#kernel); \
    ctx->n_launch[cat]++; ctx->total_launches++; } while (0)

static size_t lenK(const fdga_level_desc& d, int cls, int NP) {
    size_t nB1 = 2 * d.nK1 - 1, nB2 = 2 * d.nK2[0] - 1, nF2 = 2 * d.nK2[1], nB3 = 2 * d.nK3[0] - 1, nF3 = 2 * d.nK3[1];
    const bool mom = d.type == FDGA_LV_NL2 || d.type == FDGA_LV_NL;
    if (cls == 0) return nB1 * (mom ? NP : 1);
    if (cls == 1) return nB2 * nF2 * (d.type == FDGA_LV_NL2 ? (size_t)NP * NP : (mom ? (size_t)NP : 1));
    return nB3 * nF3 * nF3 * (mom ? NP : 1);
}

static int alloc_level(fdga_ctx* ctx, LevelBuf& lb, const fdga_level_desc& d_in) {
    memset(&lb, 0, sizeof(lb));
    fdga_level_desc d = d_in;
    lb.mbe = d.type == FDGA_LV_NL2_MBE || d.type == FDGA_LV_LOCAL_MBE;
    if (d.type == FDGA_LV_NL2_MBE) d.type = FDGA_LV_NL2;
    if (d.type == FDGA_LV_LOCAL_MBE) d.type = FDGA_LV_LOCAL;
    lb.d = d; lb.sw_dirty = true; lb.k1h_dirty = true; lb.mom_valid[0] = lb.mom_valid[1] = lb.mom_valid[2] = 0;
    int NP = ctx->g.NP;
    if (d.type == FDGA_LV_CORE) {
        lb.corelen = (size_t)(2 * d.nK3[0] - 1) * (2 * d.nK3[1]) * (2 * d.nK3[1]);
        if (d.nK3[0] <= 0 || d.nK3[1] <= 0) lb.corelen = 0;
        for (int i = 0; i < 4; i++) if (lb.corelen) { CK(cudaMalloc(&lb.core[i], lb.corelen * sizeof(C))); CK(cudaMemsetAsync(lb.core[i], 0, lb.corelen * sizeof(C), ctx->stream)); }
        return 0;
    }
    bool nl2 = d.type == FDGA_LV_NL2;
    for (int cls = 0; cls < 3; cls++) lb.len[cls] = lenK(d, cls, NP);
    lb.blocklen = 3 * (lb.len[0] + lb.len[1] + lb.len[2]);
    CK(cudaMalloc(&lb.block, lb.blocklen * sizeof(C))); CK(cudaMemsetAsync(lb.block, 0, lb.blocklen * sizeof(C), ctx->stream));
    {
        size_t off = 0;
        for (int ch = 0; ch < 3; ch++) for (int cls = 0; cls < 3; cls++) { lb.K[ch][cls] = lb.block + off; off += lb.len[cls]; }
    }
    for (int ch = 0; ch < 3; ch++) {
        if (nl2) {
            size_t n[4] = { (size_t)(2 * d.nK1 - 1), (size_t)(2 * d.nK2[0] - 1) * (2 * d.nK2[1]) * NP,
                            (size_t)(2 * d.nK2[0] - 1) * (2 * d.nK2[1]), (size_t)(2 * d.nK3[0] - 1) * (2 * d.nK3[1]) * (2 * d.nK3[1]) };
            for (int j = 0; j < 4; j++) CK(cudaMalloc(&lb.sw[ch][j], n[j] * sizeof(C)));
            CK(cudaMalloc(&lb.K1h[ch], n[0] * NP * sizeof(C)));
        } else if (d.type == FDGA_LV_NL) {
            // NL level: the table of fermionic-momentum means K2swk[W,v,P] IS K2 (dev_level aliases it); K1sw, K2sww, K3sw are P-means
            size_t n[4] = { (size_t)(2 * d.nK1 - 1), 0, (size_t)(2 * d.nK2[0] - 1) * (2 * d.nK2[1]), (size_t)(2 * d.nK3[0] - 1) * (2 * d.nK3[1]) * (2 * d.nK3[1]) };
            for (int j = 0; j < 4; j++) if (n[j]) CK(cudaMalloc(&lb.sw[ch][j], n[j] * sizeof(C)));
        }
    }
    return 0;
}
static void free_level(LevelBuf& lb) {
    cudaFree(lb.block);
    for (int ch = 0; ch < 3; ch++) for (int j = 0; j < 4; j++) cudaFree(lb.sw[ch][j]);
    for (int ch = 0; ch < 3; ch++) cudaFree(lb.K1h[ch]);
    for (int ch = 0; ch < 3; ch++) for (int j = 0; j < ML_COUNT; j++) cudaFree(lb.mom[ch][j]);
    for (int i = 0; i < 4; i++) cudaFree(lb.core[i]);
}

static DevLevel dev_level(const LevelBuf& lb) {
    DevLevel d; memset(&d, 0, sizeof(d));
    d.type = lb.d.type == FDGA_LV_NL ? (int)LV_NL2 : lb.d.type;      // evaluated through the kSW forms only (fdga_swave.cuh)
    d.mbe = lb.mbe ? 1 : 0;
    d.nK1 = lb.d.nK1; d.nK2b = lb.d.nK2[0]; d.nK2f = lb.d.nK2[1]; d.nK3b = lb.d.nK3[0]; d.nK3f = lb.d.nK3[1];
    d.U = mkC(lb.d.U_re, lb.d.U_im);
    for (int ch = 0; ch < 3; ch++) {
        d.ch[ch].K1 = lb.K[ch][0]; d.ch[ch].K2 = lb.K[ch][1]; d.ch[ch].K3 = lb.K[ch][2];
        d.ch[ch].K1sw = lb.sw[ch][0]; d.ch[ch].K2swk = lb.sw[ch][1]; d.ch[ch].K2sww = lb.sw[ch][2]; d.ch[ch].K3sw = lb.sw[ch][3];
        d.ch[ch].K1h = lb.K1h[ch];
        if (lb.d.type == FDGA_LV_NL) d.ch[ch].K2swk = lb.K[ch][1];
        for (int j = 0; j < 4; j++) d.ch[ch].K2m[j] = lb.mom[ch][j];
        d.ch[ch].K3m = lb.mom[ch][ML_K3]; d.ch[ch].K1m = lb.mom[ch][ML_K1];
    }
    for (int i = 0; i < 4; i++) d.core[i] = lb.core[i];
    if (lb.d.type == FDGA_LV_CORE && lb.corelen == 0) { d.nK3b = 0; d.nK3f = 0; }
    return d;
}
static DevLevel null_core() {
    DevLevel d; memset(&d, 0, sizeof(d)); d.type = LV_CORE; d.U = zeroC(); return d;
}
// chain of S.F starting at `from` (0 -> S.F, 1 -> S.F0)
static DevChain chain_F(fdga_ctx* ctx, int from) {
    DevChain c; memset(&c, 0, sizeof(c));
    c.L = ctx->g.L; c.NP = ctx->g.NP; c.nlev = ctx->nlev - from;
    for (int l = from; l < ctx->nlev; l++) c.lev[l - from] = dev_level(ctx->lev[l]);
    return c;
}
static DevChain chain_FL(fdga_ctx* ctx) {
    DevChain c; memset(&c, 0, sizeof(c));
    c.L = ctx->g.L; c.NP = ctx->g.NP; c.nlev = 2;
    c.lev[0] = dev_level(ctx->FL); c.lev[1] = null_core();
    return c;
}
// chain [lev0 + lev1, lev2, ...] used as the left vertex of BSE_K2! (fd) when the two NL2 levels share their meshes
static DevChain chain_F_merged(fdga_ctx* ctx) {
    DevChain c; memset(&c, 0, sizeof(c));
    c.L = ctx->g.L; c.NP = ctx->g.NP; c.nlev = ctx->nlev - 1;
    c.lev[0] = dev_level(ctx->Fsum);
    for (int l = 2; l < ctx->nlev; l++) c.lev[l - 1] = dev_level(ctx->lev[l]);
    return c;
}
static SymDev sym_dev(const SymGroup& s) {
    SymDev d; d.ncls = s.ncls; d.nmem = s.nmem; d.offsets = s.d_offsets; d.index = s.d_index; d.ops = s.d_ops; d.member_class = s.d_member_class;
    return d;
}
static C bareU(fdga_ctx* ctx) { const fdga_level_desc& d = ctx->lev[ctx->nlev - 1].d; return mkC(d.U_re, d.U_im); }

// ------------------------------------------------------------------------------------------------
// auxiliaries kept current lazily
static int refresh_swave(fdga_ctx* ctx) {
    for (int l = 0; l < ctx->nlev; l++) {          // S.FL is never evaluated at kSW: its tables are not needed
        LevelBuf& lb = ctx->lev[l];
        if ((lb.d.type != FDGA_LV_NL2 && lb.d.type != FDGA_LV_NL) || !lb.sw_dirty) continue;
        Scope sc(ctx, FDGA_T_SWAVE);
        DevLevel dl = dev_level(lb);
        SwOut out; for (int ch = 0; ch < 3; ch++) for (int j = 0; j < 4; j++) out.p[ch][j] = lb.sw[ch][j];
        if (lb.d.type == FDGA_LV_NL) {
            long long n = (long long)(2 * lb.d.nK1 - 1) + (long long)(2 * lb.d.nK2[0] - 1) * (2 * lb.d.nK2[1])
                        + (long long)(2 * lb.d.nK3[0] - 1) * (2 * lb.d.nK3[1]) * (2 * lb.d.nK3[1]);
            LAUNCH(FDGA_T_SWAVE, swave_tables_nl_kernel, dim3(nblk(n * 32, 128), 3), 128, dl, ctx->g.NP, out);
            lb.sw_dirty = false;
            continue;
        }
        long long n = (long long)(2 * lb.d.nK1 - 1) + (long long)(2 * lb.d.nK2[0] - 1) * (2 * lb.d.nK2[1]) * ctx->g.NP
                    + (long long)(2 * lb.d.nK3[0] - 1) * (2 * lb.d.nK3[1]) * (2 * lb.d.nK3[1]);
        long long n2 = (long long)(2 * lb.d.nK2[0] - 1) * (2 * lb.d.nK2[1]);
        LAUNCH(FDGA_T_SWAVE, swave_tables_kernel, dim3(nblk(n, 128), 3), 128, dl, ctx->g.NP, out);
        LAUNCH(FDGA_T_SWAVE, swave_tables2_kernel, dim3(nblk(n2, 64), 3), 64, dl, ctx->g.NP, out);
        lb.sw_dirty = false;
    }
    CK(cudaGetLastError());
    return 0;
}
static int refresh_fsum(fdga_ctx* ctx) {
    if (!ctx->has_fsum || !ctx->fsum_dirty) return 0;
    Scope sc(ctx, FDGA_T_MISC);
    LAUNCH(FDGA_T_MISC, axpby_kernel, nblk(ctx->Fsum.blocklen, 256), 256, ctx->Fsum.block, ctx->lev[0].block, 1.0, ctx->lev[1].block, 1.0, (long long)ctx->Fsum.blocklen);
    CK(cudaGetLastError());
    ctx->fsum_dirty = false; ctx->Fsum.k1h_dirty = true; ctx->Fsum.mom_valid[0] = ctx->Fsum.mom_valid[1] = ctx->Fsum.mom_valid[2] = 0;
    return 0;
}
// momentum DFT of the K1 tables of every NL2 level a column job may read (and of the merged level)
static int refresh_k1h(fdga_ctx* ctx) {
    for (int l = -1; l < ctx->nlev; l++) {
        if (l < 0 && !ctx->has_fsum) continue;
        LevelBuf& lb = l < 0 ? ctx->Fsum : ctx->lev[l];
        if (lb.d.type != FDGA_LV_NL2 || !lb.k1h_dirty) continue;
        if (l < 0 && ctx->fsum_dirty) continue;
        Scope sc(ctx, FDGA_T_SWAVE);
        K1hOut out; for (int ch = 0; ch < 3; ch++) out.p[ch] = lb.K1h[ch];
        long long n = (long long)(2 * lb.d.nK1 - 1) * ctx->g.NP;
        (void)n;
        k1_dft_kernel<<<dim3(2 * lb.d.nK1 - 1, 3), std::min(ctx->g.NP, 256), (size_t)(ctx->g.NP + ctx->g.L) * sizeof(C), ctx->stream>>>(dev_level(lb), ctx->g.L, ctx->g.NP, ctx->twL, out);
        ctx->n_launch[FDGA_T_SWAVE]++; ctx->total_launches++;
        lb.k1h_dirty = false;
    }
    CK(cudaGetLastError());
    return 0;
}
// ---- q-lane kernel: momentum-fastest copies of the vertex tables -------------------------------------------------------
static bool qlane_enabled(fdga_ctx* ctx) {
    if (ctx->swave || ctx->opt_local || ctx->opt_generic || ctx->opt_qlane == 0 || ctx->g.L > 64) return false;      // rep descriptors pack momenta in 8 / 16 bits
    if (ctx->opt_qlane == 1) return true;
    return ctx->g.NP >= 16;
}
// buffers are allocated once (device pointers inside DevChain copies must stay valid); contents are refreshed lazily
static int alloc_mom(fdga_ctx* ctx, LevelBuf& lb) {
    if (lb.d.type != FDGA_LV_NL2) return 0;
    for (int ch = 0; ch < 3; ch++) for (int j = 0; j < ML_COUNT; j++) {
        if (!(ctx->mom_mask[ch] & (1u << j)) || lb.mom[ch][j]) continue;
        const size_t n = j == ML_K1 ? lb.len[0] : (j == ML_K3 ? lb.len[2] : lb.len[1]);
        if (n + ctx->g.NP >= (size_t)1 << 31) FAIL("q-lane kernel: vertex table too large for 32-bit element offsets");
        CK(cudaMalloc(&lb.mom[ch][j], (n + ctx->g.NP) * sizeof(C)));      // + NP zeros: the row read by terms outside their Matsubara box
        CK(cudaMemsetAsync(lb.mom[ch][j] + n, 0, (size_t)ctx->g.NP * sizeof(C), ctx->stream));
    }
    return 0;
}
static int alloc_mom_all(fdga_ctx* ctx) {
    for (int l = 0; l < ctx->n_nl2; l++) if (alloc_mom(ctx, ctx->lev[l])) return 1;
    if (ctx->has_fsum && alloc_mom(ctx, ctx->Fsum)) return 1;
    return 0;
}
static int ensure_mom(fdga_ctx* ctx, LevelBuf& lb) {
    if (lb.d.type != FDGA_LV_NL2) return 0;
    MomOut out; bool any = false;
    for (int ch = 0; ch < 3; ch++) for (int j = 0; j < ML_COUNT; j++) {
        const bool need = (ctx->mom_mask[ch] & (1u << j)) && !(lb.mom_valid[ch] & (1u << j));
        out.p[ch][j] = need ? lb.mom[ch][j] : nullptr;
        if (need && !lb.mom[ch][j]) FAIL("q-lane kernel: momentum layouts not allocated");
        any = any || need;
    }
    if (!any) return 0;
    Scope sc(ctx, FDGA_T_SWAVE);
    LAUNCH(FDGA_T_SWAVE, mom_layout_kernel, dim3(nblk((long long)std::max(lb.len[0], std::max(lb.len[1], lb.len[2])), 256), 3 * ML_COUNT), 256, dev_level(lb), ctx->g.L, ctx->g.NP, out);
    CK(cudaGetLastError());
    for (int ch = 0; ch < 3; ch++) lb.mom_valid[ch] = ctx->mom_mask[ch];
    return 0;
}
// `need`: bit l = level l of the S.F chain, MOM_FSUM = the merged level; only tables a coming job reads are re-laid out
enum : unsigned { MOM_FSUM = 1u << 30, MOM_ALL = ~0u };
static int refresh_mom_all(fdga_ctx* ctx, unsigned need = MOM_ALL) {
    if (!qlane_enabled(ctx)) return 0;
    for (int l = 0; l < ctx->n_nl2; l++) if ((need & (1u << l)) && ensure_mom(ctx, ctx->lev[l])) return 1;
    if ((need & MOM_FSUM) && ctx->has_fsum && !ctx->fsum_dirty && ensure_mom(ctx, ctx->Fsum)) return 1;
    return 0;
}
static int ensure_slabs(fdga_ctx* ctx);
// the bubble `which` in the reference's layout, materialised on demand
static int ensure_pi_full(fdga_ctx* ctx, int which) {
    if (ctx->swave) FAIL("internal: the s-wave solver has no four-index bubbles");
    if (!ctx->Pi[which]) { CK(cudaMalloc(&ctx->Pi[which], ctx->lenPi * sizeof(C))); ctx->pi_full_valid[which] = false; }
    if (ctx->pi_full_valid[which]) return 0;
    if (ctx->pi_src[which] == PI_GHAT) {
        const int other = which ^ 1;                       // the pp / ph partner of the same Green function is produced alongside
        if (!ctx->Pi[other]) { CK(cudaMalloc(&ctx->Pi[other], ctx->lenPi * sizeof(C))); ctx->pi_full_valid[other] = false; }
        const bool fill_other = ctx->pi_src[other] == PI_GHAT && !ctx->pi_full_valid[other];
        C* tmp = nullptr;
        if (!fill_other) CK(cudaMalloc(&tmp, ctx->lenPi * sizeof(C)));      // partner holds explicit data: do not overwrite it
        C* pp = (which % 2 == 0) ? ctx->Pi[which] : (fill_other ? ctx->Pi[other] : tmp);
        C* ph = (which % 2 == 1) ? ctx->Pi[which] : (fill_other ? ctx->Pi[other] : tmp);
        LAUNCH(FDGA_T_BUBBLE, bubbles_product_kernel, nblk(ctx->lenPi, 256), 256, ctx->Ghat[which < 2 ? 1 : 0], pp, ph, ctx->g);
        CK(cudaGetLastError());
        if (tmp) { CK(cudaStreamSynchronize(ctx->stream)); cudaFree(tmp); }
        if (fill_other) ctx->pi_full_valid[other] = true;
    } else if (ctx->pi_src[which] == PI_NONE) {
        CK(cudaMemsetAsync(ctx->Pi[which], 0, ctx->lenPi * sizeof(C), ctx->stream));
    }
    ctx->pi_full_valid[which] = true;
    return 0;
}
// compact slabs + s-wave mean of bubble `which`, refreshed lazily
static int refresh_pi(fdga_ctx* ctx, int which) {
    if (ctx->swave) return 0;       // Pisw[which] is the bubble itself
    if (ensure_slabs(ctx)) return 1;
    if (!ctx->pi_dirty[which]) return 0;
    Scope sc(ctx, FDGA_T_MISC);
    const int nB = 2 * ctx->g.nPiB - 1, nF = 2 * ctx->g.nPiF, NP = ctx->g.NP;
    const int pp = (which % 2 == 0) ? 1 : 0, kind = pp ? 0 : 1, nsl = ctx->n_slabs[kind];
    const long long nel = (long long)nF * NP * nsl;
    if (ctx->pi_src[which] == PI_GHAT) {
        const C* Ghat = ctx->Ghat[which < 2 ? 1 : 0];
        if (nsl) LAUNCH(FDGA_T_MISC, bubble_slabs_kernel, nblk(nel, 256), 256, Ghat, ctx->PiT[which], ctx->g, pp, ctx->d_slabs[kind], nsl);
        LAUNCH(FDGA_T_MISC, bubble_swave_kernel, nblk(ctx->lenPisw, 128), 128, Ghat, ctx->Pisw[which], ctx->g, pp);
    } else if (ctx->pi_src[which] == PI_FULL) {
        if (nsl) LAUNCH(FDGA_T_MISC, pi_gather_slabs_kernel, nblk(nel, 256), 256, ctx->Pi[which], ctx->PiT[which], nB, nF, NP, ctx->d_slabs[kind], nsl);
        LAUNCH(FDGA_T_MISC, pi_swave_kernel, nblk(ctx->lenPisw, 256), 256, ctx->Pi[which], ctx->Pisw[which], nB, nF, NP);
    } else {
        if (nsl) CK(cudaMemsetAsync(ctx->PiT[which], 0, (size_t)nel * sizeof(C), ctx->stream));
        CK(cudaMemsetAsync(ctx->Pisw[which], 0, ctx->lenPisw * sizeof(C), ctx->stream));
    }
    CK(cudaGetLastError());
    ctx->pi_dirty[which] = false;
    return 0;
}

// SG finish: (all-gather of representative values) + expansion to all class members
// contiguous block partition of the class representatives over the ranks (every rank gets `chunk` slots, the
// last ones may be short or empty); the same arithmetic shards the reference's mpi_split(1:length) ranges
extern "C" int fdga_partition(int64_t nclasses, int nranks, int rank, int64_t* c0, int64_t* c1, int64_t* chunk) {
    if (nranks < 1 || rank < 0 || rank >= nranks || nclasses < 0) return 1;
    int64_t ch = (nclasses + nranks - 1) / nranks;
    int64_t a = (int64_t)rank * ch, b = a + ch;
    if (a > nclasses) a = nclasses;
    if (b > nclasses) b = nclasses;
    *c0 = a; *c1 = b; if (chunk) *chunk = ch;
    return 0;
}
static int sg_class_range(fdga_ctx* ctx, const SymGroup& s, long long& c0, long long& c1) {
    int64_t a, b;
    fdga_partition(s.ncls, ctx->nranks, ctx->rank, &a, &b, nullptr);
    c0 = a; c1 = b;
    return 0;
}
static int sg_finish(fdga_ctx* ctx, SymGroup& s, C* out) {
    if (ctx->nranks > 1) {
        Scope sc(ctx, FDGA_T_COMM);
        int rc = ctx->nccl.AllGather(s.d_repvals + (size_t)ctx->rank * s.chunk, s.d_repvals, (size_t)s.chunk * 2, /*ncclDouble*/ 8, ctx->comm, ctx->stream);
        if (rc != 0) FAIL(std::string("ncclAllGather: ") + ctx->nccl.GetErrorString(rc));
        ctx->n_launch[FDGA_T_COMM]++;
    }
    Scope sc(ctx, FDGA_T_EXPAND);
    LAUNCH(FDGA_T_EXPAND, expand_kernel, nblk(s.nmem, 256), 256, out, s.d_repvals, sym_dev(s));
    CK(cudaGetLastError());
    return 0;
}
#define NEED_SG(which) do { if (!ctx->sg[which].set) FAIL("symmetry group " #which " not set (call fdga_set_symmetry_classes)"); } while (0)

static int axpby(fdga_ctx* ctx, C* out, const C* x, double a, const C* y, double b, size_t n) {
    LAUNCH(FDGA_T_MISC, axpby_kernel, nblk(n, 256), 256, out, x, a, y, b, (long long)n);
    CK(cudaGetLastError()); return 0;
}
static int add_axpby(fdga_ctx* ctx, C* out, const C* x, double a, const C* y, double b, size_t n) {
    LAUNCH(FDGA_T_MISC, add_axpby_kernel, nblk(n, 256), 256, out, x, a, y, b, (long long)n);
    CK(cudaGetLastError()); return 0;
}
// t-channel post-fix: X_t <- (X_t + X_a) / 2   (BSE_templates.jl:35-38 etc.)
static int tfix(fdga_ctx* ctx, C* Xt, const C* Xa, size_t n) { return axpby(ctx, Xt, Xt, 0.5, Xa, 0.5, n); }

// one-launch 2-d transform over two adjacent axes (dft2_tile_kernel); false: the tile does not fit in shared memory / switched off
static bool dft2_tile(fdga_ctx* ctx, const C* in, C* out, long long pre, int n, long long post, int sgn, double scale, const C* tw, int cat) {
    // opt-in (FDGA_DFT_TILE=1): measured performance-neutral (these transforms overlap with the contraction lanes), and the two-pass
    // order of the axis kernels is the one the committed multi-GPU fingerprints were taken with
    static const bool on = getenv("FDGA_DFT_TILE") ? atoi(getenv("FDGA_DFT_TILE")) != 0 : false;
    const size_t smem = ((size_t)2 * n * n + n) * sizeof(C);
    // only worth it for many small tiles (the momentum axes of the vertex mesh): a G-sized transform (n = LG = 48, 2 N_G = 32 tiles)
    // would run on 32 CTAs and is 3-4x slower than the two axis passes with one thread per output (measured)
    static const int nmax = getenv("FDGA_DFT_TILE_NMAX") ? atoi(getenv("FDGA_DFT_TILE_NMAX")) : 16;
    if (!on || n > nmax || pre * post < 64 || smem > 200 * 1024 || pre * post > 0x7fffffffLL) return false;
    if (smem > 48 * 1024) {
        static size_t attr = 0;
        if (smem > attr) { if (cudaFuncSetAttribute(dft2_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return false; } attr = smem; }
    }
    const int thr = n * n >= 1024 ? 512 : (n * n >= 256 ? 256 : (n * n >= 64 ? 64 : 32));
    dft2_tile_kernel<<<(unsigned)(pre * post), thr, smem, ctx->stream>>>(in, out, pre, n, post, sgn, scale, tw);
    NOTE_LAUNCH("dft2_tile_kernel");
    ctx->n_launch[cat]++; ctx->total_launches++;
    return true;
}
// 2-d DFT of a G-shaped array over its momentum axes: in -> out (tmp used), out *= scale
static int dft2_G(fdga_ctx* ctx, const C* in, C* out, C* tmp, int sgn, double scale, int cat) {
    long long nGf = 2 * ctx->g.nG, LG = ctx->g.LG;
    long long n = nGf * LG * LG;
    if (dft2_tile(ctx, in, out, nGf, (int)LG, 1, sgn, scale, ctx->twLG, cat)) { CK(cudaGetLastError()); return 0; }
    LAUNCH(cat, dft_axis_kernel, nblk(n, 128), 128, in, tmp, nGf, (int)LG, LG, sgn, 1.0, ctx->twLG);
    LAUNCH(cat, dft_axis_kernel, nblk(n, 128), 128, tmp, out, nGf * LG, (int)LG, 1LL, sgn, scale, ctx->twLG);
    CK(cudaGetLastError()); return 0;
}
// 4-d DFT over the momentum axes of a [pre, L, L, L, L] array; result ends up in `a` (b = scratch)
static int dft4(fdga_ctx* ctx, C* a, C* b, long long pre, int sgn, double scale, int cat) {
    long long L = ctx->g.L, n = pre * L * L * L * L;
    // two in-place tile transforms: axes (1, 2) for every (3, 4), then axes (3, 4)
    if (dft2_tile(ctx, a, a, pre, (int)L, L * L, sgn, 1.0, ctx->twL, cat)) {
        if (dft2_tile(ctx, a, a, pre * L * L, (int)L, 1, sgn, scale, ctx->twL, cat)) { CK(cudaGetLastError()); return 0; }
        FAIL("dft4: second tile transform failed to launch");
    }
    LAUNCH(cat, dft_axis_kernel, nblk(n, 128), 128, a, b, pre, (int)L, L * L * L, sgn, 1.0, ctx->twL);
    LAUNCH(cat, dft_axis_kernel, nblk(n, 128), 128, b, a, pre * L, (int)L, L * L, sgn, 1.0, ctx->twL);
    LAUNCH(cat, dft_axis_kernel, nblk(n, 128), 128, a, b, pre * L * L, (int)L, L, sgn, 1.0, ctx->twL);
    LAUNCH(cat, dft_axis_kernel, nblk(n, 128), 128, b, a, pre * L * L * L, (int)L, 1LL, sgn, scale, ctx->twL);
    CK(cudaGetLastError()); return 0;
}

// group this rank's class representatives of a K2-shaped symmetry group into columns (W, P, k) of <= FDGA_NV reps
static int build_columns(fdga_ctx* ctx, SymGroup& s) {
    cudaFree(s.d_col_iW); cudaFree(s.d_col_iP); cudaFree(s.d_col_ik); cudaFree(s.d_col_start); cudaFree(s.d_rep_inu); cudaFree(s.d_rep_cls); cudaFree(s.d_grp_start); cudaFree(s.d_reps);
    s.d_col_iW = s.d_col_iP = s.d_col_ik = s.d_col_start = s.d_rep_inu = s.d_rep_cls = s.d_grp_start = nullptr; s.ncol = 0; s.ngrp = 0; s.d_reps = nullptr; s.nrep = 0;
    long long c0 = (long long)ctx->rank * s.chunk, c1 = c0 + s.chunk;
    if (c0 > s.ncls) c0 = s.ncls; if (c1 > s.ncls) c1 = s.ncls;
    const int nB2 = 2 * ctx->g.nK2b - 1, nF2 = 2 * ctx->g.nK2f, NP = ctx->g.NP;
    struct Rep { long long key; int inu; int cls; };
    std::vector<Rep> reps; reps.reserve(c1 - c0);
    for (long long c = c0; c < c1; c++) {
        long long idx = s.h_index[s.h_offsets[c]];
        int iW = idx % nB2; idx /= nB2; int inu = idx % nF2; idx /= nF2; int iP = idx % NP; int ik = (int)(idx / NP);
        Rep r; r.key = ((long long)ik * NP + iP) * nB2 + iW; r.inu = inu; r.cls = (int)c;
        reps.push_back(r);
    }
    {   // q-lane kernel: one warp per representative, slab-major
        std::vector<int4> list; list.reserve(reps.size());
        for (auto& r : reps) {
            long long key = r.key; const int iW = (int)(key % nB2); key /= nB2; const int iP = (int)(key % NP), ik = (int)(key / NP);
            const int L = ctx->g.L;
            list.push_back(make_int4(iW | (r.inu << 16), iP | (ik << 16), (iP % L) | ((iP / L) << 8) | ((ik % L) << 16) | ((ik / L) << 24), r.cls));
        }
        std::stable_sort(list.begin(), list.end(), [&](const int4& a, const int4& b) {
            // (P, W, nu, k): the warps of a CTA share their R rows and their K3 rows (both independent of the column momentum k)
            static const int order = getenv("FDGA_QL_ORDER") ? atoi(getenv("FDGA_QL_ORDER")) : 1;
            auto key = [&](const int4& t) {
                const long long iP = t.y & 0xffff, iW = t.x & 0xffff, ik = (t.y >> 16) & 0xffff, inu = t.x >> 16;
                return order == 0 ? ((iP * nB2 + iW) * NP + ik) * nF2 + inu : ((iP * nB2 + iW) * nF2 + inu) * NP + ik; };
            return key(a) < key(b); });
        s.nrep = (int)list.size();
        if (s.nrep) { CK(cudaMalloc(&s.d_reps, list.size() * sizeof(int4))); CK(cudaMemcpy(s.d_reps, list.data(), list.size() * sizeof(int4), cudaMemcpyHostToDevice)); }
    }
    std::stable_sort(reps.begin(), reps.end(), [](const Rep& a, const Rep& b) { return a.key < b.key; });
    std::vector<int> ciW, ciP, cik, cstart, rinu, rcls;
    for (size_t i = 0; i < reps.size();) {
        size_t j = i;
        while (j < reps.size() && reps[j].key == reps[i].key && j - i < FDGA_NV) j++;
        long long key = reps[i].key;
        ciW.push_back((int)(key % nB2)); key /= nB2; ciP.push_back((int)(key % NP)); cik.push_back((int)(key / NP));
        cstart.push_back((int)i);
        i = j;
    }
    cstart.push_back((int)reps.size());
    for (auto& r : reps) { rinu.push_back(r.inu); rcls.push_back(r.cls); }
    s.ncol = (int)ciW.size();
    if (s.ncol == 0) return 0;
    // groups of <= FDGA_WGROUP consecutive columns with the same (P, k) (the sort key makes them adjacent)
    std::vector<int> gstart;
    for (int c = 0; c < s.ncol;) {
        int e = c;
        while (e < s.ncol && e - c < FDGA_WGROUP && ciP[e] == ciP[c] && cik[e] == cik[c]) e++;
        gstart.push_back(c); c = e;
    }
    gstart.push_back(s.ncol);
    s.ngrp = (int)gstart.size() - 1;
    auto up = [&](int*& d, const std::vector<int>& h) -> int {
        CK(cudaMalloc(&d, h.size() * sizeof(int))); CK(cudaMemcpy(d, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice)); return 0; };
    if (up(s.d_col_iW, ciW) || up(s.d_col_iP, ciP) || up(s.d_col_ik, cik) || up(s.d_col_start, cstart) || up(s.d_rep_inu, rinu) || up(s.d_rep_cls, rcls) || up(s.d_grp_start, gstart)) return 1;
    return 0;
}
static ColDev col_dev(const SymGroup& s) {
    ColDev c; c.ncol = s.ncol; c.iW = s.d_col_iW; c.iP = s.d_col_iP; c.ik = s.d_col_ik; c.start = s.d_col_start; c.rep_inu = s.d_rep_inu; c.rep_cls = s.d_rep_cls; c.ngrp = s.ngrp; c.grp_start = s.d_grp_start;
    return c;
}

// (W, P) slabs of the hoisted right factor that this rank actually reads: those of its K1 representatives and of its
// K2 columns.  Everything else of Rt stays unwritten (and unread).
static int ensure_slabs(fdga_ctx* ctx) {
    if (!ctx->slabs_dirty) return 0;
    if (ctx->swave) { ctx->slabs_dirty = false; return 0; }
    if (ctx->capturing) FAIL("fdga_graph: the slab tables must be current before a recording starts (run the sequence once first)");
    ctx->epoch++;
    const Grid& g = ctx->g;
    const int nB1 = 2 * g.nK1 - 1, nB2 = 2 * g.nK2b - 1, nF2 = 2 * g.nK2f, NP = g.NP;
    for (int kind = 0; kind < 4; kind++) {
        const bool pp = (kind % 2 == 0), bubble_mesh = kind < 2;
        const int nBo = bubble_mesh ? 2 * g.nPiB - 1 : nB2;
        std::vector<unsigned char> mark((size_t)nBo * NP, 0);
        std::vector<unsigned long long> numask((size_t)nBo * NP, 0ULL);   // nu values of this rank's representatives per slab
        const SymGroup& s2 = ctx->sg[pp ? FDGA_SG_PP2 : FDGA_SG_PH2];
        if (s2.set) {
            long long c0 = std::min((long long)ctx->rank * s2.chunk, s2.ncls), c1 = std::min(c0 + s2.chunk, s2.ncls);
            for (long long c = c0; c < c1; c++) {
                long long idx = s2.h_index[s2.h_offsets[c]];
                int iW = idx % nB2; idx /= nB2; int inu = idx % nF2; idx /= nF2; int iP = idx % NP;
                int iWo = bubble_mesh ? posB(iW - (g.nK2b - 1), g.nPiB) : iW;
                mark[iWo + (size_t)nBo * iP] = 1;
                numask[iWo + (size_t)nBo * iP] |= (nF2 <= 64) ? (1ULL << inu) : ~0ULL;
            }
        }
        const SymGroup& s1 = ctx->sg[FDGA_SG_K1];
        if (bubble_mesh && s1.set) {
            long long c0 = std::min((long long)ctx->rank * s1.chunk, s1.ncls), c1 = std::min(c0 + s1.chunk, s1.ncls);
            for (long long c = c0; c < c1; c++) {
                long long idx = s1.h_index[s1.h_offsets[c]];
                int iW = idx % nB1, iP = (int)(idx / nB1);
                mark[posB(iW - (g.nK1 - 1), g.nPiB) + (size_t)nBo * iP] = 1;
            }
        }
        std::vector<int4> list;     // (iW, iP, nu mask low, nu mask high)
        for (int iP = 0; iP < NP; iP++) for (int iW = 0; iW < nBo; iW++) if (mark[iW + (size_t)nBo * iP]) {
            unsigned long long m = numask[iW + (size_t)nBo * iP];
            list.push_back(make_int4(iW, iP, (int)(unsigned)(m & 0xffffffffULL), (int)(unsigned)(m >> 32)));
        }
        cudaFree(ctx->d_slabs[kind]); ctx->d_slabs[kind] = nullptr; ctx->n_slabs[kind] = (int)list.size();
        cudaFree(ctx->d_slabmap[kind]); ctx->d_slabmap[kind] = nullptr;
        std::vector<int> map((size_t)nBo * NP, -1);
        for (size_t i = 0; i < list.size(); i++) map[list[i].x + (size_t)nBo * list[i].y] = (int)i;
        CK(cudaMalloc(&ctx->d_slabmap[kind], map.size() * sizeof(int)));
        CK(cudaMemcpy(ctx->d_slabmap[kind], map.data(), map.size() * sizeof(int), cudaMemcpyHostToDevice));
        if (!list.empty()) {
            CK(cudaMalloc(&ctx->d_slabs[kind], list.size() * sizeof(int4)));
            CK(cudaMemcpy(ctx->d_slabs[kind], list.data(), list.size() * sizeof(int4), cudaMemcpyHostToDevice));
        }
    }
    {   // compact storage of everything slab-shaped: the bubbles' slabs, the per-channel right factors, the lane right factors
        CK(cudaStreamSynchronize(ctx->stream));
        for (int i = 1; i < 3; i++) CK(cudaStreamSynchronize(ctx->lane[i]));
        const size_t slab = (size_t)(2 * g.nPiF) * NP;
        for (int i = 0; i < 4; i++) {
            cudaFree(ctx->PiT[i]); ctx->PiT[i] = nullptr;
            const size_t n = slab * ctx->n_slabs[i % 2 == 0 ? 0 : 1];
            if (n) CK(cudaMalloc(&ctx->PiT[i], n * sizeof(C)));
            ctx->pi_dirty[i] = true;
        }
        for (int ch = 0; ch < 3; ch++) {
            cudaFree(ctx->Rt3[ch]); ctx->Rt3[ch] = nullptr;
            const size_t n = slab * ctx->n_slabs[ch == FDGA_PCH ? 0 : 1];
            if (n) CK(cudaMalloc(&ctx->Rt3[ch], n * sizeof(C)));
        }
        ctx->lenRtL = (size_t)2 * std::max(g.nPiF, g.nK2f) * NP * std::max(ctx->n_slabs[2], ctx->n_slabs[3]);
        for (int i = 0; i < 3; i++) {
            cudaFree(ctx->RtL[i]); ctx->RtL[i] = nullptr;
            if (ctx->lenRtL) CK(cudaMalloc(&ctx->RtL[i], ctx->lenRtL * sizeof(C)));
        }
        ctx->Rt = ctx->RtL[0];
    }
    ctx->slabs_dirty = false;
    invalidate_rt(ctx);
    return 0;
}

// ================================================================================================
extern "C" {

const char* fdga_last_error(fdga_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int fdga_create(const fdga_dims* dims, int device, fdga_ctx** out) {
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { g_create_error = std::string("no CUDA device available (libfdga has no CPU fallback): ") + cudaGetErrorString(e); return 2; }
    if (device < 0 || device >= ndev) { g_create_error = "invalid device index"; return 2; }
    const bool swave = dims->lev[0].type == FDGA_LV_NL;
    fdga_dims dims_base = *dims;        // MBE level types -> their base types for the structural checks below
    bool any_mbe = false;
    for (int l = 0; l < dims->nlev && l < FDGA_MAX_LEVELS; l++) {
        int& t = dims_base.lev[l].type;
        if (t == FDGA_LV_NL2_MBE) { t = FDGA_LV_NL2; any_mbe = true; } else if (t == FDGA_LV_LOCAL_MBE) { t = FDGA_LV_LOCAL; any_mbe = true; }
    }
    for (int l = 1; l < dims->nlev; l++) {
        const bool m = dims->lev[l].type == FDGA_LV_NL2_MBE || dims->lev[l].type == FDGA_LV_LOCAL_MBE;
        const bool mprev = dims->lev[l - 1].type == FDGA_LV_NL2_MBE || dims->lev[l - 1].type == FDGA_LV_LOCAL_MBE;
        if (m && !mprev) { g_create_error = "dims: MBE levels must form the head of the chain (an MBE vertex below an asymptotic one is not supported)"; return 2; }
    }
    const fdga_dims* dims_orig = dims; dims = &dims_base;
    if (dims->nlev < 2 || dims->nlev > FDGA_MAX_LEVELS || (dims->lev[0].type != FDGA_LV_NL2 && !swave) || dims->lev[dims->nlev - 1].type != FDGA_LV_CORE) {
        g_create_error = "dims: need lev[0] = NL2 (or NL for the s-wave solver) and lev[nlev-1] = CORE, 2 <= nlev <= FDGA_MAX_LEVELS"; return 2; }
    for (int l = 1; l < dims->nlev - 1; l++) if (dims->lev[l].type == FDGA_LV_CORE) { g_create_error = "dims: CORE level must be last"; return 2; }
    for (int l = 1; l < dims->nlev - 1; l++) {
        const int t = dims->lev[l].type;
        if (t != FDGA_LV_NL2 && t != FDGA_LV_NL && t != FDGA_LV_LOCAL) { g_create_error = "dims: unknown level type"; return 2; }
        if (t != FDGA_LV_LOCAL && t != dims->lev[0].type) { g_create_error = "dims: the momentum-dependent levels of a chain must all be NL2 (NL2 solver) or all NL (s-wave solver)"; return 2; }
        if (t != FDGA_LV_LOCAL && dims->lev[l - 1].type == FDGA_LV_LOCAL) { g_create_error = "dims: NL2 / NL levels must precede LOCAL levels"; return 2; }
    }
    const fdga_level_desc& d0 = dims->lev[0];
    if (dims->nPiB != d0.nK1) { g_create_error = "dims: bubble bosonic mesh must equal the K1 mesh (nPiB == nK1)"; return 2; }
    if (!(d0.nK1 > d0.nK2[0] && d0.nK1 > d0.nK2[1] && d0.nK2[0] >= d0.nK3[0] && d0.nK2[1] >= d0.nK3[1] && dims->nPiF >= d0.nK2[1])) {
        g_create_error = "dims: mesh constraints violated (src/nonlocal_2/channel.jl:26-33)"; return 2; }
    if (swave && dims->LG < dims->nq) { g_create_error = "dims: the s-wave solver needs LG >= nq (Green-function mesh at least as fine as the vertex mesh)"; return 2; }
    fdga_ctx* ctx = new fdga_ctx();
    ctx->capturing = false; ctx->epoch = 0; ctx->mbe = any_mbe;
    ctx->dims = *dims; ctx->device = device; ctx->nlev = dims->nlev; ctx->swave = swave; ctx->swScratch[0] = ctx->swScratch[1] = nullptr;
    ctx->nranks = 1; ctx->rank = 0; ctx->comm = nullptr; memset(&ctx->nccl, 0, sizeof(ctx->nccl));
    ctx->profile = false; ctx->cur_cat = -1; ctx->total_launches = 0; ctx->opt_sde_own_gamma = 0; ctx->opt_generic = 0; ctx->opt_hartree_once = 0; ctx->opt_local = 0; ctx->opt_direct_k1 = 0; ctx->opt_qlane = getenv("FDGA_QLANE") ? atoi(getenv("FDGA_QLANE")) : -1; ctx->defer = false;
    memset(ctx->t_ms, 0, sizeof(ctx->t_ms)); memset(ctx->n_launch, 0, sizeof(ctx->n_launch));
    Grid& g = ctx->g;
    g.T = dims->T; g.L = dims->nq; g.NP = dims->nq * dims->nq; g.nPiB = dims->nPiB; g.nPiF = dims->nPiF;
    g.nK1 = d0.nK1; g.nK2b = d0.nK2[0]; g.nK2f = d0.nK2[1]; g.nK3b = d0.nK3[0]; g.nK3f = d0.nK3[1]; g.LG = dims->LG; g.nG = dims->nG;
#define CKC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { g_create_error = std::string(#call) + ": " + cudaGetErrorString(e_); delete ctx; return 1; } } while (0)
    CKC(cudaSetDevice(device));
    CKC(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->main_stream = ctx->stream; ctx->lane[0] = ctx->stream; ctx->cur_lane = 0; ctx->forked = false;
    { const char* e = getenv("FDGA_SERIAL"); ctx->opt_serial = e ? atoi(e) : 0; }     // debugging aid, same as FDGA_OPT_SERIAL
    {   // lane 1 carries the heaviest jobs (t channel, ph bubble): highest priority so that its CTAs are placed first
        int lo = 0, hi = 0; cudaDeviceGetStreamPriorityRange(&lo, &hi);
        CKC(cudaStreamCreateWithPriority(&ctx->lane[1], cudaStreamNonBlocking, hi));
        CKC(cudaStreamCreateWithFlags(&ctx->lane[2], cudaStreamNonBlocking));
    }
    CKC(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)); ctx->copy_pending = false;
    CKC(cudaEventCreateWithFlags(&ctx->ev_copy_ready, cudaEventDisableTiming)); CKC(cudaEventCreateWithFlags(&ctx->ev_copy_done, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&ctx->ev_cache, cudaEventDisableTiming)); ctx->cache_on_lane = false;
    for (int i = 0; i < 3; i++) CKC(cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming));
    for (int l = 0; l < ctx->nlev; l++) if (alloc_level(ctx, ctx->lev[l], dims_orig->lev[l])) { g_create_error = ctx->err; delete ctx; return 1; }
    if (any_mbe) ctx->opt_generic = 1;      // nonlinear evaluator: the piece-wise kernels (right-factor hoisting aside) do not apply
    fdga_level_desc dz = d0;
    if (alloc_level(ctx, ctx->FL, dz) || alloc_level(ctx, ctx->Fbuff, dz)) { g_create_error = ctx->err; delete ctx; return 1; }
    ctx->lenG = (size_t)2 * g.nG * g.LG * g.LG;
    for (int i = 0; i < 5; i++) { CKC(cudaMalloc(&ctx->G[i], ctx->lenG * sizeof(C))); CKC(cudaMemsetAsync(ctx->G[i], 0, ctx->lenG * sizeof(C), ctx->stream)); }
    ctx->lenPi = (size_t)(2 * g.nPiB - 1) * (2 * g.nPiF) * g.NP * g.NP;
    ctx->lenPisw = (size_t)(2 * g.nPiB - 1) * (2 * g.nPiF) * g.NP;
    if (swave) ctx->lenPi = ctx->lenPisw;      // Pi[W, w, P] (NL_MF_Pi, src/types.jl)
    for (int i = 0; i < 4; i++) {      // Pi[i] (full layout) and PiT[i] (compact slabs) are allocated on demand / by ensure_slabs
        ctx->Pi[i] = nullptr; ctx->PiT[i] = nullptr; ctx->pi_src[i] = PI_NONE; ctx->pi_full_valid[i] = false;
        CKC(cudaMalloc(&ctx->Pisw[i], ctx->lenPisw * sizeof(C)));
        if (swave) CKC(cudaMemsetAsync(ctx->Pisw[i], 0, ctx->lenPisw * sizeof(C), ctx->stream));
        ctx->pi_dirty[i] = true;
    }
    ctx->lenK3 = ctx->lev[0].len[2];
    for (int i = 0; i < 10; i++) { CKC(cudaMalloc(&ctx->cache[i], ctx->lenK3 * sizeof(C))); CKC(cudaMemsetAsync(ctx->cache[i], 0, ctx->lenK3 * sizeof(C), ctx->stream)); }
    for (int i = 0; i < 2; i++) { CKC(cudaMalloc(&ctx->L[i], ctx->lev[0].len[1] * sizeof(C))); CKC(cudaMemsetAsync(ctx->L[i], 0, ctx->lev[0].len[1] * sizeof(C), ctx->stream)); }
    ctx->Rt = nullptr; for (int i = 0; i < 3; i++) ctx->RtL[i] = nullptr;      // right factors: compact, sized by ensure_slabs
    ctx->lenRtL = 0; ctx->scratchBig = nullptr;
    CKC(cudaMalloc(&ctx->SigR2, ctx->lenG * sizeof(C)));
    ctx->lenScratch = std::max(ctx->lev[0].len[1], ctx->lenG);      // ping-pong of the 4-d transforms of the (K2-shaped) L arrays
    CKC(cudaMalloc(&ctx->scratchA, ctx->lenScratch * sizeof(C))); CKC(cudaMalloc(&ctx->scratchB, ctx->lenScratch * sizeof(C)));
    CKC(cudaMalloc(&ctx->GR, ctx->lenG * sizeof(C))); CKC(cudaMalloc(&ctx->GRm, ctx->lenG * sizeof(C)));
    for (int i = 0; i < 2; i++) CKC(cudaMalloc(&ctx->Ghat[i], (size_t)2 * g.nG * g.NP * sizeof(C)));
    CKC(cudaMalloc(&ctx->SigR, ctx->lenG * sizeof(C))); CKC(cudaMalloc(&ctx->SigTmp, ctx->lenG * sizeof(C))); CKC(cudaMalloc(&ctx->SigAcc, ctx->lenG * sizeof(C)));
    ctx->lenFlat = 3 * (ctx->lev[0].len[0] + ctx->lev[0].len[1] + ctx->lev[0].len[2]);
    CKC(cudaMalloc(&ctx->flat, ctx->lenFlat * sizeof(C))); CKC(cudaMalloc(&ctx->flat2, ctx->lenFlat * sizeof(C))); CKC(cudaMalloc(&ctx->stash, ctx->lenFlat * sizeof(C)));
    CKC(cudaMalloc(&ctx->d_occ, sizeof(double)));
    for (int i = 0; i < 3; i++) CKC(cudaMalloc(&ctx->ConvTabL[i], ctx->lev[0].len[1] * sizeof(C)));
    {   // merged level (S.F + S.F0) for the K2 left factor
        const fdga_level_desc& a = dims->lev[0]; const fdga_level_desc& b = dims->lev[1];
        ctx->has_fsum = !any_mbe && dims->nlev >= 3 && b.type == FDGA_LV_NL2 && a.nK1 == b.nK1 && a.nK2[0] == b.nK2[0] && a.nK2[1] == b.nK2[1] && a.nK3[0] == b.nK3[0] && a.nK3[1] == b.nK3[1];
        ctx->fsum_dirty = true; memset(&ctx->Fsum, 0, sizeof(ctx->Fsum));
        if (ctx->has_fsum && alloc_level(ctx, ctx->Fsum, d0)) { g_create_error = ctx->err; delete ctx; return 1; }
    }
    ctx->n_nl2 = 0; while (ctx->n_nl2 < ctx->nlev && dims->lev[ctx->n_nl2].type == FDGA_LV_NL2) ctx->n_nl2++;
    {   // momentum layouts read by any q-lane job (fdga_qlane.cuh); K1 copies are tiny and always kept (FDGA_OPT_DIRECT_K1)
        unsigned* m = ctx->mom_mask; m[0] = m[1] = m[2] = 0;
        qlane_needed_layouts<JOB_K2, CH_P>(m, true); qlane_needed_layouts<JOB_K2, CH_T>(m, true); qlane_needed_layouts<JOB_K2, CH_A>(m, true);
        qlane_needed_layouts<JOB_K2_MF, CH_P>(m, true); qlane_needed_layouts<JOB_K2_MF, CH_T>(m, true); qlane_needed_layouts<JOB_K2_MF, CH_A>(m, true);
        qlane_needed_layouts<JOB_LK2, CH_P>(m, true); qlane_needed_layouts<JOB_LK2, CH_T>(m, true); qlane_needed_layouts<JOB_LK2, CH_A>(m, true);
        qlane_needed_layouts<JOB_SDE_PP, CH_P>(m, true); qlane_needed_layouts<JOB_SDE_PH, CH_A>(m, true);
        if (qlane_enabled(ctx) && alloc_mom_all(ctx)) { g_create_error = ctx->err; delete ctx; return 1; }
    }
    CKC(cudaMalloc(&ctx->twL, g.L * sizeof(C))); CKC(cudaMalloc(&ctx->twLG, g.LG * sizeof(C)));
    twiddle_kernel<<<nblk(g.L, 64), 64, 0, ctx->stream>>>(ctx->twL, g.L);
    twiddle_kernel<<<nblk(g.LG, 64), 64, 0, ctx->stream>>>(ctx->twLG, g.LG);
    for (int i = 0; i < 3; i++) { ctx->Rt3[i] = nullptr; ctx->rt_kind[i] = -1; }
    for (int i = 0; i < 4; i++) { ctx->d_slabs[i] = nullptr; ctx->d_slabmap[i] = nullptr; ctx->n_slabs[i] = 0; } ctx->slabs_dirty = true;
    for (int i = 0; i < 3; i++) {
        CKC(cudaMalloc(&ctx->OwnTabL[i], (size_t)(2 * g.nK2f) * (2 * g.nK2b - 1) * g.NP * sizeof(C))); CKC(cudaMalloc(&ctx->RtotL[i], (size_t)(2 * g.nK2b - 1) * g.NP * sizeof(C)));
        CKC(cudaMalloc(&ctx->TtabL[i], (size_t)(2 * g.nPiF) * (2 * g.nK2f) * (2 * g.nK2b - 1) * sizeof(C)));
    }
    for (int i = 0; i < FDGA_SG_COUNT; i++) { SymGroup& s = ctx->sg[i]; s.set = false; s.d_offsets = s.d_index = nullptr; s.d_ops = nullptr; s.d_member_class = nullptr; s.d_repvals = nullptr; s.d_rep[0] = s.d_rep[1] = s.d_rep[2] = nullptr; s.ncol = 0; s.ngrp = 0; s.nrep = 0; s.d_reps = nullptr; s.d_col_iW = s.d_col_iP = s.d_col_ik = s.d_col_start = s.d_rep_inu = s.d_rep_cls = s.d_grp_start = nullptr; }
    CKC(cudaStreamSynchronize(ctx->stream));
    *out = ctx;
    return 0;
}

int fdga_destroy(fdga_ctx* ctx) {
    if (!ctx) return 0;
    auto_forget(ctx);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 1; i < 3; i++) cudaStreamSynchronize(ctx->lane[i]);
    // graphs that hold NCCL nodes must go before their communicator
    for (auto& r : ctx->graphs) if (r.live) { cudaGraphExecDestroy(r.exec); cudaGraphDestroy(r.graph); r.live = false; }
    if (ctx->comm && ctx->nccl.CommDestroy) ctx->nccl.CommDestroy(ctx->comm);
    for (int l = 0; l < ctx->nlev; l++) free_level(ctx->lev[l]);
    free_level(ctx->FL); free_level(ctx->Fbuff); if (ctx->has_fsum) free_level(ctx->Fsum);
    for (int i = 0; i < 5; i++) cudaFree(ctx->G[i]);
    for (int i = 0; i < 4; i++) { cudaFree(ctx->Pi[i]); cudaFree(ctx->PiT[i]); cudaFree(ctx->Pisw[i]); }
    for (int i = 0; i < 10; i++) cudaFree(ctx->cache[i]);
    cudaFree(ctx->L[0]); cudaFree(ctx->L[1]); cudaFree(ctx->scratchA); cudaFree(ctx->scratchB); cudaFree(ctx->scratchBig);
    cudaFree(ctx->Ghat[0]); cudaFree(ctx->Ghat[1]); cudaFree(ctx->GR); cudaFree(ctx->GRm); cudaFree(ctx->SigR); cudaFree(ctx->SigTmp); cudaFree(ctx->SigAcc);
    cudaFree(ctx->swScratch[0]); cudaFree(ctx->swScratch[1]);
    cudaFree(ctx->PiMixed[0]); cudaFree(ctx->PiMixed[1]); cudaFree(ctx->itpA); cudaFree(ctx->itpB);
    cudaFree(ctx->kryV); cudaFree(ctx->kryP); cudaFree(ctx->kryW); cudaFree(ctx->kryX); cudaFree(ctx->kryH); cudaFree(ctx->kryPart); cudaFree(ctx->kryTicket); if (ctx->kryHhost) cudaFreeHost(ctx->kryHhost);
    cudaFree(ctx->flat); cudaFree(ctx->flat2); cudaFree(ctx->stash); cudaFree(ctx->d_occ); for (int i = 0; i < 3; i++) { cudaFree(ctx->TtabL[i]); cudaFree(ctx->OwnTabL[i]); cudaFree(ctx->RtotL[i]); cudaFree(ctx->ConvTabL[i]); cudaFree(ctx->RtL[i]); } cudaFree(ctx->SigR2); cudaFree(ctx->twL); cudaFree(ctx->twLG); for (int i = 0; i < 3; i++) cudaFree(ctx->Rt3[i]); for (int i = 0; i < 4; i++) { cudaFree(ctx->d_slabs[i]); cudaFree(ctx->d_slabmap[i]); }
    for (int i = 0; i < FDGA_SG_COUNT; i++) { SymGroup& s = ctx->sg[i]; cudaFree(s.d_offsets); cudaFree(s.d_index); cudaFree(s.d_ops); cudaFree(s.d_member_class); for (int k = 0; k < 3; k++) cudaFree(s.d_rep[k]);
        cudaFree(s.d_col_iW); cudaFree(s.d_col_iP); cudaFree(s.d_col_ik); cudaFree(s.d_col_start); cudaFree(s.d_rep_inu); cudaFree(s.d_rep_cls); cudaFree(s.d_grp_start); cudaFree(s.d_reps); }
    for (auto& r : ctx->graphs) if (r.live) { cudaGraphExecDestroy(r.exec); cudaGraphDestroy(r.graph); }
    for (auto& ev : ctx->events) { cudaEventDestroy(ev.a); cudaEventDestroy(ev.b); }
    cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); cudaEventDestroy(ctx->ev_copy_ready); cudaEventDestroy(ctx->ev_copy_done);
    for (int i = 1; i < 3; i++) cudaStreamDestroy(ctx->lane[i]);
    cudaEventDestroy(ctx->ev_fork); cudaEventDestroy(ctx->ev_cache); for (int i = 0; i < 3; i++) cudaEventDestroy(ctx->ev_join[i]);
    cudaStreamDestroy(ctx->main_stream);
    delete ctx;
    return 0;
}

int fdga_set_option(fdga_ctx* ctx, int opt, int value) {
    ctx->epoch++;
    if (opt == FDGA_OPT_SDE_OWN_GAMMA) { ctx->opt_sde_own_gamma = value != 0; return 0; }
    if (opt == FDGA_OPT_GENERIC_KERNELS) {
        if (ctx->mbe && !value) FAIL("FDGA_OPT_GENERIC_KERNELS: contexts with MBE vertices only have the generic kernels");
        ctx->opt_generic = value != 0; return 0;
    }
    if (opt == FDGA_OPT_FD_HARTREE_ONCE) { ctx->opt_hartree_once = value != 0; return 0; }
    if (opt == FDGA_OPT_LOCAL_SOLVER) {
        if (value && (ctx->g.L != 1 || ctx->g.LG != 1)) FAIL("FDGA_OPT_LOCAL_SOLVER needs nq = LG = 1");
        if (value && ctx->swave) FAIL("FDGA_OPT_LOCAL_SOLVER: not for an s-wave (NL) context");
        ctx->opt_local = value != 0; invalidate_rt(ctx); return 0;
    }
    if (opt == FDGA_OPT_DIRECT_K1) { ctx->opt_direct_k1 = value != 0; return 0; }
    if (opt == FDGA_OPT_SERIAL) { ctx->opt_serial = value; return 0; }
    if (opt == FDGA_OPT_QLANE) {
        if (value < -1 || value > 1) FAIL("fdga_set_option: FDGA_OPT_QLANE takes -1, 0 or 1");
        ctx->opt_qlane = value;
        CK(cudaSetDevice(ctx->device));
        if (qlane_enabled(ctx) && alloc_mom_all(ctx)) return 1;
        return 0;
    }
    FAIL("fdga_set_option: unknown option");
}
int fdga_sync(fdga_ctx* ctx) {
    if (ctx->capturing) FAIL("fdga_sync: not inside a graph recording");
    CK(cudaSetDevice(ctx->device)); CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->copy_pending) { CK(cudaStreamSynchronize(ctx->copy_stream)); ctx->copy_pending = false; }
    return 0;
}
void* fdga_stream(fdga_ctx* ctx) { return (void*)ctx->stream; }

// ---- NCCL ----------------------------------------------------------------------------------------
static int load_nccl(NcclApi& a, std::string& err) {
    if (a.h) return 0;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { a.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (a.h) break; }
    if (!a.h) { err = std::string("dlopen libnccl failed: ") + dlerror(); return 1; }
    a.GetUniqueId = (int (*)(void*))dlsym(a.h, "ncclGetUniqueId");
    a.CommInitRank = dlsym(a.h, "ncclCommInitRank");
    a.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(a.h, "ncclAllGather");
    a.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(a.h, "ncclAllReduce");
    a.Broadcast = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(a.h, "ncclBroadcast");
    a.CommDestroy = (int (*)(void*))dlsym(a.h, "ncclCommDestroy");
    a.GroupStart = (int (*)())dlsym(a.h, "ncclGroupStart"); a.GroupEnd = (int (*)())dlsym(a.h, "ncclGroupEnd");
    a.GetErrorString = (const char* (*)(int))dlsym(a.h, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.AllGather || !a.CommDestroy || !a.GetErrorString || !a.GroupStart || !a.GroupEnd) { err = "libnccl: missing symbols"; return 1; }
    return 0;
}
int fdga_comm_unique_id(void* unique_id_128B) {
    NcclApi a; memset(&a, 0, sizeof(a)); std::string err;
    if (load_nccl(a, err)) { g_create_error = err; return 1; }
    int rc = a.GetUniqueId(unique_id_128B);
    if (rc) { g_create_error = std::string("ncclGetUniqueId: ") + a.GetErrorString(rc); return 1; }
    return 0;
}
static int rebuild_sg(fdga_ctx* ctx, int which);
int fdga_comm_init(fdga_ctx* ctx, int nranks, int rank, const void* unique_id_128B) {
    if (nranks < 1 || rank < 0 || rank >= nranks) FAIL("fdga_comm_init: bad nranks/rank");
    CK(cudaSetDevice(ctx->device));
    if (nranks > 1) {
        if (load_nccl(ctx->nccl, ctx->err)) return 1;
        typedef int (*init_t)(void**, int, UniqueId, int);
        init_t init = (init_t)ctx->nccl.CommInitRank;
        UniqueId id; memcpy(id.b, unique_id_128B, 128);
        int rc = init(&ctx->comm, nranks, id, rank);
        if (rc) FAIL(std::string("ncclCommInitRank: ") + ctx->nccl.GetErrorString(rc));
    }
    ctx->nranks = nranks; ctx->rank = rank; ctx->slabs_dirty = true;
    // re-order / re-chunk already registered symmetry groups for the new rank count
    for (int i = 0; i < FDGA_SG_COUNT; i++) if (ctx->sg[i].set && rebuild_sg(ctx, i)) return 1;
    return 0;
}

// ---- data movement ---------------------------------------------------------------------------------
static LevelBuf* which_level(fdga_ctx* ctx, int which) {
    if (which == FDGA_V_FL) return &ctx->FL;
    if (which == FDGA_V_FBUFF) return &ctx->Fbuff;
    if (which >= 0 && which < ctx->nlev) return &ctx->lev[which];
    return nullptr;
}
int fdga_set_vertex(fdga_ctx* ctx, int which, int channel, int cls, const fdga_c64* host, int64_t n) {
    CK(cudaSetDevice(ctx->device));
    LevelBuf* lb = which_level(ctx, which);
    if (!lb || lb->d.type == FDGA_LV_CORE || channel < 0 || channel > 2 || cls < 0 || cls > 2) FAIL("fdga_set_vertex: bad selector");
    if ((size_t)n != lb->len[cls]) FAIL("fdga_set_vertex: length mismatch");
    if (lb == &ctx->lev[0] && ctx->copy_pending) CK(cudaStreamWaitEvent(ctx->main_stream, ctx->ev_copy_done, 0));
    CK(cudaMemcpyAsync(lb->K[channel][cls], host, n * sizeof(C), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    lb->sw_dirty = true; lb->k1h_dirty = true; lb->mom_valid[0] = lb->mom_valid[1] = lb->mom_valid[2] = 0; ctx->fsum_dirty = true;
    invalidate_rt(ctx);
    return 0;
}
int fdga_get_vertex(fdga_ctx* ctx, int which, int channel, int cls, fdga_c64* host, int64_t n) {
    CK(cudaSetDevice(ctx->device));
    LevelBuf* lb = which_level(ctx, which);
    if (!lb || lb->d.type == FDGA_LV_CORE || channel < 0 || channel > 2 || cls < 0 || cls > 2) FAIL("fdga_get_vertex: bad selector");
    if ((size_t)n != lb->len[cls]) FAIL("fdga_get_vertex: length mismatch");
    CK(cudaMemcpyAsync(host, lb->K[channel][cls], n * sizeof(C), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int fdga_set_core(fdga_ctx* ctx, int level, int which4, const fdga_c64* host, int64_t n) {
    CK(cudaSetDevice(ctx->device));
    if (level < 0 || level >= ctx->nlev || ctx->lev[level].d.type != FDGA_LV_CORE || which4 < 0 || which4 > 3) FAIL("fdga_set_core: bad selector");
    if ((size_t)n != ctx->lev[level].corelen) FAIL("fdga_set_core: length mismatch");
    CK(cudaMemcpyAsync(ctx->lev[level].core[which4], host, n * sizeof(C), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    invalidate_rt(ctx);
    return 0;
}
#define SETGET(NAME, ARR, COUNT, LEN, DIRTY) \
int fdga_set_##NAME(fdga_ctx* ctx, int which, const fdga_c64* host, int64_t n) { \
    CK(cudaSetDevice(ctx->device)); \
    if (which < 0 || which >= COUNT) FAIL("fdga_set_" #NAME ": bad selector"); \
    if ((size_t)n != (LEN)) FAIL("fdga_set_" #NAME ": length mismatch"); \
    CK(cudaMemcpyAsync(ctx->ARR[which], host, n * sizeof(C), cudaMemcpyHostToDevice, ctx->stream)); \
    CK(cudaStreamSynchronize(ctx->stream)); DIRTY; return 0; } \
int fdga_get_##NAME(fdga_ctx* ctx, int which, fdga_c64* host, int64_t n) { \
    CK(cudaSetDevice(ctx->device)); \
    if (which < 0 || which >= COUNT) FAIL("fdga_get_" #NAME ": bad selector"); \
    if ((size_t)n != (LEN)) FAIL("fdga_get_" #NAME ": length mismatch"); \
    CK(cudaMemcpyAsync(host, ctx->ARR[which], n * sizeof(C), cudaMemcpyDeviceToHost, ctx->stream)); \
    CK(cudaStreamSynchronize(ctx->stream)); return 0; }
SETGET(green, G, 5, ctx->lenG, (void)0)
int fdga_set_bubble(fdga_ctx* ctx, int which, const fdga_c64* host, int64_t n) {
    CK(cudaSetDevice(ctx->device));
    if (which < 0 || which >= 4) FAIL("fdga_set_bubble: bad selector");
    if ((size_t)n != ctx->lenPi) FAIL("fdga_set_bubble: length mismatch");
    if (ctx->swave) {
        CK(cudaMemcpyAsync(ctx->Pisw[which], host, n * sizeof(C), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        return 0;
    }
    if (!ctx->Pi[which]) CK(cudaMalloc(&ctx->Pi[which], ctx->lenPi * sizeof(C)));
    CK(cudaMemcpyAsync(ctx->Pi[which], host, n * sizeof(C), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->pi_src[which] = PI_FULL; ctx->pi_full_valid[which] = true; ctx->pi_dirty[which] = true; invalidate_rt(ctx);
    return 0;
}
int fdga_get_bubble(fdga_ctx* ctx, int which, fdga_c64* host, int64_t n) {
    CK(cudaSetDevice(ctx->device));
    if (which < 0 || which >= 4) FAIL("fdga_get_bubble: bad selector");
    if ((size_t)n != ctx->lenPi) FAIL("fdga_get_bubble: length mismatch");
    if (ctx->swave) {
        CK(cudaMemcpyAsync(host, ctx->Pisw[which], n * sizeof(C), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        return 0;
    }
    if (ensure_pi_full(ctx, which)) return 1;
    CK(cudaMemcpyAsync(host, ctx->Pi[which], n * sizeof(C), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
SETGET(cache, cache, 10, ctx->lenK3, (void)0)
int fdga_get_L(fdga_ctx* ctx, int is_pp, fdga_c64* host, int64_t n) {
    CK(cudaSetDevice(ctx->device));
    if ((size_t)n != ctx->lev[0].len[1]) FAIL("fdga_get_L: length mismatch");
    CK(cudaMemcpyAsync(host, ctx->L[is_pp ? 0 : 1], n * sizeof(C), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream)); return 0;
}

static size_t sg_target_len(fdga_ctx* ctx, int which) {
    if (which == FDGA_SG_SIGMA) return ctx->lenG;
    if (which == FDGA_SG_K1) return ctx->lev[0].len[0];
    if (which == FDGA_SG_PP2 || which == FDGA_SG_PH2) return ctx->lev[0].len[1];
    return ctx->lev[0].len[2];
}
// (Re)build the device tables of a registered symmetry group.  The class ORDER is internal to the library (callers only see
// class members), and it decides what a rank has to touch: ranks own contiguous blocks of classes, and every slab-level kernel
// (right factor, slab_own, slab_conv) runs on the (W, P) slabs that hold a representative of the rank.  In the caller's order
// (ascending linear index: k slowest for K2) every rank touches every slab; with more than one rank the K1 / K2 classes are
// therefore sorted by the slab (P, W) of their representative, so that the slab-level work is sharded like the column work.
static int rebuild_sg(fdga_ctx* ctx, int which) {
    SymGroup& s = ctx->sg[which];
    ctx->epoch++;
    const long long nclasses = (long long)s.o_offsets.size() - 1, nmem = s.o_offsets[nclasses];
    std::vector<long long> perm(nclasses);
    for (long long c = 0; c < nclasses; c++) perm[c] = c;
    if (ctx->nranks > 1 && (which == FDGA_SG_K1 || which == FDGA_SG_PP2 || which == FDGA_SG_PH2)) {
        const Grid& g = ctx->g;
        const long long nB = which == FDGA_SG_K1 ? 2 * g.nK1 - 1 : 2 * g.nK2b - 1, nF = which == FDGA_SG_K1 ? 1 : 2 * g.nK2f;
        std::vector<long long> key(nclasses);
        for (long long c = 0; c < nclasses; c++) {
            long long idx = s.o_index[s.o_offsets[c]];
            const long long iW = idx % nB; idx /= nB; idx /= nF; const long long iP = idx % g.NP;
            key[c] = iW + nB * iP;
        }
        std::stable_sort(perm.begin(), perm.end(), [&](long long a, long long b) { return key[a] < key[b]; });
    }
    std::vector<long long> offs(nclasses + 1), idx(nmem); std::vector<unsigned char> ops(nmem); std::vector<int> mclass(nmem);
    long long pos = 0;
    for (long long c = 0; c < nclasses; c++) {
        offs[c] = pos;
        for (long long j = s.o_offsets[perm[c]]; j < s.o_offsets[perm[c] + 1]; j++) { idx[pos] = s.o_index[j]; ops[pos] = s.o_ops[j]; mclass[pos] = (int)c; pos++; }
    }
    offs[nclasses] = pos;
    cudaFree(s.d_offsets); cudaFree(s.d_index); cudaFree(s.d_ops); cudaFree(s.d_member_class); for (int k = 0; k < 3; k++) cudaFree(s.d_rep[k]);
    s.d_offsets = s.d_index = nullptr; s.d_ops = nullptr; s.d_member_class = nullptr; for (int k = 0; k < 3; k++) s.d_rep[k] = nullptr;
    s.ncls = nclasses; s.nmem = nmem; s.chunk = (nclasses + ctx->nranks - 1) / ctx->nranks;
    s.h_offsets = offs; s.h_index = idx;
    CK(cudaMalloc(&s.d_offsets, (nclasses + 1) * sizeof(long long))); CK(cudaMalloc(&s.d_index, nmem * sizeof(long long)));
    CK(cudaMalloc(&s.d_ops, nmem)); CK(cudaMalloc(&s.d_member_class, nmem * sizeof(int)));
    for (int k = 0; k < 3; k++) {
        CK(cudaMalloc(&s.d_rep[k], (size_t)s.chunk * ctx->nranks * sizeof(C)));
        CK(cudaMemset(s.d_rep[k], 0, (size_t)s.chunk * ctx->nranks * sizeof(C)));
    }
    s.d_repvals = s.d_rep[0];
    CK(cudaMemcpy(s.d_offsets, offs.data(), (nclasses + 1) * sizeof(long long), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(s.d_index, idx.data(), nmem * sizeof(long long), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(s.d_ops, ops.data(), nmem, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(s.d_member_class, mclass.data(), nmem * sizeof(int), cudaMemcpyHostToDevice));
    s.set = true; ctx->slabs_dirty = true;
    if ((which == FDGA_SG_PP2 || which == FDGA_SG_PH2) && !ctx->swave) return build_columns(ctx, s);
    return 0;
}
int fdga_set_symmetry_classes(fdga_ctx* ctx, int which, int64_t nclasses, const int64_t* offsets, const int64_t* index, const uint8_t* ops) {
    CK(cudaSetDevice(ctx->device));
    if (which < 0 || which >= FDGA_SG_COUNT) FAIL("fdga_set_symmetry_classes: bad selector");
    size_t len = sg_target_len(ctx, which);
    long long nmem = offsets[nclasses];
    if (nclasses <= 0 || offsets[0] != 0 || (size_t)nmem > len) FAIL("fdga_set_symmetry_classes: malformed class table");
    std::vector<unsigned char> seen(len, 0);
    for (int64_t c = 0; c < nclasses; c++) {
        if (offsets[c + 1] <= offsets[c]) FAIL("fdga_set_symmetry_classes: empty class");
        for (int64_t j = offsets[c]; j < offsets[c + 1]; j++) {
            if (index[j] < 0 || (size_t)index[j] >= len || seen[index[j]]) FAIL("fdga_set_symmetry_classes: index out of range or repeated");
            seen[index[j]] = 1;
        }
    }
    SymGroup& s = ctx->sg[which];
    s.o_offsets.assign(offsets, offsets + nclasses + 1);
    s.o_index.assign(index, index + nmem);
    s.o_ops.assign(ops, ops + nmem);
    return rebuild_sg(ctx, which);
}
int fdga_build_symmetry_group(int which_sg, int n0, int n1, int nq, int64_t* offsets, int64_t* index, uint8_t* ops, int64_t* nclasses) {
    if (which_sg < 0 || which_sg > FDGA_SG_NL_PH2) return 1;
    *nclasses = fdga_symgroup_build_host(which_sg, n0, n1, nq, offsets, index, ops);
    return 0;
}

int64_t fdga_length_F(fdga_ctx* ctx) { return (int64_t)ctx->lenFlat; }
static int flatten_dev(fdga_ctx* ctx, const LevelBuf& lb, C* dst) {
    CK(cudaMemcpyAsync(dst, lb.block, lb.blocklen * sizeof(C), cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
}
int fdga_flatten_F(fdga_ctx* ctx, fdga_c64* host_y) {
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(host_y, ctx->lev[0].block, ctx->lenFlat * sizeof(C), cudaMemcpyDeviceToHost, ctx->stream));   // the level block IS flatten(S.F)
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
// an asynchronous flatten may still be reading S.F: writers of the level-0 block wait for it (stream order, no host sync)
static int wait_copy(fdga_ctx* ctx) {
    if (ctx->copy_pending) CK(cudaStreamWaitEvent(ctx->main_stream, ctx->ev_copy_done, 0));
    return 0;
}
int fdga_flatten_F_async(fdga_ctx* ctx, fdga_c64* host_y) {
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventRecord(ctx->ev_copy_ready, ctx->main_stream));
    CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy_ready, 0));
    CK(cudaMemcpyAsync(host_y, ctx->lev[0].block, ctx->lenFlat * sizeof(C), cudaMemcpyDeviceToHost, ctx->copy_stream));
    CK(cudaEventRecord(ctx->ev_copy_done, ctx->copy_stream));
    ctx->copy_pending = true;
    return 0;
}
static int unflatten_dev(fdga_ctx* ctx, LevelBuf& lb, const C* src, double scale) {
    if (&lb == &ctx->lev[0] && wait_copy(ctx)) return 1;
    LAUNCH(FDGA_T_MISC, scale_copy_kernel, nblk(lb.blocklen, 256), 256, lb.block, src, scale, (long long)lb.blocklen);
    CK(cudaGetLastError());
    lb.sw_dirty = true; lb.k1h_dirty = true; lb.mom_valid[0] = lb.mom_valid[1] = lb.mom_valid[2] = 0; ctx->fsum_dirty = true;
    return 0;
}
int fdga_stash_F(fdga_ctx* ctx) { CK(cudaSetDevice(ctx->device)); return flatten_dev(ctx, ctx->lev[0], ctx->stash); }
int fdga_unstash_F(fdga_ctx* ctx) { CK(cudaSetDevice(ctx->device)); return unflatten_dev(ctx, ctx->lev[0], ctx->stash, 1.0); }
int fdga_unflatten_F(fdga_ctx* ctx, const fdga_c64* host_x, double scale) {
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(ctx->flat, host_x, ctx->lenFlat * sizeof(C), cudaMemcpyHostToDevice, ctx->stream));
    return unflatten_dev(ctx, ctx->lev[0], ctx->flat, scale);
}

// unflatten!(S.F, x * scale) for a multi-rank job whose input lives on ONE host: the root rank copies x over PCIe once, the other
// ranks receive it over NVLink (ncclBroadcast) -- instead of every rank pulling its own copy of the same vector through PCIe.
// host_x is read on the root only (may be NULL elsewhere).  Collective: every rank of the communicator must call it.
int fdga_unflatten_F_from_root(fdga_ctx* ctx, const fdga_c64* host_x, double scale, int root) {
    CK(cudaSetDevice(ctx->device));
    if (root < 0 || root >= ctx->nranks) FAIL("fdga_unflatten_F_from_root: bad root");
    if (ctx->rank == root) {
        if (!host_x) FAIL("fdga_unflatten_F_from_root: the root needs the host vector");
        CK(cudaMemcpyAsync(ctx->flat, host_x, ctx->lenFlat * sizeof(C), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (ctx->nranks > 1) {
        if (!ctx->nccl.Broadcast) FAIL("fdga_unflatten_F_from_root: ncclBroadcast not available");
        Scope sc(ctx, FDGA_T_COMM);
        int rc = ctx->nccl.Broadcast(ctx->flat, ctx->flat, ctx->lenFlat * 2, /*ncclDouble*/ 8, root, ctx->comm, ctx->stream);
        if (rc != 0) FAIL(std::string("ncclBroadcast: ") + ctx->nccl.GetErrorString(rc));
        ctx->n_launch[FDGA_T_COMM]++;
    }
    return unflatten_dev(ctx, ctx->lev[0], ctx->flat, scale);
}

// ---- Dyson / occupation / bubbles --------------------------------------------------------------------
int fdga_dyson(fdga_ctx* ctx) {
    CK(cudaSetDevice(ctx->device));
    Scope sc(ctx, FDGA_T_MISC);
    LAUNCH(FDGA_T_MISC, dyson_kernel, nblk(ctx->lenG, 256), 256, ctx->G[FDGA_G], ctx->G[FDGA_SIGMA], ctx->G[FDGA_GBARE], (long long)ctx->lenG);
    CK(cudaGetLastError()); return 0;
}
static int occupation_dev(fdga_ctx* ctx, int which) {
    LAUNCH(FDGA_T_MISC, occupation_kernel, 1, 1024, ctx->G[which], (long long)ctx->lenG, ctx->g.T, (double)(ctx->g.LG * ctx->g.LG), ctx->d_occ);
    CK(cudaGetLastError()); return 0;
}
int fdga_occupation(fdga_ctx* ctx, int which, double* occ) {
    CK(cudaSetDevice(ctx->device));
    if (which != FDGA_G && which != FDGA_G0 && which != FDGA_GBARE) FAIL("fdga_occupation: bad selector");
    if (occupation_dev(ctx, which)) return 1;
    CK(cudaMemcpyAsync(occ, ctx->d_occ, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
// set!(S.Gbare, hubbard_bare_Green(meshes(S.Gbare)...; mu, t1, t2, t3)): src/models/hubbard.jl:8-29
int fdga_set_hubbard_bare_green(fdga_ctx* ctx, double mu, double t1, double t2, double t3) {
    CK(cudaSetDevice(ctx->device));
    if (ctx->opt_local) FAIL("fdga_set_hubbard_bare_green: lattice model only (the local solver uses siam_bare_Green on the host)");
    Scope sc(ctx, FDGA_T_MISC);
    LAUNCH(FDGA_T_MISC, hubbard_bare_green_kernel, nblk(ctx->lenG, 256), 256, ctx->G[FDGA_GBARE], ctx->dims.nG, ctx->g.LG, ctx->g.T, mu, t1, t2, t3);
    CK(cudaGetLastError());
    return 0;
}
// compute_hubbard_chemical_potential(occ_target, S.Sigma, (; t1, t2, t3)): src/dyson.jl:45-57.  Roots.find_zero on the bracket
// (-4|t1|, 4|t1|) = bisection down to neighbouring floating-point numbers; occupation(mu) is one fused kernel per evaluation.
int fdga_hubbard_chemical_potential(fdga_ctx* ctx, double occ_target, double t1, double t2, double t3, double* mu_out) {
    CK(cudaSetDevice(ctx->device));
    if (ctx->opt_local) FAIL("fdga_hubbard_chemical_potential: lattice model only");
    auto f = [&](double mu, double& val) -> int {
        double occ = 0.0;
        LAUNCH(FDGA_T_MISC, occupation_mu_kernel, 1, 1024, ctx->G[FDGA_SIGMA], ctx->dims.nG, ctx->g.LG, ctx->g.T, mu, t1, t2, t3, ctx->d_occ);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&occ, ctx->d_occ, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        val = occ - occ_target;
        return 0;
    };
    double a = -4.0 * fabs(t1), b = 4.0 * fabs(t1), fa, fb;
    if (f(a, fa) || f(b, fb)) return 1;
    if (fa == 0.0) { *mu_out = a; return 0; }
    if (fb == 0.0) { *mu_out = b; return 0; }
    if ((fa < 0.0) == (fb < 0.0) || fa != fa || fb != fb)
        FAIL("fdga_hubbard_chemical_potential: the interval [-4|t1|, 4|t1|] is not a bracketing interval (occupation target out of reach)");
    for (int it = 0; it < 200; ++it) {
        const double m = a + 0.5 * (b - a);
        if (m <= a || m >= b) break;                  // a and b are neighbouring floats
        double fm; if (f(m, fm)) return 1;
        if (fm == 0.0) { a = b = m; break; }
        if ((fm < 0.0) == (fa < 0.0)) { a = m; fa = fm; } else { b = m; fb = fm; }
    }
    *mu_out = (fabs(fa) <= fabs(fb)) ? a : b;
    return 0;
}
int fdga_bubbles_real_space(fdga_ctx* ctx, int reference) {
    CK(cudaSetDevice(ctx->device));
    Scope sc(ctx, FDGA_T_BUBBLE);
    const Grid& g = ctx->g;
    int ipp = reference ? FDGA_PI0PP : FDGA_PIPP, iph = reference ? FDGA_PI0PH : FDGA_PIPH;
    const C* Gsrc = ctx->G[reference ? FDGA_G0 : FDGA_G];
    if (dft2_G(ctx, Gsrc, ctx->GR, ctx->SigTmp, -1, 1.0 / ((double)g.LG * g.LG), FDGA_T_BUBBLE)) return 1;
    if (ctx->swave) {       // bubbles_real_space!(::NL_MF_Pi; use_G_tail = true): src/nonlocal/bubble.jl:87-158
        const long long pre = (long long)(2 * g.nPiB - 1) * (2 * g.nPiF), n = pre * g.NP;
        static const bool literal = getenv("FDGA_BUBBLES_RS") ? atoi(getenv("FDGA_BUBBLES_RS")) != 0 : false;
        if (!literal) {     // back transform as a direct sum on the few inner frequencies that carry one (fdga_swave.cuh)
            const long long nb_fill = nblk(n, 128), nb_sum = nblk((long long)(2 * g.nPiB - 1) * (2 * g.nG) * g.NP * 32, 128);
            LAUNCH(FDGA_T_BUBBLE, sw_bubbles_direct_kernel, (unsigned)(nb_fill + nb_sum), 128, ctx->GR, ctx->Pisw[ipp], ctx->Pisw[iph], g, 1, ctx->twL, nb_fill);
            CK(cudaGetLastError());
            return 0;
        }
        for (int i = 0; i < 2; i++) if (!ctx->swScratch[i]) CK(cudaMalloc(&ctx->swScratch[i], ctx->lenPisw * sizeof(C)));
        LAUNCH(FDGA_T_BUBBLE, sw_bubbles_rs_kernel, nblk(n, 128), 128, ctx->GR, ctx->swScratch[0], ctx->swScratch[1], g, 1);
        for (int i = 0; i < 2; i++) {      // back transform over the two momentum axes; Pisw[] as the intermediate
            C* dst = ctx->Pisw[i == 0 ? ipp : iph];
            LAUNCH(FDGA_T_BUBBLE, dft_axis_kernel, nblk(n, 128), 128, ctx->swScratch[i], dst, pre, g.L, (long long)g.L, +1, 1.0, ctx->twL);
            LAUNCH(FDGA_T_BUBBLE, dft_axis_kernel, nblk(n, 128), 128, dst, ctx->swScratch[i], pre * g.L, g.L, 1LL, +1, 1.0, ctx->twL);
            CK(cudaMemcpyAsync(dst, ctx->swScratch[i], (size_t)n * sizeof(C), cudaMemcpyDeviceToDevice, ctx->stream));
        }
        CK(cudaGetLastError());
        return 0;
    }
    static const bool legacy = getenv("FDGA_BUBBLES_RS") ? atoi(getenv("FDGA_BUBBLES_RS")) != 0 : false;
    if (!legacy) {      // product form on the coarse-grained Green function (fdga_kernels.cuh); the slabs are filled lazily (refresh_pi)
        C* Ghat = ctx->Ghat[reference ? 1 : 0];
        LAUNCH(FDGA_T_BUBBLE, coarse_green_kernel, nblk((long long)2 * g.nG * g.NP, 128), 128, ctx->GR, Ghat, g.nG, g.LG, g.L, ctx->twL);
        CK(cudaGetLastError());
        ctx->pi_src[ipp] = ctx->pi_src[iph] = PI_GHAT; ctx->pi_full_valid[ipp] = ctx->pi_full_valid[iph] = false;
    } else {            // the reference's own route: real-space fill + 4-d back transform (A/B check)
        for (int w : {ipp, iph}) if (!ctx->Pi[w]) CK(cudaMalloc(&ctx->Pi[w], ctx->lenPi * sizeof(C)));
        if (!ctx->scratchBig) CK(cudaMalloc(&ctx->scratchBig, ctx->lenPi * sizeof(C)));
        LAUNCH(FDGA_T_BUBBLE, bubbles_rs_kernel, nblk(ctx->lenPi, 128), 128, ctx->GR, ctx->Pi[ipp], ctx->Pi[iph], g);
        CK(cudaGetLastError());
        long long pre = (long long)(2 * g.nPiB - 1) * (2 * g.nPiF);
        if (dft4(ctx, ctx->Pi[ipp], ctx->scratchBig, pre, +1, 1.0, FDGA_T_BUBBLE)) return 1;
        if (dft4(ctx, ctx->Pi[iph], ctx->scratchBig, pre, +1, 1.0, FDGA_T_BUBBLE)) return 1;
        ctx->pi_src[ipp] = ctx->pi_src[iph] = PI_FULL; ctx->pi_full_valid[ipp] = ctx->pi_full_valid[iph] = true;
    }
    ctx->pi_dirty[ipp] = ctx->pi_dirty[iph] = true;
    invalidate_rt(ctx);
    return 0;
}
int fdga_bubbles_local(fdga_ctx* ctx, int reference) {
    CK(cudaSetDevice(ctx->device));
    if (ctx->g.L != 1 || ctx->g.LG != 1) FAIL("fdga_bubbles_local: needs nq = LG = 1");
    if (ctx->swave) FAIL("fdga_bubbles_local: not for an s-wave (NL) context");
    Scope sc(ctx, FDGA_T_BUBBLE);
    int ipp = reference ? FDGA_PI0PP : FDGA_PIPP, iph = reference ? FDGA_PI0PH : FDGA_PIPH;
    for (int w : {ipp, iph}) { if (!ctx->Pi[w]) CK(cudaMalloc(&ctx->Pi[w], ctx->lenPi * sizeof(C))); ctx->pi_src[w] = PI_FULL; ctx->pi_full_valid[w] = true; }
    LAUNCH(FDGA_T_BUBBLE, bubbles_local_kernel, nblk(ctx->lenPi, 128), 128, ctx->G[reference ? FDGA_G0 : FDGA_G], ctx->Pi[ipp], ctx->Pi[iph], ctx->g);
    CK(cudaGetLastError());
    ctx->pi_dirty[ipp] = ctx->pi_dirty[iph] = true;
    invalidate_rt(ctx);
    return 0;
}
int fdga_bubbles_momentum_space(fdga_ctx* ctx, int reference) {
    CK(cudaSetDevice(ctx->device));
    if (ctx->g.LG % ctx->g.L != 0) FAIL("fdga_bubbles_momentum_space: LG must be a multiple of nq");
    if (ctx->swave) FAIL("fdga_bubbles_momentum_space: the s-wave solver's bubbles! is bubbles_real_space! (src/nonlocal/ParquetSolver.jl:294-297)");
    Scope sc(ctx, FDGA_T_BUBBLE);
    int ipp = reference ? FDGA_PI0PP : FDGA_PIPP, iph = reference ? FDGA_PI0PH : FDGA_PIPH;
    for (int w : {ipp, iph}) { if (!ctx->Pi[w]) CK(cudaMalloc(&ctx->Pi[w], ctx->lenPi * sizeof(C))); ctx->pi_src[w] = PI_FULL; ctx->pi_full_valid[w] = true; }
    LAUNCH(FDGA_T_BUBBLE, bubbles_ms_kernel, nblk(ctx->lenPi, 128), 128, ctx->G[reference ? FDGA_G0 : FDGA_G], ctx->Pi[ipp], ctx->Pi[iph], ctx->g);
    CK(cudaGetLastError());
    ctx->pi_dirty[ipp] = ctx->pi_dirty[iph] = true;
    invalidate_rt(ctx);
    return 0;
}

// ---- the vertex as a callable ----------------------------------------------------------------------------
// F(W, v, w, P, k, q, Ch, Sp; F0, gamma_p, gamma_t, gamma_a) of the chain S.F from `level` at n points (host arrays in, host
// values out).  flags: bit 0 F0, bit 1 gamma_p, bit 2 gamma_t, bit 3 gamma_a.  Synchronous; not a hot path.
int fdga_eval_vertex(fdga_ctx* ctx, int level, int ch, int sp, int flags, int swave_kq, int64_t n, const int32_t* W, const int32_t* v, const int32_t* w,
                     const int32_t* iP, const int32_t* ik, const int32_t* iq, fdga_c64* out) {
    CK(cudaSetDevice(ctx->device));
    if (level < 0 || level >= ctx->nlev || ch < 0 || ch > 2 || sp < 0 || sp > 2 || n < 0) FAIL("fdga_eval_vertex: bad selector");
    if (ctx->swave && !swave_kq) FAIL("fdga_eval_vertex: an s-wave (NL) context evaluates at k = q = kSW only");
    if (n == 0) return 0;
    for (int64_t i = 0; i < n; i++) if (iP[i] < 0 || iP[i] >= ctx->g.NP || ik[i] < 0 || ik[i] >= ctx->g.NP || iq[i] < 0 || iq[i] >= ctx->g.NP) FAIL("fdga_eval_vertex: momentum index out of range");
    if (swave_kq && refresh_swave(ctx)) return 1;
    int* d = nullptr; C* dout = nullptr;
    CK(cudaMalloc(&d, (size_t)6 * n * sizeof(int))); CK(cudaMalloc(&dout, (size_t)n * sizeof(C)));
    const int32_t* src[6] = {W, v, w, iP, ik, iq};
    for (int j = 0; j < 6; j++) CK(cudaMemcpyAsync(d + (size_t)j * n, src[j], (size_t)n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    DevChain V = chain_F(ctx, 0);
    {
        Scope sc(ctx, FDGA_T_MISC);
        if (ctx->mbe) LAUNCH(FDGA_T_MISC, eval_points_kernel<true>, nblk(n, 64), 64, V, level, ch, sp, (unsigned)flags, swave_kq, (long long)n, d, d + n, d + 2 * n, d + 3 * n, d + 4 * n, d + 5 * n, dout);
        else          LAUNCH(FDGA_T_MISC, eval_points_kernel<false>, nblk(n, 64), 64, V, level, ch, sp, (unsigned)flags, swave_kq, (long long)n, d, d + n, d + 2 * n, d + 3 * n, d + 4 * n, d + 5 * n, dout);
    }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout, (size_t)n * sizeof(C), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d); cudaFree(dout);
    if (e != cudaSuccess) FAIL(std::string("fdga_eval_vertex: ") + cudaGetErrorString(e));
    return 0;
}

// ---- K3 cache ------------------------------------------------------------------------------------------
int fdga_build_K3_cache(fdga_ctx* ctx, int mfrg, int first) {
    CK(cudaSetDevice(ctx->device));
    if (refresh_swave(ctx)) return 1;
    Scope sc(ctx, FDGA_T_CACHE);
    DevChain F0 = chain_F(ctx, 1), F = chain_F(ctx, 0);
    if (!mfrg) {
        CachePtrs cp; for (int i = 0; i < 10; i++) cp.c[i] = ctx->cache[i];
        if (ctx->lev[0].mbe) {
            const int np2 = ctx->g.NP * ctx->g.NP, thr = np2 >= 128 ? 128 : (np2 >= 64 ? 64 : 32);
            LAUNCH(FDGA_T_CACHE, build_cache_mbe_kernel, (unsigned)ctx->lenK3, thr, F0, F, cp, ctx->g, 0LL, (long long)ctx->lenK3);
        }
        else
        LAUNCH(FDGA_T_CACHE, build_cache_kernel, nblk(ctx->lenK3, 64), 64, F0, F, cp, ctx->g, 0LL, (long long)ctx->lenK3);
        CK(cudaGetLastError());
        return 0;
    }
    NEED_SG(FDGA_SG_PP3); NEED_SG(FDGA_SG_PH3);
    struct { int kind, cache, sg; } jobs[7] = { {0, FDGA_C_GPX, FDGA_SG_PP3}, {1, FDGA_C_GPP, FDGA_SG_PP3}, {2, FDGA_C_GA, FDGA_SG_PH3}, {3, FDGA_C_GT, FDGA_SG_PH3},
                                               {4, FDGA_C_FP, FDGA_SG_PP3}, {5, FDGA_C_FA, FDGA_SG_PH3}, {6, FDGA_C_FT, FDGA_SG_PH3} };
    int njobs = first ? 7 : 4;
    for (int j = 0; j < njobs; j++) {
        SymGroup& s = ctx->sg[jobs[j].sg];
        long long c0, c1; sg_class_range(ctx, s, c0, c1);
        if (c1 > c0) {
            if (ctx->mbe) LAUNCH(FDGA_T_CACHE, cache_mfrg_kernel<true>, nblk(c1 - c0, 64), 64, F0, F, jobs[j].kind, s.d_repvals, sym_dev(s), c0, c1, ctx->g);
            else          LAUNCH(FDGA_T_CACHE, cache_mfrg_kernel<false>, nblk(c1 - c0, 64), 64, F0, F, jobs[j].kind, s.d_repvals, sym_dev(s), c0, c1, ctx->g);
        }
        CK(cudaGetLastError());
        if (sg_finish(ctx, s, ctx->cache[jobs[j].cache])) return 1;
        if (jobs[j].kind == 3) { if (axpby(ctx, ctx->cache[FDGA_C_GT], ctx->cache[FDGA_C_GT], 2.0, ctx->cache[FDGA_C_GA], -1.0, ctx->lenK3)) return 1; }
        if (jobs[j].kind == 6) { if (axpby(ctx, ctx->cache[FDGA_C_FT], ctx->cache[FDGA_C_FT], 2.0, ctx->cache[FDGA_C_FA], -1.0, ctx->lenK3)) return 1; }
    }
    return 0;
}

// ---- BSE kernels -----------------------------------------------------------------------------------------
// s-wave contraction kernels (fdga_swave.cuh): one CTA per class representative when the inner sum is long, else one warp
static bool sw_cta_per_rep(int nw) {
    static const int lim = getenv("FDGA_SW_CTA_MIN") ? atoi(getenv("FDGA_SW_CTA_MIN")) : 128;
    return nw >= lim;
}
static int pi_kind(int ch, bool reference) { return ch == FDGA_PCH ? (reference ? FDGA_PI0PP : FDGA_PIPP) : (reference ? FDGA_PI0PH : FDGA_PIPH); }

extern "C++" {
template <int KIND>
static int launch_right(fdga_ctx* ctx, int ch, const DevChain& F0, const DevChain& FL, int No, int Ninner, C* Rdst = nullptr) {
    if (ensure_slabs(ctx)) return 1;
    if (!Rdst) Rdst = ctx->RtL[ctx->cur_lane];
    Scope sc(ctx, FDGA_T_RIGHT);
    const C* p0 = ctx->PiT[pi_kind(ch, true)]; const C* p1 = ctx->PiT[pi_kind(ch, false)];
    const int kind = (ch == FDGA_PCH ? 0 : 1) + (No == ctx->g.nPiB ? 0 : 2);
    const int nsl = ctx->n_slabs[kind]; const int4* sl = ctx->d_slabs[kind];
    long long n = (long long)(2 * Ninner) * ctx->g.NP * nsl;
    if (n == 0) return 0;
    const int* pimap = ctx->d_slabmap[ch == FDGA_PCH ? 0 : 1];      // the bubbles live on the bubble-mesh slab list
    if (ch == FDGA_PCH)      { if (ctx->mbe) LAUNCH(FDGA_T_RIGHT, (right_factor_kernel<CH_P, KIND, true>), nblk(n, 128), 128, F0, FL, p0, p1, Rdst, ctx->g, No, Ninner, sl, nsl, pimap); else LAUNCH(FDGA_T_RIGHT, (right_factor_kernel<CH_P, KIND>), nblk(n, 128), 128, F0, FL, p0, p1, Rdst, ctx->g, No, Ninner, sl, nsl, pimap); }
    else if (ch == FDGA_TCH) { if (ctx->mbe) LAUNCH(FDGA_T_RIGHT, (right_factor_kernel<CH_T, KIND, true>), nblk(n, 128), 128, F0, FL, p0, p1, Rdst, ctx->g, No, Ninner, sl, nsl, pimap); else LAUNCH(FDGA_T_RIGHT, (right_factor_kernel<CH_T, KIND>), nblk(n, 128), 128, F0, FL, p0, p1, Rdst, ctx->g, No, Ninner, sl, nsl, pimap); }
    else                     { if (ctx->mbe) LAUNCH(FDGA_T_RIGHT, (right_factor_kernel<CH_A, KIND, true>), nblk(n, 128), 128, F0, FL, p0, p1, Rdst, ctx->g, No, Ninner, sl, nsl, pimap); else LAUNCH(FDGA_T_RIGHT, (right_factor_kernel<CH_A, KIND>), nblk(n, 128), 128, F0, FL, p0, p1, Rdst, ctx->g, No, Ninner, sl, nsl, pimap); }
    CK(cudaGetLastError());
    return 0;
}

template <int KIND, int CH>
static void kernel_slab_own(fdga_ctx* ctx, const DevChain& V, const ColJob& job, int kind, const C* R, int cat) {
    const int PC = std::max(1, std::min(2 * ctx->g.nK2f, 2048 / job.nw));
    size_t smem = (size_t)job.nw * (1 + PC) * sizeof(C);
    // TMA staging of the slab (read twice by the SDE jobs) while it keeps the kernel's residency: opt-in until measured
    static const int tma_on = getenv("FDGA_OWN_TMA") ? atoi(getenv("FDGA_OWN_TMA")) : 0;
    const size_t slab_bytes = (size_t)job.nw * ctx->g.NP * sizeof(C);
    const int use_tma = (tma_on && slab_bytes <= 64 * 1024) ? 1 : 0;
    if (use_tma) smem += slab_bytes;
    // long inner meshes (local solver, m_Pi_nu_factor = 6: nw = 3072 at BASELINE config 2) need the opt-in shared memory carve-out
    if (smem + 2048 > 48 * 1024) cudaFuncSetAttribute(slab_own_kernel<KIND, CH>,     // static + dynamic must stay under 48 KB without the opt-in
        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::min<size_t>(smem, 227 * 1024));
    slab_own_kernel<KIND, CH><<<ctx->n_slabs[kind], 256, smem, ctx->stream>>>(V, job, ctx->d_slabs[kind], R, ctx->TtabL[ctx->cur_lane], ctx->OwnTabL[ctx->cur_lane], ctx->RtotL[ctx->cur_lane], ctx->g, PC, use_tma);
    NOTE_LAUNCH("slab_own_kernel");
    ctx->n_launch[cat]++; ctx->total_launches++;
}
// column path (fdga_column.cuh): momentum-independent table + one CTA per output column
// cross-channel K1 pieces of a column job as a per-slab momentum convolution; returns the table (or null when the job
// has no such piece / the slab does not fit in shared memory, in which case job.k1_direct is switched on)
template <int KIND, int CH>
static int launch_slab_conv(fdga_ctx* ctx, const DevChain& V, ColJob& job, int kind, const C* R, int cat, const C** tab) {
    *tab = nullptr;
    if (KIND == JOB_LK2_LOC || job.k1_direct) return 0;
    bool any = false;
    for (int l = job.lev_first; l < conv_level_end<KIND>(job); ++l) any = any || conv_level_on<KIND>(job, l);
    if (!any || ctx->n_slabs[kind] == 0) return 0;
    const Grid& g = ctx->g;
    const int nF2 = 2 * g.nK2f;
    const size_t budget = 200 * 1024;
    auto bytes = [&](int tw) { return ((size_t)2 * g.NP * (tw | 1) + (size_t)nF2 * g.NP + g.L) * sizeof(C); };
    int TW = std::max(job.nw, nF2);
    // narrow win tiles keep the slab's shared memory small, so that every slab CTA of the launch is resident at once (config 3:
    // 404 slabs, 76 KB -> 2 CTAs/SM = 1.36 waves with one 32-wide tile; 8-wide tiles -> 6 CTAs/SM, one wave, +4 % per step)
    static const int tw_cap = getenv("FDGA_CONV_TW") ? atoi(getenv("FDGA_CONV_TW")) : 8;      // tuning knob (tile width in win)
    if (tw_cap > 0) TW = std::min(TW, std::max(nF2, tw_cap));
    while (TW > nF2 && bytes(TW) > budget) TW = std::max(nF2, (TW + 1) / 2);
    if (bytes(TW) > budget) { job.k1_direct = 1; return 0; }
    if (refresh_k1h(ctx)) return 1;
    // TMA staging of the whole R slab (one cp.async.bulk per CTA) when it keeps >= 3 CTAs per SM resident; opt-in until measured
    static const int tma_on = getenv("FDGA_CONV_TMA") ? atoi(getenv("FDGA_CONV_TMA")) : 0;
    const size_t slab_bytes = (size_t)job.nw * g.NP * sizeof(C);
    const int use_tma = (tma_on && slab_bytes % 16 == 0 && bytes(TW) + slab_bytes <= 72 * 1024) ? 1 : 0;
    const size_t smem = bytes(TW) + (use_tma ? slab_bytes : 0);
    CK(cudaFuncSetAttribute(slab_conv_kernel<KIND, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // big momentum meshes: the slab's shared memory limits the SM to two CTAs, so each gets twice the threads
    static const int conv_thr_env = getenv("FDGA_CONV_THREADS") ? atoi(getenv("FDGA_CONV_THREADS")) : 0;
    const int conv_thr = conv_thr_env ? conv_thr_env : (smem > 64 * 1024 ? 512 : 256);
    slab_conv_kernel<KIND, CH><<<ctx->n_slabs[kind], conv_thr, smem, ctx->stream>>>(V, job, ctx->d_slabs[kind], R, ctx->twL, ctx->ConvTabL[ctx->cur_lane], g, TW, use_tma);
    NOTE_LAUNCH("slab_conv_kernel");
    ctx->n_launch[cat]++; ctx->total_launches++;
    *tab = ctx->ConvTabL[ctx->cur_lane];
    return 0;
}
// column path (fdga_column.cuh): momentum-independent table + one CTA per output column
template <int KIND, int CH>
static int launch_column_t(fdga_ctx* ctx, const DevChain& V, ColJob job, SymGroup& s, const C* R, int cat) {
    if (ensure_slabs(ctx)) return 1;
    Scope sc(ctx, cat);
    const C* own = nullptr; const C* rtot = nullptr; const C* conv = nullptr;
    const bool pp = (KIND == JOB_SDE_PP) || ((KIND == JOB_K2 || KIND == JOB_K2_MF || KIND == JOB_LK2 || KIND == JOB_LK2_LOC) && CH == CH_P);
    const int kind = (pp ? 0 : 1) + 2;
    job.slabmap = ctx->d_slabmap[(pp ? 0 : 1) + (job.slabW_N == ctx->g.nPiB ? 0 : 2)];      // slab numbering of R (bubble mesh or K2 mesh)
    if (KIND != JOB_LK2 && KIND != JOB_LK2_LOC) {
        // prologue: momentum-independent levels tabulated per (W, nu, w), then everything that does not depend on the
        // column momentum k reduced per slab (W, P): OwnTab[nu | W, P], Rtot[W, P]
        long long n = (long long)job.nw * (2 * ctx->g.nK2f) * (2 * ctx->g.nK2b - 1);
        LAUNCH(cat, (loc_table_kernel<KIND, CH>), nblk(n, 32), 32, V, job, ctx->g, ctx->TtabL[ctx->cur_lane]);
        if (ctx->n_slabs[kind] > 0)
            kernel_slab_own<KIND, CH>(ctx, V, job, kind, R, cat);
        own = ctx->OwnTabL[ctx->cur_lane]; rtot = ctx->RtotL[ctx->cur_lane];
    }
    if (launch_slab_conv<KIND, CH>(ctx, V, job, kind, R, cat, &conv)) return 1;
    const bool sub = ctx->profile && (KIND == JOB_K2 || KIND == JOB_K2_MF) && s.ncol > 0;   // sub-timer: the column kernel alone
    TimedEvent ev; ev.cat = FDGA_T_COLUMN_K2;
    if (sub) { cudaEventCreate(&ev.a); cudaEventCreate(&ev.b); cudaEventRecord(ev.a, ctx->stream); }
    if (KIND != JOB_LK2_LOC && qlane_enabled(ctx)) {
        // the momentum-fastest table copies are made current before the lanes fork; on one stream, right here
        if (!ctx->forked && refresh_mom_all(ctx)) return 1;
        RepDev rd; rd.nrep = s.nrep; rd.rep = s.d_reps;
        if (s.nrep > 0) {
            const size_t ring = (size_t)FDGA_QL_DEPTH * 4096 * FDGA_QL_WARPS;       // cp.async ring of qlane_consume_async
            if (ring + 4096 > 48 * 1024) {
                static bool attr_set = false;       // per template instance
                if (!attr_set) { CK(cudaFuncSetAttribute(qlane_kernel<KIND, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring)); attr_set = true; }
            }
            qlane_kernel<KIND, CH><<<nblk(s.nrep, FDGA_QL_WARPS), 32 * FDGA_QL_WARPS, ring, ctx->stream>>>(V, job, rd, R, own, rtot, conv, s.d_repvals, ctx->g);
            NOTE_LAUNCH("qlane_kernel");
            ctx->n_launch[cat]++; ctx->total_launches++;
        }
    } else
    if (s.ncol > 0) LAUNCH(cat, (column_kernel<KIND, CH>), (unsigned)s.ngrp, 128, V, job, col_dev(s), R, own, rtot, conv, s.d_repvals, ctx->g);
    if (sub) { cudaEventRecord(ev.b, ctx->stream); ctx->events.push_back(ev); ctx->n_launch[FDGA_T_COLUMN_K2]++; }
    CK(cudaGetLastError());
    return 0;
}
template <int KIND>
static int launch_column(fdga_ctx* ctx, int ch, const DevChain& V, ColJob job, SymGroup& s, const C* R, int cat) {
    if (ch == FDGA_PCH) return launch_column_t<KIND, CH_P>(ctx, V, job, s, R, cat);
    if (ch == FDGA_TCH) return launch_column_t<KIND, CH_T>(ctx, V, job, s, R, cat);
    return launch_column_t<KIND, CH_A>(ctx, V, job, s, R, cat);
}
}  // extern "C++"
static ColJob make_job(fdga_ctx* ctx, int lev_first, int nw, int Ninner, int slabN, C scale) {
    ColJob j; j.lev_first = lev_first; j.n_nl2 = ctx->n_nl2; j.own_only = ctx->opt_sde_own_gamma; j.k1_direct = ctx->opt_direct_k1; j.nw = nw; j.Ninner = Ninner;
    j.slabW_N = slabN; j.scale_re = scale.x; j.scale_im = scale.y; j.slabmap = nullptr;
    return j;
}
// right factor of channel ch with W on the bubble mesh, cached per channel (K1 and K2 share it: SURVEY App. C.3)
static int cached_right(fdga_ctx* ctx, int ch, int kind, const DevChain& F0, const DevChain& FL) {
    int tag = kind;
    if (kind == RK_MF_K2 && ch != FDGA_PCH) tag = RK_MF_K1;      // _crossing is the identity for a, t
    if (ctx->rt_kind[ch] == tag) return 0;
    int rc;
    if (kind == RK_FD) rc = launch_right<RK_FD>(ctx, ch, F0, FL, ctx->g.nPiB, ctx->g.nPiF, ctx->Rt3[ch]);
    else if (kind == RK_MF_K1) rc = launch_right<RK_MF_K1>(ctx, ch, F0, FL, ctx->g.nPiB, ctx->g.nPiF, ctx->Rt3[ch]);
    else if (kind == RK_1L) rc = launch_right<RK_1L>(ctx, ch, F0, FL, ctx->g.nPiB, ctx->g.nPiF, ctx->Rt3[ch]);
    else rc = launch_right<RK_MF_K2>(ctx, ch, F0, FL, ctx->g.nPiB, ctx->g.nPiF, ctx->Rt3[ch]);
    if (rc) return rc;
    ctx->rt_kind[ch] = tag;
    return 0;
}
static int ensure_pi(fdga_ctx* ctx, int ch) {
    if (refresh_pi(ctx, pi_kind(ch, true))) return 1;
    return refresh_pi(ctx, pi_kind(ch, false));
}
static double chsign(int ch) { return ch == FDGA_TCH ? -1.0 : 1.0; }   // BSE_templates.jl:17,25,33

// post-processing that follows the SG(...) fill of one BSE call (BSE_templates.jl:35-38 etc., BSEa_K2.jl:130-135)
static int post_fix(fdga_ctx* ctx, int kind, int ch) {
    Scope sc(ctx, FDGA_T_MISC);
    switch (kind) {
    case PK_K1:  if (ch == FDGA_TCH) return tfix(ctx, ctx->Fbuff.K[FDGA_TCH][0], ctx->Fbuff.K[FDGA_ACH][0], ctx->Fbuff.len[0]); return 0;
    case PK_LK2: ctx->FL.sw_dirty = true; invalidate_rt(ctx);
                 if (ch == FDGA_TCH) return tfix(ctx, ctx->FL.K[FDGA_TCH][1], ctx->FL.K[FDGA_ACH][1], ctx->FL.len[1]); return 0;
    case PK_K2: {
        C* out = ctx->Fbuff.K[ch][1]; size_t n = ctx->Fbuff.len[1];
        if (ch == FDGA_TCH) {
            if (add_axpby(ctx, out, ctx->FL.K[FDGA_TCH][1], 2.0, ctx->FL.K[FDGA_ACH][1], -1.0, n)) return 1;
            return tfix(ctx, out, ctx->Fbuff.K[FDGA_ACH][1], n);
        }
        return add_axpby(ctx, out, ctx->FL.K[ch][1], 1.0, nullptr, 0.0, n);
    }
    case PK_K2_NOFL: if (ch == FDGA_TCH) return tfix(ctx, ctx->Fbuff.K[FDGA_TCH][1], ctx->Fbuff.K[FDGA_ACH][1], ctx->Fbuff.len[1]); return 0;
    case PK_LK3: ctx->FL.sw_dirty = true; invalidate_rt(ctx);
                 if (ch == FDGA_TCH) return tfix(ctx, ctx->FL.K[FDGA_TCH][2], ctx->FL.K[FDGA_ACH][2], ctx->FL.len[2]); return 0;
    case PK_K3:  if (ch == FDGA_TCH) return tfix(ctx, ctx->Fbuff.K[FDGA_TCH][2], ctx->Fbuff.K[FDGA_ACH][2], ctx->Fbuff.len[2]); return 0;
    }
    return 0;
}
static int finish_or_defer(fdga_ctx* ctx, SymGroup& s, C* out, int kind, int ch) {
    if (ctx->defer) {       // fused stage: everything is finished together by flush_pending
        Pending p; p.s = &s; p.rep = s.d_repvals; p.out = out; p.kind = kind; p.ch = ch; p.expanded = false;
        ctx->pending.push_back(p); return 0;
    }
    if (sg_finish(ctx, s, out)) return 1;
    return post_fix(ctx, kind, ch);
}
// one NCCL group for all deferred all-gathers of a stage; then, when every kernel class of the stage has its three channels
// pending (the fused drivers), ONE launch applies the post-fixes to the representatives and ONE launch expands all arrays;
// otherwise the expansions and post-fixes run one by one in call order
static int flush_pending(fdga_ctx* ctx) {
    if (ctx->pending.empty()) return 0;
    if (ctx->nranks > 1) {
        Scope sc(ctx, FDGA_T_COMM);
        ctx->nccl.GroupStart();
        for (auto& p : ctx->pending) {
            int rc = ctx->nccl.AllGather(p.rep + (size_t)ctx->rank * p.s->chunk, p.rep, (size_t)p.s->chunk * 2, /*ncclDouble*/ 8, ctx->comm, ctx->stream);
            if (rc != 0) { ctx->nccl.GroupEnd(); ctx->pending.clear(); FAIL(std::string("ncclAllGather: ") + ctx->nccl.GetErrorString(rc)); }
        }
        int rc = ctx->nccl.GroupEnd();
        if (rc != 0) { ctx->pending.clear(); FAIL(std::string("ncclGroupEnd: ") + ctx->nccl.GetErrorString(rc)); }
        ctx->n_launch[FDGA_T_COMM]++;
    }
    std::vector<Pending> todo; todo.swap(ctx->pending);
    // the post-fixes need the order (kernel class, then p, a, t) whatever order the lanes were issued in
    std::stable_sort(todo.begin(), todo.end(), [](const Pending& a, const Pending& b) {
        auto rank = [](int ch) { return ch == FDGA_PCH ? 0 : (ch == FDGA_ACH ? 1 : 2); };
        return a.kind != b.kind ? a.kind < b.kind : rank(a.ch) < rank(b.ch); });
    // batched form: groups of three consecutive entries (p, a, t) of one kind
    bool batched = todo.size() % 3 == 0 && todo.size() / 3 <= FDGA_MAXFIX && todo.size() <= FDGA_MAXEXP;
    for (size_t i = 0; batched && i < todo.size(); i += 3)
        batched = todo[i].kind == todo[i + 1].kind && todo[i].kind == todo[i + 2].kind &&
                  todo[i].ch == FDGA_PCH && todo[i + 1].ch == FDGA_ACH && todo[i + 2].ch == FDGA_TCH && todo[i + 1].s == todo[i + 2].s;
    if (batched) {
        RepFixJobs fj; ExpandJobs ej; memset(&fj, 0, sizeof(fj)); memset(&ej, 0, sizeof(ej));
        long long maxcls = 0, maxmem = 0;
        const int nk = (int)todo.size() / 3;
        for (int k = 0; k < nk; k++) {
            RepFixJob& J = fj.j[k];
            const int order[3] = {FDGA_PCH, FDGA_ACH, FDGA_TCH};
            for (int i = 0; i < 3; i++) {
                const Pending& p = todo[3 * k + i];
                const int ch = order[i];
                J.rep[ch] = p.rep; J.offsets[ch] = p.s->d_offsets; J.index[ch] = p.s->d_index; J.ncls[ch] = p.s->ncls;
                J.fl[ch] = (p.kind == PK_K2) ? ctx->FL.K[ch][1] : nullptr;
                ej.j[3 * k + i].out = p.out; ej.j[3 * k + i].rep = p.rep; ej.j[3 * k + i].sg = sym_dev(*p.s);
                maxmem = std::max(maxmem, p.s->nmem);
                maxcls = std::max(maxcls, p.s->ncls);
            }
        }
        if (batched) {
            {
                Scope sc(ctx, FDGA_T_MISC);
                LAUNCH(FDGA_T_MISC, repfix_kernel, dim3(nblk(maxcls, 128), nk), 128, fj);
            }
            {
                Scope sc(ctx, FDGA_T_EXPAND);
                const unsigned nbx = (unsigned)std::min<long long>(nblk(maxmem, 256), 2048);
                LAUNCH(FDGA_T_EXPAND, expand_multi_kernel, dim3(nbx, (unsigned)todo.size()), 256, ej);
            }
            CK(cudaGetLastError());
            for (auto& p : todo) if (p.kind == PK_LK2 || p.kind == PK_LK3) { ctx->FL.sw_dirty = true; invalidate_rt(ctx); }
            return 0;
        }
    }
    for (auto& p : todo) {
        {
            Scope sc(ctx, FDGA_T_EXPAND);
            LAUNCH(FDGA_T_EXPAND, expand_kernel, nblk(p.s->nmem, 256), 256, p.out, p.rep, sym_dev(*p.s));
            CK(cudaGetLastError());
        }
        if (post_fix(ctx, p.kind, p.ch)) return 1;
    }
    return 0;
}

// ---- concurrency lanes -------------------------------------------------------------------------------------
// (profiling and the straightforward A/B kernels run serially on the main stream so that their times stay attributable)
// Concurrent lanes pay when the kernels of a stage cannot fill the GPU on their own and the right factors of the three channels
// stay L2 resident together (config 3: 65 MB each, 1.75 ms with lanes vs 2.26 ms without).  For big meshes every kernel fills the
// GPU anyway and three of them at once evict each other's slabs from L2 (nq = 16, 1 GB per right factor: 94 ms with lanes,
// 58 ms without), so the default (FDGA_OPT_SERIAL = 0) decides by the size of one bubble-shaped array.
static bool lanes_enabled(fdga_ctx* ctx) {
    if (ctx->profile || ctx->opt_generic || ctx->opt_serial == 1) return false;
    if (ctx->swave) return ctx->opt_serial == 2 || ctx->opt_serial == 0;      // small kernels: the three channels always overlap
    if (ctx->opt_serial == 2) return true;
    static const double lim_mb = getenv("FDGA_LANES_MAX_MB") ? atof(getenv("FDGA_LANES_MAX_MB")) : 160.0;
    // size of one slab-shaped array as stored (compact: only the slabs this rank reads)
    const double bytes = ctx->slabs_dirty ? (double)ctx->lenPi * sizeof(C) : (double)(2 * ctx->g.nPiF) * ctx->g.NP * std::max(ctx->n_slabs[0], ctx->n_slabs[1]) * sizeof(C);
    return bytes <= lim_mb * 1e6;
}
// everything the lanes read but do not own must be current before the fork
static int lanes_fork(fdga_ctx* ctx, unsigned mom_need = MOM_ALL) {
    if (!lanes_enabled(ctx) || ctx->forked) return 0;
    if (ctx->swave && refresh_swave(ctx)) return 1;
    for (int ch = 0; ch < 3; ch++) if (ensure_pi(ctx, ch)) return 1;
    if (refresh_fsum(ctx) || refresh_k1h(ctx) || refresh_mom_all(ctx, mom_need) || ensure_slabs(ctx)) return 1;
    CK(cudaEventRecord(ctx->ev_fork, ctx->main_stream));
    for (int i = 1; i < 3; i++) CK(cudaStreamWaitEvent(ctx->lane[i], ctx->ev_fork, 0));
    ctx->forked = true;
    return 0;
}
static void lane_use(fdga_ctx* ctx, int i) { if (ctx->forked) { ctx->cur_lane = i; ctx->stream = ctx->lane[i]; } }
static int lanes_join(fdga_ctx* ctx) {
    if (!ctx->forked) return 0;
    ctx->forked = false; ctx->cur_lane = 0; ctx->stream = ctx->main_stream;
    for (int i = 1; i < 3; i++) { CK(cudaEventRecord(ctx->ev_join[i], ctx->lane[i])); CK(cudaStreamWaitEvent(ctx->main_stream, ctx->ev_join[i], 0)); }
    return 0;
}

// BSE_K1! and the fd branch of BSE_K1_1loop! differ in the right factor only (rk_fd = RK_FD resp. RK_1L)
static int bse_K1_impl(fdga_ctx* ctx, int ch, int mfrg, int rk_fd) {
    CK(cudaSetDevice(ctx->device));
    if (ch < 0 || ch > 2) FAIL("fdga_bse_K1: bad channel");
    NEED_SG(FDGA_SG_K1);
    if (ensure_pi(ctx, ch)) return 1;
    DevChain F0 = chain_F(ctx, 1), F = chain_F(ctx, 0), FL = chain_FL(ctx);
    SymGroup& s = ctx->sg[FDGA_SG_K1]; s.d_repvals = s.d_rep[ch];
    long long c0, c1; sg_class_range(ctx, s, c0, c1);
    if (ctx->swave) {       // BSE_K1!(::NL_MF_K1, ...): src/nonlocal/BSEa/BSEa_K1.jl:2-58
        if (rk_fd != RK_FD) FAIL("fdga_bse_K1_1loop: not available for the s-wave (NL) solver");
        if (!ctx->forked && refresh_swave(ctx)) return 1;
        Scope sc(ctx, FDGA_T_K1);
        const double sc1 = ctx->g.T * chsign(ch);
        const C* p0 = ctx->Pisw[pi_kind(ch, true)]; const C* p1 = ctx->Pisw[pi_kind(ch, false)];
        const bool cta = sw_cta_per_rep(2 * ctx->g.nPiF);
        const unsigned nb = cta ? (unsigned)(c1 - c0) : nblk(c1 - c0, FDGA_SW_WARPS);
#define SWK1b(CHT, MFT, CT) LAUNCH(FDGA_T_K1, (sw_bse_k1_kernel<CHT, MFT, CT>), nb, FDGA_SW_THREADS, F0, F, FL, p0, p1, s.d_repvals, sym_dev(s), c0, c1, ctx->g, sc1)
#define SWK1(CHT, MFT) do { if (cta) SWK1b(CHT, MFT, true); else SWK1b(CHT, MFT, false); } while (0)
        if (c1 > c0) {
            if (mfrg) { if (ch == FDGA_PCH) SWK1(CH_P, true); else if (ch == FDGA_TCH) SWK1(CH_T, true); else SWK1(CH_A, true); }
            else      { if (ch == FDGA_PCH) SWK1(CH_P, false); else if (ch == FDGA_TCH) SWK1(CH_T, false); else SWK1(CH_A, false); }
        }
#undef SWK1
#undef SWK1b
        CK(cudaGetLastError());
        return finish_or_defer(ctx, s, ctx->Fbuff.K[ch][0], PK_K1, ch);
    }
    if (cached_right(ctx, ch, mfrg ? RK_MF_K1 : rk_fd, F0, FL)) return 1;
    double scale = ctx->g.T / (double)ctx->g.NP * chsign(ch);
    const DevChain& left = mfrg ? F0 : F;
    {
        Scope sc(ctx, FDGA_T_K1);
        if (c1 > c0) {
            if (ch == FDGA_PCH)      { if (ctx->mbe) LAUNCH(FDGA_T_K1, (bse_k1_kernel<CH_P, true>), (unsigned)(c1 - c0), 256, left, ctx->Rt3[ch], s.d_repvals, sym_dev(s), c0, ctx->g, scale, ctx->d_slabmap[ch == FDGA_PCH ? 0 : 1]); else LAUNCH(FDGA_T_K1, bse_k1_kernel<CH_P>, (unsigned)(c1 - c0), 256, left, ctx->Rt3[ch], s.d_repvals, sym_dev(s), c0, ctx->g, scale, ctx->d_slabmap[ch == FDGA_PCH ? 0 : 1]); }
            else if (ch == FDGA_TCH) { if (ctx->mbe) LAUNCH(FDGA_T_K1, (bse_k1_kernel<CH_T, true>), (unsigned)(c1 - c0), 256, left, ctx->Rt3[ch], s.d_repvals, sym_dev(s), c0, ctx->g, scale, ctx->d_slabmap[ch == FDGA_PCH ? 0 : 1]); else LAUNCH(FDGA_T_K1, bse_k1_kernel<CH_T>, (unsigned)(c1 - c0), 256, left, ctx->Rt3[ch], s.d_repvals, sym_dev(s), c0, ctx->g, scale, ctx->d_slabmap[ch == FDGA_PCH ? 0 : 1]); }
            else                     { if (ctx->mbe) LAUNCH(FDGA_T_K1, (bse_k1_kernel<CH_A, true>), (unsigned)(c1 - c0), 256, left, ctx->Rt3[ch], s.d_repvals, sym_dev(s), c0, ctx->g, scale, ctx->d_slabmap[ch == FDGA_PCH ? 0 : 1]); else LAUNCH(FDGA_T_K1, bse_k1_kernel<CH_A>, (unsigned)(c1 - c0), 256, left, ctx->Rt3[ch], s.d_repvals, sym_dev(s), c0, ctx->g, scale, ctx->d_slabmap[ch == FDGA_PCH ? 0 : 1]); }
        }
        CK(cudaGetLastError());
    }
    return finish_or_defer(ctx, s, ctx->Fbuff.K[ch][0], PK_K1, ch);
}
int fdga_bse_K1(fdga_ctx* ctx, int ch, int mfrg) { return bse_K1_impl(ctx, ch, mfrg, RK_FD); }
// BSE_K1_1loop!: src/nonlocal_2/BSEa/BSE_1loop.jl:2-56 (its mfRG branch is BSE_K1!'s)
int fdga_bse_K1_1loop(fdga_ctx* ctx, int ch, int mfrg) { return bse_K1_impl(ctx, ch, mfrg, RK_1L); }

// BSE_K1_new! / BSE_K2_new!: src/nonlocal_2/BSEa/BSEa_K1.jl:62-113, BSEa_K2.jl:142-216 (term-by-term kernels)
static int bse_new_impl(fdga_ctx* ctx, int ch, int mfrg, int cls) {
    CK(cudaSetDevice(ctx->device));
    if (ch < 0 || ch > 2) FAIL("fdga_bse_K*_new: bad channel");
    int which = cls == 0 ? FDGA_SG_K1 : (ch == FDGA_PCH ? FDGA_SG_PP2 : FDGA_SG_PH2);
    NEED_SG(which);
    if (ctx->opt_local) FAIL("fdga_bse_K*_new: not available for the local solver context");
    if (ctx->swave) FAIL("fdga_bse_K*_new: not available for the s-wave (NL) solver");
    if (ctx->mbe) FAIL("fdga_bse_K*_new: not available for MBE vertices (the term-by-term kernels assume the asymptotic evaluator)");
    if (ensure_pi(ctx, ch)) return 1;
    DevChain F0 = chain_F(ctx, 1), F = chain_F(ctx, 0);
    SymGroup& s = ctx->sg[which]; s.d_repvals = s.d_rep[ch];
    long long c0, c1; sg_class_range(ctx, s, c0, c1);
    const C* p0 = ctx->PiT[pi_kind(ch, true)]; const C* p1 = ctx->PiT[pi_kind(ch, false)];
    C scaleU = bareU(ctx) * (ctx->g.T / (double)ctx->g.NP * chsign(ch));      // bare_vertex(F, Sp) = +U for pSp and dSp
    const int cat = cls == 0 ? FDGA_T_K1 : FDGA_T_K2;
    {
        Scope sc(ctx, cat);
        unsigned nb = (unsigned)(c1 - c0);
        if (c1 > c0) {
#define NEWL(KER, CHT) LAUNCH(cat, KER<CHT>, nb, 256, F, F0, p0, p1, s.d_repvals, sym_dev(s), c0, ctx->g, scaleU, mfrg, ctx->d_slabmap[ch == FDGA_PCH ? 0 : 1])
            if (cls == 0) { if (ch == FDGA_PCH) NEWL(bse_k1_new_kernel, CH_P); else if (ch == FDGA_TCH) NEWL(bse_k1_new_kernel, CH_T); else NEWL(bse_k1_new_kernel, CH_A); }
            else          { if (ch == FDGA_PCH) NEWL(bse_k2_new_kernel, CH_P); else if (ch == FDGA_TCH) NEWL(bse_k2_new_kernel, CH_T); else NEWL(bse_k2_new_kernel, CH_A); }
#undef NEWL
        }
        CK(cudaGetLastError());
    }
    // post-fix: only the d -> p spin fix of the t channel (BSE_templates.jl:209-215, 244-250); no FL add
    return finish_or_defer(ctx, s, ctx->Fbuff.K[ch][cls], cls == 0 ? PK_K1 : PK_K2_NOFL, ch);
}
int fdga_bse_K1_new(fdga_ctx* ctx, int ch, int mfrg) { return bse_new_impl(ctx, ch, mfrg, 0); }
int fdga_bse_K2_new(fdga_ctx* ctx, int ch, int mfrg) { return bse_new_impl(ctx, ch, mfrg, 1); }

int fdga_bse_L_K2(fdga_ctx* ctx, int ch) {
    CK(cudaSetDevice(ctx->device));
    if (ch < 0 || ch > 2) FAIL("fdga_bse_L_K2: bad channel");
    int which = ch == FDGA_PCH ? FDGA_SG_PP2 : FDGA_SG_PH2;
    NEED_SG(which);
    if (ensure_pi(ctx, ch)) return 1;
    DevChain F0 = chain_F(ctx, 1), F = chain_F(ctx, 0), FL = chain_FL(ctx);
    SymGroup& s = ctx->sg[which]; s.d_repvals = s.d_rep[ch];
    long long c0, c1; sg_class_range(ctx, s, c0, c1);
    double scale = ctx->g.T / (double)ctx->g.NP * chsign(ch);
    if (ctx->swave) {       // BSE_L_K2!(::NL_MF_K2, ...): src/nonlocal/BSEa/BSEa_K2.jl:1-41
        if (!ctx->forked && refresh_swave(ctx)) return 1;
        Scope sc(ctx, FDGA_T_L_K2);
        const double sc1 = ctx->g.T * chsign(ch);
        const C* p0 = ctx->Pisw[pi_kind(ch, true)];
        const bool cta = sw_cta_per_rep(2 * ctx->g.nK2f);
        const unsigned nb = cta ? (unsigned)(c1 - c0) : nblk(c1 - c0, FDGA_SW_WARPS);
#define SWLK2b(CHT, CT) LAUNCH(FDGA_T_L_K2, (sw_bse_lk2_kernel<CHT, CT>), nb, FDGA_SW_THREADS, F0, F, p0, s.d_repvals, sym_dev(s), c0, c1, ctx->g, sc1)
#define SWLK2(CHT) do { if (cta) SWLK2b(CHT, true); else SWLK2b(CHT, false); } while (0)
        if (c1 > c0) { if (ch == FDGA_PCH) SWLK2(CH_P); else if (ch == FDGA_TCH) SWLK2(CH_T); else SWLK2(CH_A); }
#undef SWLK2
#undef SWLK2b
        CK(cudaGetLastError());
        return finish_or_defer(ctx, s, ctx->FL.K[ch][1], PK_LK2, ch);
    }
    if (ctx->opt_local) {       // local solver: omega over the bubble mesh, crossing on the right vertex
        if (launch_right<RK_LK2_LOC>(ctx, ch, F0, FL, ctx->g.nK2b, ctx->g.nPiF)) return 1;
        if (ctx->mbe) {      // generic per-term form
            Scope sc(ctx, FDGA_T_L_K2);
            const int* map = ctx->d_slabmap[ch == FDGA_PCH ? 2 : 3];
            if (c1 > c0) {
                if (ch == FDGA_PCH)      LAUNCH(FDGA_T_L_K2, bse_lk2_loc_kernel<CH_P>, (unsigned)(c1 - c0), 128, F, ctx->RtL[ctx->cur_lane], s.d_repvals, sym_dev(s), c0, ctx->g, scale, map);
                else if (ch == FDGA_TCH) LAUNCH(FDGA_T_L_K2, bse_lk2_loc_kernel<CH_T>, (unsigned)(c1 - c0), 128, F, ctx->RtL[ctx->cur_lane], s.d_repvals, sym_dev(s), c0, ctx->g, scale, map);
                else                     LAUNCH(FDGA_T_L_K2, bse_lk2_loc_kernel<CH_A>, (unsigned)(c1 - c0), 128, F, ctx->RtL[ctx->cur_lane], s.d_repvals, sym_dev(s), c0, ctx->g, scale, map);
            }
            CK(cudaGetLastError());
            return finish_or_defer(ctx, s, ctx->FL.K[ch][1], PK_LK2, ch);
        }
        ColJob job = make_job(ctx, 0, 2 * ctx->g.nPiF, ctx->g.nPiF, ctx->g.nK2b, mkC(scale, 0.0));
        if (launch_column<JOB_LK2_LOC>(ctx, ch, F, job, s, ctx->RtL[ctx->cur_lane], FDGA_T_L_K2)) return 1;
        return finish_or_defer(ctx, s, ctx->FL.K[ch][1], PK_LK2, ch);
    }
    if (launch_right<RK_LK2>(ctx, ch, F0, FL, ctx->g.nK2b, ctx->g.nK2f)) return 1;
    if (!ctx->opt_generic) {
        ColJob job = make_job(ctx, 0, 2 * ctx->g.nK2f, ctx->g.nK2f, ctx->g.nK2b, mkC(scale, 0.0));
        if (launch_column<JOB_LK2>(ctx, ch, F, job, s, ctx->RtL[ctx->cur_lane], FDGA_T_L_K2)) return 1;
    } else {
        Scope sc(ctx, FDGA_T_L_K2);
        if (c1 > c0) {
            if (ch == FDGA_PCH)      { if (ctx->mbe) LAUNCH(FDGA_T_L_K2, (bse_lk2_kernel<CH_P, true>), (unsigned)(c1 - c0), 128, F, ctx->Rt, s.d_repvals, sym_dev(s), c0, ctx->g, scale, ctx->d_slabmap[ch == FDGA_PCH ? 2 : 3]); else LAUNCH(FDGA_T_L_K2, bse_lk2_kernel<CH_P>, (unsigned)(c1 - c0), 128, F, ctx->Rt, s.d_repvals, sym_dev(s), c0, ctx->g, scale, ctx->d_slabmap[ch == FDGA_PCH ? 2 : 3]); }
            else if (ch == FDGA_TCH) { if (ctx->mbe) LAUNCH(FDGA_T_L_K2, (bse_lk2_kernel<CH_T, true>), (unsigned)(c1 - c0), 128, F, ctx->Rt, s.d_repvals, sym_dev(s), c0, ctx->g, scale, ctx->d_slabmap[ch == FDGA_PCH ? 2 : 3]); else LAUNCH(FDGA_T_L_K2, bse_lk2_kernel<CH_T>, (unsigned)(c1 - c0), 128, F, ctx->Rt, s.d_repvals, sym_dev(s), c0, ctx->g, scale, ctx->d_slabmap[ch == FDGA_PCH ? 2 : 3]); }
            else                     { if (ctx->mbe) LAUNCH(FDGA_T_L_K2, (bse_lk2_kernel<CH_A, true>), (unsigned)(c1 - c0), 128, F, ctx->Rt, s.d_repvals, sym_dev(s), c0, ctx->g, scale, ctx->d_slabmap[ch == FDGA_PCH ? 2 : 3]); else LAUNCH(FDGA_T_L_K2, bse_lk2_kernel<CH_A>, (unsigned)(c1 - c0), 128, F, ctx->Rt, s.d_repvals, sym_dev(s), c0, ctx->g, scale, ctx->d_slabmap[ch == FDGA_PCH ? 2 : 3]); }
        }
        CK(cudaGetLastError());
    }
    return finish_or_defer(ctx, s, ctx->FL.K[ch][1], PK_LK2, ch);
}

// BSE_K2! and the fd branch of BSE_K2_1loop! differ in the right factor only (rk_fd = RK_FD resp. RK_1L)
static int bse_K2_impl(fdga_ctx* ctx, int ch, int mfrg, int rk_fd) {
    CK(cudaSetDevice(ctx->device));
    if (ch < 0 || ch > 2) FAIL("fdga_bse_K2: bad channel");
    int which = ch == FDGA_PCH ? FDGA_SG_PP2 : FDGA_SG_PH2;
    NEED_SG(which);
    if (ensure_pi(ctx, ch)) return 1;
    DevChain F0 = chain_F(ctx, 1), F = chain_F(ctx, 0), FL = chain_FL(ctx);
    if (ctx->swave) {       // BSE_K2!(::NL_MF_K2, ...): src/nonlocal/BSEa/BSEa_K2.jl:44-106
        if (rk_fd != RK_FD) FAIL("fdga_bse_K2_1loop: not available for the s-wave (NL) solver");
        if (!ctx->forked && refresh_swave(ctx)) return 1;
        SymGroup& s = ctx->sg[which]; s.d_repvals = s.d_rep[ch];
        long long c0, c1; sg_class_range(ctx, s, c0, c1);
        {
            Scope sc(ctx, FDGA_T_K2);
            const double sc1 = ctx->g.T * chsign(ch);
            const C* p0 = ctx->Pisw[pi_kind(ch, true)]; const C* p1 = ctx->Pisw[pi_kind(ch, false)];
            const bool cta = sw_cta_per_rep(2 * ctx->g.nPiF);
            const unsigned nb = cta ? (unsigned)(c1 - c0) : nblk(c1 - c0, FDGA_SW_WARPS);
#define SWK2b(CHT, MFT, CT) LAUNCH(FDGA_T_K2, (sw_bse_k2_kernel<CHT, MFT, CT>), nb, FDGA_SW_THREADS, F0, F, FL, p0, p1, s.d_repvals, sym_dev(s), c0, c1, ctx->g, sc1)
#define SWK2(CHT, MFT) do { if (cta) SWK2b(CHT, MFT, true); else SWK2b(CHT, MFT, false); } while (0)
            if (c1 > c0) {
                if (mfrg) { if (ch == FDGA_PCH) SWK2(CH_P, true); else if (ch == FDGA_TCH) SWK2(CH_T, true); else SWK2(CH_A, true); }
                else      { if (ch == FDGA_PCH) SWK2(CH_P, false); else if (ch == FDGA_TCH) SWK2(CH_T, false); else SWK2(CH_A, false); }
            }
#undef SWK2
#undef SWK2b
            CK(cudaGetLastError());
        }
        return finish_or_defer(ctx, s, ctx->Fbuff.K[ch][1], PK_K2, ch);
    }
    if (!ctx->opt_generic) { if (cached_right(ctx, ch, mfrg ? RK_MF_K2 : rk_fd, F0, FL)) return 1; }
    else if (mfrg) { if (launch_right<RK_MF_K2>(ctx, ch, F0, FL, ctx->g.nK2b, ctx->g.nPiF)) return 1; }
    else if (rk_fd == RK_1L) { if (launch_right<RK_1L>(ctx, ch, F0, FL, ctx->g.nK2b, ctx->g.nPiF)) return 1; }
    else           { if (launch_right<RK_FD>(ctx, ch, F0, FL, ctx->g.nK2b, ctx->g.nPiF)) return 1; }
    SymGroup& s = ctx->sg[which]; s.d_repvals = s.d_rep[ch];
    long long c0, c1; sg_class_range(ctx, s, c0, c1);
    double scale = ctx->g.T / (double)ctx->g.NP * chsign(ch);
    if (!ctx->opt_generic) {
        ColJob job = make_job(ctx, mfrg ? 1 : 0, 2 * ctx->g.nPiF, ctx->g.nPiF, ctx->g.nPiB, mkC(scale, 0.0));
        if (mfrg) { if (launch_column<JOB_K2_MF>(ctx, ch, F, job, s, ctx->Rt3[ch], FDGA_T_K2)) return 1; }
        else if (ctx->has_fsum) {
            if (refresh_fsum(ctx)) return 1;
            job.n_nl2 = ctx->n_nl2 - 1;
            if (launch_column<JOB_K2>(ctx, ch, chain_F_merged(ctx), job, s, ctx->Rt3[ch], FDGA_T_K2)) return 1;
        }
        else      { if (launch_column<JOB_K2>(ctx, ch, F, job, s, ctx->Rt3[ch], FDGA_T_K2)) return 1; }
    } else {
        Scope sc(ctx, FDGA_T_K2);
        unsigned nb = (unsigned)(c1 - c0);
        if (c1 > c0) {
            if (mfrg) {
                if (ch == FDGA_PCH)      { if (ctx->mbe) LAUNCH(FDGA_T_K2, (bse_k2_kernel<CH_P, true, true>), nb, 256, F0, ctx->Rt, s.d_repvals, sym_dev(s), c0, ctx->g, scale, ctx->d_slabmap[ch == FDGA_PCH ? 2 : 3]); else LAUNCH(FDGA_T_K2, (bse_k2_kernel<CH_P, true>), nb, 256, F0, ctx->Rt, s.d_repvals, sym_dev(s), c0, ctx->g, scale, ctx->d_slabmap[ch == FDGA_PCH ? 2 : 3]); }
                else if (ch == FDGA_TCH) { if (ctx->mbe) LAUNCH(FDGA_T_K2, (bse_k2_kernel<CH_T, true, true>), nb, 256, F0, ctx->Rt, s.d_repvals, sym_dev(s), c0, ctx->g, scale, ctx->d_slabmap[ch == FDGA_PCH ? 2 : 3]); else LAUNCH(FDGA_T_K2, (bse_k2_kernel<CH_T, true>), nb, 256, F0, ctx->Rt, s.d_repvals, sym_dev(s), c0, ctx->g, scale, ctx->d_slabmap[ch == FDGA_PCH ? 2 : 3]); }
                else                     { if (ctx->mbe) LAUNCH(FDGA_T_K2, (bse_k2_kernel<CH_A, true, true>), nb, 256, F0, ctx->Rt, s.d_repvals, sym_dev(s), c0, ctx->g, scale, ctx->d_slabmap[ch == FDGA_PCH ? 2 : 3]); else LAUNCH(FDGA_T_K2, (bse_k2_kernel<CH_A, true>), nb, 256, F0, ctx->Rt, s.d_repvals, sym_dev(s), c0, ctx->g, scale, ctx->d_slabmap[ch == FDGA_PCH ? 2 : 3]); }
            } else {
                if (ch == FDGA_PCH)      { if (ctx->mbe) LAUNCH(FDGA_T_K2, (bse_k2_kernel<CH_P, false, true>), nb, 256, F, ctx->Rt, s.d_repvals, sym_dev(s), c0, ctx->g, scale, ctx->d_slabmap[ch == FDGA_PCH ? 2 : 3]); else LAUNCH(FDGA_T_K2, (bse_k2_kernel<CH_P, false>), nb, 256, F, ctx->Rt, s.d_repvals, sym_dev(s), c0, ctx->g, scale, ctx->d_slabmap[ch == FDGA_PCH ? 2 : 3]); }
                else if (ch == FDGA_TCH) { if (ctx->mbe) LAUNCH(FDGA_T_K2, (bse_k2_kernel<CH_T, false, true>), nb, 256, F, ctx->Rt, s.d_repvals, sym_dev(s), c0, ctx->g, scale, ctx->d_slabmap[ch == FDGA_PCH ? 2 : 3]); else LAUNCH(FDGA_T_K2, (bse_k2_kernel<CH_T, false>), nb, 256, F, ctx->Rt, s.d_repvals, sym_dev(s), c0, ctx->g, scale, ctx->d_slabmap[ch == FDGA_PCH ? 2 : 3]); }
                else                     { if (ctx->mbe) LAUNCH(FDGA_T_K2, (bse_k2_kernel<CH_A, false, true>), nb, 256, F, ctx->Rt, s.d_repvals, sym_dev(s), c0, ctx->g, scale, ctx->d_slabmap[ch == FDGA_PCH ? 2 : 3]); else LAUNCH(FDGA_T_K2, (bse_k2_kernel<CH_A, false>), nb, 256, F, ctx->Rt, s.d_repvals, sym_dev(s), c0, ctx->g, scale, ctx->d_slabmap[ch == FDGA_PCH ? 2 : 3]); }
            }
        }
        CK(cudaGetLastError());
    }
    return finish_or_defer(ctx, s, ctx->Fbuff.K[ch][1], PK_K2, ch);
}
int fdga_bse_K2(fdga_ctx* ctx, int ch, int mfrg) { return bse_K2_impl(ctx, ch, mfrg, RK_FD); }
// BSE_K2_1loop!: src/nonlocal_2/BSEa/BSE_1loop.jl:59-124 (its mfRG branch and its FL.K2 post-add are BSE_K2!'s)
int fdga_bse_K2_1loop(fdga_ctx* ctx, int ch, int mfrg) { return bse_K2_impl(ctx, ch, mfrg, RK_1L); }

int fdga_bse_L_K3(fdga_ctx* ctx, int ch) {
    CK(cudaSetDevice(ctx->device));
    if (ch < 0 || ch > 2) FAIL("fdga_bse_L_K3: bad channel");
    int which = ch == FDGA_PCH ? FDGA_SG_PPL3 : FDGA_SG_PHL3;
    NEED_SG(which);
    if (ensure_pi(ctx, ch)) return 1;
    SymGroup& s = ctx->sg[which]; s.d_repvals = s.d_rep[ch];
    long long c0, c1; sg_class_range(ctx, s, c0, c1);
    // (cache_G, cache_F0, sign): BSE_templates.jl:122,130,138
    int cg = ch == FDGA_ACH ? FDGA_C_GA : (ch == FDGA_PCH ? FDGA_C_GPP : FDGA_C_GT);
    int cf0 = ch == FDGA_ACH ? FDGA_C_F0A : (ch == FDGA_PCH ? FDGA_C_F0P : FDGA_C_F0T);
    double sign = ch == FDGA_ACH ? 1.0 : -1.0;
    {
        Scope sc(ctx, FDGA_T_L_K3);
        if (c1 > c0) LAUNCH(FDGA_T_L_K3, bse_lk3_kernel, nblk(c1 - c0, 64), 64, ctx->cache[cg], ctx->cache[cf0], ctx->Pisw[pi_kind(ch, true)], s.d_repvals, sym_dev(s), c0, c1, ctx->g, ctx->g.T * sign);
        CK(cudaGetLastError());
    }
    return finish_or_defer(ctx, s, ctx->FL.K[ch][2], PK_LK3, ch);
}

static int bse_K3_impl(fdga_ctx* ctx, int ch, int mfrg, bool oneloop) {
    CK(cudaSetDevice(ctx->device));
    if (ch < 0 || ch > 2) FAIL("fdga_bse_K3: bad channel");
    if (oneloop && ctx->swave) FAIL("fdga_bse_K3_1loop: not available for the s-wave (NL) solver");
    int which = ch == FDGA_PCH ? FDGA_SG_PP3 : FDGA_SG_PH3;
    NEED_SG(which);
    if (ensure_pi(ctx, ch)) return 1;
    SymGroup& s = ctx->sg[which]; s.d_repvals = s.d_rep[ch];
    long long c0, c1; sg_class_range(ctx, s, c0, c1);
    // (cache_G, cache_F, cache_F0, sign1, sign2): BSE_templates.jl:156,164,172
    int cg = ch == FDGA_ACH ? FDGA_C_GA : (ch == FDGA_PCH ? FDGA_C_GPX : FDGA_C_GT);
    int cf = ch == FDGA_ACH ? FDGA_C_FA : (ch == FDGA_PCH ? FDGA_C_FP : FDGA_C_FT);
    int cf0 = ch == FDGA_ACH ? FDGA_C_F0A : (ch == FDGA_PCH ? FDGA_C_F0P : FDGA_C_F0T);
    double s1 = ch == FDGA_ACH ? 1.0 : -1.0, s2 = ch == FDGA_TCH ? -1.0 : 1.0;
    const C* Pi0sw = ctx->Pisw[pi_kind(ch, true)]; const C* Pisw = ctx->Pisw[pi_kind(ch, false)];
    const C *FLo = ctx->FL.K[ch][2], *FLt = ctx->FL.K[FDGA_TCH][2], *FLa = ctx->FL.K[FDGA_ACH][2];
    {
        Scope sc(ctx, FDGA_T_K3);
        unsigned nb = nblk(c1 - c0, 64);
#define K3L(CHT, MFT, OL) LAUNCH(FDGA_T_K3, (bse_k3_kernel<CHT, MFT, OL>), nb, 64, FLo, FLt, FLa, ctx->cache[cg], ctx->cache[cf], ctx->cache[cf0], Pi0sw, Pisw, s.d_repvals, sym_dev(s), c0, c1, ctx->g, s1, s2)
        if (c1 > c0 && !oneloop) {
            if (mfrg) { if (ch == FDGA_PCH) K3L(CH_P, true, false); else if (ch == FDGA_TCH) K3L(CH_T, true, false); else K3L(CH_A, true, false); }
            else      { if (ch == FDGA_PCH) K3L(CH_P, false, false); else if (ch == FDGA_TCH) K3L(CH_T, false, false); else K3L(CH_A, false, false); }
        } else if (c1 > c0) {
            if (mfrg) { if (ch == FDGA_PCH) K3L(CH_P, true, true); else if (ch == FDGA_TCH) K3L(CH_T, true, true); else K3L(CH_A, true, true); }
            else      { if (ch == FDGA_PCH) K3L(CH_P, false, true); else if (ch == FDGA_TCH) K3L(CH_T, false, true); else K3L(CH_A, false, true); }
        }
#undef K3L
        CK(cudaGetLastError());
    }
    return finish_or_defer(ctx, s, ctx->Fbuff.K[ch][2], PK_K3, ch);
}
int fdga_bse_K3(fdga_ctx* ctx, int ch, int mfrg) { return bse_K3_impl(ctx, ch, mfrg, false); }
// BSE_K3_1loop!: src/nonlocal_2/BSEa/BSE_1loop.jl:123-199
int fdga_bse_K3_1loop(fdga_ctx* ctx, int ch, int mfrg) { return bse_K3_impl(ctx, ch, mfrg, true); }

int fdga_set_F_from_Fbuff(fdga_ctx* ctx) {
    CK(cudaSetDevice(ctx->device));
    if (wait_copy(ctx)) return 1;
    CK(cudaMemcpyAsync(ctx->lev[0].block, ctx->Fbuff.block, ctx->Fbuff.blocklen * sizeof(C), cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->lev[0].sw_dirty = true; ctx->lev[0].k1h_dirty = true; ctx->lev[0].mom_valid[0] = ctx->lev[0].mom_valid[1] = ctx->lev[0].mom_valid[2] = 0; ctx->fsum_dirty = true;
    return 0;
}

// ---- SDE -------------------------------------------------------------------------------------------------
// SDE!(Sigma, G, ..., F = chain[from..]): src/SDE.jl:35-48 with SDE_compute! (src/nonlocal_2/SDE.jl:154-324) per level.
// The real-space contraction, the back transform and SG_Sigma are linear in (Lpp, Lph), so the L arrays of all levels
// of the F0 chain are accumulated first (weight 1/3 for the RefVertex level, SDE.jl:305-309) and transformed ONCE:
// identical to the reference's per-level sum up to rounding.  acc += sgn * result.
static int sde_chain(fdga_ctx* ctx, C* acc, double sgn, int gwhich, bool reference, int from, bool include_U2, bool include_Hartree, bool L_only = false) {
    NEED_SG(FDGA_SG_SIGMA); NEED_SG(FDGA_SG_PP2); NEED_SG(FDGA_SG_PH2);
    const Grid& g = ctx->g;
    if (refresh_pi(ctx, reference ? FDGA_PI0PP : FDGA_PIPP) || refresh_pi(ctx, reference ? FDGA_PI0PH : FDGA_PIPH)) return 1;
    DevChain V = chain_F(ctx, 0);
    C U = bareU(ctx);
    double scale = g.T / (double)g.NP;
    size_t nK2 = ctx->lev[0].len[1];
    CK(cudaMemsetAsync(ctx->L[0], 0, nK2 * sizeof(C), ctx->stream));
    CK(cudaMemsetAsync(ctx->L[1], 0, nK2 * sizeof(C), ctx->stream));
    if (ctx->swave) {       // SDE_channel_L_pp! / ph!(::NL_MF_K2, ...), src/nonlocal/SDE.jl:3-146; every level of the chain in one pass
        if (refresh_swave(ctx)) return 1;
        if (lanes_fork(ctx)) return 1;
        for (int pp = 1; pp >= 0; pp--) {
            lane_use(ctx, pp ? 0 : 1);
            SymGroup& s = ctx->sg[pp ? FDGA_SG_PP2 : FDGA_SG_PH2];
            long long c0, c1; sg_class_range(ctx, s, c0, c1);
            const C* Pi = ctx->Pisw[pp ? (reference ? FDGA_PI0PP : FDGA_PIPP) : (reference ? FDGA_PI0PH : FDGA_PIPH)];
            {
                Scope sc(ctx, FDGA_T_SDE_L);
                const bool cta = sw_cta_per_rep(2 * g.nPiF);
                const unsigned nb = cta ? (unsigned)(c1 - c0) : nblk(c1 - c0, FDGA_SW_WARPS);
#define SWSDE(PPT, CT) LAUNCH(FDGA_T_SDE_L, (sw_sde_L_kernel<PPT, CT>), nb, FDGA_SW_THREADS, V, from, Pi, s.d_repvals, sym_dev(s), c0, c1, g, U, g.T)
                if (c1 > c0) {
                    if (pp) { if (cta) SWSDE(true, true); else SWSDE(true, false); }
                    else    { if (cta) SWSDE(false, true); else SWSDE(false, false); }
                }
#undef SWSDE
                CK(cudaGetLastError());
            }
            if (sg_finish(ctx, s, ctx->L[pp ? 0 : 1])) return 1;
        }
    } else
    if (!ctx->opt_generic) {
        // fused recursion: one column launch per bubble kind covers every level of the chain (fdga_column.cuh);
        // lanes: pp on 0, ph on 1, (G transforms + U^2 term) on 2
        if (lanes_fork(ctx, ctx->opt_sde_own_gamma ? 0u : (MOM_ALL & ~MOM_FSUM) & ~((2u << from) - 1u))) return 1;      // cross channels of the levels > from
        for (int pp = 1; pp >= 0; pp--) {
            lane_use(ctx, pp ? 0 : 1);
            SymGroup& s = ctx->sg[pp ? FDGA_SG_PP2 : FDGA_SG_PH2];
            const C* PiT = ctx->PiT[pp ? (reference ? FDGA_PI0PP : FDGA_PIPP) : (reference ? FDGA_PI0PH : FDGA_PIPH)];
            ColJob job = make_job(ctx, from, 2 * g.nPiF, g.nPiF, g.nPiB, U * scale);
            if (pp) { if (launch_column_t<JOB_SDE_PP, CH_P>(ctx, V, job, s, PiT, FDGA_T_SDE_L)) return 1; }
            else    { if (launch_column_t<JOB_SDE_PH, CH_A>(ctx, V, job, s, PiT, FDGA_T_SDE_L)) return 1; }
            if (sg_finish(ctx, s, ctx->L[pp ? 0 : 1])) return 1;
        }
    } else
    for (int level = from; level < ctx->nlev; level++) {
        double wl = (ctx->lev[level].d.type == FDGA_LV_CORE) ? 1.0 / 3.0 : 1.0;
        for (int pp = 1; pp >= 0; pp--) {
            SymGroup& s = ctx->sg[pp ? FDGA_SG_PP2 : FDGA_SG_PH2];
            long long c0, c1; sg_class_range(ctx, s, c0, c1);
            const C* PiT = ctx->PiT[pp ? (reference ? FDGA_PI0PP : FDGA_PIPP) : (reference ? FDGA_PI0PH : FDGA_PIPH)];
            {
                Scope sc(ctx, FDGA_T_SDE_L);
                if (c1 > c0) {
                    if (pp) { if (ctx->mbe) LAUNCH(FDGA_T_SDE_L, (sde_L_kernel<true, true>), (unsigned)(c1 - c0), 256, V, level, PiT, s.d_repvals, sym_dev(s), c0, g, U, scale, ctx->opt_sde_own_gamma, ctx->d_slabmap[0]); else LAUNCH(FDGA_T_SDE_L, sde_L_kernel<true>, (unsigned)(c1 - c0), 256, V, level, PiT, s.d_repvals, sym_dev(s), c0, g, U, scale, ctx->opt_sde_own_gamma, ctx->d_slabmap[0]); }
                    else    { if (ctx->mbe) LAUNCH(FDGA_T_SDE_L, (sde_L_kernel<false, true>), (unsigned)(c1 - c0), 256, V, level, PiT, s.d_repvals, sym_dev(s), c0, g, U, scale, ctx->opt_sde_own_gamma, ctx->d_slabmap[1]); else LAUNCH(FDGA_T_SDE_L, sde_L_kernel<false>, (unsigned)(c1 - c0), 256, V, level, PiT, s.d_repvals, sym_dev(s), c0, g, U, scale, ctx->opt_sde_own_gamma, ctx->d_slabmap[1]); }
                }
                CK(cudaGetLastError());
            }
            if (ctx->nranks > 1) {
                Scope sc(ctx, FDGA_T_COMM);
                int rc = ctx->nccl.AllGather(s.d_repvals + (size_t)ctx->rank * s.chunk, s.d_repvals, (size_t)s.chunk * 2, /*ncclDouble*/ 8, ctx->comm, ctx->stream);
                if (rc != 0) FAIL(std::string("ncclAllGather: ") + ctx->nccl.GetErrorString(rc));
                ctx->n_launch[FDGA_T_COMM]++;
            }
            Scope sc(ctx, FDGA_T_EXPAND);
            LAUNCH(FDGA_T_EXPAND, expand_add_kernel, nblk(s.nmem, 256), 256, ctx->L[pp ? 0 : 1], s.d_repvals, sym_dev(s), wl);
            CK(cudaGetLastError());
        }
    }
    if (L_only) return lanes_join(ctx);
    C* Sout = ctx->SigAcc;
    const long long pre = (long long)(2 * g.nK2b - 1) * (2 * g.nK2f);
    const double nrm = 1.0 / ((double)g.L * g.L * g.L * g.L);
    // L arrays are transformed in place (clobbered, as in the reference: SURVEY E4)
    for (int pp = 1; pp >= 0; pp--) {
        lane_use(ctx, pp ? 0 : 1);
        Scope sc(ctx, FDGA_T_SDE_RS);
        if (ctx->swave) {       // fft(L, momentum axis) / L^2 -> scratchA / scratchB (src/nonlocal/SDE.jl:217-218); L[] is the intermediate
            C* Lx = ctx->L[pp ? 0 : 1]; C* out = pp ? ctx->scratchA : ctx->scratchB;
            const long long n = pre * g.NP;
            if (dft2_tile(ctx, Lx, Lx, pre, g.L, 1, -1, 1.0 / ((double)g.L * g.L), ctx->twL, FDGA_T_SDE_RS)) { CK(cudaGetLastError()); continue; }
            LAUNCH(FDGA_T_SDE_RS, dft_axis_kernel, nblk(n, 128), 128, Lx, out, pre, g.L, (long long)g.L, -1, 1.0, ctx->twL);
            LAUNCH(FDGA_T_SDE_RS, dft_axis_kernel, nblk(n, 128), 128, out, Lx, pre * g.L, g.L, 1LL, -1, 1.0 / ((double)g.L * g.L), ctx->twL);
            CK(cudaGetLastError());
            continue;
        }
        if (dft4(ctx, ctx->L[pp ? 0 : 1], pp ? ctx->scratchA : ctx->scratchB, pre, -1, nrm, FDGA_T_SDE_RS)) return 1;
    }
    lane_use(ctx, 2);
    {
        Scope sc(ctx, FDGA_T_SDE_RS);
        if (dft2_G(ctx, ctx->G[gwhich], ctx->GR, ctx->SigTmp, -1, 1.0 / ((double)g.LG * g.LG), FDGA_T_SDE_RS)) return 1;
    }
    if (include_U2) {
        Scope sc(ctx, FDGA_T_SDE_U2);
        double inv = 1.0 / ((double)g.LG * g.LG);
        // GR already holds fft(G)/LG^2 ; GRm = bfft(G)/LG^2
        if (dft2_G(ctx, ctx->G[gwhich], ctx->GRm, ctx->SigTmp, +1, inv, FDGA_T_SDE_U2)) return 1;
        C fac = (U * U) * (g.T * g.T);
        sde_u2_kernel<<<(unsigned)(g.LG * g.LG), 64, (size_t)(8 * g.nG - 1) * sizeof(C), ctx->stream>>>(ctx->GR, ctx->GRm, ctx->SigR2, g.nG, g.LG, fac);
        ctx->n_launch[FDGA_T_SDE_U2]++; ctx->total_launches++;
        if (dft2_G(ctx, ctx->SigR2, ctx->GRm, ctx->SigTmp, +1, 1.0, FDGA_T_SDE_U2)) return 1;
        SymGroup& ss = ctx->sg[FDGA_SG_SIGMA];
        LAUNCH(FDGA_T_SDE_U2, symmetrize_kernel, nblk(ss.nmem, 256), 256, ctx->GRm, sym_dev(ss));
        CK(cudaGetLastError());
    }
    if (include_Hartree) {      // occupation of G for the Hartree term: single-CTA reduction, hidden on lane 2 beside the L transforms
        Scope sc(ctx, FDGA_T_MISC);
        if (occupation_dev(ctx, gwhich)) return 1;
    }
    if (lanes_join(ctx)) return 1;
    {
        Scope sc(ctx, FDGA_T_SDE_RS);
        CK(cudaMemsetAsync(ctx->SigR, 0, ctx->lenG * sizeof(C), ctx->stream));
        const int twin = std::min(g.LG, 4 * (g.L / 2) + 1);
        if (ctx->swave) LAUNCH(FDGA_T_SDE_RS, sw_sde_rs_kernel, nblk((long long)ctx->lenG, 128), 128, ctx->GR, ctx->L[0], ctx->L[1], ctx->SigR, g);
        else
        LAUNCH(FDGA_T_SDE_RS, sde_rs_kernel, nblk((long long)(2 * g.nK2f) * twin * twin * 32, 128), 128, ctx->GR, ctx->L[0], ctx->L[1], ctx->SigR, g, g.nG, g.LG, twin);
        CK(cudaGetLastError());
        if (dft2_G(ctx, ctx->SigR, Sout, ctx->SigTmp, +1, 1.0, FDGA_T_SDE_RS)) return 1;
        SymGroup& ss = ctx->sg[FDGA_SG_SIGMA];
        LAUNCH(FDGA_T_SDE_RS, symmetrize_kernel, nblk(ss.nmem, 256), 256, Sout, sym_dev(ss));
        if (include_U2) LAUNCH(FDGA_T_SDE_RS, add_axpby_kernel, nblk(ctx->lenG, 256), 256, Sout, ctx->GRm, 1.0, (const C*)nullptr, 0.0, (long long)ctx->lenG);
        CK(cudaGetLastError());
    }
    if (include_Hartree) {
        Scope sc(ctx, FDGA_T_MISC);
        LAUNCH(FDGA_T_MISC, hartree_kernel, nblk(ctx->lenG, 256), 256, Sout, ctx->d_occ, U, 1.0, (long long)ctx->lenG);
        CK(cudaGetLastError());
    }
    LAUNCH(FDGA_T_MISC, add_axpby_kernel, nblk(ctx->lenG, 256), 256, acc, Sout, sgn, (const C*)nullptr, 0.0, (long long)ctx->lenG);
    CK(cudaGetLastError());
    return 0;
}
// ---- CUDA graphs ----------------------------------------------------------------------------------------------
// A step of the iteration is ~80 small dependent launches on three lanes; recorded once as a CUDA graph it replays with one
// launch and without the per-launch gaps.  The library keeps derived tables current lazily (dirty flags on the host), so a
// recording is only valid as a steady-state cycle: the flags at its end must equal the flags at its start, and a replay is only
// accepted from that same state (and before any reallocation: `epoch`).
static bool graph_multirank() {
    static const bool mr = getenv("FDGA_GRAPH_MULTIRANK") ? atoi(getenv("FDGA_GRAPH_MULTIRANK")) != 0 : false;
    return mr;
}
static std::vector<long long> state_signature(fdga_ctx* ctx) {
    std::vector<long long> v;
    auto lvl = [&](const LevelBuf& lb) { v.push_back(lb.sw_dirty); v.push_back(lb.k1h_dirty); for (int c = 0; c < 3; c++) v.push_back(lb.mom_valid[c]); };
    for (int l = 0; l < ctx->nlev; l++) lvl(ctx->lev[l]);
    lvl(ctx->FL); lvl(ctx->Fbuff); if (ctx->has_fsum) lvl(ctx->Fsum);
    v.push_back(ctx->fsum_dirty); v.push_back(ctx->slabs_dirty);
    for (int i = 0; i < 4; i++) { v.push_back(ctx->pi_dirty[i]); v.push_back(ctx->pi_src[i]); v.push_back(ctx->pi_full_valid[i]); }
    for (int i = 0; i < 3; i++) v.push_back(ctx->rt_kind[i]);
    v.push_back(ctx->copy_pending); v.push_back(ctx->opt_serial); v.push_back(ctx->profile);
    return v;
}
static void restore_signature(fdga_ctx* ctx, const std::vector<long long>& v) {      // inverse of state_signature
    size_t i = 0;
    auto lvl = [&](LevelBuf& lb) { lb.sw_dirty = v[i++] != 0; lb.k1h_dirty = v[i++] != 0; for (int c = 0; c < 3; c++) lb.mom_valid[c] = (unsigned)v[i++]; };
    for (int l = 0; l < ctx->nlev; l++) lvl(ctx->lev[l]);
    lvl(ctx->FL); lvl(ctx->Fbuff); if (ctx->has_fsum) lvl(ctx->Fsum);
    ctx->fsum_dirty = v[i++] != 0; ctx->slabs_dirty = v[i++] != 0;
    for (int k = 0; k < 4; k++) { ctx->pi_dirty[k] = v[i++] != 0; ctx->pi_src[k] = (int)v[i++]; ctx->pi_full_valid[k] = v[i++] != 0; }
    for (int k = 0; k < 3; k++) ctx->rt_kind[k] = (int)v[i++];
}
int fdga_graph_begin(fdga_ctx* ctx) {
    CK(cudaSetDevice(ctx->device));
    if (ctx->capturing) FAIL("fdga_graph_begin: already recording");
    if (ctx->profile) FAIL("fdga_graph_begin: not while profiling (the per-kernel timers need eager launches)");
    // multi-rank: the NCCL all-gathers / broadcasts are captured with the kernels (every rank must record and replay the same
    // sequence, like any collective); opt-in until it has been exercised on more boxes
    if (ctx->nranks > 1 && !graph_multirank()) FAIL("fdga_graph_begin: single-rank contexts only (set FDGA_GRAPH_MULTIRANK=1 to record NCCL collectives too)");
    if (wait_copy(ctx)) return 1;
    ctx->copy_pending = false;
    ctx->cap_sig = state_signature(ctx);
    ctx->cap_launches0 = ctx->total_launches; memcpy(ctx->cap_n0, ctx->n_launch, sizeof(ctx->cap_n0));
    CK(cudaStreamBeginCapture(ctx->main_stream, cudaStreamCaptureModeRelaxed));
    ctx->capturing = true;
    return 0;
}
int fdga_graph_end(fdga_ctx* ctx, int* graph_id) {
    CK(cudaSetDevice(ctx->device));
    if (!ctx->capturing) FAIL("fdga_graph_end: no recording in progress");
    if (ctx->forked) lanes_join(ctx);
    if (ctx->copy_pending) { cudaStreamWaitEvent(ctx->main_stream, ctx->ev_copy_done, 0); ctx->copy_pending = false; }      // the copy stream rejoins
    ctx->capturing = false;
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(ctx->main_stream, &g);
    // a failed recording executed nothing: the lazy-state flags go back to what they were at fdga_graph_begin
    if (e != cudaSuccess || !g) { restore_signature(ctx, ctx->cap_sig); ctx->err = std::string("fdga_graph_end: cudaStreamEndCapture: ") + cudaGetErrorString(e); cudaGetLastError(); return 1; }
    if (state_signature(ctx) != ctx->cap_sig) { cudaGraphDestroy(g); restore_signature(ctx, ctx->cap_sig); FAIL("fdga_graph_end: the recorded calls do not form a steady-state cycle (lazy tables differ between start and end): run the sequence once eagerly, then record it"); }
    GraphRec r; r.graph = g; r.exec = nullptr; r.live = true; r.sig = ctx->cap_sig; r.epoch = ctx->epoch;
    r.launches = ctx->total_launches - ctx->cap_launches0;
    for (int i = 0; i < FDGA_T_COUNT; i++) r.n_launch[i] = ctx->n_launch[i] - ctx->cap_n0[i];
    e = cudaGraphInstantiate(&r.exec, g, 0);
    if (e != cudaSuccess) { cudaGraphDestroy(g); ctx->err = std::string("fdga_graph_end: cudaGraphInstantiate: ") + cudaGetErrorString(e); return 1; }
    // the recording itself executed nothing: the launch counters go back to their values at fdga_graph_begin
    ctx->total_launches = ctx->cap_launches0; memcpy(ctx->n_launch, ctx->cap_n0, sizeof(ctx->cap_n0));
    ctx->graphs.push_back(r);
    *graph_id = (int)ctx->graphs.size() - 1;
    return 0;
}
int fdga_graph_launch(fdga_ctx* ctx, int graph_id) {
    CK(cudaSetDevice(ctx->device));
    if (graph_id < 0 || graph_id >= (int)ctx->graphs.size() || !ctx->graphs[graph_id].live) FAIL("fdga_graph_launch: bad graph id");
    GraphRec& r = ctx->graphs[graph_id];
    if (ctx->capturing) FAIL("fdga_graph_launch: not inside a recording");
    if (r.epoch != ctx->epoch) FAIL("fdga_graph_launch: stale graph (device tables were rebuilt since it was recorded): record it again");
    if (wait_copy(ctx)) return 1;
    ctx->copy_pending = false;
    if (state_signature(ctx) != r.sig) FAIL("fdga_graph_launch: the context is not in the state the graph was recorded from: record it again");
    CK(cudaGraphLaunch(r.exec, ctx->main_stream));
    ctx->total_launches += r.launches;
    for (int i = 0; i < FDGA_T_COUNT; i++) ctx->n_launch[i] += r.n_launch[i];
    return 0;
}
int fdga_graph_destroy(fdga_ctx* ctx, int graph_id) {
    if (graph_id < 0 || graph_id >= (int)ctx->graphs.size() || !ctx->graphs[graph_id].live) FAIL("fdga_graph_destroy: bad graph id");
    CK(cudaSetDevice(ctx->device)); CK(cudaStreamSynchronize(ctx->main_stream));
    cudaGraphExecDestroy(ctx->graphs[graph_id].exec); cudaGraphDestroy(ctx->graphs[graph_id].graph); ctx->graphs[graph_id].live = false;
    return 0;
}


extern "C++" {
// ---- automatic graphs behind the plain entry points --------------------------------------------------------------------------
// fdga_iterate_solver, fdga_sde and the mfRG matvec are called in loops with the same arguments (fixed-point iteration, Krylov
// solver).  Once such a call is seen twice from the same lazy state it is recorded, and from then on replayed, as a CUDA graph;
// any call from another state (or after a rebuild of device tables) runs eagerly.  FDGA_AUTOGRAPH=0 switches this off.
struct AutoGraph { long long key; std::vector<int> ids; bool have_prev, bad; std::vector<long long> sig_prev; long long epoch; };      // one graph per lazy state a key was seen in (<= 4)
static std::vector<std::pair<fdga_ctx*, std::vector<AutoGraph>>> g_auto;      // per context (contexts are not thread-safe anyway)
static std::vector<AutoGraph>& auto_table(fdga_ctx* ctx) {
    for (auto& p : g_auto) if (p.first == ctx) return p.second;
    g_auto.push_back({ctx, {}});
    return g_auto.back().second;
}
static void auto_forget(fdga_ctx* ctx) {
    for (size_t i = 0; i < g_auto.size(); i++) if (g_auto[i].first == ctx) { g_auto.erase(g_auto.begin() + i); return; }
}
static bool autograph_usable(fdga_ctx* ctx) {
    static const bool on = getenv("FDGA_AUTOGRAPH") ? atoi(getenv("FDGA_AUTOGRAPH")) != 0 : true;
    return on && !ctx->capturing && !ctx->profile && (ctx->nranks == 1 || graph_multirank()) && !ctx->opt_generic;
}
template <class Fn>
static int auto_graphed(fdga_ctx* ctx, long long key, Fn body) {
    if (!autograph_usable(ctx)) return body();
    std::vector<AutoGraph>& tab = auto_table(ctx);
    AutoGraph* a = nullptr;
    for (auto& e : tab) if (e.key == key) a = &e;
    if (!a) { AutoGraph e; e.key = key; e.have_prev = false; e.bad = false; e.epoch = ctx->epoch; tab.push_back(e); a = &tab.back(); }
    if (a->epoch != ctx->epoch) {       // device tables were rebuilt: forget what was recorded
        for (int id : a->ids) fdga_graph_destroy(ctx, id);
        a->ids.clear(); a->have_prev = false; a->bad = false; a->epoch = ctx->epoch;
    }
    if (wait_copy(ctx)) return 1;
    ctx->copy_pending = false;
    const std::vector<long long> sig = state_signature(ctx);
    for (int id : a->ids) if (sig == ctx->graphs[id].sig) return fdga_graph_launch(ctx, id);
    if (a->bad || a->ids.size() >= 4 || !a->have_prev || sig != a->sig_prev) { a->sig_prev = sig; a->have_prev = true; return body(); }
    // second call in a row from this state: record it (nothing executes), then replay
    if (fdga_graph_begin(ctx)) { a->bad = true; return body(); }
    const int rc = body();
    int id = -1;
    const int rc2 = fdga_graph_end(ctx, &id);
    if (rc == 0 && rc2 == 0) { a->ids.push_back(id); return fdga_graph_launch(ctx, id); }
    // not recordable (a host synchronisation inside, or not a steady-state cycle): the recording pass changed host flags without
    // executing anything -- put them back and run the call for real
    a->bad = true;
    ctx->capturing = false; ctx->forked = false; ctx->cur_lane = 0; ctx->stream = ctx->main_stream; ctx->defer = false; ctx->pending.clear(); ctx->cache_on_lane = false;
    cudaGetLastError();
    restore_signature(ctx, sig);
    return body();
}

}  // extern "C++"

int fdga_sde_channel_L(fdga_ctx* ctx, int reference, int from) {
    CK(cudaSetDevice(ctx->device));
    if (from < 0 || from >= ctx->nlev) FAIL("fdga_sde_channel_L: bad level");
    return sde_chain(ctx, nullptr, 0.0, reference ? FDGA_G0 : FDGA_G, reference != 0, from, false, false, true);
}
static int sde_body(fdga_ctx* ctx, int strategy, int include_U2, int include_Hartree);
int fdga_sde(fdga_ctx* ctx, int strategy, int include_U2, int include_Hartree) {
    CK(cudaSetDevice(ctx->device));
    if (strategy < FDGA_SCPA || strategy > FDGA_FDPA_1LOOP) FAIL("fdga_sde: Calculation strategy unknown");      // before S.Sigma is touched (src/SDE.jl:31)
    return auto_graphed(ctx, 2000 + strategy * 4 + (include_U2 ? 2 : 0) + (include_Hartree ? 1 : 0), [&]() { return sde_body(ctx, strategy, include_U2, include_Hartree); });
}
static int sde_body(fdga_ctx* ctx, int strategy, int include_U2, int include_Hartree) {
    C* S = ctx->G[FDGA_SIGMA];
    CK(cudaMemsetAsync(S, 0, ctx->lenG * sizeof(C), ctx->stream));
    if (sde_chain(ctx, S, 1.0, FDGA_G, false, 0, include_U2, include_Hartree)) return 1;
    if (strategy == FDGA_FDPA || strategy == FDGA_FDPA_NEW || strategy == FDGA_FDPA_1LOOP) {      // src/SDE.jl:4,8,13-24
        if (sde_chain(ctx, S, -1.0, FDGA_G0, true, 1, include_U2, include_Hartree)) return 1;
        LAUNCH(FDGA_T_MISC, add_axpby_kernel, nblk(ctx->lenG, 256), 256, S, ctx->G[FDGA_SIGMA0], 1.0, (const C*)nullptr, 0.0, (long long)ctx->lenG);
        if (include_Hartree && !ctx->opt_hartree_once) {
            if (occupation_dev(ctx, FDGA_G0)) return 1;
            LAUNCH(FDGA_T_MISC, hartree_kernel, nblk(ctx->lenG, 256), 256, S, ctx->d_occ, bareU(ctx), -1.0, (long long)ctx->lenG);
        }
        CK(cudaGetLastError());
    }
    return 0;
}

// ---- drivers ---------------------------------------------------------------------------------------------
// the BSE stages of one iteration: [L_K2, L_K3] | [K1, K2, K3].  The three channels of a stage are independent up to
// their post-fixes: each runs on its own lane, and the stage ends with one batched SG finish (one NCCL group).
// with_cache: build_K3_cache! (fd flavour) is issued on the lane of the a channel right after the first fork instead of on the main
// stream before it: only the K3 kernels read the caches, so BSE_L_K2! (or BSE_K1! / BSE_K2! without an L stage) of the other lanes
// start at once and the K3 kernels wait for the cache event
static int cache_on_lane_begin(fdga_ctx* ctx) {
    if (!ctx->forked) return fdga_build_K3_cache(ctx, 0, 0);
    lane_use(ctx, FDGA_ACH);
    if (fdga_build_K3_cache(ctx, 0, 0)) return 1;
    CK(cudaEventRecord(ctx->ev_cache, ctx->stream));
    ctx->cache_on_lane = true;
    return 0;
}
static int cache_on_lane_wait(fdga_ctx* ctx) {      // every lane (or the main stream) waits for the cache before its K3 kernels
    if (!ctx->cache_on_lane) return 0;
    if (ctx->forked) { for (int i = 0; i < 3; i++) CK(cudaStreamWaitEvent(ctx->lane[i], ctx->ev_cache, 0)); }
    else CK(cudaStreamWaitEvent(ctx->main_stream, ctx->ev_cache, 0));
    ctx->cache_on_lane = false;
    return 0;
}
static int bse_stages(fdga_ctx* ctx, bool with_L, int mfrg, bool with_cache = false) {
    // issue order: the t channel carries two spin forms (twice the work), so it goes first on the high-priority lane;
    // the post-fixes, which need p, a, t (BSE_templates.jl:35-38: a before t), are ordered by flush_pending
    const int order[3] = {FDGA_TCH, FDGA_PCH, FDGA_ACH};
    ctx->defer = true;
    int rc = 0;
    if (with_cache && refresh_swave(ctx)) { ctx->defer = false; return 1; }      // the cache kernel's tables, current before the fork
    if (with_L) {
        rc = lanes_fork(ctx, 1u);      // BSE_L_K2!: cross channels of S.F itself
        if (!rc && with_cache) { rc = cache_on_lane_begin(ctx); with_cache = false; }
        for (int i = 0; i < 3 && !rc; i++) { lane_use(ctx, order[i]); rc = fdga_bse_L_K2(ctx, order[i]); }
        if (!rc) rc = cache_on_lane_wait(ctx);
        for (int i = 0; i < 3 && !rc; i++) { lane_use(ctx, order[i]); rc = fdga_bse_L_K3(ctx, order[i]); }   // reads caches and bubbles only
        if (lanes_join(ctx)) rc = 1;
        if (!rc) rc = flush_pending(ctx);
    }
    // BSE_K2!: left vertex S.F + S.F0 (merged level when available), mfRG: S.F0 only
    if (!rc) rc = lanes_fork(ctx, mfrg ? (MOM_ALL & ~MOM_FSUM & ~1u) : (ctx->has_fsum ? (MOM_ALL & ~3u) : (MOM_ALL & ~MOM_FSUM)));
    if (!rc && with_cache) rc = cache_on_lane_begin(ctx);      // no L stage (scPA): the cache is built beside BSE_K1! / BSE_K2!
    for (int i = 0; i < 3 && !rc; i++) { lane_use(ctx, order[i]); rc = fdga_bse_K1(ctx, order[i], mfrg); }
    for (int i = 0; i < 3 && !rc; i++) { lane_use(ctx, order[i]); rc = fdga_bse_K2(ctx, order[i], mfrg); }    // K1 and K2 share inputs (FL, right factor)
    if (!rc) rc = cache_on_lane_wait(ctx);
    // BSE_K3! only reads the caches, the s-wave bubbles and FL.K3 (all final after the L stage): same stage, one SG finish fewer
    for (int i = 0; i < 3 && !rc; i++) { lane_use(ctx, order[i]); rc = fdga_bse_K3(ctx, order[i], mfrg); }
    if (lanes_join(ctx)) rc = 1;
    if (!rc) rc = flush_pending(ctx);
    ctx->defer = false; ctx->pending.clear();
    return rc;
}

// one stage = the same entry point for the three channels on concurrent lanes + one batched SG finish
static int bse_stage(fdga_ctx* ctx, int (*fn)(fdga_ctx*, int, int), int mfrg) {
    const int order[3] = {FDGA_TCH, FDGA_PCH, FDGA_ACH};
    int rc = lanes_fork(ctx);
    for (int i = 0; i < 3 && !rc; i++) { lane_use(ctx, order[i]); rc = fn(ctx, order[i], mfrg); }
    if (lanes_join(ctx)) rc = 1;
    if (!rc) rc = flush_pending(ctx);
    return rc;
}
static int bse_L_K3_stage_fn(fdga_ctx* ctx, int ch, int) { return fdga_bse_L_K3(ctx, ch); }
// the stage lists of the variant strategies, src/solve.jl:26-58
static int bse_stages_variant(fdga_ctx* ctx, int strategy) {
    ctx->defer = true;
    int rc = 0;
    if (strategy == FDGA_FDPA_1LOOP) {
        rc = bse_stage(ctx, fdga_bse_K3_1loop, 0);
        if (!rc) rc = bse_stage(ctx, fdga_bse_K1_1loop, 0);
        if (!rc) rc = bse_stage(ctx, fdga_bse_K2_1loop, 0);
    } else {
        if (strategy == FDGA_FDPA_NEW) rc = bse_stage(ctx, bse_L_K3_stage_fn, 0);
        if (!rc) rc = bse_stage(ctx, fdga_bse_K3, 0);
        if (!rc) rc = bse_stage(ctx, fdga_bse_K1_new, 0);
        if (!rc) rc = bse_stage(ctx, fdga_bse_K2_new, 0);
    }
    ctx->defer = false; ctx->pending.clear();
    return rc;
}

static int iterate_body(fdga_ctx* ctx, int strategy, int update_sigma, int compute_hartree);
int fdga_iterate_solver(fdga_ctx* ctx, int strategy, int update_sigma, int compute_hartree) {
    if (strategy < FDGA_SCPA || strategy > FDGA_FDPA_1LOOP) FAIL("fdga_iterate_solver: Calculation strategy unknown");
    CK(cudaSetDevice(ctx->device));
    return auto_graphed(ctx, 1000 + strategy * 4 + (update_sigma ? 2 : 0) + (compute_hartree ? 1 : 0), [&]() { return iterate_body(ctx, strategy, update_sigma, compute_hartree); });
}
static int iterate_body(fdga_ctx* ctx, int strategy, int update_sigma, int compute_hartree) {
    if (update_sigma) { if (fdga_dyson(ctx) || (ctx->opt_local ? fdga_bubbles_local(ctx, 0) : fdga_bubbles_real_space(ctx, 0))) return 1; }
    static const bool cache_lane = getenv("FDGA_CACHE_ON_LANE") ? atoi(getenv("FDGA_CACHE_ON_LANE")) != 0 : true;
    if (strategy >= FDGA_SCPA_NEW) { if (fdga_build_K3_cache(ctx, 0, 0) || bse_stages_variant(ctx, strategy)) return 1; }
    else if (!cache_lane) { if (fdga_build_K3_cache(ctx, 0, 0) || bse_stages(ctx, strategy == FDGA_FDPA, 0)) return 1; }
    else if (bse_stages(ctx, strategy == FDGA_FDPA, 0, true)) return 1;
    if (fdga_set_F_from_Fbuff(ctx)) return 1;
    if (update_sigma) { if (fdga_sde(ctx, strategy, 1, compute_hartree ? 1 : 0)) return 1; }      // SDE!(S; strategy, include_Hartree = compute_Hartree), src/solve.jl:99
    return 0;
}

// mfRGLinearMap(S, strategy) * x on DEVICE vectors (src/mfRG.jl:34-89): y = x - flatten(BSE_lin(factor * x)) / factor.
// x_dev and y_dev may alias.  strategy: fdPA / fdPA_1loop (identical maps, src/mfRG.jl:51) or fdPA_new.
static int mfrg_matvec_body(fdga_ctx* ctx, const C* x_dev, C* y_dev, int first, int strategy);
static int mfrg_matvec_dev(fdga_ctx* ctx, const C* x_dev, C* y_dev, int first, int strategy) {
    if (strategy != FDGA_FDPA && strategy != FDGA_FDPA_NEW && strategy != FDGA_FDPA_1LOOP)
        FAIL("mfRGLinearMap: Invalid strategy. Must be fdPA or fdPA_new or fdPA_1loop.");      // src/mfRG.jl:26-28
    if (!autograph_usable(ctx)) return mfrg_matvec_body(ctx, x_dev, y_dev, first, strategy);
    // the graph works on fixed buffers (flat2 -> flat); Krylov vectors are copied in / out on the device
    if (x_dev != ctx->flat2) CK(cudaMemcpyAsync(ctx->flat2, x_dev, ctx->lenFlat * sizeof(C), cudaMemcpyDeviceToDevice, ctx->stream));
    if (auto_graphed(ctx, 3000 + strategy * 2 + (first ? 1 : 0), [&]() { return mfrg_matvec_body(ctx, ctx->flat2, ctx->flat, first, strategy); })) return 1;
    if (y_dev != ctx->flat) CK(cudaMemcpyAsync(y_dev, ctx->flat, ctx->lenFlat * sizeof(C), cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
}
static int mfrg_matvec_body(fdga_ctx* ctx, const C* x_dev, C* y_dev, int first, int strategy) {
    const double factor = 1e-2;                                  // src/mfRG.jl:37
    if (strategy != FDGA_FDPA && strategy != FDGA_FDPA_NEW && strategy != FDGA_FDPA_1LOOP)
        FAIL("mfRGLinearMap: Invalid strategy. Must be fdPA or fdPA_new or fdPA_1loop.");      // src/mfRG.jl:26-28
    if (unflatten_dev(ctx, ctx->lev[0], x_dev, factor)) return 1;
    if (fdga_build_K3_cache(ctx, 1, first)) return 1;
    if (strategy == FDGA_FDPA_NEW) {                             // src/mfRG.jl:65-83; L_K3 only reads caches and bubbles
        ctx->defer = true;
        int rc = bse_stage(ctx, bse_L_K3_stage_fn, 0);
        if (!rc) rc = bse_stage(ctx, fdga_bse_K1_new, 1);
        if (!rc) rc = bse_stage(ctx, fdga_bse_K2_new, 1);
        if (!rc) rc = bse_stage(ctx, fdga_bse_K3, 1);
        ctx->defer = false; ctx->pending.clear();
        if (rc) return 1;
    } else if (bse_stages(ctx, true, 1)) return 1;
    if (fdga_set_F_from_Fbuff(ctx)) return 1;
    LAUNCH(FDGA_T_MISC, mfrg_residual_kernel, nblk(ctx->lenFlat, 256), 256, y_dev, x_dev, ctx->lev[0].block, factor, (long long)ctx->lenFlat);
    CK(cudaGetLastError());
    return 0;
}
int fdga_mfrg_matvec_strategy(fdga_ctx* ctx, const fdga_c64* host_x, fdga_c64* host_y, int first, int strategy) {
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(ctx->flat2, host_x, ctx->lenFlat * sizeof(C), cudaMemcpyHostToDevice, ctx->stream));
    if (mfrg_matvec_dev(ctx, ctx->flat2, ctx->flat, first, strategy)) return 1;
    CK(cudaMemcpyAsync(host_y, ctx->flat, ctx->lenFlat * sizeof(C), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
// the same map for a multi-rank job whose vectors live on ONE host process: x crosses PCIe on the root only and reaches the other
// ranks over NVLink (ncclBroadcast); y is read back on the root only (host_x / host_y may be NULL elsewhere).  Collective.
int fdga_mfrg_matvec_from_root(fdga_ctx* ctx, const fdga_c64* host_x, fdga_c64* host_y, int first, int strategy, int root) {
    CK(cudaSetDevice(ctx->device));
    if (root < 0 || root >= ctx->nranks) FAIL("fdga_mfrg_matvec_from_root: bad root");
    if (ctx->rank == root) {
        if (!host_x || !host_y) FAIL("fdga_mfrg_matvec_from_root: the root needs both host vectors");
        CK(cudaMemcpyAsync(ctx->flat2, host_x, ctx->lenFlat * sizeof(C), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (ctx->nranks > 1) {
        if (!ctx->nccl.Broadcast) FAIL("fdga_mfrg_matvec_from_root: ncclBroadcast not available");
        Scope sc(ctx, FDGA_T_COMM);
        int rc = ctx->nccl.Broadcast(ctx->flat2, ctx->flat2, ctx->lenFlat * 2, /*ncclDouble*/ 8, root, ctx->comm, ctx->stream);
        if (rc != 0) FAIL(std::string("ncclBroadcast: ") + ctx->nccl.GetErrorString(rc));
        ctx->n_launch[FDGA_T_COMM]++;
    }
    if (mfrg_matvec_dev(ctx, ctx->flat2, ctx->flat, first, strategy)) return 1;
    if (ctx->rank == root) CK(cudaMemcpyAsync(host_y, ctx->flat, ctx->lenFlat * sizeof(C), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int fdga_mfrg_matvec(fdga_ctx* ctx, const fdga_c64* host_x, fdga_c64* host_y, int first) {
    return fdga_mfrg_matvec_strategy(ctx, host_x, host_y, first, FDGA_FDPA);
}

// ---- DQGMRES on the device (Saad & Wu 1996; the reference calls Krylov.dqgmres, src/mfRG.jl:147-151) ----------------
static int kry_alloc(fdga_ctx* ctx, int memory) {
    const size_t n = ctx->lenFlat;
    if (!ctx->kryW) {
        CK(cudaMalloc(&ctx->kryW, n * sizeof(C))); CK(cudaMalloc(&ctx->kryX, n * sizeof(C)));
        CK(cudaMalloc(&ctx->kryPart, (size_t)FDGA_KRY_BLOCKS * 16 * sizeof(C))); CK(cudaMalloc(&ctx->kryTicket, sizeof(unsigned int)));
        CK(cudaMemset(ctx->kryTicket, 0, sizeof(unsigned int)));
    }
    if (memory > ctx->kry_mem) {
        cudaFree(ctx->kryV); cudaFree(ctx->kryP); cudaFree(ctx->kryH); if (ctx->kryHhost) cudaFreeHost(ctx->kryHhost);
        ctx->kryV = ctx->kryP = ctx->kryH = ctx->kryHhost = nullptr; ctx->kry_mem = 0;
        CK(cudaMalloc(&ctx->kryV, (size_t)memory * n * sizeof(C)));
        CK(cudaMalloc(&ctx->kryP, (size_t)memory * n * sizeof(C)));
        CK(cudaMalloc(&ctx->kryH, (size_t)(memory + 2) * sizeof(C)));
        CK(cudaMallocHost(&ctx->kryHhost, (size_t)(memory + 2) * sizeof(C)));
        ctx->kry_mem = memory;
    }
    return 0;
}
struct cd { double re, im; };
static inline cd cmul(cd a, cd b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
static inline cd cadd(cd a, cd b) { return {a.re + b.re, a.im + b.im}; }
static inline cd cconj(cd a) { return {a.re, -a.im}; }
static inline cd cscale(cd a, double s) { return {a.re * s, a.im * s}; }
static inline double cabs_(cd a) { return hypot(a.re, a.im); }
// complex Givens rotation [c s; -conj(s) c] [a; b] = [rho; 0], c real, b real >= 0
static void sym_givens(cd a, double b, double& c, cd& s, cd& rho) {
    if (b == 0.0) { c = 1.0; s = {0.0, 0.0}; rho = a; return; }
    const double aa = cabs_(a);
    if (aa == 0.0) { c = 0.0; s = {1.0, 0.0}; rho = {b, 0.0}; return; }
    const double t = hypot(aa, b);
    const cd ph = cscale(a, 1.0 / aa);
    c = aa / t; s = cscale(ph, b / t); rho = cscale(ph, t);
}
#define KRY_MGS(vsub, hsub, vdot, outp) do { \
    kry_mgs_kernel<<<FDGA_KRY_BLOCKS, FDGA_KRY_THREADS, 0, ctx->stream>>>(ctx->kryW, vsub, hsub, vdot, outp, ctx->kryPart, ctx->kryTicket, (long long)n); \
    ctx->n_launch[FDGA_T_KRYLOV]++; ctx->total_launches++; } while (0)

// x_dev <- approximate solution of A x = b_dev (x0 = 0); b_dev is overwritten.  Vectors never leave the device.
static int dqgmres_dev(fdga_ctx* ctx, C* b_dev, C* x_dev, int strategy, int memory, double atol, double rtol, int itmax,
                       int* niter, int* solved, double* residuals, int nres) {
    const size_t n = ctx->lenFlat;
    if (memory < 1) FAIL("fdga_mfrg_dqgmres: memory must be >= 1");
    if (itmax <= 0) itmax = 2 * (int)std::min<size_t>(n, 1u << 20);
    if (kry_alloc(ctx, memory)) return 1;
    const int k = memory;
    C* V = ctx->kryV; C* P = ctx->kryP; C* Hd = ctx->kryH; cd* Hh = reinterpret_cast<cd*>(ctx->kryHhost);
    auto Vs = [&](int m) { return V + (size_t)((m - 1) % k) * n; };     // v_m, m = 1, 2, ...
    auto Ps = [&](int m) { return P + (size_t)((m - 1) % k) * n; };
    std::vector<double> cs(k); std::vector<cd> sn(k);                    // rotation i in slot (i - 1) % k
    *niter = 0; *solved = 0;
    CK(cudaMemsetAsync(x_dev, 0, n * sizeof(C), ctx->stream));
    // beta = ||b||
    CK(cudaMemcpyAsync(ctx->kryW, b_dev, n * sizeof(C), cudaMemcpyDeviceToDevice, ctx->stream));
    {
        Scope sc(ctx, FDGA_T_KRYLOV);
        KRY_MGS(nullptr, nullptr, nullptr, Hd);
        CK(cudaGetLastError());
    }
    CK(cudaMemcpyAsync(Hh, Hd, sizeof(C), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const double beta = sqrt(Hh[0].re);
    int nr = 0;
    if (residuals && nr < nres) residuals[nr++] = beta;
    if (beta == 0.0) { *solved = 1; return 0; }
    const double eps = atol + rtol * beta;
    {
        Scope sc(ctx, FDGA_T_KRYLOV);
        LAUNCH(FDGA_T_KRYLOV, scale_copy_kernel, nblk(n, 256), 256, Vs(1), ctx->kryW, 1.0 / beta, (long long)n);
        CK(cudaGetLastError());
    }
    cd gamma = {beta, 0.0};
    std::vector<cd> t(k + 3);
    bool first = true;
    for (int m = 1; m <= itmax; ++m) {
        if (mfrg_matvec_dev(ctx, Vs(m), ctx->kryW, first ? 1 : 0, strategy)) return 1;     // w = A v_m
        first = false;
        const int lo = std::max(1, m - k + 1), cnt = m - lo + 1;
        {   // incomplete modified Gram-Schmidt against v_lo .. v_m, then ||w||^2: cnt + 1 fused launches, no host sync
            Scope sc(ctx, FDGA_T_KRYLOV);
            // blocks of FDGA_KRY_B vectors: launch b subtracts block b - 1 and produces the coefficients of block b; the last
            // launch subtracts the last block and returns ||w||^2 in Hd[cnt]
            static const bool blocked = getenv("FDGA_KRY_BLOCKED") ? atoi(getenv("FDGA_KRY_BLOCKED")) != 0 : false;
            if (!blocked) {       // one fused launch per basis vector
                KRY_MGS(nullptr, nullptr, Vs(lo), Hd);
                for (int j = 1; j < cnt; ++j) KRY_MGS(Vs(lo + j - 1), Hd + (j - 1), Vs(lo + j), Hd + j);
                KRY_MGS(Vs(m), Hd + (cnt - 1), nullptr, Hd + cnt);
            }
            const int nb = blocked ? (cnt + FDGA_KRY_B - 1) / FDGA_KRY_B : -1;
            for (int b = 0; b <= nb; ++b) {
                KryBlk blk; blk.nprev = 0; blk.nnew = 0; blk.norm = (b == nb) ? 1 : 0;
                for (int p = 0; p < FDGA_KRY_B; ++p) { blk.prev_slot[p] = 0; blk.new_slot[p] = 0; }
                if (b > 0) for (int j = (b - 1) * FDGA_KRY_B; j < std::min(cnt, b * FDGA_KRY_B); ++j) blk.prev_slot[blk.nprev++] = (lo + j - 1) % k;
                if (b < nb) for (int j = b * FDGA_KRY_B; j < std::min(cnt, (b + 1) * FDGA_KRY_B); ++j) blk.new_slot[blk.nnew++] = (lo + j - 1) % k;
                kry_bmgs_kernel<<<FDGA_KRY_BLOCKS, FDGA_KRY_THREADS, 0, ctx->stream>>>(ctx->kryW, V, (long long)n, blk,
                    Hd + std::max(0, (b - 1) * FDGA_KRY_B), Hd + std::min(cnt, b * FDGA_KRY_B), ctx->kryPart, ctx->kryTicket);
                NOTE_LAUNCH("kry_bmgs_kernel");
                ctx->n_launch[FDGA_T_KRYLOV]++; ctx->total_launches++;
            }
            CK(cudaGetLastError());
        }
        CK(cudaMemcpyAsync(Hh, Hd, (size_t)(cnt + 1) * sizeof(C), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        const double hnext = sqrt(std::max(0.0, Hh[cnt].re));
        // column m of the Hessenberg matrix: rows plo .. m (+ hnext at m + 1); row m - k lies outside the window (zero)
        const int plo = std::max(1, m - k);
        int off = 0;
        if (plo < lo) { t[0] = {0.0, 0.0}; off = 1; }
        for (int j = 0; j < cnt; ++j) t[off + j] = Hh[j];                  // t[i - plo], i = plo .. m
        for (int i = plo; i < m; ++i) {                                    // previous rotations on rows (i, i + 1)
            const cd ti = t[i - plo], tj = t[i + 1 - plo];
            const double c = cs[(i - 1) % k]; const cd s = sn[(i - 1) % k];
            t[i - plo] = cadd(cscale(ti, c), cmul(s, tj));
            t[i + 1 - plo] = cadd(cscale(cmul(cconj(s), ti), -1.0), cscale(tj, c));
        }
        double c; cd s, rmm;
        sym_givens(t[m - plo], hnext, c, s, rmm);
        cs[(m - 1) % k] = c; sn[(m - 1) % k] = s;
        const cd gamma_next = cscale(cmul(cconj(s), gamma), -1.0);
        gamma = cscale(gamma, c);
        {   // p_m = (v_m - sum_{i = plo}^{m-1} t_i p_i) / r_mm ;  x += gamma_m p_m.  p_m takes the ring slot of p_{m-k}, which is
            // still an input: the kernel is elementwise (each thread reads its element of every slot before writing), and
            // p_{m-k} is the first coefficient of the first chunk, so later chunks only see the accumulated value.
            Scope sc(ctx, FDGA_T_KRYLOV);
            const double r2 = rmm.re * rmm.re + rmm.im * rmm.im;
            const C inv_r = mkC(rmm.re / r2, -rmm.im / r2), gm = mkC(gamma.re, gamma.im);
            const int np = m - plo;                                        // number of previous directions
            C* dst = Ps(m);
            int done = 0;
            do {
                KryCoefs cf; cf.nc = std::min(FDGA_KRY_CHUNK, np - done);
                for (int j = 0; j < cf.nc; ++j) { const int i = plo + done + j; cf.slot[j] = (i - 1) % k; cf.t[j] = mkC(t[i - plo].re, t[i - plo].im); }
                const int last = (done + cf.nc >= np) ? 1 : 0;
                LAUNCH(FDGA_T_KRYLOV, kry_direction_kernel, FDGA_KRY_BLOCKS, FDGA_KRY_THREADS, dst, done == 0 ? (const C*)Vs(m) : (const C*)dst, (const C*)P, (long long)n, cf, last, inv_r, gm, x_dev);
                done += cf.nc;
            } while (done < np);
            CK(cudaGetLastError());
        }
        gamma = gamma_next;
        const double rnorm = cabs_(gamma);
        *niter = m;
        if (residuals && nr < nres) residuals[nr++] = rnorm;
        if (rnorm <= eps) { *solved = 1; break; }
        if (hnext == 0.0 || m == itmax) break;
        {   // v_{m+1} = w / h_{m+1,m} into the ring slot of v_{m-k+1} (outside the next window)
            Scope sc(ctx, FDGA_T_KRYLOV);
            LAUNCH(FDGA_T_KRYLOV, scale_copy_kernel, nblk(n, 256), 256, Vs(m + 1), ctx->kryW, 1.0 / hnext, (long long)n);
            CK(cudaGetLastError());
        }
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
// Krylov.dqgmres(mfRGLinearMap(S, strategy), b; atol, rtol, itmax, memory): src/mfRG.jl:147-151.  b and x are host vectors of
// length(S.F); they cross PCIe once each.  residuals (may be NULL) receives ||b|| and the residual estimate of every iteration.
int fdga_mfrg_dqgmres(fdga_ctx* ctx, const fdga_c64* host_b, fdga_c64* host_x, int strategy, int memory, double atol, double rtol,
                      int itmax, int* niter, int* solved, double* residuals, int nres) {
    CK(cudaSetDevice(ctx->device));
    int it = 0, ok = 0;
    if (kry_alloc(ctx, memory)) return 1;
    CK(cudaMemcpyAsync(ctx->flat2, host_b, ctx->lenFlat * sizeof(C), cudaMemcpyHostToDevice, ctx->stream));
    if (dqgmres_dev(ctx, ctx->flat2, ctx->kryX, strategy, memory, atol, rtol, itmax, &it, &ok, residuals, nres)) return 1;
    CK(cudaMemcpyAsync(host_x, ctx->kryX, ctx->lenFlat * sizeof(C), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (niter) *niter = it;
    if (solved) *solved = ok;
    return 0;
}

// symmetrize_solver!(S): src/ParquetSolver.jl:246-259 (Sigma with SG_Sigma, every channel's K1 / K2 / K3 with its group)
int fdga_symmetrize_solver(fdga_ctx* ctx) {
    CK(cudaSetDevice(ctx->device));
    if (wait_copy(ctx)) return 1;
    Scope sc(ctx, FDGA_T_MISC);
    auto sym = [&](int which, C* f) -> int {
        SymGroup& g = ctx->sg[which];
        if (!g.nmem) return 0;
        LAUNCH(FDGA_T_MISC, symmetrize_kernel, nblk(g.nmem, 256), 256, f, sym_dev(g));
        return 0;
    };
    sym(FDGA_SG_SIGMA, ctx->G[FDGA_SIGMA]);
    for (int ch = 0; ch < 3; ch++) {
        const bool pp = ch == FDGA_PCH;
        sym(FDGA_SG_K1, ctx->lev[0].K[ch][0]);
        sym(pp ? FDGA_SG_PP2 : FDGA_SG_PH2, ctx->lev[0].K[ch][1]);
        sym(pp ? FDGA_SG_PP3 : FDGA_SG_PH3, ctx->lev[0].K[ch][2]);
    }
    CK(cudaGetLastError());
    ctx->lev[0].sw_dirty = true; ctx->lev[0].k1h_dirty = true; ctx->lev[0].mom_valid[0] = ctx->lev[0].mom_valid[1] = ctx->lev[0].mom_valid[2] = 0; ctx->fsum_dirty = true;
    return 0;
}

// fixed_point_preconditioned!(R, x, S; strategy, update_Sigma = false, use_preconditioner, krylov_maxiter): src/mfRG.jl:93-171
//   unflatten!(S.F, x); symmetrize_solver!(S); iterate_solver!(S; strategy, update_Sigma = false);
//   R_F = flatten(S.F) - x;  R = use_preconditioner ? dqgmres(mfRGLinearMap(S, strategy), R_F; atol = rtol = 1e-6, memory) : R_F
// As in the reference, S.F is left holding the last Krylov vector (the caller unflattens its own iterate next).
int fdga_fixed_point_preconditioned(fdga_ctx* ctx, const fdga_c64* host_x, fdga_c64* host_R, int strategy, int use_preconditioner,
                                    int krylov_maxiter, int memory, int* niter, int* solved) {
    CK(cudaSetDevice(ctx->device));
    if (strategy != FDGA_FDPA && strategy != FDGA_FDPA_NEW && strategy != FDGA_FDPA_1LOOP)
        FAIL("fdga_fixed_point_preconditioned: strategy must be fdPA, fdPA_new or fdPA_1loop");
    const size_t n = ctx->lenFlat;
    int it = 0, ok = 1;
    if (kry_alloc(ctx, std::max(1, memory))) return 1;
    CK(cudaMemcpyAsync(ctx->flat2, host_x, n * sizeof(C), cudaMemcpyHostToDevice, ctx->stream));
    if (unflatten_dev(ctx, ctx->lev[0], ctx->flat2, 1.0)) return 1;
    if (fdga_symmetrize_solver(ctx)) return 1;
    if (fdga_iterate_solver(ctx, strategy, 0, 1)) return 1;
    // R_F = flatten(S.F) - x   (flat2 <- R_F)
    LAUNCH(FDGA_T_MISC, axpby_kernel, nblk(n, 256), 256, ctx->flat2, (const C*)ctx->lev[0].block, 1.0, (const C*)ctx->flat2, -1.0, (long long)n);
    CK(cudaGetLastError());
    const C* result = ctx->flat2;
    if (use_preconditioner) {
        if (dqgmres_dev(ctx, ctx->flat2, ctx->kryX, strategy, memory, 1e-6, 1e-6, krylov_maxiter, &it, &ok, nullptr, 0)) return 1;
        result = ctx->kryX;
    }
    CK(cudaMemcpyAsync(host_R, result, n * sizeof(C), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (niter) *niter = it;
    if (solved) *solved = ok;
    return 0;
}

// ---- interpolate_vertex! / interpolate_solver!: src/interpolate.jl:1-191 ------------------------------------------------
// dst (device, frequency count Fo, D momentum components of size Lo) <- Fourier interpolation of host_in (frequency boxes
// box.ni, D components of size Li).  Memory order: frequencies fastest, then the momentum components (x fastest).
static int interp_run(fdga_ctx* ctx, C* dst, const fdga_c64* host_in, InterpBox box, int D, int Li, int Lo, int clamp) {
    if (Li < 1 || Li > 64 || Lo > 64) FAIL("fdga_interpolate: momentum mesh sizes must be in 1..64");
    const long long Fo = (long long)box.no[0] * box.no[1] * box.no[2], Fi = (long long)box.ni[0] * box.ni[1] * box.ni[2];
    long long Mi = 1, big = Fo;
    for (int d = 0; d < D; ++d) Mi *= Li;
    { long long cur = Fo * Mi; big = cur; for (int d = 0; d < D; ++d) { cur = cur / Li * Lo; big = std::max(big, cur); } }
    big = std::max(big, Fi * Mi);
    C* bufA = ctx->scratchA; C* bufB = ctx->scratchB;
    if ((size_t)big > ctx->lenScratch) {      // e.g. coarsening from a finer mesh: dedicated buffers, grown on demand
        if ((size_t)big > ctx->lenItp) {
            cudaFree(ctx->itpA); cudaFree(ctx->itpB); ctx->itpA = ctx->itpB = nullptr; ctx->lenItp = 0;
            CK(cudaMalloc(&ctx->itpA, (size_t)big * sizeof(C))); CK(cudaMalloc(&ctx->itpB, (size_t)big * sizeof(C)));
            ctx->lenItp = (size_t)big;
        }
        bufA = ctx->itpA; bufB = ctx->itpB;
    }
    // M[xo + Lo * xi] = 1/Li sum_{R = -Li/2}^{Li/2} w(R) exp(2 pi i R (xo / Lo - xi / Li)), w = 1/2 at |R| = Li/2 for even Li
    std::vector<C> M((size_t)Lo * Li);
    for (int xi = 0; xi < Li; ++xi) for (int xo = 0; xo < Lo; ++xo) {
        double re = 0.0, im = 0.0;
        for (int R = -(Li / 2); R <= Li / 2; ++R) {
            const double w = (Li % 2 == 0 && std::abs(R) == Li / 2) ? 0.5 : 1.0;
            // exact phase reduction: R * (xo * Li - xi * Lo) / (Lo * Li) turns
            const long long num = (long long)R * ((long long)xo * Li - (long long)xi * Lo), den = (long long)Lo * Li;
            const long long rem = ((num % den) + den) % den;
            const double ph = 2.0 * M_PI * (double)rem / (double)den;
            re += w * cos(ph); im += w * sin(ph);
        }
        M[xo + (size_t)Lo * xi] = mkC(re / Li, im / Li);
    }
    C* dM = ctx->SigR2;     // G-sized scratch, at least 2 nG LG^2 >= 64 * 64 entries for any realistic mesh; checked:
    if ((size_t)Lo * Li > ctx->lenG) FAIL("fdga_interpolate: interpolation matrix does not fit its scratch");
    Scope sc(ctx, FDGA_T_MISC);
    CK(cudaMemcpyAsync(dM, M.data(), M.size() * sizeof(C), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(bufB, host_in, (size_t)(Fi * Mi) * sizeof(C), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));          // M is a stack-lifetime vector, host_in belongs to the caller
    LAUNCH(FDGA_T_MISC, interp_rebox_kernel, nblk(Fo * Mi, 256), 256, (const C*)bufB, bufA, box, Mi, clamp);
    C* cur = bufA; C* other = bufB;
    long long pre = Fo, post = Mi;
    for (int d = 0; d < D; ++d) {
        post /= Li;
        C* out = (d == D - 1) ? dst : other;
        LAUNCH(FDGA_T_MISC, interp_axis_kernel, nblk(pre * Lo * post, 256), 256, (const C*)cur, out, pre, Li, Lo, post, (const C*)dM);
        pre *= Lo;
        other = cur; cur = out;
    }
    CK(cudaGetLastError());
    if (D == 0) CK(cudaMemcpyAsync(dst, bufA, (size_t)Fo * sizeof(C), cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
}
// interpolate_vertex!(Ko, Ki): Ko = class `cls` of channel `channel` of the NL2 vertex `which` of this context, Ki = host array on
// an Li x Li momentum mesh with frequency meshes nKi (cls 0: {N_K1}; 1: {N_K2 bosonic, N_K2 fermionic}; 2: {N_K3 bosonic, N_K3 fermionic})
int fdga_interpolate_vertex(fdga_ctx* ctx, int which, int channel, int cls, const fdga_c64* host_Ki, const int32_t* nKi, int Li) {
    CK(cudaSetDevice(ctx->device));
    LevelBuf* lb = which_level(ctx, which);
    if (!lb || (lb->d.type != FDGA_LV_NL2 && lb->d.type != FDGA_LV_NL) || channel < 0 || channel > 2 || cls < 0 || cls > 2) FAIL("fdga_interpolate_vertex: bad selector (needs an NL2 / NL level)");
    if (lb == &ctx->lev[0] && wait_copy(ctx)) return 1;
    const fdga_level_desc& d = lb->d;
    InterpBox b; b.nd = 1; for (int i = 0; i < 3; ++i) { b.no[i] = 1; b.ni[i] = 1; b.shift[i] = 0; }
    int D = 2;
    if (cls == 0) { b.no[0] = 2 * d.nK1 - 1; b.ni[0] = 2 * nKi[0] - 1; b.shift[0] = nKi[0] - d.nK1; }
    else if (cls == 1) {
        b.nd = 2; D = lb->d.type == FDGA_LV_NL ? 2 : 4;
        b.no[0] = 2 * d.nK2[0] - 1; b.ni[0] = 2 * nKi[0] - 1; b.shift[0] = nKi[0] - d.nK2[0];
        b.no[1] = 2 * d.nK2[1];     b.ni[1] = 2 * nKi[1];     b.shift[1] = nKi[1] - d.nK2[1];
    } else {
        b.nd = 3;
        b.no[0] = 2 * d.nK3[0] - 1; b.ni[0] = 2 * nKi[0] - 1; b.shift[0] = nKi[0] - d.nK3[0];
        b.no[1] = b.no[2] = 2 * d.nK3[1]; b.ni[1] = b.ni[2] = 2 * nKi[1]; b.shift[1] = b.shift[2] = nKi[1] - d.nK3[1];
    }
    for (int i = 0; i < 3; ++i) if (b.ni[i] < 1) FAIL("fdga_interpolate_vertex: bad input mesh sizes");
    if (interp_run(ctx, lb->K[channel][cls], host_Ki, b, D, Li, ctx->g.L, 0)) return 1;
    lb->sw_dirty = true; lb->k1h_dirty = true; lb->mom_valid[0] = lb->mom_valid[1] = lb->mom_valid[2] = 0; ctx->fsum_dirty = true;
    invalidate_rt(ctx);
    return 0;
}
// interpolate_vertex!(Go, Gi) for G-shaped arrays [nu, k] (NL_MF_G, src/interpolate.jl:62-80); clamp != 0: frequencies outside the
// input mesh take the value at its edge, as interpolate_solver! does for the self-energy (src/interpolate.jl:176-186)
int fdga_interpolate_green(fdga_ctx* ctx, int which, const fdga_c64* host_in, int nG_in, int Li, int clamp) {
    CK(cudaSetDevice(ctx->device));
    if (which < 0 || which >= 5 || nG_in < 1) FAIL("fdga_interpolate_green: bad selector");
    InterpBox b; b.nd = 1; for (int i = 0; i < 3; ++i) { b.no[i] = 1; b.ni[i] = 1; b.shift[i] = 0; }
    b.no[0] = 2 * ctx->dims.nG; b.ni[0] = 2 * nG_in; b.shift[0] = nG_in - ctx->dims.nG;
    return interp_run(ctx, ctx->G[which], host_in, b, 2, Li, ctx->g.LG, clamp);
}

// ---- outer loop of solve_using_mfRG! (src/mfRG.jl:217-372): the state updates between two vertex solves, on the device ----
// Pi_mixed = mixing * Pi + (1 - mixing) * Pi0 ; set!(S.Pi, Pi_mixed)   (src/mfRG.jl:271-276).  The mixed bubbles are kept for
// fdga_update_reference.
int fdga_mix_bubbles(fdga_ctx* ctx, double mixing) {
    CK(cudaSetDevice(ctx->device));
    for (int i = 0; i < 2; i++) if (!ctx->PiMixed[i]) CK(cudaMalloc(&ctx->PiMixed[i], ctx->lenPi * sizeof(C)));
    Scope sc(ctx, FDGA_T_MISC);
    const int pi[2] = {FDGA_PIPP, FDGA_PIPH}, pi0[2] = {FDGA_PI0PP, FDGA_PI0PH};
    if (ctx->swave) {
        for (int i = 0; i < 2; i++) {
            LAUNCH(FDGA_T_MISC, axpby_kernel, nblk(ctx->lenPi, 256), 256, ctx->PiMixed[i], (const C*)ctx->Pisw[pi[i]], mixing, (const C*)ctx->Pisw[pi0[i]], 1.0 - mixing, (long long)ctx->lenPi);
            CK(cudaMemcpyAsync(ctx->Pisw[pi[i]], ctx->PiMixed[i], ctx->lenPi * sizeof(C), cudaMemcpyDeviceToDevice, ctx->stream));
        }
        CK(cudaGetLastError());
        return 0;
    }
    for (int i = 0; i < 2; i++) {      // the mixed bubbles are not products of Green functions: they live in the full layout
        if (ensure_pi_full(ctx, pi[i]) || ensure_pi_full(ctx, pi0[i])) return 1;
        LAUNCH(FDGA_T_MISC, axpby_kernel, nblk(ctx->lenPi, 256), 256, ctx->PiMixed[i], (const C*)ctx->Pi[pi[i]], mixing, (const C*)ctx->Pi[pi0[i]], 1.0 - mixing, (long long)ctx->lenPi);
        CK(cudaMemcpyAsync(ctx->Pi[pi[i]], ctx->PiMixed[i], ctx->lenPi * sizeof(C), cudaMemcpyDeviceToDevice, ctx->stream));
        ctx->pi_src[pi[i]] = PI_FULL; ctx->pi_full_valid[pi[i]] = true;
        ctx->pi_dirty[pi[i]] = true;
    }
    CK(cudaGetLastError());
    invalidate_rt(ctx);
    return 0;
}
// "Update reference Green function, bubble, and self-energy; update reference vertex and reset target vertex" (src/mfRG.jl:336-347):
//   set!(S.Pi0pp, Pipp_mixed); set!(S.Pi0ph, Piph_mixed); set!(S.G0, S.G); set!(S.Sigma0, S.Sigma); add!(S.F0, S.F); set!(S.F, 0)
int fdga_update_reference(fdga_ctx* ctx) {
    CK(cudaSetDevice(ctx->device));
    if (!ctx->PiMixed[0]) FAIL("fdga_update_reference: call fdga_mix_bubbles first");
    if (ctx->nlev < 2 || ctx->lev[1].d.type != ctx->lev[0].d.type || ctx->lev[1].blocklen != ctx->lev[0].blocklen)
        FAIL("fdga_update_reference: add!(S.F0, S.F) needs S.F0 to be a vertex of S.F's own type on the meshes of S.F");
    if (wait_copy(ctx)) return 1;
    Scope sc(ctx, FDGA_T_MISC);
    const int pi0[2] = {FDGA_PI0PP, FDGA_PI0PH};
    for (int i = 0; i < 2; i++) {
        if (ctx->swave) { CK(cudaMemcpyAsync(ctx->Pisw[pi0[i]], ctx->PiMixed[i], ctx->lenPi * sizeof(C), cudaMemcpyDeviceToDevice, ctx->stream)); continue; }
        if (!ctx->Pi[pi0[i]]) CK(cudaMalloc(&ctx->Pi[pi0[i]], ctx->lenPi * sizeof(C)));
        CK(cudaMemcpyAsync(ctx->Pi[pi0[i]], ctx->PiMixed[i], ctx->lenPi * sizeof(C), cudaMemcpyDeviceToDevice, ctx->stream));
        ctx->pi_src[pi0[i]] = PI_FULL; ctx->pi_full_valid[pi0[i]] = true;
        ctx->pi_dirty[pi0[i]] = true;
    }
    CK(cudaMemcpyAsync(ctx->G[FDGA_G0], ctx->G[FDGA_G], ctx->lenG * sizeof(C), cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->G[FDGA_SIGMA0], ctx->G[FDGA_SIGMA], ctx->lenG * sizeof(C), cudaMemcpyDeviceToDevice, ctx->stream));
    const long long n = (long long)ctx->lev[0].blocklen;
    LAUNCH(FDGA_T_MISC, add_axpby_kernel, nblk(n, 256), 256, ctx->lev[1].block, (const C*)ctx->lev[0].block, 1.0, (const C*)nullptr, 0.0, n);
    CK(cudaGetLastError());
    CK(cudaMemsetAsync(ctx->lev[0].block, 0, (size_t)n * sizeof(C), ctx->stream));
    for (int l = 0; l < 2; l++) { ctx->lev[l].sw_dirty = true; ctx->lev[l].k1h_dirty = true; ctx->lev[l].mom_valid[0] = ctx->lev[l].mom_valid[1] = ctx->lev[l].mom_valid[2] = 0; }
    ctx->fsum_dirty = true;
    invalidate_rt(ctx);
    return 0;
}

// ---- introspection -----------------------------------------------------------------------------------------
// measured FP64 FMA throughput of this device (2 flop per DFMA), the denominator of the `fp64` roofline entry of bench.py
int fdga_measure_fp64_peak(fdga_ctx* ctx, double* tflops) {
    CK(cudaSetDevice(ctx->device));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, ctx->device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
    double* d_out = nullptr; CK(cudaMalloc(&d_out, (size_t)blocks * threads * sizeof(double)));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0, ctx->stream));
        dfma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d_out, iters, 0.999999, 1e-6);
        CK(cudaEventRecord(e1, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        float ms = 0.f; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best = std::min(best, ms);
    }
    CK(cudaGetLastError());
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
    *tflops = 2.0 * 8.0 * (double)iters * blocks * threads / (best * 1e-3) / 1e12;
    return 0;
}
int fdga_profile_enable(fdga_ctx* ctx, int on) { ctx->profile = on != 0; return 0; }
static int profile_collect(fdga_ctx* ctx) {
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    for (auto& ev : ctx->events) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ev.a, ev.b) == cudaSuccess) ctx->t_ms[ev.cat] += ms;
        cudaEventDestroy(ev.a); cudaEventDestroy(ev.b);
    }
    ctx->events.clear();
    return 0;
}
int fdga_profile_reset(fdga_ctx* ctx) {
    if (profile_collect(ctx)) return 1;
    memset(ctx->t_ms, 0, sizeof(ctx->t_ms)); memset(ctx->n_launch, 0, sizeof(ctx->n_launch)); ctx->total_launches = 0;
    return 0;
}
int fdga_kernel_time_ms(fdga_ctx* ctx, int id, double* ms, int64_t* launches) {
    if (id < 0 || id >= FDGA_T_COUNT) FAIL("fdga_kernel_time_ms: bad id");
    if (profile_collect(ctx)) return 1;
    *ms = ctx->t_ms[id]; *launches = ctx->n_launch[id];
    return 0;
}
int64_t fdga_total_launches(fdga_ctx* ctx) { return ctx->total_launches; }

}  // extern "C"
