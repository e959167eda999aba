/* ======================================================================================
 * fdga.h -- C-ABI of libfdga.so, the B200 (sm_100a) implementation of the Bethe-Salpeter /
 * K3-cache / bubble / Schwinger-Dyson hot path of jaemolihm/fdDGAsolver.jl.
 *
 * The reference has no FFI: its seam is Julia multiple dispatch on the concrete array
 * aliases of src/types.jl:105-129.  Every entry point below replaces one reference method
 * (cited as file:line relative to the reference tree) and is what a Julia `ccall` shim
 *  binds.  Plain pointers and sizes only; no torch / C++ types.
 *
 * Conventions
 *  - All arrays are column-major (first index fastest) interleaved complex double, i.e. the
 *    memory of a Julia Array{ComplexF64,N} / a numpy complex128 array in Fortran order.
 *  - The caller owns host memory; the library owns device memory.  set_* copies H2D, get_*
 *    copies D2H; kernels work on the resident device state.
 *  - Every function returns 0 on success, non-zero on error; fdga_last_error() gives the text.
 *    Nothing throws across the boundary.  A context is not thread-safe; distinct contexts are
 *    independent.  There is NO CPU fallback: fdga_create fails if no CUDA device is usable.
 *  - Matsubara meshes (MatsubaraFunctions.jl): fermionic mesh N -> indices -N..N-1 (2N points),
 *    bosonic mesh N -> indices -(N-1)..N-1 (2N-1 points).  Momentum mesh nq x nq, linear index
 *    ix + nq*iy (0-based).
 *  - Channels: 0 = pCh, 1 = tCh, 2 = aCh (flatten order of src/vertex.jl:153-167).
 * ====================================================================================== */
#ifndef FDGA_H
#define FDGA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fdga_ctx fdga_ctx;
typedef struct { double re, im; } fdga_c64;

enum { FDGA_PCH = 0, FDGA_TCH = 1, FDGA_ACH = 2 };
enum { FDGA_K1 = 0, FDGA_K2 = 1, FDGA_K3 = 2 };
/* level types of the nested vertex chain S.F -> S.F.F0 -> ... (src/vertex.jl, src/refvertex.jl) */
enum { FDGA_LV_NL2 = 0,    /* NL2_Vertex  (src/nonlocal_2/vertex.jl:1-32)  K1[W,P] K2[W,v,P,k] K3[W,v,v',P] */
       FDGA_LV_LOCAL = 1,  /* Vertex      (src/vertex.jl:7-36)             K1[W]   K2[W,v]     K3[W,v,v']   */
       FDGA_LV_CORE = 2,   /* RefVertex   (src/refvertex.jl:1-35)          U + Fp_p, Fp_x, Ft_p, Ft_x       */
       FDGA_LV_NL = 3,     /* NL_Vertex   (src/nonlocal/vertex.jl:1-37)    K1[W,P] K2[W,v,P]   K3[W,v,v',P] */
       /* multi-boson-exchange vertices (src/boson_exchange.jl:237-266, 738-770): the arrays of the base type, evaluated as
        * U + K1 + K2 + K2' + K2 K2' / (U + K1) + K3 per channel.  MBE levels form the head of the chain; a context with one runs the
        * generic per-term kernels (no `_new` / `_1loop` variants beyond what those kernels provide). */
       FDGA_LV_NL2_MBE = 4, FDGA_LV_LOCAL_MBE = 5 };

#define FDGA_MAX_LEVELS 6

typedef struct {
    int32_t type;
    int32_t nK1;         /* N of the bosonic K1 mesh                                   */
    int32_t nK2[2];      /* N of the (bosonic, fermionic) K2 meshes                    */
    int32_t nK3[2];      /* N of the (bosonic, fermionic) K3 meshes; CORE: the core box */
    double  U_re, U_im;  /* CORE only: bare vertex                                     */
} fdga_level_desc;

/* Shape of an NL2_ParquetSolver (src/nonlocal_2/ParquetSolver.jl:1-154).
 * lev[0] is S.F's own NL2 level; lev[1..nlev-1] is the chain of S.F0 (S.F.F0 === S.F0),
 * terminated by a CORE level.  S.Fbuff and S.FL have lev[0]'s grids and a null core. */
typedef struct {
    double  T;           /* temperature                                       */
    int32_t nq;          /* linear size L of the vertex / bubble momentum mesh */
    int32_t LG;          /* linear size of the G / Sigma momentum mesh         */
    int32_t nG;          /* N of the fermionic mesh of G, Sigma               */
    int32_t nPiB, nPiF;  /* N of the bubble's bosonic / fermionic meshes      */
    int32_t nlev;
    fdga_level_desc lev[FDGA_MAX_LEVELS];
} fdga_dims;

/* which-vertex selectors for set/get_vertex: a level index 0..nlev-1, or: */
enum { FDGA_V_FL = 100, FDGA_V_FBUFF = 101 };
enum { FDGA_G = 0, FDGA_G0 = 1, FDGA_GBARE = 2, FDGA_SIGMA = 3, FDGA_SIGMA0 = 4 };
enum { FDGA_PI0PP = 0, FDGA_PI0PH = 1, FDGA_PIPP = 2, FDGA_PIPH = 3 };
/* K3-shaped caches, src/nonlocal_2/ParquetSolver.jl:138-147 */
enum { FDGA_C_GPX = 0, FDGA_C_F0P = 1, FDGA_C_F0A = 2, FDGA_C_F0T = 3, FDGA_C_GPP = 4,
       FDGA_C_GA = 5, FDGA_C_GT = 6, FDGA_C_FP = 7, FDGA_C_FA = 8, FDGA_C_FT = 9 };
/* symmetry groups, src/nonlocal_2/ParquetSolver.jl:200-291 (SGxx[i] -> K(i) class) */
enum { FDGA_SG_SIGMA = 0, FDGA_SG_K1 = 1, FDGA_SG_PP2 = 2, FDGA_SG_PH2 = 3, FDGA_SG_PP3 = 4,
       FDGA_SG_PH3 = 5, FDGA_SG_PPL3 = 6, FDGA_SG_PHL3 = 7, FDGA_SG_COUNT = 8,
       /* builder-only ids (fdga_build_symmetry_group): the K2[W,v,P] groups of the s-wave solver, src/nonlocal/ParquetSolver.jl:215-250;
        * they are registered in the FDGA_SG_PP2 / FDGA_SG_PH2 slots of an s-wave context */
       FDGA_SG_NL_PP2 = 8, FDGA_SG_NL_PH2 = 9 };
/* strategies of src/solve.jl:10, src/SDE.jl:3-33 (:scPA, :fdPA, :scPA_new, :fdPA_new, :fdPA_1loop) */
enum { FDGA_SCPA = 0, FDGA_FDPA = 1, FDGA_SCPA_NEW = 2, FDGA_FDPA_NEW = 3, FDGA_FDPA_1LOOP = 4 };

/* ---- lifecycle ------------------------------------------------------------------------ */
int  fdga_create(const fdga_dims* dims, int device, fdga_ctx** out);
int  fdga_destroy(fdga_ctx* ctx);
const char* fdga_last_error(fdga_ctx* ctx);         /* ctx may be NULL: error of the last failed create */
int  fdga_sync(fdga_ctx* ctx);
/* Options.  FDGA_OPT_SDE_OWN_GAMMA: 0 (default) = SDE L kernels exactly as coded in src/nonlocal_2/SDE.jl:26-29,64-69
 * (F(...; own gamma) - F.F0(...; own gamma), which for a nested nonlocal F0 also contains F0's cross channels);
 * 1 = as the in-line comments there state (own-channel gamma of F only).  See DESIGN.md "E2". */
enum { FDGA_OPT_SDE_OWN_GAMMA = 0,
       FDGA_OPT_GENERIC_KERNELS = 1, /* 1 = use the straightforward per-term kernels instead of the column kernels (A/B check) */
       FDGA_OPT_FD_HARTREE_ONCE = 2, /* 0 (default) = fdPA SDE! exactly as coded in src/SDE.jl:13-24 (reference Hartree term
                                        subtracted twice); 1 = subtracted once (reproduces test/test_siam_fdPA.jl:84) */
       FDGA_OPT_LOCAL_SOLVER = 3     /* 1 = the context models the local ParquetSolver (src/ParquetSolver.jl) on a 1 x 1
                                        momentum mesh: BSE_L_K2! in its local form (src/BSEa/BSEa_K2.jl:1-40), bubbles! with
                                        the 1/nu tail (src/bubble.jl:9-36).  Needs nq = LG = 1. */,
       FDGA_OPT_DIRECT_K1 = 4        /* 1 = sum the cross-channel K1 terms of the column kernels term by term instead of
                                        through the per-slab momentum convolution (A/B check of slab_conv_kernel) */,
       FDGA_OPT_SERIAL = 5           /* concurrency of the three channels of a BSE stage and of the pp / ph / U^2 parts of the SDE:
                                        0 (default) = concurrent streams when one bubble-shaped array is <= 160 MB (the right
                                        factors of the channels then share the L2), one stream otherwise; 1 = always one stream;
                                        2 = always concurrent */,
       FDGA_OPT_QLANE = 6            /* contraction kernel of BSE_K2! / BSE_L_K2! / the SDE L arrays: -1 (default) = the q-lane
                                        kernel (one warp per class representative, lanes over the inner momentum, vertex tables
                                        in momentum-fastest layouts) when nq * nq >= 16, the column kernel otherwise;
                                        0 = always the column kernel; 1 = always the q-lane kernel (A/B check) */ };
int  fdga_set_option(fdga_ctx* ctx, int opt, int value);
/* one process per GPU; `unique_id` = the 128-byte ncclUniqueId obtained on rank 0 by
 * fdga_comm_unique_id and broadcast by the host (MPI in Julia, torch.distributed in tests). */
int  fdga_comm_unique_id(void* unique_id_128B);
int  fdga_comm_init(fdga_ctx* ctx, int nranks, int rank, const void* unique_id_128B);
/* how class representatives are sharded: rank r computes classes [c0, c1) and all-gathers `chunk` slots per rank
 * (host-only helper, no context; replaces the mpi_split of src/nonlocal_2/build_K3_cache.jl:35 / SG(...; mode = :hybrid)) */
int  fdga_partition(int64_t nclasses, int nranks, int rank, int64_t* c0, int64_t* c1, int64_t* chunk);

/* ---- data in / out (replace MeshFunction .data assignment / set!, src/channel.jl:80-150) -- */
int  fdga_set_vertex(fdga_ctx*, int which, int channel, int cls, const fdga_c64* host, int64_t n);
int  fdga_get_vertex(fdga_ctx*, int which, int channel, int cls, fdga_c64* host, int64_t n);
int  fdga_set_core(fdga_ctx*, int level, int which4 /*0 Fp_p 1 Fp_x 2 Ft_p 3 Ft_x*/, const fdga_c64* host, int64_t n);
int  fdga_set_green(fdga_ctx*, int which, const fdga_c64* host, int64_t n);
int  fdga_get_green(fdga_ctx*, int which, fdga_c64* host, int64_t n);
int  fdga_set_bubble(fdga_ctx*, int which, const fdga_c64* host, int64_t n);
int  fdga_get_bubble(fdga_ctx*, int which, fdga_c64* host, int64_t n);
int  fdga_set_cache(fdga_ctx*, int which, const fdga_c64* host, int64_t n);
int  fdga_get_cache(fdga_ctx*, int which, fdga_c64* host, int64_t n);
int  fdga_get_L(fdga_ctx*, int is_pp, fdga_c64* host, int64_t n);
/* SG.classes flattened (MatsubaraFunctions SymmetryGroup): CSR offsets (nclasses+1), 0-based linear
 * member indices, op bits (bit0 = sgn, bit1 = con); representative = first member of a class. */
int  fdga_set_symmetry_classes(fdga_ctx*, int which_sg, int64_t nclasses, const int64_t* offsets,
                               const int64_t* index, const uint8_t* ops);
/* Stand-alone class-table builder (host, integer only) restating SymmetryGroup(symmetries, f) for the
 * generator lists of init_sym_grp! (src/nonlocal_2/ParquetSolver.jl:200-291; generators
 * src/nonlocal/symmetries.jl, src/nonlocal_2/symmetries.jl).  n0 = N of the first mesh, n1 = N of the
 * fermionic meshes.  offsets needs len+1 entries, index/ops len entries (len = array length). */
int  fdga_build_symmetry_group(int which_sg, int n0, int n1, int nq, int64_t* offsets, int64_t* index,
                               uint8_t* ops, int64_t* nclasses);
/* flatten(S.F) / unflatten!(S.F, x*scale): src/vertex.jl:153-195, src/channel.jl:155-210 */
int64_t fdga_length_F(fdga_ctx*);
int  fdga_flatten_F(fdga_ctx*, fdga_c64* host_y);
/* same copy on a side stream, ordered after everything issued so far and overlapping what is issued next (e.g. SDE!
 * after iterate_solver!); host_y (ideally pinned) is valid after the next fdga_sync.  Later writers of S.F wait for it. */
int  fdga_flatten_F_async(fdga_ctx*, fdga_c64* host_y);
int  fdga_unflatten_F(fdga_ctx*, const fdga_c64* host_x, double scale);   /* asynchronous on the context stream */
/* collective form for a multi-rank job whose vector lives on one host: the root copies over PCIe once, the other ranks receive
 * over NVLink (ncclBroadcast); host_x is read on the root only */
int  fdga_unflatten_F_from_root(fdga_ctx*, const fdga_c64* host_x, double scale, int root);
/* device-resident copy of S.F (copy(S.F) / set!(S.F, copy)): lets a caller restart iterations without host traffic */
int  fdga_stash_F(fdga_ctx*);
int  fdga_unstash_F(fdga_ctx*);

/* ---- kernels: one per reference function ---------------------------------------------- */
/* Dyson!(S): src/dyson.jl:20-31 */
int  fdga_dyson(fdga_ctx*);
/* compute_occupation(G): src/dyson.jl:39-41; which = FDGA_G or FDGA_G0 */
int  fdga_occupation(fdga_ctx*, int which, double* occ);
/* set!(S.Gbare, hubbard_bare_Green(meshes(S.Gbare)...; mu, t1, t2, t3)) evaluated on the device: src/models/hubbard.jl:8-44 */
int  fdga_set_hubbard_bare_green(fdga_ctx*, double mu, double t1, double t2, double t3);
/* compute_hubbard_chemical_potential(occ_target, S.Sigma, hubbard_params): src/dyson.jl:45-57 (bracket -4|t1| .. 4|t1|, bisection to
 * neighbouring floats as Roots.find_zero does; one fused Gbare(mu) -> Dyson -> occupation kernel per evaluation).  Fails when the
 * bracket does not contain the target occupation. */
int  fdga_hubbard_chemical_potential(fdga_ctx*, double occ_target, double t1, double t2, double t3, double* mu);
/* bubbles_real_space!(Pipp, Piph, G): src/nonlocal_2/bubble.jl:42-122; reference != 0 -> (Pi0, G0) */
int  fdga_bubbles_real_space(fdga_ctx*, int reference);
/* bubbles!(Pipp, Piph, G) of the local solver: src/bubble.jl:9-36 (nq = LG = 1) */
int  fdga_bubbles_local(fdga_ctx*, int reference);
/* bubbles_momentum_space!: src/nonlocal_2/bubble.jl:1-37 (cross-check; needs LG % nq == 0) */
int  fdga_bubbles_momentum_space(fdga_ctx*, int reference);
/* build_K3_cache!(S): src/nonlocal_2/build_K3_cache.jl:18-94; mfrg != 0 -> build_K3_cache_mfRG!(S, first) :97-164 */
int  fdga_build_K3_cache(fdga_ctx*, int mfrg, int first);
/* BSE_L_K2!(S, Ch): src/BSE_templates.jl:47-76 -> src/nonlocal_2/BSEa/BSEa_K2.jl:1-49 */
int  fdga_bse_L_K2(fdga_ctx*, int ch);
/* BSE_L_K3!(S, Ch): src/BSE_templates.jl:117-146 -> src/nonlocal_2/BSEa/BSEa_K3.jl:1-40 */
int  fdga_bse_L_K3(fdga_ctx*, int ch);
/* BSE_K1!(S, Ch, is_mfRG): src/BSE_templates.jl:12-41 -> src/nonlocal_2/BSEa/BSEa_K1.jl:2-58 */
int  fdga_bse_K1(fdga_ctx*, int ch, int mfrg);
/* BSE_K2!(S, Ch, is_mfRG): src/BSE_templates.jl:82-111 -> src/nonlocal_2/BSEa/BSEa_K2.jl:55-138 */
int  fdga_bse_K2(fdga_ctx*, int ch, int mfrg);
/* BSE_K3!(S, Ch, is_mfRG): src/BSE_templates.jl:151-180 -> src/nonlocal_2/BSEa/BSEa_K3.jl:43-128 */
int  fdga_bse_K3(fdga_ctx*, int ch, int mfrg);
/* BSE_K1_new!(S, Ch, is_mfRG): src/BSE_templates.jl:188-218 -> src/nonlocal_2/BSEa/BSEa_K1.jl:62-113  (K1 = (U + K1 + K2') Pi U) */
int  fdga_bse_K1_new(fdga_ctx*, int ch, int mfrg);
/* BSE_K2_new!(S, Ch, is_mfRG): src/BSE_templates.jl:223-253 -> src/nonlocal_2/BSEa/BSEa_K2.jl:142-216 */
int  fdga_bse_K2_new(fdga_ctx*, int ch, int mfrg);
/* BSE_K1_1loop!(S, Ch, is_mfRG): src/BSE_templates.jl:261-291 -> src/nonlocal_2/BSEa/BSE_1loop.jl:2-56 */
int  fdga_bse_K1_1loop(fdga_ctx*, int ch, int mfrg);
/* BSE_K2_1loop!(S, Ch, is_mfRG): src/BSE_templates.jl:297-327 -> src/nonlocal_2/BSEa/BSE_1loop.jl:59-124 */
int  fdga_bse_K2_1loop(fdga_ctx*, int ch, int mfrg);
/* BSE_K3_1loop!(S, Ch, is_mfRG): src/BSE_templates.jl:332-358 -> src/nonlocal_2/BSEa/BSE_1loop.jl:123-199 */
int  fdga_bse_K3_1loop(fdga_ctx*, int ch, int mfrg);
/* set!(S.F, S.Fbuff): src/solve.jl:87 */
int  fdga_set_F_from_Fbuff(fdga_ctx*);
/* SDE!(S; strategy, include_U2, include_Hartree): src/SDE.jl:3-48 -> src/nonlocal_2/SDE.jl:154-324.  An unknown strategy is
 * rejected before S.Sigma is touched (src/SDE.jl:31). */
int  fdga_sde(fdga_ctx*, int strategy, int include_U2, int include_Hartree);
/* SDE_channel_L_pp! and SDE_channel_L_ph! (src/nonlocal_2/SDE.jl:16-33, 54-73) for the vertex chain starting at level `from`
 * (0 = S.F, 1 = S.F0, ...), summed over the levels of the chain with the weights SDE! gives them (1/3 for the RefVertex level,
 * src/nonlocal_2/SDE.jl:305-309), against the bubbles of S (reference = 0) or the reference bubbles (reference = 1).
 * Results: fdga_get_L.  (fdga_sde transforms these arrays in place afterwards, as the reference does: SURVEY E4.) */
int  fdga_sde_channel_L(fdga_ctx*, int reference, int from);
/* iterate_solver!(S; strategy, update_Sigma, compute_Hartree): src/solve.jl:4-116 (all five strategies); compute_hartree is
 * forwarded to SDE! as include_Hartree (src/solve.jl:99; false for DGammaA runs whose Sigma0 already holds the Hartree term) */
int  fdga_iterate_solver(fdga_ctx*, int strategy, int update_sigma, int compute_hartree);
/* mfRGLinearMap matvec: src/mfRG.jl:34-89 (strategy fdPA); first = is_first_iteration */
int  fdga_mfrg_matvec(fdga_ctx*, const fdga_c64* host_x, fdga_c64* host_y, int first);

/* mfRGLinearMap(S, strategy) matvec: src/mfRG.jl:20-89; strategy FDGA_FDPA, FDGA_FDPA_1LOOP (same map) or FDGA_FDPA_NEW */
int  fdga_mfrg_matvec_strategy(fdga_ctx*, const fdga_c64* host_x, fdga_c64* host_y, int first, int strategy);
/* the same map when the vectors of a multi-rank job live on ONE host process (collective): x is uploaded by `root` only and
 * broadcast over NVLink, y is read back on `root` only; host_x / host_y may be NULL on the other ranks */
int  fdga_mfrg_matvec_from_root(fdga_ctx*, const fdga_c64* host_x, fdga_c64* host_y, int first, int strategy, int root);
/* Krylov.dqgmres(mfRGLinearMap(S, strategy), b; atol, rtol, itmax, memory) as called at src/mfRG.jl:147-151, with the Krylov
 * basis, the direction vectors and the iterate resident in HBM: b and x cross PCIe once.  x0 = 0, no
 * preconditioner, modified Gram-Schmidt over the last `memory` vectors; stop when the quasi-residual estimate
 * <= atol + rtol ||b|| (solved = 1) or after itmax iterations (itmax <= 0: 2 n).  residuals may be NULL; otherwise it receives
 * ||b|| and then one estimate per iteration (at most nres entries). */
int  fdga_mfrg_dqgmres(fdga_ctx*, const fdga_c64* host_b, fdga_c64* host_x, int strategy, int memory, double atol, double rtol,
                       int itmax, int* niter, int* solved, double* residuals, int nres);
/* symmetrize_solver!(S): src/ParquetSolver.jl:246-259 */
int  fdga_symmetrize_solver(fdga_ctx*);
/* fixed_point_preconditioned!(R, x, S; strategy, update_Sigma = false, use_preconditioner, krylov_maxiter): src/mfRG.jl:93-171
 * (the function nlsolve iterates in solve_using_mfRG!, src/mfRG.jl:287); memory = 100 and atol = rtol = 1e-6 in the reference */
int  fdga_fixed_point_preconditioned(fdga_ctx*, const fdga_c64* host_x, fdga_c64* host_R, int strategy, int use_preconditioner,
                                     int krylov_maxiter, int memory, int* niter, int* solved);

/* interpolate_vertex!(Ko, Ki): src/interpolate.jl:1-165 -- Fourier interpolation between momentum meshes.  Ko = class `cls` of
 * channel `channel` of the NL2 vertex `which` of this context (its own meshes); Ki = host array on an Li x Li momentum mesh with
 * frequency meshes nKi (cls K1: {N}; K2: {N bosonic, N fermionic}; K3: {N bosonic, N fermionic}).  Frequencies outside Ki's
 * meshes are set to 0 (set!(Ko, 0) + is_inbounds || continue). */
int  fdga_interpolate_vertex(fdga_ctx*, int which, int channel, int cls, const fdga_c64* host_Ki, const int32_t* nKi, int Li);
/* the NL_MF_G method [nu, k] (src/interpolate.jl:62-80) into G / G0 / Gbare / Sigma / Sigma0 of this context; clamp != 0 extends
 * the input beyond its frequency mesh by its edge values as interpolate_solver! does for Sigma (src/interpolate.jl:176-186) */
int  fdga_interpolate_green(fdga_ctx*, int which, const fdga_c64* host_in, int nG_in, int Li, int clamp);
/* State updates of the outer loop of solve_using_mfRG! (src/mfRG.jl:217-372), device resident:
 * fdga_mix_bubbles:      Pi_mixed = mixing * Pi + (1 - mixing) * Pi0;  set!(S.Pi, Pi_mixed)                       (:271-276)
 * fdga_update_reference: set!(S.Pi0, Pi_mixed); set!(S.G0, S.G); set!(S.Sigma0, S.Sigma); add!(S.F0, S.F); set!(S.F, 0) (:336-347) */
int  fdga_mix_bubbles(fdga_ctx*, double mixing);
int  fdga_update_reference(fdga_ctx*);

/* ---- the vertex as a callable ---------------------------------------------------------------
 * F(W, v, w, P, k, q, Ch, Sp; F0, gamma_p, gamma_t, gamma_a) of the chain S.F from `level` (0: S.F, 1: S.F0, ...) at n points:
 * src/vertex.jl:209-336, src/nonlocal/vertex.jl:69-211, src/nonlocal_2/vertex.jl:207-259, src/boson_exchange.jl:349-560.
 * W, v, w: Matsubara indices (v, w may be FDGA_INF_FREQ); iP, ik, iq: linear momentum indices; swave_kq != 0: k = q = kSW;
 * flags: bit 0 F0, bit 1 gamma_p, bit 2 gamma_t, bit 3 gamma_a.  Synchronous utility (tests, conversions), not a hot path. */
#define FDGA_INF_FREQ (1 << 28)
int  fdga_eval_vertex(fdga_ctx*, int level, int channel, int spin, int flags, int swave_kq, int64_t n, const int32_t* W, const int32_t* v,
                      const int32_t* w, const int32_t* iP, const int32_t* ik, const int32_t* iq, fdga_c64* out);

/* ---- CUDA graphs ------------------------------------------------------------------------
 * Record a sequence of calls once, replay it with one launch: a step of the iteration is ~80 small dependent kernels on three
 * concurrent lanes, and the replay removes the per-launch gaps between them.
 *   fdga_graph_begin(ctx);  <any sequence of asynchronous fdga_* calls: fdga_unstash_F, fdga_unflatten_F (pinned host memory),
 *                            fdga_iterate_solver, fdga_sde, fdga_bse_*, fdga_flatten_F_async, ...>;  fdga_graph_end(ctx, &id);
 *   fdga_graph_launch(ctx, id);   ...   fdga_graph_destroy(ctx, id);
 * Nothing executes while recording.  Calls that synchronise with the host (fdga_sync, fdga_get_*, fdga_set_*, the Krylov drivers)
 * are not recordable.  The library refreshes derived tables lazily, so the recorded calls must form a steady-state cycle: run the
 * sequence once eagerly first; fdga_graph_end fails if the lazy state at its end differs from the one at fdga_graph_begin, and
 * fdga_graph_launch fails ("record it again") if the context is not in that state or device tables were rebuilt since
 * (fdga_set_symmetry_classes, fdga_set_option, fdga_comm_init).  Host pointers passed while recording are baked in.
 * Single-rank contexts only. */
int  fdga_graph_begin(fdga_ctx*);
int  fdga_graph_end(fdga_ctx*, int* graph_id);
int  fdga_graph_launch(fdga_ctx*, int graph_id);
int  fdga_graph_destroy(fdga_ctx*, int graph_id);

/* ---- introspection --------------------------------------------------------------------- */
/* accumulated device time (CUDA events on the launching stream) and launch counts per kernel id */
enum { FDGA_T_CACHE = 0, FDGA_T_L_K2 = 1, FDGA_T_L_K3 = 2, FDGA_T_K1 = 3, FDGA_T_K2 = 4, FDGA_T_K3 = 5,
       FDGA_T_SDE_L = 6, FDGA_T_SDE_RS = 7, FDGA_T_SDE_U2 = 8, FDGA_T_BUBBLE = 9, FDGA_T_RIGHT = 10,
       FDGA_T_SWAVE = 11, FDGA_T_EXPAND = 12, FDGA_T_MISC = 13, FDGA_T_COMM = 14,
       FDGA_T_COLUMN_K2 = 15,   /* the column_kernel launches of BSE_K2! alone (a sub-interval of FDGA_T_K2) */
       FDGA_T_KRYLOV = 16,      /* vector kernels of fdga_mfrg_dqgmres (orthogonalisation sweep, direction / iterate update) */
       FDGA_T_COUNT = 17 };
/* FP64 FMA throughput of the device measured with a DFMA micro-benchmark (TFLOP/s, 2 flop per FMA): roofline denominator */
int  fdga_measure_fp64_peak(fdga_ctx*, double* tflops);
int  fdga_profile_enable(fdga_ctx*, int on);
int  fdga_profile_reset(fdga_ctx*);
int  fdga_kernel_time_ms(fdga_ctx*, int kernel_id, double* ms, int64_t* launches);
int64_t fdga_total_launches(fdga_ctx*);
/* the stream all kernels of this context are launched on (cudaStream_t as void*) */
void* fdga_stream(fdga_ctx*);

#ifdef __cplusplus
}
#endif
#endif /* FDGA_H */
