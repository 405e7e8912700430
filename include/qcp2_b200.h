/* qcp2_b200.h -- C ABI of the B200 (sm_100a) post-Hartree-Fock hot path of ajay-mk/QC-P2.
 *
 * The reference has no FFI: its boundary is the set of free C++ functions declared in
 * src/integrals.h:26-34, src/mp2.h:23-27 and src/cc.h:26-55, exchanging dense row-major
 * btas::Tensor<double> objects.  Each entry point below replaces the arithmetic of one of
 * those functions; tensors cross the boundary as `const double*` + extents, row-major, last
 * index fastest, exactly the memory image of the reference's DTensor::data().
 *
 * Conventions
 *  - every call returns 0 on success, non-zero on failure; qcp2_last_error() gives the text.
 *    There is no CPU fallback: without a usable CUDA device qcp2_create() fails.
 *  - pointer mode (qcp2_set_pointer_mode): QCP2_HOST (default) = tensor arguments are host
 *    pointers, the call copies in, computes on the GPU and copies out (drop-in mode);
 *    QCP2_DEVICE = tensor arguments are device pointers on the context's device and stay
 *    resident (pipeline mode).  Scalar outputs (double*, int*) are always host pointers.
 *  - o = number of occupied spin orbitals (reference `no`), v = number of virtual spin
 *    orbitals (`nv`), n = number of spatial AOs/MOs, n2 = 2n.
 *  - `ints` is an array of 11 pointers in the order of struct int_struct
 *    (integrals.h:11-23): oooo, ovoo, oovo, ooov, ovov, ovvo, oovv, vovv, ovvv, vvvo, vvvv.
 *  - `inter` is an array of 6 pointers in the order of struct cc_intermediates
 *    (cc.h:14-16): F_ae[v,v], F_mi[o,o], F_me[o,v], W_mnij[o,o,o,o], W_mbej[o,v,v,o],
 *    W_abef[v,v,v,v].
 *  - f_oo[o,o], f_ov[o,v], f_vv[v,v] are moes.F_ii, moes.F_ia, moes.F_aa (cc.h:17-19);
 *    f_ov may be NULL, meaning zeros (the only defined reading of the UHF path, cc.cpp:62-67).
 */
#ifndef QCP2_B200_H
#define QCP2_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qcp2_ctx qcp2_ctx;

enum { QCP2_HOST = 0, QCP2_DEVICE = 1 };

/* (T) evaluation modes.  AUTO checks the permutational antisymmetry of T2, oovv, vovv and
 * ovoo on the device and picks SYMMETRIC when it holds, LITERAL otherwise. */
enum { QCP2_T_AUTO = 0, QCP2_T_LITERAL = 1, QCP2_T_SYMMETRIC = 2 };

/* ---- context ------------------------------------------------------------------------ */
/* number of CUDA devices visible to the process (0 without a driver / device) */
int qcp2_device_count(void);
int qcp2_create(int device, qcp2_ctx **ctx);
int qcp2_destroy(qcp2_ctx *ctx);
int qcp2_set_pointer_mode(qcp2_ctx *ctx, int mode);
/* stream is a cudaStream_t; NULL = the context's own stream.  Device-pointer calls are
 * enqueued on it and are synchronous only where a host scalar is returned. */
int qcp2_set_stream(qcp2_ctx *ctx, void *stream);
int qcp2_synchronize(qcp2_ctx *ctx);
const char *qcp2_last_error(const qcp2_ctx *ctx);
const char *qcp2_version(void);
/* number of kernels this library has launched on the context since creation */
long long qcp2_kernel_launches(const qcp2_ctx *ctx);

/* ---- integrals.h -------------------------------------------------------------------- */
/* transform_ao_mo (integrals.cpp:95-134): (pq|rs)[n,n,n,n] -> (ij|kl)[n,n,n,n];
 * indices k,l are rotated by C1, indices i,j by C2; C1, C2 are [AO][MO] row-major n x n. */
int qcp2_transform_ao_mo(qcp2_ctx *ctx, const double *pq_rs, const double *C1, const double *C2, int n,
                         double *ij_kl);
/* transform_to_so (integrals.cpp:136-167): three [n]^4 spatial tensors -> <ij||kl>[2n]^4 */
int qcp2_transform_to_so(qcp2_ctx *ctx, const double *mo_aa, const double *mo_bb, const double *mo_ab, int n,
                         double *so_ints);
/* slice_ints (integrals.cpp:184-207): spec is 4 characters over {o,v} */
int qcp2_slice_ints(qcp2_ctx *ctx, const double *so_ints, int n2, int o, int v, const char *spec, double *out);
/* make_denom (integrals.cpp:169-181): F_spin[n2,n2] -> denom[o,o,v,v] */
int qcp2_make_denom(qcp2_ctx *ctx, const double *F_spin, int n2, int o, int v, double *denom);

/* ---- mp2.h -------------------------------------------------------------------------- */
/* make_so_moes (mp2.cpp:11-21): eps_a[n], eps_b[n] -> F_spin[n2,n2] (diagonal, zero elsewhere) */
int qcp2_make_so_moes(qcp2_ctx *ctx, const double *eps_a, const double *eps_b, int n, double *F_spin);
/* run_mp2 (mp2.cpp:23-42): T[o,o,v,v] = oovv/denom (T may be NULL), *energy = 1/4 sum oovv^2/denom */
int qcp2_run_mp2(qcp2_ctx *ctx, const double *oovv, const double *denom, int o, int v, double *T, double *energy);
/* MP2 (mp2.cpp:45-86) from the hot-path boundary on: AO integrals + SCF coefficients and
 * orbital energies -> so_ints[n2^4] (may be NULL), T[o,o,v,v] (may be NULL), E_MP2.
 * RHF: pass Cb = Ca and eps_b = eps_a (one transform, as mp2.cpp:55); uhf != 0 performs the
 * three transforms of mp2.cpp:71-73 exactly as written. */
int qcp2_mp2(qcp2_ctx *ctx, const double *ao_ints, const double *Ca, const double *Cb, const double *eps_a,
             const double *eps_b, int n, int o, int v, int uhf, double *so_ints, double *T, double *energy);

/* ---- cc.h --------------------------------------------------------------------------- */
/* make_fock (cc.cpp:25-34): (n2-n1) x (n2-n1) diagonal block of the spin-orbital Fock matrix */
int qcp2_make_fock(qcp2_ctx *ctx, const double *eps1, const double *eps2, int n1, int n2, double *F);
/* make_tau / make_tau_bar (cc.cpp:125-156): bar != 0 selects tau-bar */
int qcp2_make_tau(qcp2_ctx *ctx, const double *Ts, const double *Td, int o, int v, int bar, double *tau);
/* multiply_Ts (cc.cpp:158-172): out[j,f,n,b] = Ts[j,f] Ts[n,b] */
int qcp2_multiply_Ts(qcp2_ctx *ctx, const double *Ts, int o, int v, double *out);
/* make_D_ia / make_D_ijab / make_D_triples (cc.cpp:73-123) */
int qcp2_make_D_ia(qcp2_ctx *ctx, const double *f_oo, const double *f_vv, int o, int v, double *D);
int qcp2_make_D_ijab(qcp2_ctx *ctx, const double *f_oo, const double *f_vv, int o, int v, double *D);
int qcp2_make_D_triples(qcp2_ctx *ctx, const double *f_oo, const double *f_vv, int o, int v, double *D);
/* update_intermediates (cc.cpp:175-255) */
int qcp2_update_intermediates(qcp2_ctx *ctx, const double *Ts, const double *Td, const double *const *ints,
                              const double *f_oo, const double *f_ov, const double *f_vv, int o, int v,
                              double *const *inter);
/* make_T1 (cc.cpp:257-292).  D_ia[o,v] is the reference's explicit denominator argument; when it
 * is NULL the denominators are rebuilt from the diagonals of f_oo, f_vv (make_D_ia). */
int qcp2_make_T1(qcp2_ctx *ctx, const double *Ts, const double *Td, const double *const *ints,
                 const double *const *inter, const double *f_oo, const double *f_ov, const double *f_vv,
                 const double *D_ia, int o, int v, double *T1_new);
/* make_T2 (cc.cpp:294-387).  Either D_ijab[o,o,v,v] or (f_oo, f_vv) must be given. */
int qcp2_make_T2(qcp2_ctx *ctx, const double *Ts, const double *Td, const double *const *ints,
                 const double *const *inter, const double *f_oo, const double *f_vv, const double *D_ijab, int o, int v,
                 double *T2_new);
/* ccsd_energy (cc.cpp:390-405) */
int qcp2_ccsd_energy(qcp2_ctx *ctx, const double *Ts, const double *Td, const double *oovv, const double *f_ov, int o,
                     int v, double *energy);
/* CCSD (cc.cpp:409-451): Jacobi iteration from T1 = 0, T2 = T2_guess, device resident.
 * Either `so_ints` ([n2]^4, sliced on the device as get_integrals cc.cpp:8-23) or `ints`
 * (11 pre-sliced blocks) must be given.  Outputs: T1[o,v], T2[o,o,v,v], *energy (the last
 * printed E_CC), *iters (index of the iteration that met |dE| < conv, or maxiter), e_hist
 * (maxiter doubles, host, may be NULL), sliced_out (11 pointers, may be NULL) receives the
 * sliced blocks (cc_results.sliced_ints). */
int qcp2_ccsd(qcp2_ctx *ctx, const double *so_ints, int n2, const double *const *ints, const double *f_oo,
              const double *f_ov, const double *f_vv, const double *T2_guess, int o, int v, int maxiter, double conv,
              double *T1, double *T2, double *energy, int *iters, double *e_hist, double *const *sliced_out);
/* seconds spent on the device in the last qcp2_ccsd call's iteration loop (CUDA events) */
double qcp2_ccsd_last_loop_seconds(const qcp2_ctx *ctx);

/* CCSD_T (cc.cpp:454-560): perturbative triples energy.  eps_o[o], eps_v[v] are the
 * diagonals of moes.F_ii / moes.F_aa. */
int qcp2_ccsd_t(qcp2_ctx *ctx, const double *T1, const double *T2, const double *oovv, const double *vovv,
                const double *ovoo, const double *eps_o, const double *eps_v, int o, int v, int mode,
                double *energy);
/* Sharded form for one rank of a multi-GPU run: evaluates the triples
 * {first, first+stride, first+2*stride, ...} < qcp2_ccsd_t_count(o, mode) of the mode's
 * enumeration (SYMMETRIC: lexicographic i<j<k; LITERAL: all ordered (i,j,k)).  Writes each
 * triple's weighted contribution to E(T) into e_triples[t] (host array of
 * qcp2_ccsd_t_count() doubles, untouched entries stay as they are; may be NULL) and the
 * partial sum into *energy.  mode must be LITERAL or SYMMETRIC here. */
int qcp2_ccsd_t_range(qcp2_ctx *ctx, const double *T1, const double *T2, const double *oovv, const double *vovv,
                      const double *ovoo, const double *eps_o, const double *eps_v, int o, int v, int mode,
                      long long first, long long stride, long long max_count, double *e_triples, double *energy);
long long qcp2_ccsd_t_count(int o, int mode);
/* returns the mode AUTO would select (device check of the antisymmetries) in *mode */
int qcp2_ccsd_t_select_mode(qcp2_ctx *ctx, const double *T2, const double *oovv, const double *vovv,
                            const double *ovoo, int o, int v, int *mode);

/* ---- multi-GPU: one process (rank) per GPU, NCCL over NVLink 5 / NVSwitch -------------
 * The reference is single-process; these entry points are the sharded form of CCSD_T
 * (cc.cpp:454-560) of SURVEY.md section 8e.  NCCL is bound at run time (libnccl.so.2).
 * Bootstrap: rank 0 calls qcp2_comm_get_unique_id and hands the 128 bytes to the other ranks
 * by any host channel (pipe, file, MPI, torch.distributed ...); every rank then calls
 * qcp2_comm_init on its own context (collective). */
#define QCP2_COMM_ID_BYTES 128
int qcp2_comm_get_unique_id(void *id, int id_bytes);
int qcp2_comm_init(qcp2_ctx *ctx, const void *id, int nranks, int rank);
int qcp2_comm_destroy(qcp2_ctx *ctx);
int qcp2_comm_rank(const qcp2_ctx *ctx, int *rank, int *nranks); /* (0, 1) without a communicator */
/* stream-ordered collectives on DEVICE buffers of doubles (no-ops / local copy on a single rank) */
int qcp2_comm_broadcast(qcp2_ctx *ctx, double *dev_buf, long long n, int root);
int qcp2_comm_allreduce_sum(qcp2_ctx *ctx, double *dev_buf, long long n);
int qcp2_comm_allgather(qcp2_ctx *ctx, const double *dev_send, double *dev_recv, long long n_per_rank);
/* CCSD_T across the communicator (collective; every rank calls it with the same o, v, mode,
 * root).  Rank `root` supplies the tensors (host or device pointers per its pointer mode); the
 * other ranks may pass NULL.  The inputs are replicated by ncclBroadcast, rank r evaluates the
 * triples {r, r+W, r+2W, ...} of the mode's enumeration, one ncclAllReduce combines the
 * per-triple energy vector and every rank returns the same, fixed-order sum in *energy:
 * bit-identical to qcp2_ccsd_t for any number of ranks.  e_triples (host, qcp2_ccsd_t_count()
 * doubles) may be NULL.  The ROOT's `mode` decides for everyone (it travels in the broadcast header), and a
 * failure on any rank before the payload broadcasts makes every rank return status 1 together.
 * Host buffers should be pinned (cudaHostAlloc / cudaHostRegister): pageable memory is staged synchronously by
 * the driver, which serialises the <vo||vv> upload with the compute it is meant to overlap.
 * Replicated host inputs: the other ranks MAY pass the seven tensors as well (host pointers; they must hold the root's values --
 * typically every rank maps one shared-memory copy).  When every rank does, each uploads the small tensors itself and the
 * <vo||vv> slabs are dealt over the ranks (slab i is uploaded and packed by rank i mod W over its own PCIe link and broadcast
 * from there), so the host-to-device traffic is spread over W links instead of the root's one.  Same energy, bit for bit.
 * QCP2_T_ROOT_UPLOAD=1 keeps the root-only upload. */
int qcp2_ccsd_t_mgpu(qcp2_ctx *ctx, const double *T1, const double *T2, const double *oovv, const double *vovv,
                     const double *ovoo, const double *eps_o, const double *eps_v, int o, int v, int mode, int root,
                     double *e_triples, double *energy);

/* Fused MP2 -> CCSD -> (T) across the communicator (collective; same n, o, v, uhf, stages, maxiter, conv, root
 * on every rank).  Rank `root` supplies the boundary inputs of qcp2_post_hf (host or device pointers per its pointer
 * mode), the others may pass NULL.  The inputs are broadcast; AO->MO, <pq||rs>, MP2 and the slicing run replicated
 * (milliseconds); in every CCSD iteration each product above 2e9 flop (the packed particle-particle ladder, the ring
 * and W_mbej products, tau.<am||ef>) is split over the ranks along its LONGER free dimension -- the ladder over the
 * a<b columns of <vv||vv>, so the big operand is streamed once in total -- each rank writes its slab into a W x R
 * row workspace and ONE in-place ncclAllGather re-assembles the product on every rank (csrc/dgemm.cu); everything
 * else is replicated, so all ranks hold identical amplitudes; the (T) triples are dealt round-robin and combined with
 * one ncclAllReduce.  Every rank returns the same energies.  If any rank fails before the first exchange (bad
 * arguments on the root, staging out of memory) every rank returns status 1 -- no rank is left inside a collective.
 * QCP2_CCSD_NO_SHARD=1 keeps the iteration replicated.  qcp2_sharded_gemms: products split so far on this context. */
int qcp2_post_hf_mgpu(qcp2_ctx *ctx, const double *ao_ints, const double *Ca, const double *Cb, const double *eps_a,
                      const double *eps_b, int n, int o, int v, int uhf, int stages, int maxiter, double conv, int root,
                      double *e_mp2, double *e_ccsd, double *e_t, int *iters, double *e_hist, double *T1, double *T2);
long long qcp2_sharded_gemms(const qcp2_ctx *ctx);
/* enable != 0: qcp2_ccsd becomes collective over the context's communicator (every rank must pass IDENTICAL inputs)
 * and splits the large products of each iteration as qcp2_post_hf_mgpu does */
int qcp2_set_ccsd_sharding(qcp2_ctx *ctx, int enable);

/* Resident (T) plan: packs/repacks the inputs once (device pointers required) so repeated
 * range evaluations (bench steps, multi-rank shards) reuse them. */
typedef struct qcp2_t_plan qcp2_t_plan;
int qcp2_t_plan_create(qcp2_ctx *ctx, const double *T1, const double *T2, const double *oovv, const double *vovv,
                       const double *ovoo, const double *eps_o, const double *eps_v, int o, int v, int mode,
                       qcp2_t_plan **plan);
int qcp2_t_plan_run(qcp2_t_plan *plan, long long first, long long stride, long long max_count,
                    double *e_triples_dev, double *energy);
/* device time of the GEMM kernels / all kernels of the last qcp2_t_plan_run (seconds, CUDA events) */
int qcp2_t_plan_timings(const qcp2_t_plan *plan, double *gemm_seconds, double *total_seconds, long long *gemm_launches,
                        long long *all_launches);
/* which kernel the plan's triples GEMM runs on (1 = the TMA/mbarrier kernel), the logical column count of X
 * (v(v-1)/2 or v^2), its padded row stride and the triples per launch; any pointer may be NULL */
int qcp2_t_plan_info(const qcp2_t_plan *plan, int *uses_tma_kernel, long long *ncols, long long *row_stride, int *triples_per_batch);
int qcp2_t_plan_destroy(qcp2_t_plan *plan);

/* Fused pipeline main.cpp:76-88 runs after the SCF: MP2 -> CCSD -> (T) with every tensor resident
 * on the device between the stages (the stepwise entry points above each return host tensors, as
 * the reference's functions do: so_ints is 7 GB at CH4/cc-pVTZ).  Inputs are the hot-path boundary
 * data of qcp2_mp2; f_oo/f_ov/f_vv are built as make_moe_tensors does (cc.cpp:36-70: diagonal
 * spin-orbital Fock blocks, F_ia = 0).  stages: 1 = MP2, 2 = +CCSD, 3 = +(T).  Optional outputs
 * (may be NULL): e_hist[maxiter], T1[o,v], T2[o,o,v,v]. */
int qcp2_post_hf(qcp2_ctx *ctx, const double *ao_ints, const double *Ca, const double *Cb, const double *eps_a,
                 const double *eps_b, int n, int o, int v, int uhf, int stages, int maxiter, double conv, double *e_mp2,
                 double *e_ccsd, double *e_t, int *iters, double *e_hist, double *T1, double *T2);

/* ---- building blocks exported for tests, benchmarks and other hosts ------------------ */
/* BTAS-style contraction (the reference's only arithmetic primitive, e.g. cc.cpp:194):
 * C[ic] = beta*C[ic] + alpha * sum A[ia]*B[ib]; labels are single characters. */
int qcp2_contract(qcp2_ctx *ctx, double alpha, const double *A, const long long *extA, const char *ia,
                  const double *B, const long long *extB, const char *ib, double beta, double *C,
                  const long long *extC, const char *ic);
/* row-major FP64 GEMM on the DMMA kernel: C[M,N] = alpha*op(A)*op(B) + beta*C,
 * transA == 0: A is [M,K] (lda >= K), else A is [K,M]; transB == 0: B is [K,N], else [N,K]. */
int qcp2_dgemm(qcp2_ctx *ctx, int transA, int transB, int M, int N, int K, double alpha, const double *A,
               long long lda, const double *B, long long ldb, double beta, double *C, long long ldc);
/* The same product with device pointers only, a strided batch, a forced kernel choice and device timing (tests and
 * tuning tools): tile < 0 lets the library choose; 0..7 force the warp-specialised TMA kernel's CTA tile
 * (80x256, 48x256, 16x256, 64x64, 112x128, 16x192, 32x256, 96x128), +16 runs the transposed problem C^T = B^T A^T on that tile; k_splits > 0
 * forces the split-K factor; the product is launched `reps` times and *seconds (may be NULL) receives the mean device
 * time of one launch.  Operands that are not 16-byte aligned take the cp.async kernel regardless of `tile`. */
int qcp2_dgemm_ex(qcp2_ctx *ctx, int transA, int transB, int M, int N, int K, double alpha, const double *A, long long lda,
                  const double *B, long long ldb, double beta, double *C, long long ldc, int batch, long long strideA,
                  long long strideB, long long strideC, int tile, int k_splits, int reps, double *seconds);
/* Record of the products the library launches (measurement only).  enable != 0 starts / clears the record; inside
 * qcp2_ccsd / qcp2_post_hf the record is restarted at every iteration, so after the call it holds the products of ONE
 * iteration.  qcp2_gemm_log_read copies up to cap records of 8 values {M, N, K, batch, beta != 0, tile_m (negative:
 * cp.async kernel), tile_n, split_k} and returns the number of records held. */
int qcp2_gemm_log(qcp2_ctx *ctx, int enable);
long long qcp2_gemm_log_read(const qcp2_ctx *ctx, long long *recs, long long cap);
/* slice_ints(transform_to_so(aa, bb, ab), o, v, spec) without the (2n)^4 tensor: the o/v block is gathered straight from
 * the spatial MO tensors (integrals.cpp:136-167 + 184-207 fused; bit-identical values).  out has the block's extents. */
int qcp2_so_block(qcp2_ctx *ctx, const double *aa, const double *bb, const double *ab, int n, int o, int v, const char *spec,
                  double *out);
/* Amplitude restart (opt-in).  The reference's CCSD() always starts from T1 = 0 and the MP2 doubles
 * (cc.cpp:417-421).  A host that has saved the T1 / T2 a previous qcp2_ccsd / qcp2_post_hf returned can resume: pass the saved
 * T2 as `T2_guess` and register the saved T1 here (host or device pointer per the pointer mode; copied); the NEXT CCSD solve
 * of this context starts from it, after which the setting is cleared.  T1 == NULL cancels.  `P2` does this with
 * QCP2_AMPLITUDES=<file>: it resumes from the file if it matches (o, v) and rewrites it after the solve. */
int qcp2_ccsd_set_start_T1(qcp2_ctx *ctx, const double *T1, int o, int v);
/* Opt-in DIIS for the CCSD loops of this context (qcp2_ccsd, qcp2_post_hf*): nvec = 0 (default) is the reference's plain
 * Jacobi iteration (cc.cpp:426-447: same trajectory, same iteration count); 2..12 extrapolates over the last nvec Jacobi
 * steps (Pulay; error vector = amplitude change).  Same fixed point, fewer iterations.  The environment variable
 * QCP2_CCSD_DIIS=nvec does the same for hosts that cannot call this (the P2 driver). */
int qcp2_set_ccsd_diis(qcp2_ctx *ctx, int nvec);
/* Formulation the last CCSD solve of this context ran (qcp2_ccsd, qcp2_post_hf*): 0 = the dense term list of cc.cpp (inputs
 * without the permutational symmetries of RHF-type integrals, e.g. the as-written UHF path), 1 = pair-packed / factorised
 * (ccsd_sym.cu), 2 = pair-packed with one block per spin class (the inputs carry the zero pattern of transform_to_so,
 * integrals.cpp:143-161); -1 before the first solve. */
int qcp2_ccsd_last_mode(const qcp2_ctx *ctx);
/* High-water mark (bytes) of the context's workspace pool since the start of the last qcp2_post_hf / qcp2_post_hf_mgpu
 * call: the peak device memory the fused pipeline needed (caller-owned device buffers not included); -1 if unknown. */
long long qcp2_pool_high_water(const qcp2_ctx *ctx);
/* products run so far on the warp-specialised TMA kernel (return value) and on the cp.async kernel (*legacy) */
long long qcp2_gemm_counts(const qcp2_ctx *ctx, long long *legacy);
/* counter-based synthetic (T) inputs of SURVEY.md section 8d, generated on the device
 * (device pointers regardless of pointer mode) */
int qcp2_synth_t_inputs(qcp2_ctx *ctx, int o, int v, unsigned long long seed, double *T1, double *T2, double *oovv,
                        double *vovv, double *ovoo, double *eps_o, double *eps_v);

#ifdef __cplusplus
}
#endif
#endif /* QCP2_B200_H */
