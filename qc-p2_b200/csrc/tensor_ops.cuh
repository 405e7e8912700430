// tensor_ops.cuh -- bandwidth-bound tensor kernels: permutation (transpose) with axpby,
// elementwise amplitude algebra, denominators, gathers and deterministic reductions.
#pragma once
#include "common.cuh"

namespace qcp2 {

// Class maps of pair-packed matrices (ccsd_sym.cu): pair q belongs to class cls[q] and sits at position loc[q] of that
// class's block; `want` selects the class a kernel works on.  Null cls: one class, natural order.
struct PairColMap {
    const int *cls = nullptr, *loc = nullptr;
    int want = 0;
};
// the packed [i<j][a<b] accumulator of the T2 equation as one block per class (finalize_t2)
struct LadderBlocks {
    const double *lp[3] = {nullptr, nullptr, nullptr};
    long long ld[3] = {0, 0, 0};
    const int *rowcls = nullptr, *rowloc = nullptr, *colcls = nullptr, *colloc = nullptr;// null: one class, natural order
};

// out[o0..o_{r-1}] = beta*out + alpha*in[sum_d o_d * istride[d]]   (out dense row-major)
void permute_axpby(Ctx &ctx, double *out, const double *in, int rank, const long long *out_ext,
                   const long long *in_stride_per_out_dim, double alpha, double beta);
// out = alpha*x + beta*out (n elements)
void axpby(Ctx &ctx, double *out, const double *x, double alpha, double beta, long long n);
void fill_zero(Ctx &ctx, double *p, long long n);

// integrals.cpp / mp2.cpp
void to_so(Ctx &ctx, const double *aa, const double *bb, const double *ab, int n, double *so);
void slice_block(Ctx &ctx, const double *so, int n2, int o, int v, const char *spec, double *out);
// spin-orbital blocks straight from the spatial MO tensors (aa, bb, ab as transform_to_so takes them), bit-identical to
// slice_ints(transform_to_so(...)) -- the (2n)^4 tensor is never built
void so_block(Ctx &ctx, const double *aa, const double *bb, const double *ab, int n, int o, int v, const char *spec, double *out);
// pair-packed rows: dst[row][q(e<f)] = <pairs[2 row] + m0, pairs[2 row + 1] + m1 || e + m2, f + m2>, e < f < d23
void so_block_packed(Ctx &ctx, const double *aa, const double *bb, const double *ab, int n, int m0, int m1, int m2, int d23,
                     const int *pairs01, int n01, double *dst, long long ld, PairColMap cm = PairColMap());
// res[0..3] as antisymmetry_defect_async for (a,b) and (e,f); res[4] = max |x| over the spin-forbidden elements
void spin_pattern_defect_async(Ctx &ctx, const double *x, const long long d[4], const int off[4], unsigned long long *res);
void so_vvvv_defect_async(Ctx &ctx, const double *aa, const double *bb, const double *ab, int n, int o, int v, unsigned long long *res);
void make_denom(Ctx &ctx, const double *F_spin, int n2, int o, int v, double *denom);
void make_so_moes(Ctx &ctx, const double *ea, const double *eb, int n, double *F);
void mp2_amplitudes_energy(Ctx &ctx, const double *oovv, const double *denom, long long n, double *T, double *e_dev);
void make_fock(Ctx &ctx, const double *e1, const double *e2, int n1, int n2, double *F);

// cc.cpp elementwise pieces
void make_tau(Ctx &ctx, const double *Ts, const double *Td, int o, int v, bool bar, double *out);
void multiply_Ts(Ctx &ctx, const double *Ts, int o, int v, double *out);
// out[j,b,n,f] = -1/2 Td[j,n,f,b] - Ts[j,f] Ts[n,b]   (cc.cpp:247 + :249 as one operand); Ts == nullptr: -1/2 Td only
void wmbej_amplitudes(Ctx &ctx, const double *Ts, const double *Td, int o, int v, double *out);
void add_offdiag(Ctx &ctx, double *F, const double *Fadd, int n);// F(a,e) += (1-d_ae) Fadd(a,e)
void make_D_ia(Ctx &ctx, const double *foo, const double *fvv, int o, int v, double *D);
void make_D_ijab(Ctx &ctx, const double *foo, const double *fvv, int o, int v, double *D);
void make_D_triples(Ctx &ctx, const double *foo, const double *fvv, int o, int v, double *D);
void divide_D_ia(Ctx &ctx, const double *r, const double *foo, const double *fvv, int o, int v, double *out);
void divide_D_ijab(Ctx &ctx, const double *r, const double *foo, const double *fvv, int o, int v, double *out);
void divide_elementwise(Ctx &ctx, const double *r, const double *D, long long n, double *out);
void ccsd_energy(Ctx &ctx, const double *Ts, const double *Td, const double *oovv, const double *fov, int o, int v,
                 double *e_dev);
// DIIS building blocks: out_dev[j] = <e_new, E[j]>, j < nvec; out = sum_k c_dev[k] X[k]
void dots(Ctx &ctx, const double *e_new, const double *E, long long stride, long long n, int nvec, double *out_dev);
void combine(Ctx &ctx, const double *X, long long stride, int nvec, const double *c_dev, long long n, double *out);
void extract_diag(Ctx &ctx, const double *F, int n, double *d);
// r[i,j,a,b] += Rt[i,a,b,j] - Rt[j,a,b,i] - Rt[i,b,a,j] + Rt[j,b,a,i]   (P(ij)P(ab) of one ring product)
void ring_combine(Ctx &ctx, const double *Rt, int o, int v, double *r);
// r = (init ? init : r) + U - U[j,i,a,b] (pair 0) or - U[i,j,b,a] (pair 1); all [o,o,v,v]
void antisym_add(Ctx &ctx, const double *U, const double *init, int o, int v, int pair, double *r);
// out[i,a,m,e] = Ts[i,e] Ts[m,a]
void ring_tt(Ctx &ctx, const double *Ts, int o, int v, double *out);

// dst[row][q(e<f)] = src[a_row][b_row][e][f] for the rows (a<b) listed in pairs01 (2 ints per row); src is [d01,d01,d23,d23]
// dst rows are `ld` doubles apart (ld >= d23 (d23-1)/2; an even ld keeps the rows 16-byte aligned for the GEMM's vector copies)
void pack_pairs(Ctx &ctx, const double *src, int d01, int d23, const int *pairs01, int n01, double *dst, long long ld,
                PairColMap cm = PairColMap());

// packed-pair pieces of the fused CCSD iteration (i<j rows, a<b columns; `ld` = row stride, even)
void tau_pack(Ctx &ctx, const double *Ts, const double *Td, int o, int v, const int *pairs_o, int npo, double *dst, long long ld,
              bool bar = false, PairColMap cm = PairColMap());// rows = any list of (i,j) pairs; bar: tau-bar (Stanton eq. 10)
void pair_trace_add(Ctx &ctx, const double *G, long long ldg, int o, PairColMap all, double *F);
// Rt: [i,a,j,b]; RtS (may be null): matrix [pos(i,a)][pos(j,b)] with row stride ldr over class-sorted composites
void ov_sort4(Ctx &ctx, const double *src, const long long d[4], int ro, int rv, int co, int cv, int V, const int *pos, long long ld,
              double alpha, double *out);
void finalize_t2(Ctx &ctx, const double *oovv, const double *Uab, const double *Uij, const double *Rt, const double *RtS,
                 const int *ovpos, long long ldr, const LadderBlocks &lb,
                 const double *foo, const double *fvv, const double *D, int o, int v, double *out);
void wmnij_pack(Ctx &ctx, const double *W, int o, const int *pairs_o, int npo, double *dst, long long ld);
void wmnij_unpack_add(Ctx &ctx, const double *WP, long long ld, int o, const int *list, int np, double alpha, double *W);
void z_t_sub(Ctx &ctx, const double *Z, long long ldz, const double *Ts, int o, int v, int nrows, double *LP, long long ld, PairColMap cm,
             PairColMap am);

// max |x + P x| over an index-pair transposition, relative to max|x| -> host values
void antisymmetry_defect(Ctx &ctx, const double *x, long long d0, long long d1, long long d2, long long d3, int pair,
                         double *defect, double *maxabs);
void antisymmetry_defect_async(Ctx &ctx, const double *x, long long d0, long long d1, long long d2, long long d3, int pair,
                               unsigned long long *res);

}// namespace qcp2
