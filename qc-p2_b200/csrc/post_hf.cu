// post_hf.cu -- AO->MO transform, MP2 and the spin-orbital CCSD iteration, device resident.
//
// Each reference contraction (a btas::contract call) is one call of qcp2::contract with the
// same index labels as the cited reference line, so the term list can be audited against
// cc.cpp line by line; the elementwise pieces are the kernels of tensor_ops.cu.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "contract.cuh"
#include "dgemm.cuh"
#include "post_hf.cuh"
#include "tensor_ops.cuh"

namespace qcp2 {

const char *const kIntNames[I_COUNT] = {"oooo", "ovoo", "oovo", "ooov", "ovov", "ovvo",
                                        "oovv", "vovv", "ovvv", "vvvo", "vvvv"};

long long int_block_size(int k, int o, int v) {
    long long n = 1;
    for (int d = 0; d < 4; ++d) n *= (kIntNames[k][d] == 'o') ? o : v;
    return n;
}
long long inter_block_size(int k, int o, int v) {
    const long long O = o, V = v;
    switch (k) {
        case X_FAE: return V * V;
        case X_FMI: return O * O;
        case X_FME: return O * V;
        case X_WMNIJ: return O * O * O * O;
        case X_WMBEJ: return O * V * V * O;
        default: return V * V * V * V;
    }
}

static TView int_view(const IntsDev &I, int k, int o, int v) {
    TView t;
    t.p = const_cast<double *>(I.p[k]);
    t.rank = 4;
    for (int d = 0; d < 4; ++d) t.ext[d] = (kIntNames[k][d] == 'o') ? o : v;
    return t;
}

// ---------------------------------------------------------------------------------
// integrals.cpp:95-134 -- four quarter transformations, 2 n^5 flop each.  Every step contracts
// the LAST index of the current tensor with a coefficient matrix and puts the new index FIRST:
//   Z1[l,p,q,r] = sum_s C1[s,l] (pq|rs)      (:121)        one dense GEMM each:
//   Z2[k,l,p,q] = sum_r C1[r,k] Z1[l,p,q,r]  (:123)        out[n][n^3] = C^T[n][n] . in^T[n][n^3]
//   Z3[j,k,l,p] = sum_q C2[q,j] Z2[k,l,p,q]  (:126)        (A = C as [K][M], B = in as [N][K])
//   out[i,j,k,l] = sum_p C2[p,i] Z3[j,k,l,p] (:129)
// after four rotations the indices are back in (ij|kl) order -- no batching, no re-layout.
// ---------------------------------------------------------------------------------
void transform_ao_mo_dev(Ctx &ctx, const double *ao, const double *C1, const double *C2, int n, double *out) {
    QCP2_REQUIRE(n >= 1 && n <= 1024, "transform_ao_mo: n out of range");
    const long long n3 = (long long) n * n * n, n4 = n3 * n;
    QCP2_REQUIRE(n3 < (1LL << 31), "transform_ao_mo: n^3 exceeds 2^31");
    DBuf x1(ctx, (size_t) n4), x2(ctx, (size_t) n4);
    auto rotate = [&](const double *in, const double *Cm, double *res) {
        GemmArgs g;
        g.M = n, g.N = (int) n3, g.K = n;
        g.A = Cm, g.lda = n;   // [K][M]
        g.B = in, g.ldb = n;   // [N][K]
        g.C = res, g.ldc = n3;
        gemm(ctx, g, false, false);
    };
    rotate(ao, C1, x1);
    rotate(x1, C1, x2);
    rotate(x2, C2, x1);
    rotate(x1, C2, out);
}

// mp2.cpp:45-86 from the AO integrals on.
void mp2_dev(Ctx &ctx, const double *ao, const double *Ca, const double *Cb, const double *ea, const double *eb, int n,
             int o, int v, bool uhf, double *so_ints, double *T, double *e_host) {
    const long long n4 = (long long) n * n * n * n, n2 = 2LL * n;
    QCP2_REQUIRE(o >= 0 && v >= 0 && o + v == n2, "MP2: o + v must equal 2n");
    DBuf so_own;
    if (!so_ints) so_own = DBuf(ctx, (size_t) (16 * n4)), so_ints = so_own;
    {
        DBuf aa(ctx, (size_t) n4);
        transform_ao_mo_dev(ctx, ao, Ca, Ca, n, aa);// mp2.cpp:55 / :71
        if (!uhf) {
            to_so(ctx, aa, aa, aa, n, so_ints);// :58
        } else {
            DBuf bb(ctx, (size_t) n4), ab(ctx, (size_t) n4);
            transform_ao_mo_dev(ctx, ao, Cb, Cb, n, bb);// :72
            transform_ao_mo_dev(ctx, ao, Ca, Cb, n, ab);// :73 (Coeff1 = Ca -> k,l ; Coeff2 = Cb -> i,j)
            to_so(ctx, aa, bb, ab, n, so_ints);         // :75
        }
    }
    DBuf F(ctx, (size_t) (n2 * n2)), oovv(ctx, (size_t) ((long long) o * o * v * v)),
            den(ctx, (size_t) ((long long) o * o * v * v)), e(ctx, 1);
    make_so_moes(ctx, ea, uhf ? eb : ea, n, F);
    slice_block(ctx, so_ints, (int) n2, o, v, "oovv", oovv);
    make_denom(ctx, F, (int) n2, o, v, den);
    mp2_amplitudes_energy(ctx, oovv, den, (long long) o * o * v * v, T, e);
    QCP2_CUDA(cudaMemcpyAsync(e_host, e.p, sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
    ctx.sync();
}

// mp2.cpp:45-86 + get_integrals (cc.cpp:8-23) for the fused pipeline: the (2n)^4 tensor of transform_to_so and the
// v^4 block are never built; each o/v block is gathered straight from the spatial MO tensors (same values, bit for bit)
void mp2_blocks_dev(Ctx &ctx, const double *ao, const double *Ca, const double *Cb, const double *ea, const double *eb, int n, int o,
                    int v, bool uhf, bool want_blocks, DBuf *blocks, IntsDev &I, DBuf spatial[3], SoSource &src, double *T,
                    double *e_host) {
    const long long n4 = (long long) n * n * n * n, n2 = 2LL * n;
    QCP2_REQUIRE(o >= 0 && v >= 0 && o + v == n2, "MP2: o + v must equal 2n");
    spatial[0] = DBuf(ctx, (size_t) n4);
    transform_ao_mo_dev(ctx, ao, Ca, Ca, n, spatial[0]);// mp2.cpp:55 / :71
    src.aa = src.bb = src.ab = spatial[0].p, src.n = n;
    if (uhf) {
        spatial[1] = DBuf(ctx, (size_t) n4), spatial[2] = DBuf(ctx, (size_t) n4);
        transform_ao_mo_dev(ctx, ao, Cb, Cb, n, spatial[1]);// :72
        transform_ao_mo_dev(ctx, ao, Ca, Cb, n, spatial[2]);// :73 (Coeff1 = Ca -> k,l ; Coeff2 = Cb -> i,j)
        src.bb = spatial[1].p, src.ab = spatial[2].p;
    }
    for (int k = 0; k < I_COUNT; ++k) {
        I.p[k] = nullptr;
        if (k == I_VVVV || (!want_blocks && k != I_OOVV)) continue;
        blocks[k] = DBuf(ctx, (size_t) std::max<long long>(int_block_size(k, o, v), 1));
        so_block(ctx, src.aa, src.bb, src.ab, n, o, v, kIntNames[k], blocks[k]);// slice_ints(transform_to_so(..), ..)
        I.p[k] = blocks[k].p;
    }
    DBuf F(ctx, (size_t) (n2 * n2)), den(ctx, (size_t) std::max<long long>((long long) o * o * v * v, 1)), e(ctx, 1);
    make_so_moes(ctx, ea, uhf ? eb : ea, n, F);
    make_denom(ctx, F, (int) n2, o, v, den);
    mp2_amplitudes_energy(ctx, I.p[I_OOVV], den, (long long) o * o * v * v, T, e);
    QCP2_CUDA(cudaMemcpyAsync(e_host, e.p, sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
    ctx.sync();
}

// ---------------------------------------------------------------------------------
// cc.cpp:175-255 -- Stanton eqs. 3-8
// ---------------------------------------------------------------------------------
void update_intermediates_dev(Ctx &ctx, const double *Ts, const double *Td, const IntsDev &I, const double *foo,
                              const double *fov, const double *fvv, int o, int v, const InterDev &X) {
    const long long O = o, V = v, o2v2 = O * O * V * V;
    const bool fused = X.p[X_WABEF] == nullptr;// the fused solver (ccsd_dev); the standalone entry point stays literal
    const TView ts(Ts, {O, V}), td(Td, {O, O, V, V});
    const TView oooo = int_view(I, I_OOOO, o, v), ooov = int_view(I, I_OOOV, o, v), oovo = int_view(I, I_OOVO, o, v),
                oovv = int_view(I, I_OOVV, o, v), vovv = int_view(I, I_VOVV, o, v), ovvv = int_view(I, I_OVVV, o, v);
    DBuf taub(ctx, (size_t) o2v2), tau(ctx, (size_t) o2v2), tt(ctx, (size_t) o2v2);
    make_tau(ctx, Ts, Td, o, v, true, taub);
    make_tau(ctx, Ts, Td, o, v, false, tau);
    if (fused) wmbej_amplitudes(ctx, Ts, Td, o, v, tt);// [j,b,n,f]
    else multiply_Ts(ctx, Ts, o, v, tt);
    const TView vtaub(taub.p, {O, O, V, V}), vtau(tau.p, {O, O, V, V}), vtt(tt.p, {O, V, O, V});

    const TView Fae(X.p[X_FAE], {V, V}), Fmi(X.p[X_FMI], {O, O}), Fme(X.p[X_FME], {O, V});
    const TView Wmnij(X.p[X_WMNIJ], {O, O, O, O}), Wmbej(X.p[X_WMBEJ], {O, V, V, O}), Wabef(X.p[X_WABEF], {V, V, V, V});

    fill_zero(ctx, Fae.p, V * V);
    if (fov) contract(ctx, -0.5, TView(fov, {O, V}), "me", ts, "ma", 1.0, Fae, "ae");// :194
    contract(ctx, 1.0, ts, "mf", ovvv, "mafe", 1.0, Fae, "ae");                      // :195
    contract(ctx, -0.5, vtaub, "mnaf", oovv, "mnef", 1.0, Fae, "ae");                // :196
    add_offdiag(ctx, Fae.p, fvv, v);                                                // :198-202

    fill_zero(ctx, Fmi.p, O * O);
    if (fov) contract(ctx, 0.5, ts, "ie", TView(fov, {O, V}), "me", 1.0, Fmi, "mi"); // :207
    contract(ctx, 1.0, ts, "ne", ooov, "mnie", 1.0, Fmi, "mi");                      // :208
    contract(ctx, 0.5, vtaub, "inef", oovv, "mnef", 1.0, Fmi, "mi");                 // :209
    add_offdiag(ctx, Fmi.p, foo, o);                                                // :212-216

    contract(ctx, 1.0, ts, "nf", oovv, "mnef", 0.0, Fme, "me");                      // :220
    if (fov) axpby(ctx, Fme.p, fov, 1.0, 1.0, O * V);                                // :221-225

    contract(ctx, 1.0, ts, "je", ooov, "mnie", 1.0, Wmnij, "mnij", oooo.p);          // :229 (+ :232 <mn||ij> in the epilogue)
    contract(ctx, -1.0, ts, "ie", ooov, "mnje", 1.0, Wmnij, "mnij");                 // :230
    contract(ctx, 0.25, vtau, "ijef", oovv, "mnef", 1.0, Wmnij, "mnij");             // :231

    if (Wabef.p) {// the fused solver never materialises W_abef (see make_T2_dev)
        contract(ctx, -1.0, ts, "mb", vovv, "amef", 1.0, Wabef, "abef", I.p[I_VVVV]);// :237 (+ :240 <ab||ef> in the epilogue)
        contract(ctx, 1.0, ts, "ma", vovv, "bmef", 1.0, Wabef, "abef");              // :238
        contract(ctx, 0.25, vtau, "mnab", oovv, "mnef", 1.0, Wabef, "abef");         // :239
    }

    contract(ctx, 1.0, ts, "jf", ovvv, "mbef", 1.0, Wmbej, "mbej", I.p[I_OVVO]);     // :245 (+ :250 <mb||ej> in the epilogue)
    contract(ctx, -1.0, ts, "nb", oovo, "mnej", 1.0, Wmbej, "mbej");                 // :246
    if (fused) {// :247 + :249 share <mn||ef> and the target: one product with -(1/2 t_jnfb + t_jf t_nb)
        contract(ctx, 1.0, TView(tt.p, {O, V, O, V}), "jbnf", oovv, "mnef", 1.0, Wmbej, "mbej");
    } else {
        contract(ctx, -0.5, td, "jnfb", oovv, "mnef", 1.0, Wmbej, "mbej");           // :247
        contract(ctx, -1.0, vtt, "jfnb", oovv, "mnef", 1.0, Wmbej, "mbej");          // :249
    }
}

// cc.cpp:257-292 -- Stanton eq. 1
void make_T1_dev(Ctx &ctx, const double *Ts, const double *Td, const IntsDev &I, const InterDev &X, const double *foo,
                 const double *fov, const double *fvv, const double *D_ia, int o, int v, double *T1_new) {
    const long long O = o, V = v;
    const TView ts(Ts, {O, V}), td(Td, {O, O, V, V});
    DBuf r(ctx, (size_t) (O * V));
    const TView R(r.p, {O, V});
    contract(ctx, 1.0, ts, "ie", TView(X.p[X_FAE], {V, V}), "ae", 0.0, R, "ia");       // :272
    contract(ctx, -1.0, ts, "ma", TView(X.p[X_FMI], {O, O}), "mi", 1.0, R, "ia");      // :273
    contract(ctx, 1.0, td, "imae", TView(X.p[X_FME], {O, V}), "me", 1.0, R, "ia");     // :274
    contract(ctx, -1.0, ts, "nf", int_view(I, I_OVOV, o, v), "naif", 1.0, R, "ia");    // :275
    contract(ctx, -0.5, td, "imef", int_view(I, I_OVVV, o, v), "maef", 1.0, R, "ia");  // :276
    contract(ctx, -0.5, td, "mnae", int_view(I, I_OOVO, o, v), "nmei", 1.0, R, "ia");  // :277
    if (fov) axpby(ctx, r.p, fov, 1.0, 1.0, O * V);                                    // :280
    if (D_ia) divide_elementwise(ctx, r, D_ia, O * V, T1_new);                         // :283-288
    else divide_D_ia(ctx, r, foo, fvv, o, v, T1_new);
}

// cc.cpp:294-387 -- Stanton eq. 2
// When X.p[X_WABEF] is null the particle-particle ladder 1/2 tau_ijef W_abef is evaluated in
// factorised form from the integrals (W_abef, v^4 doubles, is never built):
//   1/2 tau.<ab||ef>  +  1/8 (tau.<mn||ef>).tau_old  -  1/2 P(ab) (tau.<am||ef>).t_old
// with tau = tau(Ts, Td) and tau_old / t_old the amplitudes the intermediates were built from
// (Ts_old; cc.cpp:442-446 feeds make_T2 the NEW T1 but OLD intermediates).
void make_T2_dev(Ctx &ctx, const double *Ts, const double *Td, const IntsDev &I, const InterDev &X, const double *foo,
                 const double *fvv, const double *D_ijab, int o, int v, double *T2_new, const double *Ts_old) {
    const long long O = o, V = v, o2v2 = O * O * V * V;
    const TView ts(Ts, {O, V}), td(Td, {O, O, V, V});
    const TView Fae(X.p[X_FAE], {V, V}), Fmi(X.p[X_FMI], {O, O}), Fme(X.p[X_FME], {O, V});
    const TView Wmnij(X.p[X_WMNIJ], {O, O, O, O}), Wmbej(X.p[X_WMBEJ], {O, V, V, O}), Wabef(X.p[X_WABEF], {V, V, V, V});
    const TView ovov = int_view(I, I_OVOV, o, v), vvvo = int_view(I, I_VVVO, o, v), ovoo = int_view(I, I_OVOO, o, v);
    DBuf r(ctx, (size_t) o2v2), tau(ctx, (size_t) o2v2), tt(ctx, (size_t) o2v2), d1(ctx, (size_t) (V * V)),
            d2(ctx, (size_t) (O * O));
    const TView R(r.p, {O, O, V, V}), D1(d1.p, {V, V}), D2(d2.p, {O, O});
    make_tau(ctx, Ts, Td, o, v, false, tau);
    if (Wabef.p) multiply_Ts(ctx, Ts, o, v, tt);// literal path only (cc.cpp:360-363)
    const TView vtau(tau.p, {O, O, V, V}), vtt(tt.p, {O, V, O, V});

    const bool fused = Wabef.p == nullptr;// the fused solver; the standalone entry point keeps the literal term list
    DBuf ubuf;
    if (fused) ubuf = DBuf(ctx, (size_t) o2v2);
    const TView U(ubuf.p, {O, O, V, V});
    if (fused) {
        // cc.cpp:316-328: P(ab) t_ijae (F_be - 1/2 t_mb F_me).  The (a<->b) partner of each product is the same
        // product with two labels exchanged, and the two factors share t_ijae: ONE product + antisymmetrised add
        contract(ctx, 1.0, ts, "mb", Fme, "me", 0.0, D1, "be");             // :322
        axpby(ctx, d1.p, Fae.p, 1.0, -0.5, V * V);                          // F_be - 1/2 t_mb F_me
        contract(ctx, 1.0, td, "ijae", D1, "be", 0.0, U, "ijab");           // :316 + :324
        antisym_add(ctx, U.p, I.p[I_OOVV], o, v, 1, r);                     // + <ij||ab> (:312), - (a<->b) (:317 + :328)
        // cc.cpp:333-343: -P(ij) t_imab (F_mj + 1/2 t_je F_me)
        contract(ctx, 1.0, ts, "je", Fme, "me", 0.0, D2, "mj");             // :339 (stored as [m,j])
        axpby(ctx, d2.p, Fmi.p, 1.0, 0.5, O * O);                           // F_mj + 1/2 t_je F_me
        contract(ctx, -1.0, td, "imab", D2, "mj", 0.0, U, "ijab");          // :333 + :340
        antisym_add(ctx, U.p, nullptr, o, v, 0, r);                         // - (i<->j) (:334 + :343)
    } else {
    contract(ctx, 1.0, td, "ijae", Fae, "be", 1.0, R, "ijab", I.p[I_OOVV]);// :316 (+ :312 <ij||ab> in the epilogue)
    contract(ctx, -1.0, td, "ijbe", Fae, "ae", 1.0, R, "ijab");         // :317
    contract(ctx, 1.0, ts, "mb", Fme, "me", 0.0, D1, "be");             // :322
    contract(ctx, -0.5, td, "ijae", D1, "be", 1.0, R, "ijab");          // :324
    contract(ctx, 1.0, ts, "ma", Fme, "me", 0.0, D1, "ae");             // :326-327
    contract(ctx, 0.5, td, "ijbe", D1, "ae", 1.0, R, "ijab");           // :328
    contract(ctx, -1.0, td, "imab", Fmi, "mj", 1.0, R, "ijab");         // :333
    contract(ctx, 1.0, td, "jmab", Fmi, "mi", 1.0, R, "ijab");          // :334
    contract(ctx, 1.0, ts, "je", Fme, "me", 0.0, D2, "jm");             // :339
    contract(ctx, -0.5, td, "imab", D2, "jm", 1.0, R, "ijab");          // :340
    contract(ctx, 1.0, ts, "ie", Fme, "me", 0.0, D2, "im");             // :341-342
    contract(ctx, 0.5, td, "jmab", D2, "im", 1.0, R, "ijab");           // :343
    }
    contract(ctx, 0.5, vtau, "mnab", Wmnij, "mnij", 1.0, R, "ijab");    // :346
    if (Wabef.p) {
        contract(ctx, 0.5, vtau, "ijef", Wabef, "abef", 1.0, R, "ijab");// :348  particle-particle ladder
    } else {
        QCP2_REQUIRE(Ts_old != nullptr, "make_T2: factorised ladder needs the amplitudes the intermediates were built from");
        const TView to(Ts_old, {O, V}), vvvv = int_view(I, I_VVVV, o, v), vovv = int_view(I, I_VOVV, o, v),
                    oovv = int_view(I, I_OOVV, o, v);
        DBuf tau_old(ctx, (size_t) o2v2), y(ctx, (size_t) (O * O * O * O)), z(ctx, (size_t) (O * O * V * O));
        make_tau(ctx, Ts_old, Td, o, v, false, tau_old);
        const TView vtauo(tau_old.p, {O, O, V, V}), Y(y.p, {O, O, O, O}), Z(z.p, {O, O, V, O});
        contract(ctx, 0.5, vtau, "ijef", vvvv, "abef", 1.0, R, "ijab");    // 1/2 tau <ab||ef>         (cc.cpp:240 in :348)
        contract(ctx, 1.0, vtau, "ijef", oovv, "mnef", 0.0, Y, "ijmn");   // 1/8 tau <mn||ef> tau_old  (cc.cpp:239 in :348)
        contract(ctx, 0.125, Y, "ijmn", vtauo, "mnab", 1.0, R, "ijab");
        contract(ctx, 1.0, vtau, "ijef", vovv, "amef", 0.0, Z, "ijam");   // -1/2 P(ab) tau <am||ef> t (cc.cpp:237-238 in :348)
        contract(ctx, -0.5, Z, "ijam", to, "mb", 1.0, R, "ijab");
        contract(ctx, 0.5, Z, "ijbm", to, "ma", 1.0, R, "ijab");
    }
    if (Wabef.p) {// literal: eight (ov)^3 products
        contract(ctx, 1.0, td, "imae", Wmbej, "mbej", 1.0, R, "ijab");      // :354  ring terms
        contract(ctx, -1.0, td, "jmae", Wmbej, "mbei", 1.0, R, "ijab");     // :355
        contract(ctx, -1.0, td, "imbe", Wmbej, "maej", 1.0, R, "ijab");     // :356
        contract(ctx, 1.0, td, "jmbe", Wmbej, "maei", 1.0, R, "ijab");      // :357
        contract(ctx, 1.0, vtt, "iema", ovov, "mbje", 1.0, R, "ijab");      // :360
        contract(ctx, -1.0, vtt, "jema", ovov, "mbie", 1.0, R, "ijab");     // :361
        contract(ctx, -1.0, vtt, "iemb", ovov, "maje", 1.0, R, "ijab");     // :362
        contract(ctx, 1.0, vtt, "jemb", ovov, "maie", 1.0, R, "ijab");      // :363
    } else {
        // fused solver: cc.cpp:354-357 are ONE product Rt[i,a,b,j] = sum_me t_imae W_mbej evaluated at four
        // index relabelings, likewise :360-363 with (t_ie t_ma) <mb||je>; two GEMMs + one P(ij)P(ab) combine
        // Both left factors are laid out as [i,a | m,e] first, so each product is one plain (ov)^3 GEMM.
        DBuf rt(ctx, (size_t) o2v2), lf(ctx, (size_t) o2v2);
        const TView Rt(rt.p, {O, V, V, O}), Lf(lf.p, {O, V, O, V});
        {
            const long long ext[4] = {O, V, O, V}, istr[4] = {O * V * V, V, V * V, 1};// t[i,m,a,e] read as [i,a,m,e]
            permute_axpby(ctx, lf.p, Td, 4, ext, istr, 1.0, 0.0);
        }
        contract(ctx, 1.0, Lf, "iame", Wmbej, "mbej", 0.0, Rt, "iabj");
        ring_tt(ctx, Ts, o, v, lf.p);                                       // t_ie t_ma as [i,a,m,e]
        contract(ctx, 1.0, Lf, "iame", ovov, "mbje", 1.0, Rt, "iabj");
        ring_combine(ctx, rt, o, v, r);
    }
    if (fused) {
        contract(ctx, 1.0, ts, "ie", vvvo, "abej", 0.0, U, "ijab");         // :366
        antisym_add(ctx, U.p, nullptr, o, v, 0, r);                         // - (i<->j) (:367)
        contract(ctx, -1.0, ts, "ma", ovoo, "mbij", 0.0, U, "ijab");        // :370
        antisym_add(ctx, U.p, nullptr, o, v, 1, r);                         // - (a<->b) (:371)
    } else {
    contract(ctx, 1.0, ts, "ie", vvvo, "abej", 1.0, R, "ijab");         // :366
    contract(ctx, -1.0, ts, "je", vvvo, "abei", 1.0, R, "ijab");        // :367
    contract(ctx, -1.0, ts, "ma", ovoo, "mbij", 1.0, R, "ijab");        // :370
    contract(ctx, 1.0, ts, "mb", ovoo, "maij", 1.0, R, "ijab");         // :371
    }
    if (D_ijab) divide_elementwise(ctx, r, D_ijab, o2v2, T2_new);        // :374-383
    else divide_D_ijab(ctx, r, foo, fvv, o, v, T2_new);
}

// cc.cpp:409-451.  Amplitudes, integrals and intermediates never leave the device; the
// host sees one double per iteration (the energy that decides convergence).
void ccsd_dev(Ctx &ctx, const IntsDev &I_in, const double *foo, const double *fov, const double *fvv,
              const double *T2_guess, int o, int v, int maxiter, double conv, double *T1, double *T2, double *e_host,
              int *iters_host, double *e_hist_host, const SoSource *src) {
    const long long O = o, V = v, o2v2 = O * O * V * V;
    IntsDev I = I_in;
    DBuf xbuf[X_COUNT];
    InterDev X;
    for (int k = 0; k < X_COUNT; ++k) {
        X.p[k] = nullptr;
        if (k == X_WABEF) continue;// factorised ladder: the v^4 intermediate is never allocated
        xbuf[k] = DBuf(ctx, (size_t) inter_block_size(k, o, v)), X.p[k] = xbuf[k].p;
    }
    DBuf t1n(ctx, (size_t) (O * V)), t2n(ctx, (size_t) o2v2), e_dev(ctx, 1);
    // Inputs with the permutational symmetries of real (RHF-type) integrals take the pair-packed / factorised
    // iteration of ccsd_sym.cu; the check runs once, on the device (the iteration preserves the antisymmetry of T2:
    // every term enters with its P(ij) / P(ab) partner).  Anything else -- the as-written UHF integrals lose the ket
    // antisymmetry, SURVEY.md Q5 -- keeps the dense term list below.
    Trace tr(ctx, "ccsd_dev");
    SymPack sp;
    const bool sym = sym_setup(ctx, I, T2_guess, fov, o, v, sp, src);
    tr.mark(sym ? (sp.spin_blocked ? "symmetric formulation, spin-blocked pairs: checks + packing" : "symmetric formulation: checks + packing")
                : "dense formulation (symmetry checks)");
    ctx.last_ccsd_mode = sym ? (sp.spin_blocked ? 2 : 1) : 0;
    DBuf vvvv_own;
    if (!sym && I.p[I_VVVV] == nullptr) {// the dense term list reads the v^4 block itself
        QCP2_REQUIRE(src && src->aa, "ccsd: <vv||vv> missing and no spatial integrals to build it from");
        vvvv_own = DBuf(ctx, (size_t) std::max<long long>(V * V * V * V, 1));
        so_block(ctx, src->aa, src->bb, src->ab, src->n, o, v, "vvvv", vvvv_own);
        I.p[I_VVVV] = vvvv_own.p;
        tr.mark("<vv||vv> from the spatial integrals");
    }
    // the integral blocks are constant for the whole loop: their operand re-layouts are made once
    PermCache perm_cache;
    for (int k = 0; k < I_COUNT; ++k)
        if (I.p[k]) perm_cache.constants.push_back(I.p[k]);
    struct CacheGuard {
        Ctx &c;
        PermCache &pc;
        CacheGuard(Ctx &ctx_, PermCache &p) : c(ctx_), pc(p) { c.perm_cache = &pc; }
        ~CacheGuard() {
            c.perm_cache = nullptr;
            for (PermCache::Entry &e : pc.entries) c.free(e.buf);
        }
    } cache_guard(ctx, perm_cache);
    if (getenv("QCP2_CCSD_NO_PERM_CACHE")) ctx.perm_cache = nullptr;
    if (ctx.ccsd_t1_start) {// restart from checkpointed amplitudes (opt-in; the reference always starts from T1 = 0, :417-419)
        QCP2_CUDA(cudaMemcpyAsync(T1, ctx.ccsd_t1_start, (size_t) (O * V) * 8, cudaMemcpyDeviceToDevice, ctx.stream));
        ctx.ccsd_t1_start = nullptr;
    } else {
        fill_zero(ctx, T1, O * V);                                                         // :417-419
    }
    QCP2_CUDA(cudaMemcpyAsync(T2, T2_guess, (size_t) o2v2 * 8, cudaMemcpyDeviceToDevice, ctx.stream));// :421
    // events and the graph executable are released on every exit path (a throw from the body -- e.g. out of memory in a
    // workspace allocation at large v -- must not leak them)
    struct LoopResources {
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        cudaGraphExec_t exec = nullptr;
        cudaStream_t stream;
        explicit LoopResources(cudaStream_t s) : stream(s) {}
        ~LoopResources() {
            if (exec) {
                cudaStreamSynchronize(stream);
                cudaGraphExecDestroy(exec);
            }
            if (ev0) cudaEventDestroy(ev0);
            if (ev1) cudaEventDestroy(ev1);
        }
    } res(ctx.stream);
    QCP2_CUDA(cudaEventCreate(&res.ev0));
    QCP2_CUDA(cudaEventCreate(&res.ev1));
    cudaEvent_t &ev0 = res.ev0, &ev1 = res.ev1;
    QCP2_CUDA(cudaEventRecord(ev0, ctx.stream));
    double e_last = 0.0, e = 0.0;
    int it = 0;
    // One iteration (cc.cpp:442-446) is a fixed sequence of ~90 launches on fixed buffers: from the second iteration
    // on it is replayed as a CUDA graph (stream capture, incl. the stream-ordered workspace allocations), which
    // removes the launch gaps that dominate small molecules.  QCP2_CCSD_NO_GRAPH=1 keeps the eager path.
    auto body = [&]() {
        if (ctx.log_gemms) ctx.gemm_log.clear();// the record holds the products of ONE iteration
        if (sym) {
            ccsd_sym_iteration(ctx, I, foo, fov, fvv, o, v, sp, T1, T2, t1n, t2n);         // :442-446
        } else {
            update_intermediates_dev(ctx, T1, T2, I, foo, fov, fvv, o, v, X);              // :442
            make_T1_dev(ctx, T1, T2, I, X, foo, fov, fvv, nullptr, o, v, t1n);             // :445
            make_T2_dev(ctx, t1n, T2, I, X, foo, fvv, nullptr, o, v, t2n, T1);             // :446: new T1, old T2, old X
        }
        QCP2_CUDA(cudaMemcpyAsync(T1, t1n.p, (size_t) (O * V) * 8, cudaMemcpyDeviceToDevice, ctx.stream));
        QCP2_CUDA(cudaMemcpyAsync(T2, t2n.p, (size_t) o2v2 * 8, cudaMemcpyDeviceToDevice, ctx.stream));
    };
    // ---- opt-in DIIS (qcp2_set_ccsd_diis / QCP2_CCSD_DIIS=n; NOT in the reference, whose loop cc.cpp:426-447 is the plain
    // Jacobi iteration above -- off by default so that trajectories and iteration counts stay the reference's).
    // After each Jacobi step x_k = (T1, T2) with error e_k = x_k - x_{k-1}, the amplitudes are replaced by sum_k c_k x_k,
    // c minimising |sum c_k e_k| under sum c_k = 1 (Pulay), over the last `ndiis` steps.
    // Only the symmetric formulation extrapolates: the as-written UHF equations (dense term list) have more than
    // one fixed point -- measured on NH2/cc-pVDZ, DIIS converges to E_CCSD -0.15354 where the reference's Jacobi loop reaches
    // -0.15740 -- so there the request is ignored and the reference's trajectory kept.
    const int ndiis_req = ctx.ccsd_diis > 0 ? ctx.ccsd_diis : (getenv("QCP2_CCSD_DIIS") ? atoi(getenv("QCP2_CCSD_DIIS")) : 0);
    const int ndiis = sym ? ndiis_req : 0;
    const long long nx = O * V + o2v2;
    DBuf dx, de, dprev, ddots, dcoef;
    std::vector<double> Bm;// ndiis x ndiis Gram matrix of the stored errors (host)
    int dcount = 0;
    if (ndiis >= 2) {
        QCP2_REQUIRE(ndiis <= 12, "CCSD DIIS: at most 12 vectors");
        dx = DBuf(ctx, (size_t) (ndiis * nx)), de = DBuf(ctx, (size_t) (ndiis * nx)), dprev = DBuf(ctx, (size_t) nx);
        ddots = DBuf(ctx, (size_t) ndiis), dcoef = DBuf(ctx, (size_t) ndiis);
        Bm.assign((size_t) ndiis * ndiis, 0.0);
    }
    auto diis_before = [&]() {// remember x_{k-1}
        if (ndiis < 2) return;
        QCP2_CUDA(cudaMemcpyAsync(dprev.p, T1, (size_t) (O * V) * 8, cudaMemcpyDeviceToDevice, ctx.stream));
        QCP2_CUDA(cudaMemcpyAsync(dprev.p + O * V, T2, (size_t) o2v2 * 8, cudaMemcpyDeviceToDevice, ctx.stream));
    };
    auto diis_after = [&]() {
        if (ndiis < 2) return;
        const int slot = dcount % ndiis;
        double *xs = dx.p + (long long) slot * nx, *es = de.p + (long long) slot * nx;
        QCP2_CUDA(cudaMemcpyAsync(xs, T1, (size_t) (O * V) * 8, cudaMemcpyDeviceToDevice, ctx.stream));
        QCP2_CUDA(cudaMemcpyAsync(xs + O * V, T2, (size_t) o2v2 * 8, cudaMemcpyDeviceToDevice, ctx.stream));
        QCP2_CUDA(cudaMemcpyAsync(es, xs, (size_t) nx * 8, cudaMemcpyDeviceToDevice, ctx.stream));
        axpby(ctx, es, dprev.p, -1.0, 1.0, nx);// e = x_new - x_old
        ++dcount;
        const int m = std::min(dcount, ndiis);
        dots(ctx, es, de.p, nx, nx, m, ddots.p);
        std::vector<double> row((size_t) m);
        QCP2_CUDA(cudaMemcpyAsync(row.data(), ddots.p, (size_t) m * 8, cudaMemcpyDeviceToHost, ctx.stream));
        ctx.sync();
        for (int j = 0; j < m; ++j) Bm[(size_t) slot * ndiis + j] = Bm[(size_t) j * ndiis + slot] = row[(size_t) j];
        if (m < 2) return;
        // [B 1; 1 0] [c; lambda] = [0; 1], scaled by the largest diagonal element; Gaussian elimination with partial pivoting
        const int n1 = m + 1;
        std::vector<double> A((size_t) n1 * n1, 0.0), rhs((size_t) n1, 0.0);
        double scale = 0.0;
        for (int j = 0; j < m; ++j) scale = std::max(scale, std::fabs(Bm[(size_t) j * ndiis + j]));
        if (!(scale > 0.0)) return;// converged to the last bit: nothing to extrapolate
        for (int a = 0; a < m; ++a)
            for (int b = 0; b < m; ++b) A[(size_t) a * n1 + b] = Bm[(size_t) a * ndiis + b] / scale;
        for (int a = 0; a < m; ++a) A[(size_t) a * n1 + m] = A[(size_t) m * n1 + a] = 1.0;
        rhs[(size_t) m] = 1.0;
        for (int col = 0; col < n1; ++col) {
            int piv = col;
            for (int r = col + 1; r < n1; ++r)
                if (std::fabs(A[(size_t) r * n1 + col]) > std::fabs(A[(size_t) piv * n1 + col])) piv = r;
            if (std::fabs(A[(size_t) piv * n1 + col]) < 1e-14) return;// singular subspace: keep the Jacobi step
            if (piv != col) {
                for (int c2 = 0; c2 < n1; ++c2) std::swap(A[(size_t) piv * n1 + c2], A[(size_t) col * n1 + c2]);
                std::swap(rhs[(size_t) piv], rhs[(size_t) col]);
            }
            for (int r = col + 1; r < n1; ++r) {
                const double f = A[(size_t) r * n1 + col] / A[(size_t) col * n1 + col];
                for (int c2 = col; c2 < n1; ++c2) A[(size_t) r * n1 + c2] -= f * A[(size_t) col * n1 + c2];
                rhs[(size_t) r] -= f * rhs[(size_t) col];
            }
        }
        std::vector<double> sol((size_t) n1, 0.0);
        for (int r = n1 - 1; r >= 0; --r) {
            double acc = rhs[(size_t) r];
            for (int c2 = r + 1; c2 < n1; ++c2) acc -= A[(size_t) r * n1 + c2] * sol[(size_t) c2];
            sol[(size_t) r] = acc / A[(size_t) r * n1 + r];
        }
        QCP2_CUDA(cudaMemcpyAsync(dcoef.p, sol.data(), (size_t) m * 8, cudaMemcpyHostToDevice, ctx.stream));
        combine(ctx, dx.p, nx, m, dcoef.p, O * V, T1);
        combine(ctx, dx.p + O * V, nx, m, dcoef.p, o2v2, T2);
        ctx.sync();// sol is pageable
    };
    bool use_graph = getenv("QCP2_CCSD_NO_GRAPH") == nullptr && ctx.shard == nullptr;// (NCCL calls stay outside graphs)
    cudaGraphExec_t &graph_exec = res.exec;
    long long graph_launches = 0;
    for (; it < maxiter; ++it) {
        ccsd_energy(ctx, T1, T2, I.p[I_OOVV], fov, o, v, e_dev);                           // :432
        QCP2_CUDA(cudaMemcpyAsync(&e, e_dev.p, 8, cudaMemcpyDeviceToHost, ctx.stream));
        ctx.sync();
        const double de = e - e_last;
        e_last = e;
        if (e_hist_host) e_hist_host[it] = e;
        if (std::fabs(de) < conv) break;                                                   // :438 (fabs)
        diis_before();
        if (!use_graph || it == 0) {// the first iteration runs eagerly: it fills the re-layout cache and configures the kernels
            body();
            diis_after();
            continue;
        }
        if (!graph_exec) {
            const long long l0 = ctx.launches;
            cudaGraph_t graph = nullptr;
            bool ok = cudaStreamBeginCapture(ctx.stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (ok) {
                perm_cache.frozen = true;
                try {
                    body();
                } catch (const std::exception &) {
                    ok = false;
                }
                perm_cache.frozen = false;
                if (cudaStreamEndCapture(ctx.stream, &graph) != cudaSuccess || !graph) ok = false;
            }
            if (ok && cudaGraphInstantiate(&graph_exec, graph, 0) != cudaSuccess) ok = false, graph_exec = nullptr;
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();// a failed capture leaves a sticky-free error behind; the eager path below is always valid
            if (!ok) {
                ctx.launches = l0;
                use_graph = false;
                body();
                diis_after();
                continue;
            }
            graph_launches = ctx.launches - l0;
            ctx.launches = l0;
        }
        QCP2_CUDA(cudaGraphLaunch(graph_exec, ctx.stream));
        ctx.launches += graph_launches;
        diis_after();
    }
    QCP2_CUDA(cudaEventRecord(ev1, ctx.stream));
    QCP2_CUDA(cudaEventSynchronize(ev1));
    tr.mark("iteration loop");
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    ctx.ccsd_loop_seconds = ms * 1e-3;
    *e_host = e;
    if (iters_host) *iters_host = it;
}

}// namespace qcp2
