// ccsd_sym.cu -- one CCSD iteration (cc.cpp:442-446) for inputs that carry the permutational symmetries of real
// (RHF-type) spin-orbital integrals: <pq||rs> antisymmetric in (p,q) and in (r,s), T2 antisymmetric in (i,j) and (a,b).
// ccsd_dev (post_hf.cu) checks every symmetry used here once, on the device, and otherwise keeps the dense term list
// (the as-written UHF integrals lose the ket antisymmetry).
//
// Same equations (Stanton eqs. 1-8, cited line by line below), fewer executed flops and bytes:
//  * every sum over an antisymmetric index pair runs over e<f / m<n only, every result that is antisymmetric in
//    (i,j) and (a,b) is formed on i<j rows and a<b columns only: tau, <ab||ef>, <mn||ef>, <am||ef> are pair-packed;
//  * the particle-particle ladder is factorised (W_abef, v^4 doubles, is never built), and everything that multiplies
//    tau -- the ladder, its two W_abef corrections and the W_mnij term -- is ONE product of packed tau with the
//    row-concatenated operand [<ab||ef>; <am||ef>; <mn||ef>] plus three small products;
//  * the t1.t1 parts of W_mbej (cc.cpp:249) and of the ring terms (cc.cpp:360-363) are contracted one t1 at a time
//    (2 o^3 v^2 flops each instead of an (ov)^3 product);
//  * W_mbej is held as [m,e,b,j], the order its only consumer reads in place; results whose index order interleaves the
//    free indices of the two operands are written through the GEMM's offset-table epilogue (no re-layout pass);
//  * where a contraction wants T2 or tau-bar with the LAST index free, the antisymmetry t[..,a,f] = -t[..,f,a] supplies
//    that order for free (cc.cpp:196, :277);
//  * the T2 equation is assembled in one pass (finalize_t2): <ij||ab> + P(ab) U_ab + P(ij) U_ij + P(ij)P(ab) ring +
//    unpacked ladder family, divided by D_ijab.
#include <cstring>
#include <vector>

#include "contract.cuh"
#include "dgemm.cuh"
#include "post_hf.cuh"
#include "tensor_ops.cuh"

namespace qcp2 {

namespace {
TView int_view4(const IntsDev &I, int k, int o, int v) {
    TView t;
    t.p = const_cast<double *>(I.p[k]);
    t.rank = 4;
    for (int d = 0; d < 4; ++d) t.ext[d] = (kIntNames[k][d] == 'o') ? o : v;
    return t;
}
long long even(long long n) { return (n + 1) & ~1LL; }
}// namespace

bool sym_setup(Ctx &ctx, const IntsDev &I, const double *T2_guess, int o, int v, SymPack &sp, const SoSource *src) {
    const long long O = o, V = v;
    if (o < 2 || v < 2 || getenv("QCP2_CCSD_DENSE")) return false;
    Trace tr(ctx, "ccsd sym_setup");
    unsigned long long *chk = static_cast<unsigned long long *>(ctx.alloc_bytes(16 * 12));
    QCP2_CUDA(cudaMemsetAsync(chk, 0, 16 * 12, ctx.stream));
    int nchk = 0;
    auto check = [&](const double *x, long long d0, long long d1, long long d2, long long d3, int pair) {
        antisymmetry_defect_async(ctx, x, d0, d1, d2, d3, pair, chk + 2 * nchk++);
    };
    const bool direct = I.p[I_VVVV] == nullptr;// <ab||ef> is evaluated from the spatial MO tensors, never stored unpacked
    QCP2_REQUIRE(!direct || (src && src->aa && src->bb && src->ab), "ccsd: <vv||vv> missing and no spatial integrals to build it from");
    if (direct) so_vvvv_defect_async(ctx, src->aa, src->bb, src->ab, src->n, o, v, chk), nchk = 2;
    else check(I.p[I_VVVV], V, V, V, V, 0), check(I.p[I_VVVV], V, V, V, V, 1);// <ab||ef>: a<b rows, e<f columns
    check(I.p[I_OOVV], O, O, V, V, 0), check(I.p[I_OOVV], O, O, V, V, 1);// <mn||ef>: m<n, e<f
    check(I.p[I_VOVV], V, O, V, V, 1);                                  // <am||ef>: e<f
    check(I.p[I_OOOO], O, O, O, O, 0), check(I.p[I_OOOO], O, O, O, O, 1);// W_mnij: m<n, i<j
    check(I.p[I_OOOV], O, O, O, V, 0);                                  // W_mnij's t.<mn||ie> terms: m<n
    check(T2_guess, O, O, V, V, 0), check(T2_guess, O, O, V, V, 1);     // T2 (hence tau, tau-bar): i<j, a<b
    unsigned long long h[24];
    QCP2_CUDA(cudaMemcpyAsync(h, chk, 16 * 12, cudaMemcpyDeviceToHost, ctx.stream));
    ctx.sync();
    ctx.free(chk);
    for (int q = 0; q < nchk; ++q) {
        double defect, mx;
        memcpy(&defect, &h[2 * q], 8), memcpy(&mx, &h[2 * q + 1], 8);
        if (!(defect <= 1e-10 * mx)) return false;
    }
    tr.mark("permutational symmetry checks");
    // pair lists: i<j | a<b | all (a,m) | all (i,n)
    std::vector<int> pr;
    for (int i = 0; i < o; ++i)
        for (int j = i + 1; j < o; ++j) pr.push_back(i), pr.push_back(j);
    const size_t n_po = pr.size();
    for (int a = 0; a < v; ++a)
        for (int b = a + 1; b < v; ++b) pr.push_back(a), pr.push_back(b);
    const size_t n_pv = pr.size();
    for (int a = 0; a < v; ++a)
        for (int m = 0; m < o; ++m) pr.push_back(a), pr.push_back(m);
    const size_t n_am = pr.size();
    for (int i = 0; i < o; ++i)
        for (int n = 0; n < o; ++n) pr.push_back(i), pr.push_back(n);
    sp.pairs_buf = DBuf(ctx, (pr.size() * sizeof(int) + 7) / 8);
    int *pd = reinterpret_cast<int *>(sp.pairs_buf.p);
    QCP2_CUDA(cudaMemcpyAsync(pd, pr.data(), pr.size() * sizeof(int), cudaMemcpyHostToDevice, ctx.stream));
    ctx.sync();// pr is pageable
    sp.pairs_o = pd, sp.pairs_all = pd + n_am;
    sp.npo = (int) (n_po / 2), sp.npv = (int) ((n_pv - n_po) / 2);
    sp.ld = even(sp.npv), sp.ldo = even(sp.npo);
    // row-concatenated right operand of the tau product: <ab||ef> | <am||ef> | <mn||ef>, each segment starting on an even
    // row so that the matching column segments of the result stay 16-byte aligned
    sp.offZ = even(sp.npv), sp.offY = sp.offZ + even(V * O), sp.ntot = sp.offY + sp.npo, sp.ldl = even(sp.ntot);
    sp.ladB = DBuf(ctx, (size_t) (sp.ntot * sp.ld));
    QCP2_CUDA(cudaMemsetAsync(sp.ladB.p, 0, (size_t) (sp.ntot * sp.ld) * 8, ctx.stream));
    if (direct) so_block_packed(ctx, src->aa, src->bb, src->ab, src->n, o, o, o, v, pd + n_po, sp.npv, sp.ladB.p, sp.ld);
    else pack_pairs(ctx, I.p[I_VVVV], v, v, pd + n_po, sp.npv, sp.ladB.p, sp.ld);
    pack_pairs(ctx, I.p[I_VOVV], o, v, pd + n_pv, (int) (V * O), sp.ladB.p + sp.offZ * sp.ld, sp.ld);// row (a,m) = a*o + m
    pack_pairs(ctx, I.p[I_OOVV], o, v, pd, sp.npo, sp.ladB.p + sp.offY * sp.ld, sp.ld);
    sp.oovvA = DBuf(ctx, (size_t) (O * O * sp.ld));
    pack_pairs(ctx, I.p[I_OOVV], o, v, sp.pairs_all, (int) (O * O), sp.oovvA.p, sp.ld);
    sp.tauoP = DBuf(ctx, (size_t) (sp.npo * sp.ld));
    sp.ovvo_mebj = DBuf(ctx, (size_t) (O * V * V * O));
    {
        const long long ext[4] = {O, V, V, O}, istr[4] = {V * V * O, O, V * O, 1};// <mb||ej> read as [m,e,b,j]
        permute_axpby(ctx, sp.ovvo_mebj.p, I.p[I_OVVO], 4, ext, istr, 1.0, 0.0);
    }
    tr.mark("pack <vv||vv>, <vo||vv>, <oo||vv> to pairs");
    return true;
}

void ccsd_sym_iteration(Ctx &ctx, const IntsDev &I, const double *foo, const double *fov, const double *fvv, int o, int v,
                        const SymPack &sp, const double *T1, const double *T2, double *T1n, double *T2n) {
    const long long O = o, V = v, o2v2 = O * O * V * V, ld = sp.ld, ldo = sp.ldo, ldl = sp.ldl;
    const int npo = sp.npo, npv = sp.npv;
    const TView ts(T1, {O, V}), td(T2, {O, O, V, V}), tn(T1n, {O, V});
    const TView oooo = int_view4(I, I_OOOO, o, v), ooov = int_view4(I, I_OOOV, o, v), oovv = int_view4(I, I_OOVV, o, v),
                ovvv = int_view4(I, I_OVVV, o, v), ovov = int_view4(I, I_OVOV, o, v), oovo = int_view4(I, I_OOVO, o, v),
                vvvo = int_view4(I, I_VVVO, o, v), ovoo = int_view4(I, I_OVOO, o, v);
    const double *oovvP = sp.ladB.p + sp.offY * ld;

    // ---------------------------------------------------------------- amplitude-derived operands of this iteration
    DBuf taub(ctx, (size_t) o2v2), taubA(ctx, (size_t) (O * O * ld)), tt(ctx, (size_t) o2v2), lf(ctx, (size_t) o2v2);
    make_tau(ctx, T1, T2, o, v, true, taub);                                  // tau-bar, dense (F_ae)
    tau_pack(ctx, T1, T2, o, v, sp.pairs_o, npo, sp.tauoP, ld);               // tau(T1 old, T2), i<j, a<b
    tau_pack(ctx, T1, T2, o, v, sp.pairs_all, (int) (O * O), taubA, ld, true);// tau-bar, all (i,n), e<f (F_mi)
    wmbej_amplitudes(ctx, nullptr, T2, o, v, tt);                             // -1/2 t_jnfb as [j,b,n,f]
    {
        const long long ext[4] = {O, V, O, V}, istr[4] = {O * V * V, V, V * V, 1};// t[i,m,a,e] read as [i,a,m,e]
        permute_axpby(ctx, lf.p, T2, 4, ext, istr, 1.0, 0.0);
    }
    const TView Lf(lf.p, {O, V, O, V});

    // ---------------------------------------------------------------- intermediates, cc.cpp:175-255 (Stanton eqs. 3-8)
    DBuf fae(ctx, (size_t) (V * V)), fmi(ctx, (size_t) (O * O)), fme(ctx, (size_t) (O * V)), wmnij(ctx, (size_t) (O * O * O * O)),
            wm(ctx, (size_t) o2v2);
    const TView Fae(fae.p, {V, V}), Fmi(fmi.p, {O, O}), Fme(fme.p, {O, V}), Wmnij(wmnij.p, {O, O, O, O}), Wm(wm.p, {O, V, V, O});
    fill_zero(ctx, fae.p, V * V);
    if (fov) contract(ctx, -0.5, TView(fov, {O, V}), "me", ts, "ma", 1.0, Fae, "ae");          // :194
    contract(ctx, 1.0, ts, "mf", ovvv, "mafe", 1.0, Fae, "ae");                                // :195
    // :196  -1/2 taub_mnaf <mn||ef> = +1/2 taub_mnfa <mn||ef>: with a last, tau-bar is read in place as [(m,n,f)][a]
    contract(ctx, 0.5, TView(taub.p, {O, O, V, V}), "mnfa", oovv, "mnef", 1.0, Fae, "ae");
    add_offdiag(ctx, fae.p, fvv, v);                                                          // :198-202

    fill_zero(ctx, fmi.p, O * O);
    if (fov) contract(ctx, 0.5, ts, "ie", TView(fov, {O, V}), "me", 1.0, Fmi, "mi");           // :207
    contract(ctx, 1.0, ts, "ne", ooov, "mnie", 1.0, Fmi, "mi");                                // :208
    {// :209  1/2 sum_{n,ef} taub_inef <mn||ef> = sum_n sum_{e<f}: the n = n' trace of [(i,n)] x [(m,n')] on e<f pairs
        DBuf gg(ctx, (size_t) (O * O * even(O * O)));
        GemmArgs g;
        g.M = (int) (O * O), g.N = (int) (O * O), g.K = npv;
        g.A = taubA, g.lda = ld, g.B = sp.oovvA, g.ldb = ld, g.C = gg, g.ldc = even(O * O);
        gemm(ctx, g, true, false);
        pair_trace_add(ctx, gg, even(O * O), o, fmi.p);
    }
    add_offdiag(ctx, fmi.p, foo, o);                                                          // :212-216

    contract(ctx, 1.0, ts, "nf", oovv, "mnef", 0.0, Fme, "me");                                // :220
    if (fov) axpby(ctx, fme.p, fov, 1.0, 1.0, O * V);                                          // :221-225

    contract(ctx, 1.0, ts, "je", ooov, "mnie", 1.0, Wmnij, "mnij", oooo.p);                    // :229 + :232
    contract(ctx, -1.0, ts, "ie", ooov, "mnje", 1.0, Wmnij, "mnij");                           // :230
    {// :231  1/4 tau_ijef <mn||ef> = 1/2 sum_{e<f} on [i<j][m<n], unpacked with both signs
        DBuf yo(ctx, (size_t) (npo * ldo));
        GemmArgs g;
        g.M = npo, g.N = npo, g.K = npv;
        g.A = sp.tauoP, g.lda = ld, g.B = oovvP, g.ldb = ld, g.C = yo, g.ldc = ldo;
        gemm(ctx, g, true, false);
        wmnij_unpack_add(ctx, yo, ldo, o, 0.5, wmnij.p);
    }
    {// W_mbej as [m,e,b,j]
        contract(ctx, 1.0, ts, "jf", ovvv, "mbef", 1.0, Wm, "mebj", sp.ovvo_mebj.p);           // :245 + :250
        DBuf p2(ctx, (size_t) (O * O * V * O));
        const TView P2(p2.p, {O, O, V, O});
        contract(ctx, 1.0, ts, "jf", oovv, "mnef", 1.0, P2, "mnej", I.p[I_OOVO]);              // <mn||ej> + sum_f t_jf <mn||ef>
        contract(ctx, -1.0, ts, "nb", P2, "mnej", 1.0, Wm, "mebj");                            // :246 + the t.t part of :249
        contract(ctx, 1.0, TView(tt.p, {O, V, O, V}), "jbnf", oovv, "mnef", 1.0, Wm, "mebj");  // :247
    }

    // ---------------------------------------------------------------- T1 equation, cc.cpp:257-292 (Stanton eq. 1)
    {
        DBuf r1(ctx, (size_t) (O * V));
        const TView R(r1.p, {O, V});
        contract(ctx, 1.0, ts, "ie", Fae, "ae", 0.0, R, "ia");                                 // :272
        contract(ctx, -1.0, ts, "ma", Fmi, "mi", 1.0, R, "ia");                                // :273
        contract(ctx, 1.0, Lf, "iame", Fme, "me", 1.0, R, "ia");                               // :274 (t_imae read as [i,a,m,e])
        contract(ctx, -1.0, ts, "nf", ovov, "naif", 1.0, R, "ia");                             // :275
        contract(ctx, -0.5, td, "imef", ovvv, "maef", 1.0, R, "ia");                           // :276
        contract(ctx, 0.5, td, "mnea", oovo, "nmei", 1.0, R, "ia");                            // :277 with t_mnae = -t_mnea
        if (fov) axpby(ctx, r1.p, fov, 1.0, 1.0, O * V);                                       // :280
        divide_D_ia(ctx, r1, foo, fvv, o, v, T1n);                                             // :283-288
    }

    // ---------------------------------------------------------------- T2 equation, cc.cpp:294-387 (Stanton eq. 2):
    // new T1, old T2, old intermediates (cc.cpp:442-446)
    DBuf uab(ctx, (size_t) o2v2), uij(ctx, (size_t) o2v2), d1(ctx, (size_t) (V * V)), d2(ctx, (size_t) (O * O));
    const TView Uab(uab.p, {O, O, V, V}), Uij(uij.p, {O, O, V, V}), D1(d1.p, {V, V}), D2(d2.p, {O, O});
    // U_ab: the terms that enter with P(ab):  t_ijae (F_be - 1/2 t_mb F_me)  [:316-328]  - t_ma <mb||ij>  [:370-371]
    contract(ctx, 1.0, tn, "mb", Fme, "me", 0.0, D1, "be");                                    // :322
    axpby(ctx, d1.p, fae.p, 1.0, -0.5, V * V);
    contract(ctx, -1.0, tn, "ma", ovoo, "mbij", 0.0, Uab, "ijab");                             // :370
    contract(ctx, 1.0, td, "ijae", D1, "be", 1.0, Uab, "ijab");                                // :316 + :324
    // U_ij: the terms that enter with P(ij):  -t_imab (F_mj + 1/2 t_je F_me)  [:333-343]  + t_ie <ab||ej>  [:366-367]
    contract(ctx, 1.0, tn, "je", Fme, "me", 0.0, D2, "mj");                                    // :339
    axpby(ctx, d2.p, fmi.p, 1.0, 0.5, O * O);
    contract(ctx, -1.0, td, "imab", D2, "mj", 0.0, Uij, "ijab");                               // :333 + :340
    contract(ctx, 1.0, tn, "ie", vvvo, "abej", 1.0, Uij, "ijab");                              // :366

    // everything that multiplies tau = tau(new T1, T2), on [i<j] rows:
    //   [ LP | Z | Y ] = tauP . [ <ab||ef> ; <am||ef> ; <mn||ef> ]^T      (sums over e<f carry the factor 2)
    //   LP += 1/2 Y . tau_oldP                1/8 tau <mn||ef> tau_old     (cc.cpp:239 inside :348)
    //   LP += W_mnijP . tauP                  1/2 tau_mnab W_mnij          (cc.cpp:346)
    //   LP -= P(ab) Z . t_old                 -1/2 P(ab) tau <am||ef> t    (cc.cpp:237-238 inside :348)
    DBuf tauP(ctx, (size_t) (npo * ld)), lpz(ctx, (size_t) (npo * ldl)), wp(ctx, (size_t) (npo * ldo)), xab(ctx, (size_t) (npo * V * V));
    tau_pack(ctx, T1n, T2, o, v, sp.pairs_o, npo, tauP, ld);
    {
        GemmArgs g;
        g.M = npo, g.N = (int) sp.ntot, g.K = npv;
        g.A = tauP, g.lda = ld;       // [M][K]
        g.B = sp.ladB, g.ldb = ld;    // [N][K]
        g.C = lpz, g.ldc = ldl, g.c_owns_rows = true;
        gemm(ctx, g, true, false);
        double *LP = lpz.p, *Z = lpz.p + sp.offZ, *Y = lpz.p + sp.offY;
        GemmArgs g2;
        g2.M = npo, g2.N = npv, g2.K = npo, g2.alpha = 0.5, g2.beta = 1.0;
        g2.A = Y, g2.lda = ldl;           // [M][K]
        g2.B = sp.tauoP, g2.ldb = ld;     // [K][N]
        g2.C = LP, g2.ldc = ldl;
        gemm(ctx, g2, true, true);
        wmnij_pack(ctx, wmnij.p, o, sp.pairs_o, npo, wp, ldo);
        g2.alpha = 1.0, g2.A = wp, g2.lda = ldo, g2.B = tauP;
        gemm(ctx, g2, true, true);
        GemmArgs gx;// X[p][a][b] = sum_m Z[p][a,m] t_old[m,b], one batch member per pair p
        gx.M = v, gx.N = v, gx.K = o, gx.batch = npo;
        gx.A = Z, gx.lda = O, gx.strideA = ldl;
        gx.B = T1, gx.ldb = V, gx.strideB = 0;
        gx.C = xab, gx.ldc = V, gx.strideC = V * V;
        gemm(ctx, gx, true, true);
        packed_antisym_sub(ctx, xab, npo, v, LP, ldl);
    }
    // ring terms cc.cpp:354-363: Rt[i,a,b,j] = sum_me t_imae W_mbej + sum_m t_ma (sum_e t_ie <mb||je>), used at four relabelings
    DBuf rt(ctx, (size_t) o2v2), p1(ctx, (size_t) (O * O * V * O));
    const TView Rt(rt.p, {O, V, V, O}), P1(p1.p, {O, O, V, O});
    contract(ctx, 1.0, tn, "ie", ovov, "mbje", 0.0, P1, "imbj");
    contract(ctx, 1.0, tn, "ma", P1, "imbj", 0.0, Rt, "iabj");                                 // :360-363 (first writer)
    contract(ctx, 1.0, Lf, "iame", Wm, "mebj", 1.0, Rt, "iabj");                               // :354-357
    finalize_t2(ctx, I.p[I_OOVV], uab, uij, rt, lpz, ldl, foo, fvv, nullptr, o, v, T2n);       // :312, P(..) combines, :374-383
}

}// namespace qcp2
