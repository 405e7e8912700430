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
//  * W_mbej is held as [m,e,j,b], the order its only consumer reads in place; results whose index order interleaves the
//    free indices of the two operands are written through the GEMM's offset-table epilogue (no re-layout pass);
//  * where a contraction wants T2 or tau-bar with the LAST index free, the antisymmetry t[..,a,f] = -t[..,f,a] supplies
//    that order for free (cc.cpp:196, :277);
//  * the T2 equation is assembled in one pass (finalize_t2): <ij||ab> + P(ab) U_ab + P(ij) U_ij + P(ij)P(ab) ring +
//    unpacked ladder family, divided by D_ijab;
//  * spin classes: when every input carries the zero pattern of transform_to_so (integrals.cpp:143-161,
//    <pq||rs> = 0 unless s_p + s_q = s_r + s_s; checked on the device, exact zeros), a pair (x,y) has class s_x + s_y, every
//    pair-packed matrix is block diagonal in the class and only the diagonal blocks are stored and multiplied; the
//    (ov) x (ov) matrices of the ring products are held over class-sorted composites, where each has one non-zero block
//    per block row.  Without the pattern (the dense synthetic benchmark shape) there is one class and nothing changes.
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

bool sym_setup(Ctx &ctx, const IntsDev &I, const double *T2_guess, const double *fov, int o, int v, SymPack &sp, const SoSource *src) {
    const long long O = o, V = v;
    if (o < 2 || v < 2 || getenv("QCP2_CCSD_DENSE")) return false;
    Trace tr(ctx, "ccsd sym_setup");
    constexpr int kSlots = 48;
    unsigned long long *chk = static_cast<unsigned long long *>(ctx.alloc_bytes(8 * kSlots));
    QCP2_CUDA(cudaMemsetAsync(chk, 0, 8 * kSlots, ctx.stream));
    // ---- permutational symmetries the packed formulation relies on: slots [0, 20), (defect, max) pairs
    int nchk = 0;
    auto check = [&](const double *x, long long d0, long long d1, long long d2, long long d3, int pair) {
        antisymmetry_defect_async(ctx, x, d0, d1, d2, d3, pair, chk + 2 * nchk++);
    };
    const bool direct = I.p[I_VVVV] == nullptr;// <ab||ef> is evaluated from the spatial MO tensors, never stored unpacked
    QCP2_REQUIRE(!direct || (src && src->aa && src->bb && src->ab), "ccsd: <vv||vv> missing and no spatial integrals to build it from");
    unsigned long long *vv_forbidden = chk + 44;// slots 40..44 belong to the direct <vv||vv> scan
    if (direct) {
        so_vvvv_defect_async(ctx, src->aa, src->bb, src->ab, src->n, o, v, chk + 40);
    } else {
        check(I.p[I_VVVV], V, V, V, V, 0), check(I.p[I_VVVV], V, V, V, V, 1);// <ab||ef>: a<b rows, e<f columns
    }
    check(I.p[I_OOVV], O, O, V, V, 0), check(I.p[I_OOVV], O, O, V, V, 1);// <mn||ef>: m<n, e<f
    check(I.p[I_VOVV], V, O, V, V, 1);                                  // <am||ef>: e<f
    check(I.p[I_OOOO], O, O, O, O, 0), check(I.p[I_OOOO], O, O, O, O, 1);// W_mnij: m<n, i<j
    check(I.p[I_OOOV], O, O, O, V, 0);                                  // W_mnij's t.<mn||ie> terms: m<n
    check(T2_guess, O, O, V, V, 0), check(T2_guess, O, O, V, V, 1);     // T2 (hence tau, tau-bar): i<j, a<b
    // ---- spin zero pattern (integrals.cpp:143-161): one max|x| over the spin-forbidden elements per tensor, slots [20, 36) + 44
    int nspin = 0;
    const bool want_spin = getenv("QCP2_CCSD_NO_SPIN_BLOCKS") == nullptr;
    if (want_spin) {
        auto spin = [&](const double *x, const char *spec) {
            long long d[4];
            int off[4];
            for (int k = 0; k < 4; ++k) d[k] = spec[k] == 'o' ? O : V, off[k] = spec[k] == 'o' ? 0 : o;
            spin_pattern_defect_async(ctx, x, d, off, chk + 20 + nspin++);
        };
        for (int k = 0; k < I_COUNT; ++k)
            if (I.p[k]) spin(I.p[k], kIntNames[k]);
        spin(T2_guess, "oovv");
        if (fov) {// s_i != s_a must vanish: the same rule with two unit dimensions of even (alpha) parity
            const long long d[4] = {1, O, 1, V};
            const int off[4] = {0, 0, 0, o};
            spin_pattern_defect_async(ctx, fov, d, off, chk + 20 + nspin++);
        }
        if (ctx.ccsd_t1_start) {// a restart T1 must be spin-diagonal as well
            const long long d[4] = {1, O, 1, V};
            const int off[4] = {0, 0, 0, o};
            spin_pattern_defect_async(ctx, ctx.ccsd_t1_start, d, off, chk + 20 + nspin++);
        }
        QCP2_REQUIRE(nspin <= 16, "sym_setup: spin check slots");
    }
    unsigned long long h[kSlots];
    QCP2_CUDA(cudaMemcpyAsync(h, chk, 8 * kSlots, cudaMemcpyDeviceToHost, ctx.stream));
    ctx.sync();
    ctx.free(chk);
    auto as_double = [&](int slot) {
        double x;
        memcpy(&x, &h[slot], 8);
        return x;
    };
    for (int q = 0; q < nchk; ++q)
        if (!(as_double(2 * q) <= 1e-10 * as_double(2 * q + 1))) return false;
    if (direct && (!(as_double(40) <= 1e-10 * as_double(41)) || !(as_double(42) <= 1e-10 * as_double(43)))) return false;
    bool blocked = want_spin;
    for (int q = 0; q < nspin; ++q) blocked = blocked && as_double(20 + q) == 0.0;
    if (direct) blocked = blocked && as_double((int) (vv_forbidden - chk)) == 0.0;
    tr.mark("permutational symmetry + spin pattern checks");

    // ---- pair lists per class and the class / position maps of every natural pair index
    sp.spin_blocked = blocked, sp.ncls = blocked ? 3 : 1;
    auto cls_of = [&](int gx, int gy) { return blocked ? (gx & 1) + (gy & 1) : 0; };// global spin-orbital indices
    std::vector<int> lists[3][4];// per class: rows (i<j), cols (a<b), (a,m), ordered (i,n)
    std::vector<int> rowcls, rowloc, colcls, colloc, amcls, amloc, allcls, allloc;
    auto add = [&](int c, int which, int x, int y, std::vector<int> &mc, std::vector<int> &ml) {
        mc.push_back(c), ml.push_back((int) (lists[c][which].size() / 2));
        lists[c][which].push_back(x), lists[c][which].push_back(y);
    };
    for (int i = 0; i < o; ++i)
        for (int j = i + 1; j < o; ++j) add(cls_of(i, j), 0, i, j, rowcls, rowloc);
    for (int a = 0; a < v; ++a)
        for (int b = a + 1; b < v; ++b) add(cls_of(o + a, o + b), 1, a, b, colcls, colloc);
    for (int a = 0; a < v; ++a)
        for (int m = 0; m < o; ++m) add(cls_of(o + a, m), 2, a, m, amcls, amloc);
    for (int i = 0; i < o; ++i)
        for (int n = 0; n < o; ++n) add(cls_of(i, n), 3, i, n, allcls, allloc);
    sp.npo = (int) rowcls.size(), sp.npv = (int) colcls.size();
    std::vector<int> ovpos((size_t) (O * V), 0);
    if (blocked) {
        auto ovcls = [&](int i, int a) { return (i & 1) == ((o + a) & 1) ? 0 : 1 + (i & 1); };
        long long next = 0;
        for (int c = 0; c < 3; ++c) {
            sp.ov0[c] = next;
            for (int i = 0; i < o; ++i)
                for (int a = 0; a < v; ++a)
                    if (ovcls(i, a) == c) ovpos[(size_t) (i * V + a)] = (int) next++;
            sp.ovn[c] = next - sp.ov0[c];
            next = even(next);
        }
        sp.ldov = even(next);
    }
    std::vector<int> flat;
    size_t list_off[3][4], map_off[8];
    for (int c = 0; c < sp.ncls; ++c)
        for (int w = 0; w < 4; ++w) list_off[c][w] = flat.size(), flat.insert(flat.end(), lists[c][w].begin(), lists[c][w].end());
    const std::vector<int> *maps[8] = {&rowcls, &rowloc, &colcls, &colloc, &amcls, &amloc, &allcls, &allloc};
    for (int k = 0; k < 8; ++k) map_off[k] = flat.size(), flat.insert(flat.end(), maps[k]->begin(), maps[k]->end());
    const size_t ov_off = flat.size();
    flat.insert(flat.end(), ovpos.begin(), ovpos.end());
    sp.ints_buf = DBuf(ctx, (flat.size() * sizeof(int) + 7) / 8 + 1);
    int *pd = reinterpret_cast<int *>(sp.ints_buf.p);
    QCP2_CUDA(cudaMemcpyAsync(pd, flat.data(), flat.size() * sizeof(int), cudaMemcpyHostToDevice, ctx.stream));
    ctx.sync();// flat is pageable
    if (blocked) {
        sp.rowcls = pd + map_off[0], sp.rowloc = pd + map_off[1], sp.colcls = pd + map_off[2], sp.colloc = pd + map_off[3];
        sp.amcls = pd + map_off[4], sp.amloc = pd + map_off[5], sp.allcls = pd + map_off[6], sp.allloc = pd + map_off[7];
        sp.ovpos = pd + ov_off;
        // <mn||ef> as the right operand of the blocked -1/2 t_jnfb <mn||ef> product: rows (n,f), columns (m,e), both class-sorted
        sp.oovvS = DBuf(ctx, (size_t) (sp.ldov * sp.ldov));
        QCP2_CUDA(cudaMemsetAsync(sp.oovvS.p, 0, (size_t) (sp.ldov * sp.ldov) * 8, ctx.stream));
        const long long d[4] = {O, O, V, V};
        ov_sort4(ctx, I.p[I_OOVV], d, /*row (n,f)*/ 1, 3, /*col (m,e)*/ 0, 2, v, sp.ovpos, sp.ldov, 1.0, sp.oovvS.p);
    }
    // ---- per class: the row-concatenated right operand of the tau product, <ab||ef> | <am||ef> | <mn||ef> (each segment
    // starting on an even row, so the matching column segments of the result stay 16-byte aligned), all on e<f of the class
    for (int c = 0; c < sp.ncls; ++c) {
        SymPack::Cls &k = sp.cls[c];
        k.nrow = (int) (lists[c][0].size() / 2), k.ncol = (int) (lists[c][1].size() / 2), k.nam = (int) (lists[c][2].size() / 2),
        k.nall = (int) (lists[c][3].size() / 2);
        k.rows = pd + list_off[c][0], k.cols = pd + list_off[c][1], k.ams = pd + list_off[c][2], k.alls = pd + list_off[c][3];
        k.ldk = even(std::max(k.ncol, 1)), k.ldo = even(std::max(k.nrow, 1)), k.lda = even(std::max(k.nall, 1));
        k.offZ = even(k.ncol), k.offY = k.offZ + even(k.nam), k.ntot = k.offY + k.nrow, k.ldl = even(std::max<long long>(k.ntot, 1));
        if (k.ncol == 0) continue;
        PairColMap cm;
        cm.cls = sp.colcls, cm.loc = sp.colloc, cm.want = c;
        k.ladB = DBuf(ctx, (size_t) (std::max<long long>(k.ntot, 1) * k.ldk));
        QCP2_CUDA(cudaMemsetAsync(k.ladB.p, 0, (size_t) (std::max<long long>(k.ntot, 1) * k.ldk) * 8, ctx.stream));
        if (direct) so_block_packed(ctx, src->aa, src->bb, src->ab, src->n, o, o, o, v, k.cols, k.ncol, k.ladB.p, k.ldk, cm);
        else pack_pairs(ctx, I.p[I_VVVV], v, v, k.cols, k.ncol, k.ladB.p, k.ldk, cm);
        pack_pairs(ctx, I.p[I_VOVV], o, v, k.ams, k.nam, k.ladB.p + k.offZ * k.ldk, k.ldk, cm);// source row (a,m) = a*o + m
        pack_pairs(ctx, I.p[I_OOVV], o, v, k.rows, k.nrow, k.ladB.p + k.offY * k.ldk, k.ldk, cm);
        k.oovvA = DBuf(ctx, (size_t) (std::max(k.nall, 1) * k.ldk));
        pack_pairs(ctx, I.p[I_OOVV], o, v, k.alls, k.nall, k.oovvA.p, k.ldk, cm);
        k.tauoP = DBuf(ctx, (size_t) (std::max(k.nrow, 1) * k.ldk));
    }
    sp.ovvo_mejb = DBuf(ctx, (size_t) (O * V * V * O));
    {
        const long long ext[4] = {O, V, O, V}, istr[4] = {V * V * O, O, 1, V * O};// <mb||ej> read as [m,e,j,b]
        permute_axpby(ctx, sp.ovvo_mejb.p, I.p[I_OVVO], 4, ext, istr, 1.0, 0.0);
    }
    tr.mark(blocked ? "pack <vv||vv>, <vo||vv>, <oo||vv> to pairs, one block per spin class" : "pack <vv||vv>, <vo||vv>, <oo||vv> to pairs");
    return true;
}

void ccsd_sym_iteration(Ctx &ctx, const IntsDev &I, const double *foo, const double *fov, const double *fvv, int o, int v,
                        const SymPack &sp, const double *T1, const double *T2, double *T1n, double *T2n) {
    const long long O = o, V = v, o2v2 = O * O * V * V;
    auto colmap = [&](int c) {
        PairColMap m;
        m.cls = sp.colcls, m.loc = sp.colloc, m.want = c;
        return m;
    };
    const TView ts(T1, {O, V}), td(T2, {O, O, V, V}), tn(T1n, {O, V});
    const TView oooo = int_view4(I, I_OOOO, o, v), ooov = int_view4(I, I_OOOV, o, v), oovv = int_view4(I, I_OOVV, o, v),
                ovvv = int_view4(I, I_OVVV, o, v), ovov = int_view4(I, I_OVOV, o, v), oovo = int_view4(I, I_OOVO, o, v),
                vvvo = int_view4(I, I_VVVO, o, v), ovoo = int_view4(I, I_OVOO, o, v);

    // ---------------------------------------------------------------- amplitude-derived operands of this iteration
    DBuf taub(ctx, (size_t) o2v2), tt(ctx, (size_t) o2v2), lf(ctx, (size_t) o2v2);
    make_tau(ctx, T1, T2, o, v, true, taub);                                  // tau-bar, dense (F_ae)
    for (int c = 0; c < sp.ncls; ++c)                                         // tau(T1 old, T2) on i<j, a<b, one block per class
        if (sp.cls[c].nrow && sp.cls[c].ncol) tau_pack(ctx, T1, T2, o, v, sp.cls[c].rows, sp.cls[c].nrow, sp.cls[c].tauoP, sp.cls[c].ldk, false, colmap(c));
    const bool ring_blocks = sp.spin_blocked && sp.ovpos && !getenv("QCP2_CCSD_DENSE_RING");
    const long long ldr = sp.ldov;
    const long long t2d[4] = {O, O, V, V};
    DBuf lfS, ttS, wS, rtS;
    if (ring_blocks) {
        // spin-blocked ring products: the (ov) x (ov) matrices are held over class-sorted composites (SymPack::ovpos), where
        // t_imae, -1/2 t_jnfb, <mn||ef> and W_mbej are block matrices with ONE non-zero block per block row
        lfS = DBuf(ctx, (size_t) (ldr * ldr)), ttS = DBuf(ctx, (size_t) (ldr * ldr)), wS = DBuf(ctx, (size_t) (ldr * ldr)),
        rtS = DBuf(ctx, (size_t) (ldr * ldr));
        fill_zero(ctx, lfS.p, ldr * ldr), fill_zero(ctx, ttS.p, ldr * ldr), fill_zero(ctx, rtS.p, ldr * ldr);
        ov_sort4(ctx, T2, t2d, /*row (i,a)*/ 0, 2, /*col (m,e)*/ 1, 3, v, sp.ovpos, ldr, 1.0, lfS.p);   // t[i,m,a,e]
        ov_sort4(ctx, T2, t2d, /*row (j,b)*/ 0, 3, /*col (n,f)*/ 1, 2, v, sp.ovpos, ldr, -0.5, ttS.p);  // -1/2 t[j,n,f,b]
    } else {
        wmbej_amplitudes(ctx, nullptr, T2, o, v, tt);                         // -1/2 t_jnfb as [j,b,n,f]
    }
    {
        const long long ext[4] = {O, V, O, V}, istr[4] = {O * V * V, V, V * V, 1};// t[i,m,a,e] read as [i,a,m,e]
        permute_axpby(ctx, lf.p, T2, 4, ext, istr, 1.0, 0.0);
    }
    const TView Lf(lf.p, {O, V, O, V});

    // ---------------------------------------------------------------- intermediates, cc.cpp:175-255 (Stanton eqs. 3-8)
    DBuf fae(ctx, (size_t) (V * V)), fmi(ctx, (size_t) (O * O)), fme(ctx, (size_t) (O * V)), wmnij(ctx, (size_t) (O * O * O * O)),
            wm(ctx, (size_t) o2v2);
    const TView Fae(fae.p, {V, V}), Fmi(fmi.p, {O, O}), Fme(fme.p, {O, V}), Wmnij(wmnij.p, {O, O, O, O}), Wm(wm.p, {O, V, O, V});
    fill_zero(ctx, fae.p, V * V);
    if (fov) contract(ctx, -0.5, TView(fov, {O, V}), "me", ts, "ma", 1.0, Fae, "ae");          // :194
    contract(ctx, 1.0, ts, "mf", ovvv, "mafe", 1.0, Fae, "ae");                                // :195
    // :196  -1/2 taub_mnaf <mn||ef> = +1/2 taub_mnfa <mn||ef>: with a last, tau-bar is read in place as [(m,n,f)][a]
    contract(ctx, 0.5, TView(taub.p, {O, O, V, V}), "mnfa", oovv, "mnef", 1.0, Fae, "ae");
    add_offdiag(ctx, fae.p, fvv, v);                                                          // :198-202

    fill_zero(ctx, fmi.p, O * O);
    if (fov) contract(ctx, 0.5, ts, "ie", TView(fov, {O, V}), "me", 1.0, Fmi, "mi");           // :207
    contract(ctx, 1.0, ts, "ne", ooov, "mnie", 1.0, Fmi, "mi");                                // :208
    for (int c = 0; c < sp.ncls; ++c) {// :209  1/2 sum_{n,ef} taub_inef <mn||ef> = sum_n sum_{e<f}: the n = n' trace of [(i,n)] x [(m,n')]
        const SymPack::Cls &k = sp.cls[c];
        if (!k.nall || !k.ncol) continue;
        DBuf taubA(ctx, (size_t) (k.nall * k.ldk)), gg(ctx, (size_t) (k.nall * k.lda));
        tau_pack(ctx, T1, T2, o, v, k.alls, k.nall, taubA, k.ldk, true, colmap(c));// tau-bar, ordered (i,n) pairs, e<f
        GemmArgs g;
        g.M = k.nall, g.N = k.nall, g.K = k.ncol;
        g.A = taubA, g.lda = k.ldk, g.B = k.oovvA, g.ldb = k.ldk, g.C = gg, g.ldc = k.lda;
        gemm(ctx, g, true, false);
        PairColMap all;
        all.cls = sp.allcls, all.loc = sp.allloc, all.want = c;
        pair_trace_add(ctx, gg, k.lda, o, all, fmi.p);
    }
    add_offdiag(ctx, fmi.p, foo, o);                                                          // :212-216

    contract(ctx, 1.0, ts, "nf", oovv, "mnef", 0.0, Fme, "me");                                // :220
    if (fov) axpby(ctx, fme.p, fov, 1.0, 1.0, O * V);                                          // :221-225

    contract(ctx, 1.0, ts, "je", ooov, "mnie", 1.0, Wmnij, "mnij", oooo.p);                    // :229 + :232
    contract(ctx, -1.0, ts, "ie", ooov, "mnje", 1.0, Wmnij, "mnij");                           // :230
    for (int c = 0; c < sp.ncls; ++c) {// :231  1/4 tau_ijef <mn||ef> = 1/2 sum_{e<f} on [i<j][m<n], unpacked with both signs
        const SymPack::Cls &k = sp.cls[c];
        if (!k.nrow || !k.ncol) continue;
        DBuf yo(ctx, (size_t) (k.nrow * k.ldo));
        GemmArgs g;
        g.M = k.nrow, g.N = k.nrow, g.K = k.ncol;
        g.A = k.tauoP, g.lda = k.ldk, g.B = k.ladB.p + k.offY * k.ldk, g.ldb = k.ldk, g.C = yo, g.ldc = k.ldo;
        gemm(ctx, g, true, false);
        wmnij_unpack_add(ctx, yo, k.ldo, o, k.rows, k.nrow, 0.5, wmnij.p);
    }
    {// W_mbej as [m,e,j,b]
        contract(ctx, 1.0, ts, "jf", ovvv, "mbef", 1.0, Wm, "mejb", sp.ovvo_mejb.p);           // :245 + :250
        DBuf p2(ctx, (size_t) (O * O * V * O));
        const TView P2(p2.p, {O, O, V, O});
        contract(ctx, 1.0, ts, "jf", oovv, "mnef", 1.0, P2, "mnej", I.p[I_OOVO]);              // <mn||ej> + sum_f t_jf <mn||ef>
        contract(ctx, -1.0, ts, "nb", P2, "mnej", 1.0, Wm, "mejb");                            // :246 + the t.t part of :249
        if (!ring_blocks) {
            contract(ctx, 1.0, TView(tt.p, {O, V, O, V}), "jbnf", oovv, "mnef", 1.0, Wm, "mejb");// :247
        } else {
            // :247 per spin block: C[(j,b) r][(m,e) c] = ttS[r][k] . <mn||ef>S[k][c], added into W[(m,e)][(j,b)] (transposed store)
            fill_zero(ctx, wS.p, ldr * ldr);
            const long long wd[4] = {O, V, O, V};
            ov_sort4(ctx, wm.p, wd, /*row (m,e)*/ 0, 1, /*col (j,b)*/ 2, 3, v, sp.ovpos, ldr, 1.0, wS.p);
            static const int rkc[3][3] = {{0, 0, 0}, {1, 2, 1}, {2, 1, 2}};
            for (int q = 0; q < 3; ++q) {
                const int r = rkc[q][0], kk = rkc[q][1], c = rkc[q][2];
                if (!sp.ovn[r] || !sp.ovn[kk] || !sp.ovn[c]) continue;
                GemmArgs g;// computed as the transposed problem so that the store is row-major: W[c rows][r cols] += S[k][c]^T . tt[r][k]^T
                g.M = (int) sp.ovn[c], g.N = (int) sp.ovn[r], g.K = (int) sp.ovn[kk], g.beta = 1.0;
                g.A = sp.oovvS.p + sp.ov0[kk] * ldr + sp.ov0[c], g.lda = ldr;// [K][M]
                g.B = ttS.p + sp.ov0[r] * ldr + sp.ov0[kk], g.ldb = ldr;      // [N][K]
                g.C = wS.p + sp.ov0[c] * ldr + sp.ov0[r], g.ldc = ldr;
                gemm(ctx, g, false, false);
            }
        }
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

    // everything that multiplies tau = tau(new T1, T2), on [i<j] rows, one block per class:
    //   [ LP | Z | Y ] = tauP . [ <ab||ef> ; <am||ef> ; <mn||ef> ]^T      (sums over e<f carry the factor 2)
    //   LP += 1/2 Y . tau_oldP                1/8 tau <mn||ef> tau_old     (cc.cpp:239 inside :348)
    //   LP += W_mnijP . tauP                  1/2 tau_mnab W_mnij          (cc.cpp:346)
    //   LP -= P(ab) Z . t_old                 -1/2 P(ab) tau <am||ef> t    (cc.cpp:237-238 inside :348)
    DBuf lpz[3];
    LadderBlocks lb;
    lb.rowcls = sp.rowcls, lb.rowloc = sp.rowloc, lb.colcls = sp.colcls, lb.colloc = sp.colloc;
    for (int c = 0; c < sp.ncls; ++c) {
        const SymPack::Cls &k = sp.cls[c];
        lpz[c] = DBuf(ctx, (size_t) (std::max(k.nrow, 1) * k.ldl));
        lb.lp[c] = lpz[c].p, lb.ld[c] = k.ldl;
        if (!k.nrow || !k.ncol) {
            if (k.nrow) fill_zero(ctx, lpz[c].p, k.nrow * k.ldl);
            continue;
        }
        DBuf tauP(ctx, (size_t) (k.nrow * k.ldk)), wp(ctx, (size_t) (k.nrow * k.ldo));
        tau_pack(ctx, T1n, T2, o, v, k.rows, k.nrow, tauP, k.ldk, false, colmap(c));
        GemmArgs g;
        g.M = k.nrow, g.N = (int) k.ntot, g.K = k.ncol;
        g.A = tauP, g.lda = k.ldk;       // [M][K]
        g.B = k.ladB, g.ldb = k.ldk;     // [N][K]
        g.C = lpz[c], g.ldc = k.ldl, g.c_owns_rows = true;
        gemm(ctx, g, true, false);
        double *LP = lpz[c].p, *Z = lpz[c].p + k.offZ, *Y = lpz[c].p + k.offY;
        GemmArgs g2;
        g2.M = k.nrow, g2.N = k.ncol, g2.K = k.nrow, g2.alpha = 0.5, g2.beta = 1.0;
        g2.A = Y, g2.lda = k.ldl;            // [M][K]
        g2.B = k.tauoP, g2.ldb = k.ldk;      // [K][N]
        g2.C = LP, g2.ldc = k.ldl;
        gemm(ctx, g2, true, true);
        wmnij_pack(ctx, wmnij.p, o, k.rows, k.nrow, wp, k.ldo);
        g2.alpha = 1.0, g2.A = wp, g2.lda = k.ldo, g2.B = tauP;
        gemm(ctx, g2, true, true);
        PairColMap am;
        am.cls = sp.amcls, am.loc = sp.amloc, am.want = c;
        z_t_sub(ctx, Z, k.ldl, T1, o, v, k.nrow, LP, k.ldl, colmap(c), am);
    }
    // ring terms cc.cpp:354-363: Rt[i,a,b,j] = sum_me t_imae W_mbej + sum_m t_ma (sum_e t_ie <mb||je>), used at four relabelings
    DBuf rt(ctx, (size_t) o2v2), p1(ctx, (size_t) (O * O * V * O));
    const TView Rt(rt.p, {O, V, O, V}), P1(p1.p, {O, O, O, V});
    contract(ctx, 1.0, tn, "ie", ovov, "mbje", 0.0, P1, "imjb");
    contract(ctx, 1.0, tn, "ma", P1, "imjb", 0.0, Rt, "iajb");                                 // :360-363 (first writer)
    if (!ring_blocks) {
        contract(ctx, 1.0, Lf, "iame", Wm, "mejb", 1.0, Rt, "iajb");                           // :354-357
    } else {
        static const int rkc[3][3] = {{0, 0, 0}, {1, 2, 2}, {2, 1, 1}};// (class of (i,a), of (m,e), of (j,b))
        for (int q = 0; q < 3; ++q) {
            const int r = rkc[q][0], kk = rkc[q][1], c = rkc[q][2];
            if (!sp.ovn[r] || !sp.ovn[kk] || !sp.ovn[c]) continue;
            GemmArgs g;
            g.M = (int) sp.ovn[r], g.N = (int) sp.ovn[c], g.K = (int) sp.ovn[kk];
            g.A = lfS.p + sp.ov0[r] * ldr + sp.ov0[kk], g.lda = ldr;// [M][K]
            g.B = wS.p + sp.ov0[kk] * ldr + sp.ov0[c], g.ldb = ldr;  // [K][N]
            g.C = rtS.p + sp.ov0[r] * ldr + sp.ov0[c], g.ldc = ldr;
            gemm(ctx, g, true, true);
        }
    }
    finalize_t2(ctx, I.p[I_OOVV], uab, uij, rt, ring_blocks ? rtS.p : nullptr, sp.ovpos, ldr, lb, foo, fvv, nullptr, o, v,
                T2n);                                                                          // :312, P(..) combines, :374-383
}

}// namespace qcp2
