// post_hf.cuh -- device-resident implementations of the reference's hot-path functions
// (integrals.h, mp2.h, cc.h).  All pointers are device pointers on ctx.device.
#pragma once
#include "common.cuh"

namespace qcp2 {

enum IntBlock { I_OOOO, I_OVOO, I_OOVO, I_OOOV, I_OVOV, I_OVVO, I_OOVV, I_VOVV, I_OVVV, I_VVVO, I_VVVV, I_COUNT };
extern const char *const kIntNames[I_COUNT];
enum InterBlock { X_FAE, X_FMI, X_FME, X_WMNIJ, X_WMBEJ, X_WABEF, X_COUNT };

// Spatial MO integral tensors (aa, bb, ab as transform_to_so takes them; RHF: all three the same tensor): the source
// any spin-orbital block can be evaluated from without the (2n)^4 tensor (tensor_ops: so_block, so_block_packed)
struct SoSource {
    const double *aa = nullptr, *bb = nullptr, *ab = nullptr;
    int n = 0;
};

// Pair-packed constants of the symmetric CCSD iteration (ccsd_sym.cu); built once per solve by sym_setup().
// Pairs are grouped into CLASSES.  Without spin structure there is one class.  When the inputs carry the spin zero
// pattern of transform_to_so (integrals.cpp:143-161: <pq||rs> = 0 unless s_p + s_q = s_r + s_s; checked on the device,
// exact zeros) a pair (x,y) belongs to class s_x + s_y (0 = alpha-alpha, 1 = mixed, 2 = beta-beta) and every pair-packed
// matrix is block diagonal in the class: only the diagonal blocks are stored and multiplied.
struct SymPack {
    DBuf ints_buf, ovvo_mejb;
    int ncls = 1;
    bool spin_blocked = false;
    int npo = 0, npv = 0;
    struct Cls {
        int nrow = 0, ncol = 0, nam = 0, nall = 0;// i<j pairs, a<b pairs, (a,m) pairs, ordered (i,n) pairs of the class
        long long ldk = 0;                        // row stride of the [..][a<b] matrices of the class (even)
        long long ldo = 0, lda = 0;               // row strides of [..][i<j] and [..][(i,n)] matrices
        long long offZ = 0, offY = 0, ntot = 0, ldl = 0;// row segments of ladB = column segments of tauP . ladB^T
        const int *rows = nullptr, *cols = nullptr, *ams = nullptr, *alls = nullptr;// device pair lists (2 ints per pair)
        DBuf ladB, oovvA, tauoP;                  // [<ab||ef>; <am||ef>; <mn||ef>][e<f], <(m,n)||ef> all ordered pairs, tau(T1 old, T2)
    } cls[3];
    // spin-blocked only: (occupied, virtual) composites i*v + a in class-sorted order (class 0: same spin, 1: i alpha / a beta,
    // 2: i beta / a alpha; every class starts on an even position), and <mn||ef> as the matrix [pos(n,f)][pos(m,e)]
    const int *ovpos = nullptr;
    long long ov0[3] = {0, 0, 0}, ovn[3] = {0, 0, 0}, ldov = 0;
    DBuf oovvS;
    // class / position of every natural pair index (null when ncls == 1: natural order)
    const int *rowcls = nullptr, *rowloc = nullptr, *colcls = nullptr, *colloc = nullptr, *amcls = nullptr, *amloc = nullptr,
              *allcls = nullptr, *allloc = nullptr;
};

struct IntsDev {
    const double *p[I_COUNT];
};
struct InterDev {
    double *p[X_COUNT];
};
long long int_block_size(int k, int o, int v);
long long inter_block_size(int k, int o, int v);

void transform_ao_mo_dev(Ctx &ctx, const double *ao, const double *C1, const double *C2, int n, double *out);
void mp2_dev(Ctx &ctx, const double *ao, const double *Ca, const double *Cb, const double *ea, const double *eb, int n,
             int o, int v, bool uhf, double *so_ints /*device, may be null*/, double *T /*may be null*/,
             double *e_host);

void update_intermediates_dev(Ctx &ctx, const double *Ts, const double *Td, const IntsDev &I, const double *foo,
                              const double *fov, const double *fvv, int o, int v, const InterDev &X);
void make_T1_dev(Ctx &ctx, const double *Ts, const double *Td, const IntsDev &I, const InterDev &X, const double *foo,
                 const double *fov, const double *fvv, const double *D_ia, int o, int v, double *T1_new);
void make_T2_dev(Ctx &ctx, const double *Ts, const double *Td, const IntsDev &I, const InterDev &X, const double *foo,
                 const double *fvv, const double *D_ijab, int o, int v, double *T2_new,
                 const double *Ts_old = nullptr);
// ccsd_sym.cu: true if every permutational symmetry the packed formulation relies on holds (checked on the device)
// I.p[I_VVVV] may be null when `src` is given: <ab||ef> is then checked and pair-packed straight from the spatial tensors
bool sym_setup(Ctx &ctx, const IntsDev &I, const double *T2_guess, const double *fov, int o, int v, SymPack &sp,
               const SoSource *src = nullptr);
void ccsd_sym_iteration(Ctx &ctx, const IntsDev &I, const double *foo, const double *fov, const double *fvv, int o, int v,
                        const SymPack &sp, const double *T1, const double *T2, double *T1n, double *T2n);
void ccsd_dev(Ctx &ctx, const IntsDev &I, const double *foo, const double *fov, const double *fvv,
              const double *T2_guess, int o, int v, int maxiter, double conv, double *T1, double *T2, double *e_host,
              int *iters_host, double *e_hist_host, const SoSource *src = nullptr);
// the fused pipeline's front end (mp2.cpp:45-86 + cc.cpp:8-23 without the (2n)^4 tensor): AO->MO, then the ten o/v blocks
// other than <vv||vv> straight from the spatial tensors, MP2 amplitudes and energy.  `blocks` receives the buffers,
// `spatial` keeps the MO tensors alive for the CCSD setup (which packs or builds <vv||vv> from them).
void mp2_blocks_dev(Ctx &ctx, const double *ao, const double *Ca, const double *Cb, const double *ea, const double *eb, int n, int o,
                    int v, bool uhf, bool want_blocks, DBuf *blocks /*[I_COUNT]*/, IntsDev &I, DBuf spatial[3], SoSource &src,
                    double *T /*o^2 v^2*/, double *e_host);

}// namespace qcp2
