// post_hf.cuh -- device-resident implementations of the reference's hot-path functions
// (integrals.h, mp2.h, cc.h).  All pointers are device pointers on ctx.device.
#pragma once
#include "common.cuh"

namespace qcp2 {

enum IntBlock { I_OOOO, I_OVOO, I_OOVO, I_OOOV, I_OVOV, I_OVVO, I_OOVV, I_VOVV, I_OVVV, I_VVVO, I_VVVV, I_COUNT };
extern const char *const kIntNames[I_COUNT];
enum InterBlock { X_FAE, X_FMI, X_FME, X_WMNIJ, X_WMBEJ, X_WABEF, X_COUNT };

// b<c / i<j packed operands of the particle-particle ladder (valid for antisymmetric integrals only)
struct LadderPack {
    const double *vvvvP = nullptr;// [v(v-1)/2][v(v-1)/2]
    const int *pairs_o = nullptr; // device, 2 ints per i<j pair
    int npo = 0, npv = 0;
    long long ld = 0;            // row stride of the packed operands (npv rounded up to even)
    // >= 0: the caller already knows whether the doubles amplitudes (hence tau) are antisymmetric in (i,j) and
    // (a,b) (ccsd_dev reads the check back together with the iteration's energy); < 0: check tau here
    int tau_antisymmetric = -1;
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
                 const double *Ts_old = nullptr, const struct LadderPack *pack = nullptr);
void ccsd_dev(Ctx &ctx, const IntsDev &I, const double *foo, const double *fov, const double *fvv,
              const double *T2_guess, int o, int v, int maxiter, double conv, double *T1, double *T2, double *e_host,
              int *iters_host, double *e_hist_host);

}// namespace qcp2
