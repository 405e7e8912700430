// api.cu -- the C ABI of include/qcp2_b200.h: argument checking, host<->device marshalling
// for QCP2_HOST pointer mode, exception -> status translation.  No arithmetic happens here
// and there is no CPU path: every entry point ends in kernels on the context's device.
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/qcp2_b200.h"
#include "comm.cuh"
#include "contract.cuh"
#include "dgemm.cuh"
#include "post_hf.cuh"
#include "tensor_ops.cuh"
#include "triples.cuh"

using namespace qcp2;

struct qcp2_ctx {
    Ctx c;
    DBuf t1_start;// qcp2_ccsd_set_start_T1: consumed by the next CCSD solve
    Comm *comm = nullptr;// NCCL communicator (qcp2_comm_init); null = single GPU
    bool ccsd_sharding = false;// qcp2_set_ccsd_sharding: qcp2_ccsd is collective and splits its large products
};
struct qcp2_t_plan {
    qcp2_ctx *owner = nullptr;
    std::unique_ptr<TPlan> plan;
};

namespace {

// device view of an input tensor: pass-through in device mode, staged copy in host mode
struct In {
    DBuf buf;
    const double *p = nullptr;
    In() {}
    In(Ctx &c, const double *src, long long n) {
        if (!src) return;
        if (c.pointer_mode == QCP2_DEVICE) {
            p = src;
        } else {
            buf = DBuf(c, (size_t) std::max<long long>(n, 1));
            if (n > 0) QCP2_CUDA(cudaMemcpyAsync(buf.p, src, (size_t) n * 8, cudaMemcpyHostToDevice, c.stream));
            p = buf.p;
        }
    }
    operator const double *() const { return p; }
};
// device view of an output tensor; finish() copies back in host mode
struct Out {
    Ctx *c = nullptr;
    DBuf buf;
    double *p = nullptr, *host = nullptr;
    long long n = 0;
    Out() {}
    Out(Ctx &ctx, double *dst, long long n_) : c(&ctx), n(n_) {
        if (!dst) return;
        if (ctx.pointer_mode == QCP2_DEVICE) p = dst;
        else buf = DBuf(ctx, (size_t) std::max<long long>(n_, 1)), p = buf.p, host = dst;
    }
    void finish() {
        if (host && n > 0) QCP2_CUDA(cudaMemcpyAsync(host, p, (size_t) n * 8, cudaMemcpyDeviceToHost, c->stream));
    }
    operator double *() const { return p; }
};

template<class F>
int guarded(qcp2_ctx *ctx, F &&f) {
    if (!ctx) return 2;
    try {
        QCP2_CUDA(cudaSetDevice(ctx->c.device));
        f(ctx->c);
        return 0;
    } catch (const std::exception &e) {
        ctx->c.last_error = e.what();
        return 1;
    }
}

long long ipow4(long long n) { return n * n * n * n; }


}// namespace

extern "C" {

const char *qcp2_version(void) { return "qcp2_b200 0.1 (sm_100a, DMMA.8x8x4)"; }

int qcp2_device_count(void) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return ndev;
}

int qcp2_create(int device, qcp2_ctx **out) {
    if (!out) return 2;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return 3;// no CUDA device: there is deliberately no CPU fallback
    }
    if (device < 0 || device >= ndev) return 4;
    std::unique_ptr<qcp2_ctx> ctx(new qcp2_ctx());
    try {
        QCP2_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        QCP2_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10) throw Error("qcp2_b200 requires an sm_100a device (B200)");
        ctx->c.device = device;
        ctx->c.sm_count = prop.multiProcessorCount;
        QCP2_CUDA(cudaStreamCreateWithFlags(&ctx->c.own_stream, cudaStreamNonBlocking));
        ctx->c.stream = ctx->c.own_stream;
        // a private stream-ordered pool: freed workspace stays cached in it (release threshold = max) and goes
        // back to the driver in qcp2_destroy; the device's default pool is left untouched
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        QCP2_CUDA(cudaMemPoolCreate(&ctx->c.pool, &props));
        unsigned long long keep = ~0ULL;
        QCP2_CUDA(cudaMemPoolSetAttribute(ctx->c.pool, cudaMemPoolAttrReleaseThreshold, &keep));
        // split-K tile counters of the warp-specialised GEMM (zero between launches)
        ctx->c.ws_counters_cap = 1 << 16;
        QCP2_CUDA(cudaMalloc(&ctx->c.ws_counters, (size_t) ctx->c.ws_counters_cap * sizeof(unsigned int)));
        QCP2_CUDA(cudaMemset(ctx->c.ws_counters, 0, (size_t) ctx->c.ws_counters_cap * sizeof(unsigned int)));
    } catch (const std::exception &e) {
        fprintf(stderr, "qcp2_create: %s\n", e.what());
        return 1;
    }
    *out = ctx.release();
    return 0;
}

int qcp2_destroy(qcp2_ctx *ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->c.device);
    cudaStreamSynchronize(ctx->c.stream);
    if (ctx->comm) {
        try {
            comm_destroy(ctx->comm);
        } catch (...) {
        }
        ctx->comm = nullptr;
    }
    ctx->t1_start.reset();
    if (ctx->c.ws_counters) cudaFree(ctx->c.ws_counters);
    for (Ctx::OffsetTable &t : ctx->c.offset_tables) cudaFree(t.dev);
    if (ctx->c.pool) {
        cudaMemPoolTrimTo(ctx->c.pool, 0);
        cudaMemPoolDestroy(ctx->c.pool);
    }
    if (ctx->c.own_stream) cudaStreamDestroy(ctx->c.own_stream);
    delete ctx;
    return 0;
}

int qcp2_set_pointer_mode(qcp2_ctx *ctx, int mode) {
    if (!ctx || (mode != QCP2_HOST && mode != QCP2_DEVICE)) return 2;
    ctx->c.pointer_mode = mode;
    return 0;
}
int qcp2_set_stream(qcp2_ctx *ctx, void *stream) {
    if (!ctx) return 2;
    ctx->c.stream = stream ? static_cast<cudaStream_t>(stream) : ctx->c.own_stream;
    return 0;
}
int qcp2_synchronize(qcp2_ctx *ctx) {
    return guarded(ctx, [&](Ctx &c) { c.sync(); });
}
const char *qcp2_last_error(const qcp2_ctx *ctx) { return ctx ? ctx->c.last_error.c_str() : "null context"; }
long long qcp2_kernel_launches(const qcp2_ctx *ctx) { return ctx ? ctx->c.launches : 0; }

// ---- integrals.h ------------------------------------------------------------------
int qcp2_transform_ao_mo(qcp2_ctx *ctx, const double *pq_rs, const double *C1, const double *C2, int n, double *out) {
    return guarded(ctx, [&](Ctx &c) {
        QCP2_REQUIRE(pq_rs && C1 && C2 && out && n >= 1, "transform_ao_mo: null argument or n < 1");
        const long long n4 = ipow4(n);
        In ao(c, pq_rs, n4), c1(c, C1, (long long) n * n), c2(c, C2, (long long) n * n);
        Out o(c, out, n4);
        transform_ao_mo_dev(c, ao, c1, c2, n, o);
        o.finish();
        c.sync();
    });
}

int qcp2_transform_to_so(qcp2_ctx *ctx, const double *aa, const double *bb, const double *ab, int n, double *so) {
    return guarded(ctx, [&](Ctx &c) {
        QCP2_REQUIRE(aa && bb && ab && so && n >= 1, "transform_to_so: null argument or n < 1");
        const long long n4 = ipow4(n);
        In a(c, aa, n4);
        In b = (bb == aa) ? In() : In(c, bb, n4);
        In x = (ab == aa) ? In() : (ab == bb ? In() : In(c, ab, n4));
        const double *pb = (bb == aa) ? a.p : b.p;
        const double *px = (ab == aa) ? a.p : (ab == bb ? pb : x.p);
        Out o(c, so, 16 * n4);
        to_so(c, a, pb, px, n, o);
        o.finish();
        c.sync();
    });
}

int qcp2_slice_ints(qcp2_ctx *ctx, const double *so, int n2, int o, int v, const char *spec, double *out) {
    return guarded(ctx, [&](Ctx &c) {
        QCP2_REQUIRE(so && out && spec && strlen(spec) == 4, "slice_ints: bad argument");
        long long n = 1;
        for (int d = 0; d < 4; ++d) n *= spec[d] == 'o' ? o : v;
        In s(c, so, ipow4(n2));
        Out r(c, out, n);
        slice_block(c, s, n2, o, v, spec, r);
        r.finish();
        c.sync();
    });
}

int qcp2_make_denom(qcp2_ctx *ctx, const double *F_spin, int n2, int o, int v, double *denom) {
    return guarded(ctx, [&](Ctx &c) {
        QCP2_REQUIRE(F_spin && denom && o + v <= n2, "make_denom: bad argument");
        In f(c, F_spin, (long long) n2 * n2);
        Out d(c, denom, (long long) o * o * v * v);
        make_denom(c, f, n2, o, v, d);
        d.finish();
        c.sync();
    });
}

// ---- mp2.h ------------------------------------------------------------------------
int qcp2_make_so_moes(qcp2_ctx *ctx, const double *ea, const double *eb, int n, double *F) {
    return guarded(ctx, [&](Ctx &c) {
        QCP2_REQUIRE(ea && eb && F && n >= 1, "make_so_moes: bad argument");
        In a(c, ea, n), b(c, eb, n);
        Out f(c, F, 4LL * n * n);
        make_so_moes(c, a, b, n, f);
        f.finish();
        c.sync();
    });
}

int qcp2_run_mp2(qcp2_ctx *ctx, const double *oovv, const double *denom, int o, int v, double *T, double *energy) {
    return guarded(ctx, [&](Ctx &c) {
        QCP2_REQUIRE(oovv && denom && energy, "run_mp2: null argument");
        const long long n = (long long) o * o * v * v;
        In g(c, oovv, n), d(c, denom, n);
        Out t(c, T, n);
        DBuf e(c, 1);
        mp2_amplitudes_energy(c, g, d, n, t, e);
        t.finish();
        QCP2_CUDA(cudaMemcpyAsync(energy, e.p, 8, cudaMemcpyDeviceToHost, c.stream));
        c.sync();
    });
}

int qcp2_mp2(qcp2_ctx *ctx, const double *ao_ints, const double *Ca, const double *Cb, const double *eps_a,
             const double *eps_b, int n, int o, int v, int uhf, double *so_ints, double *T, double *energy) {
    return guarded(ctx, [&](Ctx &c) {
        QCP2_REQUIRE(ao_ints && Ca && eps_a && energy && n >= 1, "MP2: null argument");
        if (uhf) QCP2_REQUIRE(Cb && eps_b, "MP2: UHF needs Cb and eps_b");
        const long long n4 = ipow4(n);
        Trace tr(c, "qcp2_post_hf");
        In ao(c, ao_ints, n4), ca(c, Ca, (long long) n * n), ea(c, eps_a, n);
        In cb = uhf ? In(c, Cb, (long long) n * n) : In();
        In eb = uhf ? In(c, eps_b, n) : In();
        Out so(c, so_ints, 16 * n4), t(c, T, (long long) o * o * v * v);
        mp2_dev(c, ao, ca, uhf ? cb.p : ca.p, ea, uhf ? eb.p : ea.p, n, o, v, uhf != 0, so, t, energy);
        so.finish(), t.finish();
        c.sync();
    });
}

// ---- cc.h -------------------------------------------------------------------------
int qcp2_make_fock(qcp2_ctx *ctx, const double *e1, const double *e2, int n1, int n2, double *F) {
    return guarded(ctx, [&](Ctx &c) {
        QCP2_REQUIRE(e1 && e2 && F && n2 > n1 && n1 >= 0, "make_fock: bad argument");
        const long long need = n1 / 2 + (n2 - n1 - 1) / 2 + (n1 % 2) + 1;// highest eps index read, cc.cpp:28-31
        In a(c, e1, need), b(c, e2, need);
        Out f(c, F, (long long) (n2 - n1) * (n2 - n1));
        make_fock(c, a, b, n1, n2, f);
        f.finish();
        c.sync();
    });
}

int qcp2_make_tau(qcp2_ctx *ctx, const double *Ts, const double *Td, int o, int v, int bar, double *tau) {
    return guarded(ctx, [&](Ctx &c) {
        QCP2_REQUIRE(Ts && Td && tau, "make_tau: null argument");
        const long long n = (long long) o * o * v * v;
        In ts(c, Ts, (long long) o * v), td(c, Td, n);
        Out t(c, tau, n);
        make_tau(c, ts, td, o, v, bar != 0, t);
        t.finish();
        c.sync();
    });
}

int qcp2_multiply_Ts(qcp2_ctx *ctx, const double *Ts, int o, int v, double *out) {
    return guarded(ctx, [&](Ctx &c) {
        QCP2_REQUIRE(Ts && out, "multiply_Ts: null argument");
        In ts(c, Ts, (long long) o * v);
        Out t(c, out, (long long) o * v * o * v);
        multiply_Ts(c, ts, o, v, t);
        t.finish();
        c.sync();
    });
}

static int make_D(qcp2_ctx *ctx, const double *f_oo, const double *f_vv, int o, int v, double *D, int which) {
    return guarded(ctx, [&](Ctx &c) {
        QCP2_REQUIRE(f_oo && f_vv && D, "make_D: null argument");
        const long long O = o, V = v;
        const long long n = which == 1 ? O * V : (which == 2 ? O * O * V * V : O * O * O * V * V * V);
        In foo(c, f_oo, O * O), fvv(c, f_vv, V * V);
        Out d(c, D, n);
        if (which == 1) make_D_ia(c, foo, fvv, o, v, d);
        else if (which == 2) make_D_ijab(c, foo, fvv, o, v, d);
        else make_D_triples(c, foo, fvv, o, v, d);
        d.finish();
        c.sync();
    });
}
int qcp2_make_D_ia(qcp2_ctx *ctx, const double *f_oo, const double *f_vv, int o, int v, double *D) {
    return make_D(ctx, f_oo, f_vv, o, v, D, 1);
}
int qcp2_make_D_ijab(qcp2_ctx *ctx, const double *f_oo, const double *f_vv, int o, int v, double *D) {
    return make_D(ctx, f_oo, f_vv, o, v, D, 2);
}
int qcp2_make_D_triples(qcp2_ctx *ctx, const double *f_oo, const double *f_vv, int o, int v, double *D) {
    return make_D(ctx, f_oo, f_vv, o, v, D, 3);
}

namespace {
struct IntsIn {
    In in[I_COUNT];
    IntsDev dev;
    IntsIn(Ctx &c, const double *const *ints, int o, int v, unsigned need_mask) {
        for (int k = 0; k < I_COUNT; ++k) {
            dev.p[k] = nullptr;
            if (!ints[k]) {
                QCP2_REQUIRE(!(need_mask & (1u << k)), std::string("missing integral block ") + kIntNames[k]);
                continue;
            }
            if (!(need_mask & (1u << k)) && c.pointer_mode == QCP2_HOST) continue;// do not upload what is not read
            in[k] = In(c, ints[k], int_block_size(k, o, v));
            dev.p[k] = in[k].p;
        }
    }
};
struct InterIO {
    In in[X_COUNT];
    Out out[X_COUNT];
    InterDev dev;
};
constexpr unsigned bit(int k) { return 1u << k; }
constexpr unsigned kNeedInter = bit(I_OOOO) | bit(I_OOOV) | bit(I_OOVO) | bit(I_OOVV) | bit(I_VOVV) | bit(I_OVVV) |
                                bit(I_VVVV) | bit(I_OVVO);
constexpr unsigned kNeedT1 = bit(I_OVOV) | bit(I_OVVV) | bit(I_OOVO);
constexpr unsigned kNeedT2 = bit(I_OOVV) | bit(I_OVOV) | bit(I_VVVO) | bit(I_OVOO);
constexpr unsigned kNeedAll = (1u << I_COUNT) - 1;
}// namespace

int qcp2_update_intermediates(qcp2_ctx *ctx, const double *Ts, const double *Td, const double *const *ints,
                              const double *f_oo, const double *f_ov, const double *f_vv, int o, int v,
                              double *const *inter) {
    return guarded(ctx, [&](Ctx &c) {
        QCP2_REQUIRE(Ts && Td && ints && f_oo && f_vv && inter, "update_intermediates: null argument");
        const long long O = o, V = v;
        In ts(c, Ts, O * V), td(c, Td, O * O * V * V), foo(c, f_oo, O * O), fov(c, f_ov, O * V), fvv(c, f_vv, V * V);
        IntsIn I(c, ints, o, v, kNeedInter);
        Out outs[X_COUNT];
        InterDev X;
        for (int k = 0; k < X_COUNT; ++k) {
            QCP2_REQUIRE(inter[k], "update_intermediates: null output block");
            outs[k] = Out(c, inter[k], inter_block_size(k, o, v));
            X.p[k] = outs[k].p;
        }
        update_intermediates_dev(c, ts, td, I.dev, foo, f_ov ? fov.p : nullptr, fvv, o, v, X);
        for (int k = 0; k < X_COUNT; ++k) outs[k].finish();
        c.sync();
    });
}

int qcp2_make_T1(qcp2_ctx *ctx, const double *Ts, const double *Td, const double *const *ints,
                 const double *const *inter, const double *f_oo, const double *f_ov, const double *f_vv,
                 const double *D_ia, int o, int v, double *T1_new) {
    return guarded(ctx, [&](Ctx &c) {
        QCP2_REQUIRE(Ts && Td && ints && inter && T1_new, "make_T1: null argument");
        QCP2_REQUIRE(D_ia || (f_oo && f_vv), "make_T1: give D_ia or f_oo and f_vv");
        const long long O = o, V = v;
        In ts(c, Ts, O * V), td(c, Td, O * O * V * V), foo(c, f_oo, O * O), fov(c, f_ov, O * V), fvv(c, f_vv, V * V),
                dia(c, D_ia, O * V);
        IntsIn I(c, ints, o, v, kNeedT1);
        In xin[X_COUNT];
        InterDev X;
        for (int k = 0; k < X_COUNT; ++k) {
            X.p[k] = nullptr;
            if (k == X_FAE || k == X_FMI || k == X_FME) {
                QCP2_REQUIRE(inter[k], "make_T1: missing intermediate");
                xin[k] = In(c, inter[k], inter_block_size(k, o, v));
                X.p[k] = const_cast<double *>(xin[k].p);
            }
        }
        Out t(c, T1_new, O * V);
        make_T1_dev(c, ts, td, I.dev, X, foo, f_ov ? fov.p : nullptr, fvv, D_ia ? dia.p : nullptr, o, v, t);
        t.finish();
        c.sync();
    });
}

int qcp2_make_T2(qcp2_ctx *ctx, const double *Ts, const double *Td, const double *const *ints,
                 const double *const *inter, const double *f_oo, const double *f_vv, const double *D_ijab, int o, int v,
                 double *T2_new) {
    return guarded(ctx, [&](Ctx &c) {
        QCP2_REQUIRE(Ts && Td && ints && inter && T2_new, "make_T2: null argument");
        QCP2_REQUIRE(D_ijab || (f_oo && f_vv), "make_T2: give D_ijab or f_oo and f_vv");
        const long long O = o, V = v;
        In ts(c, Ts, O * V), td(c, Td, O * O * V * V), foo(c, f_oo, O * O), fvv(c, f_vv, V * V),
                dd(c, D_ijab, O * O * V * V);
        IntsIn I(c, ints, o, v, kNeedT2);
        In xin[X_COUNT];
        InterDev X;
        for (int k = 0; k < X_COUNT; ++k) {
            QCP2_REQUIRE(inter[k], "make_T2: missing intermediate");
            xin[k] = In(c, inter[k], inter_block_size(k, o, v));
            X.p[k] = const_cast<double *>(xin[k].p);
        }
        Out t(c, T2_new, O * O * V * V);
        make_T2_dev(c, ts, td, I.dev, X, foo, fvv, D_ijab ? dd.p : nullptr, o, v, t);
        t.finish();
        c.sync();
    });
}

int qcp2_ccsd_energy(qcp2_ctx *ctx, const double *Ts, const double *Td, const double *oovv, const double *f_ov, int o,
                     int v, double *energy) {
    return guarded(ctx, [&](Ctx &c) {
        QCP2_REQUIRE(Ts && Td && oovv && energy, "ccsd_energy: null argument");
        const long long O = o, V = v;
        In ts(c, Ts, O * V), td(c, Td, O * O * V * V), g(c, oovv, O * O * V * V), fov(c, f_ov, O * V);
        DBuf e(c, 1);
        ccsd_energy(c, ts, td, g, f_ov ? fov.p : nullptr, o, v, e);
        QCP2_CUDA(cudaMemcpyAsync(energy, e.p, 8, cudaMemcpyDeviceToHost, c.stream));
        c.sync();
    });
}

int qcp2_ccsd(qcp2_ctx *ctx, const double *so_ints, int n2, const double *const *ints, const double *f_oo,
              const double *f_ov, const double *f_vv, const double *T2_guess, int o, int v, int maxiter, double conv,
              double *T1, double *T2, double *energy, int *iters, double *e_hist, double *const *sliced_out) {
    return guarded(ctx, [&](Ctx &c) {
        QCP2_REQUIRE((so_ints || ints) && f_oo && f_vv && T2_guess && T1 && T2 && energy, "CCSD: null argument");
        QCP2_REQUIRE(maxiter >= 0, "CCSD: negative maxiter");
        const long long O = o, V = v;
        In foo(c, f_oo, O * O), fov(c, f_ov, O * V), fvv(c, f_vv, V * V), guess(c, T2_guess, O * O * V * V);
        IntsDev I;
        DBuf own[I_COUNT];
        std::unique_ptr<IntsIn> given;
        if (so_ints) {// get_integrals, cc.cpp:8-23
            QCP2_REQUIRE(o + v <= n2, "CCSD: o + v exceeds n2");
            In so(c, so_ints, ipow4(n2));
            for (int k = 0; k < I_COUNT; ++k) {
                own[k] = DBuf(c, (size_t) std::max<long long>(int_block_size(k, o, v), 1));
                slice_block(c, so, n2, o, v, kIntNames[k], own[k]);
                I.p[k] = own[k].p;
            }
            c.sync();// `so` staging buffer dies here
        } else {
            given.reset(new IntsIn(c, ints, o, v, kNeedAll));
            I = given->dev;
        }
        Out t1(c, T1, O * V), t2(c, T2, O * O * V * V);
        struct ShardGuard {
            Ctx &c;
            ShardGuard(Ctx &ctx_, Comm *s) : c(ctx_) { c.shard = s; }
            ~ShardGuard() { c.shard = nullptr; }
        } guard(c, (ctx->ccsd_sharding && ctx->comm && ctx->comm->nranks > 1) ? ctx->comm : nullptr);
        ccsd_dev(c, I, foo, f_ov ? fov.p : nullptr, fvv, guess, o, v, maxiter, conv, t1, t2, energy, iters, e_hist);
        t1.finish(), t2.finish();
        if (sliced_out)
            for (int k = 0; k < I_COUNT; ++k)
                if (sliced_out[k]) {
                    const long long n = int_block_size(k, o, v);
                    QCP2_CUDA(cudaMemcpyAsync(sliced_out[k], I.p[k], (size_t) n * 8,
                                              c.pointer_mode == QCP2_DEVICE ? cudaMemcpyDeviceToDevice
                                                                            : cudaMemcpyDeviceToHost,
                                              c.stream));
                }
        c.sync();
    });
}

double qcp2_ccsd_last_loop_seconds(const qcp2_ctx *ctx) { return ctx ? ctx->c.ccsd_loop_seconds : 0.0; }

// ---- (T) --------------------------------------------------------------------------
long long qcp2_ccsd_t_count(int o, int mode) { return t_count(o, mode); }

int qcp2_ccsd_t_select_mode(qcp2_ctx *ctx, const double *T2, const double *oovv, const double *vovv,
                            const double *ovoo, int o, int v, int *mode) {
    return guarded(ctx, [&](Ctx &c) {
        QCP2_REQUIRE(T2 && oovv && vovv && ovoo && mode, "(T): null argument");
        const long long O = o, V = v;
        In t2(c, T2, O * O * V * V), g(c, oovv, O * O * V * V), w(c, vovv, V * O * V * V), x(c, ovoo, O * V * O * O);
        *mode = t_select_mode(c, t2, g, w, x, o, v);
    });
}

// Shared body of qcp2_ccsd_t_range / qcp2_ccsd_t_mgpu once the small tensors are on the device.
// vovv_dev: device <vo||vv> (classic path) or null; vovv_host + streamed: the root's host tensor is
// streamed slab by slab (SYMMETRIC only).  Returns the per-triple vector on the device in `et`.
namespace {
struct TEval {
    std::unique_ptr<TPlan> plan;
    DBuf et;
    long long nt = 0;
    double partial = 0.0;
};
void t_evaluate(Ctx &c, TEval &ev_, const double *t1, const double *t2, const double *g, const double *vovv_dev, const double *x,
                const double *eo, const double *ev, int o, int v, int m, const TStream *ts, long long first, long long stride,
                long long max_count, bool want_vector, Trace &tr) {
    ev_.plan.reset(t_plan_create(c, t1, t2, g, vovv_dev, x, eo, ev, o, v, m, ts));
    tr.mark(ts ? "plan (pack b<c, workspaces, slab stream enqueued)" : "plan (pack b<c, workspaces)");
    ev_.nt = ev_.plan->ntriples;
    if (want_vector) {
        ev_.et = DBuf(c, (size_t) std::max<long long>(ev_.nt, 1));
        fill_zero(c, ev_.et, ev_.nt);
    }
    t_plan_run(*ev_.plan, first, stride, max_count, want_vector ? ev_.et.p : nullptr, &ev_.partial);
    tr.mark("run triples");
}
bool stream_allowed(const Ctx &c, int mode, int o, int v) {
    return c.pointer_mode == QCP2_HOST && mode != QCP2_T_LITERAL && v >= 2 && t_count(o, QCP2_T_SYMMETRIC) > 0 &&
           !getenv("QCP2_T_NO_STREAM");
}
}// namespace

int qcp2_ccsd_t_range(qcp2_ctx *ctx, const double *T1, const double *T2, const double *oovv, const double *vovv,
                      const double *ovoo, const double *eps_o, const double *eps_v, int o, int v, int mode,
                      long long first, long long stride, long long max_count, double *e_triples, double *energy) {
    return guarded(ctx, [&](Ctx &c) {
        QCP2_REQUIRE(T1 && T2 && oovv && vovv && ovoo && eps_o && eps_v && energy, "(T): null argument");
        const long long O = o, V = v;
        Trace tr(c, "qcp2_ccsd_t_range");
        In t1(c, T1, O * V), t2(c, T2, O * O * V * V), g(c, oovv, O * O * V * V), x(c, ovoo, O * V * O * O), eo(c, eps_o, O),
                ev(c, eps_v, V);
        tr.mark("stage the small inputs");
        TEval E;
        bool done = false;
        int m = mode;
        // host pointer mode, symmetric inputs: <vo||vv> is streamed -- H2D, b<c packing and the evaluation of the
        // triples overlap, and the unpacked tensor never exists on the device
        if (stream_allowed(c, mode, o, v) && (mode == QCP2_T_SYMMETRIC || t_select_mode(c, t2, g, nullptr, x, o, v) == QCP2_T_SYMMETRIC)) {
            TStream ts;
            ts.vovv_host = vovv;
            t_evaluate(c, E, t1, t2, g, nullptr, x, eo, ev, o, v, QCP2_T_SYMMETRIC, &ts, first, stride, max_count, e_triples != nullptr, tr);
            done = true;
            if (mode == QCP2_T_AUTO) {// AUTO promised a check of <vo||vv> too: it ran slab by slab
                double defect, mx;
                t_plan_vovv_defect(*E.plan, &defect, &mx);
                if (!(defect <= 1e-10 * mx)) done = false, m = QCP2_T_LITERAL, E = TEval();
            }
        }
        In w;
        if (!done) {
            w = In(c, vovv, V * O * V * V);
            tr.mark("stage <vo||vv> (H2D in host pointer mode)");
            if (m == QCP2_T_AUTO) m = t_select_mode(c, t2, g, w, x, o, v);
            tr.mark("select mode");
            t_evaluate(c, E, t1, t2, g, w, x, eo, ev, o, v, m, nullptr, first, stride, max_count, e_triples != nullptr, tr);
        }
        *energy = E.partial;
        const long long nt = E.nt;
        if (e_triples && nt > 0) {
            // only the evaluated entries are written back; the rest of e_triples is left untouched
            std::vector<double> h((size_t) nt);
            QCP2_CUDA(cudaMemcpyAsync(h.data(), E.et.p, (size_t) nt * 8, cudaMemcpyDeviceToHost, c.stream));
            c.sync();
            long long cnt = first < nt ? (nt - first + stride - 1) / stride : 0;
            if (max_count >= 0) cnt = std::min(cnt, max_count);
            for (long long q = 0; q < cnt; ++q) e_triples[first + q * stride] = h[(size_t) (first + q * stride)];
        }
        c.sync();
    });
}

int qcp2_ccsd_t(qcp2_ctx *ctx, const double *T1, const double *T2, const double *oovv, const double *vovv,
                const double *ovoo, const double *eps_o, const double *eps_v, int o, int v, int mode,
                double *energy) {
    return qcp2_ccsd_t_range(ctx, T1, T2, oovv, vovv, ovoo, eps_o, eps_v, o, v, mode, 0, 1, -1, nullptr, energy);
}

/* ---- multi-GPU (one rank per GPU) --------------------------------------------------- */
int qcp2_comm_get_unique_id(void *id, int id_bytes) {
    if (!id || id_bytes < QCP2_COMM_ID_BYTES) return 2;
    try {
        comm_unique_id(id);
        return 0;
    } catch (const std::exception &e) {
        fprintf(stderr, "qcp2_comm_get_unique_id: %s\n", e.what());
        return 1;
    }
}

int qcp2_comm_init(qcp2_ctx *ctx, const void *id, int nranks, int rank) {
    return guarded(ctx, [&](Ctx &c) {
        QCP2_REQUIRE(!ctx->comm, "comm_init: the context already has a communicator");
        ctx->comm = comm_init(c, id, nranks, rank);
    });
}

int qcp2_comm_destroy(qcp2_ctx *ctx) {
    return guarded(ctx, [&](Ctx &c) {
        c.sync();
        comm_destroy(ctx->comm);
        ctx->comm = nullptr;
    });
}

int qcp2_comm_rank(const qcp2_ctx *ctx, int *rank, int *nranks) {
    if (!ctx) return 2;
    if (rank) *rank = ctx->comm ? ctx->comm->rank : 0;
    if (nranks) *nranks = ctx->comm ? ctx->comm->nranks : 1;
    return 0;
}

int qcp2_comm_broadcast(qcp2_ctx *ctx, double *dev_buf, long long n, int root) {
    return guarded(ctx, [&](Ctx &c) {
        QCP2_REQUIRE(dev_buf || n == 0, "comm_broadcast: null buffer");
        if (ctx->comm) comm_broadcast(c, *ctx->comm, dev_buf, n, root);
    });
}

int qcp2_comm_allreduce_sum(qcp2_ctx *ctx, double *dev_buf, long long n) {
    return guarded(ctx, [&](Ctx &c) {
        QCP2_REQUIRE(dev_buf || n == 0, "comm_allreduce_sum: null buffer");
        if (ctx->comm) comm_allreduce_sum(c, *ctx->comm, dev_buf, n);
    });
}

int qcp2_comm_allgather(qcp2_ctx *ctx, const double *dev_send, double *dev_recv, long long n_per_rank) {
    return guarded(ctx, [&](Ctx &c) {
        QCP2_REQUIRE((dev_send && dev_recv) || n_per_rank == 0, "comm_allgather: null buffer");
        Comm single;
        comm_allgather(c, ctx->comm ? *ctx->comm : single, dev_send, dev_recv, n_per_rank);
    });
}

int qcp2_ccsd_t_mgpu(qcp2_ctx *ctx, const double *T1, const double *T2, const double *oovv, const double *vovv,
                     const double *ovoo, const double *eps_o, const double *eps_v, int o, int v, int mode, int root,
                     double *e_triples, double *energy) {
    return guarded(ctx, [&](Ctx &c) {
        Comm single;
        Comm &cm = ctx->comm ? *ctx->comm : single;
        QCP2_REQUIRE(energy && o >= 0 && v >= 0, "(T) mgpu: null argument");
        QCP2_REQUIRE(root >= 0 && root < cm.nranks, "(T) mgpu: root out of range");
        const bool is_root = cm.rank == root;
        Trace tr(c, "qcp2_ccsd_t_mgpu");
        const long long O = o, V = v;
        enum { kT1, kT2, kG, kW, kX, kEo, kEv };
        const long long sizes[7] = {O * V, O * O * V * V, O * O * V * V, V * O * V * V, O * V * O * O, O, V};
        const double *src[7] = {T1, T2, oovv, vovv, ovoo, eps_o, eps_v};
        // device images: the root stages (HOST mode) or uses in place (DEVICE mode) its tensors; the other
        // ranks receive into caller-provided device buffers (DEVICE mode) or scratch.  <vo||vv> is decided below.
        DBuf scratch[7];
        In staged[7];
        double *dev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        auto place = [&](int k) {
            if (is_root) {
                staged[k] = In(c, src[k], sizes[k]);
                dev[k] = const_cast<double *>(staged[k].p);
            } else if (c.pointer_mode == QCP2_DEVICE && src[k]) {
                dev[k] = const_cast<double *>(src[k]);
            } else {
                scratch[k] = DBuf(c, (size_t) std::max<long long>(sizes[k], 1));
                dev[k] = scratch[k].p;
            }
        };
        // header from the root: [0] evaluation mode, [1] 1 = <vo||vv> is streamed from the root's host tensor,
        // [2] the mode the root was CALLED with (the other ranks' `mode` argument is ignored, so a mismatch cannot
        // send the ranks down different collective sequences).  The root's pointer mode decides streaming.
        // Everything that can fail before the first payload collective (argument checks, staging, the symmetry
        // probe) runs inside this try block on every rank; the outcomes are then summed over the ranks so that all
        // of them leave together instead of one rank throwing while the others block in ncclBroadcast.
        DBuf hdr(c, 4);
        double hh[4] = {0.0, 0.0, 0.0, 0.0};
        std::string local_err;
        // Replicated host inputs: a rank other than the root may pass the tensors as well (HOST mode; they must hold the root's
        // values, e.g. mappings of one shared-memory copy).  When EVERY rank does, each uploads the small tensors itself and the
        // <vo||vv> slabs are dealt over the ranks (slab i by rank i % W over its own PCIe link), instead of everything
        // crossing the root's link and NVLink.
        const bool have_all = T1 && T2 && oovv && vovv && ovoo && eps_o && eps_v && c.pointer_mode == QCP2_HOST &&
                              !getenv("QCP2_T_ROOT_UPLOAD");
        double vote = have_all ? 1.0 : 0.0;
        if (is_root && !(T1 && T2 && oovv && vovv && ovoo && eps_o && eps_v)) local_err = "(T) mgpu: the root rank must supply every tensor";
        comm_agree(c, cm, local_err, "(T) mgpu", &vote);
        const bool replicated = cm.nranks > 1 && vote == (double) cm.nranks;
        try {
            if (replicated) {
                // every rank holds the host tensors: rank r uploads the r-th of W equal chunks of T2, <oo||vv> and <ov||oo> over its
                // own PCIe link and ONE in-place ncclAllGather per tensor assembles them everywhere (NVLink); T1 and the orbital
                // energies (kilobytes) are uploaded by every rank
                const int W = cm.nranks;
                for (int k = 0; k < 7; ++k) {
                    if (k == kW) continue;
                    if (k == kT1 || k == kEo || k == kEv) {
                        staged[k] = In(c, src[k], sizes[k]);
                        dev[k] = const_cast<double *>(staged[k].p);
                        continue;
                    }
                    const long long per = (sizes[k] + W - 1) / W, lo = std::min(sizes[k], per * cm.rank), hi = std::min(sizes[k], lo + per);
                    scratch[k] = DBuf(c, (size_t) std::max<long long>(per * W, 1));
                    dev[k] = scratch[k].p;
                    if (hi > lo)
                        QCP2_CUDA(cudaMemcpyAsync(dev[k] + lo, src[k] + lo, (size_t) (hi - lo) * 8, cudaMemcpyHostToDevice, c.stream));
                }
            } else {
                for (int k = 0; k < 7; ++k)
                    if (k != kW) place(k);
            }
            c.sync();
        } catch (const std::exception &e) {
            local_err = e.what();
            cudaGetLastError();
        }
        comm_agree(c, cm, local_err, "(T) mgpu, staging");
        if (replicated) {
            const int W = cm.nranks;
            for (int k : {(int) kT2, (int) kG, (int) kX}) {
                const long long per = (sizes[k] + W - 1) / W;
                comm_allgather(c, cm, dev[k] + per * cm.rank, dev[k], per);
            }
        }
        try {
            if (is_root) {
                if (!replicated) c.sync();
                int m = mode;
                bool streamed = false;
                if (stream_allowed(c, mode, o, v) &&
                    (mode == QCP2_T_SYMMETRIC || t_select_mode(c, dev[kT2], dev[kG], nullptr, dev[kX], o, v) == QCP2_T_SYMMETRIC))
                    streamed = true, m = QCP2_T_SYMMETRIC;
                if (!streamed) {
                    place(kW);
                    if (m == QCP2_T_AUTO) m = t_select_mode(c, dev[kT2], dev[kG], dev[kW], dev[kX], o, v);
                }
                QCP2_REQUIRE(m == QCP2_T_LITERAL || m == QCP2_T_SYMMETRIC, "(T) mgpu: unknown evaluation mode");
                hh[0] = (double) m, hh[1] = streamed ? 1.0 : 0.0, hh[2] = (double) mode;
            }
            c.sync();
        } catch (const std::exception &e) {
            local_err = e.what();
            cudaGetLastError();
        }
        comm_agree(c, cm, local_err, "(T) mgpu, mode selection");
        if (is_root) QCP2_CUDA(cudaMemcpyAsync(hdr.p, hh, 32, cudaMemcpyHostToDevice, c.stream));
        comm_broadcast(c, cm, hdr.p, 4, root);
        if (!replicated)
            for (int k = 0; k < 7; ++k)
                if (k != kW) comm_broadcast(c, cm, dev[k], sizes[k], root);
        QCP2_CUDA(cudaMemcpyAsync(hh, hdr.p, 32, cudaMemcpyDeviceToHost, c.stream));
        c.sync();
        tr.mark("stage + broadcast the small inputs");
        int m = (int) hh[0];
        bool streamed = hh[1] != 0.0;
        const int root_mode = (int) hh[2];// what the root was asked for; identical on every rank
        QCP2_REQUIRE(m == QCP2_T_LITERAL || m == QCP2_T_SYMMETRIC, "(T) mgpu: mode broadcast failed");
        TEval E;
        bool done = false;
        if (streamed) {
            TStream ts;
            ts.vovv_host = (is_root || replicated) ? vovv : nullptr, ts.comm = &cm, ts.root = root, ts.dealt = replicated;
            t_evaluate(c, E, dev[kT1], dev[kT2], dev[kG], nullptr, dev[kX], dev[kEo], dev[kEv], o, v, QCP2_T_SYMMETRIC, &ts, cm.rank,
                       cm.nranks, -1, true, tr);
            done = true;
            if (root_mode == QCP2_T_AUTO) {// the slab-wise <vo||vv> check, known to every rank through the broadcast
                double defect, mx;
                t_plan_vovv_defect(*E.plan, &defect, &mx);
                if (!(defect <= 1e-10 * mx)) done = false, m = QCP2_T_LITERAL, E = TEval();
            }
        }
        if (!done) {
            std::string err;
            try {
                if (!dev[kW]) place(kW);
                c.sync();
            } catch (const std::exception &e) {
                err = e.what();
                cudaGetLastError();
            }
            comm_agree(c, cm, err, "(T) mgpu, <vo||vv> staging");
            comm_broadcast(c, cm, dev[kW], sizes[kW], root);
            tr.mark("stage + broadcast <vo||vv>");
            t_evaluate(c, E, dev[kT1], dev[kT2], dev[kG], dev[kW], dev[kX], dev[kEo], dev[kEv], o, v, m, nullptr, cm.rank, cm.nranks, -1,
                       true, tr);// static round-robin: rank r owns r, r+W, ...
        }
        const long long nt = E.nt;
        // every slot is written by exactly one rank, so the sum is exact (x + 0 + ... + 0) and the
        // fixed-order host sum below is bit-identical for any number of ranks
        comm_allreduce_sum(c, cm, E.et.p, nt);
        std::vector<double> h((size_t) std::max<long long>(nt, 1), 0.0);
        if (nt > 0) QCP2_CUDA(cudaMemcpyAsync(h.data(), E.et.p, (size_t) nt * 8, cudaMemcpyDeviceToHost, c.stream));
        c.sync();
        tr.mark("all-reduce + read back");
        double total = 0.0;
        for (long long t = 0; t < nt; ++t) total += h[(size_t) t];
        if (e_triples)
            for (long long t = 0; t < nt; ++t) e_triples[t] = h[(size_t) t];
        *energy = total;
    });
}

int qcp2_t_plan_create(qcp2_ctx *ctx, const double *T1, const double *T2, const double *oovv, const double *vovv,
                       const double *ovoo, const double *eps_o, const double *eps_v, int o, int v, int mode,
                       qcp2_t_plan **plan) {
    if (plan) *plan = nullptr;
    return guarded(ctx, [&](Ctx &c) {
        QCP2_REQUIRE(plan && T1 && T2 && oovv && vovv && ovoo && eps_o && eps_v, "(T) plan: null argument");
        QCP2_REQUIRE(c.pointer_mode == QCP2_DEVICE, "(T) plan: needs QCP2_DEVICE pointer mode (resident inputs)");
        int m = mode;
        if (m == QCP2_T_AUTO) m = t_select_mode(c, T2, oovv, vovv, ovoo, o, v);
        std::unique_ptr<qcp2_t_plan> p(new qcp2_t_plan());
        p->owner = ctx;
        p->plan.reset(t_plan_create(c, T1, T2, oovv, vovv, ovoo, eps_o, eps_v, o, v, m));
        *plan = p.release();
    });
}

int qcp2_t_plan_run(qcp2_t_plan *plan, long long first, long long stride, long long max_count,
                    double *e_triples_dev, double *energy) {
    if (!plan) return 2;
    return guarded(plan->owner, [&](Ctx &) {
        QCP2_REQUIRE(energy, "(T) plan: null energy pointer");
        t_plan_run(*plan->plan, first, stride, max_count, e_triples_dev, energy);
    });
}

int qcp2_t_plan_timings(const qcp2_t_plan *plan, double *gemm_seconds, double *total_seconds,
                        long long *gemm_launches, long long *all_launches) {
    if (!plan) return 2;
    if (gemm_seconds) *gemm_seconds = plan->plan->last_gemm_seconds;
    if (total_seconds) *total_seconds = plan->plan->last_total_seconds;
    if (gemm_launches) *gemm_launches = plan->plan->last_gemm_launches;
    if (all_launches) *all_launches = plan->plan->last_all_launches;
    return 0;
}

int qcp2_t_plan_info(const qcp2_t_plan *plan, int *uses_tma_kernel, long long *ncols, long long *row_stride, int *triples_per_batch) {
    if (!plan) return 2;
    if (uses_tma_kernel) *uses_tma_kernel = plan->plan->use_tma ? 1 : 0;
    if (ncols) *ncols = plan->plan->ncols;
    if (row_stride) *row_stride = plan->plan->ld;
    if (triples_per_batch) *triples_per_batch = plan->plan->tb;
    return 0;
}

int qcp2_t_plan_destroy(qcp2_t_plan *plan) {
    if (!plan) return 0;
    cudaSetDevice(plan->owner->c.device);
    delete plan;
    return 0;
}

namespace {
// Body of qcp2_post_hf / qcp2_post_hf_mgpu.  cm == nullptr: single GPU.  With a communicator the call is
// collective: the root's boundary inputs are broadcast, MP2 / slicing run replicated (milliseconds), the large
// products of every CCSD iteration are row-sharded (dgemm.cu) and the (T) triples dealt round-robin.
void post_hf_impl(Ctx &c, Comm *cm, int root, const double *ao_ints, const double *Ca, const double *Cb, const double *eps_a,
                  const double *eps_b, int n, int o, int v, int uhf, int stages, int maxiter, double conv, double *e_mp2,
                  double *e_ccsd, double *e_t, int *iters, double *e_hist, double *T1, double *T2) {
    const bool multi = cm && cm->nranks > 1;
    const bool is_root = !multi || cm->rank == root;
    QCP2_REQUIRE(e_mp2 && n >= 1, "post_hf: null argument");
    QCP2_REQUIRE(stages >= 1 && stages <= 3, "post_hf: stages must be 1, 2 or 3");
    if (stages >= 2) QCP2_REQUIRE(e_ccsd && maxiter >= 0, "post_hf: CCSD outputs missing");
    if (stages >= 3) QCP2_REQUIRE(e_t, "post_hf: (T) output missing");
    const long long n4 = ipow4(n), O = o, V = v, o2v2 = O * O * V * V;
    const int n2 = 2 * n;
    Trace tr(c, multi ? "qcp2_post_hf_mgpu" : "qcp2_post_hf");
    // boundary inputs on the device: staged / used in place on the root, received into scratch elsewhere
    DBuf scratch[5];
    auto place = [&](const double *src, long long count, int slot) -> In {
        if (is_root) return In(c, src, count);
        scratch[slot] = DBuf(c, (size_t) std::max<long long>(count, 1));
        In in;
        in.p = scratch[slot].p;
        return in;
    };
    In ao, ca, ea, cb, eb;
    {// root-only checks and the staging can fail on one rank only: agree on the outcome before the first broadcast
        std::string err;
        try {
            if (is_root) QCP2_REQUIRE(ao_ints && Ca && eps_a, "post_hf: null argument");
            if (uhf && is_root) QCP2_REQUIRE(Cb && eps_b, "post_hf: UHF needs Cb and eps_b");
            ao = place(ao_ints, n4, 0), ca = place(Ca, (long long) n * n, 1), ea = place(eps_a, n, 2);
            if (uhf) cb = place(Cb, (long long) n * n, 3), eb = place(eps_b, n, 4);
            if (multi) c.sync();
        } catch (const std::exception &e) {
            err = e.what();
            cudaGetLastError();
        }
        if (multi) comm_agree(c, *cm, err, "post_hf_mgpu");
        else if (!err.empty()) throw Error(err);
    }
    if (multi) {
        comm_broadcast(c, *cm, const_cast<double *>(ao.p), n4, root);
        comm_broadcast(c, *cm, const_cast<double *>(ca.p), (long long) n * n, root);
        comm_broadcast(c, *cm, const_cast<double *>(ea.p), n, root);
        if (uhf) comm_broadcast(c, *cm, const_cast<double *>(cb.p), (long long) n * n, root), comm_broadcast(c, *cm, const_cast<double *>(eb.p), n, root);
    }
    const double *pea = ea.p, *peb = uhf ? eb.p : ea.p;
    DBuf t2_mp2(c, (size_t) std::max<long long>(o2v2, 1));
    IntsDev I;
    DBuf blocks[I_COUNT], spatial[3];
    SoSource src;
    QCP2_CUDA(cudaMemPoolSetAttribute(c.pool, cudaMemPoolAttrUsedMemHigh, &c.pool_reset_value));// peak-memory report of this call
    tr.mark("stage AO integrals");
    // mp2.cpp:45-86 and get_integrals (cc.cpp:8-23) without the (2n)^4 tensor and without <vv||vv>: every block is gathered
    // straight from the spatial MO tensors, which stay alive for the CCSD setup (it packs or builds <vv||vv> from them)
    mp2_blocks_dev(c, ao, ca, uhf ? cb.p : ca.p, pea, peb, n, o, v, uhf != 0, stages >= 2, blocks, I, spatial, src, t2_mp2, e_mp2);
    tr.mark("AO->MO, the o/v blocks, MP2");
    (void) n2;
    if (stages < 2) return;
    DBuf foo(c, (size_t) (O * O)), fvv(c, (size_t) (V * V)), F(c, (size_t) ((long long) n2 * n2));
    make_so_moes(c, pea, peb, n, F);// same diagonal as make_fock(eps_a, eps_b, 0, no+nv), cc.cpp:25-34
    {
        const long long oo[2] = {O, O}, so2[2] = {n2, 1};
        permute_axpby(c, foo, F.p, 2, oo, so2, 1.0, 0.0);
        const long long vv[2] = {V, V};
        permute_axpby(c, fvv, F.p + O * n2 + O, 2, vv, so2, 1.0, 0.0);
    }
    Out t1(c, T1, O * V), t2(c, T2, o2v2);
    DBuf t1own, t2own;
    double *t1p = t1.p, *t2p = t2.p;
    if (!t1p) t1own = DBuf(c, (size_t) std::max<long long>(O * V, 1)), t1p = t1own.p;
    if (!t2p) t2own = DBuf(c, (size_t) std::max<long long>(o2v2, 1)), t2p = t2own.p;
    {
        struct ShardGuard {// row-sharded products for the duration of the iteration loop only
            Ctx &c;
            ShardGuard(Ctx &ctx_, Comm *s) : c(ctx_) { c.shard = s; }
            ~ShardGuard() { c.shard = nullptr; }
        } guard(c, multi && !getenv("QCP2_CCSD_NO_SHARD") ? cm : nullptr);
        ccsd_dev(c, I, foo, nullptr, fvv, t2_mp2, o, v, maxiter, conv, t1p, t2p, e_ccsd, iters, e_hist, &src);// cc.cpp:409-451
    }
    for (DBuf &sb : spatial) sb.reset();
    t1.finish(), t2.finish();
    tr.mark("CCSD iterations");
    if (stages >= 3) {// cc.cpp:454-560
        DBuf eo(c, (size_t) std::max(o, 1)), ev(c, (size_t) std::max(v, 1));
        extract_diag(c, foo, o, eo);
        extract_diag(c, fvv, v, ev);
        const int mode = t_select_mode(c, t2p, I.p[I_OOVV], I.p[I_VOVV], I.p[I_OVOO], o, v);
        std::unique_ptr<TPlan> plan(t_plan_create(c, t1p, t2p, I.p[I_OOVV], I.p[I_VOVV], I.p[I_OVOO], eo, ev, o, v, mode));
        if (!multi) {
            t_plan_run(*plan, 0, 1, -1, nullptr, e_t);
        } else {// the amplitudes and blocks are already replicated: round-robin triples, one all-reduce, fixed-order sum
            const long long nt = plan->ntriples;
            DBuf et(c, (size_t) std::max<long long>(nt, 1));
            fill_zero(c, et, nt);
            double partial = 0.0;
            t_plan_run(*plan, cm->rank, cm->nranks, -1, et.p, &partial);
            comm_allreduce_sum(c, *cm, et.p, nt);
            std::vector<double> h((size_t) std::max<long long>(nt, 1), 0.0);
            if (nt > 0) QCP2_CUDA(cudaMemcpyAsync(h.data(), et.p, (size_t) nt * 8, cudaMemcpyDeviceToHost, c.stream));
            c.sync();
            double total = 0.0;
            for (long long t = 0; t < nt; ++t) total += h[(size_t) t];
            *e_t = total;
        }
        tr.mark("(T)");
    }
    c.sync();
}
}// namespace

int qcp2_post_hf(qcp2_ctx *ctx, const double *ao_ints, const double *Ca, const double *Cb, const double *eps_a,
                 const double *eps_b, int n, int o, int v, int uhf, int stages, int maxiter, double conv, double *e_mp2,
                 double *e_ccsd, double *e_t, int *iters, double *e_hist, double *T1, double *T2) {
    return guarded(ctx, [&](Ctx &c) {
        post_hf_impl(c, nullptr, 0, ao_ints, Ca, Cb, eps_a, eps_b, n, o, v, uhf, stages, maxiter, conv, e_mp2, e_ccsd, e_t, iters, e_hist,
                     T1, T2);
    });
}

int qcp2_post_hf_mgpu(qcp2_ctx *ctx, const double *ao_ints, const double *Ca, const double *Cb, const double *eps_a,
                      const double *eps_b, int n, int o, int v, int uhf, int stages, int maxiter, double conv, int root,
                      double *e_mp2, double *e_ccsd, double *e_t, int *iters, double *e_hist, double *T1, double *T2) {
    return guarded(ctx, [&](Ctx &c) {
        QCP2_REQUIRE(!ctx->comm || (root >= 0 && root < ctx->comm->nranks), "post_hf_mgpu: root out of range");
        post_hf_impl(c, ctx->comm, root, ao_ints, Ca, Cb, eps_a, eps_b, n, o, v, uhf, stages, maxiter, conv, e_mp2, e_ccsd, e_t, iters,
                     e_hist, T1, T2);
    });
}

long long qcp2_sharded_gemms(const qcp2_ctx *ctx) { return ctx ? ctx->c.sharded_gemms : 0; }

int qcp2_set_ccsd_sharding(qcp2_ctx *ctx, int enable) {
    if (!ctx) return 2;
    ctx->ccsd_sharding = enable != 0;
    return 0;
}

// ---- building blocks ----------------------------------------------------------------
int qcp2_contract(qcp2_ctx *ctx, double alpha, const double *A, const long long *extA, const char *ia,
                  const double *B, const long long *extB, const char *ib, double beta, double *C,
                  const long long *extC, const char *ic) {
    return guarded(ctx, [&](Ctx &c) {
        QCP2_REQUIRE(A && B && C && ia && ib && ic && extA && extB && extC, "contract: null argument");
        QCP2_REQUIRE(strlen(ia) <= 8 && strlen(ib) <= 8 && strlen(ic) <= 8, "contract: rank > 8");
        TView a, b, cc;
        a.rank = (int) strlen(ia), b.rank = (int) strlen(ib), cc.rank = (int) strlen(ic);
        for (int k = 0; k < a.rank; ++k) a.ext[k] = extA[k];
        for (int k = 0; k < b.rank; ++k) b.ext[k] = extB[k];
        for (int k = 0; k < cc.rank; ++k) cc.ext[k] = extC[k];
        In ain(c, A, a.size()), bin(c, B, b.size());
        // C is read when beta != 0
        DBuf cbuf;
        double *cp = C;
        if (c.pointer_mode == QCP2_HOST) {
            cbuf = DBuf(c, (size_t) std::max<long long>(cc.size(), 1));
            QCP2_CUDA(cudaMemcpyAsync(cbuf.p, C, (size_t) cc.size() * 8, cudaMemcpyHostToDevice, c.stream));
            cp = cbuf.p;
        }
        a.p = const_cast<double *>(ain.p), b.p = const_cast<double *>(bin.p), cc.p = cp;
        contract(c, alpha, a, ia, b, ib, beta, cc, ic);
        if (c.pointer_mode == QCP2_HOST)
            QCP2_CUDA(cudaMemcpyAsync(C, cp, (size_t) cc.size() * 8, cudaMemcpyDeviceToHost, c.stream));
        c.sync();
    });
}

int qcp2_dgemm(qcp2_ctx *ctx, int transA, int transB, int M, int N, int K, double alpha, const double *A,
               long long lda, const double *B, long long ldb, double beta, double *C, long long ldc) {
    return guarded(ctx, [&](Ctx &c) {
        QCP2_REQUIRE(A && B && C && M >= 0 && N >= 0 && K >= 0, "dgemm: bad argument");
        QCP2_REQUIRE(lda >= (transA ? M : K) && ldb >= (transB ? K : N) && ldc >= N, "dgemm: leading dimension too small");
        const long long na = (long long) (transA ? K : M) * lda, nb = (long long) (transB ? N : K) * ldb,
                        nc = (long long) M * ldc;
        In a(c, A, na), b(c, B, nb);
        DBuf cbuf;
        double *cp = C;
        if (c.pointer_mode == QCP2_HOST) {
            cbuf = DBuf(c, (size_t) std::max<long long>(nc, 1));
            QCP2_CUDA(cudaMemcpyAsync(cbuf.p, C, (size_t) nc * 8, cudaMemcpyHostToDevice, c.stream));
            cp = cbuf.p;
        }
        GemmArgs g;
        g.M = M, g.N = N, g.K = K, g.alpha = alpha, g.beta = beta;
        g.A = a, g.lda = lda, g.B = b, g.ldb = ldb, g.C = cp, g.ldc = ldc;
        gemm(c, g, !transA, !transB);
        if (c.pointer_mode == QCP2_HOST) QCP2_CUDA(cudaMemcpyAsync(C, cp, (size_t) nc * 8, cudaMemcpyDeviceToHost, c.stream));
        c.sync();
    });
}

int qcp2_dgemm_ex(qcp2_ctx *ctx, int transA, int transB, int M, int N, int K, double alpha, const double *A, long long lda,
                  const double *B, long long ldb, double beta, double *C, long long ldc, int batch, long long strideA,
                  long long strideB, long long strideC, int tile, int k_splits, int reps, double *seconds) {
    return guarded(ctx, [&](Ctx &c) {
        QCP2_REQUIRE(A && B && C && M >= 0 && N >= 0 && K >= 0 && batch >= 1, "dgemm_ex: bad argument");
        QCP2_REQUIRE(c.pointer_mode == QCP2_DEVICE, "dgemm_ex: device pointers only");
        GemmArgs g;
        g.M = M, g.N = N, g.K = K, g.alpha = alpha, g.beta = beta, g.batch = batch;
        g.A = A, g.lda = lda, g.strideA = strideA, g.B = B, g.ldb = ldb, g.strideB = strideB, g.C = C, g.ldc = ldc, g.strideC = strideC;
        g.force_tile = tile, g.force_split = k_splits;
        cudaEvent_t e0, e1;
        QCP2_CUDA(cudaEventCreate(&e0));
        QCP2_CUDA(cudaEventCreate(&e1));
        QCP2_CUDA(cudaEventRecord(e0, c.stream));
        for (int r = 0; r < std::max(reps, 1); ++r) gemm(c, g, !transA, !transB);
        QCP2_CUDA(cudaEventRecord(e1, c.stream));
        QCP2_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaEventDestroy(e0), cudaEventDestroy(e1);
        if (seconds) *seconds = 1e-3 * ms / std::max(reps, 1);
        c.sync();
    });
}

int qcp2_gemm_log(qcp2_ctx *ctx, int enable) {
    if (!ctx) return 2;
    ctx->c.log_gemms = enable != 0;
    ctx->c.gemm_log.clear();
    return 0;
}

long long qcp2_gemm_log_read(const qcp2_ctx *ctx, long long *recs, long long cap) {
    if (!ctx) return 0;
    const long long n = (long long) ctx->c.gemm_log.size();
    for (long long k = 0; recs && k < n && k < cap; ++k) {
        const Ctx::GemmRec &r = ctx->c.gemm_log[(size_t) k];
        const long long v[8] = {r.M, r.N, r.K, r.batch, r.beta, r.tile_m, r.tile_n, r.split};
        for (int q = 0; q < 8; ++q) recs[8 * k + q] = v[q];
    }
    return n;
}

int qcp2_so_block(qcp2_ctx *ctx, const double *aa, const double *bb, const double *ab, int n, int o, int v, const char *spec, double *out) {
    return guarded(ctx, [&](Ctx &c) {
        QCP2_REQUIRE(aa && bb && ab && spec && out && n >= 1 && strlen(spec) == 4, "so_block: bad argument");
        const long long n4 = ipow4(n);
        long long total = 1;
        for (int d = 0; d < 4; ++d) total *= (spec[d] == 'o') ? o : v;
        In a(c, aa, n4), b(c, bb, n4), m(c, ab, n4);
        Out res(c, out, total);
        so_block(c, a, b, m, n, o, v, spec, res);
        res.finish();
        c.sync();
    });
}

int qcp2_ccsd_set_start_T1(qcp2_ctx *ctx, const double *T1, int o, int v) {
    return guarded(ctx, [&](Ctx &c) {
        QCP2_REQUIRE(o >= 0 && v >= 0, "ccsd_set_start_T1: negative extent");
        c.ccsd_t1_start = nullptr;
        ctx->t1_start.reset();
        if (!T1) return;
        const long long n = (long long) o * v;
        ctx->t1_start = DBuf(c, (size_t) std::max<long long>(n, 1));
        QCP2_CUDA(cudaMemcpyAsync(ctx->t1_start.p, T1, (size_t) n * 8,
                                  c.pointer_mode == QCP2_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c.stream));
        c.sync();
        c.ccsd_t1_start = ctx->t1_start.p;
    });
}

int qcp2_set_ccsd_diis(qcp2_ctx *ctx, int nvec) {
    if (!ctx || nvec < 0 || nvec == 1 || nvec > 12) return 2;
    ctx->c.ccsd_diis = nvec;
    return 0;
}

int qcp2_ccsd_last_mode(const qcp2_ctx *ctx) { return ctx ? ctx->c.last_ccsd_mode : -1; }

long long qcp2_pool_high_water(const qcp2_ctx *ctx) {
    if (!ctx || !ctx->c.pool) return -1;
    unsigned long long hi = 0;
    if (cudaMemPoolGetAttribute(ctx->c.pool, cudaMemPoolAttrUsedMemHigh, &hi) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return (long long) hi;
}

long long qcp2_gemm_counts(const qcp2_ctx *ctx, long long *legacy) {
    if (!ctx) return 0;
    if (legacy) *legacy = ctx->c.legacy_gemms;
    return ctx->c.ws_gemms;
}

int qcp2_synth_t_inputs(qcp2_ctx *ctx, int o, int v, unsigned long long seed, double *T1, double *T2, double *oovv,
                        double *vovv, double *ovoo, double *eps_o, double *eps_v) {
    return guarded(ctx, [&](Ctx &c) {
        synth_t_inputs(c, o, v, seed, T1, T2, oovv, vovv, ovoo, eps_o, eps_v);
        c.sync();
    });
}

}// extern "C"
