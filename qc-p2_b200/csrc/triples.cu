// triples.cu -- (T) energy, cc.cpp:454-560, restated per occupied triple (i,j,k):
//
//   A[a,:]  = [ T2[j,k,a,:] | -T2[i,k,a,:] | -T2[j,i,a,:] | -ovoo[:,a,j,k] | ovoo[:,a,i,k] | ovoo[:,a,j,i] ]
//   B[:,bc] = [ vovv[:,i,b,c] ; vovv[:,j,b,c] ; vovv[:,k,b,c] ; T2[i,:,b,c] ; T2[j,:,b,c] ; T2[k,:,b,c] ]
//   X = A.B                                   (cc.cpp:501,505,509,513,517,521: one GEMM, K = 3(v+o))
//   W[a,b,c] = X[a,b,c] - X[b,a,c] - X[c,b,a] (the other twelve contract calls, cc.cpp:502-523)
//   Y[a,b,c] = t_ia<jk||bc> - t_ja<ik||bc> - t_ka<ji||bc>,  V = Y - Y[b,a,c] - Y[c,b,a]   (cc.cpp:479-489)
//   E(T) = 1/36 sum_ijk sum_abc W (W + V) / D_ijkabc                                       (cc.cpp:541-556)
//
// LITERAL mode sums all o^3 ordered triples and all v^3 (a,b,c) and assumes nothing about
// the inputs (this is what the as-written UHF path needs).  SYMMETRIC mode
// requires the physical antisymmetries of T2, <oo||vv>, <vo||vv>, <ov||oo>, packs b<c
// columns, visits i<j<k and a<b<c only and weights by 36/36.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/qcp2_b200.h"
#include "comm.cuh"
#include "tensor_ops.cuh"
#include "triples.cuh"

namespace qcp2 {

namespace {

constexpr int kBlock = 256;
constexpr int kBTile = 8;// b values per energy block (one per warp)

#define GRID_STRIDE(lin, n) \
    for (long long lin = (long long) blockIdx.x * blockDim.x + threadIdx.x; lin < (n); lin += (long long) gridDim.x * blockDim.x)

__host__ __device__ __forceinline__ long long pair_index(long long b, long long c, long long v) {
    return b * v - b * (b + 1) / 2 + (c - b - 1);// b < c
}

// dst[(r1,r0) or (r0,r1)][q] = src[r0][r1][b][c], rows `ld` doubles apart (ld even: 16-byte rows for the bulk copies);
// packed: q = p(b<c) (symmetric mode), else q = b v + c (literal mode: every (b,c))
__global__ void pack_bc_kernel(const double *__restrict__ src, long long R0, long long R1, int swap01, int v, int packed,
                               long long ld, double *__restrict__ dst) {
    const long long n = R0 * R1 * v * v;
    GRID_STRIDE(lin, n) {
        long long r = lin;
        const long long c = r % v;
        r /= v;
        const long long b = r % v;
        r /= v;
        if (packed && b >= c) continue;
        const long long r1 = r % R1, r0 = r / R1;
        const long long row = swap01 ? r1 * R0 + r0 : r0 * R1 + r1;
        dst[row * ld + (packed ? pair_index(b, c, v) : b * v + c)] = src[lin];
    }
}

// one streamed slab: src[e][b][c] (v^3 doubles, unpacked) -> dst[e][p(b<c)], and the antisymmetry
// defect max |x[e,b,c] + x[e,c,b]| / max |x| of the slab accumulated into res[0..1]
__global__ void pack_slab_kernel(const double *__restrict__ src, int v, double *__restrict__ dst, long long np,
                                 unsigned long long *__restrict__ res) {
    const long long n = (long long) v * v * v;// np = row stride of dst (>= v(v-1)/2, even)
    double defect = 0.0, mx = 0.0;
    GRID_STRIDE(lin, n) {
        long long r = lin;
        const long long c = r % v;
        r /= v;
        const long long b = r % v, e = r / v;
        if (b > c) continue;
        const double x = src[lin];
        if (b == c) {
            defect = fmax(defect, fabs(2.0 * x)), mx = fmax(mx, fabs(x));
            continue;
        }
        const double y = src[(e * v + c) * v + b];
        defect = fmax(defect, fabs(x + y)), mx = fmax(mx, fmax(fabs(x), fabs(y)));
        dst[e * np + pair_index(b, c, v)] = x;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        defect = fmax(defect, __shfl_xor_sync(0xffffffffu, defect, s));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, s));
    }
    if ((threadIdx.x & 31) == 0) {// non-negative doubles order like their bit patterns
        atomicMax(res, (unsigned long long) __double_as_longlong(defect));
        atomicMax(res + 1, (unsigned long long) __double_as_longlong(mx));
    }
}

struct BuildAArgs {
    const double *T2, *ovoo;
    const int *ijk;
    double *A;
    int o, v, kpad, count;
    int tiled, mtiles;// tiled: A[t][m-tile][k-block][TG_BM][TG_ALD] (the TMA kernel's stage image)
    int merged;       // the three o-row (hole) segments form ONE block of 3 o rows behind segment 3 (K padded once, not three times)
    int kb_begin[GEMM_MAX_SEG + 1];
};

__global__ void build_A_kernel(BuildAArgs p) {
    const long long O = p.o, V = p.v;
    const int nkb = p.kpad / GEMM_BK;
    const long long per = p.tiled ? (long long) p.mtiles * nkb * TG_BM * TG_ALD : (long long) p.v * p.kpad;
    const long long n = per * p.count;
    GRID_STRIDE(lin, n) {
        const int t = (int) (lin / per);
        long long rem = lin - (long long) t * per;
        long long a;
        int col;
        if (p.tiled) {
            const int c = (int) (rem % TG_ALD);
            rem /= TG_ALD;
            const int r = (int) (rem % TG_BM);
            rem /= TG_BM;
            const int kb = (int) (rem % nkb);
            const int mt = (int) (rem / nkb);
            a = (long long) mt * TG_BM + r;
            col = kb * GEMM_BK + c;
            if (c >= GEMM_BK || a >= V) {
                p.A[lin] = 0.0;
                continue;
            }
        } else {
            a = rem / p.kpad;
            col = (int) (rem - a * p.kpad);
        }
        const int i = p.ijk[3 * t], j = p.ijk[3 * t + 1], k = p.ijk[3 * t + 2];
        int s = 0;
#pragma unroll
        for (int q = 1; q < GEMM_MAX_SEG; ++q)
            if (col >= p.kb_begin[q] * GEMM_BK) s = q;
        long long r = col - p.kb_begin[s] * GEMM_BK;
        if (p.merged && s >= 3) {// row r of the merged hole block belongs to sub-segment r / o
            s = 3 + (int) (r / O);
            r -= (long long) (s - 3) * O;
            if (s > 5) s = 5, r = O;// padding rows
        }
        double val = 0.0;
        if (s < 3) {
            if (r < V) {
                const long long x = s == 0 ? j : (s == 1 ? i : j), y = s == 0 ? k : (s == 1 ? k : i);
                val = p.T2[((x * O + y) * V + a) * V + r];
                if (s != 0) val = -val;
            }
        } else if (r < O) {
            const long long x = s == 3 ? j : (s == 4 ? i : j), y = s == 3 ? k : (s == 4 ? k : i);
            val = p.ovoo[((r * V + a) * O + x) * O + y];
            if (s == 3) val = -val;
        }
        p.A[lin] = val;
    }
}

struct EnergyArgs {
    const double *X;        // [count][v][ncols]
    const double *T1;       // [o][v]
    const double *oovv;     // packed [o][o][np] or dense [o][o][v][v]
    const double *eo, *ev;
    const int *ijk;
    double *partial;        // [count][blocks_per_triple]
    int o, v, nbt;
    long long ncols;
};

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) x += __shfl_xor_sync(0xffffffffu, x, s);
    return x;
}

// One block per (a, tile of 8 b values); warp w owns b = 8*bt + w, lanes sweep c.
template<bool PACKED>
__global__ void __launch_bounds__(kBlock) t_energy_kernel(EnergyArgs p) {
    __shared__ double ws[kBlock / 32];
    const int t = blockIdx.z;
    const int a = blockIdx.x / p.nbt, bt = blockIdx.x % p.nbt;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = bt * kBTile + warp;
    const long long V = p.v, O = p.o, nc = p.ncols;
    const int i = p.ijk[3 * t], j = p.ijk[3 * t + 1], k = p.ijk[3 * t + 2];
    const double *X = p.X + (long long) t * V * nc;
    double sum = 0.0;
    const bool active = PACKED ? (b < p.v && a < b) : (b < p.v);
    if (active) {
        const long long blk = PACKED ? nc : V * V;
        const double *Ojk = p.oovv + (j * O + k) * blk, *Oik = p.oovv + (i * O + k) * blk,
                     *Oji = p.oovv + (j * O + i) * blk;
        const double tia = p.T1[i * V + a], tja = p.T1[j * V + a], tka = p.T1[k * V + a];
        const double tib = p.T1[i * V + b], tjb = p.T1[j * V + b], tkb = p.T1[k * V + b];
        const double dab = p.eo[i] + p.eo[j] + p.eo[k] - p.ev[a] - p.ev[b];
        if (PACKED) {
            const long long pab = pair_index(a, b, V);
            const double ojk_ab = Ojk[pab], oik_ab = Oik[pab], oji_ab = Oji[pab];
            const double *Xa = X + a * nc, *Xb = X + b * nc;
            for (int c = b + 1 + lane; c < p.v; c += 32) {
                const long long pbc = pair_index(b, c, V), pac = pair_index(a, c, V);
                const double W = Xa[pbc] - Xb[pac] + X[c * nc + pab];
                const double tic = p.T1[i * V + c], tjc = p.T1[j * V + c], tkc = p.T1[k * V + c];
                const double Ya = tia * Ojk[pbc] - tja * Oik[pbc] - tka * Oji[pbc];
                const double Yb = tib * Ojk[pac] - tjb * Oik[pac] - tkb * Oji[pac];
                const double Yc = tic * ojk_ab - tjc * oik_ab - tkc * oji_ab;
                const double Vv = Ya - Yb + Yc;
                sum += W * (W + Vv) / (dab - p.ev[c]);
            }
        } else {
            const double *Xab = X + a * nc + b * V, *Xba = X + b * nc + a * V;// X rows are nc (>= v^2) apart
            const double *ojk_bc = Ojk + b * V, *oik_bc = Oik + b * V, *oji_bc = Oji + b * V;
            const double *ojk_ac = Ojk + a * V, *oik_ac = Oik + a * V, *oji_ac = Oji + a * V;
            for (int c = lane; c < p.v; c += 32) {
                const double W = Xab[c] - Xba[c] - X[c * nc + b * V + a];
                const double tic = p.T1[i * V + c], tjc = p.T1[j * V + c], tkc = p.T1[k * V + c];
                const double Ya = tia * ojk_bc[c] - tja * oik_bc[c] - tka * oji_bc[c];
                const double Yb = tib * ojk_ac[c] - tjb * oik_ac[c] - tkb * oji_ac[c];
                const double Yc = tic * Ojk[b * V + a] - tjc * Oik[b * V + a] - tkc * Oji[b * V + a];
                const double Vv = Ya - Yb - Yc;
                sum += W * (W + Vv) / (dab - p.ev[c]);
            }
        }
    }
    sum = warp_sum(sum);
    if (lane == 0) ws[warp] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kBlock / 32; ++w) s += ws[w];
        p.partial[(long long) t * gridDim.x + blockIdx.x] = s;
    }
}

// Symmetric-mode energy sweep, tiled.  For fixed a the three X terms of
//   W[a,b,c] = X[a,p(b,c)] - X[b,p(a,c)] + X[c,p(a,b)]          (a < b < c)
// are the packed row a plus M_a - M_a^T of the square matrix M_a[r][s] = X[r, p(a,s)] (r,s > a),
// whose rows are contiguous.  One block (4 warps) owns one work item (a, bt <= ct): a 32x32 tile
// pair of M_a.  The (b,c) tile of M_a and of row a is read straight from global memory, the (c,b)
// tile is transposed through shared memory -- every element of X is fetched exactly once, with
// coalesced requests.  Column data (three <oo||vv> values, three T1 values, eps_c) live in
// registers, row data in shared memory; all offsets inside one X slice are 32-bit.  Work items are
// enumerated on the host, so no empty blocks are launched.
constexpr int kEBlock = 128;
struct EnergySymArgs {
    const double *X, *T1, *oovvP, *eo, *ev;
    const int *ijk;
    const int *items;// [n_items][3] = (a, bt, ct)
    double *partial;
    int o, v;
    long long np;
};

template<int EB, int MINB>
__global__ void __launch_bounds__(EB, MINB) t_energy_sym_kernel(EnergySymArgs p) {
    __shared__ double Mt[32][33];
    __shared__ double qb[3][32], tbs[3][32], evb[32];
    __shared__ double ws[EB / 32];
    const int t = blockIdx.y;
    const int a = p.items[3 * blockIdx.x], bt = p.items[3 * blockIdx.x + 1], ct = p.items[3 * blockIdx.x + 2];
    const int s0 = a + 1, v = p.v;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const long long V = p.v, O = p.o, np = p.np;
    const int i = p.ijk[3 * t], j = p.ijk[3 * t + 1], k = p.ijk[3 * t + 2];
    const double *X = p.X + (long long) t * V * np;
    const double *O0 = p.oovvP + (j * O + k) * np, *O1 = p.oovvP + (i * O + k) * np, *O2 = p.oovvP + (j * O + i) * np;
    const int inp = (int) np;
    const int pa = (int) pair_index(a, s0, V) - s0;// p(a, s) = pa + s
    const int b0 = s0 + bt * 32, c0 = s0 + ct * 32, c = c0 + tx;
    const bool c_ok = c < v;
#pragma unroll
    for (int rr = ty; rr < 32; rr += EB / 32) {// M_a[c rows][b cols], transposed on read-out
        const int row = c0 + rr, col = b0 + tx;
        Mt[rr][tx] = (row < v && col < v) ? X[row * inp + pa + col] : 0.0;
    }
    if (threadIdx.x < 96) {// (EB >= 128)
        const int w = threadIdx.x >> 5;
        const double *Owp = w == 0 ? O0 : (w == 1 ? O1 : O2);
        const long long ow = w == 0 ? i : (w == 1 ? j : k);
        const bool ok = b0 + tx < v;
        qb[w][tx] = ok ? Owp[pa + b0 + tx] : 0.0;
        tbs[w][tx] = ok ? p.T1[ow * V + b0 + tx] : 0.0;
    } else {
        evb[tx] = (b0 + tx < v) ? p.ev[b0 + tx] : 0.0;
    }
    const double qc0 = c_ok ? O0[pa + c] : 0.0, qc1 = c_ok ? O1[pa + c] : 0.0, qc2 = c_ok ? O2[pa + c] : 0.0;
    const double tc0 = c_ok ? p.T1[i * V + c] : 0.0, tc1 = c_ok ? p.T1[j * V + c] : 0.0, tc2 = c_ok ? p.T1[k * V + c] : 0.0;
    const double tia = p.T1[i * V + a], tja = p.T1[j * V + a], tka = p.T1[k * V + a];
    const double dac = p.eo[i] + p.eo[j] + p.eo[k] - p.ev[a] - (c_ok ? p.ev[c] : 0.0);
    const double *Xa = X + (long long) a * np;
    __syncthreads();
    double sum = 0.0;
#pragma unroll
    for (int r = ty; r < 32; r += EB / 32) {
        const int b = b0 + r;
        if (b < c && c_ok) {
            const int pbc = b * v - ((b * (b + 1)) >> 1) + (c - b - 1);
            const double W = Xa[pbc] - X[b * inp + pa + c] + Mt[tx][r];
            const double Ya = tia * O0[pbc] - tja * O1[pbc] - tka * O2[pbc];
            const double Yb = tbs[0][r] * qc0 - tbs[1][r] * qc1 - tbs[2][r] * qc2;
            const double Yc = tc0 * qb[0][r] - tc1 * qb[1][r] - tc2 * qb[2][r];
            sum += W * (W + (Ya - Yb + Yc)) / (dac - evb[r]);
        }
    }
    sum = warp_sum(sum);
    if (tx == 0) ws[ty] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < EB / 32; ++w) s += ws[w];
        p.partial[(long long) t * gridDim.x + blockIdx.x] = s;
    }
}

// one block per triple: fixed-order sum of the block partials -> weighted triple energy
__global__ void t_finish_kernel(const double *__restrict__ partial, int bpt, double weight, double *__restrict__ e_run,
                                double *__restrict__ e_triples, const long long *__restrict__ gidx) {
    __shared__ double ws[kBlock / 32];
    const int t = blockIdx.x;
    double s = 0.0;
    for (int q = threadIdx.x; q < bpt; q += blockDim.x) s += partial[(long long) t * bpt + q];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
#pragma unroll
        for (int w = 0; w < kBlock / 32; ++w) tot += ws[w];
        tot *= weight;
        e_run[t] = tot;
        if (e_triples) e_triples[gidx[t]] = tot;
    }
}

// ---- synthetic inputs ---------------------------------------
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    unsigned long long z = x + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double u_rand(unsigned long long seed, unsigned long long tag, unsigned long long idx) {
    const unsigned long long z = splitmix64(seed ^ (tag << 56) ^ idx);
    return (double) (z >> 11) * 0x1.0p-52 - 1.0;
}

// kind 0: plain [n]; 1: antisymmetric in (0,1) and (2,3) of [o,o,v,v]; 2: antisymmetric in
// the last pair of [d0,d1,d2,d2]
__global__ void synth_kernel(double *__restrict__ out, int kind, long long d0, long long d1, long long d2,
                             long long d3, unsigned long long seed, unsigned long long tag, double scale) {
    const long long n = d0 * d1 * d2 * d3;
    GRID_STRIDE(lin, n) {
        double val;
        if (kind == 0) {
            val = scale * u_rand(seed, tag, (unsigned long long) lin);
        } else {
            long long r = lin;
            long long i3 = r % d3;
            r /= d3;
            long long i2 = r % d2;
            r /= d2;
            long long i1 = r % d1;
            long long i0 = r / d1;
            double sign = 1.0;
            if (i2 == i3) sign = 0.0;
            if (i2 > i3) {
                const long long tmp = i2;
                i2 = i3, i3 = tmp, sign = -sign;
            }
            if (kind == 1) {
                if (i0 == i1) sign = 0.0;
                if (i0 > i1) {
                    const long long tmp = i0;
                    i0 = i1, i1 = tmp, sign = -sign;
                }
            }
            const unsigned long long canon = (unsigned long long) (((i0 * d1 + i1) * d2 + i2) * d3 + i3);
            val = sign == 0.0 ? 0.0 : sign * scale * u_rand(seed, tag, canon);
        }
        out[lin] = val;
    }
}
__global__ void ramp_kernel(double *__restrict__ out, int n, double lo, double span) {
    GRID_STRIDE(i, (long long) n) out[i] = lo + span * (double) i / (double) (n > 1 ? n - 1 : 1);
}

}// namespace

void synth_t_inputs(Ctx &ctx, int o, int v, unsigned long long seed, double *T1, double *T2, double *oovv, double *vovv,
                    double *ovoo, double *eps_o, double *eps_v) {
    auto launch = [&](double *out, int kind, long long d0, long long d1, long long d2, long long d3,
                      unsigned long long tag, double scale) {
        const long long n = d0 * d1 * d2 * d3;
        if (n == 0 || !out) return;
        synth_kernel<<<grid_for(n, kBlock, ctx.sm_count), kBlock, 0, ctx.stream>>>(out, kind, d0, d1, d2, d3, seed, tag,
                                                                                 scale);
        ctx.launched("synth");
    };
    launch(T1, 0, 1, 1, o, v, 1, 0.05);
    launch(T2, 1, o, o, v, v, 2, 0.02);
    launch(oovv, 1, o, o, v, v, 3, 0.1);
    launch(vovv, 2, v, o, v, v, 4, 0.1);
    launch(ovoo, 2, o, v, o, o, 5, 0.1);
    if (eps_o) {
        ramp_kernel<<<1, kBlock, 0, ctx.stream>>>(eps_o, o, -2.0, 1.5);
        ctx.launched("ramp");
    }
    if (eps_v) {
        ramp_kernel<<<grid_for(v, kBlock, ctx.sm_count), kBlock, 0, ctx.stream>>>(eps_v, v, 0.1, 2.9);
        ctx.launched("ramp");
    }
}

long long t_count(int o, int mode) {
    const long long O = o;
    return mode == QCP2_T_SYMMETRIC ? O * (O - 1) * (O - 2) / 6 : O * O * O;
}

int t_select_mode(Ctx &ctx, const double *T2, const double *oovv, const double *vovv, const double *ovoo, int o,
                  int v) {
    const double tol = 1e-10;
    double d, m;
    auto ok = [&](const double *x, long long d0, long long d1, long long d2, long long d3, int pair) {
        antisymmetry_defect(ctx, x, d0, d1, d2, d3, pair, &d, &m);
        return d <= tol * m;
    };
    // vovv == nullptr: the caller checks <vo||vv> itself (streamed plans do it slab by slab)
    const bool sym = ok(T2, o, o, v, v, 0) && ok(T2, o, o, v, v, 1) && ok(oovv, o, o, v, v, 0) &&
                     ok(oovv, o, o, v, v, 1) && (!vovv || ok(vovv, v, o, v, v, 1)) && ok(ovoo, o, v, o, o, 1);
    return sym ? QCP2_T_SYMMETRIC : QCP2_T_LITERAL;
}

TPlan::~TPlan() {
    if (slab_thread.joinable()) slab_thread.join();
    if (ctx) {
        if (copy) cudaStreamSynchronize(copy);
        cudaStreamSynchronize(ctx->stream);
        if (aux) cudaStreamSynchronize(aux);
    }
    for (cudaEvent_t e : slab_ready) cudaEventDestroy(e);
    if (vovv_defect) cudaFree(vovv_defect);
    if (copy) cudaStreamDestroy(copy);
    for (int s = 0; s < 2; ++s) {
        if (gemm_done[s]) cudaEventDestroy(gemm_done[s]);
        if (energy_done[s]) cudaEventDestroy(energy_done[s]);
    }
    if (t_begin) cudaEventDestroy(t_begin);
    if (t_end) cudaEventDestroy(t_end);
    for (cudaEvent_t e : gemm_ev) cudaEventDestroy(e);
    if (seg_dev) cudaFree(seg_dev);
    if (ijk_dev) cudaFree(ijk_dev);
    if (items_dev) cudaFree(items_dev);
    if (aux) cudaStreamDestroy(aux);
}

// streamed <vo||vv>, body of the helper thread: every slab in index order on the copy stream -- upload + pack on the owning rank,
// broadcast with a communicator, event; after the last slab the antisymmetry verdict is exchanged.  Same order on every rank.
static void slab_thread_main(TPlan *Pp) {
    TPlan &P = *Pp;
    Ctx &ctx = *P.ctx;
    const long long O = P.o, V = P.v, np = P.ld;
    const TStream &ts = P.src;
    const bool multi = P.src_multi, dealt = multi && ts.dealt;
    const int W = multi ? ts.comm->nranks : 1, me = multi ? ts.comm->rank : 0;
    auto owner = [&](long long i) { return dealt ? (int) (i % W) : ts.root; };
    try {
        QCP2_CUDA(cudaSetDevice(ctx.device));
        for (long long i = 0; i < O; ++i) {
            double *slabP = P.vovvP.p + i * V * np;
            if (P.src_uploads && (!multi || owner(i) == me)) {
                // vovv[e][i][b][c]: v rows of v^2 doubles, o*v^2 apart -> stage[e][b][c]
                QCP2_CUDA(cudaMemcpy2DAsync(P.stage.p, (size_t) (V * V) * 8, ts.vovv_host + i * V * V, (size_t) (O * V * V) * 8,
                                            (size_t) (V * V) * 8, (size_t) V, cudaMemcpyHostToDevice, P.copy));
                pack_slab_kernel<<<grid_for(V * V * V, kBlock, ctx.sm_count, 4), kBlock, 0, P.copy>>>(P.stage.p, P.v, slabP, np, P.vovv_defect);
                QCP2_CUDA(cudaGetLastError());
                ++P.slab_launches;
            }
            if (multi) comm_broadcast(ctx, *ts.comm, slabP, V * np, owner(i), P.copy);
            QCP2_CUDA(cudaEventRecord(P.slab_ready[(size_t) i], P.copy));
            if (i == O - 1 && multi) {// every rank learns the verdict on the (b,c) antisymmetry: the root's, or (dealt) every packer's
                if (dealt) comm_allgather(ctx, *ts.comm, reinterpret_cast<const double *>(P.vovv_defect), P.defect_all.p, 2, P.copy);
                else comm_broadcast_bytes(ctx, *ts.comm, P.vovv_defect, 16, ts.root, P.copy);
            }
            {
                std::lock_guard<std::mutex> lk(P.slab_mutex);
                P.slabs_enqueued = i + 1;
            }
            P.slab_cv.notify_all();
        }
    } catch (const std::exception &e) {
        {
            std::lock_guard<std::mutex> lk(P.slab_mutex);
            P.slab_error = e.what();
        }
        P.slab_cv.notify_all();
    }
}
void t_plan_wait_slab(TPlan &P, long long i) {
    std::unique_lock<std::mutex> lk(P.slab_mutex);
    P.slab_cv.wait(lk, [&] { return P.slabs_enqueued > i || !P.slab_error.empty(); });
    if (!P.slab_error.empty()) throw Error("(T) slab pipeline: " + P.slab_error);
}
void t_plan_join_slabs(TPlan &P) {
    if (P.slab_thread.joinable()) P.slab_thread.join();
    P.ctx->launches += P.slab_launches;
    P.slab_launches = 0;
    if (!P.slab_error.empty()) throw Error("(T) slab pipeline: " + P.slab_error);
}

TPlan *t_plan_create(Ctx &ctx, const double *T1, const double *T2, const double *oovv, const double *vovv,
                     const double *ovoo, const double *eps_o, const double *eps_v, int o, int v, int mode,
                     const TStream *ts) {
    QCP2_REQUIRE(mode == QCP2_T_LITERAL || mode == QCP2_T_SYMMETRIC, "(T): mode must be LITERAL or SYMMETRIC");
    QCP2_REQUIRE((vovv != nullptr) != (ts != nullptr), "(T): give either a device vovv or a stream source");
    if (ts) QCP2_REQUIRE(mode == QCP2_T_SYMMETRIC && v >= 2, "(T): streamed <vo||vv> needs SYMMETRIC mode");
    QCP2_REQUIRE(o >= 1 && v >= 1, "(T): o and v must be positive");
    std::unique_ptr<TPlan> P(new TPlan());
    P->ctx = &ctx, P->o = o, P->v = v, P->mode = mode;
    const long long O = o, V = v;
    const bool sym = mode == QCP2_T_SYMMETRIC;
    P->ncols = sym ? V * (V - 1) / 2 : V * V;
    P->ld = (P->ncols + 1) & ~1LL;// even row stride: every packed row starts on a 16-byte boundary (v = 2, 3 mod 4 included)
    P->ntriples = t_count(o, mode);
    P->T1 = T1, P->T2 = T2, P->oovv = oovv, P->vovv = vovv, P->ovoo = ovoo;
    P->eo = DBuf(ctx, (size_t) o), P->ev = DBuf(ctx, (size_t) v);
    QCP2_CUDA(cudaMemcpyAsync(P->eo.p, eps_o, (size_t) o * 8, cudaMemcpyDeviceToDevice, ctx.stream));
    QCP2_CUDA(cudaMemcpyAsync(P->ev.p, eps_v, (size_t) v * 8, cudaMemcpyDeviceToDevice, ctx.stream));
    {
        const char *env = getenv("QCP2_T_GEMM");// "cpasync" forces the non-TMA kernel (A/B experiments)
        P->merged_o = !(env && std::string(env) == "cpasync");// the TMA kernel fetches B row by row, so the three hole segments
    }                                                          // (o rows each) can share one padded block: K pad 3 x (16 - o % 16) -> once
    for (int s = 0; s < GEMM_MAX_SEG; ++s) {
        long long rows = s < 3 ? V : O;
        if (P->merged_o && s >= 3) rows = s == 3 ? 3 * O : 0;
        P->kb_begin[s + 1] = P->kb_begin[s] + (int) ((rows + GEMM_BK - 1) / GEMM_BK);
    }
    P->kpad = P->kb_begin[GEMM_MAX_SEG] * GEMM_BK;
    if (P->ncols > 0) {
        // Both modes own re-laid copies [i][e][ld] / [i][m][ld] of <vo||vv> and T2 (symmetric: b<c packed; literal: all
        // (b,c)), so the B rows of every triple are contiguous, 16-byte aligned rows for the TMA kernel's bulk copies.
        // The pad column (odd v(v-1)/2 or v^2) is zero and is simply computed along.
        const long long np = P->ld;
        P->vovvP = DBuf(ctx, (size_t) (O * V * np));
        P->T2P = DBuf(ctx, (size_t) (O * O * np));
        if (np != P->ncols) {
            QCP2_CUDA(cudaMemsetAsync(P->vovvP.p, 0, (size_t) (O * V * np) * 8, ctx.stream));
            QCP2_CUDA(cudaMemsetAsync(P->T2P.p, 0, (size_t) (O * O * np) * 8, ctx.stream));
        }
        if (sym) {
            P->oovvP = DBuf(ctx, (size_t) (O * O * np));
            if (np != P->ncols) QCP2_CUDA(cudaMemsetAsync(P->oovvP.p, 0, (size_t) (O * O * np) * 8, ctx.stream));
        }
        auto pack = [&](const double *src, long long R0, long long R1, int swap, double *dst) {
            pack_bc_kernel<<<grid_for(R0 * R1 * V * V, kBlock, ctx.sm_count), kBlock, 0, ctx.stream>>>(src, R0, R1, swap, v, sym ? 1 : 0,
                                                                                                   np, dst);
            ctx.launched("pack_bc");
        };
        if (!ts) pack(vovv, V, O, 1, P->vovvP);// [e][i][b][c] -> [i][e][q]
        pack(T2, O, O, 0, P->T2P);    // [i][m][b][c] -> [i][m][q]
        if (sym) pack(oovv, O, O, 0, P->oovvP);
        if (ts) {
            const bool multi = ts->comm && ts->comm->nranks > 1;
            const bool dealt = multi && ts->dealt;
            const int W = multi ? ts->comm->nranks : 1, me = multi ? ts->comm->rank : 0;
            auto owner = [&](long long i) { return dealt ? (int) (i % W) : ts->root; };// the rank that uploads and packs slab i
            const bool is_root = !multi || dealt || me == ts->root;                  // "this rank uploads slabs"
            if (is_root) QCP2_REQUIRE(ts->vovv_host != nullptr, "(T): a rank that uploads slabs must supply the host <vo||vv>");
            {// highest priority: the short pack / broadcast kernels must not queue behind the GEMM's CTAs
                int least = 0, greatest = 0;
                QCP2_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
                QCP2_CUDA(cudaStreamCreateWithPriority(&P->copy, cudaStreamNonBlocking, greatest));
            }
            QCP2_CUDA(cudaMalloc(&P->vovv_defect, 16));
            QCP2_CUDA(cudaMemsetAsync(P->vovv_defect, 0, 16, ctx.stream));
            if (is_root) P->stage = DBuf(ctx, (size_t) (V * V * V));
            if (dealt) P->defect_all = DBuf(ctx, (size_t) (2 * W)), P->defect_ranks = W;
            // the copy stream starts once the stream-ordered allocations above exist
            cudaEvent_t born;
            QCP2_CUDA(cudaEventCreateWithFlags(&born, cudaEventDisableTiming));
            QCP2_CUDA(cudaEventRecord(born, ctx.stream));
            QCP2_CUDA(cudaStreamWaitEvent(P->copy, born, 0));
            // The slab pipeline is enqueued by a helper thread (see TPlan): a long queue of NCCL broadcasts behind the first uploads
            // makes the host block in ncclBroadcast until the uploads drain (8 ranks: 0.37 s), which must not keep the main thread
            // from launching the triples whose slabs have arrived.
            P->src = *ts, P->src_multi = multi, P->src_uploads = is_root;
            for (long long i = 0; i < O; ++i) {
                cudaEvent_t ready;
                QCP2_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
                P->slab_ready.push_back(ready);
            }
            QCP2_CUDA(cudaStreamSynchronize(ctx.stream));// the packed buffers and `stage` exist before the helper thread touches them
            P->slab_thread = std::thread(slab_thread_main, P.get());
            cudaEventDestroy(born);
        }
    }
    // batch size: enough CTAs for ~48 full waves, bounded by a 4 GB X buffer
    const long long tiles = ((V + TileT::kBM - 1) / TileT::kBM) * ((std::max<long long>(P->ncols, 1) + TileT::kBN - 1) / TileT::kBN);
    long long tb = (48LL * ctx.sm_count + tiles - 1) / tiles;
    const long long xbytes = std::max<long long>(V * P->ld * 8, 8);
    tb = std::min(tb, std::max<long long>(1, (4LL << 30) / xbytes));
    tb = std::max<long long>(1, std::min<long long>(tb, std::min<long long>(4096, std::max<long long>(P->ntriples, 1))));
    if (const char *env = getenv("QCP2_T_BATCH")) tb = std::max<long long>(1, std::min<long long>(atoll(env), std::max<long long>(P->ntriples, 1)));// experiments
    P->tb = (int) tb;
    {
        if (sym) {// work items (a, bt <= ct) of the tiled energy sweep
            QCP2_REQUIRE(V * P->ld < (1LL << 31), "(T): v * v(v-1)/2 must stay below 2^31");
            std::vector<int> items;
            for (int a = 0; a + 2 < v; ++a) {
                const int nt = (v - a - 1 + 31) / 32;
                for (int ct = 0; ct < nt; ++ct)
                    for (int bt = 0; bt <= ct; ++bt) items.push_back(a), items.push_back(bt), items.push_back(ct);
            }
            P->sym_items = (int) (items.size() / 3);
            if (P->sym_items > 0) {
                QCP2_CUDA(cudaMalloc(&P->items_dev, items.size() * sizeof(int)));
                QCP2_CUDA(cudaMemcpy(P->items_dev, items.data(), items.size() * sizeof(int), cudaMemcpyHostToDevice));
            }
        }
        P->blocks_per_triple = sym ? std::max(P->sym_items, 1) : v * ((v + kBTile - 1) / kBTile);
    }
    {
        const char *env = getenv("QCP2_T_GEMM");// "cpasync" forces the non-TMA kernel (A/B experiments)
        const bool want_tma = !(env && std::string(env) == "cpasync");
        const bool rows_aligned = ((reinterpret_cast<uintptr_t>(P->vovvP.p) & 15) == 0) && ((reinterpret_cast<uintptr_t>(P->T2P.p) & 15) == 0);
        P->use_tma = want_tma && rows_aligned && P->ncols > 0;// both modes; the cp.async SEG kernel stays as an A/B option
        QCP2_REQUIRE(P->use_tma == P->merged_o || P->ncols == 0, "(T): packed rows are not 16-byte aligned");
        const long long mtiles = (V + TG_BM - 1) / TG_BM;
        P->a_per_triple = P->use_tma ? mtiles * (P->kpad / GEMM_BK) * TG_BM * TG_ALD : V * P->kpad;
    }
    for (int s = 0; s < 2; ++s) {
        P->Abuf[s] = DBuf(ctx, (size_t) (tb * P->a_per_triple));
        P->Xbuf[s] = DBuf(ctx, (size_t) (tb * V * std::max<long long>(P->ld, 1)));
        P->partial[s] = DBuf(ctx, (size_t) (tb * P->blocks_per_triple));
        QCP2_CUDA(cudaEventCreateWithFlags(&P->gemm_done[s], cudaEventDisableTiming));
        QCP2_CUDA(cudaEventCreateWithFlags(&P->energy_done[s], cudaEventDisableTiming));
    }
    QCP2_CUDA(cudaEventCreate(&P->t_begin));
    QCP2_CUDA(cudaEventCreate(&P->t_end));
    QCP2_CUDA(cudaStreamCreateWithFlags(&P->aux, cudaStreamNonBlocking));
    ctx.sync();
    return P.release();
}

void t_plan_vovv_defect(TPlan &P, double *defect, double *maxabs) {
    QCP2_REQUIRE(P.copy && P.vovv_defect, "(T): not a streamed plan");
    QCP2_CUDA(cudaStreamSynchronize(P.copy));
    if (P.defect_ranks > 0) {// dealt upload: the worst record over the packing ranks
        std::vector<double> h((size_t) (2 * P.defect_ranks));
        QCP2_CUDA(cudaMemcpy(h.data(), P.defect_all.p, h.size() * 8, cudaMemcpyDeviceToHost));
        *defect = 0.0, *maxabs = 0.0;
        for (int r = 0; r < P.defect_ranks; ++r) *defect = std::max(*defect, h[(size_t) (2 * r)]), *maxabs = std::max(*maxabs, h[(size_t) (2 * r + 1)]);
        return;
    }
    unsigned long long h[2];
    QCP2_CUDA(cudaMemcpy(h, P.vovv_defect, 16, cudaMemcpyDeviceToHost));
    long long b0 = (long long) h[0], b1 = (long long) h[1];
    memcpy(defect, &b0, 8);
    memcpy(maxabs, &b1, 8);
}

static void triple_of(long long t, int o, bool sym, int *ijk) {
    if (!sym) {
        ijk[2] = (int) (t % o), ijk[1] = (int) ((t / o) % o), ijk[0] = (int) (t / ((long long) o * o));
        return;
    }
    // lexicographic i<j<k: walk i then j (o is small)
    long long rem = t;
    for (int i = 0; i < o - 2; ++i) {
        const long long ni = (long long) (o - i - 1) * (o - i - 2) / 2;
        if (rem >= ni) {
            rem -= ni;
            continue;
        }
        for (int j = i + 1; j < o - 1; ++j) {
            const long long nj = o - j - 1;
            if (rem >= nj) {
                rem -= nj;
                continue;
            }
            ijk[0] = i, ijk[1] = j, ijk[2] = j + 1 + (int) rem;
            return;
        }
    }
    throw Error("(T): triple index out of range");
}

void t_plan_run(TPlan &P, long long first, long long stride, long long max_count, double *e_triples_dev,
                double *energy_host) {
    Ctx &ctx = *P.ctx;
    QCP2_REQUIRE(first >= 0 && stride >= 1, "(T): bad triple range");
    long long count = first < P.ntriples ? (P.ntriples - first + stride - 1) / stride : 0;
    if (max_count >= 0) count = std::min(count, max_count);
    *energy_host = 0.0;
    P.last_gemm_seconds = P.last_total_seconds = 0.0;
    P.last_gemm_launches = 0;
    const long long launches0 = ctx.launches;
    if (count == 0 || P.ncols == 0) {
        if (!P.slab_ready.empty()) t_plan_join_slabs(P);// a rank without triples still takes part in every slab exchange
        return;
    }
    const bool sym = P.mode == QCP2_T_SYMMETRIC;
    const long long O = P.o, V = P.v, nc = P.ld;// row stride of the packed tensors and of X (the pad column is computed along)

    // host-side descriptors for the whole run
    std::vector<int> ijk((size_t) (3 * count));
    std::vector<SegDesc> seg((size_t) count);
    std::vector<long long> gidx((size_t) count);
    // processing order: the enumeration's own, except for streamed plans, which visit the triples by
    // ascending largest occupied index so that work exists as soon as the first slabs have arrived
    std::vector<long long> order((size_t) count);
    for (long long q = 0; q < count; ++q) order[(size_t) q] = first + q * stride;
    const bool streamed = !P.slab_ready.empty();
    if (streamed) {
        std::vector<int> key((size_t) (3 * count));
        std::vector<long long> pos((size_t) count);
        for (long long q = 0; q < count; ++q) triple_of(order[(size_t) q], P.o, sym, &key[(size_t) (3 * q)]), pos[(size_t) q] = q;
        std::stable_sort(pos.begin(), pos.end(), [&](long long a, long long b) {
            const int *x = &key[(size_t) (3 * a)], *y = &key[(size_t) (3 * b)];
            return x[2] != y[2] ? x[2] < y[2] : (x[1] != y[1] ? x[1] < y[1] : x[0] < y[0]);
        });
        std::vector<long long> sorted((size_t) count);
        for (long long q = 0; q < count; ++q) sorted[(size_t) q] = order[(size_t) pos[(size_t) q]];
        order.swap(sorted);
    }
    bool aligned = true;
    for (long long q = 0; q < count; ++q) {
        const long long t = order[(size_t) q];
        gidx[(size_t) q] = t;
        int *x = &ijk[(size_t) (3 * q)];
        triple_of(t, P.o, sym, x);
        for (int s = 0; s < 3; ++s) {
            const long long occ = x[s];
            seg[(size_t) q].B[s] = P.vovvP.p + occ * V * nc;
            seg[(size_t) q].B[3 + s] = P.T2P.p + occ * O * nc;
        }
        for (int s = 0; s < GEMM_MAX_SEG; ++s)
            aligned = aligned && ((reinterpret_cast<uintptr_t>(seg[(size_t) q].B[s]) & 15) == 0);
    }
    if (P.run_capacity < count) {
        if (P.seg_dev) cudaFree(P.seg_dev);
        if (P.ijk_dev) cudaFree(P.ijk_dev);
        P.seg_dev = nullptr, P.ijk_dev = nullptr;
        QCP2_CUDA(cudaMalloc(&P.seg_dev, (size_t) count * sizeof(SegDesc)));
        QCP2_CUDA(cudaMalloc(&P.ijk_dev, (size_t) count * 3 * sizeof(int)));
        P.run_capacity = count;
    }
    DBuf e_run(ctx, (size_t) count);
    long long *gidx_dev = static_cast<long long *>(ctx.alloc_bytes((size_t) count * 8));
    QCP2_CUDA(cudaMemcpyAsync(P.seg_dev, seg.data(), (size_t) count * sizeof(SegDesc), cudaMemcpyHostToDevice, ctx.stream));
    QCP2_CUDA(cudaMemcpyAsync(P.ijk_dev, ijk.data(), (size_t) count * 3 * sizeof(int), cudaMemcpyHostToDevice, ctx.stream));
    QCP2_CUDA(cudaMemcpyAsync(gidx_dev, gidx.data(), (size_t) count * 8, cudaMemcpyHostToDevice, ctx.stream));

    const long long nbatch = (count + P.tb - 1) / P.tb;
    while ((long long) P.gemm_ev.size() < 2 * nbatch) {
        cudaEvent_t e;
        QCP2_CUDA(cudaEventCreate(&e));
        P.gemm_ev.push_back(e);
    }
    const double weight = sym ? 1.0 : 1.0 / 36.0;// 6 (ijk) x 6 (abc) / 36 for sorted triples
    QCP2_CUDA(cudaEventRecord(P.t_begin, ctx.stream));
    for (long long bq = 0; bq < nbatch; ++bq) {
        const int s = (int) (bq & 1);
        const long long off = bq * P.tb;
        const int cnt = (int) std::min<long long>(P.tb, count - off);
        if (bq >= 2) QCP2_CUDA(cudaStreamWaitEvent(ctx.stream, P.energy_done[s], 0));
        if (streamed) {// the batch's B rows: slabs up to its largest occupied index (the copy stream is in order)
            int kmax = 0;
            for (int q = 0; q < cnt; ++q) kmax = std::max(kmax, ijk[(size_t) (3 * (off + q) + 2)]);
            t_plan_wait_slab(P, kmax);// host side: the event below must have been recorded (enqueued) by the slab thread
            QCP2_CUDA(cudaStreamWaitEvent(ctx.stream, P.slab_ready[(size_t) kmax], 0));
        }
        BuildAArgs ba;
        ba.T2 = P.T2, ba.ovoo = P.ovoo, ba.ijk = P.ijk_dev + 3 * off, ba.A = P.Abuf[s];
        ba.o = P.o, ba.v = P.v, ba.kpad = P.kpad, ba.count = cnt;
        ba.tiled = P.use_tma ? 1 : 0, ba.mtiles = (P.v + TG_BM - 1) / TG_BM, ba.merged = P.merged_o ? 1 : 0;
        for (int q = 0; q <= GEMM_MAX_SEG; ++q) ba.kb_begin[q] = P.kb_begin[q];
        build_A_kernel<<<grid_for((long long) cnt * P.a_per_triple, kBlock, ctx.sm_count), kBlock, 0, ctx.stream>>>(ba);
        ctx.launched("build_A");

        QCP2_CUDA(cudaEventRecord(P.gemm_ev[(size_t) (2 * bq)], ctx.stream));
        if (P.use_tma) {
            TmaGemmArgs g;
            g.At = P.Abuf[s], g.strideA = P.a_per_triple;
            g.C = P.Xbuf[s], g.ldc = nc, g.strideC = V * nc;
            g.M = P.v, g.N = (int) nc, g.nkb = P.kpad / GEMM_BK;
            g.seg = P.seg_dev + off, g.nseg = GEMM_MAX_SEG;
            for (int q = 0; q <= GEMM_MAX_SEG; ++q) g.seg_kb_begin[q] = P.kb_begin[q];
            for (int q = 0; q < GEMM_MAX_SEG; ++q) g.seg_rows[q] = q < 3 ? P.v : P.o, g.seg_ldb[q] = nc;
            if (P.merged_o) g.seg_rows[3] = 3 * P.o, g.seg_rows[4] = g.seg_rows[5] = 0, g.merged_sub_rows = P.o;
            launch_t_gemm_tma(ctx, g, cnt);
        } else {
            GemmArgs g;
            g.M = P.v, g.N = (int) nc, g.K = P.kpad, g.batch = cnt;
            g.A = P.Abuf[s], g.lda = P.kpad, g.strideA = V * P.kpad;
            g.C = P.Xbuf[s], g.ldc = nc, g.strideC = V * nc;
            g.seg = P.seg_dev + off, g.nseg = GEMM_MAX_SEG, g.seg_aligned = aligned;
            for (int q = 0; q <= GEMM_MAX_SEG; ++q) g.seg_kb_begin[q] = P.kb_begin[q];
            for (int q = 0; q < GEMM_MAX_SEG; ++q) {
                g.seg_rows[q] = q < 3 ? P.v : P.o;
                g.seg_ldb[q] = nc;
            }
            gemm(ctx, g, true, true);
        }
        QCP2_CUDA(cudaEventRecord(P.gemm_ev[(size_t) (2 * bq + 1)], ctx.stream));
        ++P.last_gemm_launches;
        QCP2_CUDA(cudaEventRecord(P.gemm_done[s], ctx.stream));

        QCP2_CUDA(cudaStreamWaitEvent(P.aux, P.gemm_done[s], 0));
        static const bool skip_energy = getenv("QCP2_T_SKIP_ENERGY") != nullptr;// experiments only: GEMM-alone timing
        if (skip_energy) {
            QCP2_CUDA(cudaEventRecord(P.energy_done[s], P.aux));
            continue;
        }
        EnergyArgs ea;
        ea.X = P.Xbuf[s], ea.T1 = P.T1, ea.oovv = sym ? P.oovvP.p : P.oovv, ea.eo = P.eo, ea.ev = P.ev;
        ea.ijk = P.ijk_dev + 3 * off, ea.partial = P.partial[s];
        ea.o = P.o, ea.v = P.v, ea.nbt = (P.v + kBTile - 1) / kBTile, ea.ncols = nc;
        if (sym) {
            EnergySymArgs es;
            es.X = P.Xbuf[s], es.T1 = P.T1, es.oovvP = P.oovvP.p, es.eo = P.eo, es.ev = P.ev;
            es.ijk = P.ijk_dev + 3 * off, es.partial = P.partial[s], es.o = P.o, es.v = P.v, es.np = nc;
            es.items = P.items_dev;
            if (P.sym_items > 0) {
                dim3 grid((unsigned) P.sym_items, (unsigned) cnt, 1);
                // <= 56 registers (9 resident blocks); measured under ncu on 5 triples of o=40/v=400: 85 registers
                // (no occupancy hint) 638 us, 56 -> 417 us, 40 -> 449 us, 32 (spills) -> 417 us: past ~50 % occupancy the
                // sweep is bound by the latency of its dependent loads, not by residency
                t_energy_sym_kernel<128, 9><<<grid, 128, 0, P.aux>>>(es);
            } else {
                QCP2_CUDA(cudaMemsetAsync(P.partial[s].p, 0, (size_t) cnt * sizeof(double), P.aux));
            }
        } else {
            dim3 grid((unsigned) (P.v * ((P.v + kBTile - 1) / kBTile)), 1, (unsigned) cnt);
            t_energy_kernel<false><<<grid, kBlock, 0, P.aux>>>(ea);
        }
        ctx.launched("t_energy");
        t_finish_kernel<<<cnt, kBlock, 0, P.aux>>>(P.partial[s], P.blocks_per_triple, weight, e_run.p + off, e_triples_dev,
                                                  gidx_dev + off);
        ctx.launched("t_finish");
        QCP2_CUDA(cudaEventRecord(P.energy_done[s], P.aux));
    }
    if (streamed) t_plan_join_slabs(P);// every slab has been enqueued; the communicator is the main thread's again
    for (int s = 0; s < 2 && s < nbatch; ++s) QCP2_CUDA(cudaStreamWaitEvent(ctx.stream, P.energy_done[s], 0));
    QCP2_CUDA(cudaEventRecord(P.t_end, ctx.stream));
    std::vector<double> e_host((size_t) count);
    QCP2_CUDA(cudaMemcpyAsync(e_host.data(), e_run.p, (size_t) count * 8, cudaMemcpyDeviceToHost, ctx.stream));
    ctx.sync();
    ctx.free(gidx_dev);
    double tot = 0.0;// summed in the enumeration's order: independent of batching, scheduling and the rank count
    if (streamed) {
        std::vector<long long> by_index((size_t) count);
        for (long long q = 0; q < count; ++q) by_index[(size_t) q] = q;
        std::sort(by_index.begin(), by_index.end(), [&](long long a, long long b) { return gidx[(size_t) a] < gidx[(size_t) b]; });
        for (long long q : by_index) tot += e_host[(size_t) q];
    } else {
        for (double x : e_host) tot += x;
    }
    *energy_host = tot;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, P.t_begin, P.t_end);
    P.last_total_seconds = ms * 1e-3;
    for (long long bq = 0; bq < nbatch; ++bq) {
        cudaEventElapsedTime(&ms, P.gemm_ev[(size_t) (2 * bq)], P.gemm_ev[(size_t) (2 * bq + 1)]);
        P.last_gemm_seconds += ms * 1e-3;
    }
    P.last_all_launches = ctx.launches - launches0;
}

}// namespace qcp2
