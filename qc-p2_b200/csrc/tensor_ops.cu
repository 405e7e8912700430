// tensor_ops.cu -- bandwidth-bound kernels of the post-HF path (see tensor_ops.cuh).
// Every kernel is a coalesced grid-stride sweep; reductions are two-stage with a fixed
// grid so the summation order (and therefore the energy bits) is reproducible run to run.
#include <algorithm>
#include <cstring>

#include "tensor_ops.cuh"

namespace qcp2 {

static constexpr int kBlock = 256;
static constexpr int kRedBlocks = 592;// 4 x 148: fixed, so reductions are deterministic

// ---------------------------------------------------------------------------------
// permutation
// ---------------------------------------------------------------------------------
struct PermArgs {
    int rank;
    long long ext[8];
    long long istr[8];
    long long n;
};

__global__ void permute_generic_kernel(double *__restrict__ out, const double *__restrict__ in, PermArgs a,
                                       double alpha, double beta) {
    for (long long lin = (long long) blockIdx.x * blockDim.x + threadIdx.x; lin < a.n;
         lin += (long long) gridDim.x * blockDim.x) {
        long long r = lin, off = 0;
#pragma unroll 1
        for (int d = a.rank - 1; d >= 0; --d) {
            const long long e = a.ext[d];
            const long long q = r / e;
            off += (r - q * e) * a.istr[d];
            r = q;
        }
        const double val = alpha * in[off];
        out[lin] = (beta == 0.0) ? val : val + beta * out[lin];
    }
}

// tiled transpose between the input-contiguous output dim `di` and the last output dim
struct PermTileArgs {
    int rank;          // rank of the "rest" dims
    long long ext[8];  // rest extents
    long long istr[8]; // rest input strides
    long long ostr[8]; // rest output strides
    long long ext_i, ext_o;     // extents of di and of the last dim
    long long istr_o, ostr_i;   // input stride of the last dim; output stride of di
    long long tiles_i, tiles_o, nrest;
};

__global__ void permute_tiled_kernel(double *__restrict__ out, const double *__restrict__ in, PermTileArgs a,
                                     double alpha, double beta) {
    __shared__ double tile[32][33];
    const long long ntiles = a.tiles_i * a.tiles_o * a.nrest;
    for (long long b = blockIdx.x; b < ntiles; b += gridDim.x) {
        long long r = b;
        const long long to = r % a.tiles_o;
        r /= a.tiles_o;
        const long long ti = r % a.tiles_i;
        r /= a.tiles_i;
        long long ibase = 0, obase = 0;
#pragma unroll 1
        for (int d = a.rank - 1; d >= 0; --d) {
            const long long e = a.ext[d];
            const long long q = r / e;
            const long long x = r - q * e;
            ibase += x * a.istr[d], obase += x * a.ostr[d];
            r = q;
        }
        const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;// 32 x 8
        {
            const long long ii = ti * 32 + tx;
#pragma unroll
            for (int yy = 0; yy < 32; yy += 8) {
                const long long oo = to * 32 + yy + ty;
                if (ii < a.ext_i && oo < a.ext_o) tile[yy + ty][tx] = in[ibase + ii + oo * a.istr_o];
            }
        }
        __syncthreads();
        {
            const long long oo = to * 32 + tx;
#pragma unroll
            for (int yy = 0; yy < 32; yy += 8) {
                const long long ii = ti * 32 + yy + ty;
                if (ii < a.ext_i && oo < a.ext_o) {
                    double *dst = out + obase + ii * a.ostr_i + oo;
                    const double val = alpha * tile[tx][yy + ty];
                    *dst = (beta == 0.0) ? val : val + beta * (*dst);
                }
            }
        }
        __syncthreads();
    }
}

void permute_axpby(Ctx &ctx, double *out, const double *in, int rank, const long long *out_ext,
                   const long long *in_stride, double alpha, double beta) {
    QCP2_REQUIRE(rank >= 1 && rank <= 8, "permute: rank out of range");
    // drop unit dims, merge neighbours that are contiguous in the input as well
    long long ext[8], istr[8];
    int r = 0;
    for (int d = 0; d < rank; ++d) {
        if (out_ext[d] == 1) continue;
        if (r > 0 && istr[r - 1] == in_stride[d] * out_ext[d]) {
            ext[r - 1] *= out_ext[d];
            istr[r - 1] = in_stride[d];
        } else {
            ext[r] = out_ext[d], istr[r] = in_stride[d], ++r;
        }
    }
    long long n = 1;
    for (int d = 0; d < rank; ++d) n *= out_ext[d];
    if (n == 0) return;
    if (r == 0) ext[0] = 1, istr[0] = 1, r = 1;
    int di = -1;
    for (int d = 0; d < r; ++d)
        if (istr[d] == 1) di = d;
    if (di >= 0 && di != r - 1 && ext[di] >= 8 && ext[r - 1] >= 8) {
        PermTileArgs a{};
        long long ostr[8];
        ostr[r - 1] = 1;
        for (int d = r - 2; d >= 0; --d) ostr[d] = ostr[d + 1] * ext[d + 1];
        a.rank = 0;
        a.nrest = 1;
        for (int d = 0; d < r - 1; ++d) {
            if (d == di) continue;
            a.ext[a.rank] = ext[d], a.istr[a.rank] = istr[d], a.ostr[a.rank] = ostr[d];
            a.nrest *= ext[d];
            ++a.rank;
        }
        a.ext_i = ext[di], a.ext_o = ext[r - 1];
        a.istr_o = istr[r - 1], a.ostr_i = ostr[di];
        a.tiles_i = (a.ext_i + 31) / 32, a.tiles_o = (a.ext_o + 31) / 32;
        const long long ntiles = a.tiles_i * a.tiles_o * a.nrest;
        const unsigned grid = (unsigned) std::min<long long>(ntiles, (long long) ctx.sm_count * 32);
        permute_tiled_kernel<<<grid, 256, 0, ctx.stream>>>(out, in, a, alpha, beta);
        ctx.launched("permute_tiled");
        return;
    }
    PermArgs a{};
    a.rank = r, a.n = n;
    for (int d = 0; d < r; ++d) a.ext[d] = ext[d], a.istr[d] = istr[d];
    permute_generic_kernel<<<grid_for(n, kBlock, ctx.sm_count), kBlock, 0, ctx.stream>>>(out, in, a, alpha, beta);
    ctx.launched("permute_generic");
}

// ---------------------------------------------------------------------------------
// elementwise
// ---------------------------------------------------------------------------------
#define GRID_STRIDE(lin, n) \
    for (long long lin = (long long) blockIdx.x * blockDim.x + threadIdx.x; lin < (n); lin += (long long) gridDim.x * blockDim.x)

__global__ void axpby_kernel(double *__restrict__ out, const double *__restrict__ x, double alpha, double beta,
                             long long n) {
    GRID_STRIDE(i, n) out[i] = (beta == 0.0) ? alpha * x[i] : alpha * x[i] + beta * out[i];
}
void axpby(Ctx &ctx, double *out, const double *x, double alpha, double beta, long long n) {
    if (n <= 0) return;
    axpby_kernel<<<grid_for(n, kBlock, ctx.sm_count), kBlock, 0, ctx.stream>>>(out, x, alpha, beta, n);
    ctx.launched("axpby");
}
void fill_zero(Ctx &ctx, double *p, long long n) {
    if (n > 0) QCP2_CUDA(cudaMemsetAsync(p, 0, (size_t) n * sizeof(double), ctx.stream));
}

// integrals.cpp:136-167 as a gather.  One block per (i,j) row pair; a warp owns a row k and its
// lanes sweep the spatial index L, each producing the adjacent pair (l = 2L, 2L+1) as one 16-byte
// store: of the two spins of l exactly one can match, so a pair is (val, 0) or (0, val), and a
// whole row is zero when the spin of k matches neither i nor j.  No integer division, reads walk
// rows of the spatial tensors (contiguous in L), writes are fully coalesced.
__global__ void __launch_bounds__(256) to_so_kernel(const double *__restrict__ aa, const double *__restrict__ bb,
                                                    const double *__restrict__ ab, int n, double *__restrict__ so) {
    const int n2 = 2 * n;
    const int i = blockIdx.x / n2, j = blockIdx.x - i * n2;
    const int si = i & 1, sj = j & 1, I = i >> 1, J = j >> 1;
    double2 *out = reinterpret_cast<double2 *>(so + (size_t) blockIdx.x * n2 * n2);
    const size_t n3 = (size_t) n * n * n, nn = (size_t) n * n;
    const double *same = si ? bb : aa;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int k = warp; k < n2; k += nwarps) {
        const int sk = k & 1, K = k >> 1;
        double2 *row = out + (size_t) k * n;
        const double *p_ikj = nullptr, *p_jki = nullptr;// rows [I,K,J,:] and [J,K,I,:]
        int slot = 0;                                   // which element of the pair carries the value
        double sign = 1.0;
        if (si == sj) {
            if (sk == si) p_ikj = same + I * n3 + K * nn + (size_t) J * n, p_jki = same + J * n3 + K * nn + (size_t) I * n, slot = si;// :143-149
        } else if (sk == si) {
            p_ikj = ab + I * n3 + K * nn + (size_t) J * n, slot = sj;                                                                  // :151-155
        } else {
            p_ikj = ab + J * n3 + K * nn + (size_t) I * n, slot = si, sign = -1.0;                                                     // :157-161
        }
        if (!p_ikj) {
            for (int L = lane; L < n; L += 32) row[L] = make_double2(0.0, 0.0);// never-written spin patterns (zero-initialised in the reference)
            continue;
        }
        for (int L0 = lane; L0 < n; L0 += 128) {// four independent load pairs in flight per lane before the first store
            double va[4], vb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int L = L0 + 32 * u;
                va[u] = L < n ? p_ikj[L] : 0.0;
                vb[u] = (p_jki && L < n) ? p_jki[L] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int L = L0 + 32 * u;
                if (L < n) {
                    const double val = sign * (va[u] - vb[u]);
                    row[L] = slot ? make_double2(0.0, val) : make_double2(val, 0.0);
                }
            }
        }
    }
}
// scalar form for output buffers that are not 16-byte aligned
__global__ void to_so_scalar_kernel(const double *__restrict__ aa, const double *__restrict__ bb,
                                    const double *__restrict__ ab, int n, double *__restrict__ so) {
    const int n2 = 2 * n;
    const int i = blockIdx.x / n2, j = blockIdx.x - i * n2;
    const int si = i & 1, sj = j & 1, I = i >> 1, J = j >> 1;
    double *out = so + (size_t) blockIdx.x * n2 * n2;
    const size_t n3 = (size_t) n * n * n, nn = (size_t) n * n;
    const double *same = si ? bb : aa;
    for (int kl = threadIdx.x; kl < n2 * n2; kl += blockDim.x) {
        const int k = kl / n2, l = kl - k * n2;
        const int sk = k & 1, sl = l & 1, K = k >> 1, L = l >> 1;
        const size_t ikjl = I * n3 + K * nn + (size_t) J * n + L, jkil = J * n3 + K * nn + (size_t) I * n + L;
        double val = 0.0;
        if (si == sk && sj == sl) val = (si == sj) ? same[ikjl] - same[jkil] : ab[ikjl];// :143-155
        else if (si == sl && sj == sk && si != sj) val = -ab[jkil];                    // :157-161
        out[kl] = val;// the remaining spin patterns are the zero-initialised ones
    }
}
void to_so(Ctx &ctx, const double *aa, const double *bb, const double *ab, int n, double *so) {
    if (n <= 0) return;
    if ((reinterpret_cast<uintptr_t>(so) & 15) == 0) to_so_kernel<<<4u * n * n, kBlock, 0, ctx.stream>>>(aa, bb, ab, n, so);
    else to_so_scalar_kernel<<<4u * n * n, kBlock, 0, ctx.stream>>>(aa, bb, ab, n, so);
    ctx.launched("to_so");
}

// Column maps of the pair-packed matrices: the natural pair index q(e<f) belongs to class cm.cls[q] and sits at column
// cm.loc[q] of that class's block; a kernel working on class cm.want touches only the pairs of that class.
// Null maps: one class, natural order.
__device__ __forceinline__ bool pair_col(const PairColMap &m, size_t q, long long &col) {
    if (!m.cls) {
        col = (long long) q;
        return true;
    }
    if (m.cls[q] != m.want) return false;
    col = m.loc[q];
    return true;
}
// ---- spin-orbital blocks straight from the spatial MO tensors (never building the (2n)^4 tensor) ----
// <ij||kl> exactly as integrals.cpp:143-161 writes it (the ten unmatched spin patterns are the zero-initialised ones)
__device__ __forceinline__ double so_element(const double *__restrict__ aa, const double *__restrict__ bb, const double *__restrict__ ab,
                                             int n, int i, int j, int k, int l) {
    const int si = i & 1, sj = j & 1, sk = k & 1, sl = l & 1;
    const size_t I = i >> 1, J = j >> 1, K = k >> 1, L = l >> 1;
    const size_t n3 = (size_t) n * n * n, nn = (size_t) n * n;
    const size_t ikjl = I * n3 + K * nn + J * n + L, jkil = J * n3 + K * nn + I * n + L;
    if (si == sk && sj == sl) {
        if (si == sj) {
            const double *same = si ? bb : aa;
            return same[ikjl] - same[jkil];// :143-149
        }
        return ab[ikjl];                   // :151-155
    }
    if (si == sl && sj == sk && si != sj) return -ab[jkil];// :157-161
    return 0.0;
}
// out[p][q][r][s] = <p+m0, q+m1 || r+m2, s+m3>: the slice_ints block (integrals.cpp:184-207) without its source tensor
__global__ void __launch_bounds__(256) so_block_kernel(const double *__restrict__ aa, const double *__restrict__ bb,
                                                       const double *__restrict__ ab, int n, int e1, int e2, int e3, int m0, int m1,
                                                       int m2, int m3, double *__restrict__ out) {
    const int p = blockIdx.x / e1, q = blockIdx.x - p * e1;
    double *dst = out + (size_t) blockIdx.x * e2 * e3;
    const int total = e2 * e3;
    for (int rs = blockIdx.y * blockDim.x + threadIdx.x; rs < total; rs += gridDim.y * blockDim.x) {
        const int r = rs / e3, sidx = rs - r * e3;
        dst[rs] = so_element(aa, bb, ab, n, p + m0, q + m1, r + m2, sidx + m3);
    }
}
void so_block(Ctx &ctx, const double *aa, const double *bb, const double *ab, int n, int o, int v, const char *spec, double *out) {
    QCP2_REQUIRE(spec && strlen(spec) == 4, "so_block: spec must have 4 characters");
    int e[4], m[4];
    for (int d = 0; d < 4; ++d) {
        QCP2_REQUIRE(spec[d] == 'o' || spec[d] == 'v', "so_block: spec characters must be 'o' or 'v'");
        e[d] = spec[d] == 'o' ? o : v;
        m[d] = spec[d] == 'v' ? o : 0;
    }
    QCP2_REQUIRE(o + v <= 2 * n, "so_block: o + v exceeds the spin-orbital count");
    if ((long long) e[0] * e[1] * e[2] * e[3] == 0) return;
    const long long rows = (long long) e[0] * e[1], panel = (long long) e[2] * e[3];
    const unsigned ny = (unsigned) std::max<long long>(1, std::min<long long>((panel + 255) / 256, (8LL * ctx.sm_count + rows - 1) / rows));
    so_block_kernel<<<dim3((unsigned) rows, ny), 256, 0, ctx.stream>>>(aa, bb, ab, n, e[1], e[2], e[3], m[0], m[1], m[2], m[3], out);
    ctx.launched("so_block");
}
// dst[row][q(e<f)] = <p_row + m0, q_row + m1 || e + m2, f + m2> for the (p,q) rows listed in pairs01 (2 ints per row): the
// pair-packed form of a block (pack_pairs of slice_ints of transform_to_so) without any of the three intermediates
__global__ void so_block_packed_kernel(const double *__restrict__ aa, const double *__restrict__ bb, const double *__restrict__ ab,
                                       int n, const int *__restrict__ pairs01, int m0, int m1, int m2, int d23,
                                       double *__restrict__ dst, long long ld, PairColMap cm) {
    const int p = pairs01[2 * blockIdx.x] + m0, q = pairs01[2 * blockIdx.x + 1] + m1;
    double *d = dst + (size_t) blockIdx.x * (size_t) ld;
    for (int e = blockIdx.y; e + 1 < d23; e += gridDim.y) {
        const size_t q0 = (size_t) e * d23 - (size_t) e * (e + 1) / 2 - (e + 1);
        for (int f = e + 1 + threadIdx.x; f < d23; f += blockDim.x) {
            long long col;
            if (pair_col(cm, q0 + f, col)) d[col] = so_element(aa, bb, ab, n, p, q, e + m2, f + m2);
        }
    }
}
void so_block_packed(Ctx &ctx, const double *aa, const double *bb, const double *ab, int n, int m0, int m1, int m2, int d23,
                     const int *pairs01, int n01, double *dst, long long ld, PairColMap cm) {
    if (n01 <= 0 || d23 < 2) return;
    const unsigned ny = n01 >= 8 * ctx.sm_count ? 1u : (unsigned) std::min(d23 - 1, (8 * ctx.sm_count + n01 - 1) / n01);
    so_block_packed_kernel<<<dim3((unsigned) n01, ny), d23 >= 2 * kBlock ? kBlock : 64, 0, ctx.stream>>>(aa, bb, ab, n, pairs01, m0, m1, m2,
                                                                                                     d23, dst, ld, cm);
    ctx.launched("so_block_packed");
}
// antisymmetry of the <vv||vv> block in (a,b) and in (e,f), evaluated from the spatial tensors: res[0..1] = max |x + P_ab x|, max |x|,
// res[2..3] the same for (e,f)  (bit patterns of non-negative doubles, accumulated with atomicMax; zero res first)
__global__ void so_vvvv_defect_kernel(const double *__restrict__ aa, const double *__restrict__ bb, const double *__restrict__ ab, int n,
                                      int o, int v, unsigned long long *__restrict__ res) {
    const long long total = (long long) v * v * v * v;
    double d0 = 0.0, d1 = 0.0, mx = 0.0, forb = 0.0;
    GRID_STRIDE(lin, total) {
        long long r = lin;
        const int f = (int) (r % v);
        r /= v;
        const int e = (int) (r % v);
        r /= v;
        const int b = (int) (r % v);
        const int a = (int) (r / v);
        if (a > b || e > f) continue;// every unordered pair of pairs once
        const double x = so_element(aa, bb, ab, n, a + o, b + o, e + o, f + o);
        d0 = fmax(d0, fabs(x + so_element(aa, bb, ab, n, b + o, a + o, e + o, f + o)));
        d1 = fmax(d1, fabs(x + so_element(aa, bb, ab, n, a + o, b + o, f + o, e + o)));
        mx = fmax(mx, fabs(x));
        if ((((a + o) & 1) + ((b + o) & 1)) != (((e + o) & 1) + ((f + o) & 1))) forb = fmax(forb, fabs(x));
    }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) {
        forb = fmax(forb, __shfl_xor_sync(0xffffffffu, forb, sft));
        d0 = fmax(d0, __shfl_xor_sync(0xffffffffu, d0, sft));
        d1 = fmax(d1, __shfl_xor_sync(0xffffffffu, d1, sft));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, sft));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(res, (unsigned long long) __double_as_longlong(d0));
        atomicMax(res + 1, (unsigned long long) __double_as_longlong(mx));
        atomicMax(res + 2, (unsigned long long) __double_as_longlong(d1));
        atomicMax(res + 3, (unsigned long long) __double_as_longlong(mx));
        atomicMax(res + 4, (unsigned long long) __double_as_longlong(forb));
    }
}
// max |x| over the elements of a 4-index block whose spins do not conserve (s0 + s1 != s2 + s3; spin of index k of
// dimension d = (k + off[d]) & 1: even = alpha, odd = beta, integrals.cpp:143-161) -> res[0] (atomicMax of the bit pattern)
__global__ void spin_pattern_kernel(const double *__restrict__ x, long long d0, long long d1, long long d2, long long d3, int f0, int f1,
                                    int f2, int f3, unsigned long long *__restrict__ res) {
    const long long n = d0 * d1 * d2 * d3;
    double mx = 0.0;
    GRID_STRIDE(lin, n) {
        long long r = lin;
        const int i3 = (int) (r % d3);
        r /= d3;
        const int i2 = (int) (r % d2);
        r /= d2;
        const int i1 = (int) (r % d1);
        const int i0 = (int) (r / d1);
        if ((((i0 + f0) & 1) + ((i1 + f1) & 1)) != (((i2 + f2) & 1) + ((i3 + f3) & 1))) mx = fmax(mx, fabs(x[lin]));
    }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, sft));
    if ((threadIdx.x & 31) == 0) atomicMax(res, (unsigned long long) __double_as_longlong(mx));
}
void spin_pattern_defect_async(Ctx &ctx, const double *x, const long long d[4], const int off[4], unsigned long long *res) {
    const long long n = d[0] * d[1] * d[2] * d[3];
    if (n == 0) return;
    spin_pattern_kernel<<<grid_for(n, kBlock, ctx.sm_count), kBlock, 0, ctx.stream>>>(x, d[0], d[1], d[2], d[3], off[0], off[1], off[2], off[3], res);
    ctx.launched("spin_pattern");
}

void so_vvvv_defect_async(Ctx &ctx, const double *aa, const double *bb, const double *ab, int n, int o, int v, unsigned long long *res) {
    const long long total = (long long) v * v * v * v;
    if (total == 0) return;
    so_vvvv_defect_kernel<<<grid_for(total, kBlock, ctx.sm_count), kBlock, 0, ctx.stream>>>(aa, bb, ab, n, o, v, res);
    ctx.launched("so_vvvv_defect");
}

// integrals.cpp:184-207.  blockIdx.x = (p,q) pair of the block's first two indices, blockIdx.y = a
// chunk of kSliceRows rows r of the (r,s) panel.  A warp copies four rows at a time (four
// independent loads in flight per lane before the first store), lanes sweep the contiguous s.
constexpr int kSliceRows = 32;
__global__ void __launch_bounds__(256) slice_kernel(const double *__restrict__ so, int n2, int e1, int e2, int e3, int m0,
                                                    int m1, int m2, int m3, double *__restrict__ out) {
    const int p = blockIdx.x / e1, q = blockIdx.x - p * e1;
    const double *src = so + ((size_t) (p + m0) * n2 + (q + m1)) * n2 * n2 + (size_t) m2 * n2 + m3;
    double *dst = out + (size_t) blockIdx.x * e2 * e3;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;// 8 warps x 4 rows = kSliceRows
    const int r0 = blockIdx.y * kSliceRows + warp * 4;
    if (r0 >= e2) return;
    const int nr = min(4, e2 - r0);
    const double *s0 = src + (size_t) r0 * n2;
    double *d0 = dst + (size_t) r0 * e3;
    if (nr == 4) {
        for (int s = lane; s < e3; s += 32) {
            const double a = s0[s], b = s0[n2 + s], c = s0[2 * n2 + s], d = s0[3 * n2 + s];
            d0[s] = a, d0[e3 + s] = b, d0[2 * e3 + s] = c, d0[3 * e3 + s] = d;
        }
    } else {
        for (int r = 0; r < nr; ++r)
            for (int s = lane; s < e3; s += 32) d0[r * e3 + s] = s0[r * n2 + s];
    }
}
void slice_block(Ctx &ctx, const double *so, int n2, int o, int v, const char *spec, double *out) {
    QCP2_REQUIRE(spec && strlen(spec) == 4, "slice_ints: spec must have 4 characters");
    int e[4], m[4];
    for (int d = 0; d < 4; ++d) {
        QCP2_REQUIRE(spec[d] == 'o' || spec[d] == 'v', "slice_ints: spec characters must be 'o' or 'v'");
        e[d] = spec[d] == 'o' ? o : v;
        m[d] = spec[d] == 'v' ? o : 0;
    }
    QCP2_REQUIRE(o + v <= n2, "slice_ints: o + v exceeds the spin-orbital count");
    const long long total = (long long) e[0] * e[1] * e[2] * e[3];
    if (total == 0) return;
    const dim3 grid((unsigned) (e[0] * e[1]), (unsigned) ((e[2] + kSliceRows - 1) / kSliceRows));
    slice_kernel<<<grid, 256, 0, ctx.stream>>>(so, n2, e[1], e[2], e[3], m[0], m[1], m[2], m[3], out);
    ctx.launched("slice");
}

__global__ void denom_from_F_kernel(const double *__restrict__ F, long long n2, int o, int v,
                                    double *__restrict__ d) {
    const long long total = (long long) o * o * v * v;
    GRID_STRIDE(lin, total) {
        long long r = lin;
        const int b = (int) (r % v);
        r /= v;
        const int a = (int) (r % v);
        r /= v;
        const int j = (int) (r % o);
        const int i = (int) (r / o);
        d[lin] = F[i * n2 + i] + F[j * n2 + j] - F[(o + a) * n2 + (o + a)] - F[(o + b) * n2 + (o + b)];
    }
}
void make_denom(Ctx &ctx, const double *F_spin, int n2, int o, int v, double *denom) {
    const long long total = (long long) o * o * v * v;
    if (total == 0) return;
    denom_from_F_kernel<<<grid_for(total, kBlock, ctx.sm_count), kBlock, 0, ctx.stream>>>(F_spin, n2, o, v, denom);
    ctx.launched("make_denom");
}

__global__ void so_moes_kernel(const double *__restrict__ ea, const double *__restrict__ eb, int n,
                               double *__restrict__ F) {
    const long long n2 = 2LL * n, total = n2 * n2;
    GRID_STRIDE(lin, total) {
        const long long i = lin / n2, j = lin % n2;
        F[lin] = (i == j) ? ((i & 1) ? eb[i >> 1] : ea[i >> 1]) : 0.0;
    }
}
void make_so_moes(Ctx &ctx, const double *ea, const double *eb, int n, double *F) {
    const long long total = 4LL * n * n;
    if (total == 0) return;
    so_moes_kernel<<<grid_for(total, kBlock, ctx.sm_count), kBlock, 0, ctx.stream>>>(ea, eb, n, F);
    ctx.launched("make_so_moes");
}

__global__ void fock_kernel(const double *__restrict__ e1, const double *__restrict__ e2, int n1, int dim,
                            double *__restrict__ F) {
    const long long total = (long long) dim * dim;
    GRID_STRIDE(lin, total) {
        const long long i = lin / dim, j = lin % dim;
        double val = 0.0;
        if (i == j) {
            const long long base = (n1 % 2 == 0) ? n1 / 2 + i / 2 : n1 / 2 + i / 2 + 1;// cc.cpp:28-31
            val = (i % 2 == 0) ? e1[base] : e2[base];
        }
        F[lin] = val;
    }
}
void make_fock(Ctx &ctx, const double *e1, const double *e2, int n1, int n2, double *F) {
    const int dim = n2 - n1;
    if (dim <= 0) return;
    fock_kernel<<<grid_for((long long) dim * dim, kBlock, ctx.sm_count), kBlock, 0, ctx.stream>>>(e1, e2, n1, dim, F);
    ctx.launched("make_fock");
}

// ---------------------------------------------------------------------------------
// deterministic reductions
// ---------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}
__device__ __forceinline__ double block_sum(double v) {
    __shared__ double ws[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) ws[w] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? ws[threadIdx.x] : 0.0;
    if (w == 0) v = warp_sum(v);
    return v;// valid in thread 0
}
// out[pos[x_ro * V + x_rv] * ld + pos[x_co * V + x_cv]] = alpha * src[x0,x1,x2,x3]: a 4-index tensor re-laid out as a matrix
// over (occupied, virtual) composite indices in the class-sorted order pos[] (ccsd_sym.cu, spin-blocked ring products);
// ro / rv / co / cv select which of the four source indices form the row and column composites.
__global__ void ov_sort4_kernel(const double *__restrict__ src, long long d0, long long d1, long long d2, long long d3, int ro, int rv,
                                int co, int cv, int V, const int *__restrict__ pos, long long ld, double alpha, double *__restrict__ out) {
    const long long n = d0 * d1 * d2 * d3;
    GRID_STRIDE(lin, n) {
        long long r = lin;
        int x[4];
        x[3] = (int) (r % d3);
        r /= d3;
        x[2] = (int) (r % d2);
        r /= d2;
        x[1] = (int) (r % d1);
        x[0] = (int) (r / d1);
        out[(long long) pos[x[ro] * V + x[rv]] * ld + pos[x[co] * V + x[cv]]] = alpha * src[lin];
    }
}
void ov_sort4(Ctx &ctx, const double *src, const long long d[4], int ro, int rv, int co, int cv, int V, const int *pos, long long ld,
              double alpha, double *out) {
    const long long n = d[0] * d[1] * d[2] * d[3];
    if (n == 0) return;
    ov_sort4_kernel<<<grid_for(n, kBlock, ctx.sm_count), kBlock, 0, ctx.stream>>>(src, d[0], d[1], d[2], d[3], ro, rv, co, cv, V, pos, ld, alpha, out);
    ctx.launched("ov_sort4");
}

// DIIS (opt-in; not in the reference, cc.cpp:426-447 is a plain Jacobi loop): out[j] = <e_new, E[j]> for the nvec stored error
// vectors (rows of E, `stride` doubles apart), deterministic two-stage reduction
__global__ void dots_kernel(const double *__restrict__ e_new, const double *__restrict__ E, long long stride, long long n,
                            double *__restrict__ partial) {
    const double *ej = E + (long long) blockIdx.y * stride;
    double s = 0.0;
    GRID_STRIDE(lin, n) s += e_new[lin] * ej[lin];
    s = block_sum(s);
    if (threadIdx.x == 0) partial[(long long) blockIdx.y * gridDim.x + blockIdx.x] = s;
}
__global__ void finish_rows_kernel(const double *__restrict__ partial, int n, double *__restrict__ out) {
    double v = 0.0;
    for (int k = threadIdx.x; k < n; k += blockDim.x) v += partial[(long long) blockIdx.x * n + k];
    v = block_sum(v);
    if (threadIdx.x == 0) out[blockIdx.x] = v;
}
void dots(Ctx &ctx, const double *e_new, const double *E, long long stride, long long n, int nvec, double *out_dev) {
    if (nvec <= 0) return;
    DBuf partial(ctx, (size_t) kRedBlocks * nvec);
    dots_kernel<<<dim3(kRedBlocks, (unsigned) nvec), kBlock, 0, ctx.stream>>>(e_new, E, stride, n, partial);
    ctx.launched("diis_dots");
    finish_rows_kernel<<<nvec, kBlock, 0, ctx.stream>>>(partial, kRedBlocks, out_dev);
    ctx.launched("finish_rows");
}
// out = sum_k c[k] X[k]  (rows of X `stride` doubles apart; c on the device)
__global__ void combine_kernel(const double *__restrict__ X, long long stride, int nvec, const double *__restrict__ c, long long n,
                               double *__restrict__ out) {
    GRID_STRIDE(lin, n) {
        double s = 0.0;
        for (int k = 0; k < nvec; ++k) s += c[k] * X[(long long) k * stride + lin];
        out[lin] = s;
    }
}
void combine(Ctx &ctx, const double *X, long long stride, int nvec, const double *c_dev, long long n, double *out) {
    if (n <= 0 || nvec <= 0) return;
    combine_kernel<<<grid_for(n, kBlock, ctx.sm_count), kBlock, 0, ctx.stream>>>(X, stride, nvec, c_dev, n, out);
    ctx.launched("diis_combine");
}

__global__ void finish_sum_kernel(const double *__restrict__ partial, int n, double *__restrict__ out) {
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) v += partial[i];
    v = block_sum(v);
    if (threadIdx.x == 0) out[0] = v;
}

__global__ void mp2_kernel(const double *__restrict__ oovv, const double *__restrict__ denom, long long n,
                           double *__restrict__ T, double *__restrict__ partial) {
    double e = 0.0;
    GRID_STRIDE(i, n) {
        const double g = oovv[i], d = denom[i];
        if (T) T[i] = g / d;
        e += 0.25 * (g * g) / d;// mp2.cpp:33-34
    }
    e = block_sum(e);
    if (threadIdx.x == 0) partial[blockIdx.x] = e;
}
void mp2_amplitudes_energy(Ctx &ctx, const double *oovv, const double *denom, long long n, double *T,
                           double *e_dev) {
    DBuf partial(ctx, kRedBlocks);
    mp2_kernel<<<kRedBlocks, kBlock, 0, ctx.stream>>>(oovv, denom, n, T, partial);
    ctx.launched("mp2");
    finish_sum_kernel<<<1, kBlock, 0, ctx.stream>>>(partial, kRedBlocks, e_dev);
    ctx.launched("finish_sum");
}

__global__ void ccsd_energy_kernel(const double *__restrict__ ts, const double *__restrict__ td,
                                   const double *__restrict__ oovv, const double *__restrict__ fov, int o, int v,
                                   double *__restrict__ partial) {
    const long long n = (long long) o * o * v * v, n1 = (long long) o * v;
    double e = 0.0;
    GRID_STRIDE(lin, n) {
        long long r = lin;
        const int b = (int) (r % v);
        r /= v;
        const int a = (int) (r % v);
        r /= v;
        const int j = (int) (r % o);
        const int i = (int) (r / o);
        const double g = oovv[lin];
        e += 0.25 * g * td[lin] + 0.5 * g * ts[i * v + a] * ts[j * v + b];// cc.cpp:399
        if (fov != nullptr && lin < n1) e += fov[lin] * ts[lin];         // cc.cpp:396
    }
    e = block_sum(e);
    if (threadIdx.x == 0) partial[blockIdx.x] = e;
}
void ccsd_energy(Ctx &ctx, const double *Ts, const double *Td, const double *oovv, const double *fov, int o, int v,
                 double *e_dev) {
    DBuf partial(ctx, kRedBlocks);
    ccsd_energy_kernel<<<kRedBlocks, kBlock, 0, ctx.stream>>>(Ts, Td, oovv, fov, o, v, partial);
    ctx.launched("ccsd_energy");
    finish_sum_kernel<<<1, kBlock, 0, ctx.stream>>>(partial, kRedBlocks, e_dev);
    ctx.launched("finish_sum");
}

// ---------------------------------------------------------------------------------
// cc.cpp elementwise pieces
// ---------------------------------------------------------------------------------
__global__ void tau_kernel(const double *__restrict__ ts, const double *__restrict__ td, int o, int v, int bar,
                           double *__restrict__ out) {
    const long long n = (long long) o * o * v * v;
    GRID_STRIDE(lin, n) {
        long long r = lin;
        const int b = (int) (r % v);
        r /= v;
        const int a = (int) (r % v);
        r /= v;
        const int j = (int) (r % o);
        const int i = (int) (r / o);
        const double x = ts[i * v + a] * ts[j * v + b], y = ts[i * v + b] * ts[j * v + a];
        out[lin] = bar ? td[lin] + 0.5 * (x - y) : td[lin] + x - y;// Stanton eqs. 10 / 9
    }
}
void make_tau(Ctx &ctx, const double *Ts, const double *Td, int o, int v, bool bar, double *out) {
    const long long n = (long long) o * o * v * v;
    if (n == 0) return;
    tau_kernel<<<grid_for(n, kBlock, ctx.sm_count), kBlock, 0, ctx.stream>>>(Ts, Td, o, v, bar ? 1 : 0, out);
    ctx.launched("make_tau");
}

__global__ void multiply_ts_kernel(const double *__restrict__ ts, int o, int v, double *__restrict__ out) {
    const long long n = (long long) o * v * o * v, ov = (long long) o * v;
    GRID_STRIDE(lin, n) out[lin] = ts[lin / ov] * ts[lin % ov];// out[j,f,n,b] = t[j,f] t[n,b]
}
// out[j,b,n,f] = -1/2 Td[j,n,f,b] - Ts[j,f] Ts[n,b]: the amplitude factor of the last two terms of
// Stanton eq. 8 (cc.cpp:247 and :249 share <mn||ef> and the target), already in (free | contracted) order
__global__ void wmbej_amp_kernel(const double *__restrict__ ts, const double *__restrict__ td, int o, int v,
                                 double *__restrict__ out) {
    __shared__ double tile[32][33];
    // one block = one (j, n) pair and a 32x32 (f, b) tile: Td[j,n,f,b] is read with b fastest, written with f fastest
    const int jn = blockIdx.z, j = jn / o, n = jn - j * o;
    const int f0 = blockIdx.y * 32, b0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;// 32 x 8
    const double *src = td + ((size_t) j * o + n) * v * v;
    for (int r = ty; r < 32; r += 8) {
        const int f = f0 + r, b = b0 + tx;
        if (f < v && b < v) tile[r][tx] = -0.5 * src[(size_t) f * v + b] - (ts ? ts[j * v + f] * ts[n * v + b] : 0.0);
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int b = b0 + r, f = f0 + tx;
        if (f < v && b < v) out[(((size_t) j * v + b) * o + n) * v + f] = tile[tx][r];
    }
}
void wmbej_amplitudes(Ctx &ctx, const double *Ts, const double *Td, int o, int v, double *out) {
    if (o <= 0 || v <= 0) return;
    const dim3 grid((unsigned) ((v + 31) / 32), (unsigned) ((v + 31) / 32), (unsigned) (o * o));
    wmbej_amp_kernel<<<grid, 256, 0, ctx.stream>>>(Ts, Td, o, v, out);
    ctx.launched("wmbej_amplitudes");
}

void multiply_Ts(Ctx &ctx, const double *Ts, int o, int v, double *out) {
    const long long n = (long long) o * v * o * v;
    if (n == 0) return;
    multiply_ts_kernel<<<grid_for(n, kBlock, ctx.sm_count), kBlock, 0, ctx.stream>>>(Ts, o, v, out);
    ctx.launched("multiply_Ts");
}

__global__ void add_offdiag_kernel(double *__restrict__ F, const double *__restrict__ add, int n) {
    const long long total = (long long) n * n;
    GRID_STRIDE(lin, total) {
        if (lin / n != lin % n) F[lin] += add[lin];
    }
}
void add_offdiag(Ctx &ctx, double *F, const double *Fadd, int n) {
    if (n <= 0) return;
    add_offdiag_kernel<<<grid_for((long long) n * n, kBlock, ctx.sm_count), kBlock, 0, ctx.stream>>>(F, Fadd, n);
    ctx.launched("add_offdiag");
}

__global__ void d_ia_kernel(const double *__restrict__ r, const double *__restrict__ foo,
                            const double *__restrict__ fvv, int o, int v, double *__restrict__ out) {
    const long long n = (long long) o * v;
    GRID_STRIDE(lin, n) {
        const int i = (int) (lin / v), a = (int) (lin % v);
        const double d = foo[(long long) i * o + i] - fvv[(long long) a * v + a];
        out[lin] = r ? r[lin] / d : d;
    }
}
__global__ void d_ijab_kernel(const double *__restrict__ r, const double *__restrict__ foo,
                              const double *__restrict__ fvv, int o, int v, double *__restrict__ out) {
    const long long n = (long long) o * o * v * v;
    GRID_STRIDE(lin, n) {
        long long q = lin;
        const int b = (int) (q % v);
        q /= v;
        const int a = (int) (q % v);
        q /= v;
        const int j = (int) (q % o);
        const int i = (int) (q / o);
        const double d = foo[(long long) i * o + i] + foo[(long long) j * o + j] - fvv[(long long) a * v + a] -
                         fvv[(long long) b * v + b];
        out[lin] = r ? r[lin] / d : d;
    }
}
__global__ void d_triples_kernel(const double *__restrict__ foo, const double *__restrict__ fvv, int o, int v,
                                 double *__restrict__ out) {
    const long long n = (long long) o * o * o * v * v * v;
    GRID_STRIDE(lin, n) {
        long long q = lin;
        const int c = (int) (q % v);
        q /= v;
        const int b = (int) (q % v);
        q /= v;
        const int a = (int) (q % v);
        q /= v;
        const int k = (int) (q % o);
        q /= o;
        const int j = (int) (q % o);
        const int i = (int) (q / o);
        out[lin] = foo[(long long) i * o + i] + foo[(long long) j * o + j] + foo[(long long) k * o + k] -
                   fvv[(long long) a * v + a] - fvv[(long long) b * v + b] - fvv[(long long) c * v + c];
    }
}
void make_D_ia(Ctx &ctx, const double *foo, const double *fvv, int o, int v, double *D) {
    divide_D_ia(ctx, nullptr, foo, fvv, o, v, D);
}
void make_D_ijab(Ctx &ctx, const double *foo, const double *fvv, int o, int v, double *D) {
    divide_D_ijab(ctx, nullptr, foo, fvv, o, v, D);
}
void make_D_triples(Ctx &ctx, const double *foo, const double *fvv, int o, int v, double *D) {
    const long long n = (long long) o * o * o * v * v * v;
    if (n == 0) return;
    d_triples_kernel<<<grid_for(n, kBlock, ctx.sm_count), kBlock, 0, ctx.stream>>>(foo, fvv, o, v, D);
    ctx.launched("make_D_triples");
}
void divide_D_ia(Ctx &ctx, const double *r, const double *foo, const double *fvv, int o, int v, double *out) {
    const long long n = (long long) o * v;
    if (n == 0) return;
    d_ia_kernel<<<grid_for(n, kBlock, ctx.sm_count), kBlock, 0, ctx.stream>>>(r, foo, fvv, o, v, out);
    ctx.launched("D_ia");
}
void divide_D_ijab(Ctx &ctx, const double *r, const double *foo, const double *fvv, int o, int v, double *out) {
    const long long n = (long long) o * o * v * v;
    if (n == 0) return;
    d_ijab_kernel<<<grid_for(n, kBlock, ctx.sm_count), kBlock, 0, ctx.stream>>>(r, foo, fvv, o, v, out);
    ctx.launched("D_ijab");
}

__global__ void divide_kernel(const double *__restrict__ r, const double *__restrict__ D, long long n, double *__restrict__ out) {
    GRID_STRIDE(i, n) out[i] = r[i] / D[i];
}
void divide_elementwise(Ctx &ctx, const double *r, const double *D, long long n, double *out) {
    if (n <= 0) return;
    divide_kernel<<<grid_for(n, kBlock, ctx.sm_count), kBlock, 0, ctx.stream>>>(r, D, n, out);
    ctx.launched("divide");
}

__global__ void ring_combine_kernel(const double *__restrict__ Rt, int o, int v, double *__restrict__ r) {
    const long long n = (long long) o * o * v * v;
    GRID_STRIDE(lin, n) {
        long long q = lin;
        const int b = (int) (q % v);
        q /= v;
        const int a = (int) (q % v);
        q /= v;
        const int j = (int) (q % o);
        const int i = (int) (q / o);
        auto at = [&](int x, int c, int d, int y) { return Rt[(((long long) x * v + c) * v + d) * o + y]; };
        r[lin] += at(i, a, b, j) - at(j, a, b, i) - at(i, b, a, j) + at(j, b, a, i);
    }
}
void ring_combine(Ctx &ctx, const double *Rt, int o, int v, double *r) {
    const long long n = (long long) o * o * v * v;
    if (n == 0) return;
    ring_combine_kernel<<<grid_for(n, kBlock, ctx.sm_count), kBlock, 0, ctx.stream>>>(Rt, o, v, r);
    ctx.launched("ring_combine");
}

// r[i,j,a,b] = (init ? init : r)[i,j,a,b] + U[i,j,a,b] - U[j,i,a,b] (pair 0) / - U[i,j,b,a] (pair 1):
// the P(ij) / P(ab) partner of a product is the same product with two labels exchanged
__global__ void antisym_add_kernel(const double *__restrict__ U, const double *__restrict__ init, int o, int v, int pair,
                                   double *__restrict__ r) {
    const long long n = (long long) o * o * v * v, vv = (long long) v * v;
    GRID_STRIDE(lin, n) {
        long long q = lin;
        const int b = (int) (q % v);
        q /= v;
        const int a = (int) (q % v);
        q /= v;
        const int j = (int) (q % o);
        const int i = (int) (q / o);
        const long long other = pair == 0 ? ((long long) j * o + i) * vv + (long long) a * v + b
                                          : ((long long) i * o + j) * vv + (long long) b * v + a;
        r[lin] = (init ? init[lin] : r[lin]) + U[lin] - U[other];
    }
}
void antisym_add(Ctx &ctx, const double *U, const double *init, int o, int v, int pair, double *r) {
    const long long n = (long long) o * o * v * v;
    if (n == 0) return;
    antisym_add_kernel<<<grid_for(n, kBlock, ctx.sm_count), kBlock, 0, ctx.stream>>>(U, init, o, v, pair, r);
    ctx.launched("antisym_add");
}

// out[i,a,m,e] = Ts[i,e] Ts[m,a]: multiply_Ts (cc.cpp:158-172) in the (free | contracted) order of cc.cpp:360
__global__ void ring_tt_kernel(const double *__restrict__ ts, int o, int v, double *__restrict__ out) {
    const long long n = (long long) o * v * o * v;
    GRID_STRIDE(lin, n) {
        long long q = lin;
        const int e = (int) (q % v);
        q /= v;
        const int m = (int) (q % o);
        q /= o;
        const int a = (int) (q % v);
        const int i = (int) (q / v);
        out[lin] = ts[i * v + e] * ts[m * v + a];
    }
}
void ring_tt(Ctx &ctx, const double *Ts, int o, int v, double *out) {
    const long long n = (long long) o * v * o * v;
    if (n == 0) return;
    ring_tt_kernel<<<grid_for(n, kBlock, ctx.sm_count), kBlock, 0, ctx.stream>>>(Ts, o, v, out);
    ctx.launched("ring_tt");
}

__global__ void pack_pairs_kernel(const double *__restrict__ src, int d01, int d23, const int *__restrict__ pairs01,
                                  double *__restrict__ dst, long long ld, PairColMap cm) {
    const int a = pairs01[2 * blockIdx.x], b = pairs01[2 * blockIdx.x + 1];
    const double *s = src + ((size_t) a * d01 + b) * d23 * d23;
    double *d = dst + (size_t) blockIdx.x * (size_t) ld;
    for (int e = blockIdx.y; e + 1 < d23; e += gridDim.y) {// blockIdx.y strides the rows so few pairs still fill the GPU
        const size_t q0 = (size_t) e * d23 - (size_t) e * (e + 1) / 2 - (e + 1);// q(e,f) = q0 + f
        for (int f = e + 1 + threadIdx.x; f < d23; f += blockDim.x) {
            long long col;
            if (pair_col(cm, q0 + f, col)) d[col] = s[(size_t) e * d23 + f];
        }
    }
}
void pack_pairs(Ctx &ctx, const double *src, int d01, int d23, const int *pairs01, int n01, double *dst, long long ld, PairColMap cm) {
    if (n01 <= 0 || d23 < 2) return;
    const unsigned ny = n01 >= 8 * ctx.sm_count ? 1u : (unsigned) std::min(d23 - 1, (8 * ctx.sm_count + n01 - 1) / n01);
    pack_pairs_kernel<<<dim3((unsigned) n01, ny), d23 >= 2 * kBlock ? kBlock : 64, 0, ctx.stream>>>(src, d01, d23, pairs01, dst, ld, cm);
    ctx.launched("pack_pairs");
}

// dst[p(i<j)][q(a<b)] = Td[i,j,a,b] + t_ia t_jb - t_ib t_ja (Stanton eq. 9 on the packed pairs only)
__global__ void tau_pack_kernel(const double *__restrict__ ts, const double *__restrict__ td, int o, int v,
                                const int *__restrict__ pairs_o, double *__restrict__ dst, long long ld, double f, PairColMap cm) {
    const int i = pairs_o[2 * blockIdx.x], j = pairs_o[2 * blockIdx.x + 1];
    const double *s = td + ((size_t) i * o + j) * v * v;
    const double *ti = ts + (size_t) i * v, *tj = ts + (size_t) j * v;
    double *d = dst + (size_t) blockIdx.x * (size_t) ld;
    for (int a = blockIdx.y; a + 1 < v; a += gridDim.y) {
        const size_t q0 = (size_t) a * v - (size_t) a * (a + 1) / 2 - (a + 1);
        const double tia = f * ti[a], tja = f * tj[a];
        for (int b = a + 1 + threadIdx.x; b < v; b += blockDim.x) {
            long long col;
            if (pair_col(cm, q0 + b, col)) d[col] = s[(size_t) a * v + b] + tia * tj[b] - ti[b] * tja;
        }
    }
}
void tau_pack(Ctx &ctx, const double *Ts, const double *Td, int o, int v, const int *pairs_o, int npo, double *dst, long long ld,
              bool bar, PairColMap cm) {
    if (npo <= 0 || v < 2) return;
    const unsigned ny = (unsigned) std::min(v - 1, (8 * ctx.sm_count + npo - 1) / npo);
    tau_pack_kernel<<<dim3((unsigned) npo, ny), v >= 2 * kBlock ? kBlock : 64, 0, ctx.stream>>>(Ts, Td, o, v, pairs_o, dst, ld,
                                                                                             bar ? 0.5 : 1.0, cm);
    ctx.launched("tau_pack");
}

// F[m][i] += sum_n G[(i,n)][(m,n)] over the (i,n), (m,n) pairs of class all.want (G: that class's block of the all-pairs
// product, cc.cpp:209; all.cls / all.loc over i*o + n).  Null maps: one class, G is [o*o][ldg].
__global__ void pair_trace_add_kernel(const double *__restrict__ G, long long ldg, int o, PairColMap all, double *__restrict__ F) {
    const int n2 = o * o;
    GRID_STRIDE(lin, (long long) n2) {
        const int m = (int) (lin / o), i = (int) (lin % o);
        double s = 0.0;
        for (int n = 0; n < o; ++n) {
            long long r, c;
            if (pair_col(all, (size_t) i * o + n, r) && pair_col(all, (size_t) m * o + n, c)) s += G[r * ldg + c];
        }
        F[lin] += s;
    }
}
void pair_trace_add(Ctx &ctx, const double *G, long long ldg, int o, PairColMap all, double *F) {
    if (o <= 0) return;
    pair_trace_add_kernel<<<grid_for((long long) o * o, kBlock, ctx.sm_count), kBlock, 0, ctx.stream>>>(G, ldg, o, all, F);
    ctx.launched("pair_trace_add");
}

// The last step of the fused T2 equation (cc.cpp:312-383 assembled in one pass):
//   r = <ij||ab> + P(ab) Uab + P(ij) Uij + P(ij)P(ab) Rt[i,a,j,b] + s(ij) s(ab) LP[i<j][a<b];   T2_new = r / D_ijab
__global__ void finalize_t2_kernel(const double *__restrict__ oovv, const double *__restrict__ Uab, const double *__restrict__ Uij,
                                   const double *__restrict__ Rt, const double *__restrict__ RtS, const int *__restrict__ ovpos, long long ldr,
                                   LadderBlocks lb,
                                   const double *__restrict__ foo, const double *__restrict__ fvv, const double *__restrict__ D,
                                   int o, int v, double *__restrict__ out) {
    const long long n = (long long) o * o * v * v, vv = (long long) v * v;
    GRID_STRIDE(lin, n) {
        long long q = lin;
        const int b = (int) (q % v);
        q /= v;
        const int a = (int) (q % v);
        q /= v;
        const int j = (int) (q % o);
        const int i = (int) (q / o);
        const long long ijba = ((long long) i * o + j) * vv + (long long) b * v + a;
        const long long jiab = ((long long) j * o + i) * vv + (long long) a * v + b;
        auto at = [&](int x, int c, int d, int y) { return Rt[(((long long) x * v + c) * o + y) * v + d]; };// Rt[i,a,j,b] at (i,a,b,j)
        double x = oovv[lin] + (Uab[lin] - Uab[ijba]) + (Uij[lin] - Uij[jiab]);
        x += at(i, a, b, j) - at(j, a, b, i) - at(i, b, a, j) + at(j, b, a, i);
        if (RtS) {// the (ov)^3 part of the ring terms, held as a matrix over class-sorted (i,a) x (j,b) composites
            const long long ia = ovpos[i * v + a], ib = ovpos[i * v + b], ja = ovpos[j * v + a], jb = ovpos[j * v + b];
            x += RtS[ia * ldr + jb] - RtS[ja * ldr + ib] - RtS[ib * ldr + ja] + RtS[jb * ldr + ia];
        }
        if (i != j && a != b) {
            int i0 = i, j0 = j, a0 = a, b0 = b;
            double sign = 1.0;
            if (i0 > j0) {
                const int t = i0;
                i0 = j0, j0 = t, sign = -sign;
            }
            if (a0 > b0) {
                const int t = a0;
                a0 = b0, b0 = t, sign = -sign;
            }
            const long long pij = (long long) i0 * o - (long long) i0 * (i0 + 1) / 2 + (j0 - i0 - 1);
            const long long pab = (long long) a0 * v - (long long) a0 * (a0 + 1) / 2 + (b0 - a0 - 1);
            // the packed ladder family: class blocks [rows of the class][columns of the class]; other classes hold exact zeros
            if (!lb.rowcls) {
                x += sign * lb.lp[0][pij * lb.ld[0] + pab];
            } else {
                const int c = lb.rowcls[pij];
                if (lb.colcls[pab] == c) x += sign * lb.lp[c][(long long) lb.rowloc[pij] * lb.ld[c] + lb.colloc[pab]];
            }
        }
        const double d = D ? D[lin]
                           : foo[(long long) i * o + i] + foo[(long long) j * o + j] - fvv[(long long) a * v + a] - fvv[(long long) b * v + b];
        out[lin] = x / d;
    }
}
void finalize_t2(Ctx &ctx, const double *oovv, const double *Uab, const double *Uij, const double *Rt, const double *RtS,
                 const int *ovpos, long long ldr, const LadderBlocks &lb,
                 const double *foo, const double *fvv, const double *D, int o, int v, double *out) {
    const long long n = (long long) o * o * v * v;
    if (n == 0) return;
    finalize_t2_kernel<<<grid_for(n, kBlock, ctx.sm_count), kBlock, 0, ctx.stream>>>(oovv, Uab, Uij, Rt, RtS, ovpos, ldr, lb, foo, fvv, D, o, v, out);
    ctx.launched("finalize_t2");
}

// WP[p(i<j)][q(m<n)] = W[m,n,i,j]  (W_mnij as the left factor of sum_{m<n} W_mnij tau_mnab on packed pairs)
__global__ void wmnij_pack_kernel(const double *__restrict__ W, int o, const int *__restrict__ pairs_o, int npo,
                                  double *__restrict__ dst, long long ld) {
    const long long n = (long long) npo * npo;
    GRID_STRIDE(lin, n) {
        const int p = (int) (lin / npo), q = (int) (lin % npo);
        const int i = pairs_o[2 * p], j = pairs_o[2 * p + 1], m = pairs_o[2 * q], nn = pairs_o[2 * q + 1];
        dst[(long long) p * ld + q] = W[(((long long) m * o + nn) * o + i) * o + j];
    }
}
void wmnij_pack(Ctx &ctx, const double *W, int o, const int *pairs_o, int npo, double *dst, long long ld) {
    if (npo <= 0) return;
    wmnij_pack_kernel<<<grid_for((long long) npo * npo, kBlock, ctx.sm_count), kBlock, 0, ctx.stream>>>(W, o, pairs_o, npo, dst, ld);
    ctx.launched("wmnij_pack");
}

// W[m,n,i,j] += s(mn) s(ij) alpha WP[p][q] for the pairs (i<j) = list[p], (m<n) = list[q] of ONE class (the packed
// 1/4 tau_ijef <mn||ef> term of W_mnij, cc.cpp:231); every element of W is touched by at most one (p, q, sign) combination
__global__ void wmnij_unpack_add_kernel(const double *__restrict__ WP, long long ld, int o, const int *__restrict__ list, int np,
                                        double alpha, double *__restrict__ W) {
    const long long n = (long long) np * np;
    GRID_STRIDE(lin, n) {
        const int p = (int) (lin / np), q = (int) (lin % np);
        const long long i = list[2 * p], j = list[2 * p + 1], m = list[2 * q], nn = list[2 * q + 1];
        const double x = alpha * WP[(long long) p * ld + q];
        W[((m * o + nn) * o + i) * o + j] += x;
        W[((nn * o + m) * o + i) * o + j] -= x;
        W[((m * o + nn) * o + j) * o + i] -= x;
        W[((nn * o + m) * o + j) * o + i] += x;
    }
}
void wmnij_unpack_add(Ctx &ctx, const double *WP, long long ld, int o, const int *list, int np, double alpha, double *W) {
    if (np <= 0) return;
    wmnij_unpack_add_kernel<<<grid_for((long long) np * np, kBlock, ctx.sm_count), kBlock, 0, ctx.stream>>>(WP, ld, o, list, np, alpha, W);
    ctx.launched("wmnij_unpack_add");
}

// LP[p][col(a<b)] -= sum_m ( Z[p][(a,m)] t[m,b] - Z[p][(b,m)] t[m,a] ) for the rows p of one class: the
// -1/2 P(ab) (tau.<am||ef>).t_old term (cc.cpp:237-238 inside :348) applied directly on the packed pairs.  Z's columns are
// the (a,m) pairs of the same class (am.cls / am.loc over a*o + m); pairs of other classes are exact zeros and are skipped.
__global__ void z_t_sub_kernel(const double *__restrict__ Z, long long ldz, const double *__restrict__ ts, int o, int v,
                               double *__restrict__ LP, long long ld, PairColMap cm, PairColMap am) {
    const double *z = Z + (size_t) blockIdx.x * (size_t) ldz;
    double *d = LP + (size_t) blockIdx.x * (size_t) ld;
    for (int a = blockIdx.y; a + 1 < v; a += gridDim.y) {
        const size_t q0 = (size_t) a * v - (size_t) a * (a + 1) / 2 - (a + 1);
        for (int b = a + 1 + threadIdx.x; b < v; b += blockDim.x) {
            long long col;
            if (!pair_col(cm, q0 + b, col)) continue;
            double acc = 0.0;
            for (int m = 0; m < o; ++m) {
                long long ca, cb;
                if (pair_col(am, (size_t) a * o + m, ca)) acc += z[ca] * ts[(size_t) m * v + b];
                if (pair_col(am, (size_t) b * o + m, cb)) acc -= z[cb] * ts[(size_t) m * v + a];
            }
            d[col] -= acc;
        }
    }
}
void z_t_sub(Ctx &ctx, const double *Z, long long ldz, const double *Ts, int o, int v, int nrows, double *LP, long long ld, PairColMap cm,
             PairColMap am) {
    if (nrows <= 0 || v < 2) return;
    const unsigned ny = (unsigned) std::min(v - 1, (8 * ctx.sm_count + nrows - 1) / nrows);
    z_t_sub_kernel<<<dim3((unsigned) nrows, ny), v >= 2 * kBlock ? kBlock : 64, 0, ctx.stream>>>(Z, ldz, Ts, o, v, LP, ld, cm, am);
    ctx.launched("z_t_sub");
}


__global__ void diag_kernel(const double *__restrict__ F, int n, double *__restrict__ d) {
    GRID_STRIDE(i, (long long) n) d[i] = F[i * n + i];
}
void extract_diag(Ctx &ctx, const double *F, int n, double *d) {
    if (n <= 0) return;
    diag_kernel<<<grid_for(n, kBlock, ctx.sm_count), kBlock, 0, ctx.stream>>>(F, n, d);
    ctx.launched("extract_diag");
}

// ---------------------------------------------------------------------------------
// antisymmetry check (selects the symmetric (T) path)
// ---------------------------------------------------------------------------------
__global__ void antisym_kernel(const double *__restrict__ x, long long d0, long long d1, long long d2, long long d3,
                               int pair, unsigned long long *__restrict__ res) {
    const long long n = d0 * d1 * d2 * d3;
    double defect = 0.0, mx = 0.0;
    GRID_STRIDE(lin, n) {
        long long r = lin;
        const long long i3 = r % d3;
        r /= d3;
        const long long i2 = r % d2;
        r /= d2;
        const long long i1 = r % d1;
        const long long i0 = r / d1;
        const long long other = pair == 0 ? ((i1 * d1 + i0) * d2 + i2) * d3 + i3 : ((i0 * d1 + i1) * d2 + i3) * d3 + i2;
        const double a = x[lin];
        defect = fmax(defect, fabs(a + x[other]));
        mx = fmax(mx, fabs(a));
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
// asynchronous form: accumulates (max) into res[0..1] on the device; the caller zeroes res and reads it
// back with whatever synchronisation point it already has
void antisymmetry_defect_async(Ctx &ctx, const double *x, long long d0, long long d1, long long d2, long long d3, int pair,
                               unsigned long long *res) {
    if (pair == 0) QCP2_REQUIRE(d0 == d1, "antisymmetry check: extents differ");
    else QCP2_REQUIRE(d2 == d3, "antisymmetry check: extents differ");
    const long long n = d0 * d1 * d2 * d3;
    if (n > 0) {
        antisym_kernel<<<grid_for(n, kBlock, ctx.sm_count), kBlock, 0, ctx.stream>>>(x, d0, d1, d2, d3, pair, res);
        ctx.launched("antisym");
    }
}

void antisymmetry_defect(Ctx &ctx, const double *x, long long d0, long long d1, long long d2, long long d3, int pair,
                         double *defect, double *maxabs) {
    if (pair == 0) QCP2_REQUIRE(d0 == d1, "antisymmetry check: extents differ");
    else QCP2_REQUIRE(d2 == d3, "antisymmetry check: extents differ");
    unsigned long long *res = static_cast<unsigned long long *>(ctx.alloc_bytes(16));
    QCP2_CUDA(cudaMemsetAsync(res, 0, 16, ctx.stream));
    const long long n = d0 * d1 * d2 * d3;
    if (n > 0) {
        antisym_kernel<<<grid_for(n, kBlock, ctx.sm_count), kBlock, 0, ctx.stream>>>(x, d0, d1, d2, d3, pair, res);
        ctx.launched("antisym");
    }
    unsigned long long h[2];
    QCP2_CUDA(cudaMemcpyAsync(h, res, 16, cudaMemcpyDeviceToHost, ctx.stream));
    ctx.sync();
    ctx.free(res);
    long long b0 = (long long) h[0], b1 = (long long) h[1];
    memcpy(defect, &b0, 8);
    memcpy(maxabs, &b1, 8);
}

}// namespace qcp2
