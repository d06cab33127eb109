// Column kernels: everything the reference does along z or pointwise in
// (kx, ky) on spectral data, fused per "4-group" of columns
//   { (b, a), (b, ny-a), (nx-b, a), (nx-b, ny-a) }
// so that diffx/diffy (which couple Hermitian slot k with slot n-k,
// reference sta3dfft.f90:304-377, via mpi_reverse.f90:332-398) are local
// register operations and no mirrored data is ever re-read from HBM.  One
// block owns one group: the 4 columns (nz+1 doubles each, contiguous in
// memory) are staged in shared memory, all z transforms and pointwise operators
// run there, results are written back once.
//
// Spectral arrays are [kx][kyl][pz] with kyl the rank-local index of the
// *paired* ky order  ky' = 0, ny/2, 1, ny-1, 2, ny-2, ...  (kyl = 2a'+sy), so
// the y-mirror of a column is its neighbour kyl^1 and a slab of ky' is closed
// under mirroring (replaces mpi_reverse entirely).
//
// Work split inside a block (NT = nz/2 threads):
//  * pointwise stages are "z-owner": thread t owns rows z = t, t + NT (and
//    thread 0 also row nz) of ALL four slots, so every x/y mirror coupling is a
//    register operation and the hyperbolic functions of a row are evaluated once
//    and kept in registers for the whole kernel;
//  * DST-I / DCT-I of length nz follow the reference's own reduction to a real FFT of
//    length nz (stafft.f90:410-550: pre-process, forfft, post-process), two columns per
//    complex FFT of length nz (same block_cfft as the x/y passes).  The post-processing
//    recurrence x(2j+1) = x(2j-1) +- sqrt2*wk(.) (stafft.f90:466-471, 526-533) is a prefix
//    sum: done as a per-thread scan of 4 + a warp-shuffle scan, i.e. O(log n) roundings
//    instead of the reference's O(n).  Two 4-slot fields are transformed concurrently
//    (4 FFTs of nz/8 threads each = the whole block).
//
// The N-sized tables of the reference (phim, phip, thetam, thetap, dthetam,
// dthetap, green, filt: inversion_utils.f90:281-369,484-542) are never
// stored: they are recomputed per group from exp(-kl*zp), exp(-kl*zm).
#pragma once

#include "fft_core.cuh"

namespace ps3d {

struct SpecGeom {
    int nx, nyl, nz, pz;
    int has00;              // this rank owns the (kx,ky) = (0,0) column (at kx = 0, kyl = 0)
    const double* kxd;      // [nx/2+1]  x wavenumber used by diffx for slot pair b (0 at b = 0, nx/2)
    const double* kyd;      // [nyl/2]   y wavenumber used by diffy for local pair a' (0 for the (0, ny/2) pair)
    const double* k2l2;     // [nx/2+1][nyl]  k^2+l^2   (inversion_utils.f90:240-258)
    const double* k2l2i;    // [nx/2+1][nyl]  1/(k^2+l^2), 0 at (0,0)
    const double* zm;       // [nz+1] upper - z    (inversion_utils.f90:293-297)
    const double* zp;       // [nz+1] z - lower
    const double* rkz;      // [nz+1] pi kz / Lz, rkz[0] = 0   (sta3dfft.f90:106-107)
    const double* gamtop;   // [nz+1] (inversion_utils.f90:362)
    const double* gambot;   // [nz+1]
    double Lz, dzi, hdzi;
    const double2* tw;
    int ntw;
};

template <int NZ>
struct ZCfg {
    static constexpr int TPF = NZ / 8;            // threads per complex FFT of length NZ (two columns)
    static constexpr int NT = 4 * TPF;            // 4 FFTs = two 4-slot fields at a time  (= NZ/2)
    static constexpr int LC = NZ + 16;            // column buffer stride (rows 0..NZ, swizzled in 16-row blocks)
    static constexpr int BUF = 4 * LC;            // doubles per 4-slot field buffer
    static constexpr int NSIN = NZ / 2 + 2;       // sin(m pi/nz), m = 0..nz/2
    static constexpr int SCR = 4 * 2 * NZ + NSIN + 64 + 32;   // FFT scratch (4 x re/im x NZ) + sine table + warp totals + kept scalars
    static constexpr int ZI = 3;                  // rows per thread: t, t+NT, and NZ (thread 0 only)
};

// row z of a column buffer: XOR swizzle inside 16-row blocks, so that both the unit-stride (z-owner)
// and the stride-8 (transform output: thread u owns rows 8u..8u+7) accesses are bank-conflict free
__device__ __forceinline__ int cz(int z) { return z ^ ((z >> 4) & 15); }

// ---- group bookkeeping -----------------------------------------------------
struct Grp {
    int b, ap;
    bool dupx;          // b == 0 or b == nx/2: the x-mirror slot is the column itself (inactive)
    bool g00;           // this group holds the (0,0) column in slot 0
    double kx, ky;      // diff wavenumbers
    double k2[2], k2i[2];
    long long off[4];   // column offsets (doubles) of slots s = 2*sx + sy
    long long col[4];   // column index kx*nyl + kyl of the slots (per-column tables of the steppers)
};

__device__ __forceinline__ Grp make_grp(const SpecGeom& g, int gid) {
    Grp r;
    const int npair = g.nyl / 2;
    r.b = gid / npair;
    r.ap = gid - r.b * npair;
    r.dupx = (r.b == 0) || (2 * r.b == g.nx);
    r.g00 = g.has00 && r.b == 0 && r.ap == 0;
    r.kx = __ldg(&g.kxd[r.b]);
    r.ky = __ldg(&g.kyd[r.ap]);
    const int kxm = r.dupx ? r.b : g.nx - r.b;
#pragma unroll
    for (int sy = 0; sy < 2; ++sy) {
        const int kyl = 2 * r.ap + sy;
        r.k2[sy] = __ldg(&g.k2l2[(long long)r.b * g.nyl + kyl]);
        r.k2i[sy] = __ldg(&g.k2l2i[(long long)r.b * g.nyl + kyl]);
        r.col[sy] = (long long)r.b * g.nyl + kyl;
        r.col[2 + sy] = (long long)kxm * g.nyl + kyl;
        r.off[sy] = r.col[sy] * g.pz;
        r.off[2 + sy] = r.col[2 + sy] * g.pz;
    }
    return r;
}

// number of active slots: 2 when the x-mirror is the column itself
__device__ __forceinline__ int nslots(const Grp& r) { return r.dupx ? 2 : 4; }

// ---- z-owner helpers ----------------------------------------------------------
// row owned by this thread in iteration `it` (-1: none)
template <int NZ>
__device__ __forceinline__ int my_row(int it) {
    constexpr int NT = ZCfg<NZ>::NT;
    const int t = threadIdx.x;
    if (it < 2) return t + it * NT;
    return (t == 0) ? NZ : -1;
}

// Four-slot register tile of one row.
struct Row4 { double v[4]; };

template <int NZ>
__device__ __forceinline__ Row4 row_load_g(const double* __restrict__ src, const Grp& r, int z) {
    Row4 x;
#pragma unroll
    for (int s = 0; s < 4; ++s) x.v[s] = (s < 2 || !r.dupx) ? src[r.off[s] + z] : 0.0;
    return x;
}
template <int NZ>
__device__ __forceinline__ void row_store_g(double* __restrict__ dst, const Grp& r, int z, const Row4& x) {
#pragma unroll
    for (int s = 0; s < 4; ++s)
        if (s < 2 || !r.dupx) dst[r.off[s] + z] = x.v[s];
}
template <int NZ>
__device__ __forceinline__ Row4 row_load_s(const double* buf, int z) {
    constexpr int LC = ZCfg<NZ>::LC;
    Row4 x;
    const int zz = cz(z);
#pragma unroll
    for (int s = 0; s < 4; ++s) x.v[s] = buf[s * LC + zz];
    return x;
}
template <int NZ>
__device__ __forceinline__ void row_store_s(double* buf, int z, const Row4& x) {
    constexpr int LC = ZCfg<NZ>::LC;
    const int zz = cz(z);
#pragma unroll
    for (int s = 0; s < 4; ++s) buf[s * LC + zz] = x.v[s];
}

// d/dx, d/dy of a row (sta3dfft.f90:325-329): slot s = 2*sx + sy,
//   sx = 0 (kx = b):  ds = -kx f(nx-b);   sx = 1 (kx = nx-b): ds = +kx f(b);   same in y with sy.
__device__ __forceinline__ Row4 ddx(const Row4& f, const Grp& r) {
    Row4 d;
    d.v[0] = -r.kx * f.v[2]; d.v[1] = -r.kx * f.v[3]; d.v[2] = r.kx * f.v[0]; d.v[3] = r.kx * f.v[1];
    return d;
}
__device__ __forceinline__ Row4 ddy(const Row4& f, const Grp& r) {
    Row4 d;
    d.v[0] = -r.ky * f.v[1]; d.v[1] = r.ky * f.v[0]; d.v[2] = -r.ky * f.v[3]; d.v[3] = r.ky * f.v[2];
    return d;
}

// ---- harmonic (Laplace) functions on the fly -------------------------------
// block constants per sy (inversion_utils.f90:494-503, 529-530)
struct Hyp { double kl, ef, div, k2if, Q, R; bool lin; };

__device__ __forceinline__ Hyp make_hyp(const SpecGeom& g, const Grp& r, int sy) {
    Hyp h;
    h.lin = r.g00 && sy == 0;                 // (0,0): phim = zm/Lz, phip = zp/Lz, thetas = 0
    h.kl = sqrt(r.k2[sy]);
    h.ef = exp(-(h.kl * g.Lz));
    h.div = h.lin ? 0.0 : 1.0 / (1.0 - h.ef * h.ef);
    h.k2if = 0.5 * r.k2i[sy];
    h.Q = h.div * (1.0 + h.ef * h.ef);
    h.R = h.div * 2.0 * h.ef;
    return h;
}

// per-thread table for the rows this thread owns: phim, phip for sy = 0, 1 (inversion_utils.f90:505-519)
// (only phim/phip stay in registers for the whole kernel; exp(-kl zp), exp(-kl zm) are recovered where the
//  theta functions need them: ep = phim + ef phip, em = phip + ef phim)
template <int NZ>
struct HypRows {
    double phim[3][2], phip[3][2];
};

template <int NZ>
__device__ __forceinline__ void hyp_rows(HypRows<NZ>& T, const Hyp (&h)[2], const SpecGeom& g, const Grp& r) {
    const bool same = (r.k2[0] == r.k2[1]) && !r.g00;
#pragma unroll
    for (int it = 0; it < 3; ++it) {
        const int z = my_row<NZ>(it);
        if (z < 0) {
            // never read, but keep every register defined on every path: ptxas 12.9 was seen to emit spill
            // loads without the matching stores for values that are undefined on some lanes
            T.phim[it][0] = T.phim[it][1] = T.phip[it][0] = T.phip[it][1] = 0.0;
            continue;
        }
        const double zm = __ldg(&g.zm[z]), zp = __ldg(&g.zp[z]);
#pragma unroll
        for (int sy = 0; sy < 2; ++sy) {
            if (sy == 1 && same) {
                T.phim[it][1] = T.phim[it][0]; T.phip[it][1] = T.phip[it][0];
                continue;
            }
            if (h[sy].lin) {
                T.phim[it][sy] = zm / g.Lz;
                T.phip[it][sy] = zp / g.Lz;
            } else {
                const double ep = exp(-(h[sy].kl * zp));
                const double em = exp(-(h[sy].kl * zm));
                T.phim[it][sy] = h[sy].div * (ep - h[sy].ef * em);
                T.phip[it][sy] = h[sy].div * (em - h[sy].ef * ep);
            }
        }
    }
}

// thetam, thetap, dthetam, dthetap of one row (inversion_utils.f90:518-541)
__device__ __forceinline__ void hyp_theta(const Hyp& h, double zm, double zp, double phim,
                                          double phip, double& thm, double& thp, double& dthm, double& dthp) {
    if (h.lin) { thm = thp = dthm = dthp = 0.0; return; }
    const double ep = phim + h.ef * phip, em = phip + h.ef * phim;     // exp(-kl zp), exp(-kl zm)
    const double Lm = h.kl * zm;
    const double Lp = h.kl * zp;
    const double dphim = -h.kl * h.div * (ep + h.ef * em);
    const double dphip = h.kl * h.div * (em + h.ef * ep);
    thm = h.k2if * (h.R * Lm * phip - h.Q * Lp * phim);
    thp = h.k2if * (h.R * Lp * phim - h.Q * Lm * phip);
    dthm = -h.k2if * ((h.Q * Lp - 1.0) * dphim - h.R * Lm * dphip);
    dthp = -h.k2if * ((h.Q * Lm - 1.0) * dphip - h.R * Lp * dphim);
}

// ---- z transforms on 4-slot shared-memory fields ------------------------------
enum { XF_DST = 0, XF_DCT = 1 };

// shared scratch of the transforms
template <int NZ>
struct ZScr {
    double* fft;      // [4][2][NZ]
    double* sintab;   // [NZ/2 + 1]  sin(m pi / NZ)
    double* wt;       // [64] warp totals of the group scan / sum
    double* keep;     // [32] block-wide scalars parked between stages (keeps them out of registers)
};
template <int NZ>
__device__ __forceinline__ ZScr<NZ> make_scr(double* base) {
    ZScr<NZ> s;
    s.fft = base; s.sintab = base + 8 * NZ; s.wt = s.sintab + ZCfg<NZ>::NSIN; s.keep = s.wt + 64;
    return s;
}
// fill the sine table (once per block; followed by a barrier in the caller)
template <int NZ>
__device__ __forceinline__ void scr_init(const ZScr<NZ>& sc, const SpecGeom& g) {
    const int step = g.ntw / (2 * NZ);                      // tw[m] = exp(2 pi i m / ntw), ntw >= 2 NZ
    for (int m = threadIdx.x; m <= NZ / 2; m += blockDim.x) sc.sintab[m] = __ldg(&g.tw[m * step]).y;
}
template <int NZ>
__device__ __forceinline__ double sin_j(const ZScr<NZ>& sc, int j) { return sc.sintab[(j <= NZ / 2) ? j : NZ - j]; }
template <int NZ>
__device__ __forceinline__ double cos_j(const ZScr<NZ>& sc, int j) {
    return (j <= NZ / 2) ? sc.sintab[NZ / 2 - j] : -sc.sintab[j - NZ / 2];
}

// inclusive scan of (a, b) over the TPF consecutive threads of one FFT (u = index inside the FFT)
template <int TPF>
__device__ __forceinline__ void group_scan2(double& a, double& b, int u, double* wt) {
    constexpr int W = (TPF < 32) ? TPF : 32;
#pragma unroll
    for (int d = 1; d < W; d <<= 1) {
        const double ta = __shfl_up_sync(0xffffffffu, a, d, W);
        const double tb = __shfl_up_sync(0xffffffffu, b, d, W);
        if ((u & (W - 1)) >= d) { a += ta; b += tb; }
    }
    if (TPF > 32) {
        const int w = u >> 5;
        if ((u & 31) == 31) { wt[2 * w] = a; wt[2 * w + 1] = b; }
        __syncthreads();
        for (int v = 0; v < w; ++v) { a += wt[2 * v]; b += wt[2 * v + 1]; }
    }
}
// sum of (a, b) over the TPF threads of one FFT, result in every thread
template <int TPF>
__device__ __forceinline__ void group_sum2(double& a, double& b, int u, double* wt) {
    constexpr int W = (TPF < 32) ? TPF : 32;
#pragma unroll
    for (int d = W >> 1; d > 0; d >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, d, W);
        b += __shfl_xor_sync(0xffffffffu, b, d, W);
    }
    if (TPF > 32) {
        const int w = u >> 5;
        if ((u & 31) == 0) { wt[2 * w] = a; wt[2 * w + 1] = b; }
        __syncthreads();
        a = 0.0; b = 0.0;
        for (int v = 0; v < TPF / 32; ++v) { a += wt[2 * v]; b += wt[2 * v + 1]; }
    }
}

// DST-I (rows 1..NZ-1; rows 0 and NZ neither read nor written, stafft.f90:509-513) or DCT-I (rows 0..NZ) of the
// four slots of X0 and, concurrently, of X1 (may be null), scaled sqrt(2/NZ).  Reference algorithm
// (stafft.f90:410-550) with two columns per complex FFT:
//   DST: y_j = (x_j - x_{n-j})/2 + sin(j pi/n)(x_j + x_{n-j});  Y = FFT(y);
//        X_1 = Y_0/2, X_{2k} = -Im Y_k, X_{2k+1} = X_{2k-1} + Re Y_k            (all times sqrt(2/n))
//   DCT: y_0 = (x_0 + x_n)/2, y_j = (x_j + x_{n-j})/2 - sin(j pi/n)(x_j - x_{n-j});
//        X_1 = x_0/2 - x_n/2 + sum_j x_j cos(j pi/n), X_0 = Y_0, X_{2k} = Re Y_k, X_n = Y_{n/2},
//        X_{2k+1} = X_{2k-1} - Im Y_k
// The caller must have a barrier between its last write of X0/X1 and this call; ends with a barrier.
template <int NZ>
__device__ __forceinline__ void xform2(double* X0, int kind0, double* X1, int kind1, const ZScr<NZ>& sc,
                                       const SpecGeom& g) {
    constexpr int n = NZ, TPF = ZCfg<NZ>::TPF, LC = ZCfg<NZ>::LC;
    const int t = threadIdx.x;
    const int fg = t / (2 * TPF), f = (t / TPF) & 1, u = t - (t / TPF) * TPF, fft = 2 * fg + f;
    double* X = fg ? X1 : X0;
    const int kind = fg ? kind1 : kind0;
    const bool act = (X != nullptr);
    double* xa = (act ? X : X0) + (2 * f) * LC;
    double* xb = xa + LC;
    double* sre = sc.fft + fft * 2 * n;
    double* sim = sre + n;
    double* wt = sc.wt + fft * 8;
    const IxSwz ix;

    // ---- pre-process: two real sequences -> one complex sequence
    double vr[8], vi[8];
    double sa = 0.0, sb = 0.0;                 // DCT: X_1 partial sums
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int j = u + e * TPF;
        double yr = 0.0, yi = 0.0;
        if (act) {
            if (j == 0) {
                if (kind == XF_DCT) {
                    const double a0 = xa[cz(0)], an = xa[cz(n)], b0 = xb[cz(0)], bn = xb[cz(n)];
                    yr = 0.5 * (a0 + an); yi = 0.5 * (b0 + bn);
                    sa += 0.5 * (a0 - an); sb += 0.5 * (b0 - bn);
                }
            } else {
                const double aj = xa[cz(j)], an = xa[cz(n - j)], bj = xb[cz(j)], bn = xb[cz(n - j)];
                const double sn = sin_j<NZ>(sc, j);
                if (kind == XF_DST) {
                    yr = 0.5 * (aj - an) + sn * (aj + an);
                    yi = 0.5 * (bj - bn) + sn * (bj + bn);
                } else {
                    yr = 0.5 * (aj + an) - sn * (aj - an);
                    yi = 0.5 * (bj + bn) - sn * (bj - bn);
                    const double cs = cos_j<NZ>(sc, j);
                    sa += aj * cs; sb += bj * cs;
                }
            }
        }
        vr[e] = yr; vi[e] = yi;
    }
    group_sum2<TPF>(sa, sb, u, wt);            // (only meaningful for DCT; executed uniformly)

    block_cfft<n, false>(vr, vi, u, true, sre, sim, ix, g.tw, g.ntw / n);
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int idx = ix(u + e * TPF);
        sre[idx] = vr[e]; sim[idx] = vi[e];
    }
    __syncthreads();

    // ---- post-process: thread u owns k = 4u .. 4u+3, i.e. output rows 8u .. 8u+7
    double oa[4], ob[4], ea[4], eb[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int k = 4 * u + c;
        if (k == 0) {
            const double p = sre[ix(0)], q = sim[ix(0)];
            if (kind == XF_DST) { oa[c] = 0.5 * p; ob[c] = 0.5 * q; ea[c] = 0.0; eb[c] = 0.0; }
            else { oa[c] = sa; ob[c] = sb; ea[c] = p; eb[c] = q; }
        } else {
            const int ik = ix(k), im = ix(n - k);
            const double p = sre[ik], q = sim[ik], r = sre[im], s = sim[im];
            const double reA = 0.5 * (p + r), imA = 0.5 * (q - s), reB = 0.5 * (q + s), imB = 0.5 * (r - p);
            if (kind == XF_DST) { oa[c] = reA; ob[c] = reB; ea[c] = -imA; eb[c] = -imB; }
            else { oa[c] = -imA; ob[c] = -imB; ea[c] = reA; eb[c] = reB; }
        }
    }
#pragma unroll
    for (int c = 1; c < 4; ++c) { oa[c] += oa[c - 1]; ob[c] += ob[c - 1]; }
    double ta = oa[3], tb = ob[3];
    group_scan2<TPF>(ta, tb, u, wt + 4);
    const double pa = ta - oa[3], pb = tb - ob[3];      // exclusive prefix of this thread
    const double scl = sqrt(2.0 / (double)n);
    double nyq_a = 0.0, nyq_b = 0.0;
    if (u == 0 && kind == XF_DCT) { nyq_a = sre[ix(n / 2)]; nyq_b = sim[ix(n / 2)]; }
    if (act) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int k = 4 * u + c;
            xa[cz(2 * k + 1)] = scl * (oa[c] + pa);
            xb[cz(2 * k + 1)] = scl * (ob[c] + pb);
            if (k > 0 || kind == XF_DCT) { xa[cz(2 * k)] = scl * ea[c]; xb[cz(2 * k)] = scl * eb[c]; }
        }
        if (u == 0 && kind == XF_DCT) { xa[cz(n)] = scl * nyq_a; xb[cz(n)] = scl * nyq_b; }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------
// Operator-mode kernel: one z-operation on one field (drop-in for the public
// module procedures fftsine, fftcosine, diffx, diffy, central_diffz,
// field_combine_semi_spectral, field_decompose_semi_spectral).
// ---------------------------------------------------------------------------
enum { ZOP_SINE = 0, ZOP_COSINE, ZOP_COMBINE, ZOP_DECOMPOSE, ZOP_DIFFZ, ZOP_DIFFX, ZOP_DIFFY, ZOP_POISSON };

template <int NZ>
constexpr size_t zop_smem_bytes() { return (size_t)(ZCfg<NZ>::BUF + ZCfg<NZ>::SCR) * sizeof(double); }

template <int NZ>
__global__ void __launch_bounds__(ZCfg<NZ>::NT) k_zop(SpecGeom g, int op, const double* __restrict__ in,
                                                       double* __restrict__ out) {
    PS_SMEM(double, sm);
    constexpr int LC = ZCfg<NZ>::LC, BUF = ZCfg<NZ>::BUF;
    double* X = sm;
    const ZScr<NZ> scr = make_scr<NZ>(X + BUF);
    scr_init<NZ>(scr, g);
    const Grp r = make_grp(g, blockIdx.x);
    Hyp h[2];
    HypRows<NZ> T;
    if (op == ZOP_COMBINE || op == ZOP_DECOMPOSE) {
        h[0] = make_hyp(g, r, 0); h[1] = make_hyp(g, r, 1);
        hyp_rows<NZ>(T, h, g, r);
    }
    if (op == ZOP_DIFFX || op == ZOP_DIFFY) {
#pragma unroll
        for (int it = 0; it < 3; ++it) {
            const int z = my_row<NZ>(it);
            if (z < 0) continue;
            const Row4 x = row_load_g<NZ>(in, r, z);
            row_store_g<NZ>(out, r, z, op == ZOP_DIFFX ? ddx(x, r) : ddy(x, r));
        }
        return;
    }
    // stage the columns
#pragma unroll
    for (int it = 0; it < 3; ++it) {
        const int z = my_row<NZ>(it);
        if (z < 0) continue;
        Row4 x = row_load_g<NZ>(in, r, z);
        if (op == ZOP_DECOMPOSE && z >= 1 && z < NZ) {
            // subtract the harmonic part (inversion_utils.f90:571); boundary rows straight from memory
            const Row4 x0 = row_load_g<NZ>(in, r, 0), xn = row_load_g<NZ>(in, r, NZ);
#pragma unroll
            for (int s = 0; s < 4; ++s) x.v[s] -= x0.v[s] * T.phim[it][s & 1] + xn.v[s] * T.phip[it][s & 1];
        }
        row_store_s<NZ>(X, z, x);
    }
    __syncthreads();
    if (op == ZOP_DIFFZ) {
#pragma unroll
        for (int it = 0; it < 3; ++it) {
            const int z = my_row<NZ>(it);
            if (z < 0) continue;
            Row4 d;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const double* c = X + s * LC;
                d.v[s] = (z == 0) ? g.dzi * (c[cz(1)] - c[cz(0)]) : (z == NZ) ? g.dzi * (c[cz(NZ)] - c[cz(NZ - 1)]) : (c[cz(z + 1)] - c[cz(z - 1)]) * g.hdzi;
            }
            row_store_g<NZ>(out, r, z, d);
        }
        return;
    }
    xform2<NZ>(X, (op == ZOP_COSINE || op == ZOP_POISSON) ? XF_DCT : XF_DST, nullptr, XF_DST, scr, g);
    if (op == ZOP_POISSON) {
        // pressure Poisson solve between two cosine transforms (fields_derived.f90:125-148):
        // rs <- green * rs with green(kz) = -1/(k^2+l^2+rkz^2), green(0) = -1/(k^2+l^2) (inversion_utils.f90:283-288)
#pragma unroll
        for (int it = 0; it < 3; ++it) {
            const int z = my_row<NZ>(it);
            if (z < 0) continue;
            Row4 x = row_load_s<NZ>(X, z);
            const double rk = __ldg(&g.rkz[z]);
#pragma unroll
            for (int s = 0; s < 4; ++s) x.v[s] *= (z == 0) ? -r.k2i[s & 1] : -1.0 / (r.k2[s & 1] + rk * rk);
            row_store_s<NZ>(X, z, x);
        }
        __syncthreads();
        xform2<NZ>(X, XF_DCT, nullptr, XF_DST, scr, g);
    }
#pragma unroll
    for (int it = 0; it < 3; ++it) {
        const int z = my_row<NZ>(it);
        if (z < 0) continue;
        Row4 x = row_load_s<NZ>(X, z);
        if (op == ZOP_SINE && z == NZ) { x.v[0] = x.v[1] = x.v[2] = x.v[3] = 0.0; }     // stafft.f90:546-549
        if (op == ZOP_COMBINE && z >= 1 && z < NZ) {
#pragma unroll
            for (int s = 0; s < 4; ++s) x.v[s] += X[s * LC + cz(0)] * T.phim[it][s & 1] + X[s * LC + cz(NZ)] * T.phip[it][s & 1];
        }
        row_store_g<NZ>(out, r, z, x);
    }
}

// ---------------------------------------------------------------------------
// vor2vel, spectral part (reference inversion.f90:23-226 minus the six
// fftxys2p calls): svor -> svor (solenoidal), semi-spectral vorticity (input of
// the inverse x/y passes that give `vor`), svel.
// ---------------------------------------------------------------------------
template <int NZ>
constexpr size_t v2v_smem_bytes() { return (size_t)(4 * ZCfg<NZ>::BUF + ZCfg<NZ>::SCR) * sizeof(double); }

struct V2VArgs {
    double* svor0; double* svor1; double* svor2;         // in/out
    double* wsem0; double* wsem1; double* wsem2;         // semi-spectral vorticity (out)
    double* svel0; double* svel1; double* svel2;         // semi-spectral velocity (out)
};

// The solenoidal projection of inversion.f90:39-76 on one row (all four slots):
//   D = B_x - A_y;  A <- k2l2i (E_x + D_y),  B <- k2l2i (E_y - D_x);  the (0,0) column keeps its values.
// It only mixes slots with the same k^2 + l^2 and has no z dependence, so it commutes with the z transforms and
// with adding/removing the harmonic part: the kernel applies it once to the mixed-spectral rows (-> new svor)
// and once to the semi-spectral rows (-> input of the inverse x/y passes) instead of transforming the
// projected fields again.
__device__ __forceinline__ void project_row(Row4& fa, Row4& fb, const Row4& fe, const Grp& r) {
    const Row4 bx = ddx(fb, r), ay = ddy(fa, r);
    Row4 d;
#pragma unroll
    for (int s = 0; s < 4; ++s) d.v[s] = bx.v[s] - ay.v[s];
    const Row4 ex = ddx(fe, r), ey = ddy(fe, r), dx_ = ddx(d, r), dy_ = ddy(d, r);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        if (r.g00 && s == 0) continue;
        fa.v[s] = r.k2i[s & 1] * (ex.v[s] + dy_.v[s]);
        fb.v[s] = r.k2i[s & 1] * (ey.v[s] - dx_.v[s]);
    }
}

template <int NZ>
__global__ void __launch_bounds__(ZCfg<NZ>::NT, (NZ == 512) ? 2 : 1) k_vor2vel_spec(SpecGeom g, V2VArgs a) {
    PS_SMEM(double, sm);
    constexpr int LC = ZCfg<NZ>::LC, BUF = ZCfg<NZ>::BUF;
    double* A = sm;
    double* B = A + BUF;
    double* C = B + BUF;
    double* E = C + BUF;
    const ZScr<NZ> scr = make_scr<NZ>(E + BUF);
    scr_init<NZ>(scr, g);
    const Grp r = make_grp(g, blockIdx.x);
    Hyp h[2];
    h[0] = make_hyp(g, r, 0); h[1] = make_hyp(g, r, 1);
    HypRows<NZ> T;
    hyp_rows<NZ>(T, h, g, r);

    // stage svor
#pragma unroll
    for (int it = 0; it < 3; ++it) {
        const int z = my_row<NZ>(it);
        if (z < 0) continue;
        row_store_s<NZ>(A, z, row_load_g<NZ>(a.svor0, r, z));
        row_store_s<NZ>(B, z, row_load_g<NZ>(a.svor1, r, z));
        row_store_s<NZ>(C, z, row_load_g<NZ>(a.svor2, r, z));
    }
    __syncthreads();

    // C -> semi-spectral zeta (inversion.f90:45, :142-144): DST + harmonic part; this is also the
    // semi-spectral zeta that feeds the inverse x/y passes (:81).  A (xi) rides along: its sine sum is the
    // interior of combine(xi_old), used by the semi-spectral projection below.
    xform2<NZ>(C, XF_DST, A, XF_DST, scr, g);
#pragma unroll
    for (int it = 0; it < 3; ++it) {
        const int z = my_row<NZ>(it);
        if (z < 0) continue;
        Row4 c = row_load_s<NZ>(C, z);
        if (z >= 1 && z < NZ) {
#pragma unroll
            for (int s = 0; s < 4; ++s) c.v[s] += C[s * LC + cz(0)] * T.phim[it][s & 1] + C[s * LC + cz(NZ)] * T.phip[it][s & 1];
            row_store_s<NZ>(C, z, c);
        }
        row_store_g<NZ>(a.wsem2, r, z, c);
    }
    __syncthreads();
    // E = decompose(central_diffz(C)) (:46-47): FD, harmonic part removed, then DST (with eta in B)
    {
        double e0[4], en[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const double* c = C + s * LC;
            e0[s] = g.dzi * (c[cz(1)] - c[cz(0)]);
            en[s] = g.dzi * (c[cz(NZ)] - c[cz(NZ - 1)]);
        }
#pragma unroll
        for (int it = 0; it < 3; ++it) {
            const int z = my_row<NZ>(it);
            if (z < 0) continue;
            Row4 e;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const double* c = C + s * LC;
                if (z == 0) e.v[s] = e0[s];
                else if (z == NZ) e.v[s] = en[s];
                else e.v[s] = (c[cz(z + 1)] - c[cz(z - 1)]) * g.hdzi - (e0[s] * T.phim[it][s & 1] + en[s] * T.phip[it][s & 1]);
            }
            row_store_s<NZ>(E, z, e);
        }
    }
    __syncthreads();
    xform2<NZ>(E, XF_DST, B, XF_DST, scr, g);

    // Solenoidal projection (:39-76), twice (see project_row):
    //  * semi-spectral rows (combine(xi_old), combine(eta_old), dzeta/dz) -> wsem0, wsem1, the vorticity that the
    //    inverse x/y passes take to physical space (:80-82);
    //  * mixed-spectral rows (xi_old, eta_old re-read from memory, E) -> new svor, and the source of the w
    //    inversion D2 = A_y - B_x (:86-90) -> E buffer.
#pragma unroll
    for (int it = 0; it < 3; ++it) {
        const int z = my_row<NZ>(it);
        if (z < 0) continue;
        Row4 sa = row_load_s<NZ>(A, z), sb = row_load_s<NZ>(B, z), se;
        Row4 fa = row_load_g<NZ>(a.svor0, r, z), fb = row_load_g<NZ>(a.svor1, r, z);
        const Row4 fe = row_load_s<NZ>(E, z);
        if (z >= 1 && z < NZ) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const double* c = C + s * LC;
                sa.v[s] += A[s * LC + cz(0)] * T.phim[it][s & 1] + A[s * LC + cz(NZ)] * T.phip[it][s & 1];
                sb.v[s] += B[s * LC + cz(0)] * T.phim[it][s & 1] + B[s * LC + cz(NZ)] * T.phip[it][s & 1];
                se.v[s] = (c[cz(z + 1)] - c[cz(z - 1)]) * g.hdzi;
            }
        } else {
            se = fe;            // rows 0, NZ of E still hold the one-sided differences
        }
        project_row(sa, sb, se, r);
        row_store_g<NZ>(a.wsem0, r, z, sa);
        row_store_g<NZ>(a.wsem1, r, z, sb);
        project_row(fa, fb, fe, r);
        row_store_g<NZ>(a.svor0, r, z, fa);
        row_store_g<NZ>(a.svor1, r, z, fb);
        const Row4 ay2 = ddy(fa, r), bx2 = ddx(fb, r);
        Row4 d;
#pragma unroll
        for (int s = 0; s < 4; ++s) d.v[s] = ay2.v[s] - bx2.v[s];
        row_store_s<NZ>(E, z, d);
    }
    __syncthreads();
    // boundary values of D2 (:96-104), before anything is overwritten (parked in shared memory: they are
    // needed again only in the last stage)
    if (threadIdx.x < 4) {
        const int s = threadIdx.x;
        scr.keep[s] = E[s * LC + cz(0)];
        scr.keep[4 + s] = E[s * LC + cz(NZ)];
    }

    // invert Laplacian (:108-122): E <- green * D2 (rows 1..nz-1), A <- rkz * E (cosine series of dw/dz)
#pragma unroll
    for (int it = 0; it < 3; ++it) {
        const int z = my_row<NZ>(it);
        if (z < 0) continue;
        Row4 as, ds = row_load_s<NZ>(E, z);
        if (z >= 1 && z < NZ) {
            const double rk = __ldg(&g.rkz[z]);
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const double green = -1.0 / (r.k2[s & 1] + rk * rk);
                ds.v[s] = green * ds.v[s];
                as.v[s] = rk * ds.v[s];
            }
            row_store_s<NZ>(E, z, ds);
        } else {
            as.v[0] = as.v[1] = as.v[2] = as.v[3] = 0.0;
        }
        row_store_s<NZ>(A, z, as);
    }
    // horizontally averaged flow from the (0,0) column (:150-165): cosine transform in B slots 0, 1
    if (r.g00) {
#pragma unroll
        for (int it = 0; it < 3; ++it) {
            const int z = my_row<NZ>(it);
            if (z < 0) continue;
            // the (0,0) column of svor is not touched by the projection (:72-76): read it back from memory
            Row4 m;
            m.v[0] = m.v[1] = m.v[2] = m.v[3] = 0.0;
            if (z >= 1 && z < NZ) {                             // :153-154
                const double rkzi = 1.0 / __ldg(&g.rkz[z]);
                m.v[0] = -rkzi * a.svor1[r.off[0] + z];
                m.v[1] = rkzi * a.svor0[r.off[0] + z];
            }
            row_store_s<NZ>(B, z, m);
        }
    }
    __syncthreads();
    xform2<NZ>(A, XF_DCT, E, XF_DST, scr, g);     // (:128-129)
    if (r.g00) xform2<NZ>(B, XF_DCT, nullptr, XF_DST, scr, g);

    // w = E + boundary part, dw/dz = es + as (:96-104, :136-139);
    // u = k2l2i (es_x + cs_y), v = k2l2i (es_y - cs_x), (0,0) <- ubar, vbar (:169-213)
    double d0[4], dn[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) { d0[s] = scr.keep[s]; dn[s] = scr.keep[4 + s]; }
    double a00 = 0.0, a0n = 0.0, b00 = 0.0, b0n = 0.0;
    if (r.g00) {
        a00 = a.svor0[r.off[0]]; a0n = a.svor0[r.off[0] + NZ];
        b00 = a.svor1[r.off[0]]; b0n = a.svor1[r.off[0] + NZ];
    }
#pragma unroll
    for (int it = 0; it < 3; ++it) {
        const int z = my_row<NZ>(it);
        if (z < 0) continue;
        const Row4 as = row_load_s<NZ>(A, z), ds = row_load_s<NZ>(E, z), cs = row_load_s<NZ>(C, z);
        const double zm = __ldg(&g.zm[z]), zp = __ldg(&g.zp[z]);
        Row4 es, w;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const int sy = s & 1;
            double thm, thp, dthm, dthp;
            hyp_theta(h[sy], zm, zp, T.phim[it][sy], T.phip[it][sy], thm, thp, dthm, dthp);
            es.v[s] = d0[s] * dthm + dn[s] * dthp + as.v[s];
            w.v[s] = (z == 0 || z == NZ) ? 0.0 : ds.v[s] + d0[s] * thm + dn[s] * thp;
        }
        const Row4 ex = ddx(es, r), ey = ddy(es, r), cx = ddx(cs, r), cy = ddy(cs, r);
        Row4 u, v;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            u.v[s] = r.k2i[s & 1] * (ex.v[s] + cy.v[s]);
            v.v[s] = r.k2i[s & 1] * (ey.v[s] - cx.v[s]);
        }
        if (r.g00) {
            const double gt = __ldg(&g.gamtop[z]), gb = __ldg(&g.gambot[z]);
            u.v[0] = B[cz(z)] + b0n * gt - b00 * gb;               // ubar (:163)
            v.v[0] = B[LC + cz(z)] - a0n * gt + a00 * gb;          // vbar (:164)
        }
        row_store_g<NZ>(a.svel0, r, z, u);
        row_store_g<NZ>(a.svel1, r, z, v);
        row_store_g<NZ>(a.svel2, r, z, w);
    }
}

// ---------------------------------------------------------------------------
// vorticity tendency, spectral part (reference inversion.f90:298-371 after the
// fftxyp2s calls).  Inputs are the x/y-transformed fluxes r, q, p
// (semi-spectral); central_diffz commutes with the x/y FFT, so dq/dz and dp/dz
// are formed here instead of through two more 2-D FFTs.  The reference
// decomposes r, dq/dz, dp/dz, q, p separately (five sine transforms) and then
// combines them with diffx/diffy; those horizontal derivatives have no z
// dependence and only mix slots that share k^2 + l^2 (hence phim, phip), so
// they commute with the decomposition and the curl is taken first:
//   svorts = decompose( r_y - q_z,  p_z - r_x,  q_x - p_y )      (three sine transforms).
// ---------------------------------------------------------------------------
template <int NZ>
constexpr size_t src_smem_bytes() { return (size_t)(3 * ZCfg<NZ>::BUF + ZCfg<NZ>::SCR) * sizeof(double); }

struct SrcArgs {
    const double* r; const double* q; const double* p;   // semi-spectral fluxes
    double* s0; double* s1; double* s2;                  // svorts (mixed spectral)
};

// semi-spectral curl of one row from the rows z-1, z, z+1 of q, p and row z of r; dz = 1/(2 dz) in the
// interior, 1/dz with (lo, hi) = (z, z+1) or (z-1, z) at the boundaries (inversion_utils.f90:653-680)
__device__ __forceinline__ void curl_row(const Row4& fr, const Row4& fq, const Row4& fp, const Row4& qlo,
                                         const Row4& qhi, const Row4& plo, const Row4& phi, double dz,
                                         const Grp& r, Row4& s0, Row4& s1, Row4& s2) {
    const Row4 ry = ddy(fr, r), rx = ddx(fr, r), qx = ddx(fq, r), py = ddy(fp, r);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        s0.v[s] = ry.v[s] - (qhi.v[s] - qlo.v[s]) * dz;      // dr/dy - dq/dz (:341-347)
        s1.v[s] = (phi.v[s] - plo.v[s]) * dz - rx.v[s];      // dp/dz - dr/dx (:353-359)
        s2.v[s] = qx.v[s] - py.v[s];                         // dq/dx - dp/dy (:363-367)
    }
}

template <int NZ>
__global__ void __launch_bounds__(ZCfg<NZ>::NT, (NZ == 512) ? 2 : 1) k_source_spec(SpecGeom g, SrcArgs a) {
    PS_SMEM(double, sm);
    constexpr int BUF = ZCfg<NZ>::BUF;
    double* S0 = sm;
    double* S1 = S0 + BUF;
    double* S2 = S1 + BUF;
    const ZScr<NZ> scr = make_scr<NZ>(S2 + BUF);
    scr_init<NZ>(scr, g);
    const Grp r = make_grp(g, blockIdx.x);
    Hyp h[2];
    h[0] = make_hyp(g, r, 0); h[1] = make_hyp(g, r, 1);
    HypRows<NZ> T;
    hyp_rows<NZ>(T, h, g, r);

    // boundary rows of the curl (every thread: they define the harmonic part that is removed from its rows)
    Row4 b0[3], bn[3];
    {
        const Row4 r0 = row_load_g<NZ>(a.r, r, 0), q0 = row_load_g<NZ>(a.q, r, 0), p0 = row_load_g<NZ>(a.p, r, 0);
        const Row4 q1 = row_load_g<NZ>(a.q, r, 1), p1 = row_load_g<NZ>(a.p, r, 1);
        curl_row(r0, q0, p0, q0, q1, p0, p1, g.dzi, r, b0[0], b0[1], b0[2]);
        const Row4 rn = row_load_g<NZ>(a.r, r, NZ), qn = row_load_g<NZ>(a.q, r, NZ), pn = row_load_g<NZ>(a.p, r, NZ);
        const Row4 qm = row_load_g<NZ>(a.q, r, NZ - 1), pm = row_load_g<NZ>(a.p, r, NZ - 1);
        curl_row(rn, qn, pn, qm, qn, pm, pn, g.dzi, r, bn[0], bn[1], bn[2]);
    }
#pragma unroll
    for (int it = 0; it < 3; ++it) {
        const int z = my_row<NZ>(it);
        if (z < 0) continue;
        Row4 s[3];
        if (z == 0) { s[0] = b0[0]; s[1] = b0[1]; s[2] = b0[2]; }
        else if (z == NZ) { s[0] = bn[0]; s[1] = bn[1]; s[2] = bn[2]; }
        else {
            const Row4 fr = row_load_g<NZ>(a.r, r, z), fq = row_load_g<NZ>(a.q, r, z), fp = row_load_g<NZ>(a.p, r, z);
            const Row4 qlo = row_load_g<NZ>(a.q, r, z - 1), qhi = row_load_g<NZ>(a.q, r, z + 1);
            const Row4 plo = row_load_g<NZ>(a.p, r, z - 1), phi = row_load_g<NZ>(a.p, r, z + 1);
            curl_row(fr, fq, fp, qlo, qhi, plo, phi, g.hdzi, r, s[0], s[1], s[2]);
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    s[c].v[k] -= b0[c].v[k] * T.phim[it][k & 1] + bn[c].v[k] * T.phip[it][k & 1];
        }
        row_store_s<NZ>(S0, z, s[0]);
        row_store_s<NZ>(S1, z, s[1]);
        row_store_s<NZ>(S2, z, s[2]);
    }
    __syncthreads();
    xform2<NZ>(S0, XF_DST, S1, XF_DST, scr, g);
#pragma unroll
    for (int it = 0; it < 3; ++it) {
        const int z = my_row<NZ>(it);
        if (z < 0) continue;
        row_store_g<NZ>(a.s0, r, z, row_load_s<NZ>(S0, z));
        row_store_g<NZ>(a.s1, r, z, row_load_s<NZ>(S1, z));
    }
    xform2<NZ>(S2, XF_DST, nullptr, XF_DST, scr, g);
#pragma unroll
    for (int it = 0; it < 3; ++it) {
        const int z = my_row<NZ>(it);
        if (z < 0) continue;
        row_store_g<NZ>(a.s2, r, z, row_load_s<NZ>(S2, z));
    }
}

}  // namespace ps3d
