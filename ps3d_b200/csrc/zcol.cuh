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
// Two instantiations of every kernel (template parameter GEN):
//  * GEN = false, the fast path: pairs (a, ny-a) with a >= 1.  Both columns of a pair share k^2 + l^2, hence
//    one set of harmonic functions phi-/phi+/theta per block.
//  * GEN = true: the pair (0, ny/2) (different k^2 + l^2 per column, holds the (0,0) column with its linear
//    harmonic functions and the mean-flow work).  nx/2 + 1 blocks on the rank that owns ky = 0.
//
// Work split inside a block (NT = nz/2 threads):
//  * pointwise stages are "z-owner": thread t owns rows z = t, t + NT (and
//    thread 0 also row nz) of ALL four slots, so every x/y mirror coupling is a
//    register operation;
//  * DST-I / DCT-I of length nz follow the reference's own reduction to a real FFT of
//    length nz (stafft.f90:410-550: pre-process, forfft, post-process), two columns per
//    complex FFT of length nz (same block_cfft as the x/y passes), IN PLACE: the two column
//    buffers are the re/im exchange scratch of their own FFT.  The post-processing
//    recurrence x(2j+1) = x(2j-1) +- sqrt2*wk(.) (stafft.f90:466-471, 526-533) is a prefix
//    sum: done as a per-thread scan of 4 + a warp-shuffle scan, i.e. O(log n) roundings
//    instead of the reference's O(n).  Two 4-slot fields are transformed concurrently
//    (4 FFTs of nz/8 threads each = the whole block).
//
// The N-sized tables of the reference (phim, phip, thetam, thetap, dthetam,
// dthetap, green, filt: inversion_utils.f90:281-369,484-542) are never
// stored in HBM: phim/phip of the block's (kx, ky) are built once per block in shared memory from
// exp(-kl*zp), exp(-kl*zm), the thetas are derived from them where they are used.
//
// Shared memory at nz = 512: 4 column buffers of 4 x 513 doubles + 10.8 KB of tables = 74.7 KB, registers
// capped at 80: three blocks (24 warps) per SM.
#pragma once

#include "fft_core.cuh"

namespace ps3d {

struct SpecGeom {
    int nx, nyl, nz, pz;
    int has00;              // this rank owns the (kx,ky) = (0,0) column (at kx = 0, kyl = 0), i.e. the pair (0, ny/2)
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
    int ap0, npf;           // group mapping of a launch: block id -> b = id / npf, a' = ap0 + id % npf
};

template <int NZ>
struct ZCfg {
    static constexpr int TPF = NZ / 8;            // threads per complex FFT of length NZ (two columns)
    static constexpr int NT = 4 * TPF;            // 4 FFTs = two 4-slot fields at a time  (= NZ/2)
        // column buffer: rows 0..NZ/2 in a lower plane, rows NZ/2+1..NZ mirrored in an upper plane that starts at OFFU
    // (see cz); both XOR-swizzled in 16-row blocks (the swizzle moves row NZ/2 past NZ/2 unless NZ/32 is a multiple of 16)
    static constexpr int OFFU = (((NZ >> 5) & 15) != 0) ? NZ / 2 + 16 : NZ / 2 + 1;
    static constexpr int LC = OFFU + NZ / 2;
    static constexpr int BUF = 4 * LC;            // doubles per 4-slot field buffer
    static constexpr int NSIN = NZ / 2 + 2;       // sin(m pi/nz), m = 0..nz/2
    static constexpr int PST = NZ + 1;            // stride of one phim / phip plane
    static constexpr int ZI = 3;                  // rows per thread: t, t+NT, and NZ (thread 0 only)
    // tables behind the column buffers: sines, phim/phip planes (one per distinct k^2+l^2), warp totals (WT),
    // parked scalars (24), rows 0 / NZ of the columns whose transform input is handed over pre-processed (32)
    static constexpr int WT = (NT / 8 > 40) ? NT / 8 : 40;      // >= 4 per warp (put_pre's cosine sums), >= 32 (xform scans)
    static constexpr int aux(bool gen) { return NSIN + (gen ? 4 : 2) * PST + WT + 24 + 32; }
    // blocks per SM the launch bounds ask for: what shared memory allows, at no fewer than 80 registers
    static constexpr int ctas(int nbuf, bool gen) {
        const int bytes = (nbuf * BUF + aux(gen)) * 8 + 1024;
        int c = 233472 / bytes;
        if (c > 768 / NT) c = 768 / NT;
        if (c > 8) c = 8;
        return c < 1 ? 1 : c;
    }
};

// Row z of a column buffer.  Two planes: rows 0..NZ/2 at swz(z), rows NZ/2+1..NZ MIRRORED at OFFU + swz(NZ - z),
// swz = XOR swizzle inside 16-row blocks.  A thread owns the row pair (t, NZ - t) (my_row), i.e. the same position
// of the two planes: its accesses are unit-stride and 16-aligned in both planes, and the stride-8 accesses of the
// transform output (thread u owns rows 8u..8u+7) stay bank-conflict free in either plane through the swizzle.
__device__ __forceinline__ int swz16(int m) { return m ^ ((m >> 4) & 15); }
template <int NZ>
__device__ __forceinline__ int cz(int z) {
    return (2 * z <= NZ) ? swz16(z) : ZCfg<NZ>::OFFU + swz16(NZ - z);
}

// ---- group bookkeeping -----------------------------------------------------
template <bool GEN> __device__ __forceinline__ constexpr int sy_of(int s) { return GEN ? (s & 1) : 0; }

struct Grp {
    int b, ap;
    bool dupx;          // b == 0 or b == nx/2: the x-mirror slot is the column itself (inactive)
    bool g00;           // this group holds the (0,0) column in slot 0
    double kx, ky;      // diff wavenumbers
    double k2[2], k2i[2];   // per sy (fast path: [0] only)
    long long off[4];   // column offsets (doubles) of slots s = 2*sx + sy
};

template <bool GEN>
__device__ __forceinline__ Grp make_grp(const SpecGeom& g, int gid) {
    Grp r;
    r.b = gid / g.npf;
    r.ap = g.ap0 + (gid - r.b * g.npf);
    r.dupx = (r.b == 0) || (2 * r.b == g.nx);
    r.g00 = GEN && g.has00 && r.b == 0 && r.ap == 0;
    r.kx = __ldg(&g.kxd[r.b]);
    r.ky = __ldg(&g.kyd[r.ap]);
    const int kxm = r.dupx ? r.b : g.nx - r.b;
#pragma unroll
    for (int sy = 0; sy < 2; ++sy) {
        const int kyl = 2 * r.ap + sy;
        r.k2[sy] = __ldg(&g.k2l2[(long long)r.b * g.nyl + kyl]);
        r.k2i[sy] = __ldg(&g.k2l2i[(long long)r.b * g.nyl + kyl]);
        r.off[sy] = ((long long)r.b * g.nyl + kyl) * g.pz;
        r.off[2 + sy] = ((long long)kxm * g.nyl + kyl) * g.pz;
    }
    return r;
}

// ---- z-owner helpers ----------------------------------------------------------
// row owned by this thread in iteration `it` (-1: none)
template <int NZ>
__device__ __forceinline__ int my_row(int it) {
    constexpr int NT = ZCfg<NZ>::NT;
    const int t = threadIdx.x;
    if (it == 0) return t;                               // rows 0 .. NZ/2-1
    if (it == 1) return (t == 0) ? NZ / 2 : NZ - t;      // the mirror row (thread 0: the self-mirrored row NZ/2)
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
    const int zz = cz<NZ>(z);
#pragma unroll
    for (int s = 0; s < 4; ++s) x.v[s] = buf[s * LC + zz];
    return x;
}
template <int NZ>
__device__ __forceinline__ void row_store_s(double* buf, int z, const Row4& x) {
    constexpr int LC = ZCfg<NZ>::LC;
    const int zz = cz<NZ>(z);
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

// ---- shared tables behind the column buffers ---------------------------------
template <int NZ>
struct ZScr {
    double* sintab;   // [NZ/2 + 1]  sin(m pi / NZ)
    double* phim;     // [planes][PST]  harmonic functions of this block's (kx, ky): plane sy (fast path: one plane)
    double* phip;
    double* wt;       // [WT] warp totals of the group scan / sum
    double* keep;     // [24] block-wide scalars parked between stages (keeps them out of registers)
    double* park;     // [4 buffers][4 slots][2] rows 0 and NZ of a DST whose input was handed over pre-processed
};
template <int NZ, bool GEN>
__device__ __forceinline__ ZScr<NZ> make_scr(double* base) {
    typedef ZCfg<NZ> Z;
    ZScr<NZ> s;
    s.sintab = base;
    s.phim = s.sintab + Z::NSIN;
    s.phip = s.phim + (GEN ? 2 : 1) * Z::PST;
    s.wt = s.phip + (GEN ? 2 : 1) * Z::PST;
    s.keep = s.wt + Z::WT;
    s.park = s.keep + 24;
    return s;
}
// sine table (once per block; followed by a barrier in the caller)
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

// FFT twiddles of the z transforms from the sine table (fft_core.cuh, "Twiddle sources"): the butterflies of the
// twiddled passes need W^m with m = s * (b / s) < NZ/8, and (cos, sin)(2 pi m / NZ) = (sintab[NZ/2 - 2m], sintab[2m]).
// W^2m and W^4m by squaring (two and four roundings more, ~1e-16): no table look-up in global memory on the
// critical path of these latency-bound kernels.
template <int NZ>
struct TwSin {
    const double* st;
    __device__ __forceinline__ void get(int /*pass*/, int s, int p, double2& w1, double2& w2, double2& w4) const {
        const int m2 = 2 * s * p;
        w1 = make_double2(st[NZ / 2 - m2], st[m2]);
        w2 = make_double2(w1.x * w1.x - w1.y * w1.y, 2.0 * (w1.x * w1.y));
        w4 = make_double2(w2.x * w2.x - w2.y * w2.y, 2.0 * (w2.x * w2.y));
    }
};

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

// phim, phip of the rows this thread owns -> shared planes (inversion_utils.f90:505-519); the caller provides
// the barrier before other threads read them
template <int NZ, bool GEN>
__device__ __forceinline__ void phi_fill(const ZScr<NZ>& sc, const SpecGeom& g, const Grp& r) {
    constexpr int PST = ZCfg<NZ>::PST;
    const bool same = GEN && (r.k2[0] == r.k2[1]) && !r.g00;
#pragma unroll
    for (int sy = 0; sy < (GEN ? 2 : 1); ++sy) {
        const Hyp h = make_hyp(g, r, sy);
#pragma unroll
        for (int it = 0; it < 3; ++it) {
            const int z = my_row<NZ>(it);
            if (z < 0) continue;
            double pm, pp;
            if (sy == 1 && same) {
                pm = sc.phim[z]; pp = sc.phip[z];           // written by this very thread
            } else if (h.lin) {
                pm = __ldg(&g.zm[z]) / g.Lz;
                pp = __ldg(&g.zp[z]) / g.Lz;
            } else {
                const double ep = exp(-(h.kl * __ldg(&g.zp[z])));
                const double em = exp(-(h.kl * __ldg(&g.zm[z])));
                pm = h.div * (ep - h.ef * em);
                pp = h.div * (em - h.ef * ep);
            }
            sc.phim[sy * PST + z] = pm;
            sc.phip[sy * PST + z] = pp;
        }
    }
}
template <int NZ, bool GEN>
__device__ __forceinline__ double phim_of(const ZScr<NZ>& sc, int z, int s) { return sc.phim[sy_of<GEN>(s) * ZCfg<NZ>::PST + z]; }
template <int NZ, bool GEN>
__device__ __forceinline__ double phip_of(const ZScr<NZ>& sc, int z, int s) { return sc.phip[sy_of<GEN>(s) * ZCfg<NZ>::PST + z]; }

// thetam, thetap, dthetam, dthetap of one row (inversion_utils.f90:518-541); exp(-kl zp), exp(-kl zm) are
// recovered from phim, phip:  ep = phim + ef phip, em = phip + ef phim
struct Theta { double thm, thp, dthm, dthp; };
__device__ __forceinline__ Theta hyp_theta(const Hyp& h, double zm, double zp, double phim, double phip) {
    Theta t;
    if (h.lin) { t.thm = t.thp = t.dthm = t.dthp = 0.0; return t; }
    const double ep = phim + h.ef * phip, em = phip + h.ef * phim;
    const double Lm = h.kl * zm;
    const double Lp = h.kl * zp;
    const double dphim = -h.kl * h.div * (ep + h.ef * em);
    const double dphip = h.kl * h.div * (em + h.ef * ep);
    t.thm = h.k2if * (h.R * Lm * phip - h.Q * Lp * phim);
    t.thp = h.k2if * (h.R * Lp * phim - h.Q * Lm * phip);
    t.dthm = -h.k2if * ((h.Q * Lp - 1.0) * dphim - h.R * Lm * dphip);
    t.dthp = -h.k2if * ((h.Q * Lm - 1.0) * dphip - h.R * Lp * dphim);
    return t;
}

// ---- z transforms on 4-slot shared-memory fields ------------------------------
enum { XF_DST = 0, XF_DCT = 1 };

// Barrier over the TPF threads of ONE FFT (its exchanges, the regrouping of its spectrum and its scan touch only its
// own two column buffers and its own slice of wt): a named barrier when the group is whole warps, else the block's.
// PS3D_ZBAR_BLOCK (build switch) falls back to block-wide barriers everywhere.  Every thread of the block makes the
// same sequence of calls (inactive groups included), so the test-only emulator can map it onto __syncthreads.
template <int TPF>
struct ZBar {
    int id;         // 1 + index of the FFT inside the block
    __device__ __forceinline__ void operator()() const {
#if defined(PS3D_EMU) || defined(PS3D_ZBAR_BLOCK)
        __syncthreads();
#else
        if (TPF >= 32 && TPF % 32 == 0) asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(TPF) : "memory");
        else __syncthreads();
#endif
    }
};

// inclusive scan of (a, b) over the TPF consecutive threads of one FFT (u = index inside the FFT)
template <int TPF, class BAR>
__device__ __forceinline__ void group_scan2(double& a, double& b, int u, double* wt, const BAR& bar) {
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
        bar();
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

// DST-I (rows 1..NZ-1; rows 0 and NZ keep their values, stafft.f90:509-513) or DCT-I (rows 0..NZ) of the
// four slots of X0 and, concurrently, of X1 (may be null), scaled sqrt(2/NZ).  Reference algorithm
// (stafft.f90:410-550) with two columns per complex FFT:
//   DST: y_j = (x_j - x_{n-j})/2 + sin(j pi/n)(x_j + x_{n-j});  Y = FFT(y);
//        X_1 = Y_0/2, X_{2k} = -Im Y_k, X_{2k+1} = X_{2k-1} + Re Y_k            (all times sqrt(2/n))
//   DCT: y_0 = (x_0 + x_n)/2, y_j = (x_j + x_{n-j})/2 - sin(j pi/n)(x_j - x_{n-j});
//        X_1 = x_0/2 - x_n/2 + sum_j x_j cos(j pi/n), X_0 = Y_0, X_{2k} = Re Y_k, X_n = Y_{n/2},
//        X_{2k+1} = X_{2k-1} - Im Y_k
// In place: once the pre-processed sequence is in registers, the two column buffers of an FFT serve as the
// re/im planes of its Stockham exchanges (indices 0..n-1); rows 0 and n of a DST are carried across in registers.
// The caller must have a barrier between its last write of X0/X1 and this call; ends with a barrier.
// Second half of a transform: complex FFT of the pre-processed sequence held in (vr, vi), Hermitian split into the
// two columns, post-processing recurrence, rows written back in the column layout (cz).  (sa, sb): DCT: X_1 of the
// two columns; DST, u == 0: row 0.  (na, nb): DST, u == 0: row n.  Entered with every read of the buffers done
// (they become the exchange scratch); ends with a barrier.
template <int NZ>
__device__ __forceinline__ void xform_tail(double (&vr)[8], double (&vi)[8], double sa, double sb, double na, double nb,
                                           double* xa, double* xb, int kind, bool act, int u, double* wt, const ZScr<NZ>& sc) {
    constexpr int n = NZ, TPF = ZCfg<NZ>::TPF;
    const IxSwz ix;
    const ZBar<TPF> bar{1 + (int)(threadIdx.x / TPF)};
    block_cfft<n, false>(vr, vi, u, act, xa, xb, ix, TwSin<NZ>{sc.sintab}, bar);
    bar();
    if (act) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int idx = ix(u + e * TPF);
            xa[idx] = vr[e]; xb[idx] = vi[e];
        }
    }
    bar();

    // ---- post-process: thread u owns k = 4u .. 4u+3, i.e. output rows 8u .. 8u+7
    double oa[4], ob[4], ea[4], eb[4];
    double nyq_a = 0.0, nyq_b = 0.0;
    if (act) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int k = 4 * u + c;
            if (k == 0) {
                const double p = xa[ix(0)], q = xb[ix(0)];
                if (kind == XF_DST) { oa[c] = 0.5 * p; ob[c] = 0.5 * q; ea[c] = sa; eb[c] = sb; }
                else { oa[c] = sa; ob[c] = sb; ea[c] = p; eb[c] = q; }
            } else {
                const int ik = ix(k), im = ix(n - k);
                const double p = xa[ik], q = xb[ik], r = xa[im], s = xb[im];
                const double reA = 0.5 * (p + r), imA = 0.5 * (q - s), reB = 0.5 * (q + s), imB = 0.5 * (r - p);
                if (kind == XF_DST) { oa[c] = reA; ob[c] = reB; ea[c] = -imA; eb[c] = -imB; }
                else { oa[c] = -imA; ob[c] = -imB; ea[c] = reA; eb[c] = reB; }
            }
        }
        if (u == 0 && kind == XF_DCT) { nyq_a = xa[ix(n / 2)]; nyq_b = xb[ix(n / 2)]; }
    } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) { oa[c] = ob[c] = ea[c] = eb[c] = 0.0; }
    }
#pragma unroll
    for (int c = 1; c < 4; ++c) { oa[c] += oa[c - 1]; ob[c] += ob[c - 1]; }
    double ta = oa[3], tb = ob[3];
    group_scan2<TPF>(ta, tb, u, wt + 4, bar);
    if (TPF <= 32) bar();                      // every read of the spectrum is done before the rows are written
    const double pa = ta - oa[3], pb = tb - ob[3];      // exclusive prefix of this thread
    const double scl = sqrt(2.0 / (double)n);
    if (act) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int k = 4 * u + c;
            xa[cz<NZ>(2 * k + 1)] = scl * (oa[c] + pa);
            xb[cz<NZ>(2 * k + 1)] = scl * (ob[c] + pb);
            // row 0 of a DST: the value it had on entry (unscaled)
            const double se = (k == 0 && kind == XF_DST) ? 1.0 : scl;
            xa[cz<NZ>(2 * k)] = se * ea[c]; xb[cz<NZ>(2 * k)] = se * eb[c];
        }
        if (u == 0) {
            if (kind == XF_DCT) { xa[cz<NZ>(n)] = scl * nyq_a; xb[cz<NZ>(n)] = scl * nyq_b; }
            else { xa[cz<NZ>(n)] = na; xb[cz<NZ>(n)] = nb; }
        }
    }
    __syncthreads();
}

template <int NZ>
__device__ __forceinline__ void xform2(double* X0, int kind0, double* X1, int kind1, const ZScr<NZ>& sc) {
    constexpr int n = NZ, TPF = ZCfg<NZ>::TPF, LC = ZCfg<NZ>::LC;
    const int t = threadIdx.x;
    const int fg = t / (2 * TPF), f = (t / TPF) & 1, u = t - (t / TPF) * TPF, fft = 2 * fg + f;
    double* X = fg ? X1 : X0;
    const int kind = fg ? kind1 : kind0;
    const bool act = (X != nullptr);
    double* xa = (act ? X : X0) + (2 * f) * LC;     // (inactive threads never dereference these)
    double* xb = xa + LC;
    double* wt = sc.wt + fft * 8;

    // ---- pre-process: two real sequences -> one complex sequence
    double vr[8], vi[8];
    double sa = 0.0, sb = 0.0;                 // DCT: X_1 partial sums;  DST, u == 0: row 0 carried across
    double na = 0.0, nb = 0.0;                 // DST, u == 0: row n carried across (it lies inside the scratch range)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int j = u + e * TPF;
        double yr = 0.0, yi = 0.0;
        if (act) {
            if (j == 0) {
                const double a0 = xa[cz<NZ>(0)], b0 = xb[cz<NZ>(0)];
                if (kind == XF_DCT) {
                    const double an = xa[cz<NZ>(n)], bn = xb[cz<NZ>(n)];
                    yr = 0.5 * (a0 + an); yi = 0.5 * (b0 + bn);
                    sa += 0.5 * (a0 - an); sb += 0.5 * (b0 - bn);
                } else {
                    sa = a0; sb = b0;
                    na = xa[cz<NZ>(n)]; nb = xb[cz<NZ>(n)];
                }
            } else {
                const double aj = xa[cz<NZ>(j)], an = xa[cz<NZ>(n - j)], bj = xb[cz<NZ>(j)], bn = xb[cz<NZ>(n - j)];
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
    {
        // DCT: block-wide sums for X_1.  DST: (sa, sb) of thread u == 0 hold row 0 and must survive as they
        // are, so the sum runs on a copy and is discarded.
        double ta = (kind == XF_DCT) ? sa : 0.0, tb = (kind == XF_DCT) ? sb : 0.0;
        group_sum2<TPF>(ta, tb, u, wt);        // executed uniformly (barrier inside when TPF > 32)
        if (kind == XF_DCT) { sa = ta; sb = tb; }
    }
    if (TPF <= 32) __syncthreads();            // every pre-process read is done before the buffers become scratch
    xform_tail<NZ>(vr, vi, sa, sb, na, nb, xa, xb, kind, act, u, wt, sc);
}

// ---- transforms whose input is handed over pre-processed (the hot kernels) ------------------------------------
// The stage that PRODUCES the input of a transform owns the row pair (t, n - t) of all four slots (my_row), which is
// exactly what the pre-processing step needs: it writes y_t, y_{n-t} instead of x_t, x_{n-t} (put_pre), at the plain
// positions t, n - t of the column buffers (unit stride for the producer and for the first FFT pass alike: no bank
// conflicts at any alignment).  Against pre-processing inside the transform this saves, per element, two mirrored
// shared-memory reads (the conflict-laden ones) and the sine look-up.  Rows 0 and n of a DST (kept by the transform)
// are parked in `park`, the cosine sums of a DCT (X_1) are reduced over the block in put_pre_sums.
//
// lo = row t, hi = row n - t.  Thread 0 (rows 0, n/2, n) passes hi = row n/2 and hands rows 0 and n over separately,
// where it computes them (park_rows: the values a DST carries through; a DCT of this file has zero boundary rows),
// so that no thread holds them in registers.  sn, cs = sin, cos(t pi / n) of the calling thread.  ps: running
// cosine sums of a DCT.
template <int NZ>
__device__ __forceinline__ void put_pre(double* X, int kind, double sn, double cs, const Row4& lo, const Row4& hi,
                                        double (&ps)[4]) {
    constexpr int LC = ZCfg<NZ>::LC;
    const int t = threadIdx.x;
    if (t != 0) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const double a = lo.v[s], b = hi.v[s];
            if (kind == XF_DST) {
                const double d = 0.5 * (a - b), e = sn * (a + b);
                X[s * LC + t] = d + e; X[s * LC + NZ - t] = e - d;
            } else {
                const double d = 0.5 * (a + b), e = sn * (a - b);
                X[s * LC + t] = d - e; X[s * LC + NZ - t] = d + e;
                ps[s] += cs * (a - b);                       // cos((n - t) pi / n) = -cos(t pi / n)
            }
        }
    } else {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            // y_0 = 0 (DST) / (x_0 + x_n)/2 = 0 (DCT with zero boundary rows);  y_{n/2} = 2 x_{n/2} (DST) / x_{n/2} (DCT)
            X[s * LC] = 0.0;
            X[s * LC + NZ / 2] = (kind == XF_DST) ? 2.0 * hi.v[s] : hi.v[s];
        }
    }
}
// row 0 (top = 0) or row n (top = 1) of the four columns of a DST input: parked, restored by the transform
__device__ __forceinline__ void park_row(double* park, int top, const Row4& x) {
#pragma unroll
    for (int s = 0; s < 4; ++s) park[2 * s + top] = x.v[s];
}
// warp totals of the cosine sums -> wt[warp][slot]; xform2p adds them up after the caller's barrier
template <int NZ>
__device__ __forceinline__ void put_pre_sums(const double (&ps)[4], const ZScr<NZ>& sc) {
    constexpr int NT = ZCfg<NZ>::NT, W = (NT < 32) ? NT : 32;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        double v = ps[s];
#pragma unroll
        for (int d = W >> 1; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d, W);
        if ((threadIdx.x & (W - 1)) == 0) sc.wt[(threadIdx.x / W) * 4 + s] = v;
    }
}

// xform2 for pre-processed input (see put_pre).  pk0 / pk1: the parked rows of X0 / X1 (DST); a DCT takes its X_1
// from the warp totals in sc.wt.  The caller has a barrier between put_pre / put_pre_sums and this call.
// Inlined at every call site.  (Measured: as a real function call -- to shrink the 105 KB body of vor2vel, whose
// three resident blocks per SM run different phases: 13 % "no instruction" stalls in ncu r02i -- the kernels get
// SLOWER, vor2vel 4.89 -> 5.70 ms, source 2.77 -> 3.14 ms, same-box A/B profiles/r02o_ab_noinline.log: the values
// that live across the call are spilled around it.)
#define PS_XFORM_CALL __device__ __forceinline__
template <int NZ>
PS_XFORM_CALL void xform2p_(double* X0, int kind0, const double* pk0, double* X1, int kind1, const double* pk1,
                            double* sintab, double* wtab) {
    ZScr<NZ> sc;
    sc.sintab = sintab; sc.wt = wtab; sc.phim = sc.phip = sc.keep = sc.park = nullptr;
    constexpr int TPF = ZCfg<NZ>::TPF, LC = ZCfg<NZ>::LC, NT = ZCfg<NZ>::NT, NW = (NT < 32) ? 1 : NT / 32;
    const int t = threadIdx.x;
    const int fg = t / (2 * TPF), f = (t / TPF) & 1, u = t - (t / TPF) * TPF, fft = 2 * fg + f;
    double* X = fg ? X1 : X0;
    const double* pk = fg ? pk1 : pk0;
    const int kind = fg ? kind1 : kind0;
    const bool act = (X != nullptr);
    double* xa = (act ? X : X0) + (2 * f) * LC;
    double* xb = xa + LC;
    double vr[8], vi[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        vr[e] = act ? xa[u + e * TPF] : 0.0;
        vi[e] = act ? xb[u + e * TPF] : 0.0;
    }
    double sa = 0.0, sb = 0.0, na = 0.0, nb = 0.0;
    if (act && u == 0) {
        if (kind == XF_DST) {
            sa = pk[4 * f]; na = pk[4 * f + 1]; sb = pk[4 * f + 2]; nb = pk[4 * f + 3];
        } else {
            for (int w = 0; w < NW; ++w) { sa += sc.wt[4 * w + 2 * f]; sb += sc.wt[4 * w + 2 * f + 1]; }
        }
    }
    __syncthreads();                           // every read of the input (and of the warp totals) is done
    xform_tail<NZ>(vr, vi, sa, sb, na, nb, xa, xb, kind, act, u, sc.wt + fft * 8, sc);
}

template <int NZ>
__device__ __forceinline__ void xform2p(double* X0, int kind0, const double* pk0, double* X1, int kind1, const double* pk1,
                                        const ZScr<NZ>& sc) {
    xform2p_<NZ>(X0, kind0, pk0, X1, kind1, pk1, sc.sintab, sc.wt);
}

// ---------------------------------------------------------------------------
// Operator-mode kernel: one z-operation on one field (drop-in for the public
// module procedures fftsine, fftcosine, diffx, diffy, central_diffz,
// field_combine_semi_spectral, field_decompose_semi_spectral).  General
// instantiation for every group (this is the host-boundary path, not the hot loop).
// ---------------------------------------------------------------------------
enum { ZOP_SINE = 0, ZOP_COSINE, ZOP_COMBINE, ZOP_DECOMPOSE, ZOP_DIFFZ, ZOP_DIFFX, ZOP_DIFFY, ZOP_POISSON,
       ZOP_DIFFZ_SPEC };     // spectral d/dz of a mixed-spectral field (diffz, inversion_utils.f90:683-719; buoyancy build)

// dphim, dphip of one row (inversion_utils.f90:520-521; (0,0): -1/Lz, +1/Lz, :333-336) from phim, phip
__device__ __forceinline__ void hyp_dphi(const Hyp& h, double Lz, double phim, double phip, double& dpm, double& dpp) {
    if (h.lin) { dpm = -1.0 / Lz; dpp = 1.0 / Lz; return; }
    const double ep = phim + h.ef * phip, em = phip + h.ef * phim;
    dpm = -h.kl * h.div * (ep + h.ef * em);
    dpp = h.kl * h.div * (em + h.ef * ep);
}

template <int NZ>
constexpr size_t zop_smem_bytes() { return (size_t)(ZCfg<NZ>::BUF + ZCfg<NZ>::aux(true)) * sizeof(double); }

template <int NZ>
__global__ void __launch_bounds__(ZCfg<NZ>::NT) k_zop(SpecGeom g, int op, const double* __restrict__ in,
                                                       double* __restrict__ out) {
    PS_SMEM(double, sm);
    constexpr int LC = ZCfg<NZ>::LC, BUF = ZCfg<NZ>::BUF;
    constexpr bool GEN = true;
    double* X = sm;
    const ZScr<NZ> scr = make_scr<NZ, GEN>(X + BUF);
    scr_init<NZ>(scr, g);
    const Grp r = make_grp<GEN>(g, blockIdx.x);
    if (op == ZOP_COMBINE || op == ZOP_DECOMPOSE || op == ZOP_DIFFZ_SPEC) phi_fill<NZ, GEN>(scr, g, r);
    if (op == ZOP_DIFFZ_SPEC) {
        // diffz (inversion_utils.f90:683-719): ds = fs(0) dphim + fs(nz) dphip + cosine(rkz * fs), then decompose
        double f0[4], fn[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            f0[s] = (s < 2 || !r.dupx) ? in[r.off[s]] : 0.0;
            fn[s] = (s < 2 || !r.dupx) ? in[r.off[s] + NZ] : 0.0;
        }
#pragma unroll
        for (int it = 0; it < 3; ++it) {
            const int z = my_row<NZ>(it);
            if (z < 0) continue;
            Row4 x = row_load_g<NZ>(in, r, z);
            const double rk = (z >= 1 && z < NZ) ? __ldg(&g.rkz[z]) : 0.0;          // as(0) = as(nz) = 0
#pragma unroll
            for (int s = 0; s < 4; ++s) x.v[s] *= rk;
            row_store_s<NZ>(X, z, x);
        }
        __syncthreads();
        xform2<NZ>(X, XF_DCT, nullptr, XF_DST, scr);
        Hyp h[2];
        h[0] = make_hyp(g, r, 0); h[1] = make_hyp(g, r, 1);
#pragma unroll
        for (int it = 0; it < 3; ++it) {
            const int z = my_row<NZ>(it);
            if (z < 0) continue;
            Row4 x = row_load_s<NZ>(X, z);
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                double dpm, dpp;
                hyp_dphi(h[s & 1], g.Lz, phim_of<NZ, GEN>(scr, z, s), phip_of<NZ, GEN>(scr, z, s), dpm, dpp);
                x.v[s] += f0[s] * dpm + fn[s] * dpp;
            }
            row_store_s<NZ>(X, z, x);
        }
        __syncthreads();
        // field_decompose_semi_spectral (:563-592): remove the harmonic part defined by the boundary rows, sine transform
        double d0[4], dn[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) { d0[s] = X[s * LC + cz<NZ>(0)]; dn[s] = X[s * LC + cz<NZ>(NZ)]; }
#pragma unroll
        for (int it = 0; it < 3; ++it) {
            const int z = my_row<NZ>(it);
            if (z < 1 || z >= NZ) continue;
            Row4 x = row_load_s<NZ>(X, z);
#pragma unroll
            for (int s = 0; s < 4; ++s) x.v[s] -= d0[s] * phim_of<NZ, GEN>(scr, z, s) + dn[s] * phip_of<NZ, GEN>(scr, z, s);
            row_store_s<NZ>(X, z, x);
        }
        __syncthreads();
        xform2<NZ>(X, XF_DST, nullptr, XF_DST, scr);
#pragma unroll
        for (int it = 0; it < 3; ++it) {
            const int z = my_row<NZ>(it);
            if (z < 0) continue;
            row_store_g<NZ>(out, r, z, row_load_s<NZ>(X, z));
        }
        return;
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
            // (phim/phip of this thread's own rows: written by itself above)
            const Row4 x0 = row_load_g<NZ>(in, r, 0), xn = row_load_g<NZ>(in, r, NZ);
#pragma unroll
            for (int s = 0; s < 4; ++s) x.v[s] -= x0.v[s] * phim_of<NZ, GEN>(scr, z, s) + xn.v[s] * phip_of<NZ, GEN>(scr, z, s);
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
                d.v[s] = (z == 0) ? g.dzi * (c[cz<NZ>(1)] - c[cz<NZ>(0)]) : (z == NZ) ? g.dzi * (c[cz<NZ>(NZ)] - c[cz<NZ>(NZ - 1)]) : (c[cz<NZ>(z + 1)] - c[cz<NZ>(z - 1)]) * g.hdzi;
            }
            row_store_g<NZ>(out, r, z, d);
        }
        return;
    }
    xform2<NZ>(X, (op == ZOP_COSINE || op == ZOP_POISSON) ? XF_DCT : XF_DST, nullptr, XF_DST, scr);
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
        xform2<NZ>(X, XF_DCT, nullptr, XF_DST, scr);
    }
#pragma unroll
    for (int it = 0; it < 3; ++it) {
        const int z = my_row<NZ>(it);
        if (z < 0) continue;
        Row4 x = row_load_s<NZ>(X, z);
        if (op == ZOP_SINE && z == NZ) { x.v[0] = x.v[1] = x.v[2] = x.v[3] = 0.0; }     // stafft.f90:546-549
        if (op == ZOP_COMBINE && z >= 1 && z < NZ) {
#pragma unroll
            for (int s = 0; s < 4; ++s)
                x.v[s] += X[s * LC + cz<NZ>(0)] * phim_of<NZ, GEN>(scr, z, s) + X[s * LC + cz<NZ>(NZ)] * phip_of<NZ, GEN>(scr, z, s);
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
constexpr size_t v2v_smem_bytes(bool gen) { return (size_t)(4 * ZCfg<NZ>::BUF + ZCfg<NZ>::aux(gen)) * sizeof(double); }

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
template <bool GEN>
__device__ __forceinline__ void project_row(Row4& fa, Row4& fb, const Row4& fe, const Grp& r) {
    const Row4 bx = ddx(fb, r), ay = ddy(fa, r);
    Row4 d;
#pragma unroll
    for (int s = 0; s < 4; ++s) d.v[s] = bx.v[s] - ay.v[s];
    const Row4 ex = ddx(fe, r), ey = ddy(fe, r), dx_ = ddx(d, r), dy_ = ddy(d, r);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        if (GEN && r.g00 && s == 0) continue;
        fa.v[s] = r.k2i[sy_of<GEN>(s)] * (ex.v[s] + dy_.v[s]);
        fb.v[s] = r.k2i[sy_of<GEN>(s)] * (ey.v[s] - dx_.v[s]);
    }
}

// sin, cos(t pi / NZ) of the calling thread (t < NZ/2) from the twiddle table: the two constants of put_pre
template <int NZ>
__device__ __forceinline__ void my_sincos(const SpecGeom& g, double& sn, double& cs) {
    const int step = g.ntw / (2 * NZ), t = threadIdx.x;
    sn = __ldg(&g.tw[t * step]).y;
    cs = __ldg(&g.tw[(NZ / 2 - t) * step]).y;
}

template <int NZ, bool GEN>
__global__ void __launch_bounds__(ZCfg<NZ>::NT, ZCfg<NZ>::ctas(4, GEN)) k_vor2vel_spec(SpecGeom g, V2VArgs a) {
    PS_SMEM(double, sm);
    constexpr int LC = ZCfg<NZ>::LC, BUF = ZCfg<NZ>::BUF;
    double* A = sm;
    double* B = A + BUF;
    double* C = B + BUF;
    double* E = C + BUF;
    const ZScr<NZ> scr = make_scr<NZ, GEN>(E + BUF);
    double* pkA = scr.park;
    double* pkB = pkA + 8;
    double* pkC = pkB + 8;
    double* pkE = pkC + 8;
    scr_init<NZ>(scr, g);
    const Grp r = make_grp<GEN>(g, blockIdx.x);
    phi_fill<NZ, GEN>(scr, g, r);
    double sn, cs;
    my_sincos<NZ>(g, sn, cs);
    const int t = threadIdx.x;
    const int z0 = my_row<NZ>(0), z1 = my_row<NZ>(1);          // the row pair of this thread (thread 0: rows 0, NZ/2, NZ)
    double nops[4] = {0.0, 0.0, 0.0, 0.0};
    Row4 zero4;
    zero4.v[0] = zero4.v[1] = zero4.v[2] = zero4.v[3] = 0.0;

    // svor from memory straight into the pre-processed input of the three sine transforms (thread 0: rows 0 and NZ,
    // which the transforms carry through, go to the park)
    {
        const Row4 cl = row_load_g<NZ>(a.svor2, r, z0), ch = row_load_g<NZ>(a.svor2, r, z1);
        const Row4 al = row_load_g<NZ>(a.svor0, r, z0), ah = row_load_g<NZ>(a.svor0, r, z1);
        put_pre<NZ>(C, XF_DST, sn, cs, cl, ch, nops);
        put_pre<NZ>(A, XF_DST, sn, cs, al, ah, nops);
        const Row4 bl = row_load_g<NZ>(a.svor1, r, z0), bh = row_load_g<NZ>(a.svor1, r, z1);
        put_pre<NZ>(B, XF_DST, sn, cs, bl, bh, nops);
        if (t == 0) {
            park_row(pkC, 0, cl); park_row(pkA, 0, al); park_row(pkB, 0, bl);
            park_row(pkC, 1, row_load_g<NZ>(a.svor2, r, NZ));
            park_row(pkA, 1, row_load_g<NZ>(a.svor0, r, NZ));
            park_row(pkB, 1, row_load_g<NZ>(a.svor1, r, NZ));
        }
    }
    __syncthreads();

    // C -> semi-spectral zeta (inversion.f90:45, :142-144): DST + harmonic part; this is also the
    // semi-spectral zeta that feeds the inverse x/y passes (:81).  A (xi) rides along: its sine sum is the
    // interior of combine(xi_old), used by the semi-spectral projection below.
    xform2p<NZ>(C, XF_DST, pkC, A, XF_DST, pkA, scr);
#pragma unroll
    for (int it = 0; it < 3; ++it) {
        const int z = my_row<NZ>(it);
        if (z < 0) continue;
        Row4 c = row_load_s<NZ>(C, z);
        if (z >= 1 && z < NZ) {
#pragma unroll
            for (int s = 0; s < 4; ++s)
                c.v[s] += C[s * LC + cz<NZ>(0)] * phim_of<NZ, GEN>(scr, z, s) + C[s * LC + cz<NZ>(NZ)] * phip_of<NZ, GEN>(scr, z, s);
            row_store_s<NZ>(C, z, c);
        }
        row_store_g<NZ>(a.wsem2, r, z, c);
    }
    __syncthreads();
    // E = decompose(central_diffz(C)) (:46-47): FD, harmonic part removed -> pre-processed input of its DST (with
    // eta in B); rows 0 and NZ (the one-sided differences) ride through the transform
    {
        double e0[4], en[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const double* c = C + s * LC;
            e0[s] = g.dzi * (c[cz<NZ>(1)] - c[cz<NZ>(0)]);
            en[s] = g.dzi * (c[cz<NZ>(NZ)] - c[cz<NZ>(NZ - 1)]);
        }
        Row4 er[2];
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int z = it ? z1 : z0;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const double* c = C + s * LC;
                if (z == 0) er[it].v[s] = e0[s];
                else er[it].v[s] = (c[cz<NZ>(z + 1)] - c[cz<NZ>(z - 1)]) * g.hdzi
                                   - (e0[s] * phim_of<NZ, GEN>(scr, z, s) + en[s] * phip_of<NZ, GEN>(scr, z, s));
            }
        }
        put_pre<NZ>(E, XF_DST, sn, cs, er[0], er[1], nops);
        if (t == 0) {
#pragma unroll
            for (int s = 0; s < 4; ++s) { pkE[2 * s] = e0[s]; pkE[2 * s + 1] = en[s]; }
        }
    }
    __syncthreads();
    xform2p<NZ>(E, XF_DST, pkE, B, XF_DST, pkB, scr);

    // Solenoidal projection (:39-76), twice (see project_row):
    //  * semi-spectral rows (combine(xi_old), combine(eta_old), dzeta/dz) -> wsem0, wsem1, the vorticity that the
    //    inverse x/y passes take to physical space (:80-82);
    //  * mixed-spectral rows (xi_old, eta_old re-read from memory, E) -> new svor, and the source of the w
    //    inversion D2 = A_y - B_x (:86-90).
    // The Laplacian inversion (:108-122) follows in registers: ds = green * D2 (rows 1..nz-1; sine series of w),
    // as = rkz * ds (cosine series of dw/dz); both go to the buffers as pre-processed transform input once every
    // thread has read its rows.
    Row4 dsr[2], asr[2];
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
                const double pm = phim_of<NZ, GEN>(scr, z, s), pp = phip_of<NZ, GEN>(scr, z, s);
                sa.v[s] += A[s * LC + cz<NZ>(0)] * pm + A[s * LC + cz<NZ>(NZ)] * pp;
                sb.v[s] += B[s * LC + cz<NZ>(0)] * pm + B[s * LC + cz<NZ>(NZ)] * pp;
                se.v[s] = (c[cz<NZ>(z + 1)] - c[cz<NZ>(z - 1)]) * g.hdzi;
            }
        } else {
            se = fe;            // rows 0, NZ of E still hold the one-sided differences
        }
        project_row<GEN>(sa, sb, se, r);
        row_store_g<NZ>(a.wsem0, r, z, sa);
        row_store_g<NZ>(a.wsem1, r, z, sb);
        project_row<GEN>(fa, fb, fe, r);
        row_store_g<NZ>(a.svor0, r, z, fa);
        row_store_g<NZ>(a.svor1, r, z, fb);
        const Row4 ay2 = ddy(fa, r), bx2 = ddx(fb, r);
        Row4 d, as;
#pragma unroll
        for (int s = 0; s < 4; ++s) d.v[s] = ay2.v[s] - bx2.v[s];
        if (z >= 1 && z < NZ) {
            const double rk = __ldg(&g.rkz[z]);
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const double green = -1.0 / (r.k2[sy_of<GEN>(s)] + rk * rk);
                d.v[s] = green * d.v[s];
                as.v[s] = rk * d.v[s];
            }
        } else {
            // boundary values of D2 (:96-104): needed again in the last stage (keep); as rows 0 / NZ of the sine
            // series they ride through its transform (park: nobody else touches pkE between the two transforms)
#pragma unroll
            for (int s = 0; s < 4; ++s) scr.keep[(z == 0 ? 0 : 4) + s] = d.v[s];
            park_row(pkE, z == 0 ? 0 : 1, d);
            as = zero4;
        }
        if (it < 2) { dsr[it] = d; asr[it] = as; }
    }
    __syncthreads();          // every read of A, B, E is done: they take the input of the last two transforms
    {
        double ps[4] = {0.0, 0.0, 0.0, 0.0};
        put_pre<NZ>(E, XF_DST, sn, cs, dsr[0], dsr[1], nops);
        put_pre<NZ>(A, XF_DCT, sn, cs, asr[0], asr[1], ps);
        put_pre_sums<NZ>(ps, scr);
    }
    // horizontally averaged flow from the (0,0) column (:150-165): cosine transform in B slots 0, 1
    if (GEN && r.g00) {
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
    xform2p<NZ>(A, XF_DCT, pkA, E, XF_DST, pkE, scr);     // (:128-129)
    if (GEN && r.g00) xform2<NZ>(B, XF_DCT, nullptr, XF_DST, scr);

    // w = E + boundary part, dw/dz = es + as (:96-104, :136-139);
    // u = k2l2i (es_x + cs_y), v = k2l2i (es_y - cs_x), (0,0) <- ubar, vbar (:169-213)
    Hyp h[GEN ? 2 : 1];
#pragma unroll
    for (int sy = 0; sy < (GEN ? 2 : 1); ++sy) h[sy] = make_hyp(g, r, sy);
    double a00 = 0.0, a0n = 0.0, b00 = 0.0, b0n = 0.0;
    if (GEN && r.g00) {
        a00 = a.svor0[r.off[0]]; a0n = a.svor0[r.off[0] + NZ];
        b00 = a.svor1[r.off[0]]; b0n = a.svor1[r.off[0] + NZ];
    }
#pragma unroll
    for (int it = 0; it < 3; ++it) {
        const int z = my_row<NZ>(it);
        if (z < 0) continue;
        const Row4 as = row_load_s<NZ>(A, z), ds = row_load_s<NZ>(E, z), cs_ = row_load_s<NZ>(C, z);
        const double zm = __ldg(&g.zm[z]), zp = __ldg(&g.zp[z]);
        Theta th[GEN ? 2 : 1];
#pragma unroll
        for (int sy = 0; sy < (GEN ? 2 : 1); ++sy)
            th[sy] = hyp_theta(h[sy], zm, zp, phim_of<NZ, GEN>(scr, z, sy), phip_of<NZ, GEN>(scr, z, sy));
        Row4 es, w;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const Theta& q = th[sy_of<GEN>(s)];
            const double d0 = scr.keep[s], dn = scr.keep[4 + s];
            es.v[s] = d0 * q.dthm + dn * q.dthp + as.v[s];
            w.v[s] = (z == 0 || z == NZ) ? 0.0 : ds.v[s] + d0 * q.thm + dn * q.thp;
        }
        const Row4 ex = ddx(es, r), ey = ddy(es, r), cx = ddx(cs_, r), cy = ddy(cs_, r);
        Row4 u, v;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            u.v[s] = r.k2i[sy_of<GEN>(s)] * (ex.v[s] + cy.v[s]);
            v.v[s] = r.k2i[sy_of<GEN>(s)] * (ey.v[s] - cx.v[s]);
        }
        if (GEN && r.g00) {
            const double gt = __ldg(&g.gamtop[z]), gb = __ldg(&g.gambot[z]);
            u.v[0] = B[cz<NZ>(z)] + b0n * gt - b00 * gb;               // ubar (:163)
            v.v[0] = B[LC + cz<NZ>(z)] - a0n * gt + a00 * gb;          // vbar (:164)
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
constexpr size_t src_smem_bytes(bool gen) { return (size_t)(4 * ZCfg<NZ>::BUF + ZCfg<NZ>::aux(gen)) * sizeof(double); }

struct SrcArgs {
    const double* r; const double* q; const double* p;   // semi-spectral fluxes
    double* s0; double* s1; double* s2;                  // svorts (mixed spectral)
    // Crank-Nicolson update folded into the last stage (cn2.f90:120-135, 162-173 with the combine -> vdiss ->
    // decompose pairs collapsed, see k_cn2_update): upd < 0: store svorts;  upd = 0: vortsm = svor + c1 S,
    // svor = fac (vortsm + c1 S);  upd = 1: svor = fac (vortsm + c1 S).  svorts is then not stored at all.
    //   upd = 11..14: substep one..four of impl-diff-rk4 (impl_rk4.f90:212-364, pairs collapsed, see k_rk4_update):
    //     S' = pq S;  11: svori = svor, svor = mq (svori + c1 S'), svorf = svori + c2 S';  12, 13: svor = mq (svori + c1 S'),
    //     svorf += c2 S';  14: svor = mq (svorf + c1 S').  vortsm doubles as svori, wb is svorf.
    int upd;
    double c2;                    // rk4: second stage coefficient
    double* wb[3];                // rk4: svorf
    const double* mq;             // rk4: emq per column
    const double* pq;             // rk4: epq (or filt(0,:,:) in substep one) per column
    double c1;                    // dt/2 (cn2); first stage coefficient (rk4)
    double* svor[3];
    double* vortsm[3];
    const double* f2d;            // vdiss * filt2d per column
    const double* filtz;          // z part of the filter, [nz+1]
    const double* vd;             // vdiss per column: the (0,0) column has filt = 1 (inversion_utils.f90:275-277)
};

// semi-spectral curl of one row, first two components, from row z of r and the rows z-1/z+1 (lo/hi) of q, p;
// dz = 1/(2 dz) in the interior, 1/dz with (lo, hi) = (z, z+1) or (z-1, z) at the boundaries
// (inversion_utils.f90:653-680)
__device__ __forceinline__ void curl01_row(const Row4& fr, const Row4& qlo, const Row4& qhi, const Row4& plo,
                                           const Row4& phi, double dz, const Grp& r, Row4& s0, Row4& s1) {
    const Row4 ry = ddy(fr, r), rx = ddx(fr, r);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        s0.v[s] = ry.v[s] - (qhi.v[s] - qlo.v[s]) * dz;      // dr/dy - dq/dz (:341-347)
        s1.v[s] = (phi.v[s] - plo.v[s]) * dz - rx.v[s];      // dp/dz - dr/dx (:353-359)
    }
}
// third component from row z of q, p
__device__ __forceinline__ Row4 curl2_row(const Row4& fq, const Row4& fp, const Grp& r) {
    const Row4 qx = ddx(fq, r), py = ddy(fp, r);
    Row4 s2;
#pragma unroll
    for (int s = 0; s < 4; ++s) s2.v[s] = qx.v[s] - py.v[s]; // dq/dx - dp/dy (:363-367)
    return s2;
}

// prefetch of the rows this thread owns of one 4-slot field into a column buffer (asynchronous copies; the same
// thread reads them back after ps_cp_async_wait, so no barrier is involved)
template <int NZ>
__device__ __forceinline__ void rows_prefetch(double* buf, const double* __restrict__ src, const Grp& r) {
    constexpr int LC = ZCfg<NZ>::LC;
#pragma unroll
    for (int it = 0; it < 3; ++it) {
        const int z = my_row<NZ>(it);
        if (z < 0) continue;
        const int zz = cz<NZ>(z);
#pragma unroll
        for (int s = 0; s < 4; ++s)
            if (s < 2 || !r.dupx) ps_cp_async8(buf + s * LC + zz, src + r.off[s] + z);
    }
}

// last stage of the source kernel for one component: svorts row -> memory, or the Crank-Nicolson update with it
template <int NZ, bool GEN, bool RK>
__device__ __forceinline__ void src_finish(const SrcArgs& a, int comp, const double* S, const double* X, double* sv_out,
                                           const SpecGeom& g, const Grp& r, const double (&f2)[4]) {
#pragma unroll
    for (int it = 0; it < 3; ++it) {
        const int z = my_row<NZ>(it);
        if (z < 0) continue;
        const Row4 sr = row_load_s<NZ>(S, z);
        if (a.upd < 0) { row_store_g<NZ>(sv_out, r, z, sr); continue; }
        const Row4 x = row_load_s<NZ>(X, z);             // the prefetched operand: svor / vortsm (cn2), svor / svori / svorf (rk4)
        if (RK) {
            const int st = a.upd - 10;
            Row4 wbv, out, wbo;
            if (st == 2 || st == 3) wbv = row_load_g<NZ>(a.wb[comp], r, z);
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const long long col = r.off[s] / g.pz;
                const double sp = __ldg(&a.pq[col]) * sr.v[s];
                out.v[s] = __ldg(&a.mq[col]) * (x.v[s] + a.c1 * sp);
                wbo.v[s] = ((st == 1) ? x.v[s] : wbv.v[s]) + a.c2 * sp;
            }
            if (st == 1) row_store_g<NZ>(a.vortsm[comp], r, z, x);
            if (st <= 3) row_store_g<NZ>(a.wb[comp], r, z, wbo);
            row_store_g<NZ>(a.svor[comp], r, z, out);
            continue;
        }
        const double fz = __ldg(&a.filtz[z]);
        Row4 sm, out;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            double fac = f2[s] * fz;
            if (GEN && r.g00 && s == 0) fac = f2[s];      // (0,0): vdiss only
            sm.v[s] = (a.upd == 0) ? x.v[s] + a.c1 * sr.v[s] : x.v[s];
            out.v[s] = fac * (sm.v[s] + a.c1 * sr.v[s]);
        }
        if (a.upd == 0) row_store_g<NZ>(a.vortsm[comp], r, z, sm);
        row_store_g<NZ>(a.svor[comp], r, z, out);
    }
}

// RK: the instantiation that carries an impl-diff-rk4 substep (upd = 11..14); the other one (upd < 11: plain or cn2)
// is the kernel of the headline path and stays free of the rk4 operands
template <int NZ, bool GEN, bool RK = false>
__global__ void __launch_bounds__(ZCfg<NZ>::NT, ZCfg<NZ>::ctas(4, GEN)) k_source_spec(SpecGeom g, SrcArgs a) {
    PS_SMEM(double, sm);
    constexpr int BUF = ZCfg<NZ>::BUF;
    double* R = sm;           // curl component 0 (r itself stays in the registers of its row owner)
    double* Q = R + BUF;      // q, then curl component 2
    double* P = Q + BUF;      // p, then the operand of the update
    double* T = P + BUF;      // curl component 1
    const ZScr<NZ> scr = make_scr<NZ, GEN>(T + BUF);
    double* pkR = scr.park;
    double* pkQ = pkR + 8;
    double* pkT = pkQ + 8;
    scr_init<NZ>(scr, g);
    const Grp r = make_grp<GEN>(g, blockIdx.x);
    phi_fill<NZ, GEN>(scr, g, r);
    double sn, cs;
    my_sincos<NZ>(g, sn, cs);
    const int t = threadIdx.x;
    const int z0 = my_row<NZ>(0), z1 = my_row<NZ>(1);
    double nops[4] = {0.0, 0.0, 0.0, 0.0};
    Row4 zero4;
    zero4.v[0] = zero4.v[1] = zero4.v[2] = zero4.v[3] = 0.0;

    // stage q and p (their neighbour rows are needed for d/dz); every flux element is read from memory exactly once,
    // the rows of r are used by their owner only
#pragma unroll
    for (int it = 0; it < 3; ++it) {
        const int z = my_row<NZ>(it);
        if (z < 0) continue;
        row_store_s<NZ>(Q, z, row_load_g<NZ>(a.q, r, z));
        row_store_s<NZ>(P, z, row_load_g<NZ>(a.p, r, z));
    }
    const Row4 rl = row_load_g<NZ>(a.r, r, z0), rh = row_load_g<NZ>(a.r, r, z1);
    __syncthreads();
    // boundary rows of the curl (they define the harmonic part removed from every interior row): thread 0 owns
    // both; parked in shared memory as keep[(c*2 + top)*4 + slot]
    if (t == 0) {
#pragma unroll
        for (int top = 0; top < 2; ++top) {
            const int z = top ? NZ : 0, zl = top ? NZ - 1 : 0, zh = top ? NZ : 1;
            Row4 c0, c1;
            curl01_row(top ? row_load_g<NZ>(a.r, r, NZ) : rl, row_load_s<NZ>(Q, zl), row_load_s<NZ>(Q, zh), row_load_s<NZ>(P, zl),
                       row_load_s<NZ>(P, zh), g.dzi, r, c0, c1);
            const Row4 c2 = curl2_row(row_load_s<NZ>(Q, z), row_load_s<NZ>(P, z), r);
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                scr.keep[(0 + top) * 4 + s] = c0.v[s];
                scr.keep[(2 + top) * 4 + s] = c1.v[s];
                scr.keep[(4 + top) * 4 + s] = c2.v[s];
            }
            // rows 0 and NZ of the three sine transforms
            park_row(pkR, top, c0); park_row(pkT, top, c1); park_row(pkQ, top, c2);
        }
    }
    __syncthreads();
    // the three curl components of this thread's row pair, harmonic part removed; components 0 and 1 go to R and T
    // as pre-processed transform input at once (both buffers are free), component 2 takes the place of q once every
    // thread has read its neighbour rows
    Row4 c2r[2];
    {
        Row4 c0r[2], c1r[2];
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int z = it ? z1 : z0;
            if (z == 0) {
#pragma unroll
                for (int s = 0; s < 4; ++s) { c0r[it].v[s] = scr.keep[s]; c1r[it].v[s] = scr.keep[8 + s]; c2r[it].v[s] = scr.keep[16 + s]; }
            } else {
                curl01_row(it ? rh : rl, row_load_s<NZ>(Q, z - 1), row_load_s<NZ>(Q, z + 1),
                           row_load_s<NZ>(P, z - 1), row_load_s<NZ>(P, z + 1), g.hdzi, r, c0r[it], c1r[it]);
                c2r[it] = curl2_row(row_load_s<NZ>(Q, z), row_load_s<NZ>(P, z), r);
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const double pm = phim_of<NZ, GEN>(scr, z, s), pp = phip_of<NZ, GEN>(scr, z, s);
                    c0r[it].v[s] -= scr.keep[0 + s] * pm + scr.keep[4 + s] * pp;
                    c1r[it].v[s] -= scr.keep[8 + s] * pm + scr.keep[12 + s] * pp;
                    c2r[it].v[s] -= scr.keep[16 + s] * pm + scr.keep[20 + s] * pp;
                }
            }
        }
        put_pre<NZ>(R, XF_DST, sn, cs, c0r[0], c0r[1], nops);
        put_pre<NZ>(T, XF_DST, sn, cs, c1r[0], c1r[1], nops);
    }
    __syncthreads();          // every neighbour row of q, p has been read
    put_pre<NZ>(Q, XF_DST, sn, cs, c2r[0], c2r[1], nops);
    // per-slot factor of the update: vdiss * filt2d of the column ((0,0): vdiss, filt = 1)
    double f2[4] = {0.0, 0.0, 0.0, 0.0};
    if (!RK && a.upd >= 0) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const long long col = r.off[s] / g.pz;
            f2[s] = (GEN && r.g00 && s == 0) ? __ldg(&a.vd[0]) : __ldg(&a.f2d[col]);
        }
    }
    // operand of the update: svor (cn2 first update, rk4 substep one), vortsm = svori (cn2 iterations, rk4 two / three),
    // svorf (rk4 substep four)
    const bool op_svor = (a.upd == 0 || a.upd == 11), op_wb = RK && (a.upd == 14);
    const double* x0 = op_svor ? a.svor[0] : op_wb ? a.wb[0] : a.vortsm[0];
    const double* x1 = op_svor ? a.svor[1] : op_wb ? a.wb[1] : a.vortsm[1];
    const double* x2 = op_svor ? a.svor[2] : op_wb ? a.wb[2] : a.vortsm[2];
    // component 2 first (the half-filled transform round); its update operand lands in P, free since the curl
    if (a.upd >= 0) rows_prefetch<NZ>(P, x2, r);
    __syncthreads();
    xform2p<NZ>(Q, XF_DST, pkQ, nullptr, XF_DST, pkQ, scr);
    if (a.upd >= 0) ps_cp_async_wait();
    src_finish<NZ, GEN, RK>(a, 2, Q, P, a.s2, g, r, f2);
    // components 0, 1: operands into P and Q (this thread only ever touches its own rows of them from here on)
    if (a.upd >= 0) { rows_prefetch<NZ>(P, x0, r); rows_prefetch<NZ>(Q, x1, r); }
    xform2p<NZ>(R, XF_DST, pkR, T, XF_DST, pkT, scr);
    if (a.upd >= 0) ps_cp_async_wait();
    src_finish<NZ, GEN, RK>(a, 0, R, P, a.s0, g, r, f2);
    src_finish<NZ, GEN, RK>(a, 1, T, Q, a.s1, g, r, f2);
}

}  // namespace ps3d
