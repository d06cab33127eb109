// Column kernels: everything the reference does along z or pointwise in
// (kx, ky) on spectral data, fused per "4-group" of columns
//   { (b, a), (b, ny-a), (nx-b, a), (nx-b, ny-a) }
// so that diffx/diffy (which couple Hermitian slot k with slot n-k,
// reference sta3dfft.f90:304-377, via mpi_reverse.f90:332-398) are local
// register/shared-memory operations and no mirrored data is ever re-read from
// HBM.  One block owns one group: the 4 columns (nz+1 doubles each, contiguous
// in memory) are staged in shared memory, all z transforms and pointwise
// operators run there, results are written back once.
//
// Spectral arrays are [kx][kyl][pz] with kyl the rank-local index of the
// *paired* ky order  ky' = 0, ny/2, 1, ny-1, 2, ny-2, ...  (kyl = 2a'+sy), so
// the y-mirror of a column is its neighbour kyl^1 and a slab of ky' is closed
// under mirroring (replaces mpi_reverse entirely).
//
// DST-I / DCT-I of length nz (reference stafft.f90:410-550 conventions) are
// done two columns at a time as one complex FFT of length 2 nz over the odd /
// even extensions: accurate (no post-processing recurrence) and built on the
// same block_cfft as the x/y passes.
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
    static constexpr int M = 2 * NZ;              // pair-FFT length
    static constexpr int TPF = M / 8;             // threads per pair FFT
    static constexpr int NT = 2 * TPF;            // 2 pair FFTs = 4 slots of one field
    static constexpr int LC = NZ + 2;             // column buffer stride (rows 0..NZ used)
    static constexpr int PL = padded_len(M);
    static constexpr int SCR = 2 * 2 * PL;        // doubles of FFT scratch
    static constexpr int BUF = 4 * LC;            // doubles per 4-slot field buffer
};

// ---- group bookkeeping -----------------------------------------------------
struct Grp {
    int b, ap;
    bool dupx;          // b == 0 or b == nx/2: the x-mirror slot is the column itself (inactive)
    bool g00;           // this group holds the (0,0) column in slot 0
    double kx, ky;      // diff wavenumbers
    double k2[2], k2i[2];
    long long off[4];   // column offsets (doubles) of slots s = 2*sx + sy
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
        r.off[sy] = ((long long)r.b * g.nyl + kyl) * g.pz;
        r.off[2 + sy] = ((long long)kxm * g.nyl + kyl) * g.pz;
    }
    return r;
}

__device__ __forceinline__ bool slot_active(const Grp& r, int s) { return !(r.dupx && s >= 2); }

template <int NZ>
__device__ __forceinline__ void col_load(double* buf, const double* __restrict__ src, const Grp& r) {
    constexpr int LC = ZCfg<NZ>::LC;
    for (int i = threadIdx.x; i < 4 * LC; i += blockDim.x) {
        const int s = i / LC, z = i - s * LC;
        buf[i] = (z <= NZ && slot_active(r, s)) ? src[r.off[s] + z] : 0.0;
    }
}

template <int NZ>
__device__ __forceinline__ void col_store(double* __restrict__ dst, const double* buf, const Grp& r) {
    constexpr int LC = ZCfg<NZ>::LC;
    for (int i = threadIdx.x; i < 4 * LC; i += blockDim.x) {
        const int s = i / LC, z = i - s * LC;
        if (z <= NZ && slot_active(r, s)) dst[r.off[s] + z] = buf[i];
    }
}

// ---- harmonic (Laplace) functions on the fly -------------------------------
// EP/EM[sy][i] = exp(-kl zp_i), exp(-kl zm_i)   (inversion_utils.f90:505-509)
template <int NZ>
__device__ __forceinline__ void hyp_tables(double* EP, double* EM, const SpecGeom& g, const Grp& r) {
    constexpr int LC = ZCfg<NZ>::LC;
    for (int i = threadIdx.x; i < 2 * LC; i += blockDim.x) {
        const int sy = i / LC, z = i - sy * LC;
        double ep = 0.0, em = 0.0;
        if (z <= NZ) {
            const double kl = sqrt(r.k2[sy]);
            ep = exp(-(kl * __ldg(&g.zp[z])));
            em = exp(-(kl * __ldg(&g.zm[z])));
        }
        EP[i] = ep; EM[i] = em;
    }
}

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

__device__ __forceinline__ void hyp_phi(const Hyp& h, const SpecGeom& g, double ep, double em, int z,
                                        double& phim, double& phip) {
    if (h.lin) {
        phim = __ldg(&g.zm[z]) / g.Lz;
        phip = __ldg(&g.zp[z]) / g.Lz;
    } else {
        phim = h.div * (ep - h.ef * em);
        phip = h.div * (em - h.ef * ep);
    }
}

// thetam, thetap, dthetam, dthetap (inversion_utils.f90:518-541)
__device__ __forceinline__ void hyp_theta(const Hyp& h, const SpecGeom& g, double ep, double em, int z,
                                          double& thm, double& thp, double& dthm, double& dthp) {
    if (h.lin) { thm = thp = dthm = dthp = 0.0; return; }
    const double Lm = h.kl * __ldg(&g.zm[z]);
    const double Lp = h.kl * __ldg(&g.zp[z]);
    const double phim = h.div * (ep - h.ef * em);
    const double phip = h.div * (em - h.ef * ep);
    const double dphim = -h.kl * h.div * (ep + h.ef * em);
    const double dphip = h.kl * h.div * (em + h.ef * ep);
    thm = h.k2if * (h.R * Lm * phip - h.Q * Lp * phim);
    thp = h.k2if * (h.R * Lp * phim - h.Q * Lm * phip);
    dthm = -h.k2if * ((h.Q * Lp - 1.0) * dphim - h.R * Lm * dphip);
    dthp = -h.k2if * ((h.Q * Lm - 1.0) * dphip - h.R * Lp * dphim);
}

// ---- z transforms on a 4-slot shared-memory field ---------------------------
// DST-I of rows 1..NZ-1 of each slot (scaled sqrt(2/NZ)); rows 0 and NZ are
// neither read nor written (stafft.f90:509-513).  Ends with a barrier.
template <int NZ>
__device__ __forceinline__ void dst4(double* X, double* scr, const SpecGeom& g) {
    constexpr int M = ZCfg<NZ>::M, TPF = ZCfg<NZ>::TPF, LC = ZCfg<NZ>::LC, PL = ZCfg<NZ>::PL;
    const int t = threadIdx.x;
    const int f = t / TPF, u = t - f * TPF;
    const bool active = f < 2;
    const double* x0 = X + (2 * f) * LC;
    const double* x1 = x0 + LC;
    double vr[8], vi[8];
    if (active) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int j = u + e * (M / 8);
            double a = 0.0, c = 0.0;
            if (j > 0 && j < NZ) { a = x0[j]; c = x1[j]; }
            else if (j > NZ) { a = -x0[M - j]; c = -x1[M - j]; }
            vr[e] = a; vi[e] = c;
        }
    }
    double* sre = scr + (active ? f : 0) * 2 * PL;
    double* sim = sre + PL;
    block_cfft<M, false>(vr, vi, u, active, sre, sim, g.tw, g.ntw / M);
    const double sc = rsqrt((double)M);     // 1/sqrt(2 nz)
    if (active) {
        double* y0 = X + (2 * f) * LC;
        double* y1 = y0 + LC;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k = u + e * (M / 8);
            if (k >= 1 && k < NZ) { y0[k] = -vi[e] * sc; y1[k] = vr[e] * sc; }
        }
    }
    __syncthreads();
}

// DCT-I of rows 0..NZ of each slot (scaled sqrt(2/NZ)).  Ends with a barrier.
template <int NZ>
__device__ __forceinline__ void dct4(double* X, double* scr, const SpecGeom& g) {
    constexpr int M = ZCfg<NZ>::M, TPF = ZCfg<NZ>::TPF, LC = ZCfg<NZ>::LC, PL = ZCfg<NZ>::PL;
    const int t = threadIdx.x;
    const int f = t / TPF, u = t - f * TPF;
    const bool active = f < 2;
    const double* x0 = X + (2 * f) * LC;
    const double* x1 = x0 + LC;
    double vr[8], vi[8];
    if (active) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int j = u + e * (M / 8);
            const int jj = (j <= NZ) ? j : M - j;
            vr[e] = x0[jj]; vi[e] = x1[jj];
        }
    }
    double* sre = scr + (active ? f : 0) * 2 * PL;
    double* sim = sre + PL;
    block_cfft<M, false>(vr, vi, u, active, sre, sim, g.tw, g.ntw / M);
    const double sc = rsqrt((double)M);
    if (active) {
        double* y0 = X + (2 * f) * LC;
        double* y1 = y0 + LC;
#pragma unroll
        for (int e = 0; e < 5; ++e) {
            const int k = u + e * (M / 8);
            if (k <= NZ) { y0[k] = vr[e] * sc; y1[k] = vi[e] * sc; }
        }
    }
    __syncthreads();
}

// X(1..NZ-1) += / -= X(0) phim + X(NZ) phip   (inversion_utils.f90:571, 643)
template <int NZ, int SIGN>
__device__ __forceinline__ void harmonic(double* X, const double* EP, const double* EM,
                                         const SpecGeom& g, const Grp& r) {
    constexpr int LC = ZCfg<NZ>::LC;
    const Hyp h0 = make_hyp(g, r, 0), h1 = make_hyp(g, r, 1);
    for (int i = threadIdx.x; i < 4 * LC; i += blockDim.x) {
        const int s = i / LC, z = i - s * LC, sy = s & 1;
        if (z >= 1 && z < NZ) {
            double phim, phip;
            hyp_phi(sy ? h1 : h0, g, EP[sy * LC + z], EM[sy * LC + z], z, phim, phip);
            const double hpart = X[s * LC] * phim + X[s * LC + NZ] * phip;
            X[i] = (SIGN > 0) ? X[i] + hpart : X[i] - hpart;
        }
    }
    __syncthreads();
}

// field_combine_semi_spectral in place (inversion_utils.f90:617-647)
template <int NZ>
__device__ __forceinline__ void combine4(double* X, double* scr, const double* EP, const double* EM,
                                         const SpecGeom& g, const Grp& r) {
    dst4<NZ>(X, scr, g);
    harmonic<NZ, +1>(X, EP, EM, g, r);
}

// field_decompose_semi_spectral in place (inversion_utils.f90:563-592)
template <int NZ>
__device__ __forceinline__ void decompose4(double* X, double* scr, const double* EP, const double* EM,
                                           const SpecGeom& g, const Grp& r) {
    harmonic<NZ, -1>(X, EP, EM, g, r);
    dst4<NZ>(X, scr, g);
}

// D = central_diffz(S)  (inversion_utils.f90:653-673).  Ends with a barrier.
template <int NZ>
__device__ __forceinline__ void diffz4(double* D, const double* S, const SpecGeom& g) {
    constexpr int LC = ZCfg<NZ>::LC;
    for (int i = threadIdx.x; i < 4 * LC; i += blockDim.x) {
        const int s = i / LC, z = i - s * LC;
        if (z == 0) D[i] = g.dzi * (S[i + 1] - S[i]);
        else if (z == NZ) D[i] = g.dzi * (S[i] - S[i - 1]);
        else if (z < NZ) D[i] = (S[i + 1] - S[i - 1]) * g.hdzi;
    }
    __syncthreads();
}

// d/dx and d/dy of a 4-slot field at element i = s*LC + z (sta3dfft.f90:325-329):
//   slot sx = 0 (kx = b):    ds = -kx * f(nx-b) ;  slot sx = 1 (kx = nx-b): ds = +kx * f(b)
template <int NZ>
__device__ __forceinline__ double ddx(const double* F, int i, int s, const Grp& r) {
    constexpr int LC = ZCfg<NZ>::LC;
    return (s & 2) ? r.kx * F[i - 2 * LC] : -r.kx * F[i + 2 * LC];
}
template <int NZ>
__device__ __forceinline__ double ddy(const double* F, int i, int s, const Grp& r) {
    constexpr int LC = ZCfg<NZ>::LC;
    return (s & 1) ? r.ky * F[i - LC] : -r.ky * F[i + LC];
}

// ---------------------------------------------------------------------------
// Operator-mode kernel: one z-operation on one field (drop-in for the public
// module procedures fftsine, fftcosine, diffx, diffy, central_diffz,
// field_combine_semi_spectral, field_decompose_semi_spectral).
// ---------------------------------------------------------------------------
enum { ZOP_SINE = 0, ZOP_COSINE, ZOP_COMBINE, ZOP_DECOMPOSE, ZOP_DIFFZ, ZOP_DIFFX, ZOP_DIFFY };

template <int NZ>
constexpr size_t zop_smem_bytes() { return (size_t)(3 * ZCfg<NZ>::BUF + ZCfg<NZ>::SCR) * sizeof(double); }

template <int NZ>
__global__ void __launch_bounds__(ZCfg<NZ>::NT) k_zop(SpecGeom g, int op, const double* __restrict__ in,
                                                       double* __restrict__ out) {
    PS_SMEM(double, sm);
    constexpr int LC = ZCfg<NZ>::LC, BUF = ZCfg<NZ>::BUF;
    double* X = sm;
    double* Y = X + BUF;
    double* EP = Y + BUF;
    double* EM = EP + 2 * LC;
    double* scr = EP + BUF;
    const Grp r = make_grp(g, blockIdx.x);
    col_load<NZ>(X, in, r);
    if (op == ZOP_COMBINE || op == ZOP_DECOMPOSE) hyp_tables<NZ>(EP, EM, g, r);
    __syncthreads();
    if (op == ZOP_SINE) {
        dst4<NZ>(X, scr, g);
        for (int s = threadIdx.x; s < 4; s += blockDim.x) X[s * LC + NZ] = 0.0;   // stafft.f90:546-549
        __syncthreads();
        col_store<NZ>(out, X, r);
    } else if (op == ZOP_COSINE) {
        dct4<NZ>(X, scr, g);
        col_store<NZ>(out, X, r);
    } else if (op == ZOP_COMBINE) {
        combine4<NZ>(X, scr, EP, EM, g, r);
        col_store<NZ>(out, X, r);
    } else if (op == ZOP_DECOMPOSE) {
        decompose4<NZ>(X, scr, EP, EM, g, r);
        col_store<NZ>(out, X, r);
    } else if (op == ZOP_DIFFZ) {
        diffz4<NZ>(Y, X, g);
        col_store<NZ>(out, Y, r);
    } else {
        for (int i = threadIdx.x; i < 4 * LC; i += blockDim.x) {
            const int s = i / LC;
            Y[i] = (op == ZOP_DIFFX) ? ddx<NZ>(X, i, s, r) : ddy<NZ>(X, i, s, r);
        }
        __syncthreads();
        col_store<NZ>(out, Y, r);
    }
}

// ---------------------------------------------------------------------------
// vor2vel, spectral part (reference inversion.f90:23-226 minus the six
// fftxys2p calls): svor -> svor (solenoidal), semi-spectral vorticity (input of
// the inverse x/y passes that give `vor`), svel.
// ---------------------------------------------------------------------------
template <int NZ>
constexpr size_t v2v_smem_bytes() { return (size_t)(6 * ZCfg<NZ>::BUF + ZCfg<NZ>::SCR + 16) * sizeof(double); }

struct V2VArgs {
    double* svor0; double* svor1; const double* svor2;   // in/out, in/out, in
    double* wsem0; double* wsem1; double* wsem2;         // semi-spectral vorticity (out)
    double* svel0; double* svel1; double* svel2;         // semi-spectral velocity (out)
};

template <int NZ>
__global__ void __launch_bounds__(ZCfg<NZ>::NT) k_vor2vel_spec(SpecGeom g, V2VArgs a) {
    PS_SMEM(double, sm);
    constexpr int LC = ZCfg<NZ>::LC, BUF = ZCfg<NZ>::BUF;
    double* A = sm;
    double* B = A + BUF;
    double* C = B + BUF;
    double* D = C + BUF;
    double* E = D + BUF;
    double* EP = E + BUF;
    double* EM = EP + 2 * LC;
    double* scr = EP + BUF;
    double* bnd = scr + ZCfg<NZ>::SCR;       // [8] D(0), D(nz) per slot
    const int tid = threadIdx.x, nt = blockDim.x;
    const Grp r = make_grp(g, blockIdx.x);

    col_load<NZ>(A, a.svor0, r);
    col_load<NZ>(B, a.svor1, r);
    col_load<NZ>(C, a.svor2, r);
    hyp_tables<NZ>(EP, EM, g, r);
    __syncthreads();

    // D = B_x - A_y   (inversion.f90:39-42)
    for (int i = tid; i < 4 * LC; i += nt) {
        const int s = i / LC;
        D[i] = ddx<NZ>(B, i, s, r) - ddy<NZ>(A, i, s, r);
    }
    // C -> semi-spectral zeta (inversion.f90:45); E = C_z, decomposed (:46-47)
    combine4<NZ>(C, scr, EP, EM, g, r);
    diffz4<NZ>(E, C, g);
    decompose4<NZ>(E, scr, EP, EM, g, r);

    // A = k2l2i (E_x + D_y), B = k2l2i (E_y - D_x); (0,0) column keeps its mean (:55-76)
    for (int i = tid; i < 4 * LC; i += nt) {
        const int s = i / LC;
        if (r.g00 && s == 0) continue;
        const double k2i = r.k2i[s & 1];
        A[i] = k2i * (ddx<NZ>(E, i, s, r) + ddy<NZ>(D, i, s, r));
        B[i] = k2i * (ddy<NZ>(E, i, s, r) - ddx<NZ>(D, i, s, r));
    }
    __syncthreads();
    col_store<NZ>(a.svor0, A, r);
    col_store<NZ>(a.svor1, B, r);

    // source of the w inversion: D = A_y - B_x (mixed spectral) (:86-90)
    for (int i = tid; i < 4 * LC; i += nt) {
        const int s = i / LC;
        D[i] = ddy<NZ>(A, i, s, r) - ddx<NZ>(B, i, s, r);
    }
    // horizontally averaged flow from the (0,0) column (:150-165) -> E slots 0 (ubar), 1 (vbar)
    if (r.g00) {
        for (int i = tid; i < 4 * LC; i += nt) {
            const int s = i / LC, z = i - s * LC;
            double v = 0.0;
            if (z >= 1 && z < NZ) {
                const double rkzi = 1.0 / __ldg(&g.rkz[z]);
                if (s == 0) v = -rkzi * B[z];
                else if (s == 1) v = rkzi * A[z];
            }
            E[i] = v;
        }
        __syncthreads();
        dct4<NZ>(E, scr, g);
        for (int z = tid; z <= NZ; z += nt) {
            const double gt = __ldg(&g.gamtop[z]), gb = __ldg(&g.gambot[z]);
            E[z] = E[z] + B[NZ] * gt - B[0] * gb;            // ubar
            E[LC + z] = E[LC + z] - A[NZ] * gt + A[0] * gb;  // vbar
        }
    }
    __syncthreads();

    // vorticity to semi-spectral space for the inverse x/y passes (:80-82)
    combine4<NZ>(A, scr, EP, EM, g, r);
    combine4<NZ>(B, scr, EP, EM, g, r);
    col_store<NZ>(a.wsem0, A, r);
    col_store<NZ>(a.wsem1, B, r);
    col_store<NZ>(a.wsem2, C, r);
    for (int s = tid; s < 4; s += nt) { bnd[s] = D[s * LC]; bnd[4 + s] = D[s * LC + NZ]; }
    __syncthreads();

    // invert Laplacian (:108-122): D <- green * D (rows 1..nz-1), A <- rkz * D
    for (int i = tid; i < 4 * LC; i += nt) {
        const int s = i / LC, z = i - s * LC;
        double as = 0.0;
        if (z >= 1 && z < NZ) {
            const double rk = __ldg(&g.rkz[z]);
            const double green = -1.0 / (r.k2[s & 1] + rk * rk);
            const double d = green * D[i];
            D[i] = d;
            as = rk * d;
        }
        A[i] = as;
    }
    __syncthreads();
    dct4<NZ>(A, scr, g);     // (:128)
    dst4<NZ>(D, scr, g);     // (:129)
    // w = D + boundary part, dw/dz = es + as  (:96-104, :136-139); B <- dw/dz
    {
        const Hyp h0 = make_hyp(g, r, 0), h1 = make_hyp(g, r, 1);
        for (int i = tid; i < 4 * LC; i += nt) {
            const int s = i / LC, z = i - s * LC, sy = s & 1;
            if (z > NZ) continue;
            double thm, thp, dthm, dthp;
            hyp_theta(sy ? h1 : h0, g, EP[sy * LC + z], EM[sy * LC + z], z, thm, thp, dthm, dthp);
            const double d0 = bnd[s], dn = bnd[4 + s];
            B[i] = d0 * dthm + dn * dthp + A[i];
            D[i] = (z == 0 || z == NZ) ? 0.0 : D[i] + d0 * thm + dn * thp;
        }
    }
    __syncthreads();
    // u = k2l2i (es_x + cs_y), v = k2l2i (es_y - cs_x), (0,0) <- ubar, vbar (:169-213)
    for (int i = tid; i < 4 * LC; i += nt) {
        const int s = i / LC, z = i - s * LC;
        if (z > NZ || !slot_active(r, s)) continue;
        const double k2i = r.k2i[s & 1];
        double u = k2i * (ddx<NZ>(B, i, s, r) + ddy<NZ>(C, i, s, r));
        double v = k2i * (ddy<NZ>(B, i, s, r) - ddx<NZ>(C, i, s, r));
        if (r.g00 && s == 0) { u = E[z]; v = E[LC + z]; }
        a.svel0[r.off[s] + z] = u;
        a.svel1[r.off[s] + z] = v;
        a.svel2[r.off[s] + z] = D[i];
    }
}

// ---------------------------------------------------------------------------
// vorticity tendency, spectral part (reference inversion.f90:298-371 after the
// fftxyp2s calls).  Inputs are the x/y-transformed fluxes r, q, p
// (semi-spectral); central_diffz commutes with the x/y FFT, so dq/dz and dp/dz
// are formed here instead of through two more 2-D FFTs.
// ---------------------------------------------------------------------------
template <int NZ>
constexpr size_t src_smem_bytes() { return (size_t)(6 * ZCfg<NZ>::BUF + ZCfg<NZ>::SCR) * sizeof(double); }

struct SrcArgs {
    const double* r; const double* q; const double* p;   // semi-spectral fluxes
    double* s0; double* s1; double* s2;                  // svorts (mixed spectral)
};

template <int NZ>
__global__ void __launch_bounds__(ZCfg<NZ>::NT) k_source_spec(SpecGeom g, SrcArgs a) {
    PS_SMEM(double, sm);
    constexpr int LC = ZCfg<NZ>::LC, BUF = ZCfg<NZ>::BUF;
    double* R = sm;
    double* Q = R + BUF;
    double* P = Q + BUF;
    double* DQ = P + BUF;
    double* DP = DQ + BUF;
    double* EP = DP + BUF;
    double* EM = EP + 2 * LC;
    double* scr = EP + BUF;
    const int tid = threadIdx.x, nt = blockDim.x;
    const Grp r = make_grp(g, blockIdx.x);
    col_load<NZ>(R, a.r, r);
    col_load<NZ>(Q, a.q, r);
    col_load<NZ>(P, a.p, r);
    hyp_tables<NZ>(EP, EM, g, r);
    __syncthreads();
    diffz4<NZ>(DQ, Q, g);
    diffz4<NZ>(DP, P, g);
    decompose4<NZ>(R, scr, EP, EM, g, r);
    decompose4<NZ>(Q, scr, EP, EM, g, r);
    decompose4<NZ>(P, scr, EP, EM, g, r);
    decompose4<NZ>(DQ, scr, EP, EM, g, r);
    decompose4<NZ>(DP, scr, EP, EM, g, r);
    for (int i = tid; i < 4 * LC; i += nt) {
        const int s = i / LC, z = i - s * LC;
        if (z > NZ || !slot_active(r, s)) continue;
        a.s0[r.off[s] + z] = ddy<NZ>(R, i, s, r) - DQ[i];                      // dr/dy - dq/dz
        a.s1[r.off[s] + z] = DP[i] - ddx<NZ>(R, i, s, r);                      // dp/dz - dr/dx
        a.s2[r.off[s] + z] = ddx<NZ>(Q, i, s, r) - ddy<NZ>(P, i, s, r);        // dq/dx - dp/dy
    }
}

}  // namespace ps3d
