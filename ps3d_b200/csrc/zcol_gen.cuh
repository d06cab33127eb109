// Column kernels for nz that is not a power of two (nz = 2^a 3^b 5^c, even): the algorithms of zcol.cuh -- one
// block per 4-group of columns, columns staged in shared memory, harmonic functions built per block, DST-I / DCT-I
// by the reference's reduction to a real FFT of length nz with two columns per complex FFT (stafft.f90:410-550) --
// written with plain loops over z and the mixed-radix transform of gen_fft.cuh.  Coverage path for the grids the
// reference's factorisen accepts (stafft.f90:128-187); the configurations of BASELINE.json (powers of two) run
// through zcol.cuh.  Every group uses the general form (per-column k^2 + l^2, (0,0) column included).
#pragma once

#include "gen_fft.cuh"
#include "zcol.cuh"

namespace ps3d {

struct GenZ {
    GenPlan plan;           // factorisation of nz
    const double2* tw;      // [nz]   exp(2 pi i m / nz)
    const double* sinz;     // [nz+1] sin(pi j / nz)
    const double* cosz;     // [nz+1] cos(pi j / nz)
};

// shared-memory layout behind the `nbuf` column buffers (doubles)
struct GScr {
    int nz, lc;             // lc = nz + 1: column stride (no swizzle)
    double2* S;             // [2 nz] transform scratch (the transformed buffer itself is the ping-pong partner)
    double* phim;           // [2][lc] per sy
    double* phip;           // [2][lc]
    double* edge;           // [8]  rows 0 and nz of the four columns across an in-place transform
    double* sums;           // [4]  DCT: X_1 of the four columns
    double* keep;           // [24] block-wide scalars parked between stages
};
inline size_t gen_col_smem_bytes(int nbuf, int nz) {
    return (size_t)(nbuf * 4 * (nz + 1) + 4 * nz + 4 * (nz + 1) + 8 + 4 + 24 + 4) * sizeof(double);
}
__device__ __forceinline__ GScr gscr_make(double* base, int nz) {
    GScr s;
    s.nz = nz; s.lc = nz + 1;
    s.S = reinterpret_cast<double2*>(base);            // base is 16-byte aligned: 4 * lc doubles per buffer
    s.phim = base + 4 * nz;
    s.phip = s.phim + 2 * s.lc;
    s.edge = s.phip + 2 * s.lc;
    s.sums = s.edge + 8;
    s.keep = s.sums + 4;
    return s;
}

__device__ __forceinline__ Row4 grow_load_s(const double* buf, int lc, int z) {
    Row4 x;
#pragma unroll
    for (int s = 0; s < 4; ++s) x.v[s] = buf[s * lc + z];
    return x;
}
__device__ __forceinline__ void grow_store_s(double* buf, int lc, int z, const Row4& x) {
#pragma unroll
    for (int s = 0; s < 4; ++s) buf[s * lc + z] = x.v[s];
}

// phim, phip of every row for sy = 0, 1 (inversion_utils.f90:505-519)
__device__ __forceinline__ void gphi_fill(const GScr& sc, const SpecGeom& g, const Grp& r) {
    const bool same = (r.k2[0] == r.k2[1]) && !r.g00;
    for (int sy = 0; sy < 2; ++sy) {
        const Hyp h = make_hyp(g, r, sy);
        for (int z = threadIdx.x; z <= sc.nz; z += blockDim.x) {
            double pm, pp;
            if (sy == 1 && same) {
                pm = sc.phim[z]; pp = sc.phip[z];
            } else if (h.lin) {
                pm = __ldg(&g.zm[z]) / g.Lz;
                pp = __ldg(&g.zp[z]) / g.Lz;
            } else {
                const double ep = exp(-(h.kl * __ldg(&g.zp[z])));
                const double em = exp(-(h.kl * __ldg(&g.zm[z])));
                pm = h.div * (ep - h.ef * em);
                pp = h.div * (em - h.ef * ep);
            }
            sc.phim[sy * sc.lc + z] = pm;
            sc.phip[sy * sc.lc + z] = pp;
        }
    }
}

// DST-I (rows 1..nz-1; rows 0 and nz keep their values) or DCT-I (rows 0..nz) of the four columns of X, scaled
// sqrt(2/nz): the algorithm of xform2 in zcol.cuh with loops.  The columns become the second FFT buffer once the
// pre-processed sequences sit in the scratch.  Needs a barrier between the caller's last write of X and the call;
// ends with a barrier.  Block of GEN_COL_THREADS threads (one warp per column in the reductions and scans).
constexpr int GEN_COL_THREADS = 256;

__device__ __forceinline__ void xform_gen(double* X, int kind, const GScr& sc, const GenZ& gz) {
    const int n = sc.nz, lc = sc.lc, t = threadIdx.x;
    const int warp = t >> 5, lane = t & 31;
    const int col = warp & 3;                          // warps 4..7 shadow 0..3 (uniform shuffles), results unused
    const bool own = warp < 4;
    const double* xc = X + col * lc;
    // DCT: X_1 = x_0/2 - x_n/2 + sum_j x_j cos(j pi/n) (stafft.f90:440-448), one warp per column
    {
        double sum = 0.0;
        if (kind == XF_DCT) {
            for (int j = 1 + lane; j < n; j += 32) sum += xc[j] * __ldg(&gz.cosz[j]);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d, 32);
        if (own && lane == 0) sc.sums[col] = sum + 0.5 * (xc[0] - xc[n]);
    }
    if (t < 8) sc.edge[t] = X[(t >> 1) * lc + ((t & 1) ? n : 0)];
    // pre-process: pair f = columns (2f, 2f+1) -> complex sequence f
    for (int w = t; w < 2 * n; w += blockDim.x) {
        const int f = w & 1, j = w >> 1;
        const double* xa = X + (2 * f) * lc;
        const double* xb = xa + lc;
        double yr = 0.0, yi = 0.0;
        if (j == 0) {
            if (kind == XF_DCT) { yr = 0.5 * (xa[0] + xa[n]); yi = 0.5 * (xb[0] + xb[n]); }
        } else {
            const double aj = xa[j], an = xa[n - j], bj = xb[j], bn = xb[n - j];
            const double sn = __ldg(&gz.sinz[j]);
            if (kind == XF_DST) {
                yr = 0.5 * (aj - an) + sn * (aj + an);
                yi = 0.5 * (bj - bn) + sn * (bj + bn);
            } else {
                yr = 0.5 * (aj + an) - sn * (aj - an);
                yi = 0.5 * (bj + bn) - sn * (bj - bn);
            }
        }
        sc.S[w] = make_double2(yr, yi);
    }
    __syncthreads();
    double2* XB = reinterpret_cast<double2*>(X);
    const double2* Y = gen_cfft<false>(sc.S, XB, 2, gz.plan, gz.tw);
    if (Y != sc.S) {
        // the result must not alias the rows that are written next
        for (int w = t; w < 2 * n; w += blockDim.x) sc.S[w] = XB[w];
        __syncthreads();
        Y = sc.S;
    }
    // post-process, one warp per column c = 2f + hh: spectrum of the real sequence hh of pair f
    //   A_k = (Y_k + conj Y_{n-k})/2 (hh = 0),  B_k = (Y_k - conj Y_{n-k})/(2i) (hh = 1)
    //   DST: X_1 = re_0/2, X_{2k} = -im_k, X_{2k+1} = X_{2k-1} + re_k;  DCT: X_0 = re_0, X_1 = sums, X_{2k} = re_k,
    //   X_{2k+1} = X_{2k-1} - im_k, X_n = re_{n/2}
    {
        const int f = col >> 1, hh = col & 1;
        double* xo = X + col * lc;
        const int K = n / 2 - 1;                       // increments k = 1..K
        const int ch = (K + 31) / 32;
        const int k0 = 1 + lane * ch, k1 = (k0 + ch - 1 < K) ? k0 + ch - 1 : K;
        double part = 0.0;
        for (int k = k0; k <= k1; ++k) {
            const double2 yk = Y[k * 2 + f], ym = Y[(n - k) * 2 + f];
            const double re = hh ? 0.5 * (yk.y + ym.y) : 0.5 * (yk.x + ym.x);
            const double im = hh ? 0.5 * (ym.x - yk.x) : 0.5 * (yk.y - ym.y);
            part += (kind == XF_DST) ? re : -im;
        }
        double incl = part;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double up = __shfl_up_sync(0xffffffffu, incl, d, 32);
            if (lane >= d) incl += up;
        }
        const double2 y0 = Y[f];
        const double re0 = hh ? y0.y : y0.x;
        const double scl = sqrt(2.0 / (double)n);
        double run = ((kind == XF_DST) ? 0.5 * re0 : sc.sums[col]) + (incl - part);     // X_{2 k0 - 1}
        if (own) {
            if (lane == 0) {
                xo[1] = scl * run;
                if (kind == XF_DCT) {
                    const double2 yh = Y[(n / 2) * 2 + f];
                    xo[0] = scl * re0;
                    xo[n] = scl * (hh ? yh.y : yh.x);
                } else {
                    xo[0] = sc.edge[2 * col]; xo[n] = sc.edge[2 * col + 1];
                }
            }
            for (int k = k0; k <= k1; ++k) {
                const double2 yk = Y[k * 2 + f], ym = Y[(n - k) * 2 + f];
                const double re = hh ? 0.5 * (yk.y + ym.y) : 0.5 * (yk.x + ym.x);
                const double im = hh ? 0.5 * (ym.x - yk.x) : 0.5 * (yk.y - ym.y);
                run += (kind == XF_DST) ? re : -im;
                xo[2 * k] = scl * ((kind == XF_DST) ? -im : re);
                xo[2 * k + 1] = scl * run;
            }
        }
    }
    __syncthreads();
}

// ---- operator mode ------------------------------------------------------------
__global__ void __launch_bounds__(GEN_COL_THREADS) k_zop_gen(SpecGeom g, GenZ gz, int op, const double* __restrict__ in,
                                                             double* __restrict__ out) {
    PS_SMEM(double, sm);
    const int nz = g.nz, lc = nz + 1;
    double* X = sm;
    const GScr sc = gscr_make(X + 4 * lc, nz);
    const Grp r = make_grp<true>(g, blockIdx.x);
    if (op == ZOP_DIFFX || op == ZOP_DIFFY) {
        for (int z = threadIdx.x; z <= nz; z += blockDim.x) {
            const Row4 x = row_load_g<0>(in, r, z);
            row_store_g<0>(out, r, z, op == ZOP_DIFFX ? ddx(x, r) : ddy(x, r));
        }
        return;
    }
    if (op == ZOP_COMBINE || op == ZOP_DECOMPOSE || op == ZOP_DIFFZ_SPEC) gphi_fill(sc, g, r);
    for (int z = threadIdx.x; z <= nz; z += blockDim.x) grow_store_s(X, lc, z, row_load_g<0>(in, r, z));
    __syncthreads();
    if (op == ZOP_DIFFZ_SPEC) {
        // diffz (inversion_utils.f90:683-719), see k_zop: ds = fs(0) dphim + fs(nz) dphip + cosine(rkz fs), then decompose
        double f0[4], fn[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) { f0[s] = X[s * lc]; fn[s] = X[s * lc + nz]; }
        __syncthreads();
        for (int z = threadIdx.x; z <= nz; z += blockDim.x) {
            const double rk = (z >= 1 && z < nz) ? __ldg(&g.rkz[z]) : 0.0;
#pragma unroll
            for (int s = 0; s < 4; ++s) X[s * lc + z] *= rk;
        }
        __syncthreads();
        xform_gen(X, XF_DCT, sc, gz);
        Hyp h[2];
        h[0] = make_hyp(g, r, 0); h[1] = make_hyp(g, r, 1);
        for (int z = threadIdx.x; z <= nz; z += blockDim.x) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                double dpm, dpp;
                hyp_dphi(h[s & 1], g.Lz, sc.phim[(s & 1) * lc + z], sc.phip[(s & 1) * lc + z], dpm, dpp);
                X[s * lc + z] += f0[s] * dpm + fn[s] * dpp;
            }
        }
        __syncthreads();
        double d0[4], dn[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) { d0[s] = X[s * lc]; dn[s] = X[s * lc + nz]; }
        __syncthreads();
        for (int z = 1 + threadIdx.x; z < nz; z += blockDim.x) {
#pragma unroll
            for (int s = 0; s < 4; ++s) X[s * lc + z] -= d0[s] * sc.phim[(s & 1) * lc + z] + dn[s] * sc.phip[(s & 1) * lc + z];
        }
        __syncthreads();
        xform_gen(X, XF_DST, sc, gz);
        for (int z = threadIdx.x; z <= nz; z += blockDim.x) row_store_g<0>(out, r, z, grow_load_s(X, lc, z));
        return;
    }
    if (op == ZOP_DIFFZ) {
        for (int z = threadIdx.x; z <= nz; z += blockDim.x) {
            Row4 d;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const double* c = X + s * lc;
                d.v[s] = (z == 0) ? g.dzi * (c[1] - c[0]) : (z == nz) ? g.dzi * (c[nz] - c[nz - 1]) : (c[z + 1] - c[z - 1]) * g.hdzi;
            }
            row_store_g<0>(out, r, z, d);
        }
        return;
    }
    if (op == ZOP_DECOMPOSE) {
        // subtract the harmonic part (inversion_utils.f90:571)
        for (int z = 1 + threadIdx.x; z < nz; z += blockDim.x) {
#pragma unroll
            for (int s = 0; s < 4; ++s)
                X[s * lc + z] -= X[s * lc] * sc.phim[(s & 1) * lc + z] + X[s * lc + nz] * sc.phip[(s & 1) * lc + z];
        }
        __syncthreads();
    }
    xform_gen(X, (op == ZOP_COSINE || op == ZOP_POISSON) ? XF_DCT : XF_DST, sc, gz);
    if (op == ZOP_POISSON) {
        // fields_derived.f90:125-148, green: inversion_utils.f90:283-288
        for (int z = threadIdx.x; z <= nz; z += blockDim.x) {
            const double rk = __ldg(&g.rkz[z]);
#pragma unroll
            for (int s = 0; s < 4; ++s) X[s * lc + z] *= (z == 0) ? -r.k2i[s & 1] : -1.0 / (r.k2[s & 1] + rk * rk);
        }
        __syncthreads();
        xform_gen(X, XF_DCT, sc, gz);
    }
    for (int z = threadIdx.x; z <= nz; z += blockDim.x) {
        Row4 x = grow_load_s(X, lc, z);
        if (op == ZOP_SINE && z == nz) { x.v[0] = x.v[1] = x.v[2] = x.v[3] = 0.0; }     // stafft.f90:546-549
        if (op == ZOP_COMBINE && z >= 1 && z < nz) {
#pragma unroll
            for (int s = 0; s < 4; ++s) x.v[s] += X[s * lc] * sc.phim[(s & 1) * lc + z] + X[s * lc + nz] * sc.phip[(s & 1) * lc + z];
        }
        row_store_g<0>(out, r, z, x);
    }
}

// ---- vor2vel, spectral part (inversion.f90:23-226 minus the fftxys2p calls; see k_vor2vel_spec) ----------------
__global__ void __launch_bounds__(GEN_COL_THREADS) k_vor2vel_gen(SpecGeom g, GenZ gz, V2VArgs a) {
    PS_SMEM(double, sm);
    const int nz = g.nz, lc = nz + 1, BUF = 4 * lc;
    double* A = sm;
    double* B = A + BUF;
    double* C = B + BUF;
    double* E = C + BUF;
    const GScr sc = gscr_make(E + BUF, nz);
    const Grp r = make_grp<true>(g, blockIdx.x);
    gphi_fill(sc, g, r);
    for (int z = threadIdx.x; z <= nz; z += blockDim.x) {
        grow_store_s(A, lc, z, row_load_g<0>(a.svor0, r, z));
        grow_store_s(B, lc, z, row_load_g<0>(a.svor1, r, z));
        grow_store_s(C, lc, z, row_load_g<0>(a.svor2, r, z));
    }
    __syncthreads();
    // semi-spectral zeta (:45, :142-144) and the sine sums of xi, eta
    xform_gen(C, XF_DST, sc, gz);
    xform_gen(A, XF_DST, sc, gz);
    xform_gen(B, XF_DST, sc, gz);
    for (int z = threadIdx.x; z <= nz; z += blockDim.x) {
        Row4 c = grow_load_s(C, lc, z);
        if (z >= 1 && z < nz) {
#pragma unroll
            for (int s = 0; s < 4; ++s) c.v[s] += C[s * lc] * sc.phim[(s & 1) * lc + z] + C[s * lc + nz] * sc.phip[(s & 1) * lc + z];
        }
        row_store_g<0>(a.wsem2, r, z, c);
        // the interior rows of C are updated after every thread has read the boundary rows: they are not modified
        if (z >= 1 && z < nz) grow_store_s(C, lc, z, c);
    }
    __syncthreads();
    // E = decompose(central_diffz(C)) (:46-47)
    for (int z = threadIdx.x; z <= nz; z += blockDim.x) {
        Row4 e;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const double* c = C + s * lc;
            const double e0 = g.dzi * (c[1] - c[0]), en = g.dzi * (c[nz] - c[nz - 1]);
            if (z == 0) e.v[s] = e0;
            else if (z == nz) e.v[s] = en;
            else e.v[s] = (c[z + 1] - c[z - 1]) * g.hdzi - (e0 * sc.phim[(s & 1) * lc + z] + en * sc.phip[(s & 1) * lc + z]);
        }
        grow_store_s(E, lc, z, e);
    }
    __syncthreads();
    xform_gen(E, XF_DST, sc, gz);
    // solenoidal projection of the semi-spectral and of the mixed-spectral rows (:39-76), D2 -> E (:86-90)
    for (int z = threadIdx.x; z <= nz; z += blockDim.x) {
        Row4 sa = grow_load_s(A, lc, z), sb = grow_load_s(B, lc, z), se;
        Row4 fa = row_load_g<0>(a.svor0, r, z), fb = row_load_g<0>(a.svor1, r, z);
        const Row4 fe = grow_load_s(E, lc, z);
        if (z >= 1 && z < nz) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const double* c = C + s * lc;
                const double pm = sc.phim[(s & 1) * lc + z], pp = sc.phip[(s & 1) * lc + z];
                sa.v[s] += A[s * lc] * pm + A[s * lc + nz] * pp;
                sb.v[s] += B[s * lc] * pm + B[s * lc + nz] * pp;
                se.v[s] = (c[z + 1] - c[z - 1]) * g.hdzi;
            }
        } else {
            se = fe;
        }
        project_row<true>(sa, sb, se, r);
        row_store_g<0>(a.wsem0, r, z, sa);
        row_store_g<0>(a.wsem1, r, z, sb);
        project_row<true>(fa, fb, fe, r);
        row_store_g<0>(a.svor0, r, z, fa);
        row_store_g<0>(a.svor1, r, z, fb);
        const Row4 ay2 = ddy(fa, r), bx2 = ddx(fb, r);
        Row4 d;
#pragma unroll
        for (int s = 0; s < 4; ++s) d.v[s] = ay2.v[s] - bx2.v[s];
        grow_store_s(E, lc, z, d);
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        const int s = threadIdx.x;
        sc.keep[s] = E[s * lc];
        sc.keep[4 + s] = E[s * lc + nz];
    }
    // invert the Laplacian (:108-122); mean flow of the (0,0) column (:150-165) -> B slots 0, 1
    for (int z = threadIdx.x; z <= nz; z += blockDim.x) {
        Row4 as, ds = grow_load_s(E, lc, z);
        if (z >= 1 && z < nz) {
            const double rk = __ldg(&g.rkz[z]);
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                ds.v[s] = -1.0 / (r.k2[s & 1] + rk * rk) * ds.v[s];
                as.v[s] = rk * ds.v[s];
            }
            grow_store_s(E, lc, z, ds);
        } else {
            as.v[0] = as.v[1] = as.v[2] = as.v[3] = 0.0;
        }
        grow_store_s(A, lc, z, as);
        if (r.g00) {
            Row4 m;
            m.v[0] = m.v[1] = m.v[2] = m.v[3] = 0.0;
            if (z >= 1 && z < nz) {
                const double rkzi = 1.0 / __ldg(&g.rkz[z]);
                m.v[0] = -rkzi * a.svor1[r.off[0] + z];
                m.v[1] = rkzi * a.svor0[r.off[0] + z];
            }
            grow_store_s(B, lc, z, m);
        }
    }
    __syncthreads();
    xform_gen(A, XF_DCT, sc, gz);      // (:128-129)
    xform_gen(E, XF_DST, sc, gz);
    if (r.g00) xform_gen(B, XF_DCT, sc, gz);
    // w, dw/dz, u, v (:96-104, :136-139, :169-213)
    Hyp h[2];
    h[0] = make_hyp(g, r, 0); h[1] = make_hyp(g, r, 1);
    double a00 = 0.0, a0n = 0.0, b00 = 0.0, b0n = 0.0;
    if (r.g00) {
        a00 = a.svor0[r.off[0]]; a0n = a.svor0[r.off[0] + nz];
        b00 = a.svor1[r.off[0]]; b0n = a.svor1[r.off[0] + nz];
    }
    for (int z = threadIdx.x; z <= nz; z += blockDim.x) {
        const Row4 as = grow_load_s(A, lc, z), ds = grow_load_s(E, lc, z), cs = grow_load_s(C, lc, z);
        const double zm = __ldg(&g.zm[z]), zp = __ldg(&g.zp[z]);
        Theta th[2];
        for (int sy = 0; sy < 2; ++sy) th[sy] = hyp_theta(h[sy], zm, zp, sc.phim[sy * lc + z], sc.phip[sy * lc + z]);
        Row4 es, w;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const Theta& q = th[s & 1];
            const double d0 = sc.keep[s], dn = sc.keep[4 + s];
            es.v[s] = d0 * q.dthm + dn * q.dthp + as.v[s];
            w.v[s] = (z == 0 || z == nz) ? 0.0 : ds.v[s] + d0 * q.thm + dn * q.thp;
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
            u.v[0] = B[z] + b0n * gt - b00 * gb;               // ubar (:163)
            v.v[0] = B[lc + z] - a0n * gt + a00 * gb;          // vbar (:164)
        }
        row_store_g<0>(a.svel0, r, z, u);
        row_store_g<0>(a.svel1, r, z, v);
        row_store_g<0>(a.svel2, r, z, w);
    }
}

// ---- vorticity tendency, spectral part (inversion.f90:298-371 after the fftxyp2s calls; see k_source_spec) -----
__global__ void __launch_bounds__(GEN_COL_THREADS) k_source_gen(SpecGeom g, GenZ gz, SrcArgs a) {
    PS_SMEM(double, sm);
    const int nz = g.nz, lc = nz + 1, BUF = 4 * lc;
    double* R = sm;           // r, then curl component 0
    double* Q = R + BUF;      // q, then curl component 2
    double* P = Q + BUF;      // p
    double* T = P + BUF;      // curl component 1
    const GScr sc = gscr_make(T + BUF, nz);
    const Grp r = make_grp<true>(g, blockIdx.x);
    gphi_fill(sc, g, r);
    for (int z = threadIdx.x; z <= nz; z += blockDim.x) {
        grow_store_s(R, lc, z, row_load_g<0>(a.r, r, z));
        grow_store_s(Q, lc, z, row_load_g<0>(a.q, r, z));
        grow_store_s(P, lc, z, row_load_g<0>(a.p, r, z));
    }
    __syncthreads();
    if (threadIdx.x == 0 || threadIdx.x == 32) {
        const bool top = (threadIdx.x != 0);
        const int z = top ? nz : 0, zl = top ? nz - 1 : 0, zh = top ? nz : 1;
        Row4 c0, c1;
        curl01_row(grow_load_s(R, lc, z), grow_load_s(Q, lc, zl), grow_load_s(Q, lc, zh), grow_load_s(P, lc, zl),
                   grow_load_s(P, lc, zh), g.dzi, r, c0, c1);
        const Row4 c2 = curl2_row(grow_load_s(Q, lc, z), grow_load_s(P, lc, z), r);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            sc.keep[(0 + top) * 4 + s] = c0.v[s];
            sc.keep[(2 + top) * 4 + s] = c1.v[s];
            sc.keep[(4 + top) * 4 + s] = c2.v[s];
        }
    }
    __syncthreads();
    for (int z = threadIdx.x; z <= nz; z += blockDim.x) {
        Row4 s0, s1;
        if (z == 0 || z == nz) {
            const int top = (z == nz);
#pragma unroll
            for (int s = 0; s < 4; ++s) { s0.v[s] = sc.keep[(0 + top) * 4 + s]; s1.v[s] = sc.keep[(2 + top) * 4 + s]; }
        } else {
            curl01_row(grow_load_s(R, lc, z), grow_load_s(Q, lc, z - 1), grow_load_s(Q, lc, z + 1),
                       grow_load_s(P, lc, z - 1), grow_load_s(P, lc, z + 1), g.hdzi, r, s0, s1);
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const double pm = sc.phim[(s & 1) * lc + z], pp = sc.phip[(s & 1) * lc + z];
                s0.v[s] -= sc.keep[0 + s] * pm + sc.keep[4 + s] * pp;
                s1.v[s] -= sc.keep[8 + s] * pm + sc.keep[12 + s] * pp;
            }
        }
        grow_store_s(R, lc, z, s0);
        grow_store_s(T, lc, z, s1);
    }
    __syncthreads();
    for (int z = threadIdx.x; z <= nz; z += blockDim.x) {
        Row4 s2 = curl2_row(grow_load_s(Q, lc, z), grow_load_s(P, lc, z), r);
        if (z >= 1 && z < nz) {
#pragma unroll
            for (int s = 0; s < 4; ++s)
                s2.v[s] -= sc.keep[16 + s] * sc.phim[(s & 1) * lc + z] + sc.keep[20 + s] * sc.phip[(s & 1) * lc + z];
        }
        grow_store_s(Q, lc, z, s2);
    }
    __syncthreads();
    xform_gen(R, XF_DST, sc, gz);
    xform_gen(T, XF_DST, sc, gz);
    xform_gen(Q, XF_DST, sc, gz);
    for (int z = threadIdx.x; z <= nz; z += blockDim.x) {
        row_store_g<0>(a.s0, r, z, grow_load_s(R, lc, z));
        row_store_g<0>(a.s1, r, z, grow_load_s(T, lc, z));
        row_store_g<0>(a.s2, r, z, grow_load_s(Q, lc, z));
    }
}

}  // namespace ps3d
