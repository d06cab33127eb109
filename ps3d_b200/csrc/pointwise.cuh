// Streaming kernels: layout repack at the host boundary, the time-stepper
// updates (reference src/stepper/cn2.f90:92-181, impl_rk4.f90:76-364, using
// the identity decompose(c(ky,kx) * combine(q)) == c*q so that no z transform
// is needed), the mean-vorticity fix (field_diagnostics.f90:584-619), the
// adapt()/diagnostic reductions (advance.f90:171-313,
// field_diagnostics.f90:85-206,405-579) and the per-point strain eigenvalue
// sweep (advance.f90:222-276, utils/jacobi.f90:19-51,215-305).
//
// All reductions are two-stage with a fixed grid and fixed tree, i.e.
// deterministic and independent of scheduling (no floating-point atomics).
#pragma once

#include "rt.h"

namespace ps3d {

// ---- host-boundary repack ---------------------------------------------------
// natural Fortran order  f(0:nz, y, x)  <->  internal [x][rowmap(y)][pz]
__global__ void k_repack_in(const double* __restrict__ nat, double* __restrict__ dev, int nxl, int ny, int nzp,
                            int pz, const int* __restrict__ rowmap) {
    const long long n = (long long)nxl * ny * pz;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int z = (int)(i % pz);
        const long long xy = i / pz;
        const int y = (int)(xy % ny);
        const long long x = xy / ny;
        const int yy = rowmap ? rowmap[y] : y;
        dev[(x * ny + yy) * pz + z] = (z < nzp) ? nat[(x * ny + y) * nzp + z] : 0.0;
    }
}

__global__ void k_repack_out(const double* __restrict__ dev, double* __restrict__ nat, int nxl, int ny, int nzp,
                             int pz, const int* __restrict__ rowmap) {
    const long long n = (long long)nxl * ny * nzp;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int z = (int)(i % nzp);
        const long long xy = i / nzp;
        const int y = (int)(xy % ny);
        const long long x = xy / ny;
        const int yy = rowmap ? rowmap[y] : y;
        nat[i] = dev[(x * ny + yy) * pz + z];
    }
}

// ---- time steppers ------------------------------------------------------------
struct StepArgs {
    double* svor[3];
    double* svorts[3];
    double* wa[3];          // cn2: vortsm      rk4: svori
    double* wb[3];          // rk4: svorf
    const double* f2d;      // cn2: vdiss*filt2d (per column)   rk4: unused
    const double* filtz;    // cn2: z part of the filter, [nz+1], 1 at rows 0 and nz
    const double* mq;       // rk4: emq (per column)
    const double* pq;       // rk4: epq or filt(0,:,:) (per column)
    const double* vd;       // cn2: vdiss (per column) — used for the (0,0) column where filt = 1
    double c1, c2;          // cn2: dt/2.   rk4: stage coefficients
    int stage;              // cn2: 0 = first update (defines vortsm), 1 = iteration. rk4: 1..4
    long long ncol;         // nx*nyl
    int nz, pz;
    int has00;
    int ncomp = 3;          // components updated: 3 (vorticity) or 1 (buoyancy: slot 0 of the pointer arrays)
};

// cn2.f90:120-135 / :162-173 with combine -> vdiss -> decompose collapsed.
__global__ void k_cn2_update(StepArgs a) {
    const long long n = a.ncol * a.pz;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int z = (int)(i % a.pz);
        if (z > a.nz) continue;
        const long long col = i / a.pz;
        double fac = __ldg(&a.f2d[col]) * __ldg(&a.filtz[z]);
        if (a.has00 && col == 0) fac = __ldg(&a.vd[0]);      // filt(:,0,0) = 1 (inversion_utils.f90:275-277)
        for (int c = 0; c < a.ncomp; ++c) {
            const double s = a.svorts[c][i];
            double sm;
            if (a.stage == 0) { sm = a.svor[c][i] + a.c1 * s; a.wa[c][i] = sm; }
            else sm = a.wa[c][i];
            a.svor[c][i] = fac * (sm + a.c1 * s);
        }
    }
}

// impl_rk4.f90:212-364, substeps one..four, combine/decompose pairs collapsed.
__global__ void k_rk4_update(StepArgs a) {
    const long long n = a.ncol * a.pz;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int z = (int)(i % a.pz);
        if (z > a.nz) continue;
        const long long col = i / a.pz;
        const double mq = __ldg(&a.mq[col]), pq = __ldg(&a.pq[col]);
        for (int c = 0; c < a.ncomp; ++c) {
            const double s = pq * a.svorts[c][i];
            a.svorts[c][i] = s;
            if (a.stage == 1) {
                const double qi = a.svor[c][i];
                a.wa[c][i] = qi;
                a.svor[c][i] = mq * (qi + a.c1 * s);
                a.wb[c][i] = qi + a.c2 * s;
            } else if (a.stage == 4) {
                a.svor[c][i] = mq * (a.wb[c][i] + a.c1 * s);
            } else {
                a.svor[c][i] = mq * (a.wa[c][i] + a.c1 * s);
                a.wb[c][i] = a.wb[c][i] + a.c2 * s;
            }
        }
    }
}

// per-column factor tables for the steppers
//   mode 0 (cn2.f90:59):       o1 = 1/(1 + dfac*vhdis),  o2 = o1 * filt2d
//   mode 1 (impl_rk4.f90:43-47,87-89): e = exp(dfac*vhdis); o1 = 1/e (emq), o2 = e*filt2d (epq)
//   mode 2: o1 = o1^2 (emq**2, :151)     mode 3: o2 = o2^2 (epq**2, :185)
__global__ void k_step_factors(int mode, double dfac, const double* __restrict__ vhdis, const double* __restrict__ filt2d,
                               double* __restrict__ o1, double* __restrict__ o2, long long ncol) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ncol; i += (long long)gridDim.x * blockDim.x) {
        if (mode == 0) {
            const double v = 1.0 / (1.0 + dfac * vhdis[i]);
            o1[i] = v; o2[i] = v * filt2d[i];
        } else if (mode == 1) {
            const double e = exp(dfac * vhdis[i]);
            o1[i] = 1.0 / e; o2[i] = e * filt2d[i];
        } else if (mode == 2) {
            o1[i] = o1[i] * o1[i];
        } else {
            o2[i] = o2[i] * o2[i];
        }
    }
}

// ---- block reduction helper -----------------------------------------------------
// op 0 = sum, 1 = max.  All threads call; result valid in thread 0.  `red` has blockDim.x doubles.
__device__ __forceinline__ double block_reduce(double v, int op, double* red) {
    const int t = threadIdx.x;
    __syncthreads();
    red[t] = v;
    __syncthreads();
    for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
        if (t < s) red[t] = op ? fmax(red[t], red[t + s]) : red[t] + red[t + s];
        __syncthreads();
    }
    return red[0];
}

// final stage: out[q] = reduce_{b} partial[b*nq + q] in a fixed order
__global__ void k_reduce_final(const double* __restrict__ partial, int nblk, int nq, unsigned opmask,
                               double* __restrict__ out) {
    PS_SMEM(double, red);
    for (int q = 0; q < nq; ++q) {
        const int op = (opmask >> q) & 1;
        double v = op ? -1.0e300 : 0.0;
        for (int b = threadIdx.x; b < nblk; b += blockDim.x) {
            const double p = partial[(long long)b * nq + q];
            v = op ? fmax(v, p) : v + p;
        }
        v = block_reduce(v, op, red);
        if (threadIdx.x == 0) out[q] = v;
    }
}

// ---- mean vorticity (field_diagnostics.f90:584-619) -------------------------------
// savg = (svor(0)+svor(nz))/2 + 1/nz * sum_k dst(svor)(k).  The sum over the sine
// transform is the dot product with wz[j] = sqrt(2/nz) sum_k sin(pi j k/nz)
// (= sqrt(2/nz) cot(pi j/(2 nz)) for odd j, 0 for even j), precomputed on the host.
// mode 0: ini_mean = savg.  mode 1: svor(0), svor(nz) += ini_mean - savg.
__global__ void k_vor_mean(double* svor0, double* svor1, const double* __restrict__ wz, int nz, double fnzi,
                           double* ini_mean, int mode) {
    PS_SMEM(double, red);
    for (int c = 0; c < 2; ++c) {
        double* col = c ? svor1 : svor0;
        double v = 0.0;
        for (int j = 1 + threadIdx.x; j < nz; j += blockDim.x) v += __ldg(&wz[j]) * col[j];
        v = block_reduce(v, 0, red);
        if (threadIdx.x == 0) {
            const double savg = 0.5 * (col[0] + col[nz]) + fnzi * v;
            if (mode == 0) ini_mean[c] = savg;
            else { const double d = ini_mean[c] - savg; col[0] += d; col[nz] += d; }
        }
        __syncthreads();
    }
}

// ---- field reductions ------------------------------------------------------------
enum { RQ_MAXW2 = 0, RQ_SUMW2, RQ_SUMW0, RQ_SUMW1, RQ_SUMW2C, RQ_MAXU, RQ_MAXV, RQ_MAXWV, RQ_SUMU2, RQ_SUMUW,
       RQ_SUMUH, RQ_SUMWH, RQ_MAXWH, RQ_N };
constexpr unsigned RQ_OPMASK = (1u << RQ_MAXW2) | (1u << RQ_MAXU) | (1u << RQ_MAXV) | (1u << RQ_MAXWV) | (1u << RQ_MAXWH);

struct FieldPtrs { const double* vor[3]; const double* vel[3]; };

// trapezoid-weighted sums / maxima over the physical fields
__global__ void k_field_reduce(FieldPtrs f, long long ncol, int nz, int pz, double* __restrict__ partial) {
    PS_SMEM(double, red);
    double acc[RQ_N];
#pragma unroll
    for (int q = 0; q < RQ_N; ++q) acc[q] = ((RQ_OPMASK >> q) & 1) ? -1.0e300 : 0.0;
    // two consecutive z per thread (16-byte loads; pz is even and every column starts 16-byte aligned), 32-bit
    // index arithmetic
    const unsigned hp = (unsigned)pz >> 1;
    const unsigned long long n2 = (unsigned long long)ncol * hp;
    for (unsigned long long i2 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i2 < n2;
         i2 += (unsigned long long)gridDim.x * blockDim.x) {
        const int z0 = 2 * (int)((n2 >> 32) ? (i2 % hp) : ((unsigned)i2 % hp));
        if (z0 > nz) continue;
        const double2 a2 = reinterpret_cast<const double2*>(f.vor[0])[i2], b2 = reinterpret_cast<const double2*>(f.vor[1])[i2];
        const double2 c2 = reinterpret_cast<const double2*>(f.vor[2])[i2], u2 = reinterpret_cast<const double2*>(f.vel[0])[i2];
        const double2 v2 = reinterpret_cast<const double2*>(f.vel[1])[i2], w2_ = reinterpret_cast<const double2*>(f.vel[2])[i2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int z = z0 + h;
            if (z > nz) continue;
            const double w = (z == 0 || z == nz) ? 0.5 : 1.0;
            const double a = h ? a2.y : a2.x, b = h ? b2.y : b2.x, c = h ? c2.y : c2.x;
            const double u = h ? u2.y : u2.x, v = h ? v2.y : v2.x, ww = h ? w2_.y : w2_.x;
            const double w2 = a * a + b * b + c * c;
            acc[RQ_MAXW2] = fmax(acc[RQ_MAXW2], fabs(w2));
            acc[RQ_SUMW2] += w * w2;
            acc[RQ_SUMW0] += w * a; acc[RQ_SUMW1] += w * b; acc[RQ_SUMW2C] += w * c;
            acc[RQ_MAXU] = fmax(acc[RQ_MAXU], u); acc[RQ_MAXV] = fmax(acc[RQ_MAXV], v); acc[RQ_MAXWV] = fmax(acc[RQ_MAXWV], ww);
            acc[RQ_SUMU2] += w * (u * u + v * v + ww * ww);
            acc[RQ_SUMUW] += w * (u * a + v * b + ww * c);
            acc[RQ_SUMUH] += w * (u * u + v * v);                          // field_diagnostics.f90:135-142
            acc[RQ_SUMWH] += w * (a * a + b * b);                          // :215-222
            acc[RQ_MAXWH] = fmax(acc[RQ_MAXWH], a * a + b * b);           // :237 (sqrt taken on the host)
        }
    }
    for (int q = 0; q < RQ_N; ++q) {
        const double r = block_reduce(acc[q], (RQ_OPMASK >> q) & 1, red);
        if (threadIdx.x == 0) partial[(long long)blockIdx.x * RQ_N + q] = r;
    }
}

// ---- the remaining scalars of the field-statistics file (field_diagnostics_netcdf.f90:257-439) ------------
// minima / maxima of the vorticity components (:356-358, :392-394), their maxima on the upper and lower
// surface (:398-403), max horizontal speed^2 on the upper surface (:404-405), and the sums behind the rms of the
// upper-surface zeta and horizontal divergence (:301, :305).  Minima are carried as maxima of the negated value
// so that one max/sum op mask serves the two-stage reduction.
enum { SQ_NMIN0 = 0, SQ_NMIN1, SQ_NMIN2, SQ_MAX0, SQ_MAX1, SQ_MAX2, SQ_US0, SQ_US1, SQ_US2, SQ_LS0, SQ_LS1, SQ_LS2,
       SQ_USUH2, SQ_USZ2, SQ_USDEL2, SQ_N };
constexpr unsigned SQ_OPMASK = (1u << SQ_USZ2 | 1u << SQ_USDEL2) ^ ((1u << SQ_N) - 1u);    // 1 = max, 0 = sum

__global__ void k_field_stats(FieldPtrs f, const double* __restrict__ delta, long long ncol, int nz, int pz,
                              double* __restrict__ partial) {
    PS_SMEM(double, red);
    double acc[SQ_N];
#pragma unroll
    for (int q = 0; q < SQ_N; ++q) acc[q] = ((SQ_OPMASK >> q) & 1) ? -1.0e300 : 0.0;
    const long long n = ncol * pz;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int z = (int)(i % pz);
        if (z > nz) continue;
        const double a = f.vor[0][i], b = f.vor[1][i], c = f.vor[2][i];
        acc[SQ_NMIN0] = fmax(acc[SQ_NMIN0], -a); acc[SQ_NMIN1] = fmax(acc[SQ_NMIN1], -b); acc[SQ_NMIN2] = fmax(acc[SQ_NMIN2], -c);
        acc[SQ_MAX0] = fmax(acc[SQ_MAX0], a); acc[SQ_MAX1] = fmax(acc[SQ_MAX1], b); acc[SQ_MAX2] = fmax(acc[SQ_MAX2], c);
        if (z == nz) {
            const double u = f.vel[0][i], v = f.vel[1][i], d = delta[i];
            acc[SQ_US0] = fmax(acc[SQ_US0], a); acc[SQ_US1] = fmax(acc[SQ_US1], b); acc[SQ_US2] = fmax(acc[SQ_US2], c);
            acc[SQ_USUH2] = fmax(acc[SQ_USUH2], u * u + v * v);
            acc[SQ_USZ2] += c * c;
            acc[SQ_USDEL2] += d * d;
        } else if (z == 0) {
            acc[SQ_LS0] = fmax(acc[SQ_LS0], a); acc[SQ_LS1] = fmax(acc[SQ_LS1], b); acc[SQ_LS2] = fmax(acc[SQ_LS2], c);
        }
    }
    for (int q = 0; q < SQ_N; ++q) {
        const double r = block_reduce(acc[q], (SQ_OPMASK >> q) & 1, red);
        if (threadIdx.x == 0) partial[(long long)blockIdx.x * SQ_N + q] = r;
    }
}

// ---- kinetic-energy spectrum (genspec.f90:64-106) ---------------------------------------------------------
// u, v (cosine series in z) and w (sine series) fully spectral; bin m = int(nint(|k|) / dk) collects
// |u|^2 + |v|^2 + |w|^2 and a count.  k2l2 is the [nx/2+1][nyl] table of rkx^2 + rky^2 (rkx(kx) = rkx(nx-kx)).
__global__ void k_spec_bin(const double* __restrict__ u, const double* __restrict__ v, const double* __restrict__ w,
                           const double* __restrict__ k2l2, const double* __restrict__ rkz, int nx, int nyl, int nz,
                           int pz, double dki, int nb, double* spec, double* num) {
    const long long n = (long long)nx * nyl * pz;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int z = (int)(i % pz);
        if (z > nz) continue;
        const long long col = i / pz;
        const int kx = (int)(col / nyl), kyl = (int)(col - (long long)kx * nyl);
        const int b = (kx <= nx - kx) ? kx : nx - kx;
        const double rk = __ldg(&rkz[z]);
        const double kmag = floor(sqrt(__ldg(&k2l2[(long long)b * nyl + kyl]) + rk * rk) + 0.5);    // nint (:79)
        const int m = (int)(kmag * dki);                                                            // :100
        if (m >= nb) continue;              // cannot happen with the host's bin count; never write out of bounds
        const double a = u[i], bb = v[i], c = w[i];
        atomicAdd(&spec[m], a * a + bb * bb + c * c);
        atomicAdd(&num[m], 1.0);
    }
}

// right-hand side of the pressure Poisson equation (fields_derived.f90:97-113) from the five strain fields and omega
struct StrainPtrsFwd;
__global__ void k_pressure_rhs(const double* __restrict__ dudx, const double* __restrict__ dudy, const double* __restrict__ dvdy,
                               const double* __restrict__ dwdx, const double* __restrict__ dwdy, const double* __restrict__ xi,
                               const double* __restrict__ eta, const double* __restrict__ zeta, double* __restrict__ out,
                               long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double ux = dudx[i], uy = dudy[i], vy = dvdy[i], wx = dwdx[i], wy = dwdy[i];
        const double wz = -(ux + vy);
        out[i] = 2.0 * (ux * vy - uy * (zeta[i] + uy) + vy * wz - wy * (wy - xi[i]) + wz * ux - wx * (wx + eta[i]));
    }
}

// ---- buoyancy build (ENABLE_BUOYANCY): flux products, tendency assembly, buoyancy-frequency maximum ----
__global__ void k_mul(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = a[i] * b[i];
}
// sbuoys = -(d1 + d2 + d3) - bfsq * w   (inversion.f90:248-290 in mixed-spectral space, see do_buoyancy_tendency)
__global__ void k_btend(const double* __restrict__ d1, const double* __restrict__ d2, const double* __restrict__ d3,
                        const double* __restrict__ w, double bfsq, double* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = -d1[i] - d2[i] - bfsq * w[i] - d3[i];
}
// partial[b] = max over the grid of (db/dx)^2 + (db/dy)^2 + (db/dz + bfsq)^2   (advance.f90:160-165)
__global__ void k_bfmax(const double* __restrict__ xp, const double* __restrict__ yp, const double* __restrict__ zp,
                        double bfsq, long long ncol, int nz, int pz, double* __restrict__ partial) {
    PS_SMEM(double, red);
    double m = -1.0e300;
    const long long n = ncol * pz;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int z = (int)(i % pz);
        if (z > nz) continue;
        const double a = xp[i], b = yp[i], c = zp[i] + bfsq;
        m = fmax(m, a * a + b * b + c * c);
    }
    const double r = block_reduce(m, 1, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = r;
}

__global__ void k_add(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = a[i] + b[i];
}
// out = a + b + s * c   (pressure source of the buoyancy build, fields_derived.f90:108-112: pres + db/dz + f_cor(3) zeta)
__global__ void k_add_axpy(const double* __restrict__ a, const double* __restrict__ b, double s, const double* __restrict__ c,
                           double* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = a[i] + b[i] + s * c[i];
}

// ---- Jacobi eigenvalues of the symmetrised strain (jacobi.f90) -----------------
__device__ __forceinline__ void givens(double aij, double di, double dj, double& s, double& t, double& tau) {
    const double eps = 2.220446049250313e-16;
    const double g = 100.0 * fabs(aij);
    const double h = dj - di;
    if (fabs(h) + g == fabs(h)) {
        t = aij / (h + copysign(eps, h));
    } else {
        const double theta = 0.5 * h / (aij + copysign(eps, aij));
        t = 1.0 / (fabs(theta) + sqrt(1.0 + theta * theta));
        if (theta < 0.0) t = -t;
    }
    const double c = 1.0 / sqrt(1.0 + t * t);
    s = t * c;
    tau = s / (1.0 + c);
}

// returns max |eigenvalue| (advance.f90:261-264)
__device__ __forceinline__ double jacobi_max_abs(double d1, double a12, double a13, double d2, double a23, double d3) {
    double b1 = d1, b2 = d2, b3 = d3;
    double sm = fabs(a12) + fabs(a13) + fabs(a23);
    int guard = 0;
    while (sm > 1.0e-15 && guard < 100) {
        double z1 = 0.0, z2 = 0.0, z3 = 0.0, s, t, tau, h, g, hh;
        // (1,2)
        givens(a12, d1, d2, s, t, tau);
        h = t * a12; z1 -= h; z2 += h; d1 -= h; d2 += h; a12 = 0.0;
        g = a13; hh = a23; a13 = g - s * (hh + g * tau); a23 = hh + s * (g - hh * tau);
        // (1,3)
        givens(a13, d1, d3, s, t, tau);
        h = t * a13; z1 -= h; z3 += h; d1 -= h; d3 += h; a13 = 0.0;
        g = a12; hh = a23; a12 = g - s * (hh + g * tau); a23 = hh + s * (g - hh * tau);
        // (2,3)
        givens(a23, d2, d3, s, t, tau);
        h = t * a23; z2 -= h; z3 += h; d2 -= h; d3 += h; a23 = 0.0;
        g = a12; hh = a13; a12 = g - s * (hh + g * tau); a13 = hh + s * (g - hh * tau);
        b1 += z1; b2 += z2; b3 += z3;
        d1 = b1; d2 = b2; d3 = b3;
        sm = fabs(a12) + fabs(a13) + fabs(a23);
        ++guard;
    }
    return fmax(fmax(fabs(d1), fabs(d2)), fabs(d3));
}

// max |eigenvalue| of the symmetric, traceless strain matrix in closed form (trigonometric solution of the
// depressed cubic lambda^3 - (p2/2) lambda - det = 0).  The extreme eigenvalue is the well-conditioned one:
// where acos is ill-conditioned (|r| -> 1) the extreme root is stationary in the angle, so the result is
// accurate to a few ulp; it agrees with the reference's cyclic Jacobi sweeps (atol 1e-15, jacobi.f90:13)
// to round-off at ~1/10 of their cost.  `strict` in k_strain selects the literal Jacobi restatement.
__device__ __forceinline__ double sym3_max_abs(double a, double d, double e, double b, double f, double c) {
    const double p1 = d * d + e * e + f * f;
    const double p2 = a * a + b * b + c * c + 2.0 * p1;
    if (p2 == 0.0) return 0.0;
    const double p = sqrt(p2 * (1.0 / 6.0));
    const double pi_ = 1.0 / p;
    const double A = a * pi_, B = b * pi_, C = c * pi_, D = d * pi_, E = e * pi_, F = f * pi_;
    double r = 0.5 * (A * (B * C - F * F) - D * (D * C - E * F) + E * (D * F - B * E));
    r = fmin(1.0, fmax(-1.0, r));
    const double phi = acos(r) * (1.0 / 3.0);
    const double l1 = 2.0 * p * cos(phi);                                    // largest
    const double l3 = 2.0 * p * cos(phi + 2.0943951023931954923084289221863);  // smallest
    return fmax(fabs(l1), fabs(l3));
}

struct StrainPtrs { const double* dudx; const double* dudy; const double* dvdy; const double* dwdx; const double* dwdy;
                    const double* vor[3]; };

// One pass over the five strain fields and omega for the two remaining reductions of adapt:
//  * max |eigenvalue| of the symmetrised strain over all points / points with iz = nz / iz = 0
//    (advance.f90:222-276) -> partial[b*5 + {0,1,2}];
//  * get_char_vorticity (field_diagnostics.f90:501-545): sums over the cell-averaged |omega| of the cells with
//    sum |omega_bar| > vortrms -> partial[b*5 + {3,4}].  vortrms = sqrt(<|omega|^2>) (advance.f90:180-183) is taken
//    from the first reduction ON THE DEVICE (red[RQ_SUMW2], already reduced over the ranks), with the same two
//    correctly rounded operations the host would do: no host round trip between the reductions.
// (red == nullptr: strain only, as before.)
__global__ void k_strain(StrainPtrs f, long long ncol, int nz, int pz, int strict, const double* __restrict__ red,
                         double ncell, double* __restrict__ partial) {
    PS_SMEM(double, redbuf);
    double gg = 0.0, us = 0.0, ls = 0.0, l1 = 0.0, l2 = 0.0;
    const double vortrms = red ? sqrt(red[RQ_SUMW2] / ncell) : 0.0;
    // two consecutive z per thread (16-byte loads), 32-bit index arithmetic
    const unsigned hp = (unsigned)pz >> 1;
    const unsigned long long n2 = (unsigned long long)ncol * hp;
    for (unsigned long long i2 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i2 < n2;
         i2 += (unsigned long long)gridDim.x * blockDim.x) {
        const int z0 = 2 * (int)((n2 >> 32) ? (i2 % hp) : ((unsigned)i2 % hp));
        if (z0 > nz) continue;
        const double2 ux2 = reinterpret_cast<const double2*>(f.dudx)[i2], uy2 = reinterpret_cast<const double2*>(f.dudy)[i2];
        const double2 vy2 = reinterpret_cast<const double2*>(f.dvdy)[i2], wx2 = reinterpret_cast<const double2*>(f.dwdx)[i2];
        const double2 wy2 = reinterpret_cast<const double2*>(f.dwdy)[i2];
        const double2 p0 = reinterpret_cast<const double2*>(f.vor[0])[i2], p1 = reinterpret_cast<const double2*>(f.vor[1])[i2];
        const double2 p2 = reinterpret_cast<const double2*>(f.vor[2])[i2];
        // omega one level below the pair (cell averages of get_char_vorticity)
        double q0 = 0.0, q1 = 0.0, q2 = 0.0;
        if (red && z0 >= 1) { q0 = f.vor[0][2 * i2 - 1]; q1 = f.vor[1][2 * i2 - 1]; q2 = f.vor[2][2 * i2 - 1]; }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int z = z0 + h;
            if (z > nz) continue;
            const double ux = h ? ux2.y : ux2.x, uy = h ? uy2.y : uy2.x, vy = h ? vy2.y : vy2.x;
            const double wx = h ? wx2.y : wx2.x, wy = h ? wy2.y : wy2.x;
            const double o0 = h ? p0.y : p0.x, o1 = h ? p1.y : p1.x, o2 = h ? p2.y : p2.x;
            // advance.f90:252-257
            const double s12 = uy + 0.5 * o2, s13 = wx + 0.5 * o1, s23 = wy - 0.5 * o0;
            // Only the maximum over the grid is wanted: the matrix is symmetric and traceless, so lambda_max^2 <= 2/3 |S|_F^2,
            // and a point whose bound does not exceed this thread's running maximum cannot raise it.  The eigenvalues
            // are evaluated only where it can (and on the two surfaces, which have their own maxima): same result,
            // the pass becomes a streaming read.
            const double s33 = -(ux + vy);
            const double fro = ux * ux + vy * vy + s33 * s33 + 2.0 * (s12 * s12 + s13 * s13 + s23 * s23);
            const bool surf = (z == nz) || (z == 0);
            if (surf || fro * (2.0 / 3.0) * (1.0 + 1.0e-9) > gg * gg) {
                const double l = strict ? jacobi_max_abs(ux, s12, s13, vy, s23, s33)
                                        : sym3_max_abs(ux, s12, s13, vy, s23, s33);
                gg = fmax(gg, l);
                if (z == nz) us = fmax(us, l);
                if (z == 0) ls = fmax(ls, l);
            }
            if (red && z >= 1) {
                const double m0 = h ? p0.x : q0, m1 = h ? p1.x : q1, m2 = h ? p2.x : q2;
                const double v1 = 0.5 * fabs(m0 + o0);
                const double v2 = 0.5 * fabs(m1 + o1);
                const double v3 = 0.5 * fabs(m2 + o2);
                if (v1 + v2 + v3 > vortrms) {
                    l1 += v1 + v2 + v3;
                    l2 += v1 * v1 + v2 * v2 + v3 * v3;
                }
            }
        }
    }
    const double r0 = block_reduce(gg, 1, redbuf);
    const double r1 = block_reduce(us, 1, redbuf);
    const double r2 = block_reduce(ls, 1, redbuf);
    const double r3 = block_reduce(l1, 0, redbuf);
    const double r4 = block_reduce(l2, 0, redbuf);
    if (threadIdx.x == 0) {
        double* p = partial + (long long)blockIdx.x * 5;
        p[0] = r0; p[1] = r1; p[2] = r2; p[3] = r3; p[4] = r4;
    }
}

// after a device-side all-reduce of the same values as sums (s) and as maxima (m): out[q] = opmask bit q ? m[q] : s[q]
__global__ void k_select_reduced(const double* __restrict__ s, const double* __restrict__ m, int n, unsigned opmask,
                                 double* __restrict__ out) {
    const int q = threadIdx.x;
    if (q < n) out[q] = ((opmask >> q) & 1) ? m[q] : s[q];
}

}  // namespace ps3d
