// Mixed-radix complex FFT for lengths that are not a power of two (n = 2^a 3^b 5^c, the lengths the reference's
// factorisen accepts: stafft.f90:128-187 tries 6, 4, 2, 3, 5).  Stockham autosort, decimation in frequency, one
// pass per radix 4 / 2 / 3 / 5 (a 6 of the reference is a 2 and a 3 here), NB interleaved sequences per block,
// ping-pong between two shared-memory arrays, every thread of the block looping over (butterfly, sequence) work
// items.  This is the coverage path: the power-of-two lengths of the configurations in BASELINE.json run through
// the register-blocked radix-8 kernels of fft_core.cuh.
//
// Same pass formula as fft_core.cuh: butterfly b in [0, n/R), p = b / s, q = b % s, inputs x[b + k n/R],
// outputs (times W_n^(s p j)) at x'[q + s (R p + j)], s = product of the radices of the previous passes.
#pragma once

#include "rt.h"

namespace ps3d {

struct GenPlan {
    int n;
    int nfac;
    int radix[24];
};

// host: n = prod radix, radices 4 first (fewest passes), then 2, 3, 5; returns false if n has another prime factor
inline bool gen_plan_make(int n, GenPlan& p) {
    p.n = n; p.nfac = 0;
    int m = n;
    while (m % 4 == 0) { p.radix[p.nfac++] = 4; m /= 4; }
    while (m % 2 == 0) { p.radix[p.nfac++] = 2; m /= 2; }
    while (m % 3 == 0) { p.radix[p.nfac++] = 3; m /= 3; }
    while (m % 5 == 0) { p.radix[p.nfac++] = 5; m /= 5; }
    return m == 1 && n >= 2 && p.nfac <= 24;
}

// in-place DFT of length R in {2, 3, 4, 5} (forward: exp(-2 pi i jk/R))
template <bool INV>
__device__ __forceinline__ void small_dft(int R, double* xr, double* xi) {
    if (R == 2) {
        const double ar = xr[0], ai = xi[0];
        xr[0] = ar + xr[1]; xi[0] = ai + xi[1];
        xr[1] = ar - xr[1]; xi[1] = ai - xi[1];
    } else if (R == 4) {
        const double t0r = xr[0] + xr[2], t0i = xi[0] + xi[2], t1r = xr[0] - xr[2], t1i = xi[0] - xi[2];
        const double t2r = xr[1] + xr[3], t2i = xi[1] + xi[3], dr = xr[1] - xr[3], di = xi[1] - xi[3];
        const double t3r = INV ? -di : di, t3i = INV ? dr : -dr;         // (x1 - x3) * (-i) forward, * (+i) inverse
        xr[0] = t0r + t2r; xi[0] = t0i + t2i;
        xr[1] = t1r + t3r; xi[1] = t1i + t3i;
        xr[2] = t0r - t2r; xi[2] = t0i - t2i;
        xr[3] = t1r - t3r; xi[3] = t1i - t3i;
    } else if (R == 3) {
        const double s3 = 0.86602540378443864676372317075294;            // sin(2 pi / 3)
        const double ar = xr[1] + xr[2], ai = xi[1] + xi[2];
        const double br = xr[1] - xr[2], bi = xi[1] - xi[2];
        const double mr = xr[0] - 0.5 * ar, mi = xi[0] - 0.5 * ai;
        // forward: X1 = m - i s3 b, X2 = m + i s3 b
        const double ur = INV ? -s3 * bi : s3 * bi, ui = INV ? s3 * br : -s3 * br;
        xr[0] += ar; xi[0] += ai;
        xr[1] = mr + ur; xi[1] = mi + ui;
        xr[2] = mr - ur; xi[2] = mi - ui;
    } else {   // R == 5
        const double c1 = 0.30901699437494742410229341718282, c2 = -0.80901699437494742410229341718282;   // cos 2pi/5, cos 4pi/5
        const double s1 = 0.95105651629515357211643933337938, s2 = 0.58778525229247312916870595463907;    // sin 2pi/5, sin 4pi/5
        const double a1r = xr[1] + xr[4], a1i = xi[1] + xi[4], a2r = xr[2] + xr[3], a2i = xi[2] + xi[3];
        const double b1r = xr[1] - xr[4], b1i = xi[1] - xi[4], b2r = xr[2] - xr[3], b2i = xi[2] - xi[3];
        const double r1r = xr[0] + c1 * a1r + c2 * a2r, r1i = xi[0] + c1 * a1i + c2 * a2i;
        const double r2r = xr[0] + c2 * a1r + c1 * a2r, r2i = xi[0] + c2 * a1i + c1 * a2i;
        const double i1r = s1 * b1r + s2 * b2r, i1i = s1 * b1i + s2 * b2i;
        const double i2r = s2 * b1r - s1 * b2r, i2i = s2 * b1i - s1 * b2i;
        // forward: X1 = r1 - i i1, X4 = r1 + i i1, X2 = r2 - i i2, X3 = r2 + i i2;  (-i)(a + ib) = b - ia
        const double u1r = INV ? -i1i : i1i, u1i = INV ? i1r : -i1r;
        const double u2r = INV ? -i2i : i2i, u2i = INV ? i2r : -i2r;
        xr[0] += a1r + a2r; xi[0] += a1i + a2i;
        xr[1] = r1r + u1r; xi[1] = r1i + u1i;
        xr[4] = r1r - u1r; xi[4] = r1i - u1i;
        xr[2] = r2r + u2r; xi[2] = r2i + u2i;
        xr[3] = r2r - u2r; xi[3] = r2i - u2i;
    }
}

// one pass of radix R (compile time, so that the butterfly lives in registers)
template <bool INV, int R>
__device__ __forceinline__ void gen_pass(const double2* src, double2* dst, int n, int NB, int s,
                                         const double2* __restrict__ tw) {
    const int nb = n / R;
    for (int w = threadIdx.x; w < nb * NB; w += blockDim.x) {
        const int bf = w / NB, f = w - bf * NB;
        const int p = bf / s, q = bf - p * s;
        double xr[5], xi[5];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const double2 c = src[(bf + k * nb) * NB + f];
            xr[k] = c.x; xi[k] = c.y;
        }
        small_dft<INV>(R, xr, xi);
        const int m1 = s * p;                              // s p j < n for j < R
#pragma unroll
        for (int j = 0; j < R; ++j) {
            double re = xr[j], im = xi[j];
            if (j > 0 && m1 > 0) {
                const double2 t = __ldg(&tw[m1 * j]);
                const double wr = t.x, wi = INV ? t.y : -t.y;
                const double tr = re * wr - im * wi;
                im = re * wi + im * wr; re = tr;
            }
            dst[(q + s * (R * p + j)) * NB + f] = make_double2(re, im);
        }
    }
}

// Transform of NB interleaved sequences of length plan.n: element i of sequence f at [i * NB + f].  `a` holds the
// input, `b` is the second buffer; returns the buffer that holds the (natural-order) result.  tw[m] =
// exp(2 pi i m / n), m = 0..n-1.  Every thread of the block must call this; ends with a barrier.  The caller must
// have a barrier between filling `a` and the call.
template <bool INV>
__device__ __forceinline__ double2* gen_cfft(double2* a, double2* b, int NB, const GenPlan& plan,
                                             const double2* __restrict__ tw) {
    const int n = plan.n;
    int s = 1;
    double2* src = a;
    double2* dst = b;
    for (int pass = 0; pass < plan.nfac; ++pass) {
        const int R = plan.radix[pass];
        if (R == 4) gen_pass<INV, 4>(src, dst, n, NB, s, tw);
        else if (R == 2) gen_pass<INV, 2>(src, dst, n, NB, s, tw);
        else if (R == 3) gen_pass<INV, 3>(src, dst, n, NB, s, tw);
        else gen_pass<INV, 5>(src, dst, n, NB, s, tw);
        __syncthreads();
        double2* t = src; src = dst; dst = t;
        s *= R;
    }
    return src;
}

}  // namespace ps3d
