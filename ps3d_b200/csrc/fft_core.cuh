// Block-cooperative FP64 complex FFT of power-of-two length N, Stockham
// autosort (decimation in frequency), radix-8 passes with a radix-4/2 tail.
//
// Each participating thread `u` in [0, N/8) owns eight complex values per pass:
//   v[e] = x[u + e*N/8],  e = 0..7
// both on entry (caller supplies the first pass's inputs in registers, e.g.
// straight from global memory) and on exit (natural-order spectrum, so the
// caller can store straight to global memory).  Between passes the data goes
// through a padded shared-memory scratch line (re/im split, conflict-free for
// the stride-R Stockham scatter).  No tensor cores: the work is ~2.5 N log2 N
// flops per real line against 16 B of HBM traffic per point, i.e. HBM-bound.
//
// Replaces the per-stage out-of-place sweeps of forrdx*/revrdx*
// (reference src/fft/stafft.f90:561-1757) — same transform, different
// factorisation (8·8·8… instead of 4·4·…·2).
#pragma once

#include "rt.h"

namespace ps3d {

// Scratch layouts of the Stockham exchanges (put/get of complex element i of this thread's FFT):
//  IxSwz  : one FFT per scratch line, re and im in two separate planes of doubles (sre, sim), lanes of a warp
//           are consecutive butterflies u.  XOR swizzle b0^=i4, b1^=i5, b2^=i6, b3^=i6: the bits that vary
//           across a half-warp in pass 1 (i3..i6), pass 2 (i0..i2,i6) and pass 3 (i0..i3) always map onto 16
//           distinct 8-byte columns (shared memory: 32 4-byte banks).  No padding.  Used by the z-column
//           kernels, whose two column buffers are the two planes.
//  IxIlv  : NF = 8 FFTs interleaved as complex numbers: element i of FFT f is the double2 at 8*i + f of one
//           array (sre; sim unused), lanes = f fastest then u.  A quarter-warp (the unit of a 128-bit shared
//           access) is the 8 FFTs at one element index: one contiguous 128-byte row, conflict-free in the
//           stride-8 scatter and the unit-stride gather alike, half the shared-memory instructions of split
//           planes and no swizzle arithmetic.  Used by the x/y line kernels.
struct IxSwz {
    __device__ __forceinline__ int operator()(int i) const { return i ^ (((i >> 4) & 7) | (((i >> 6) & 1) << 3)); }
    __device__ __forceinline__ void put(double* sre, double* sim, int i, double re, double im) const {
        const int k = (*this)(i);
        sre[k] = re; sim[k] = im;
    }
    __device__ __forceinline__ void get(const double* sre, const double* sim, int i, double& re, double& im) const {
        const int k = (*this)(i);
        re = sre[k]; im = sim[k];
    }
};
template <int NF>
struct IxIlv {
    int f;
    // (NF = 4, the 8-z tiles of N = 1024: a quarter-warp is two element indices x 4 FFTs = two 64-byte half rows;
    //  they collide (2-way) only in the stride-8 scatter of the first pass, i = 8 u + j.  A swizzle that avoids it
    //  costs more in index registers -- the kernel is at its 64-register cap -- than the conflict does.)
    __device__ __forceinline__ int at(int i) const { return NF * i + f; }
    __device__ __forceinline__ void put(double* sre, double*, int i, double re, double im) const {
        reinterpret_cast<double2*>(sre)[at(i)] = make_double2(re, im);
    }
    __device__ __forceinline__ void get(const double* sre, const double*, int i, double& re, double& im) const {
        const double2 c = reinterpret_cast<const double2*>(sre)[at(i)];
        re = c.x; im = c.y;
    }
};

template <bool INV>
__device__ __forceinline__ void radix4(double& r0, double& i0, double& r1, double& i1,
                                       double& r2, double& i2, double& r3, double& i3) {
    const double t0r = r0 + r2, t0i = i0 + i2;
    const double t1r = r0 - r2, t1i = i0 - i2;
    const double t2r = r1 + r3, t2i = i1 + i3;
    const double dr = r1 - r3, di = i1 - i3;
    // t3 = (c1 - c3) * (-i) forward, * (+i) inverse
    const double t3r = INV ? -di : di;
    const double t3i = INV ? dr : -dr;
    r0 = t0r + t2r; i0 = t0i + t2i;
    r1 = t1r + t3r; i1 = t1i + t3i;
    r2 = t0r - t2r; i2 = t0i - t2i;
    r3 = t1r - t3r; i3 = t1i - t3i;
}

// In-place radix-R DFT of x[0..R-1] (forward: exp(-2 pi i jk/R)).
template <int R, bool INV>
__device__ __forceinline__ void radix(double* xr, double* xi) {
    if (R == 2) {
        const double ar = xr[0], ai = xi[0];
        xr[0] = ar + xr[1]; xi[0] = ai + xi[1];
        xr[1] = ar - xr[1]; xi[1] = ai - xi[1];
    } else if (R == 4) {
        radix4<INV>(xr[0], xi[0], xr[1], xi[1], xr[2], xi[2], xr[3], xi[3]);
    } else {
        const double h = 0.70710678118654752440084436210485;
        // even and odd radix-4 sub-transforms
        radix4<INV>(xr[0], xi[0], xr[2], xi[2], xr[4], xi[4], xr[6], xi[6]);
        radix4<INV>(xr[1], xi[1], xr[3], xi[3], xr[5], xi[5], xr[7], xi[7]);
        // E_j in slots 0,2,4,6 ; O_j in slots 1,3,5,7
        // O_1 *= w, O_2 *= w^2, O_3 *= w^3 with w = exp(-+ i pi/4)
        double o1r, o1i, o2r, o2i, o3r, o3i;
        if (!INV) {
            o1r = (xr[3] + xi[3]) * h; o1i = (xi[3] - xr[3]) * h;
            o2r = xi[5];               o2i = -xr[5];
            o3r = (xi[7] - xr[7]) * h; o3i = -(xr[7] + xi[7]) * h;
        } else {
            o1r = (xr[3] - xi[3]) * h; o1i = (xr[3] + xi[3]) * h;
            o2r = -xi[5];              o2i = xr[5];
            o3r = -(xr[7] + xi[7]) * h; o3i = (xr[7] - xi[7]) * h;
        }
        const double e0r = xr[0], e0i = xi[0], e1r = xr[2], e1i = xi[2];
        const double e2r = xr[4], e2i = xi[4], e3r = xr[6], e3i = xi[6];
        const double o0r = xr[1], o0i = xi[1];
        xr[0] = e0r + o0r; xi[0] = e0i + o0i;
        xr[4] = e0r - o0r; xi[4] = e0i - o0i;
        xr[1] = e1r + o1r; xi[1] = e1i + o1i;
        xr[5] = e1r - o1r; xi[5] = e1i - o1i;
        xr[2] = e2r + o2r; xi[2] = e2i + o2i;
        xr[6] = e2r - o2r; xi[6] = e2i - o2i;
        xr[3] = e3r + o3r; xi[3] = e3i + o3i;
        xr[7] = e3r - o3r; xi[7] = e3i - o3i;
    }
}

// Twiddle sources.  Only the radix-8 passes that are not the last pass carry twiddles (the radix-4/2 tail is
// always the last pass), and butterfly b of such a pass needs W^m, W^2m, ..., W^7m with m = s * (b / s),
// W = exp(-+2 pi i / N).  A source hands out W^m, W^2m, W^4m as (cos, sin) pairs; the other powers are single
// products in fft_pass.
//  TwGlobal : read-only global table tw[j] = exp(2 pi i j / NTW), NTW = twscale * N; all three powers come
//             from the table (correctly rounded).  Used by the x/y line kernels.
//  TwSin    : (zcol.cuh) W^m from the block's shared sine table, W^2m and W^4m by squaring.
struct TwGlobal {
    const double2* __restrict__ tw;
    int twscale;
    __device__ __forceinline__ void get(int /*pass*/, int s, int p, double2& w1, double2& w2, double2& w4) const {
        const int m1 = s * p * twscale;
        w1 = __ldg(&tw[m1]); w2 = __ldg(&tw[2 * m1]); w4 = __ldg(&tw[4 * m1]);
    }
};
// One Stockham DIF pass of radix R on the thread's eight values.
//   butterfly g (g < 8/R) has index b = u + g*N/8 in [0, N/R); its inputs are
//   v[g + k*(8/R)], its outputs overwrite the same registers (output j at
//   v[g + j*(8/R)]) and belong at x'[q + s*(R*p + j)], p = b / s, q = b % s.
// `s` is the product of the radices of the previous passes, `pass` their number.
template <int N, int R, bool INV, class TW>
__device__ __forceinline__ void fft_pass(double (&vr)[8], double (&vi)[8], int u, int s, int pass, const TW& tws) {
    constexpr int G = 8 / R;
#pragma unroll
    for (int g = 0; g < G; ++g) {
        double xr[R], xi[R];
#pragma unroll
        for (int k = 0; k < R; ++k) { xr[k] = vr[g + k * G]; xi[k] = vi[g + k * G]; }
        radix<R, INV>(xr, xi);
        if (R == 8 && s * R < N) {
            const int b = u + g * (N / 8);
            const int p = b / s;
            double wr[8], wi[8];
            {
                double2 w1, w2, w4;
                tws.get(pass, s, p, w1, w2, w4);
                wr[1] = w1.x; wi[1] = INV ? w1.y : -w1.y;
                wr[2] = w2.x; wi[2] = INV ? w2.y : -w2.y;
                wr[4] = w4.x; wi[4] = INV ? w4.y : -w4.y;
                wr[3] = wr[1] * wr[2] - wi[1] * wi[2]; wi[3] = wr[1] * wi[2] + wi[1] * wr[2];
                wr[5] = wr[1] * wr[4] - wi[1] * wi[4]; wi[5] = wr[1] * wi[4] + wi[1] * wr[4];
                wr[6] = wr[2] * wr[4] - wi[2] * wi[4]; wi[6] = wr[2] * wi[4] + wi[2] * wr[4];
                wr[7] = wr[3] * wr[4] - wi[3] * wi[4]; wi[7] = wr[3] * wi[4] + wi[3] * wr[4];
            }
#pragma unroll
            for (int j = 1; j < R; ++j) {
                const double tr = xr[j] * wr[j] - xi[j] * wi[j];
                const double ti = xr[j] * wi[j] + xi[j] * wr[j];
                xr[j] = tr; xi[j] = ti;
            }
        }
#pragma unroll
        for (int k = 0; k < R; ++k) { vr[g + k * G] = xr[k]; vi[g + k * G] = xi[k]; }
    }
}

template <int N, int R, class IX>
__device__ __forceinline__ void fft_scatter(const double (&vr)[8], const double (&vi)[8], int u, int s,
                                            double* __restrict__ sre, double* __restrict__ sim, const IX& ix) {
    constexpr int G = 8 / R;
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const int b = u + g * (N / 8);
        const int p = b / s;
        const int q = b - p * s;
#pragma unroll
        for (int j = 0; j < R; ++j) {
            ix.put(sre, sim, q + s * (R * p + j), vr[g + j * G], vi[g + j * G]);
        }
    }
}

template <int N, class IX>
__device__ __forceinline__ void fft_gather(double (&vr)[8], double (&vi)[8], int u,
                                           const double* __restrict__ sre, const double* __restrict__ sim, const IX& ix) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        ix.get(sre, sim, u + e * (N / 8), vr[e], vi[e]);
    }
}

// radix of the last pass for length N = 8^a * tail
__host__ __device__ constexpr int fft_tail(int n) { return (n % 8 == 0 && n > 8) ? fft_tail(n / 8) : n; }
// tail is 8 (pure), 16 -> handled as 8 then 2, 32 -> 8 then 4, ...
__host__ __device__ constexpr bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

// Full transform.  All N/8 threads of the FFT must call this together (it
// contains __syncthreads(), so every thread of the BLOCK must take part, with
// `active` false for threads that own no FFT: they neither read nor write the
// scratch).  sre/sim: this FFT's scratch line (N doubles each for IxSwz).  The
// scratch must not be in use by other threads when the call is entered.
// Barrier over the threads that share the exchange scratch: the whole block (default), or a named barrier over
// one group of a block that runs several independent FFT groups (line_tma.cuh).
struct BarBlock { __device__ __forceinline__ void operator()() const { __syncthreads(); } };

template <int N, bool INV, class IX, class TW, class BAR = BarBlock>
__device__ __forceinline__ void block_cfft(double (&vr)[8], double (&vi)[8], int u, bool active,
                                           double* __restrict__ sre, double* __restrict__ sim, const IX& ix,
                                           const TW& tws, const BAR& bar = BAR()) {
    static_assert(is_pow2(N) && N >= 8, "power-of-two lengths >= 8 only");
    int s = 1;
    constexpr int L = (N == 8) ? 3 : (N == 16) ? 4 : (N == 32) ? 5 : (N == 64) ? 6 : (N == 128) ? 7 :
                      (N == 256) ? 8 : (N == 512) ? 9 : (N == 1024) ? 10 : (N == 2048) ? 11 : (N == 4096) ? 12 : -1;
    static_assert(L > 0, "unsupported FFT length");
    constexpr int N8 = L / 3;            // number of radix-8 passes
    constexpr int TAIL = 1 << (L % 3);   // 1, 2 or 4: radix of the last pass when L is not a multiple of 3
#pragma unroll
    for (int pass = 0; pass < N8; ++pass) {
        if (active) fft_pass<N, 8, INV>(vr, vi, u, s, pass, tws);
        const bool last = (pass == N8 - 1) && (TAIL == 1);
        if (!last) {
            if (pass > 0) bar();                         // WAR: everyone has gathered
            if (active) fft_scatter<N, 8>(vr, vi, u, s, sre, sim, ix);
            bar();
            if (active) fft_gather<N>(vr, vi, u, sre, sim, ix);
        }
        s *= 8;
    }
    if (TAIL == 2) {
        if (active) fft_pass<N, 2, INV>(vr, vi, u, s, N8, tws);
    } else if (TAIL == 4) {
        if (active) fft_pass<N, 4, INV>(vr, vi, u, s, N8, tws);
    }
}

}  // namespace ps3d
