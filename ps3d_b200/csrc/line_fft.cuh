// Batched real FFT along a strided (x or y) axis: one sweep of
// fftxyp2s / fftxys2p (reference src/fft/sta3dfft.f90:136-260), with the
// reference's four global transposes (fft_pencil.f90:283-330) replaced by
// strided, z-coalesced tile access, and diffx/diffy (sta3dfft.f90:304-377) or
// the u x omega products (inversion.f90:327-350) folded into the load.
//
// Tile = 8 consecutive z (one 64-byte segment per row) x N rows.  The 8 real
// lines are transformed as 4 complex FFTs (two real lines per complex FFT:
// c = a + i b, separated afterwards by Hermitian symmetry), each by N/8
// threads: block = N/2 threads, lane = (f = t&3, u = t>>2) so that one warp
// load instruction covers 8 rows x 64 B.
//
// Output packing is the reference's (stafft.f90:55-58): row k holds Re X_k,
// row N-k holds Im X_k (k = 1..N/2-1), rows 0 and N/2 the real DC/Nyquist
// terms, all scaled 1/sqrt(N); X_k = sum_j x_j exp(-2 pi i jk/N).
#pragma once

#include "fft_core.cuh"

namespace ps3d {

enum { PRO_PLAIN = 0, PRO_DIFF = 1, PRO_CROSS = 2 };

struct LineArgs {
    const double* in0;        // PLAIN/DIFF: the field.  CROSS: a
    const double* in1;        // CROSS: b      (value = a*b - c*d)
    const double* in2;        // CROSS: c
    const double* in3;        // CROSS: d
    double add1, add3;        // CROSS: constants added to b and d (f_cor)
    double* out;
    long long in_os, out_os;  // stride between consecutive outer lines (doubles)
    const long long* in_rowoff;   // [N] offset of row k from the tile base (doubles)
    const long long* out_rowoff;  // [N]
    int nzc;                  // z-chunks per line (pz / 8)
    const double* kdiff;      // DIFF: wavenumber per k = 0..N/2 (0 at k = 0 and N/2)
    double scale;             // 1/sqrt(N)
    const double2* tw;
    int twscale;
};

__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, double x, double y) { *reinterpret_cast<double2*>(p) = make_double2(x, y); }

template <int N, int PRO>
__global__ void __launch_bounds__(N / 2) k_line_fwd(LineArgs a) {
    PS_SMEM(double, sm);
    constexpr int PL = padded_len(N);
    const int t = threadIdx.x, f = t & 3, u = t >> 2;
    const int o = blockIdx.x / a.nzc, zc = blockIdx.x - o * a.nzc;
    const long long ibase = (long long)o * a.in_os + zc * 8 + 2 * f;
    double vr[8], vi[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const long long off = ibase + __ldg(&a.in_rowoff[u + e * (N / 8)]);
        if (PRO == PRO_CROSS) {
            const double2 x0 = ld2(a.in0 + off), x1 = ld2(a.in1 + off);
            const double2 x2 = ld2(a.in2 + off), x3 = ld2(a.in3 + off);
            vr[e] = x0.x * (x1.x + a.add1) - x2.x * (x3.x + a.add3);
            vi[e] = x0.y * (x1.y + a.add1) - x2.y * (x3.y + a.add3);
        } else {
            const double2 x = ld2(a.in0 + off);
            vr[e] = x.x; vi[e] = x.y;
        }
    }
    double* sre = sm + f * 2 * PL;
    double* sim = sre + PL;
    block_cfft<N, false>(vr, vi, u, true, sre, sim, a.tw, a.twscale);
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int idx = padi(u + e * (N / 8));
        sre[idx] = vr[e]; sim[idx] = vi[e];
    }
    __syncthreads();
    const long long obase = (long long)o * a.out_os + zc * 8 + 2 * f;
    const double sc = a.scale, hs = 0.5 * a.scale;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int k = u + e * (N / 8);
        if (k == 0) {
            st2(a.out + obase + __ldg(&a.out_rowoff[0]), sre[0] * sc, sim[0] * sc);
            const int ih = padi(N / 2);
            st2(a.out + obase + __ldg(&a.out_rowoff[N / 2]), sre[ih] * sc, sim[ih] * sc);
        } else {
            const int ik = padi(k), im = padi(N - k);
            const double p = sre[ik], q = sim[ik], r = sre[im], s = sim[im];
            // A_k = (C_k + conj C_{N-k})/2, B_k = (C_k - conj C_{N-k})/(2i)
            st2(a.out + obase + __ldg(&a.out_rowoff[k]), (p + r) * hs, (q + s) * hs);       // Re A, Re B
            st2(a.out + obase + __ldg(&a.out_rowoff[N - k]), (q - s) * hs, (r - p) * hs);   // Im A, Im B
        }
    }
}

template <int N, int PRO>
__global__ void __launch_bounds__(N / 2) k_line_inv(LineArgs a) {
    PS_SMEM(double, sm);
    constexpr int PL = padded_len(N);
    const int t = threadIdx.x, f = t & 3, u = t >> 2;
    const int o = blockIdx.x / a.nzc, zc = blockIdx.x - o * a.nzc;
    const long long ibase = (long long)o * a.in_os + zc * 8 + 2 * f;
    double* sre = sm + f * 2 * PL;
    double* sim = sre + PL;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int k = u + e * (N / 8);
        if (k == 0) {
            double2 x0 = ld2(a.in0 + ibase + __ldg(&a.in_rowoff[0]));
            double2 xh = ld2(a.in0 + ibase + __ldg(&a.in_rowoff[N / 2]));
            if (PRO == PRO_DIFF) { x0.x = x0.y = 0.0; xh.x = xh.y = 0.0; }
            sre[0] = x0.x; sim[0] = x0.y;
            const int ih = padi(N / 2);
            sre[ih] = xh.x; sim[ih] = xh.y;
        } else {
            const double2 xk = ld2(a.in0 + ibase + __ldg(&a.in_rowoff[k]));
            const double2 xm = ld2(a.in0 + ibase + __ldg(&a.in_rowoff[N - k]));
            double Ar = xk.x, Ai = xm.x, Br = xk.y, Bi = xm.y;
            if (PRO == PRO_DIFF) {
                // d/dx: X_k -> i kappa X_k  (sta3dfft.f90:325-329)
                const double kap = __ldg(&a.kdiff[k]);
                const double ar = -kap * Ai, ai = kap * Ar, br = -kap * Bi, bi = kap * Br;
                Ar = ar; Ai = ai; Br = br; Bi = bi;
            }
            const int ik = padi(k), im = padi(N - k);
            sre[ik] = Ar - Bi; sim[ik] = Ai + Br;     // C_k     = A + i B
            sre[im] = Ar + Bi; sim[im] = Br - Ai;     // C_{N-k} = conj(A) + i conj(B)
        }
    }
    __syncthreads();
    double vr[8], vi[8];
    fft_gather<N>(vr, vi, u, sre, sim);
    __syncthreads();
    block_cfft<N, true>(vr, vi, u, true, sre, sim, a.tw, a.twscale);
    const long long obase = (long long)o * a.out_os + zc * 8 + 2 * f;
    const double sc = a.scale;
#pragma unroll
    for (int e = 0; e < 8; ++e)
        st2(a.out + obase + __ldg(&a.out_rowoff[u + e * (N / 8)]), vr[e] * sc, vi[e] * sc);
}

template <int N>
constexpr size_t line_smem_bytes() { return (size_t)4 * 2 * padded_len(N) * sizeof(double); }

}  // namespace ps3d
