// Batched real FFT along a strided (x or y) axis: one sweep of
// fftxyp2s / fftxys2p (reference src/fft/sta3dfft.f90:136-260), with the
// reference's four global transposes (fft_pencil.f90:283-330) replaced by
// strided, z-coalesced tile access, and diffx/diffy (sta3dfft.f90:304-377) or
// the u x omega products (inversion.f90:327-350) folded into the load.
//
// Tile = 16 consecutive z (one full 128-byte line per row) x N rows.  The 16
// real lines are transformed as 8 complex FFTs (two real lines per complex
// FFT: c = a + i b, separated afterwards by Hermitian symmetry), each by N/8
// threads: block = N threads, lane = (f = t&7, u = t>>3) so that one warp
// load instruction covers 4 rows x 128 B (the minimum number of L1 wavefronts).
// The 8 FFTs are interleaved, as complex numbers, in the shared-memory scratch (IxIlv<8>).
//
// Output packing is the reference's (stafft.f90:55-58): row k holds Re X_k,
// row N-k holds Im X_k (k = 1..N/2-1), rows 0 and N/2 the real DC/Nyquist
// terms, all scaled 1/sqrt(N); X_k = sum_j x_j exp(-2 pi i jk/N).
#pragma once

#include "fft_core.cuh"

namespace ps3d {

enum { PRO_PLAIN = 0, PRO_DIFF = 1, PRO_CROSS = 2 };
constexpr int LINE_ZC = 16;            // z values per tile row (pz is a multiple of this)
constexpr int LINE_NF = LINE_ZC / 2;   // complex FFTs per tile
// N = 1024: tiles of 8 z (64-byte rows, 4 complex FFTs, 512 threads, 64 KB of scratch) so that two blocks fit an
// SM like at N = 512; a 16-z tile would need 128 KB and 1024 threads, one block per SM, with nothing to overlap
// its load and store phases
// Sweeps that scatter their rows into peer memory keep 16-z tiles at every N: 128-byte stores over NVLink.
__host__ __device__ constexpr int line_zc(int n) { return n >= 1024 ? 8 : LINE_ZC; }
__host__ __device__ constexpr int line_threads(int n, int zc) { return (zc / 2) * n / 8; }

// Row k of a line -> (destination block d, offset in doubles from the tile base), computed arithmetically
// and branch-free so that the unrolled global accesses of a tile can all be issued back to back.
//   kp = k, or for paired == 1 the position of ky = k in the paired order ky' = 0, n/2, 1, n-1, 2, n-2, ...
//   d  = kp >> logblk (owner rank of the row), kl = kp & (2^logblk - 1) (row inside the owner's slab)
//   offset = (self < 0 ? d : self) * nb + kl * stride
// One rank, or reading/writing a local buffer laid out as P blocks: self = -1 and the block index is the
// owner d (for a plain row map logblk = 30, so d = 0 and kl = k).  Peer-memory scatter (the sweep stores
// straight into rank d's receive buffer, see LineArgs::outp): self = this rank, the block it fills there.
struct RowMap {
    long long stride;
    long long nb;          // doubles per block (nxl * nyl * pz)
    int paired, n, logblk, self;
    int blk;               // rows per block (= 2^logblk when that is a power of two; the mixed-radix kernels divide by it)
};
__device__ __forceinline__ long long row_off(const RowMap& m, int k, int& d) {
    const int h = m.n >> 1;
    int kq = (k < h) ? 2 * k : 2 * (m.n - k) + 1;
    kq = (k == 0) ? 0 : kq;
    kq = (k == h) ? 1 : kq;
    const int kp = m.paired ? kq : k;
    d = kp >> m.logblk;
    const int kl = kp & ((1 << m.logblk) - 1);
    const int blk = (m.self < 0) ? d : m.self;
    return (long long)blk * m.nb + (long long)kl * m.stride;
}
__device__ __forceinline__ long long row_off(const RowMap& m, int k) { int d; return row_off(m, k, d); }
// the same map with a block size that need not be a power of two (line_gen.cuh)
__device__ __forceinline__ long long row_off_div(const RowMap& m, int k, int& d) {
    const int h = m.n >> 1;
    int kq = (k < h) ? 2 * k : 2 * (m.n - k) + 1;
    kq = (k == 0) ? 0 : kq;
    kq = (k == h) ? 1 : kq;
    const int kp = m.paired ? kq : k;
    d = kp / m.blk;
    const int kl = kp - d * m.blk;
    const int blk = (m.self < 0) ? d : m.self;
    return (long long)blk * m.nb + (long long)kl * m.stride;
}
__device__ __forceinline__ long long row_off_div(const RowMap& m, int k) { int d; return row_off_div(m, k, d); }

struct LineArgs {
    const double* in0;        // PLAIN/DIFF: the field.  CROSS: a
    const double* in1;        // CROSS: b      (value = a*b - c*d)
    const double* in2;        // CROSS: c
    const double* in3;        // CROSS: d
    double add1, add3;        // CROSS: constants added to b and d (f_cor)
    double* out;
    double* outp[8];          // peer-memory scatter: base of rank d's receive buffer (all = out otherwise)
    long long in_os, out_os;  // stride between consecutive outer lines (doubles)
    RowMap in_map, out_map;   // offset of row k from the tile base
    int ntiles;               // tiles of this launch (outer lines x nzc); a block loops over tiles blockIdx.x, +gridDim.x, ...
    int nzc;                  // z-chunks handled by this launch
    int zc0;                  // first z-chunk of this launch
    int in_zc0, out_zc0;      // z-chunk held at offset 0 of the input / output array (0 for full arrays; = zc0 for
                              // the compact, L2-resident intermediate of a chunked 2-D FFT)
    int final_store;          // 1: streaming stores (result leaves the cache), 0: keep in L2 for the next sweep
    int scatter_fence;        // peer-memory scatter: system-scope fence before the block retires
    const double* kdiff;      // DIFF: wavenumber per k = 0..N/2 (0 at k = 0 and N/2)
    double scale;             // 1/sqrt(N)
    const double2* tw;
    int twscale;
};

__device__ __forceinline__ double* row_dst(const struct LineArgs& a, int k);
// The row offsets of a tile do not depend on the tile, so the compiler hoists all of them out of the
// persistent tile loop and then spills them (64-bit each) to local memory; an opaque copy of the thread's
// row index keeps the (cheap, branch-free) offset arithmetic inside the loop and the kernel spill-free.
__device__ __forceinline__ int tile_local(int u) {
#ifndef PS3D_EMU
    asm volatile("" : "+r"(u));
#endif
    return u;
}
// sweep inputs are dead after the load: streaming (evict-first) loads
__device__ __forceinline__ double2 ld2(const double* p) { return __ldcs(reinterpret_cast<const double2*>(p)); }
__device__ __forceinline__ void st2(double* p, double x, double y) { *reinterpret_cast<double2*>(p) = make_double2(x, y); }
__device__ __forceinline__ void st2f(int fin, double* p, double x, double y) {
    if (fin) __stcs(reinterpret_cast<double2*>(p), make_double2(x, y));
    else *reinterpret_cast<double2*>(p) = make_double2(x, y);
}

__device__ __forceinline__ double* row_dst(const LineArgs& a, int k) {
    int d;
    const long long off = row_off(a.out_map, k, d);
    return a.outp[d & 7] + off;
}
__device__ __forceinline__ double* row_dst_div(const LineArgs& a, int k) {
    int d;
    const long long off = row_off_div(a.out_map, k, d);
    return a.outp[d & 7] + off;
}

template <int N, int PRO, int ZC = line_zc(N)>
__global__ void __launch_bounds__(line_threads(N, ZC), (N >= 64 && N <= 1024) ? 1024 / line_threads(N, ZC) : 1) k_line_fwd(LineArgs a) {
    PS_SMEM(double, sm);
    constexpr int NF = ZC / 2;
    const int t = threadIdx.x, f = t & (NF - 1), u = t / NF;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    const int o = tile / a.nzc, zc = a.zc0 + (tile - o * a.nzc);
    const long long ibase = (long long)o * a.in_os + (zc - a.in_zc0) * ZC + 2 * f;
    const int ul = tile_local(u);
    double vr[8], vi[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const long long off = ibase + row_off(a.in_map, ul + e * (N / 8));
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
    double* sre = sm;
    double* sim = sm + NF * N;
    const IxIlv<NF> ix{f};
    block_cfft<N, false>(vr, vi, u, true, sre, sim, ix, TwGlobal{a.tw, a.twscale});
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 8; ++e) ix.put(sre, sim, u + e * (N / 8), vr[e], vi[e]);
    __syncthreads();
    const long long obase = (long long)o * a.out_os + (zc - a.out_zc0) * ZC + 2 * f;
    const double sc = a.scale, hs = 0.5 * a.scale;
#pragma unroll(PRO == PRO_CROSS ? 2 : 4)
    for (int e = 0; e < 4; ++e) {
        const int k = u + e * (N / 8);
        if (k == 0) {
            double p, q;
            ix.get(sre, sim, 0, p, q);
            st2f(a.final_store, row_dst(a, 0) + obase, p * sc, q * sc);
            ix.get(sre, sim, N / 2, p, q);
            st2f(a.final_store, row_dst(a, N / 2) + obase, p * sc, q * sc);
        } else {
            double p, q, r, s;
            ix.get(sre, sim, k, p, q);
            ix.get(sre, sim, N - k, r, s);
            // A_k = (C_k + conj C_{N-k})/2, B_k = (C_k - conj C_{N-k})/(2i)
            st2f(a.final_store, row_dst(a, k) + obase, (p + r) * hs, (q + s) * hs);       // Re A, Re B
            st2f(a.final_store, row_dst(a, N - k) + obase, (q - s) * hs, (r - p) * hs);   // Im A, Im B
        }
    }
    __syncthreads();          // scratch is reused by the next tile of a persistent launch
    }
    if (a.out_map.self >= 0 && a.scatter_fence) __threadfence_system();     // peer-memory scatter: stores visible to the owner GPU
}

template <int N, int PRO, int ZC = line_zc(N)>
__global__ void __launch_bounds__(line_threads(N, ZC), (N >= 64 && N <= 1024) ? 1024 / line_threads(N, ZC) : 1) k_line_inv(LineArgs a) {
    PS_SMEM(double, sm);
    constexpr int NF = ZC / 2;
    const int t = threadIdx.x, f = t & (NF - 1), u = t / NF;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    const int o = tile / a.nzc, zc = a.zc0 + (tile - o * a.nzc);
    const long long ibase = (long long)o * a.in_os + (zc - a.in_zc0) * ZC + 2 * f;
    double* sre = sm;
    double* sim = sm + NF * N;
    const IxIlv<NF> ix{f};
    // rows k and N-k (k = 0: rows 0 and N/2); all eight loads are issued before the first use
    double2 xa[4], xb[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int k = u + e * (N / 8);
        const int kb = (k == 0) ? N / 2 : N - k;
        xa[e] = ld2(a.in0 + ibase + row_off(a.in_map, k));
        xb[e] = ld2(a.in0 + ibase + row_off(a.in_map, kb));
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int k = u + e * (N / 8);
        const bool k0 = (k == 0);
        const int kb = k0 ? N / 2 : N - k;
        double Ar = xa[e].x, Ai = xb[e].x, Br = xa[e].y, Bi = xb[e].y;
        if (PRO == PRO_DIFF) {
            // d/dx: X_k -> i kappa X_k  (sta3dfft.f90:325-329); kappa = 0 at k = 0 and N/2 (:321-335)
            const double kap = __ldg(&a.kdiff[k]);
            const double ar = -kap * Ai, ai = kap * Ar, br = -kap * Bi, bi = kap * Br;
            Ar = ar; Ai = ai; Br = br; Bi = bi;
        }
        // generic k: C_k = A + i B, C_{N-k} = conj(A) + i conj(B);  k = 0: the two rows are the real DC and
        // Nyquist terms of both lines, C_0 = a_0 + i b_0, C_{N/2} = a_{N/2} + i b_{N/2}
        double rk = Ar - Bi, qk = Ai + Br, rm = Ar + Bi, qm = Br - Ai;
        if (k0) {
            const bool z = (PRO == PRO_DIFF);
            rk = z ? 0.0 : xa[e].x; qk = z ? 0.0 : xa[e].y;
            rm = z ? 0.0 : xb[e].x; qm = z ? 0.0 : xb[e].y;
        }
        ix.put(sre, sim, k, rk, qk);
        ix.put(sre, sim, kb, rm, qm);
    }
    __syncthreads();
    double vr[8], vi[8];
    fft_gather<N>(vr, vi, u, sre, sim, ix);
    __syncthreads();
    block_cfft<N, true>(vr, vi, u, true, sre, sim, ix, TwGlobal{a.tw, a.twscale});
    const long long obase = (long long)o * a.out_os + (zc - a.out_zc0) * ZC + 2 * f;
    const double sc = a.scale;
#pragma unroll
    for (int e = 0; e < 8; ++e)
        st2f(a.final_store, row_dst(a, u + e * (N / 8)) + obase, vr[e] * sc, vi[e] * sc);
    __syncthreads();          // scratch is reused by the next tile of a persistent launch
    }
    if (a.out_map.self >= 0 && a.scatter_fence) __threadfence_system();     // peer-memory scatter: stores visible to the owner GPU
}

template <int N, int ZC = line_zc(N)>
constexpr size_t line_smem_bytes() { return (size_t)(ZC / 2) * 2 * N * sizeof(double); }

}  // namespace ps3d
