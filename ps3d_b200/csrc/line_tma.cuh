// TMA-staged x/y line sweeps: the same transforms as line_fft.cuh (one sweep of fftxyp2s / fftxys2p, reference
// src/fft/sta3dfft.f90:136-260, with diffx/diffy :304-377 folded into the inverse load), restructured so that the
// global loads never wait for a block's compute phases.
//
// Why: the register-staged sweeps of line_fft.cuh are load-latency bound.  A block issues its 64 KB of loads, waits
// for them, then runs three FFT passes and stores; with two blocks per SM the bytes in flight average ~16 KB per SM
// where HBM3e needs ~45 KB (Little's law at ~1 us).  Here ONE persistent block per SM runs G independent FFT
// groups (each = the thread layout of one line_fft tile, its own exchange scratch, a named barrier) and NL landing
// buffers that the TMA unit fills with whole tiles (cp.async.bulk.tensor, completion on an mbarrier).  The load of
// tile i+NL is issued the moment the group that owns tile i has copied it out of its landing buffer, i.e. a tile
// is always in flight while the groups compute.  The inverse sweep builds its FFT input straight from the landing
// rows k and N-k (16 shared loads per thread instead of 8 global loads + 8 stores + 8 loads of the put/gather
// exchange of line_fft.cuh).
//
// Shared memory (N = 512: 64 KB tiles): G = 2 scratch + NL = 1 landing = 192 KB; registers capped at 64 (1024
// threads), as in line_fft.cuh.  Tensor maps are rank 4, ascending strides:
//   y sweeps: (z, row inside block, outer line, block)     x sweeps: (z, outer line, row, 1)
// so that the box {ZC, rows, 1, 1} / {ZC, 1, rows, 1} lands as dense [row][ZC] in both cases.
#pragma once
#ifndef PS3D_EMU

#include <cuda.h>
#include "line_fft.cuh"

namespace ps3d {

struct TmaArgs {
    int mode;          // 0: y sweep, coordinates (z, r, o, blk);  1: x sweep, coordinates (z, o, r, 0)
    int nops;          // bulk tensor copies per tile
    int rows_per_op;   // rows of one copy (<= 256)
    int blkrows;       // y sweeps: rows per block of the source array (n on one rank, nyl for received slab blocks)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// named barrier over one FFT group (ids 1..G; 0 is __syncthreads)
template <int T>
struct BarGroup {
    int id;
    __device__ __forceinline__ void operator()() const { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(T) : "memory"); }
};

// FFT groups per block and landing buffers for line length N at ZC z values per tile
__host__ __device__ constexpr int tma_groups(int n, int zc) { return 1024 / line_threads(n, zc); }
__host__ __device__ constexpr int tma_landing(int n, int zc) {
    return (n * zc * 8 >= 65536) ? 1 : (n * zc * 8 >= 32768) ? 2 : 4;
}
template <int N, int ZC>
constexpr size_t line_tma_smem_bytes() {
    return (size_t)(tma_groups(N, ZC) + tma_landing(N, ZC)) * N * ZC * sizeof(double) + 8 * tma_groups(N, ZC) * (tma_groups(N, ZC) + tma_landing(N, ZC)) + 64;
}

// ROT = false: NL landing buffers + one exchange scratch per group (a tile is copied out of its landing buffer, which
//   is refilled at once).
// ROT = true : the NB = G + NL buffers rotate; tile i lands in buffer i % NB, is copied to registers, and the SAME
//   buffer then serves as the exchange scratch of its group; it is refilled (tile i + NB) when the group has finished
//   its last exchange.  A load is then issued most of a tile time before it is needed instead of less than half.
template <int N, bool INV, int PRO, int ZC, bool ROT>
__global__ void __launch_bounds__(1024, 1) k_line_tma(LineArgs a, TmaArgs ta, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) unsigned char smraw[];
    constexpr int T = line_threads(N, ZC), NF = ZC / 2, G = tma_groups(N, ZC), NL = tma_landing(N, ZC);
    constexpr int TILE = N * ZC;                       // doubles per tile
    constexpr int NB = G + NL;                         // ROT: rotating buffers
    constexpr int NBAR = ROT ? G * NB : G;             // mbarriers: tile i -> full[i % NBAR] (one group, one buffer each)
    constexpr int AHEAD = ROT ? NB : NL;               // tiles in flight / owned
    static_assert(G * T == 1024 && G <= 15 && G % NL == 0, "one block = 1024 threads = G FFT groups; NL divides G");
    double* land = reinterpret_cast<double*>(smraw);
    double* scr0 = land + (size_t)NL * TILE;
    uint64_t* full = reinterpret_cast<uint64_t*>(land + (size_t)NB * TILE);
    const int grp = threadIdx.x / T, t = threadIdx.x - grp * T, f = t & (NF - 1), u = t / NF;
    double* sre = scr0 + (size_t)grp * TILE;           // (ROT: set per tile)
    double* sim = sre + NF * N;
    IxIlv<NF> ix{f};
    const BarGroup<T> bar{1 + grp};
    const int nseq = (a.ntiles > (int)blockIdx.x) ? (a.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    auto issue = [&](int i) {
        const int tile = blockIdx.x + i * gridDim.x;
        const int o = tile / a.nzc, zc = a.zc0 + (tile - o * a.nzc);
        const int slot = ROT ? i % NB : i % NL;
        // one mbarrier per (FFT group, buffer) combination (tile i belongs to group i % G): a barrier's phases are then
        // waited for strictly in order by one group, which the parity wait needs (waiting one phase ahead of
        // the current one would succeed at once)
        uint64_t* fb = &full[i % NBAR];
        mbar_expect_tx(fb, (uint32_t)(TILE * sizeof(double)));
        for (int j = 0; j < ta.nops; ++j) {
            const int r = j * ta.rows_per_op;
            double* dst = land + (size_t)slot * TILE + (size_t)r * ZC;
            if (ta.mode == 0) {
                const int blk = r / ta.blkrows;
                tma_load_4d(dst, &tmap, fb, zc * ZC, r - blk * ta.blkrows, o, blk);
            } else {
                tma_load_4d(dst, &tmap, fb, zc * ZC, o, r, 0);
            }
        }
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < NBAR; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int i = 0; i < AHEAD && i < nseq; ++i) issue(i);

    for (int i = grp; i < nseq; i += G) {
        const int tile = blockIdx.x + i * gridDim.x;
        const int o = tile / a.nzc, zc = a.zc0 + (tile - o * a.nzc);
        const int slot = ROT ? i % NB : i % NL;
        mbar_wait(&full[i % NBAR], (uint32_t)((i / NBAR) & 1));
        if (ROT) { sre = land + (size_t)slot * TILE; sim = sre + NF * N; }
        const double2* L = reinterpret_cast<const double2*>(land + (size_t)slot * TILE) + f;    // row r: L[r * NF]
        const int paired = a.in_map.paired;
        constexpr int H = N / 2;
        double vr[8], vi[8];
        if (!INV) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const double2 x = L[(u + e * (N / 8)) * NF];
                vr[e] = x.x; vi[e] = x.y;
            }
        } else {
            // FFT input at position p = u + e N/8 straight from the landing rows k = min(p, N-p) and N-k:
            //   p < N/2: C_k = A + i B;   p > N/2: C_{N-k} = conj(A) + i conj(B);   p = 0, N/2: the real DC / Nyquist rows
            // with A = (row_k.x, row_{N-k}.x), B = (row_k.y, row_{N-k}.y) (packing of line_fft.cuh)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int p = u + e * (N / 8);
                const bool lo = (p < H), edge = (p == 0) || (p == H);
                const int k = lo ? p : N - p;                 // 0 < k < N/2 unless edge (then k = p)
                const int kb = edge ? k : N - k;
                // landing row of ky = k: paired order 0, N/2, 1, N-1, 2, ... for the inverse y sweep
                int ra = k, rb = kb;
                if (paired) {
                    ra = (k == 0) ? 0 : (k == H) ? 1 : 2 * k;
                    rb = edge ? ra : 2 * k + 1;               // kb = N - k > N/2  ->  2 (N - kb) + 1
                }
                const double2 xa = L[ra * NF], xb = L[rb * NF];
                double Ar = xa.x, Ai = xb.x, Br = xa.y, Bi = xb.y;
                if (PRO == PRO_DIFF) {
                    const double kap = __ldg(&a.kdiff[k]);
                    const double ar = -kap * Ai, ai = kap * Ar, br = -kap * Bi, bi = kap * Br;
                    Ar = ar; Ai = ai; Br = br; Bi = bi;
                }
                double re = lo ? Ar - Bi : Ar + Bi;
                double im = lo ? Ai + Br : Br - Ai;
                if (edge) {
                    const bool z = (PRO == PRO_DIFF);
                    re = z ? 0.0 : xa.x; im = z ? 0.0 : xa.y;
                }
                vr[e] = re; vi[e] = im;
            }
        }
        bar();                                   // every thread of the group has copied its part of the landing tile
        if (!ROT && t == 0 && i + NL < nseq) issue(i + NL);
        const int ul = tile_local(u);
        if (!INV) {
            block_cfft<N, false>(vr, vi, u, true, sre, sim, ix, TwGlobal{a.tw, a.twscale}, bar);
            bar();
#pragma unroll
            for (int e = 0; e < 8; ++e) ix.put(sre, sim, u + e * (N / 8), vr[e], vi[e]);
            bar();
            const long long obase = (long long)o * a.out_os + (zc - a.out_zc0) * ZC + 2 * f;
            const double sc = a.scale, hs = 0.5 * a.scale;
#pragma unroll 4
            for (int e = 0; e < 4; ++e) {
                const int k = ul + e * (N / 8);
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
                    st2f(a.final_store, row_dst(a, k) + obase, (p + r) * hs, (q + s) * hs);
                    st2f(a.final_store, row_dst(a, N - k) + obase, (q - s) * hs, (r - p) * hs);
                }
                if (ROT && e == 3) {             // every read of the buffer is done: refill it
                    bar();
                    if (t == 0 && i + NB < nseq) issue(i + NB);
                }
            }
        } else {
            block_cfft<N, true>(vr, vi, u, true, sre, sim, ix, TwGlobal{a.tw, a.twscale}, bar);
            if (ROT) {                           // the last gather is behind every thread: refill the buffer
                bar();
                if (t == 0 && i + NB < nseq) issue(i + NB);
            }
            const long long obase = (long long)o * a.out_os + (zc - a.out_zc0) * ZC + 2 * f;
            const double sc = a.scale;
#pragma unroll
            for (int e = 0; e < 8; ++e)
                st2f(a.final_store, row_dst(a, ul + e * (N / 8)) + obase, vr[e] * sc, vi[e] * sc);
        }
        // (the barrier after the landing copy of the group's next tile orders this tile's last scratch reads before
        //  the next tile's first scratch writes)
    }
    if (a.out_map.self >= 0 && a.scatter_fence) __threadfence_system();     // peer-memory scatter: stores visible to the owner GPU
}

}  // namespace ps3d
#endif
