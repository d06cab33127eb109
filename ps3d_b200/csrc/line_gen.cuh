// x/y sweeps for line lengths that are not a power of two (n = 2^a 3^b 5^c, even): the same tile (16 z x n rows),
// the same two-real-lines-per-complex-FFT packing, prologues (diffx/diffy, cross products) and output packing as
// line_fft.cuh (reference sta3dfft.f90:136-260, 304-377; stafft.f90:55-58), with the mixed-radix transform of
// gen_fft.cuh looping through shared memory instead of the register-blocked radix-8 passes.  Coverage path for
// the grids the reference's factorisen accepts (slab blocks of any size: the row maps divide instead of shifting).
#pragma once

#include "gen_fft.cuh"
#include "line_fft.cuh"

namespace ps3d {

constexpr int GEN_THREADS = 256;

struct GenLine {
    GenPlan plan;
    const double2* tw;      // [n] exp(2 pi i m / n)
};

inline size_t line_gen_smem_bytes(int n) { return (size_t)2 * LINE_NF * n * sizeof(double2); }

template <int PRO>
__global__ void __launch_bounds__(GEN_THREADS) k_line_gen_fwd(LineArgs a, GenLine gl) {
    PS_SMEM(double, sm);
    const int n = gl.plan.n;
    double2* A = reinterpret_cast<double2*>(sm);
    double2* B = A + LINE_NF * n;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        const int o = tile / a.nzc, zc = a.zc0 + (tile - o * a.nzc);
        const long long ibase = (long long)o * a.in_os + (zc - a.in_zc0) * LINE_ZC;
        for (int w = threadIdx.x; w < LINE_NF * n; w += blockDim.x) {
            const int f = w & (LINE_NF - 1), row = w / LINE_NF;
            const long long off = ibase + 2 * f + row_off_div(a.in_map, row);
            double2 v;
            if (PRO == PRO_CROSS) {
                const double2 x0 = ld2(a.in0 + off), x1 = ld2(a.in1 + off), x2 = ld2(a.in2 + off), x3 = ld2(a.in3 + off);
                v.x = x0.x * (x1.x + a.add1) - x2.x * (x3.x + a.add3);
                v.y = x0.y * (x1.y + a.add1) - x2.y * (x3.y + a.add3);
            } else {
                v = ld2(a.in0 + off);
            }
            A[w] = v;                                   // element `row` of FFT f at row * NF + f
        }
        __syncthreads();
        const double2* C = gen_cfft<false>(A, B, LINE_NF, gl.plan, gl.tw);
        const long long obase = (long long)o * a.out_os + (zc - a.out_zc0) * LINE_ZC;
        const double sc = a.scale, hs = 0.5 * a.scale;
        const int h = n / 2;
        for (int w = threadIdx.x; w < LINE_NF * (h + 1); w += blockDim.x) {
            const int f = w & (LINE_NF - 1), k = w / LINE_NF;
            const long long ob = obase + 2 * f;
            if (k == 0) {
                const double2 c0 = C[f];
                st2f(a.final_store, row_dst_div(a, 0) + ob, c0.x * sc, c0.y * sc);
            } else if (k == h) {
                const double2 ch = C[h * LINE_NF + f];
                st2f(a.final_store, row_dst_div(a, h) + ob, ch.x * sc, ch.y * sc);
            } else {
                const double2 ck = C[k * LINE_NF + f], cm = C[(n - k) * LINE_NF + f];
                // A_k = (C_k + conj C_{n-k})/2, B_k = (C_k - conj C_{n-k})/(2i)
                st2f(a.final_store, row_dst_div(a, k) + ob, (ck.x + cm.x) * hs, (ck.y + cm.y) * hs);       // Re A, Re B
                st2f(a.final_store, row_dst_div(a, n - k) + ob, (ck.y - cm.y) * hs, (cm.x - ck.x) * hs);   // Im A, Im B
            }
        }
        __syncthreads();          // the buffers are reused by the next tile
    }
    if (a.out_map.self >= 0 && a.scatter_fence) __threadfence_system();     // peer-memory scatter (see line_fft.cuh)
}

template <int PRO>
__global__ void __launch_bounds__(GEN_THREADS) k_line_gen_inv(LineArgs a, GenLine gl) {
    PS_SMEM(double, sm);
    const int n = gl.plan.n;
    double2* A = reinterpret_cast<double2*>(sm);
    double2* B = A + LINE_NF * n;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        const int o = tile / a.nzc, zc = a.zc0 + (tile - o * a.nzc);
        const long long ibase = (long long)o * a.in_os + (zc - a.in_zc0) * LINE_ZC;
        const int h = n / 2;
        for (int w = threadIdx.x; w < LINE_NF * (h + 1); w += blockDim.x) {
            const int f = w & (LINE_NF - 1), k = w / LINE_NF;
            const long long ib = ibase + 2 * f;
            if (k == 0 || k == h) {
                // real DC / Nyquist terms of both lines; d/dx: kappa = 0 there (sta3dfft.f90:321-335)
                double2 c = ld2(a.in0 + ib + row_off_div(a.in_map, k));
                if (PRO == PRO_DIFF) c = make_double2(0.0, 0.0);
                A[k * LINE_NF + f] = c;
            } else {
                const double2 xa = ld2(a.in0 + ib + row_off_div(a.in_map, k)), xb = ld2(a.in0 + ib + row_off_div(a.in_map, n - k));
                double Ar = xa.x, Ai = xb.x, Br = xa.y, Bi = xb.y;
                if (PRO == PRO_DIFF) {
                    // d/dx: X_k -> i kappa X_k (sta3dfft.f90:325-329)
                    const double kap = __ldg(&a.kdiff[k]);
                    const double ar = -kap * Ai, ai = kap * Ar, br = -kap * Bi, bi = kap * Br;
                    Ar = ar; Ai = ai; Br = br; Bi = bi;
                }
                // C_k = A + i B, C_{n-k} = conj(A) + i conj(B)
                A[k * LINE_NF + f] = make_double2(Ar - Bi, Ai + Br);
                A[(n - k) * LINE_NF + f] = make_double2(Ar + Bi, Br - Ai);
            }
        }
        __syncthreads();
        const double2* C = gen_cfft<true>(A, B, LINE_NF, gl.plan, gl.tw);
        const long long obase = (long long)o * a.out_os + (zc - a.out_zc0) * LINE_ZC;
        const double sc = a.scale;
        for (int w = threadIdx.x; w < LINE_NF * n; w += blockDim.x) {
            const int f = w & (LINE_NF - 1), row = w / LINE_NF;
            const double2 c = C[w];
            st2f(a.final_store, row_dst_div(a, row) + obase + 2 * f, c.x * sc, c.y * sc);
        }
        __syncthreads();
    }
    if (a.out_map.self >= 0 && a.scatter_fence) __threadfence_system();
}

}  // namespace ps3d
