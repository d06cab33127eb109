// libps3d_cuda: context, host-side orchestration and the C ABI (include/ps3d_cuda.h).
//
// Host logic restated from the reference (file:line under /root/reference/src):
//   init          <- mpi/mpi_layout.f90:53, utils/parameters.f90:61, fft/sta3dfft.f90:53
//   init_inversion<- inversion/inversion_utils.f90:222-455 (2-D / 1-D tables only)
//   vor2vel/source<- inversion/inversion.f90:23-226, 298-388
//   adapt         <- stepper/advance.f90:109-410, utils/rolling_mean.f90:36-69
//   steppers      <- stepper/cn2.f90:40-181, stepper/impl_rk4.f90:37-207
#include <cstdarg>
#include <map>
#include <vector>
#include <algorithm>
#include <cmath>

#include "../../include/ps3d_cuda.h"
#include "rt.h"
#include "fft_core.cuh"
#include "line_fft.cuh"
#include "line_tma.cuh"
#include "zcol.cuh"
#include "line_gen.cuh"
#include "zcol_gen.cuh"
#include "pointwise.cuh"
#include "transport.h"

namespace ps3d {

std::string g_last_error;
void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

#ifndef PS3D_EMU
// ---- PS3D_TRACE: in-situ kernel times (rt.h) ----
int g_trace = 0;
struct TraceRec { const char* name; cudaEvent_t a, b; };
static std::vector<TraceRec> g_trace_recs;
void ps_trace_begin(const char* name, cudaStream_t s) {
    TraceRec r{name, nullptr, nullptr};
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    cudaEventRecord(r.a, s);
    g_trace_recs.push_back(r);
}
void ps_trace_end(cudaStream_t s) { cudaEventRecord(g_trace_recs.back().b, s); }
static void trace_dump() {
    if (!g_trace || g_trace_recs.empty()) return;
    cudaDeviceSynchronize();
    std::map<std::string, std::pair<int, double>> agg;
    double tot = 0.0;
    for (auto& r : g_trace_recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { auto& a = agg[r.name]; a.first++; a.second += ms; tot += ms; }
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    std::vector<std::pair<std::string, std::pair<int, double>>> v(agg.begin(), agg.end());
    std::sort(v.begin(), v.end(), [](auto& x, auto& y) { return x.second.second > y.second.second; });
    fprintf(stderr, "PS3D_TRACE: %zu launches, %.3f ms between event pairs\n", g_trace_recs.size(), tot);
    for (auto& e : v)
        fprintf(stderr, "PS3D_TRACE %-60s n=%6d total=%10.3f ms avg=%9.4f ms (%5.1f%%)\n", e.first.c_str(), e.second.first,
                e.second.second, e.second.second / e.second.first, 100.0 * e.second.second / tot);
    g_trace_recs.clear();
}
#endif

struct StatusError { int code; };
[[noreturn]] static void fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    throw StatusError{code};
}

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    void alloc(size_t count) { release(); p = (T*)ps_malloc(count * sizeof(T)); n = count; }
    void release() { if (p) ps_free(p); p = nullptr; n = 0; }
    void upload(const std::vector<T>& h, ps_stream_t s) {
        if (h.size() != n) alloc(h.size());
        ps_h2d(p, h.data(), h.size() * sizeof(T), s);
        ps_sync(s);
    }
};

static const int RED_BLOCKS = 1184;   // 148 SMs x 8 (upper bound; small grids use one block per 256 elements)
static const int RED_THREADS = 256;

struct RollingMean {   // utils/rolling_mean.f90:36-69
    std::vector<double> history;
    int inew = 1, iold = 1, length = 0;
    double sma = 0.0;
    bool filled = false;
    // allocated once (advance.f90: rollmean%alloc in the stepper set-up); the window cannot change afterwards
    bool alloc(int n) { if (history.empty()) { history.assign(n, 0.0); length = n; } return n == length; }
    double get_next(double vnew) {
        if (filled) {
            const double vold = history[iold - 1];
            iold = iold % length + 1;
            sma = sma + (vnew - vold) / (double)length;
            history[inew - 1] = vnew;
            inew = inew % length + 1;
        } else {
            history[inew - 1] = vnew;
            double s = 0.0;
            for (int i = 0; i < inew; ++i) s += history[i];
            sma = s / (double)inew;
            filled = (length == inew);
            inew = inew % length + 1;
        }
        return sma;
    }
};

struct Ctx {
    int nx = 0, ny = 0, nz = 0, nzp = 0, pz = 0;
    int rank = 0, nranks = 1, nxl = 0, nyl = 0;
    double lower[3], extent[3], upper[3], dx[3];
    long long ncell = 0;
    size_t nint = 0;     // doubles per internal field (nxl*ny*pz == nx*nyl*pz)
    size_t nnat = 0;     // doubles per host field on this rank (nxl*ny*nzp)
    ps_stream_t stream = 0;
    bool inversion_ready = false, diffusion_ready = false;
    int filtering = 0, nnu = 3, stepper = PS3D_STEPPER_CN2;
    bool stepper_ready = false;
    double vvisc = 0.0;
    RollingMean rollmean;
    long long launches = 0;
                                         // (measured slower on B200: 87.7 vs 77.3 ms/step at 512^3 cn2 -> off by default)
    int red_blocks = RED_BLOCKS;         // blocks of the two-stage reductions: fixed per grid size -> deterministic sums
    int num_sms = 148;
    int p2p_ctas_per_sm = -1;            // PS3D_P2P_CTAS: blocks per SM of the persistent scatter sweeps (0 = full grid, -1 = auto)
    int scatter_fence = 1;               // PS3D_NO_SCATTER_FENCE=1: rely on kernel completion for the visibility of peer stores
    int fuse_update = 1;                 // PS3D_NO_FUSED_UPDATE=1: cn2 update as a separate kernel after the source kernel
    int l2_chunks = 0;                   // PS3D_L2_CHUNKS: z-chunks per launch of the L2-blocked 2-D FFT (0 = off)
    int keep_velx = 1;                   // PS3D_NO_KEEP_VELX=1: adapt re-does the x sweeps of u, v, w for its d/dy fields
    int strict_jacobi = 0;               // PS3D_STRICT_JACOBI=1: literal cyclic Jacobi (jacobi.f90) instead of the closed form
#ifndef PS3D_EMU
    // TMA-staged line sweeps (line_tma.cuh): PS3D_LINE_TMA=0 falls back to the register-staged sweeps of line_fft.cuh,
    // 1 (default on one rank): lines of 512 and 1024, 2: every length from 128
    int line_tma = 1;
    int tma_rot = 1;                                   // PS3D_TMA_ROT=0: separate landing buffer instead of rotating buffers
    int tma_zc8 = 0;                                   // PS3D_TMA_ZC=8: N = 512 sweeps as four groups of 8-z tiles
    void* encode_tiled = nullptr;                      // cuTensorMapEncodeTiled through the runtime's driver entry point
    std::map<std::pair<const void*, int>, CUtensorMap> tmaps;    // (input array, sweep kind) -> tensor map
    long long tma_launches = 0;
#endif
    double rk4_dfac = 0.0;               // impl-diff-rk4: 0.5 * pref * dt of the last set_diffusion (vdiss = rk4_dfac * vhdis)
    bool rk4_dfac_set = false;
    double last_advance_ms = 0.0;
    double last_diag[16] = {};           // diagnostics of the last adapt (advance.f90: set_netcdf_field_diagnostic)
    bool have_diag = false;
    bool svorts_stale = false;           // the last source call carried a stepper update and did not store svorts

    DevBuf<double> svor[3], vor[3], vel[3], svel[3], svorts[3], wa[3], wb[3], W[9];
    // one rank: the x-transformed velocity of the last vor2vel ([x][ky'][pz], the intermediate between the two sweeps
    // of fftxys2p) is kept: adapt's d/dy fields (advance.f90:199-217) then need only their y sweep
    DevBuf<double> velx[3];
    bool velx_valid = false;
    // physics.f90:88-92, 163-169: planetary vorticity and squared buoyancy frequency (ps3d_cuda_set_physics)
    double f_cor[3] = {0.0, 0.0, 0.0};
    double bfsq = 0.0;
    // ENABLE_BUOYANCY build (configure.ac:228-245), switched on at run time by ps3d_cuda_enable_buoyancy:
    // sbuoy (mixed spectral b'), buoy (physical), sbuoys (tendency), bsm / sbuoyi, sbuoyf, bsem (semi-spectral b')
    bool buoyancy = false;
    DevBuf<double> sbuoy, buoy, sbuoys, bsm, sbuoyf, bsem;
    DevBuf<double> bhdis, bfac1, bfac2;
    int bnnu = 3, bpretype = -1, bwin = 0;
    double bvisc = 0.0, rk4_dbac = 0.0;
    bool bdiffusion_ready = false;
    RollingMean buoy_rollmean;
    double last_bdiag[4] = {};          // bfmax, rmb, bval
    Transport tr;
    ps_stream_t comm_stream = 0;          // NCCL all-to-alls run here, overlapped with the sweeps of other fields
#ifndef PS3D_EMU
    cudaEvent_t ev_first[8] = {}, ev_a2a[8] = {}, ev_second[8] = {};
#endif
    DevBuf<double> redS, redM;            // all-reduce landing buffers (sum / max)
    PeerMailPtrs mail;                    // every rank's mailbox (peer-mapped), valid when tr.p2p
    DevBuf<double> stage;                 // natural-layout staging for the host boundary
    // streamed upload (ps3d_cuda_upload_vorticity_begin / _end): three staging fields filled by a copy stream while
    // the compute stream works on the previous state
    DevBuf<double> stage3[2];                        // two staging sets: a second upload may be queued while the first waits for _end
    ps_stream_t copy_stream = 0;
    int upload_head = 0, upload_pending = 0;         // FIFO of queued uploads (depth <= 2)
#ifndef PS3D_EMU
    cudaEvent_t ev_copy[2] = {nullptr, nullptr};
#endif
    DevBuf<double> kxl, kyline, kxd, kyd, k2l2, k2l2i, zm, zp, rkz, gamtop, gambot;
    DevBuf<double> filt2d, filtz, vhdis, fac1, fac2, wz, ini_mean, partial, red;
    DevBuf<double2> tw;
    // lengths that are not a power of two (line_gen.cuh, zcol_gen.cuh): per-axis plan and tables
    bool gen[3] = {false, false, false};            // x, y, z
    GenPlan plan[3];
    DevBuf<double2> gtw[3];                         // [n] exp(2 pi i m / n)
    DevBuf<double> gsinz, gcosz;                    // [nz+1] sin, cos(pi j / nz)
    DevBuf<int> permy;
    int ntw = 0;
    std::vector<double> h_rkx, h_rky, h_rkz, h_k2l2, h_filt2d;   // host copies ([kx][kyl] local order)
    std::vector<int> h_ky_of_kyl;
    double* h_red = nullptr;              // host landing buffer for reductions
#ifndef PS3D_EMU
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
#endif

    SpecGeom geom() const {
        SpecGeom g;
        g.nx = nx; g.nyl = nyl; g.nz = nz; g.pz = pz;
        g.has00 = (rank == 0);
        g.kxd = kxd.p; g.kyd = kyd.p; g.k2l2 = k2l2.p; g.k2l2i = k2l2i.p;
        g.zm = zm.p; g.zp = zp.p; g.rkz = rkz.p; g.gamtop = gamtop.p; g.gambot = gambot.p;
        g.Lz = extent[2]; g.dzi = 1.0 / dx[2]; g.hdzi = 0.5 * (1.0 / dx[2]);
        g.tw = tw.p; g.ntw = ntw;
        g.ap0 = 0; g.npf = nyl / 2;               // all groups (operator mode)
        return g;
    }
    int ngroups() const { return (nx / 2 + 1) * (nyl / 2); }
    GenZ genz() const { GenZ z; z.plan = plan[2]; z.tw = gtw[2].p; z.sinz = gsinz.p; z.cosz = gcosz.p; return z; }
    // the hot-loop column kernels run as two launches: the fast instantiation on the pairs (a, ny-a), a >= 1,
    // and the general one on the pair (0, ny/2) of the rank that owns ky = 0 (zcol.cuh)
    SpecGeom geom_fast() const { SpecGeom g = geom(); g.ap0 = (rank == 0) ? 1 : 0; g.npf = nyl / 2 - g.ap0; return g; }
    SpecGeom geom_gen() const { SpecGeom g = geom(); g.ap0 = 0; g.npf = 1; return g; }
};

static Ctx* g_ctx = nullptr;

static Ctx& ctx() {
    if (!g_ctx) fail(PS3D_ERR_NOT_INITIALISED, "ps3d_cuda_init has not been called");
    return *g_ctx;
}

static bool pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

// ---------------------------------------------------------------------------
// kernel launch helpers with compile-time size dispatch
// ---------------------------------------------------------------------------
#ifndef PS3D_EMU
template <class K>
static void allow_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024)
        PS_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}
#else
template <class K>
static void allow_smem(K, size_t) {}
#endif

#define PS_FOR_LINE_SIZES(X) X(8) X(16) X(32) X(64) X(128) X(256) X(512) X(1024)
#define PS_FOR_Z_SIZES(X) X(8) X(16) X(32) X(64) X(128) X(256) X(512) X(1024)

template <int N, int ZC>
static void launch_line_nz(Ctx& c, bool inv, int pro, const LineArgs& a, int ntiles, ps_stream_t stream, int max_ctas) {
    const size_t sm = line_smem_bytes<N, ZC>();
    const dim3 grid(max_ctas > 0 ? std::min(ntiles, max_ctas) : ntiles), block(line_threads(N, ZC));
    if (!inv) {
        if (pro == PRO_CROSS) { allow_smem(k_line_fwd<N, PRO_CROSS, ZC>, sm); PS_LAUNCH((k_line_fwd<N, PRO_CROSS, ZC>), grid, block, sm, stream, a); }
        else { allow_smem(k_line_fwd<N, PRO_PLAIN, ZC>, sm); PS_LAUNCH((k_line_fwd<N, PRO_PLAIN, ZC>), grid, block, sm, stream, a); }
    } else {
        if (pro == PRO_DIFF) { allow_smem(k_line_inv<N, PRO_DIFF, ZC>, sm); PS_LAUNCH((k_line_inv<N, PRO_DIFF, ZC>), grid, block, sm, stream, a); }
        else { allow_smem(k_line_inv<N, PRO_PLAIN, ZC>, sm); PS_LAUNCH((k_line_inv<N, PRO_PLAIN, ZC>), grid, block, sm, stream, a); }
    }
    ++c.launches;
}

// z values per tile of the sweep kernel for line length n: line_zc(n), except 16 for peer-memory scatter sweeps
static int sweep_zc(int n, bool scatter) { return scatter ? LINE_ZC : line_zc(n); }

template <int N>
static void launch_line_n(Ctx& c, bool inv, int pro, const LineArgs& a, int ntiles, ps_stream_t stream, int max_ctas) {
    if (line_zc(N) != LINE_ZC && a.out_map.self >= 0) launch_line_nz<N, LINE_ZC>(c, inv, pro, a, ntiles, stream, max_ctas);
    else launch_line_nz<N, line_zc(N)>(c, inv, pro, a, ntiles, stream, max_ctas);
}

// line lengths that are not a power of two (coverage path, one rank)
static void launch_line_gen(Ctx& c, int axis, bool inv, int pro, const LineArgs& a, int ntiles, ps_stream_t stream, int max_ctas) {
    GenLine gl;
    gl.plan = c.plan[axis]; gl.tw = c.gtw[axis].p;
    const size_t sm = line_gen_smem_bytes(gl.plan.n);
    const dim3 grid(std::min(ntiles, max_ctas > 0 ? max_ctas : 4 * c.num_sms)), block(GEN_THREADS);
    if (!inv) {
        if (pro == PRO_CROSS) { allow_smem(k_line_gen_fwd<PRO_CROSS>, sm); PS_LAUNCH((k_line_gen_fwd<PRO_CROSS>), grid, block, sm, stream, a, gl); }
        else { allow_smem(k_line_gen_fwd<PRO_PLAIN>, sm); PS_LAUNCH((k_line_gen_fwd<PRO_PLAIN>), grid, block, sm, stream, a, gl); }
    } else {
        if (pro == PRO_DIFF) { allow_smem(k_line_gen_inv<PRO_DIFF>, sm); PS_LAUNCH((k_line_gen_inv<PRO_DIFF>), grid, block, sm, stream, a, gl); }
        else { allow_smem(k_line_gen_inv<PRO_PLAIN>, sm); PS_LAUNCH((k_line_gen_inv<PRO_PLAIN>), grid, block, sm, stream, a, gl); }
    }
    ++c.launches;
}

static void launch_line(Ctx& c, int n, bool inv, int pro, const LineArgs& a, int ntiles, ps_stream_t stream, int max_ctas) {
    switch (n) {
#define X(NN) case NN: launch_line_n<NN>(c, inv, pro, a, ntiles, stream, max_ctas); break;
        PS_FOR_LINE_SIZES(X)
#undef X
        default: fail(PS3D_ERR_UNSUPPORTED_SIZE, "line length %d not supported (power of two, 8..1024)", n);
    }
}

#ifndef PS3D_EMU
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Tensor map of a sweep input (line_tma.cuh).  kind 0: y sweep over a physical array [xl][y][pz]; 1: y sweep over
// slab blocks [d][xl][kyl][pz] (one rank: one block, kyl = ny); 2: x sweep over [x][kyl][pz].  Returns null (and
// switches the TMA path off) if the driver refuses the descriptor.
static const CUtensorMap* sweep_tmap(Ctx& c, const double* base, int kind, int zc, TmaArgs& ta) {
    const int n = (kind == 2) ? c.nx : c.ny;
    const int blkrows = (kind == 1) ? c.nyl : n;
    const int rpo = std::min(256, (kind == 2) ? n : blkrows);
    ta.mode = (kind == 2) ? 1 : 0;
    ta.rows_per_op = rpo; ta.nops = n / rpo; ta.blkrows = blkrows;
    auto key = std::make_pair((const void*)base, kind * 100 + zc);
    auto it = c.tmaps.find(key);
    if (it != c.tmaps.end()) return &it->second;
    cuuint64_t dim[4], str[3];
    cuuint32_t box[4], es[4] = {1, 1, 1, 1};
    const cuuint64_t rowb = (cuuint64_t)c.pz * 8;
    dim[0] = c.pz;
    if (kind == 2) {
        dim[1] = c.nyl; dim[2] = c.nx; dim[3] = 1;
        str[0] = rowb; str[1] = rowb * c.nyl; str[2] = rowb * c.nyl * c.nx;
        box[0] = zc; box[1] = 1; box[2] = rpo; box[3] = 1;
    } else {
        const int nblk = n / blkrows;
        dim[1] = blkrows; dim[2] = c.nxl; dim[3] = nblk;
        str[0] = rowb; str[1] = rowb * blkrows; str[2] = rowb * blkrows * c.nxl;
        box[0] = zc; box[1] = rpo; box[2] = 1; box[3] = 1;
    }
    alignas(64) CUtensorMap tm;
    const CUresult rc = ((EncodeTiledFn)c.encode_tiled)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void*)base, dim, str, box, es,
                                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        fprintf(stderr, "libps3d_cuda: cuTensorMapEncodeTiled failed (%d) for sweep kind %d: register-staged sweeps from here on\n", (int)rc, kind);
        c.line_tma = 0;
        return nullptr;
    }
    return &c.tmaps.emplace(key, tm).first->second;
}

template <int N, int ZC, bool ROT>
static void launch_line_tma_r(Ctx& c, bool inv, int pro, const LineArgs& a, const TmaArgs& ta, const CUtensorMap& tm, int nctas,
                              ps_stream_t stream) {
    const size_t sm = line_tma_smem_bytes<N, ZC>();
    const dim3 grid(nctas), block(1024);
    if (!inv) { allow_smem(k_line_tma<N, false, PRO_PLAIN, ZC, ROT>, sm); PS_LAUNCH((k_line_tma<N, false, PRO_PLAIN, ZC, ROT>), grid, block, sm, stream, a, ta, tm); }
    else if (pro == PRO_DIFF) { allow_smem(k_line_tma<N, true, PRO_DIFF, ZC, ROT>, sm); PS_LAUNCH((k_line_tma<N, true, PRO_DIFF, ZC, ROT>), grid, block, sm, stream, a, ta, tm); }
    else { allow_smem(k_line_tma<N, true, PRO_PLAIN, ZC, ROT>, sm); PS_LAUNCH((k_line_tma<N, true, PRO_PLAIN, ZC, ROT>), grid, block, sm, stream, a, ta, tm); }
    ++c.launches; ++c.tma_launches;
}
template <int N, int ZC>
static void launch_line_tma_n(Ctx& c, bool inv, int pro, const LineArgs& a, const TmaArgs& ta, const CUtensorMap& tm, int nctas,
                              ps_stream_t stream) {
    if (c.tma_rot) launch_line_tma_r<N, ZC, true>(c, inv, pro, a, ta, tm, nctas, stream);
    else launch_line_tma_r<N, ZC, false>(c, inv, pro, a, ta, tm, nctas, stream);
}
#endif

// one x or y sweep.  axis: 0 = x, 1 = y.
struct Sweep { int axis; bool inv; int pro; const double* in[4]; double add1, add3; double* out;
               int scatter = -1;      // >= 0: store straight into the peers' receive buffer `scatter` (peer memory)
               // z-chunk range of this launch and the compact (L2-resident) intermediate of a chunked 2-D FFT:
               int zc0 = 0, nzc = -1;            // chunks [zc0, zc0 + nzc) of pz/16 (nzc < 0: all)
               int in_pitch = 0, out_pitch = 0;  // doubles per column of the input / output array (0: pz)
               int in_zc0 = 0, out_zc0 = 0;      // chunk stored at offset 0 of the input / output array
               int final_store = 1;
               bool on_comm_stream = false;      // launch on the communication stream (peer-memory scatter sweeps)
               int max_ctas = 0; };              // > 0: persistent launch with at most this many blocks

static void run_sweep(Ctx& c, const Sweep& s) {
    LineArgs a;
    a.in0 = s.in[0]; a.in1 = s.in[1]; a.in2 = s.in[2]; a.in3 = s.in[3];
    a.add1 = s.add1; a.add3 = s.add3;
    a.out = s.out;
    const int nline = (s.axis == 1) ? c.ny : c.nx;
    int zcl = c.gen[s.axis] ? LINE_ZC : sweep_zc(nline, s.scatter >= 0);     // z values per tile of this sweep's kernel
#ifndef PS3D_EMU
    // TMA-staged sweep (line_tma.cuh): whole arrays, plain or derivative prologue, tiles of at most 64 KB
    const bool tma = c.line_tma && !c.gen[s.axis] && s.pro != PRO_CROSS && s.nzc < 0 && !s.in_pitch && !s.out_pitch &&
                     nline >= (c.line_tma >= 2 ? 128 : 512) && (long long)nline * zcl * 8 <= 65536;   // (measured: no gain below 512)
    if (tma && c.tma_zc8 && nline == 512 && s.scatter < 0) zcl = 8;            // PS3D_TMA_ZC=8: four groups of 8-z tiles
#endif
    a.nzc = (s.nzc < 0) ? c.pz / zcl : s.nzc;
    a.zc0 = s.zc0; a.in_zc0 = s.in_zc0; a.out_zc0 = s.out_zc0; a.final_store = s.final_store;
    a.scatter_fence = c.scatter_fence;
    const long long ipz = s.in_pitch ? s.in_pitch : c.pz, opz = s.out_pitch ? s.out_pitch : c.pz;
    a.tw = c.tw.p;
    int n, nouter;
    int lognyl = 0, lognxl = 0;
    while ((1 << lognyl) < c.nyl) ++lognyl;
    while ((1 << lognxl) < c.nxl) ++lognxl;
    const long long nbi = (long long)c.nxl * c.nyl * ipz, nbo = (long long)c.nxl * c.nyl * opz;
    const int self = (s.scatter >= 0) ? c.rank : -1;
    for (int d = 0; d < 8; ++d) a.outp[d] = (s.scatter >= 0 && d < c.nranks) ? c.tr.peer_t2[s.scatter][d] : s.out;
    if (s.axis == 1) {
        n = c.ny; nouter = c.nxl;
        // physical side [xl][y][pz]; spectral side: P blocks [d][xl][kyl][pz] (== [kx][kyl][pz] after the exchange)
        a.in_os = (s.inv ? (long long)c.nyl : (long long)c.ny) * ipz;
        a.out_os = (s.inv ? (long long)c.ny : (long long)c.nyl) * opz;
        const RowMap phys_in{ipz, 0, 0, n, 30, -1, 1 << 30}, phys_out{opz, 0, 0, n, 30, -1, 1 << 30};
        const RowMap spec_in{ipz, nbi, 1, n, lognyl, -1, c.nyl};
        const RowMap spec_out{opz, nbo, 1, n, lognyl, self, c.nyl};
        a.in_map = s.inv ? spec_in : phys_in;
        a.out_map = s.inv ? phys_out : spec_out;
        a.kdiff = c.kyline.p;
    } else {
        n = c.nx; nouter = c.nyl;
        a.in_os = ipz; a.out_os = opz;
        const RowMap xin{(long long)c.nyl * ipz, 0, 0, n, 30, -1, 1 << 30};
        const RowMap xout{(long long)c.nyl * opz, nbo, 0, n, lognxl, self, c.nxl};      // block d = x / nxl
        a.in_map = xin; a.out_map = xout;
        a.kdiff = c.kxl.p;
    }
    a.scale = 1.0 / std::sqrt((double)n);
    a.twscale = c.ntw / n;
    a.ntiles = nouter * a.nzc;
#ifndef PS3D_EMU
    if (tma) {
        TmaArgs ta;
        const int kind = (s.axis == 0) ? 2 : (s.inv ? 1 : 0);
        const CUtensorMap* tm = sweep_tmap(c, s.in[0], kind, zcl, ta);
        if (tm) {
            const int nctas = std::min(a.ntiles, s.max_ctas > 0 ? std::min(s.max_ctas, c.num_sms) : c.num_sms);
            ps_stream_t st = s.on_comm_stream ? c.comm_stream : c.stream;
            switch (n) {
                case 128: launch_line_tma_n<128, 16>(c, s.inv, s.pro, a, ta, *tm, nctas, st); return;
                case 256: launch_line_tma_n<256, 16>(c, s.inv, s.pro, a, ta, *tm, nctas, st); return;
                case 512:
                    if (zcl == 8) launch_line_tma_n<512, 8>(c, s.inv, s.pro, a, ta, *tm, nctas, st);
                    else launch_line_tma_n<512, 16>(c, s.inv, s.pro, a, ta, *tm, nctas, st);
                    return;
                case 1024: launch_line_tma_n<1024, 8>(c, s.inv, s.pro, a, ta, *tm, nctas, st); return;
                default: break;
            }
        }
    }
#endif
    if (c.gen[s.axis]) launch_line_gen(c, s.axis, s.inv, s.pro, a, a.ntiles, s.on_comm_stream ? c.comm_stream : c.stream, s.max_ctas);
    else launch_line(c, n, s.inv, s.pro, a, a.ntiles, s.on_comm_stream ? c.comm_stream : c.stream, s.max_ctas);
}

// ---- slab exchange: P equal contiguous blocks, block d of `send` goes to rank d and lands as block
// `rank` of its `recv` (replaces the four transpose_to_pencil calls per 2-D FFT of the reference) ----
static void exchange(Ctx& c, const double* send, double* recv, ps_stream_t stream) {
    Transport& t = c.tr;
    const size_t nb = (size_t)c.nxl * c.nyl * c.pz;         // doubles per block
    ++t.n_alltoall;
    t.bytes_sent += (double)nb * 8.0 * (t.nranks - 1);
    if (t.a2a_cb) {
        ps_sync(stream);
        if (t.a2a_cb(send, recv, nb * sizeof(double), t.user) != 0) fail(PS3D_ERR_DEVICE, "all-to-all callback failed");
        return;
    }
#ifndef PS3D_EMU
    if (t.have_nccl()) {
        int rc = t.nccl.GroupStart();
        for (int d = 0; d < t.nranks && rc == 0; ++d) {
            rc = t.nccl.Send(send + (size_t)d * nb, nb, NcclApi::kFloat64, d, t.comm, (void*)stream);
            if (rc == 0) rc = t.nccl.Recv(recv + (size_t)d * nb, nb, NcclApi::kFloat64, d, t.comm, (void*)stream);
        }
        const int rc2 = t.nccl.GroupEnd();
        if (rc != 0 || rc2 != 0) fail(PS3D_ERR_DEVICE, "NCCL all-to-all failed: %s", t.nccl.GetErrorString(rc ? rc : rc2));
        return;
    }
#endif
    fail(PS3D_ERR_NOT_INITIALISED, "nranks > 1 but neither an NCCL id nor a transport callback was given");
}

// small all-reduce of host values: vals[i] <- sum or max over ranks according to bit i of opmask
static void allreduce_host(Ctx& c, double* vals, int n, unsigned opmask) {
    Transport& t = c.tr;
    if (t.nranks == 1) return;
    double s[64], m[64];
    for (int i = 0; i < n; ++i) { s[i] = vals[i]; m[i] = vals[i]; }
    if (t.ar_cb) {
        if (t.ar_cb(s, n, 0, t.user) != 0 || t.ar_cb(m, n, 1, t.user) != 0) fail(PS3D_ERR_DEVICE, "all-reduce callback failed");
    } else {
#ifndef PS3D_EMU
        if (!t.have_nccl()) fail(PS3D_ERR_NOT_INITIALISED, "no transport for the all-reduce");
        ps_h2d(c.redS.p, s, n * sizeof(double), c.stream);
        ps_h2d(c.redM.p, m, n * sizeof(double), c.stream);
        int rc = t.nccl.AllReduce(c.redS.p, c.redS.p, n, NcclApi::kFloat64, NcclApi::kSum, t.comm, (void*)c.stream);
        if (rc == 0) rc = t.nccl.AllReduce(c.redM.p, c.redM.p, n, NcclApi::kFloat64, NcclApi::kMax, t.comm, (void*)c.stream);
        if (rc != 0) fail(PS3D_ERR_DEVICE, "NCCL all-reduce failed: %s", t.nccl.GetErrorString(rc));
        ps_d2h(s, c.redS.p, n * sizeof(double), c.stream);
        ps_d2h(m, c.redM.p, n * sizeof(double), c.stream);
        ps_sync(c.stream);
#else
        fail(PS3D_ERR_NOT_INITIALISED, "no transport for the all-reduce");
#endif
    }
    for (int i = 0; i < n; ++i) vals[i] = ((opmask >> i) & 1) ? m[i] : s[i];
}

#ifndef PS3D_EMU
// all ranks' preceding work on the compute stream is complete (and its peer stores visible) before any rank continues
static void cross_rank_barrier(Ctx& c, ps_stream_t stream) {
    if (c.tr.p2p && c.tr.peer_sync) {
        PS_LAUNCH((k_peer_barrier), dim3(1), dim3(32), 0, stream, c.mail, c.rank, c.nranks, ++c.tr.bar_epoch);
        ++c.launches;
        return;
    }
    const int rc = c.tr.nccl.AllReduce(c.redM.p + 32, c.redM.p + 32, 1, NcclApi::kFloat64, NcclApi::kSum, c.tr.comm, (void*)stream);
    if (rc != 0) fail(PS3D_ERR_DEVICE, "NCCL barrier failed: %s", c.tr.nccl.GetErrorString(rc));
}

// exchange CUDA IPC handles of the two receive buffers and map every peer's (NVSwitch: any-to-any peer access)
static void setup_p2p(Ctx& c) {
    Transport& t = c.tr;
    if (!t.have_nccl() || getenv("PS3D_NO_P2P")) return;
    const int P = t.nranks;
    constexpr int NB = Transport::NPEERBUF;
    double* local[NB] = {c.W[6].p, c.W[8].p, c.velx[0].p, c.velx[1].p, c.velx[2].p, c.W[5].p, c.W[7].p};
    cudaIpcMemHandle_t mine[NB];
    int ok = 1;
    for (int b = 0; b < NB; ++b)
        if (cudaIpcGetMemHandle(&mine[b], local[b]) != cudaSuccess) ok = 0;
    (void)cudaGetLastError();
    const size_t hb = sizeof(mine);
    DevBuf<unsigned char> all;
    all.alloc(hb * P);
    std::vector<unsigned char> host(hb * P, 0);
    ps_h2d(all.p + hb * t.rank, mine, hb, c.stream);
    int rc = t.nccl.AllGather(all.p + hb * t.rank, all.p, hb, NcclApi::kUint8, t.comm, (void*)c.stream);
    if (rc != 0) fail(PS3D_ERR_DEVICE, "NCCL all-gather of IPC handles failed: %s", t.nccl.GetErrorString(rc));
    ps_d2h(host.data(), all.p, hb * P, c.stream);
    ps_sync(c.stream);
    for (int p = 0; p < P && ok; ++p) {
        for (int b = 0; b < NB; ++b) {
            if (p == t.rank) { t.peer_t2[b][p] = local[b]; continue; }
            cudaIpcMemHandle_t h;
            memcpy(&h, host.data() + hb * p + sizeof(h) * b, sizeof(h));
            void* ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; (void)cudaGetLastError(); break; }
            t.ipc_opened[b][p] = ptr;
            t.peer_t2[b][p] = (double*)ptr;
        }
    }
    // every rank must take the same path
    double flag = ok ? 0.0 : 1.0;
    allreduce_host(c, &flag, 1, 1u);
    t.p2p = (flag == 0.0);
    if (t.p2p)
        for (int p = 0; p < P; ++p) c.mail.m[p] = reinterpret_cast<PeerMail*>(t.peer_t2[0][p] + c.nint);
    t.peer_sync = getenv("PS3D_NO_PEER_SYNC") ? 0 : 1;
    t.rot_bufs = (getenv("PS3D_ROT_BUFS") && atoi(getenv("PS3D_ROT_BUFS")) == 2) ? 2 : 4;
    // (measured at 4 GPUs, 512^3: 21.5 ms/step split-phase vs 20.8 with one barrier kernel per 2-D FFT -- the three extra
    //  one-warp kernels per exchange cost more stream time than the decoupling saves; off by default)
    t.split_phase = (getenv("PS3D_SPLIT_PHASE") && atoi(getenv("PS3D_SPLIT_PHASE"))) ? 1 : 0;
    all.release();
}
#endif

// A batch of 2-D FFTs as a two-sweep pipeline with the slab exchange in between:
//   one rank : first sweep -> tmp -> second sweep
//   P ranks  : first sweep(i) -> t1 ([d][xl][kyl][pz] blocks) -> all-to-all(i) on the comm stream -> t2 -> second
//              sweep(i); the exchange of field i overlaps the first sweep of field i+1 and the second sweep of
//              field i-1 (two buffer pairs, CUDA events between the two streams).
// forward (fftxyp2s): first = y sweep, second = x sweep; inverse (fftxys2p): first = x, second = y.
static void fft2d_batch(Ctx& c, int n, Sweep* first, Sweep* second) {
    if (c.nranks == 1) {
        // (the L2-blocked variant assumes 16-z tiles on both axes)
        const int nzc = c.pz / LINE_ZC, G = (line_zc(c.nx) == LINE_ZC && line_zc(c.ny) == LINE_ZC && !c.gen[0] && !c.gen[1]) ? c.l2_chunks : 0;
        if (G <= 0 || 2 * G > nzc) {       // the two ping-pong intermediates of 16 G levels must fit the work field
            for (int i = 0; i < n; ++i) {
                if (!first[i].out) first[i].out = c.W[5].p;          // (a caller may keep the intermediate: do_vor2vel)
                run_sweep(c, first[i]);
                second[i].in[0] = first[i].out;
                run_sweep(c, second[i]);
            }
            return;
        }
        // L2-blocked: the two sweeps of a 2-D FFT run chunk by chunk (G z-chunks of 16 levels at a time) through
        // a compact intermediate [row][row][16 G] that is rewritten every chunk and therefore stays in the
        // 126 MB L2: one HBM read and one HBM write per 2-D FFT instead of two of each
        for (int i = 0; i < n; ++i) {
            for (int z0 = 0, k = 0; z0 < nzc; z0 += G, ++k) {
                double* t = (k & 1) ? c.W[5].p + (size_t)c.nx * c.ny * LINE_ZC * G : c.W[5].p;
                const int g = std::min(G, nzc - z0);
                Sweep a = first[i], b = second[i];
                a.out = t; a.zc0 = z0; a.nzc = g; a.out_pitch = LINE_ZC * G; a.out_zc0 = z0; a.final_store = 0;
                b.in[0] = t; b.zc0 = z0; b.nzc = g; b.in_pitch = LINE_ZC * G; b.in_zc0 = z0;
                run_sweep(c, a);
                run_sweep(c, b);
            }
        }
        return;
    }
    double* t1[2] = {c.W[5].p, c.W[7].p};
    double* t2[2] = {c.W[6].p, c.W[8].p};
#ifndef PS3D_EMU
    if (c.tr.p2p) {
        // fused sweep + exchange: the first sweep of field i stores every row straight into the owner's
        // receive buffer over NVLink (peer memory), a cross-rank barrier, then the second sweep.  Two
        // receive buffers; the barrier of field i+1 orders the reuse of buffer i&1 (peers enter it only
        // after their second sweep of field i), the barrier at batch start orders reuse across batches.
        // two streams: B (comm_stream) runs the NVLink-bound scatter sweeps as persistent launches with one
        // block per SM and the barriers, A (stream) the HBM-bound second sweeps in the SM slots left free, so
        // the scatter of field i+1 overlaps the second sweep of field i.
        //   B: s1(i) ; wait[A: s2(i-1) done] ; barrier(i)         A: wait[B: barrier(i)] ; s2(i)
        // s1(i+1) writes the peers' buffer (i+1)&1, last read by their s2(i-1): they enter barrier(i) only after it.
        PS_CUDA_TRY(cudaEventRecord(c.ev_first[0], c.stream));
        PS_CUDA_TRY(cudaStreamWaitEvent(c.comm_stream, c.ev_first[0], 0));
        if (c.tr.peer_sync && c.tr.split_phase) {
            // Split-phase protocol (no stream ever blocks on a full barrier): per receive buffer b every rank counts
            // its exchanges k = 1, 2, ...;
            //   B: wait[done(b) >= k-1 from every rank] ; s1(i) -> peers' buffer b ; signal arr(b) = k to every rank
            //   A: wait[arr(b) >= k from every rank] ; s2(i) ; signal done(b) = k to every rank
            // The NVLink-bound first sweeps run back to back on B, skew between the ranks is absorbed by A, which has
            // slack (its sweeps are 2-4x shorter).  The flags a rank sends itself order its own two streams.
            // The kept velocity buffers (>= 2) are read until the next vor2vel: their done flag is sent there.
            for (int i = 0, j = 0; i < n; ++i) {
                if (!Transport::is_kept(first[i].scatter)) first[i].scatter = Transport::rot_index(j++ % c.tr.rot_bufs);
                const int b = first[i].scatter;
                const unsigned long long k = ++c.tr.use_count[b];
                first[i].out = c.tr.peer_t2[b][c.rank];
                first[i].on_comm_stream = true;
                first[i].max_ctas = (c.p2p_ctas_per_sm >= 0 ? c.p2p_ctas_per_sm : 2) * c.num_sms;
                if (k > 1) {
                    PS_LAUNCH((k_peer_wait), dim3(1), dim3(32), 0, c.comm_stream, c.mail, c.rank, c.nranks, 1, b, k - 1);
                    ++c.launches;
                }
                run_sweep(c, first[i]);
                ++c.tr.n_alltoall;
                c.tr.bytes_sent += (double)c.nxl * c.nyl * c.pz * 8.0 * (c.nranks - 1);
                PS_LAUNCH((k_peer_signal), dim3(1), dim3(32), 0, c.comm_stream, c.mail, c.rank, c.nranks, 0, b, k);
                PS_LAUNCH((k_peer_wait), dim3(1), dim3(32), 0, c.stream, c.mail, c.rank, c.nranks, 0, b, k);
                second[i].in[0] = first[i].out;
                run_sweep(c, second[i]);
                if (!Transport::is_kept(b)) {
                    PS_LAUNCH((k_peer_signal), dim3(1), dim3(32), 0, c.stream, c.mail, c.rank, c.nranks, 1, b, k);
                    ++c.launches;
                }
                c.launches += 2;
            }
            return;
        }
        cross_rank_barrier(c, c.comm_stream);
        const int NR = c.tr.rot_bufs;
        for (int i = 0, j = 0; i < n; ++i) {
            // (a caller may name a dedicated receive buffer that it keeps: do_vor2vel); the others rotate through NR
            // buffers: s1(i+1) overwrites the buffer last read by a second sweep no later than s2(i+1-NR), and every
            // rank enters barrier(i) only after its s2(i+1-NR)
            if (!Transport::is_kept(first[i].scatter)) first[i].scatter = Transport::rot_index(j++ % NR);
            first[i].out = c.tr.peer_t2[first[i].scatter][c.rank];
            first[i].on_comm_stream = true;
            // persistent launch, two blocks per SM: the system-scope fence that ends a scatter sweep is then paid once
            // per block instead of once per tile (measured, 512^3 cn2 on 4 GPUs: full grid 27.4 ms/step, persistent
            // 22.6, full grid without the fence 22.2)
            first[i].max_ctas = (c.p2p_ctas_per_sm >= 0 ? c.p2p_ctas_per_sm : 2) * c.num_sms;
            run_sweep(c, first[i]);
            ++c.tr.n_alltoall;
            c.tr.bytes_sent += (double)c.nxl * c.nyl * c.pz * 8.0 * (c.nranks - 1);
            if (i + 1 - NR >= 0) PS_CUDA_TRY(cudaStreamWaitEvent(c.comm_stream, c.ev_second[i + 1 - NR], 0));
            cross_rank_barrier(c, c.comm_stream);
            PS_CUDA_TRY(cudaEventRecord(c.ev_a2a[i], c.comm_stream));
            PS_CUDA_TRY(cudaStreamWaitEvent(c.stream, c.ev_a2a[i], 0));
            second[i].in[0] = first[i].out;
            run_sweep(c, second[i]);
            PS_CUDA_TRY(cudaEventRecord(c.ev_second[i], c.stream));
        }
        return;
    }
#endif
#ifndef PS3D_EMU
    const bool overlap = c.tr.have_nccl() && !c.tr.a2a_cb;
#else
    const bool overlap = false;
#endif
    if (!overlap) {
        for (int i = 0; i < n; ++i) {
            first[i].out = t1[0];
            run_sweep(c, first[i]);
            exchange(c, t1[0], t2[0], c.stream);
            second[i].in[0] = t2[0];
            run_sweep(c, second[i]);
        }
        return;
    }
#ifndef PS3D_EMU
    if (n > 8) fail(PS3D_ERR_BAD_ARGUMENT, "fft2d_batch: at most 8 fields");
    for (int i = 0; i <= n; ++i) {
        if (i < n) {
            first[i].out = t1[i & 1];
            run_sweep(c, first[i]);
            PS_CUDA_TRY(cudaEventRecord(c.ev_first[i], c.stream));
            PS_CUDA_TRY(cudaStreamWaitEvent(c.comm_stream, c.ev_first[i], 0));
            if (i >= 2) PS_CUDA_TRY(cudaStreamWaitEvent(c.comm_stream, c.ev_second[i - 2], 0));   // t2[i&1] free again
            exchange(c, t1[i & 1], t2[i & 1], c.comm_stream);
            PS_CUDA_TRY(cudaEventRecord(c.ev_a2a[i], c.comm_stream));
        }
        if (i >= 1) {
            const int j = i - 1;
            PS_CUDA_TRY(cudaStreamWaitEvent(c.stream, c.ev_a2a[j], 0));
            second[j].in[0] = t2[j & 1];
            run_sweep(c, second[j]);
            PS_CUDA_TRY(cudaEventRecord(c.ev_second[j], c.stream));
        }
    }
#endif
}

static Sweep sweep_plain(int axis, bool inv, bool diff, const double* in, double* out) {
    return Sweep{axis, inv, diff ? PRO_DIFF : PRO_PLAIN, {in, nullptr, nullptr, nullptr}, 0.0, 0.0, out};
}

// fftxyp2s on internal layouts: physical [xl][y][pz] -> semi-spectral [kx][kyl][pz]
static void fft2d_fwd(Ctx& c, const double* in, double* out) {
    Sweep a = sweep_plain(1, false, false, in, nullptr), b = sweep_plain(0, false, false, nullptr, out);
    fft2d_batch(c, 1, &a, &b);
}

// fftxys2p (optionally of d/dx or d/dy of the input): [kx][kyl][pz] -> [xl][y][pz]
static void fft2d_inv(Ctx& c, const double* in, double* out, bool dx, bool dy) {
    Sweep a = sweep_plain(0, true, dx, in, nullptr), b = sweep_plain(1, true, dy, nullptr, out);
    fft2d_batch(c, 1, &a, &b);
}

template <int NZ>
static void launch_zop_n(Ctx& c, int op, const double* in, double* out) {
    const size_t sm = zop_smem_bytes<NZ>();
    allow_smem(k_zop<NZ>, sm);
    PS_LAUNCH((k_zop<NZ>), dim3(c.ngroups()), dim3(ZCfg<NZ>::NT), sm, c.stream, c.geom(), op, in, out);
    ++c.launches;
}
static void launch_zop(Ctx& c, int op, const double* in, double* out) {
    if (c.gen[2]) {
        const size_t sm = gen_col_smem_bytes(1, c.nz);
        allow_smem(k_zop_gen, sm);
        PS_LAUNCH((k_zop_gen), dim3(c.ngroups()), dim3(GEN_COL_THREADS), sm, c.stream, c.geom(), c.genz(), op, in, out);
        ++c.launches;
        return;
    }
    switch (c.nz) {
#define X(NN) case NN: launch_zop_n<NN>(c, op, in, out); break;
        PS_FOR_Z_SIZES(X)
#undef X
        default: fail(PS3D_ERR_UNSUPPORTED_SIZE, "nz = %d not supported (power of two, 8..1024)", c.nz);
    }
}

template <int NZ>
static void launch_v2v_n(Ctx& c, const V2VArgs& a) {
    const SpecGeom gf = c.geom_fast();
    if (gf.npf > 0) {
        const size_t sm = v2v_smem_bytes<NZ>(false);
        allow_smem(k_vor2vel_spec<NZ, false>, sm);
        PS_LAUNCH((k_vor2vel_spec<NZ, false>), dim3((c.nx / 2 + 1) * gf.npf), dim3(ZCfg<NZ>::NT), sm, c.stream, gf, a);
        ++c.launches;
    }
    if (c.rank == 0) {
        const size_t sm = v2v_smem_bytes<NZ>(true);
        allow_smem(k_vor2vel_spec<NZ, true>, sm);
        PS_LAUNCH((k_vor2vel_spec<NZ, true>), dim3(c.nx / 2 + 1), dim3(ZCfg<NZ>::NT), sm, c.stream, c.geom_gen(), a);
        ++c.launches;
    }
}
static void launch_v2v(Ctx& c, const V2VArgs& a) {
    if (c.gen[2]) {
        const size_t sm = gen_col_smem_bytes(4, c.nz);
        allow_smem(k_vor2vel_gen, sm);
        PS_LAUNCH((k_vor2vel_gen), dim3(c.ngroups()), dim3(GEN_COL_THREADS), sm, c.stream, c.geom(), c.genz(), a);
        ++c.launches;
        return;
    }
    switch (c.nz) {
#define X(NN) case NN: launch_v2v_n<NN>(c, a); break;
        PS_FOR_Z_SIZES(X)
#undef X
        default: fail(PS3D_ERR_UNSUPPORTED_SIZE, "nz = %d not supported (power of two, 8..1024)", c.nz);
    }
}

template <int NZ, bool RK>
static void launch_src_nr(Ctx& c, const SrcArgs& a) {
    const SpecGeom gf = c.geom_fast();
    if (gf.npf > 0) {
        const size_t sm = src_smem_bytes<NZ>(false);
        allow_smem(k_source_spec<NZ, false, RK>, sm);
        PS_LAUNCH((k_source_spec<NZ, false, RK>), dim3((c.nx / 2 + 1) * gf.npf), dim3(ZCfg<NZ>::NT), sm, c.stream, gf, a);
        ++c.launches;
    }
    if (c.rank == 0) {
        const size_t sm = src_smem_bytes<NZ>(true);
        allow_smem(k_source_spec<NZ, true, RK>, sm);
        PS_LAUNCH((k_source_spec<NZ, true, RK>), dim3(c.nx / 2 + 1), dim3(ZCfg<NZ>::NT), sm, c.stream, c.geom_gen(), a);
        ++c.launches;
    }
}
template <int NZ>
static void launch_src_n(Ctx& c, const SrcArgs& a) {
    if (a.upd >= 11) launch_src_nr<NZ, true>(c, a);
    else launch_src_nr<NZ, false>(c, a);
}
static void launch_src(Ctx& c, const SrcArgs& a) {
    if (c.gen[2]) {
        const size_t sm = gen_col_smem_bytes(4, c.nz);
        allow_smem(k_source_gen, sm);
        PS_LAUNCH((k_source_gen), dim3(c.ngroups()), dim3(GEN_COL_THREADS), sm, c.stream, c.geom(), c.genz(), a);
        ++c.launches;
        return;
    }
    switch (c.nz) {
#define X(NN) case NN: launch_src_n<NN>(c, a); break;
        PS_FOR_Z_SIZES(X)
#undef X
        default: fail(PS3D_ERR_UNSUPPORTED_SIZE, "nz = %d not supported (power of two, 8..1024)", c.nz);
    }
}

static int stream_blocks(size_t n) { return (int)std::min<size_t>((n + 255) / 256, (size_t)148 * 16); }

// ---------------------------------------------------------------------------
// host boundary: natural Fortran layout <-> internal layouts
// ---------------------------------------------------------------------------
// Host layouts: physical fields f(0:nz, 0:ny-1, x-slab of this rank); spectral fields f(0:nz, ky, 0:nx-1) with ky in
// natural order on one rank, and on P > 1 ranks the rank's slab of the paired order ky' (0, ny/2, 1, ny-1, ...).
static void to_device(Ctx& c, const double* host, double* dev, bool spectral) {
    const bool slab = spectral && c.nranks > 1;
    const int d0 = slab ? c.nx : c.nxl, d1 = slab ? c.nyl : c.ny;
    ps_h2d(c.stage.p, host, c.nnat * sizeof(double), c.stream);
    PS_LAUNCH((k_repack_in), dim3(stream_blocks(c.nint)), dim3(256), 0, c.stream, (const double*)c.stage.p, dev, d0,
              d1, c.nzp, c.pz, (spectral && !slab) ? (const int*)c.permy.p : (const int*)nullptr);
    ++c.launches;
}
static void to_host(Ctx& c, const double* dev, double* host, bool spectral) {
    const bool slab = spectral && c.nranks > 1;
    const int d0 = slab ? c.nx : c.nxl, d1 = slab ? c.nyl : c.ny;
    PS_LAUNCH((k_repack_out), dim3(stream_blocks(c.nnat)), dim3(256), 0, c.stream, dev, c.stage.p, d0, d1, c.nzp,
              c.pz, (spectral && !slab) ? (const int*)c.permy.p : (const int*)nullptr);
    ++c.launches;
    ps_d2h(host, c.stage.p, c.nnat * sizeof(double), c.stream);
    ps_sync(c.stream);
}

// ---------------------------------------------------------------------------
// init
// ---------------------------------------------------------------------------
static void do_init(int nx, int ny, int nz, const double* lower, const double* extent, int rank, int nranks,
                    const void* nccl_id) {
    if (g_ctx) fail(PS3D_ERR_BAD_ARGUMENT, "ps3d_cuda_init called twice without ps3d_cuda_finalise");
    if (!lower || !extent) fail(PS3D_ERR_BAD_ARGUMENT, "null lower/extent");
    // Powers of two in 8..1024 run through the register-blocked kernels; other even lengths with prime factors
    // 2, 3, 5 only (the lengths factorisen accepts, stafft.f90:128-187) through the mixed-radix coverage kernels
    // (shared-memory limits: nx, ny <= 896, nz <= 1200).
    bool gen_axis[3];
    GenPlan gplan[3];
    {
        const int nn[3] = {nx, ny, nz};
        const int lim[3] = {896, 896, 1200};
        for (int i = 0; i < 3; ++i) {
            const bool fast = pow2(nn[i]) && nn[i] >= 8 && nn[i] <= 1024;
            gen_axis[i] = !fast;
            if (fast) continue;
            if (nn[i] < 6 || (nn[i] & 1) || nn[i] > lim[i] || !gen_plan_make(nn[i], gplan[i]))
                fail(PS3D_ERR_UNSUPPORTED_SIZE,
                     "grid %dx%dx%d: supported are powers of two in 8..1024 and even lengths 2^a 3^b 5^c (nx, ny <= 896, "
                     "nz <= 1200)", nx, ny, nz);
        }
    }
    if (nranks < 1 || rank < 0 || rank >= nranks || nx % nranks || (ny / 2) % nranks)
        fail(PS3D_ERR_BAD_ARGUMENT, "bad rank layout %d/%d for %dx%d", rank, nranks, nx, ny);
    if (nranks > PS3D_MAX_RANKS)
        fail(PS3D_ERR_BAD_ARGUMENT, "nranks = %d: at most %d ranks (the GPUs of one NVSwitch box; peer tables are fixed-size)", nranks, PS3D_MAX_RANKS);
    for (int i = 0; i < 3; ++i)
        if (!(extent[i] > 0.0)) fail(PS3D_ERR_BAD_ARGUMENT, "domain extent must be positive (sta2dfft.f90:59-74)");
#ifndef PS3D_EMU
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        fail(PS3D_ERR_NO_DEVICE, "no CUDA device: libps3d_cuda has no CPU path");
    {
        // one process per GPU: PS3D_DEVICE, else the launcher's LOCAL_RANK, else rank modulo the device count
        const char* e = getenv("PS3D_DEVICE");
        if (!e) e = getenv("LOCAL_RANK");
        int dev = e ? atoi(e) : rank % ndev;
        if (dev < 0 || dev >= ndev) fail(PS3D_ERR_BAD_ARGUMENT, "device %d out of range (%d devices)", dev, ndev);
        PS_CUDA_TRY(cudaSetDevice(dev));
    }
#endif
    Ctx* c = new Ctx();
    g_ctx = c;
    c->nx = nx; c->ny = ny; c->nz = nz; c->nzp = nz + 1;
#ifndef PS3D_EMU
    g_trace = getenv("PS3D_TRACE") ? atoi(getenv("PS3D_TRACE")) : 0;
#endif
    c->strict_jacobi = getenv("PS3D_STRICT_JACOBI") ? atoi(getenv("PS3D_STRICT_JACOBI")) : 0;
    c->l2_chunks = getenv("PS3D_L2_CHUNKS") ? atoi(getenv("PS3D_L2_CHUNKS")) : 0;
    c->fuse_update = getenv("PS3D_NO_FUSED_UPDATE") ? 0 : 1;
    c->keep_velx = getenv("PS3D_NO_KEEP_VELX") ? 0 : 1;
    c->scatter_fence = getenv("PS3D_NO_SCATTER_FENCE") ? 0 : 1;
    c->p2p_ctas_per_sm = getenv("PS3D_P2P_CTAS") ? atoi(getenv("PS3D_P2P_CTAS")) : -1;
    c->pz = (c->nzp + LINE_ZC - 1) / LINE_ZC * LINE_ZC;
    c->rank = rank; c->nranks = nranks;
    c->nxl = nx / nranks; c->nyl = ny / nranks;
    for (int i = 0; i < 3; ++i) {
        c->lower[i] = lower[i]; c->extent[i] = extent[i];
        c->upper[i] = lower[i] + extent[i];
    }
    c->dx[0] = extent[0] / (double)nx; c->dx[1] = extent[1] / (double)ny; c->dx[2] = extent[2] / (double)nz;
    c->ncell = (long long)nx * ny * nz;
    c->nint = (size_t)c->nxl * ny * c->pz;
    c->nnat = (size_t)c->nxl * ny * c->nzp;
#ifndef PS3D_EMU
    {
        int lo = 0, hi = 0;
        PS_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        PS_CUDA_TRY(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, hi));
        cudaDeviceProp prop;
        int dev = 0;
        PS_CUDA_TRY(cudaGetDevice(&dev));
        PS_CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
        c->num_sms = prop.multiProcessorCount;
    }
    {
        // cuTensorMapEncodeTiled without linking libcuda: the runtime hands out the driver entry point.
        // Default: on for one rank (with several ranks the scatter sweeps and the second sweeps of neighbouring
        // fields share the SMs, which one 192 KB block per SM does not allow); PS3D_LINE_TMA=0/1 overrides.
        const char* e = getenv("PS3D_LINE_TMA");
        c->line_tma = e ? atoi(e) : (nranks == 1 ? 1 : 0);
        c->tma_zc8 = getenv("PS3D_TMA_ZC") && atoi(getenv("PS3D_TMA_ZC")) == 8;
        c->tma_rot = getenv("PS3D_TMA_ROT") ? atoi(getenv("PS3D_TMA_ROT")) : 1;
        cudaDriverEntryPointQueryResult qr;
        void* fn = nullptr;
        if (c->line_tma && (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess ||
                            qr != cudaDriverEntryPointSuccess || !fn)) {
            (void)cudaGetLastError();
            c->line_tma = 0;
        }
        c->encode_tiled = fn;
    }
    PS_CUDA_TRY(cudaEventCreate(&c->ev0));
    PS_CUDA_TRY(cudaEventCreate(&c->ev1));
    PS_CUDA_TRY(cudaMallocHost((void**)&c->h_red, 64 * sizeof(double)));
#else
    c->h_red = (double*)calloc(64, sizeof(double));
#endif
    c->tr.rank = rank; c->tr.nranks = nranks;
#ifndef PS3D_EMU
    if (nranks > 1) {
        {
            int lo = 0, hi = 0;
            PS_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            PS_CUDA_TRY(cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, lo));
        }
        for (int i = 0; i < 8; ++i) {
            PS_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_first[i], cudaEventDisableTiming));
            PS_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_a2a[i], cudaEventDisableTiming));
            PS_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_second[i], cudaEventDisableTiming));
        }
    }
    if (nranks > 1 && nccl_id) {
        if (!c->tr.nccl.load()) throw StatusError{PS3D_ERR_DEVICE};
        NcclApi::UniqueId id;
        memcpy(id.internal, nccl_id, sizeof id.internal);
        const int rc = c->tr.nccl.CommInitRank(&c->tr.comm, nranks, id, rank);
        if (rc != 0) fail(PS3D_ERR_DEVICE, "ncclCommInitRank failed: %s", c->tr.nccl.GetErrorString(rc));
    }
#else
    (void)nccl_id;
#endif
    ps_stream_t s = c->stream;

    for (int i = 0; i < 3; ++i) {
        c->gen[i] = gen_axis[i];
        if (!gen_axis[i]) continue;
        c->plan[i] = gplan[i];
        const int n = gplan[i].n;
        const long double pi_l = 3.14159265358979323846264338327950288L;
        std::vector<double2> tw(n);
        for (int m = 0; m < n; ++m) {
            const long double ang = 2.0L * pi_l * (long double)m / (long double)n;
            tw[m] = make_double2((double)cosl(ang), (double)sinl(ang));
        }
        c->gtw[i].upload(tw, c->stream);
        if (i == 2) {
            std::vector<double> sn(n + 1), cs(n + 1);
            for (int j = 0; j <= n; ++j) {
                const long double ang = pi_l * (long double)j / (long double)n;
                sn[j] = (double)sinl(ang); cs[j] = (double)cosl(ang);
            }
            sn[0] = 0.0; sn[n] = 0.0; cs[n / 2] = 0.0;
            c->gsinz.upload(sn, c->stream); c->gcosz.upload(cs, c->stream);
        }
    }
    // twiddles of the power-of-two kernels: exp(2 pi i m / ntw), ntw = max(nx, ny, 2 nz) over the power-of-two axes
    c->ntw = std::max(std::max(gen_axis[0] ? 16 : nx, gen_axis[1] ? 16 : ny), gen_axis[2] ? 16 : 2 * nz);
    {
        std::vector<double2> tw(c->ntw);
        for (int m = 0; m < c->ntw; ++m) {
            const long double ang = 2.0L * 3.14159265358979323846264338327950288L * (long double)m / (long double)c->ntw;
            tw[m] = make_double2((double)cosl(ang), (double)sinl(ang));
        }
        // exact values on the axes / diagonals
        const int q = c->ntw / 4;
        tw[0] = make_double2(1.0, 0.0); tw[q] = make_double2(0.0, 1.0);
        tw[2 * q] = make_double2(-1.0, 0.0); tw[3 * q] = make_double2(0.0, -1.0);
        c->tw.upload(tw, s);
    }
    // wavenumbers: sta2dfft.f90:59-66, sta3dfft.f90:89-108, deriv1d.f90:17-22
    const double pi = std::acos(-1.0);
    c->h_rkx.assign(nx, 0.0); c->h_rky.assign(ny, 0.0); c->h_rkz.assign(nz + 1, 0.0);
    {
        const double scx = pi / extent[0], scy = pi / extent[1], scz = pi / extent[2];
        for (int k = 1; k < nx / 2; ++k) { c->h_rkx[k] = scx * (double)(2 * k); c->h_rkx[nx - k] = c->h_rkx[k]; }
        c->h_rkx[nx / 2] = scx * (double)nx;
        for (int k = 1; k < ny / 2; ++k) { c->h_rky[k] = scy * (double)(2 * k); c->h_rky[ny - k] = c->h_rky[k]; }
        c->h_rky[ny / 2] = scy * (double)ny;
        for (int k = 1; k <= nz; ++k) c->h_rkz[k] = scz * (double)k;
        std::vector<double> kxl(nx / 2 + 1, 0.0), kyl(ny / 2 + 1, 0.0);
        for (int k = 1; k < nx / 2; ++k) kxl[k] = c->h_rkx[k];     // 0 at k = 0 and nx/2 (sta3dfft.f90:321-335)
        for (int k = 1; k < ny / 2; ++k) kyl[k] = c->h_rky[k];
        c->kxl.upload(kxl, s); c->kyline.upload(kyl, s); c->kxd.upload(kxl, s);
        c->rkz.upload(c->h_rkz, s);
    }
    // paired ky order: ky' = 2a + sy;  a = 0: (0, ny/2);  a >= 1: (a, ny - a)
    {
        std::vector<int> perm(ny);
        perm[0] = 0; perm[ny / 2] = 1;
        for (int k = 1; k < ny / 2; ++k) { perm[k] = 2 * k; perm[ny - k] = 2 * k + 1; }
        c->permy.upload(perm, s);
        c->h_ky_of_kyl.assign(c->nyl, 0);
        std::vector<double> kyd(c->nyl / 2, 0.0);
        for (int k = 0; k < ny; ++k) {
            const int kp = perm[k];
            if (kp / c->nyl == rank) c->h_ky_of_kyl[kp % c->nyl] = k;
        }
        for (int ap = 0; ap < c->nyl / 2; ++ap) {
            const int a = (rank * c->nyl) / 2 + ap;
            kyd[ap] = (a == 0) ? 0.0 : c->h_rky[a];
        }
        c->kyd.upload(kyd, s);
    }
    c->stage.alloc(c->nnat);
    // (the first receive buffer carries the peer-memory mailbox behind its field: one IPC mapping serves both)
    for (int i = 0; i < (nranks > 1 ? 9 : 6); ++i) c->W[i].alloc(c->nint + (i == 6 ? sizeof(PeerMail) / sizeof(double) + 1 : 0));
    c->redS.alloc(64); c->redM.alloc(64);
    c->red_blocks = (int)std::min<size_t>((size_t)RED_BLOCKS, (c->nint + RED_THREADS - 1) / RED_THREADS);
    c->partial.alloc((size_t)RED_BLOCKS * 16);
    c->red.alloc(64);
#ifndef PS3D_EMU
    if (nranks > 1) {
        // the kept x-transformed velocity (do_vor2vel) is a peer-written receive buffer on several ranks: it must exist
        // before the IPC handles are exchanged
        if (c->keep_velx && nccl_id && !getenv("PS3D_NO_P2P")) for (int i = 0; i < 3; ++i) c->velx[i].alloc(c->nint);
        setup_p2p(*c);
    }
#endif
}

static void do_init_inversion(int filtering) {
    Ctx& c = ctx();
    if (c.inversion_ready) return;                       // inversion_utils.f90:227-229
    if (filtering != PS3D_FILTER_HOU_LI && filtering != PS3D_FILTER_23_RULE)
        fail(PS3D_ERR_BAD_ARGUMENT, "unknown filtering id %d", filtering);
    c.filtering = filtering;
    const int nx = c.nx, ny = c.ny, nz = c.nz, nyl = c.nyl;
    ps_stream_t s = c.stream;
    // k2l2, k2l2i (inversion_utils.f90:240-258), one row per slot pair b = min(kx, nx-kx)
    std::vector<double> k2((size_t)(nx / 2 + 1) * nyl), k2i(k2.size());
    for (int b = 0; b <= nx / 2; ++b)
        for (int kl = 0; kl < nyl; ++kl) {
            const int ky = c.h_ky_of_kyl[kl];
            double v = c.h_rkx[b] * c.h_rkx[b] + c.h_rky[ky] * c.h_rky[ky];
            double vi;
            if (b == 0 && ky == 0) { v = 0.0; vi = 0.0; } else vi = 1.0 / v;
            k2[(size_t)b * nyl + kl] = v; k2i[(size_t)b * nyl + kl] = vi;
        }
    c.k2l2.upload(k2, s); c.k2l2i.upload(k2i, s);
    // full [kx][kyl] copies for the steppers
    c.h_k2l2.assign((size_t)nx * nyl, 0.0);
    for (int kx = 0; kx < nx; ++kx)
        for (int kl = 0; kl < nyl; ++kl)
            c.h_k2l2[(size_t)kx * nyl + kl] = k2[(size_t)std::min(kx, nx - kx) * nyl + kl];
    // de-aliasing filter, separable parts (inversion_utils.f90:377-455)
    double rkxmax = 0, rkymax = 0, rkzmax = 0;
    for (double v : c.h_rkx) rkxmax = std::max(rkxmax, v);
    for (double v : c.h_rky) rkymax = std::max(rkymax, v);
    for (double v : c.h_rkz) rkzmax = std::max(rkzmax, v);
    std::vector<double> filtz(nz + 1, 1.0);
    c.h_filt2d.assign((size_t)nx * nyl, 0.0);
    if (filtering == PS3D_FILTER_HOU_LI) {
        const double kxmaxi = 1.0 / rkxmax, kymaxi = 1.0 / rkymax, kzmaxi = 1.0 / rkzmax;
        for (int kx = 0; kx < nx; ++kx)
            for (int kl = 0; kl < nyl; ++kl) {
                const double skx = -36.0 * std::pow(kxmaxi * c.h_rkx[kx], 36);
                const double sky = -36.0 * std::pow(kymaxi * c.h_rky[c.h_ky_of_kyl[kl]], 36);
                c.h_filt2d[(size_t)kx * nyl + kl] = std::exp(skx + sky);
            }
        for (int kz = 1; kz < nz; ++kz) filtz[kz] = std::exp(-36.0 * std::pow(kzmaxi * c.h_rkz[kz], 36));
    } else {
        const double f23 = 2.0 / 3.0;
        for (int kx = 0; kx < nx; ++kx)
            for (int kl = 0; kl < nyl; ++kl)
                c.h_filt2d[(size_t)kx * nyl + kl] =
                    ((c.h_rkx[kx] <= f23 * rkxmax) ? 1.0 : 0.0) * ((c.h_rky[c.h_ky_of_kyl[kl]] <= f23 * rkymax) ? 1.0 : 0.0);
        for (int kz = 1; kz < nz; ++kz) filtz[kz] = (c.h_rkz[kz] <= f23 * rkzmax) ? 1.0 : 0.0;
    }
    if (c.rank == 0) c.h_filt2d[0] = 1.0;                // filt(:,0,0) = 1 (inversion_utils.f90:275-277)
    c.filt2d.upload(c.h_filt2d, s); c.filtz.upload(filtz, s);
    // zm, zp (inversion_utils.f90:293-297); gamtop, gambot (:350-369)
    std::vector<double> zm(nz + 1), zp(nz + 1), gt(nz + 1), gb(nz + 1), wz(nz + 1, 0.0);
    for (int iz = 0; iz <= nz; ++iz) {
        const double z = c.lower[2] + c.dx[2] * (double)iz;
        zm[iz] = c.upper[2] - z;
        zp[iz] = z - c.lower[2];
    }
    for (int iz = 0; iz <= nz; ++iz) {
        const double ph = zp[iz] / c.extent[2];
        gt[iz] = 0.5 * c.extent[2] * (ph * ph - 1.0 / 3.0);
    }
    for (int iz = 0; iz <= nz; ++iz) gb[iz] = gt[nz - iz];
    // weights of sum_k dst(x)(k) (field_diagnostics.f90:592-596)
    for (int j = 1; j < nz; j += 2) {
        const long double a = 3.14159265358979323846264338327950288L * (long double)j / (2.0L * (long double)nz);
        wz[j] = (double)(std::sqrt(2.0L / (long double)nz) * (cosl(a) / sinl(a)));
    }
    c.zm.upload(zm, s); c.zp.upload(zp, s); c.gamtop.upload(gt, s); c.gambot.upload(gb, s); c.wz.upload(wz, s);
    for (int i = 0; i < 3; ++i) {
        c.svor[i].alloc(c.nint); c.vor[i].alloc(c.nint); c.vel[i].alloc(c.nint);
        c.svel[i].alloc(c.nint); c.svorts[i].alloc(c.nint);
    }
    c.ini_mean.alloc(2);
    c.vhdis.alloc((size_t)nx * nyl); c.fac1.alloc((size_t)nx * nyl); c.fac2.alloc((size_t)nx * nyl);
    c.inversion_ready = true;
}

static Ctx& ready() {
    Ctx& c = ctx();
    if (!c.inversion_ready) fail(PS3D_ERR_NOT_INITIALISED, "Error: Inversion not initialised!");
    return c;
}

// get_viscosity + init_dissipation (inversion_utils.f90:157-218): returns the (hyper)viscosity, fills hdis[kx][kyl]
static double viscosity_table(Ctx& c, int nnu, double prediss, int lscale, double te, double en, std::vector<double>& vh) {
    double rkxmax = 0, rkymax = 0;
    for (double v : c.h_rkx) rkxmax = std::max(rkxmax, v);
    for (double v : c.h_rky) rkymax = std::max(rkymax, v);
    const double K2max = std::pow(std::max(rkxmax, rkymax), 2);
    const double rkmsi = 1.0 / K2max;
    double vis;
    if (lscale == PS3D_LSCALE_KOLMOGOROV) vis = prediss * std::pow(K2max * te / en, 1.0 / 3.0) * std::pow(rkmsi, nnu);
    else if (lscale == PS3D_LSCALE_GEOPHYSICAL) vis = prediss * std::pow(rkmsi, nnu);
    else fail(PS3D_ERR_BAD_ARGUMENT, "We only support 'Kolmogorov' or 'geophysical'");
    vh.resize(c.h_k2l2.size());
    for (size_t i = 0; i < vh.size(); ++i) vh[i] = (nnu == 1) ? vis * c.h_k2l2[i] : vis * std::pow(c.h_k2l2[i], nnu);
    return vis;
}

static double do_init_diffusion(int nnu, double prediss, int lscale, double te, double en) {
    Ctx& c = ready();
    std::vector<double> vh;
    const double vis = viscosity_table(c, nnu, prediss, lscale, te, en, vh);
    c.vvisc = vis; c.nnu = nnu;
    c.vhdis.upload(vh, c.stream);
    c.diffusion_ready = true;
    return vis;
}

static void do_finalise() {
    if (!g_ctx) return;
    Ctx* c = g_ctx;
    ps_sync(c->stream);
#ifndef PS3D_EMU
    trace_dump();
#endif
    DevBuf<double>* groups[] = {c->svor, c->vor, c->vel, c->svel, c->svorts, c->wa, c->wb, c->velx};
    for (auto* g : groups) for (int i = 0; i < 3; ++i) g[i].release();
    for (int i = 0; i < 9; ++i) c->W[i].release();
    c->redS.release(); c->redM.release();
#ifndef PS3D_EMU
    for (int b = 0; b < Transport::NPEERBUF; ++b) for (int p = 0; p < 8; ++p) if (c->tr.ipc_opened[b][p]) cudaIpcCloseMemHandle(c->tr.ipc_opened[b][p]);
    if (c->tr.comm) c->tr.nccl.CommDestroy(c->tr.comm);
#endif
#ifndef PS3D_EMU
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    for (int i = 0; i < 2; ++i) if (c->ev_copy[i]) cudaEventDestroy(c->ev_copy[i]);
#endif
    c->stage3[0].release(); c->stage3[1].release();
    DevBuf<double>* singles[] = {&c->stage, &c->kxl, &c->kyline, &c->kxd, &c->kyd, &c->k2l2, &c->k2l2i, &c->zm, &c->zp,
                                 &c->rkz, &c->gamtop, &c->gambot, &c->filt2d, &c->filtz, &c->vhdis, &c->fac1, &c->fac2,
                                 &c->wz, &c->ini_mean, &c->partial, &c->red};
    for (auto* b : singles) b->release();
    DevBuf<double>* bb[] = {&c->sbuoy, &c->buoy, &c->sbuoys, &c->bsm, &c->sbuoyf, &c->bsem, &c->bhdis, &c->bfac1, &c->bfac2};
    for (auto* b : bb) b->release();
    c->tw.release(); c->permy.release();
    for (int i = 0; i < 3; ++i) c->gtw[i].release();
    c->gsinz.release(); c->gcosz.release();
#ifndef PS3D_EMU
    if (c->h_red) cudaFreeHost(c->h_red);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    for (int i = 0; i < 8; ++i) {
        if (c->ev_first[i]) cudaEventDestroy(c->ev_first[i]);
        if (c->ev_a2a[i]) cudaEventDestroy(c->ev_a2a[i]);
        if (c->ev_second[i]) cudaEventDestroy(c->ev_second[i]);
    }
    if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
    if (c->stream) cudaStreamDestroy(c->stream);
#else
    free(c->h_red);
#endif
    delete c;
    g_ctx = nullptr;
}

// ---------------------------------------------------------------------------
// resident-mode operators
// ---------------------------------------------------------------------------
static void do_vor2vel(Ctx& c) {
    V2VArgs a;
    a.svor0 = c.svor[0].p; a.svor1 = c.svor[1].p; a.svor2 = c.svor[2].p;
    a.wsem0 = c.W[0].p; a.wsem1 = c.W[1].p; a.wsem2 = c.W[2].p;
    a.svel0 = c.svel[0].p; a.svel1 = c.svel[1].p; a.svel2 = c.svel[2].p;
    launch_v2v(c, a);
    // one rank: the intermediate of the two sweeps is kept; several ranks with peer memory: the velocity fields are
    // scattered into dedicated receive buffers (peer buffers 2..4 = velx) that nobody overwrites until the next vor2vel
    const bool keep1 = (c.nranks == 1 && c.l2_chunks <= 0 && c.keep_velx);
    const bool keepP = (c.nranks > 1 && c.tr.p2p && c.keep_velx && c.velx[0].p);
    const bool keep = keep1 || keepP;
#ifndef PS3D_EMU
    if (keepP && c.tr.peer_sync && c.tr.split_phase) {
        // every earlier reader of the kept velocity buffers (the second sweeps of the last vor2vel, adapt's d/dy sweeps)
        // is behind this point of the compute stream: the peers may overwrite them
        for (int b = 2; b <= 4; ++b)
            if (c.tr.use_count[b] > 0) {
                PS_LAUNCH((k_peer_signal), dim3(1), dim3(32), 0, c.stream, c.mail, c.rank, c.nranks, 1, b, c.tr.use_count[b]);
                ++c.launches;
            }
    }
#endif
    Sweep f[6], g[6];
    for (int i = 0; i < 3; ++i) {
        f[i] = sweep_plain(0, true, false, c.W[i].p, nullptr);        g[i] = sweep_plain(1, true, false, nullptr, c.vor[i].p);
        if (keep1 && !c.velx[i].p) c.velx[i].alloc(c.nint);
        f[3 + i] = sweep_plain(0, true, false, c.svel[i].p, keep1 ? c.velx[i].p : nullptr);
        if (keepP) f[3 + i].scatter = 2 + i;
        g[3 + i] = sweep_plain(1, true, false, nullptr, c.vel[i].p);
    }
    fft2d_batch(c, 6, f, g);
    c.velx_valid = keep;
}

static StepArgs step_args(Ctx& c);
static void vor_mean(Ctx& c, int mode);
static void launch_zop(Ctx& c, int op, const double* in, double* out);

// buoyancy_tendency (inversion.f90:232-292).  The reference takes each flux divergence back to physical space, sums
// there and decomposes the sum; combine and decompose are linear and mutually inverse, so the sum is formed in
// mixed-spectral space:  sbuoys = -[diffx(F1) + diffy(F2) + diffz(F3)] - bfsq * decompose(w),  F_i = decompose(u_i b').
static void do_buoyancy_tendency(Ctx& c) {
    const long long n = (long long)c.nint;
    const int nb = stream_blocks(c.nint);
    launch_zop(c, ZOP_COMBINE, c.sbuoy.p, c.bsem.p);                       // field_combine_physical(sbuoy, buoy) (:247)
    fft2d_inv(c, c.bsem.p, c.buoy.p, false, false);
    const int dop[3] = {ZOP_DIFFX, ZOP_DIFFY, ZOP_DIFFZ_SPEC};
    for (int i = 0; i < 3; ++i) {
        PS_LAUNCH((k_mul), dim3(nb), dim3(256), 0, c.stream, (const double*)c.vel[i].p, (const double*)c.buoy.p, c.W[3].p, n);
        ++c.launches;
        fft2d_fwd(c, c.W[3].p, c.W[4].p);
        launch_zop(c, ZOP_DECOMPOSE, c.W[4].p, c.W[3].p);
        launch_zop(c, dop[i], c.W[3].p, c.W[i].p);
    }
    launch_zop(c, ZOP_DECOMPOSE, c.svel[2].p, c.W[3].p);                   // mixed-spectral w (vel(:,:,:,3) of :288)
    PS_LAUNCH((k_btend), dim3(nb), dim3(256), 0, c.stream, (const double*)c.W[0].p, (const double*)c.W[1].p,
              (const double*)c.W[2].p, (const double*)c.W[3].p, c.bfsq, c.sbuoys.p, n);
    ++c.launches;
}

static StepArgs buoy_step_args(Ctx& c);

// upd < 0: svorts only.  upd = 0 / 1 (cn2, power-of-two nz): the Crank-Nicolson update of cn2.f90:120-135 /
// :162-173 rides on the last stage of the source kernel, svorts is not stored (see SrcArgs)
// upd = 11..14 (impl-diff-rk4, power-of-two nz): substep one..four rides on it the same way (c1, c2 = the stage
// coefficients, pq = epq or filt(0,:,:)); the buoyancy updates stay separate kernels (rk4_update_buoy)
static void do_source(Ctx& c, int upd = -1, double dt2 = 0.0, double c2 = 0.0, const double* pq = nullptr) {
    const double* fc = c.f_cor;              // vor + f_cor (inversion.f90:310-314; `vor` itself stays relative here)
    if (c.buoyancy) do_buoyancy_tendency(c); // inversion.f90:381-383; also leaves the semi-spectral b' in bsem
    const double *u = c.vel[0].p, *v = c.vel[1].p, *w = c.vel[2].p;
    const double *xi = c.vor[0].p, *eta = c.vor[1].p, *zeta = c.vor[2].p;
    // r = u*eta - v*xi ; q = w*xi - u*zeta ; p = v*zeta - w*eta   (inversion.f90:327,336,350)
    Sweep f[3] = {Sweep{1, false, PRO_CROSS, {u, eta, v, xi}, fc[1], fc[0], nullptr},
                  Sweep{1, false, PRO_CROSS, {w, xi, u, zeta}, fc[0], fc[2], nullptr},
                  Sweep{1, false, PRO_CROSS, {v, zeta, w, eta}, fc[2], fc[1], nullptr}};
    Sweep g[3];
    for (int i = 0; i < 3; ++i) g[i] = sweep_plain(0, false, false, nullptr, c.W[i].p);
    fft2d_batch(c, 3, f, g);
    if (c.buoyancy) {
        // r = u eta - v xi + b (inversion.f90:329-331): the x/y transform is linear, so b joins r in semi-spectral space
        PS_LAUNCH((k_add), dim3(stream_blocks(c.nint)), dim3(256), 0, c.stream, (const double*)c.W[0].p, (const double*)c.bsem.p,
                  c.W[0].p, (long long)c.nint);
        ++c.launches;
    }
    SrcArgs a;
    a.r = c.W[0].p; a.q = c.W[1].p; a.p = c.W[2].p;
    a.s0 = c.svorts[0].p; a.s1 = c.svorts[1].p; a.s2 = c.svorts[2].p;
    a.upd = upd; a.c1 = dt2; a.c2 = c2;
    const StepArgs st = step_args(c);
    for (int i = 0; i < 3; ++i) { a.svor[i] = st.svor[i]; a.vortsm[i] = st.wa[i]; a.wb[i] = st.wb[i]; }
    a.f2d = st.f2d; a.filtz = st.filtz; a.vd = st.vd;
    a.mq = st.mq; a.pq = pq ? pq : st.pq;
    launch_src(c, a);
    c.svorts_stale = (upd >= 0);
    if (upd >= 0 && upd < 11) vor_mean(c, 1);          // cn2: adjust_vorticity_mean after every update (cn2.f90:137, 175); rk4: once, at the end
}

static void vor_mean(Ctx& c, int mode) {
    if (c.rank != 0) return;
    PS_LAUNCH((k_vor_mean), dim3(1), dim3(256), 256 * sizeof(double), c.stream, c.svor[0].p, c.svor[1].p,
              (const double*)c.wz.p, c.nz, 1.0 / (double)c.nz, c.ini_mean.p, mode);
    ++c.launches;
}

static FieldPtrs field_ptrs(Ctx& c) {
    FieldPtrs f;
    for (int i = 0; i < 3; ++i) { f.vor[i] = c.vor[i].p; f.vel[i] = c.vel[i].p; }
    return f;
}

// reduces the physical fields into c.red[0..RQ_N)
static void field_reduce(Ctx& c) {
    const long long ncol = (long long)c.nxl * c.ny;
    PS_LAUNCH((k_field_reduce), dim3(c.red_blocks), dim3(RED_THREADS), RED_THREADS * sizeof(double), c.stream,
              field_ptrs(c), ncol, c.nz, c.pz, c.partial.p);
    PS_LAUNCH((k_reduce_final), dim3(1), dim3(RED_THREADS), RED_THREADS * sizeof(double), c.stream,
              (const double*)c.partial.p, c.red_blocks, (int)RQ_N, RQ_OPMASK, c.red.p);
    c.launches += 2;
}

static void do_set_diffusion(Ctx& c, double dt, double pref) {
    if (!c.diffusion_ready) fail(PS3D_ERR_NOT_INITIALISED, "init_diffusion has not been called");
    const long long ncol = (long long)c.nx * c.nyl;
    if (c.stepper == PS3D_STEPPER_CN2) {
        const double dfac = (c.nnu == 1) ? dt : pref * dt;          // cn2.f90:50-56
        PS_LAUNCH((k_step_factors), dim3(stream_blocks(ncol)), dim3(256), 0, c.stream, 0, dfac, (const double*)c.vhdis.p,
                  (const double*)c.filt2d.p, c.fac1.p, c.fac2.p, ncol);
        ++c.launches;
    } else {
        // impl_rk4.f90:43-47: vdiss = dfac * vhdis is the persistent state; emq / epq are rebuilt from it at the start
        // of every impl_rk4_step (:87-89) because the step squares them in place (:151, :185)
        c.rk4_dfac = 0.5 * pref * dt;
        c.rk4_dfac_set = true;
    }
}

// bdiss (cn2.f90:61-75 / impl_rk4.f90:48-50)
static void do_set_diffusion_buoyancy(Ctx& c, double dt, double bf) {
    if (!c.buoyancy) return;
    if (!c.bdiffusion_ready) fail(PS3D_ERR_NOT_INITIALISED, "init_diffusion_buoyancy has not been called");
    const long long ncol = (long long)c.nx * c.nyl;
    if (c.stepper == PS3D_STEPPER_CN2) {
        const double dbac = (c.bnnu == 1) ? dt : bf * dt;
        PS_LAUNCH((k_step_factors), dim3(stream_blocks(ncol)), dim3(256), 0, c.stream, 0, dbac, (const double*)c.bhdis.p,
                  (const double*)c.filt2d.p, c.bfac1.p, c.bfac2.p, ncol);
        ++c.launches;
    } else {
        c.rk4_dbac = 0.5 * bf * dt;
    }
}

static void rk4_factors(Ctx& c) {                               // impl_rk4.f90:87-89
    if (!c.rk4_dfac_set) fail(PS3D_ERR_NOT_INITIALISED, "set_diffusion has not been called");
    const long long ncol = (long long)c.nx * c.nyl;
    PS_LAUNCH((k_step_factors), dim3(stream_blocks(ncol)), dim3(256), 0, c.stream, 1, c.rk4_dfac, (const double*)c.vhdis.p,
              (const double*)c.filt2d.p, c.fac1.p, c.fac2.p, ncol);
    ++c.launches;
    if (c.buoyancy) {                                           // bpq, bmq (:91-95)
        PS_LAUNCH((k_step_factors), dim3(stream_blocks(ncol)), dim3(256), 0, c.stream, 1, c.rk4_dbac, (const double*)c.bhdis.p,
                  (const double*)c.filt2d.p, c.bfac1.p, c.bfac2.p, ncol);
        ++c.launches;
    }
}

static StepArgs step_args(Ctx& c) {
    StepArgs a;
    for (int i = 0; i < 3; ++i) { a.svor[i] = c.svor[i].p; a.svorts[i] = c.svorts[i].p; a.wa[i] = c.wa[i].p; a.wb[i] = c.wb[i].p; }
    a.f2d = c.fac2.p; a.filtz = c.filtz.p; a.mq = c.fac1.p; a.pq = c.fac2.p; a.vd = c.fac1.p;
    a.c1 = a.c2 = 0.0; a.stage = 0;
    a.ncol = (long long)c.nx * c.nyl; a.nz = c.nz; a.pz = c.pz; a.has00 = (c.rank == 0);
    return a;
}

static StepArgs buoy_step_args(Ctx& c) {
    StepArgs a = step_args(c);
    a.svor[0] = c.sbuoy.p; a.svorts[0] = c.sbuoys.p; a.wa[0] = c.bsm.p; a.wb[0] = c.sbuoyf.p;
    a.f2d = c.bfac2.p; a.mq = c.bfac1.p; a.pq = c.bfac2.p; a.vd = c.bfac1.p;
    a.ncomp = 1;
    return a;
}

static void do_adapt(Ctx& c, double t, double t_limit, double alpha, int pretype, int win, double* dt_out, double* diag);

static void do_stepper_setup(Ctx& c, int stepper) {
    if (stepper != PS3D_STEPPER_CN2 && stepper != PS3D_STEPPER_IMPL_RK4)
        fail(PS3D_ERR_BAD_ARGUMENT, "unknown stepper id %d", stepper);
    c.stepper = stepper;
    for (int i = 0; i < 3; ++i) {
        if (!c.wa[i].p) c.wa[i].alloc(c.nint);
        if (stepper == PS3D_STEPPER_IMPL_RK4 && !c.wb[i].p) c.wb[i].alloc(c.nint);
    }
    if (c.buoyancy && stepper == PS3D_STEPPER_IMPL_RK4 && !c.sbuoyf.p) c.sbuoyf.alloc(c.nint);    // impl_rk4.f90:67-72
    c.stepper_ready = true;
}

static void cn2_update(Ctx& c, double dt2, int stage) {
    StepArgs a = step_args(c);
    a.c1 = dt2; a.stage = stage;
    PS_LAUNCH((k_cn2_update), dim3(stream_blocks(c.nint)), dim3(256), 0, c.stream, a);
    ++c.launches;
    vor_mean(c, 1);
}

// sbuoy update of cn2_step (cn2.f90:107-117 stage 0, :151-160 stage 1), combine -> bdiss -> decompose collapsed
static void cn2_update_buoy(Ctx& c, double dt2, int stage) {
    if (!c.buoyancy) return;
    StepArgs a = buoy_step_args(c);
    a.c1 = dt2; a.stage = stage;
    PS_LAUNCH((k_cn2_update), dim3(stream_blocks(c.nint)), dim3(256), 0, c.stream, a);
    ++c.launches;
}

static void rk4_update_buoy(Ctx& c, int stage, double c1, double c2, const double* pq) {
    if (!c.buoyancy) return;                           // sbuoy (impl_rk4.f90:99-105, 124-131, 154-164, 187-195)
    StepArgs b = buoy_step_args(c);
    b.c1 = c1; b.c2 = c2; b.stage = stage; b.pq = (stage == 1) ? pq : c.bfac2.p;
    PS_LAUNCH((k_rk4_update), dim3(stream_blocks(c.nint)), dim3(256), 0, c.stream, b);
    ++c.launches;
}

static void rk4_update(Ctx& c, int stage, double c1, double c2, const double* pq) {
    rk4_update_buoy(c, stage, c1, c2, pq);
    StepArgs a = step_args(c);
    a.c1 = c1; a.c2 = c2; a.stage = stage; a.pq = pq;
    PS_LAUNCH((k_rk4_update), dim3(stream_blocks(c.nint)), dim3(256), 0, c.stream, a);
    ++c.launches;
}

static void square_factor(Ctx& c, int mode) {                  // emq = emq**2 / epq = epq**2 (impl_rk4.f90:151,185)
    const long long ncol = (long long)c.nx * c.nyl;
    PS_LAUNCH((k_step_factors), dim3(stream_blocks(ncol)), dim3(256), 0, c.stream, mode, 0.0, (const double*)nullptr,
              (const double*)nullptr, c.fac1.p, c.fac2.p, ncol);
    ++c.launches;
    if (c.buoyancy) {                                          // bmq = bmq**2 / bpq = bpq**2 (:155, :188)
        PS_LAUNCH((k_step_factors), dim3(stream_blocks(ncol)), dim3(256), 0, c.stream, mode, 0.0, (const double*)nullptr,
                  (const double*)nullptr, c.bfac1.p, c.bfac2.p, ncol);
        ++c.launches;
    }
}

// the cn2 update can ride on the source kernel (power-of-two nz; PS3D_NO_FUSED_UPDATE=1 keeps it separate)
static bool cn2_fused(const Ctx& c) { return c.stepper == PS3D_STEPPER_CN2 && !c.gen[2] && c.fuse_update; }
// ... and so can the four substeps of impl-diff-rk4
static bool rk4_fused(const Ctx& c) { return c.stepper == PS3D_STEPPER_IMPL_RK4 && !c.gen[2] && c.fuse_update; }

// first_update_done: the source call before this step already carried the first update (ps3d_cuda_advance)
static void do_step(Ctx& c, double* t, double dt, bool first_update_done = false) {
    if (!c.stepper_ready) fail(PS3D_ERR_NOT_INITIALISED, "stepper_setup has not been called");
    // bstep%step consumes the tendency of the current state (advance.f90:95-102).  The time loop does not keep svorts
    // when an update rode on the source kernel, so a step without a source call since the last one has nothing to consume
    if (!first_update_done && c.svorts_stale)
        fail(PS3D_ERR_NOT_INITIALISED, "ps3d_cuda_step: call ps3d_cuda_source for the current state first (advance.f90:95-102)");
    if (c.stepper == PS3D_STEPPER_CN2) {
        const double dt2 = 0.5 * dt;                       // cn2.f90:101
        if (!first_update_done) { cn2_update_buoy(c, dt2, 0); cn2_update(c, dt2, 0); }     // :107-137
        for (int iter = 0; iter < 2; ++iter) {             // niter = 2 (:34, :143-177)
            do_vor2vel(c);
            if (cn2_fused(c)) {
                do_source(c, 1, dt2);
                cn2_update_buoy(c, dt2, 1);                // (its operands sbuoys / bsm are untouched by the source kernel)
            } else {
                do_source(c);
                cn2_update_buoy(c, dt2, 1);
                cn2_update(c, dt2, 1);
            }
        }
        *t += dt;
    } else {
        const double dt2 = 0.5 * dt, dt3 = dt / 3.0, dt6 = dt / 6.0;   // impl_rk4.f90:82-84
        if (rk4_fused(c)) {
            // every substep rides on the source kernel that produces its tendency (the factor tables are squared
            // before the call that uses them: neither vor2vel nor the tendency read them)
            if (!first_update_done) {
                rk4_factors(c);                                // epq, emq (:87-89)
                rk4_update(c, 1, dt2, dt6, c.filt2d.p);        // (the source call before this step was a plain one)
            }
            do_vor2vel(c);
            do_source(c, 12, dt2, dt3, c.fac2.p);
            rk4_update_buoy(c, 2, dt2, dt3, c.fac2.p);
            *t += dt2;
            square_factor(c, 2);                               // emq = emq**2 (:151)
            do_vor2vel(c);
            do_source(c, 13, dt, dt3, c.fac2.p);
            rk4_update_buoy(c, 3, dt, dt3, c.fac2.p);
            *t += dt2;
            square_factor(c, 3);                               // epq = epq**2 (:185)
            do_vor2vel(c);
            do_source(c, 14, dt6, 0.0, c.fac2.p);
            rk4_update_buoy(c, 4, dt6, 0.0, c.fac2.p);
            vor_mean(c, 1);                                    // :205
            return;
        }
        rk4_factors(c);                                    // epq, emq (:87-89)
        // substep one filters the source with filt(0,:,:) (:227-229)
        rk4_update(c, 1, dt2, dt6, c.filt2d.p);
        do_vor2vel(c);
        do_source(c);
        *t += dt2;
        rk4_update(c, 2, dt2, dt3, c.fac2.p);
        do_vor2vel(c);
        do_source(c);
        *t += dt2;
        square_factor(c, 2);                               // emq = emq**2 (:151)
        rk4_update(c, 3, dt, dt3, c.fac2.p);
        do_vor2vel(c);
        do_source(c);
        square_factor(c, 3);                               // epq = epq**2 (:185)
        rk4_update(c, 4, dt6, 0.0, c.fac2.p);
        vor_mean(c, 1);                                    // :205
    }
}

// the five velocity-gradient fields of advance.f90:199-217 -> W[0..4] = du/dx, du/dy, dw/dx, dv/dy, dw/dy
static void strain_fields(Ctx& c) {
    const int comp[5] = {0, 0, 2, 1, 2};
    const bool ddx_[5] = {true, false, true, false, false};
    if (c.velx_valid) {
        // d/dy fields: the x sweep of u, v, w was done by vor2vel and kept; only the y sweep with the derivative
        // prologue remains (3 sweeps instead of 6)
        for (int i = 0; i < 5; ++i) {
            if (ddx_[i]) continue;
            Sweep g = sweep_plain(1, true, true, c.velx[comp[i]].p, c.W[i].p);
            run_sweep(c, g);
        }
        Sweep f[2], g[2];
        int k = 0;
        for (int i = 0; i < 5; ++i) {
            if (!ddx_[i]) continue;
            f[k] = sweep_plain(0, true, true, c.svel[comp[i]].p, nullptr);
            g[k] = sweep_plain(1, true, false, nullptr, c.W[i].p);
            ++k;
        }
        fft2d_batch(c, 2, f, g);
        return;
    }
    Sweep f[5], g[5];
    for (int i = 0; i < 5; ++i) {
        f[i] = sweep_plain(0, true, ddx_[i], c.svel[comp[i]].p, nullptr);
        g[i] = sweep_plain(1, true, !ddx_[i], nullptr, c.W[i].p);
    }
    fft2d_batch(c, 5, f, g);
}

// device-side all-reduce of red[0..n): sums and maxima according to opmask, result back in red (no host round trip)
static void allreduce_dev(Ctx& c, double* red, int n, unsigned opmask) {
    Transport& t = c.tr;
    if (t.nranks == 1) return;
#ifndef PS3D_EMU
    if (t.p2p && t.peer_sync && !t.ar_cb && n <= 32) {
        PS_LAUNCH((k_peer_allreduce), dim3(1), dim3(32), 0, c.stream, c.mail, c.rank, c.nranks, ++t.ar_epoch, n, opmask, red);
        ++c.launches;
        return;
    }
    if (t.have_nccl() && !t.ar_cb) {
        ps_d2d(c.redS.p, red, n * sizeof(double), c.stream);
        ps_d2d(c.redM.p, red, n * sizeof(double), c.stream);
        int rc = t.nccl.AllReduce(c.redS.p, c.redS.p, n, NcclApi::kFloat64, NcclApi::kSum, t.comm, (void*)c.stream);
        if (rc == 0) rc = t.nccl.AllReduce(c.redM.p, c.redM.p, n, NcclApi::kFloat64, NcclApi::kMax, t.comm, (void*)c.stream);
        if (rc != 0) fail(PS3D_ERR_DEVICE, "NCCL all-reduce failed: %s", t.nccl.GetErrorString(rc));
        PS_LAUNCH((k_select_reduced), dim3(1), dim3(64), 0, c.stream, (const double*)c.redS.p, (const double*)c.redM.p, n, opmask, red);
        ++c.launches;
        return;
    }
#endif
    // host-supplied collectives (ps3d_cuda_set_transport): through the host
    double h[64];
    ps_d2h(h, red, n * sizeof(double), c.stream);
    ps_sync(c.stream);
    allreduce_host(c, h, n, opmask);
    ps_h2d(red, h, n * sizeof(double), c.stream);
    ps_sync(c.stream);
}

// bfmax (advance.f90:147-168): the three gradient components of b' in physical space, max of |grad b + N^2 z^|^2
// -> red[RQ_N + 5] (the fourth root is taken on the host)
static void do_bfmax(Ctx& c) {
    const long long ncol = (long long)c.nxl * c.ny;
    launch_zop(c, ZOP_COMBINE, c.sbuoy.p, c.bsem.p);                 // field_combine_semi_spectral(sbuoy) (:150)
    fft2d_inv(c, c.bsem.p, c.W[0].p, true, false);                   // diffx + fftxys2p (:151-152)
    fft2d_inv(c, c.bsem.p, c.W[1].p, false, true);                   // diffy + fftxys2p (:154-155)
    launch_zop(c, ZOP_DIFFZ, c.bsem.p, c.W[3].p);                    // central_diffz (:157)
    fft2d_inv(c, c.W[3].p, c.W[2].p, false, false);                  // (:158)
    PS_LAUNCH((k_bfmax), dim3(c.red_blocks), dim3(RED_THREADS), RED_THREADS * sizeof(double), c.stream,
              (const double*)c.W[0].p, (const double*)c.W[1].p, (const double*)c.W[2].p, c.bfsq, ncol, c.nz, c.pz, c.partial.p);
    PS_LAUNCH((k_reduce_final), dim3(1), dim3(RED_THREADS), RED_THREADS * sizeof(double), c.stream,
              (const double*)c.partial.p, c.red_blocks, 1, 1u, c.red.p + RQ_N + 5);
    c.launches += 2;
    allreduce_dev(c, c.red.p + RQ_N + 5, 1, 1u);                     // buf(1) of the MPI_MAX reduction (:291-305)
}

static void do_adapt(Ctx& c, double t, double t_limit, double alpha, int pretype, int win, double* dt_out, double* diag) {
    const long long ncol = (long long)c.nxl * c.ny;
    if (c.buoyancy) do_bfmax(c);
    // first reduction (advance.f90:171-185, 285-313) -> red[0..RQ_N), reduced over the ranks on the device
    field_reduce(c);
    allreduce_dev(c, c.red.p, RQ_N, RQ_OPMASK);             // advance.f90:299-305, field_diagnostics.f90:418-424
    // velocity strain (advance.f90:199-217): derivative folded into the inverse sweeps
    strain_fields(c);
    StrainPtrs sp;
    sp.dudx = c.W[0].p; sp.dudy = c.W[1].p; sp.dwdx = c.W[2].p; sp.dvdy = c.W[3].p; sp.dwdy = c.W[4].p;
    for (int i = 0; i < 3; ++i) sp.vor[i] = c.vor[i].p;
    // second reduction: strain maxima and the characteristic vorticity (needs vortrms of the first, read on the
    // device) in one pass -> red[RQ_N .. RQ_N + 5) = ggmax, usggmax, lsggmax, vorl1, vorl2
    PS_LAUNCH((k_strain), dim3(c.red_blocks), dim3(RED_THREADS), RED_THREADS * sizeof(double), c.stream, sp, ncol, c.nz,
              c.pz, c.strict_jacobi, (const double*)c.red.p, (double)c.ncell, c.partial.p);
    PS_LAUNCH((k_reduce_final), dim3(1), dim3(RED_THREADS), RED_THREADS * sizeof(double), c.stream,
              (const double*)c.partial.p, c.red_blocks, 5, 7u, c.red.p + RQ_N);
    c.launches += 2;
    allreduce_dev(c, c.red.p + RQ_N, 5, 7u);                // sums: vorl1, vorl2 (field_diagnostics.f90:529-535); max: strain
    // the one device -> host read of the step
    ps_d2h(c.h_red, c.red.p, (RQ_N + 6) * sizeof(double), c.stream);
    ps_sync(c.stream);
    double r1[RQ_N];
    for (int i = 0; i < RQ_N; ++i) r1[i] = c.h_red[i];
    const double vortmax = std::sqrt(r1[RQ_MAXW2]);
    const double vortrms = std::sqrt(r1[RQ_SUMW2] / (double)c.ncell);
    const double* h2 = c.h_red + RQ_N;
    const double small = 1.0e-12, cflmax = 0.8;                    // constants.f90:63-65
    // vorl1 starts from `small` on every rank before the reduction (field_diagnostics.f90:509,529-535)
    const double vorl1 = small * (double)c.nranks + h2[3], vorl2 = h2[4];
    const double vorch = vorl2 / vorl1;
    const double bfmax = c.buoyancy ? std::sqrt(std::sqrt(h2[5])) : 0.0;    // advance.f90:145, 167
    const double ggmax = std::max(2.220446049250313e-16, h2[0]);   // ggmax = epsilon(ggmax) (advance.f90:222)
    const double usggmax = std::max(0.0, h2[1]), lsggmax = std::max(0.0, h2[2]);
    const double umax = r1[RQ_MAXU], vmax = r1[RQ_MAXV], wmax = r1[RQ_MAXWV];
    const double dtcfl = cflmax * std::min(std::min(c.dx[0] / (umax + small), c.dx[1] / (vmax + small)),
                                           c.dx[2] / (wmax + small));
    const double dt = std::min(std::min(alpha / (ggmax + small), alpha / (bfmax + small)), std::min(dtcfl, t_limit - t));
    if (!c.rollmean.alloc(win))
        fail(PS3D_ERR_BAD_ARGUMENT, "roll_mean_win_size changed from %d to %d (rolling_mean.f90: allocated once)", c.rollmean.length, win);
    double rmb = 0.0;
    if (c.buoyancy) {                                              // advance.f90:358-362
        if (!c.bdiffusion_ready) fail(PS3D_ERR_NOT_INITIALISED, "init_diffusion_buoyancy has not been called");
        if (!c.buoy_rollmean.alloc(c.bwin))
            fail(PS3D_ERR_BAD_ARGUMENT, "buoyancy roll_mean_win_size changed (rolling_mean.f90: allocated once)");
        rmb = c.buoy_rollmean.get_next(bfmax);
    }
    const double rmv = c.rollmean.get_next(ggmax);
    auto prefactor = [&](int type) -> double {                     // get_diffusion_pre_factor (advance.f90:381-410)
        switch (type) {
            case PS3D_PRE_CONSTANT: return 1.0;
            case PS3D_PRE_VORCH: return vorch;
            case PS3D_PRE_BFMAX: return bfmax;
            case PS3D_PRE_ROLL_MEAN_MAX_STRAIN: return rmv;
            case PS3D_PRE_MAX_STRAIN: return ggmax;
            case PS3D_PRE_US_MAX_STRAIN: return usggmax;
            case PS3D_PRE_ROLL_MEAN_BFMAX: if (c.buoyancy) return rmb; break;      // (#ifdef ENABLE_BUOYANCY, :395-398)
            default: break;
        }
        fail(PS3D_ERR_BAD_ARGUMENT, "We only support 'constant', 'vorch', 'bfmax', 'roll-mean-max-strain', "
                                    "'roll-mean-bfmax', 'max-strain' and us-max-strain");
    };
    const double pref = prefactor(pretype);                        // vval (:370)
    const double bval = c.buoyancy ? prefactor(c.bpretype) : 0.0;  // bval (:372-374)
    c.last_bdiag[0] = bfmax; c.last_bdiag[1] = rmb; c.last_bdiag[2] = bval; c.last_bdiag[3] = c.bvisc;
    if (diag) {
        const double ncelli = 1.0 / (double)c.ncell;
        diag[PS3D_D_VORTMAX] = vortmax; diag[PS3D_D_VORTRMS] = vortrms; diag[PS3D_D_VORCH] = vorch;
        diag[PS3D_D_VORMEAN_X] = r1[RQ_SUMW0] * ncelli; diag[PS3D_D_VORMEAN_Y] = r1[RQ_SUMW1] * ncelli;
        diag[PS3D_D_VORMEAN_Z] = r1[RQ_SUMW2C] * ncelli;
        diag[PS3D_D_BFMAX] = bfmax; diag[PS3D_D_GGMAX] = ggmax; diag[PS3D_D_UMAX] = umax; diag[PS3D_D_VMAX] = vmax;
        diag[PS3D_D_WMAX] = wmax; diag[PS3D_D_USGGMAX] = usggmax; diag[PS3D_D_LSGGMAX] = lsggmax;
        diag[PS3D_D_RMV] = rmv; diag[PS3D_D_DT] = dt; diag[PS3D_D_PREFACTOR] = pref;
    }
    {
        const double ncelli = 1.0 / (double)c.ncell;
        double* d = c.last_diag;
        d[PS3D_D_VORTMAX] = vortmax; d[PS3D_D_VORTRMS] = vortrms; d[PS3D_D_VORCH] = vorch;
        d[PS3D_D_VORMEAN_X] = r1[RQ_SUMW0] * ncelli; d[PS3D_D_VORMEAN_Y] = r1[RQ_SUMW1] * ncelli;
        d[PS3D_D_VORMEAN_Z] = r1[RQ_SUMW2C] * ncelli;
        d[PS3D_D_BFMAX] = bfmax; d[PS3D_D_GGMAX] = ggmax; d[PS3D_D_UMAX] = umax; d[PS3D_D_VMAX] = vmax;
        d[PS3D_D_WMAX] = wmax; d[PS3D_D_USGGMAX] = usggmax; d[PS3D_D_LSGGMAX] = lsggmax;
        d[PS3D_D_RMV] = rmv; d[PS3D_D_DT] = dt; d[PS3D_D_PREFACTOR] = pref;
        c.have_diag = true;
    }
    *dt_out = dt;
    if (c.stepper_ready && c.diffusion_ready) {                                  // advance.f90:375
        do_set_diffusion(c, dt, pref);
        do_set_diffusion_buoyancy(c, dt, bval);
    }
}

static void do_upload_vorticity(Ctx& c, const double* vor_phys) {
    // utils.f90:160-165
    for (int i = 0; i < 3; ++i) {
        to_device(c, vor_phys + (size_t)i * c.nnat, c.vor[i].p, false);
        fft2d_fwd(c, c.vor[i].p, c.W[0].p);
        launch_zop(c, ZOP_DECOMPOSE, c.W[0].p, c.svor[i].p);
    }
    vor_mean(c, 0);
    ps_sync(c.stream);
}

// streamed form of do_upload_vorticity: _begin queues the three host -> device copies on the copy stream and
// returns; _end makes the compute stream wait for them, then repacks and decomposes (utils.f90:160-165)
static void do_upload_begin(Ctx& c, const double* vor_phys) {
    if (c.upload_pending >= 2) fail(PS3D_ERR_BAD_ARGUMENT, "upload_vorticity_begin: two uploads are already queued (call upload_vorticity_end)");
    const int slot = (c.upload_head + c.upload_pending) & 1;
    if (!c.stage3[slot].p) c.stage3[slot].alloc(3 * c.nnat);
#ifndef PS3D_EMU
    if (!c.copy_stream) {
        PS_CUDA_TRY(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) PS_CUDA_TRY(cudaEventCreateWithFlags(&c.ev_copy[i], cudaEventDisableTiming));
    }
#endif
    ps_h2d(c.stage3[slot].p, vor_phys, 3 * c.nnat * sizeof(double), c.copy_stream);
#ifndef PS3D_EMU
    PS_CUDA_TRY(cudaEventRecord(c.ev_copy[slot], c.copy_stream));
#endif
    ++c.upload_pending;
}

static void do_upload_end(Ctx& c) {
    if (c.upload_pending < 1) fail(PS3D_ERR_BAD_ARGUMENT, "upload_vorticity_end without upload_vorticity_begin");
    const int slot = c.upload_head;
#ifndef PS3D_EMU
    PS_CUDA_TRY(cudaStreamWaitEvent(c.stream, c.ev_copy[slot], 0));
#endif
    for (int i = 0; i < 3; ++i) {
        PS_LAUNCH((k_repack_in), dim3(stream_blocks(c.nint)), dim3(256), 0, c.stream, (const double*)(c.stage3[slot].p + (size_t)i * c.nnat),
                  c.vor[i].p, c.nxl, c.ny, c.nzp, c.pz, (const int*)nullptr);
        ++c.launches;
        fft2d_fwd(c, c.vor[i].p, c.W[0].p);
        launch_zop(c, ZOP_DECOMPOSE, c.W[0].p, c.svor[i].p);
    }
    vor_mean(c, 0);
    ps_sync(c.stream);
    c.upload_head ^= 1;
    --c.upload_pending;
}

// pressure (fields_derived.f90:67-157), lazily: needs the five strain fields
// pressure (fields_derived.f90:67-157), evaluated on demand from the current svel / vor -> W[0]
static void do_pressure(Ctx& c) {
    strain_fields(c);
    const long long n = (long long)c.nint;
    // W[0..4] = du/dx, du/dy, dw/dx, dv/dy, dw/dy
    PS_LAUNCH((k_pressure_rhs), dim3(stream_blocks(c.nint)), dim3(256), 0, c.stream, (const double*)c.W[0].p,
              (const double*)c.W[1].p, (const double*)c.W[3].p, (const double*)c.W[2].p, (const double*)c.W[4].p,
              (const double*)c.vor[0].p, (const double*)c.vor[1].p, (const double*)c.vor[2].p, c.W[0].p, n);
    ++c.launches;
    if (c.buoyancy) {
        // pres = pres + dbdz + f_cor(3) * zeta (fields_derived.f90:108-112); central_diffz commutes with the x/y
        // transform, so db/dz is taken on the semi-spectral b' and brought to physical space
        launch_zop(c, ZOP_COMBINE, c.sbuoy.p, c.bsem.p);
        launch_zop(c, ZOP_DIFFZ, c.bsem.p, c.W[2].p);
        fft2d_inv(c, c.W[2].p, c.W[1].p, false, false);
        PS_LAUNCH((k_add_axpy), dim3(stream_blocks(c.nint)), dim3(256), 0, c.stream, (const double*)c.W[0].p,
                  (const double*)c.W[1].p, c.f_cor[2], (const double*)c.vor[2].p, c.W[0].p, n);
        ++c.launches;
    }
    fft2d_fwd(c, c.W[0].p, c.W[1].p);
    launch_zop(c, ZOP_POISSON, c.W[1].p, c.W[2].p);
    fft2d_inv(c, c.W[2].p, c.W[0].p, false, false);
}

// horizontal_divergence (fields_derived.f90:161-182): delta = u_x + v_y -> W[0]
static void do_delta(Ctx& c) {
    fft2d_inv(c, c.svel[0].p, c.W[0].p, true, false);
    fft2d_inv(c, c.svel[1].p, c.W[1].p, false, true);
    PS_LAUNCH((k_add), dim3(stream_blocks(c.nint)), dim3(256), 0, c.stream, (const double*)c.W[0].p, (const double*)c.W[1].p,
              c.W[0].p, (long long)c.nint);
    ++c.launches;
}

}  // namespace ps3d

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
using namespace ps3d;

#define PS_API_BEGIN try {
#define PS_API_END                                                        \
    return PS3D_OK;                                                       \
    } catch (const StatusError& e) { return e.code; }                     \
    catch (const DeviceError&) { return PS3D_ERR_DEVICE; }                \
    catch (const std::exception& e) { set_error("%s", e.what()); return PS3D_ERR_DEVICE; }

extern "C" {

const char* ps3d_cuda_last_error(void) { return g_last_error.c_str(); }

int ps3d_cuda_init(int nx, int ny, int nz, const double lower[3], const double extent[3], int rank, int nranks,
                   const void* nccl_id) {
    PS_API_BEGIN
    const bool fresh = (g_ctx == nullptr);
    try { do_init(nx, ny, nz, lower, extent, rank, nranks, nccl_id); }
    catch (...) { if (fresh && g_ctx) { try { do_finalise(); } catch (...) {} } throw; }
    PS_API_END
}

int ps3d_cuda_init_inversion(int filtering_id) { PS_API_BEGIN do_init_inversion(filtering_id); PS_API_END }

int ps3d_cuda_init_diffusion(int nnu, double prediss, int length_scale_id, double te, double en, double* nu_out) {
    PS_API_BEGIN
    const double v = do_init_diffusion(nnu, prediss, length_scale_id, te, en);
    if (nu_out) *nu_out = v;
    PS_API_END
}

int ps3d_cuda_finalise(void) { PS_API_BEGIN do_finalise(); PS_API_END }

int ps3d_cuda_fftxyp2s(const double* fp, double* fs) {
    PS_API_BEGIN
    Ctx& c = ctx();
    to_device(c, fp, c.W[0].p, false);
    fft2d_fwd(c, c.W[0].p, c.W[1].p);
    to_host(c, c.W[1].p, fs, true);
    PS_API_END
}

int ps3d_cuda_fftxys2p(const double* fs, double* fp) {
    PS_API_BEGIN
    Ctx& c = ctx();
    to_device(c, fs, c.W[0].p, true);
    fft2d_inv(c, c.W[0].p, c.W[1].p, false, false);
    to_host(c, c.W[1].p, fp, false);
    PS_API_END
}

static void zop_host(int op, const double* in, double* out) {
    Ctx& c = ready();
    to_device(c, in, c.W[0].p, true);
    launch_zop(c, op, c.W[0].p, c.W[1].p);
    to_host(c, c.W[1].p, out, true);
}

int ps3d_cuda_fftsine(double* fs) { PS_API_BEGIN zop_host(ZOP_SINE, fs, fs); PS_API_END }
int ps3d_cuda_fftcosine(double* fs) { PS_API_BEGIN zop_host(ZOP_COSINE, fs, fs); PS_API_END }
int ps3d_cuda_diffx(const double* fs, double* ds) { PS_API_BEGIN zop_host(ZOP_DIFFX, fs, ds); PS_API_END }
int ps3d_cuda_diffy(const double* fs, double* ds) { PS_API_BEGIN zop_host(ZOP_DIFFY, fs, ds); PS_API_END }
int ps3d_cuda_central_diffz(const double* fs, double* ds) { PS_API_BEGIN zop_host(ZOP_DIFFZ, fs, ds); PS_API_END }
int ps3d_cuda_diffz(const double* fs, double* ds) {        // mixed-spectral in, mixed-spectral out (buoyancy build)
    PS_API_BEGIN
    zop_host(ZOP_DIFFZ_SPEC, fs, ds);
    PS_API_END
}
int ps3d_cuda_field_combine_semi_spectral(double* sf) { PS_API_BEGIN zop_host(ZOP_COMBINE, sf, sf); PS_API_END }
int ps3d_cuda_field_decompose_semi_spectral(double* sfc) { PS_API_BEGIN zop_host(ZOP_DECOMPOSE, sfc, sfc); PS_API_END }

int ps3d_cuda_field_combine_physical(const double* sf, double* fc) {
    PS_API_BEGIN
    Ctx& c = ready();
    to_device(c, sf, c.W[0].p, true);
    launch_zop(c, ZOP_COMBINE, c.W[0].p, c.W[1].p);
    fft2d_inv(c, c.W[1].p, c.W[0].p, false, false);
    to_host(c, c.W[0].p, fc, false);
    PS_API_END
}

int ps3d_cuda_field_decompose_physical(const double* fc, double* sf) {
    PS_API_BEGIN
    Ctx& c = ready();
    to_device(c, fc, c.W[0].p, false);
    fft2d_fwd(c, c.W[0].p, c.W[1].p);
    launch_zop(c, ZOP_DECOMPOSE, c.W[1].p, c.W[0].p);
    to_host(c, c.W[0].p, sf, true);
    PS_API_END
}

int ps3d_cuda_upload_vorticity(const double* vor_phys) { PS_API_BEGIN do_upload_vorticity(ready(), vor_phys); PS_API_END }
int ps3d_cuda_upload_vorticity_begin(const double* vor_phys) {
    PS_API_BEGIN
    if (!vor_phys) fail(PS3D_ERR_BAD_ARGUMENT, "vor_phys is null");
    do_upload_begin(ready(), vor_phys);
    PS_API_END
}
int ps3d_cuda_upload_vorticity_end(void) { PS_API_BEGIN do_upload_end(ready()); PS_API_END }
int ps3d_cuda_vor2vel(void) { PS_API_BEGIN Ctx& c = ready(); do_vor2vel(c); ps_sync(c.stream); PS_API_END }
int ps3d_cuda_source(void) { PS_API_BEGIN Ctx& c = ready(); do_source(c); ps_sync(c.stream); PS_API_END }

int ps3d_cuda_adapt(double t, double t_limit, double alpha, int pretype_id, int win, double* dt, double diag_out[16]) {
    PS_API_BEGIN
    if (!dt) fail(PS3D_ERR_BAD_ARGUMENT, "dt is null");
    if (win < 1) fail(PS3D_ERR_BAD_ARGUMENT, "roll_mean_win_size must be >= 1");
    do_adapt(ready(), t, t_limit, alpha, pretype_id, win, dt, diag_out);
    PS_API_END
}

int ps3d_cuda_set_physics(const double f_cor[3], double bfsq) {
    PS_API_BEGIN
    Ctx& c = ctx();
    if (!f_cor) fail(PS3D_ERR_BAD_ARGUMENT, "f_cor is null");
    for (int i = 0; i < 3; ++i) c.f_cor[i] = f_cor[i];
    c.bfsq = bfsq;
    PS_API_END
}

int ps3d_cuda_enable_buoyancy(void) {
    PS_API_BEGIN
    Ctx& c = ready();
    if (c.buoyancy) return PS3D_OK;
    DevBuf<double>* f[] = {&c.sbuoy, &c.buoy, &c.sbuoys, &c.bsm, &c.bsem};
    for (auto* b : f) { b->alloc(c.nint); ps_memset(b->p, 0, c.nint * sizeof(double), c.stream); }
    const size_t ncol = (size_t)c.nx * c.nyl;
    c.bhdis.alloc(ncol); c.bfac1.alloc(ncol); c.bfac2.alloc(ncol);
    if (c.stepper_ready && c.stepper == PS3D_STEPPER_IMPL_RK4) c.sbuoyf.alloc(c.nint);
    ps_sync(c.stream);
    c.buoyancy = true;
    PS_API_END
}

int ps3d_cuda_upload_buoyancy(const double* buoy_phys) {
    PS_API_BEGIN
    Ctx& c = ready();
    if (!c.buoyancy) fail(PS3D_ERR_NOT_INITIALISED, "enable_buoyancy has not been called");
    if (!buoy_phys) fail(PS3D_ERR_BAD_ARGUMENT, "buoy_phys is null");
    to_device(c, buoy_phys, c.buoy.p, false);                 // utils.f90:158 field_decompose_physical(buoy, sbuoy)
    fft2d_fwd(c, c.buoy.p, c.W[0].p);
    launch_zop(c, ZOP_DECOMPOSE, c.W[0].p, c.sbuoy.p);
    ps_sync(c.stream);
    PS_API_END
}

int ps3d_cuda_init_diffusion_buoyancy(int nnu, double prediss, int length_scale_id, double te, double en, int pretype_id,
                                      int roll_mean_win_size, double* nu_out) {
    PS_API_BEGIN
    Ctx& c = ready();
    if (!c.buoyancy) fail(PS3D_ERR_NOT_INITIALISED, "enable_buoyancy has not been called");
    if (roll_mean_win_size < 1) fail(PS3D_ERR_BAD_ARGUMENT, "roll_mean_win_size must be >= 1");
    if (pretype_id < PS3D_PRE_CONSTANT || pretype_id > PS3D_PRE_ROLL_MEAN_BFMAX) fail(PS3D_ERR_BAD_ARGUMENT, "unknown pretype id %d", pretype_id);
    std::vector<double> vh;
    c.bvisc = viscosity_table(c, nnu, prediss, length_scale_id, te, en, vh);     // inversion_utils.f90:144-151
    c.bnnu = nnu; c.bpretype = pretype_id; c.bwin = roll_mean_win_size;
    c.bhdis.upload(vh, c.stream);
    c.bdiffusion_ready = true;
    if (nu_out) *nu_out = c.bvisc;
    PS_API_END
}

int ps3d_cuda_set_diffusion_buoyancy(double dt, double bf) {
    PS_API_BEGIN
    Ctx& c = ready();
    if (!c.buoyancy) fail(PS3D_ERR_NOT_INITIALISED, "enable_buoyancy has not been called");
    do_set_diffusion_buoyancy(c, dt, bf);
    ps_sync(c.stream);
    PS_API_END
}

int ps3d_cuda_buoyancy_diag(double out[4]) {
    PS_API_BEGIN
    Ctx& c = ready();
    if (!out) fail(PS3D_ERR_BAD_ARGUMENT, "out is null");
    for (int i = 0; i < 4; ++i) out[i] = c.last_bdiag[i];
    PS_API_END
}

int ps3d_cuda_stepper_setup(int stepper_id) { PS_API_BEGIN do_stepper_setup(ready(), stepper_id); PS_API_END }
int ps3d_cuda_set_diffusion(double dt, double pref) { PS_API_BEGIN Ctx& c = ready(); do_set_diffusion(c, dt, pref); ps_sync(c.stream); PS_API_END }
int ps3d_cuda_step(double* t, double dt) { PS_API_BEGIN Ctx& c = ready(); if (!t) fail(PS3D_ERR_BAD_ARGUMENT, "t is null"); do_step(c, t, dt); ps_sync(c.stream); PS_API_END }

int ps3d_cuda_advance(double* t, double t_limit, double alpha, int pretype_id, int win, double* dt_out, double diag_out[16]) {
    PS_API_BEGIN
    Ctx& c = ready();
    if (!t) fail(PS3D_ERR_BAD_ARGUMENT, "t is null");
    if (win < 1) fail(PS3D_ERR_BAD_ARGUMENT, "roll_mean_win_size must be >= 1");
    if (!c.stepper_ready || !c.diffusion_ready) fail(PS3D_ERR_NOT_INITIALISED, "stepper_setup / init_diffusion missing");
#ifndef PS3D_EMU
    PS_CUDA_TRY(cudaEventRecord(c.ev0, c.stream));
#endif
    double dt = 0.0;
    do_vor2vel(c);                                     // advance.f90:85
    do_adapt(c, *t, t_limit, alpha, pretype_id, win, &dt, diag_out);   // :88
    if (cn2_fused(c)) {
        do_source(c, 0, 0.5 * dt);                     // :95 + the first update of cn2_step (cn2.f90:120-137)
        cn2_update_buoy(c, 0.5 * dt, 0);               // cn2.f90:107-117
        do_step(c, t, dt, true);                       // :102
    } else if (rk4_fused(c)) {
        rk4_factors(c);                                // impl_rk4.f90:87-95
        do_source(c, 11, 0.5 * dt, dt / 6.0, c.filt2d.p);       // :95 + substep one (impl_rk4.f90:212-241)
        rk4_update_buoy(c, 1, 0.5 * dt, dt / 6.0, c.filt2d.p);
        do_step(c, t, dt, true);                       // :102
    } else {
        do_source(c);                                  // :95
        do_step(c, t, dt);                             // :102
    }
#ifndef PS3D_EMU
    PS_CUDA_TRY(cudaEventRecord(c.ev1, c.stream));
    PS_CUDA_TRY(cudaEventSynchronize(c.ev1));
    float ms = 0.f;
    PS_CUDA_TRY(cudaEventElapsedTime(&ms, c.ev0, c.ev1));
    c.last_advance_ms = ms;
#else
    ps_sync(c.stream);
#endif
    if (dt_out) *dt_out = dt;
    PS_API_END
}

static DevBuf<double>* field_by_id(Ctx& c, int id, bool& spectral) {
    switch (id) {
        case PS3D_F_SVOR: spectral = true; return c.svor;
        case PS3D_F_VOR: spectral = false; return c.vor;
        case PS3D_F_VEL: spectral = false; return c.vel;
        case PS3D_F_SVEL: spectral = true; return c.svel;
        case PS3D_F_SVORTS: spectral = true; return c.svorts;
        default: return nullptr;
    }
}

int ps3d_cuda_download(int field_id, int comp, double* host) {
    PS_API_BEGIN
    Ctx& c = ready();
    if (!host) fail(PS3D_ERR_BAD_ARGUMENT, "host pointer is null");
    if (field_id == PS3D_F_PRES || field_id == PS3D_F_DELTA) {
        // output-only quantities of adapt (advance.f90:278-282), evaluated lazily at output cadence
        if (field_id == PS3D_F_DELTA) do_delta(c); else do_pressure(c);
        to_host(c, c.W[0].p, host, false);
        return PS3D_OK;
    }
    if (field_id == PS3D_F_SBUOY || field_id == PS3D_F_BUOY || field_id == PS3D_F_SBUOYS) {
        if (!c.buoyancy) fail(PS3D_ERR_NOT_INITIALISED, "enable_buoyancy has not been called");
        if (field_id == PS3D_F_BUOY) {                 // field_combine_physical(sbuoy, buoy) (field_diagnostics_netcdf.f90:275)
            launch_zop(c, ZOP_COMBINE, c.sbuoy.p, c.bsem.p);
            fft2d_inv(c, c.bsem.p, c.buoy.p, false, false);
        }
        to_host(c, field_id == PS3D_F_SBUOY ? c.sbuoy.p : field_id == PS3D_F_BUOY ? c.buoy.p : c.sbuoys.p, host,
                field_id != PS3D_F_BUOY);
        return PS3D_OK;
    }
    bool spectral = false;
    DevBuf<double>* f = field_by_id(c, field_id, spectral);
    if (!f || comp < 0 || comp > 2) fail(PS3D_ERR_BAD_ARGUMENT, "bad field id %d / component %d", field_id, comp);
    // the time loop does not keep svorts when the update rides on the source kernel: re-evaluate the last tendency
    // from vel, vor (still those of the last source call) when somebody asks for it
    if (field_id == PS3D_F_SVORTS && c.svorts_stale) do_source(c);
    to_host(c, f[comp].p, host, spectral);
    PS_API_END
}

int ps3d_cuda_upload(int field_id, int comp, const double* host) {
    PS_API_BEGIN
    Ctx& c = ready();
    if ((field_id == PS3D_F_SBUOY || field_id == PS3D_F_SBUOYS) && host) {
        if (!c.buoyancy) fail(PS3D_ERR_NOT_INITIALISED, "enable_buoyancy has not been called");
        to_device(c, host, field_id == PS3D_F_SBUOY ? c.sbuoy.p : c.sbuoys.p, true);
        ps_sync(c.stream);
        return PS3D_OK;
    }
    bool spectral = false;
    DevBuf<double>* f = field_by_id(c, field_id, spectral);
    if (!f || comp < 0 || comp > 2 || !host) fail(PS3D_ERR_BAD_ARGUMENT, "bad field id %d / component %d", field_id, comp);
    to_device(c, host, f[comp].p, spectral);
    if (field_id == PS3D_F_SVEL) c.velx_valid = false;
    ps_sync(c.stream);
    PS_API_END
}

int ps3d_cuda_diagnostics(double out[8]) {
    PS_API_BEGIN
    Ctx& c = ready();
    if (!out) fail(PS3D_ERR_BAD_ARGUMENT, "out is null");
    field_reduce(c);
    ps_d2h(c.h_red, c.red.p, RQ_N * sizeof(double), c.stream);
    ps_sync(c.stream);
    allreduce_host(c, c.h_red, RQ_N, RQ_OPMASK);
    const double ncelli = 1.0 / (double)c.ncell;
    for (int i = 0; i < 8; ++i) out[i] = 0.0;
    out[0] = 0.5 * c.h_red[RQ_SUMU2] * ncelli;      // field_diagnostics.f90:93-103
    out[1] = 0.5 * c.h_red[RQ_SUMW2] * ncelli;      // :177-187
    out[2] = c.h_red[RQ_SUMUW] * ncelli;            // plotting/plot_vor_vel_he_evolution.py:62-65
    out[3] = 0.5 * c.h_red[RQ_SUMUH] * ncelli;      // get_horizontal_kinetic_energy (field_diagnostics.f90:128)
    out[4] = out[0] - out[3];                       // get_vertical_kinetic_energy (:153)
    out[5] = 0.5 * c.h_red[RQ_SUMWH] * ncelli;      // get_horizontal_enstrophy (:211)
    out[6] = out[1] - out[5];                       // get_vertical_enstrophy (:248)
    out[7] = std::sqrt(c.h_red[RQ_MAXWH]);          // get_max_horizontal_enstrophy (:233)
    PS_API_END
}

int ps3d_cuda_field_stats(double out[40]) {
    PS_API_BEGIN
    Ctx& c = ready();
    if (!out) fail(PS3D_ERR_BAD_ARGUMENT, "out is null");
    if (!c.have_diag) fail(PS3D_ERR_NOT_INITIALISED, "field_stats needs the diagnostics of adapt (advance.f90:188-193, 315-321)");
    for (int i = 0; i < PS3D_NC_COUNT; ++i) out[i] = 0.0;
    // summed values (field_diagnostics_netcdf.f90:292-330)
    field_reduce(c);
    ps_d2h(c.h_red, c.red.p, RQ_N * sizeof(double), c.stream);
    ps_sync(c.stream);
    double r1[RQ_N];
    for (int i = 0; i < RQ_N; ++i) r1[i] = c.h_red[i];
    allreduce_host(c, r1, RQ_N, RQ_OPMASK);
    const double ncelli = 1.0 / (double)c.ncell;
    out[PS3D_NC_KE] = 0.5 * r1[RQ_SUMU2] * ncelli;
    out[PS3D_NC_EN] = 0.5 * r1[RQ_SUMW2] * ncelli;
    out[PS3D_NC_KEXY] = 0.5 * r1[RQ_SUMUH] * ncelli;
    out[PS3D_NC_KEZ] = out[PS3D_NC_KE] - out[PS3D_NC_KEXY];
    out[PS3D_NC_ENXY] = 0.5 * r1[RQ_SUMWH] * ncelli;
    out[PS3D_NC_ENZ] = out[PS3D_NC_EN] - out[PS3D_NC_ENXY];
    out[PS3D_NC_HEMAX] = std::sqrt(r1[RQ_MAXWH]);
    // minima, maxima and the two surface rms values (:301-305, :352-413); delta = u_x + v_y -> W[0]
    do_delta(c);
    const long long ncol = (long long)c.nxl * c.ny;
    PS_LAUNCH((k_field_stats), dim3(c.red_blocks), dim3(RED_THREADS), RED_THREADS * sizeof(double), c.stream,
              field_ptrs(c), (const double*)c.W[0].p, ncol, c.nz, c.pz, c.partial.p);
    PS_LAUNCH((k_reduce_final), dim3(1), dim3(RED_THREADS), RED_THREADS * sizeof(double), c.stream,
              (const double*)c.partial.p, c.red_blocks, (int)SQ_N, SQ_OPMASK, c.red.p);
    c.launches += 2;
    ps_d2h(c.h_red, c.red.p, SQ_N * sizeof(double), c.stream);
    ps_sync(c.stream);
    double r2[SQ_N];
    for (int i = 0; i < SQ_N; ++i) r2[i] = c.h_red[i];
    allreduce_host(c, r2, SQ_N, SQ_OPMASK);
    out[PS3D_NC_OXMIN] = -r2[SQ_NMIN0]; out[PS3D_NC_OYMIN] = -r2[SQ_NMIN1]; out[PS3D_NC_OZMIN] = -r2[SQ_NMIN2];
    out[PS3D_NC_OXMAX] = r2[SQ_MAX0]; out[PS3D_NC_OYMAX] = r2[SQ_MAX1]; out[PS3D_NC_OZMAX] = r2[SQ_MAX2];
    out[PS3D_NC_USOXMAX] = r2[SQ_US0]; out[PS3D_NC_USOYMAX] = r2[SQ_US1]; out[PS3D_NC_USOZMAX] = r2[SQ_US2];
    out[PS3D_NC_LSOXMAX] = r2[SQ_LS0]; out[PS3D_NC_LSOYMAX] = r2[SQ_LS1]; out[PS3D_NC_LSOZMAX] = r2[SQ_LS2];
    out[PS3D_NC_USUHMAX] = std::sqrt(r2[SQ_USUH2]);
    const double nxy = (double)c.nx * (double)c.ny;
    out[PS3D_NC_USZRMS] = std::sqrt(r2[SQ_USZ2] / nxy);
    out[PS3D_NC_USDELRMS] = std::sqrt(r2[SQ_USDEL2] / nxy);
    const double fcor3 = c.f_cor[2];                 // physics.f90:163-169 (ps3d_cuda_set_physics; 0 by default)
    out[PS3D_NC_ROMIN] = out[PS3D_NC_OZMIN] / fcor3; // field_diagnostics.f90:297-298
    out[PS3D_NC_ROMAX] = out[PS3D_NC_OZMAX] / fcor3; // :313-314
    // handed over by adapt (advance.f90:188-193, 315-321, 366)
    const double* d = c.last_diag;
    out[PS3D_NC_OMAX] = d[PS3D_D_VORTMAX]; out[PS3D_NC_ORMS] = d[PS3D_D_VORTRMS]; out[PS3D_NC_OCHAR] = d[PS3D_D_VORCH];
    out[PS3D_NC_OXMEAN] = d[PS3D_D_VORMEAN_X]; out[PS3D_NC_OYMEAN] = d[PS3D_D_VORMEAN_Y];
    out[PS3D_NC_OZMEAN] = d[PS3D_D_VORMEAN_Z];
    out[PS3D_NC_GMAX] = d[PS3D_D_GGMAX]; out[PS3D_NC_BFMAX] = d[PS3D_D_BFMAX];
    out[PS3D_NC_UMAX] = d[PS3D_D_UMAX]; out[PS3D_NC_VMAX] = d[PS3D_D_VMAX]; out[PS3D_NC_WMAX] = d[PS3D_D_WMAX];
    out[PS3D_NC_USGMAX] = d[PS3D_D_USGGMAX]; out[PS3D_NC_LSGMAX] = d[PS3D_D_LSGGMAX];
    out[PS3D_NC_RGMAX] = d[PS3D_D_RMV];
    PS_API_END
}

int ps3d_cuda_genspec(int nmax, double* spec, double* num, int* nbins, double* dk_out) {
    PS_API_BEGIN
    Ctx& c = ready();
    if (!nbins || !dk_out) fail(PS3D_ERR_BAD_ARGUMENT, "nbins / dk is null");
    // kmax = maxval(nint(|k|)) (genspec.f90:74-84): |k| is largest at the Nyquist wavenumbers
    const double kx = c.h_rkx[c.nx / 2], ky = c.h_rky[c.ny / 2], kz = c.h_rkz[c.nz];
    const int kmax = (int)std::floor(std::sqrt(kx * kx + ky * ky + kz * kz) + 0.5);
    const double dk = (double)kmax / std::sqrt(0.25 * (double)c.nx * c.nx + 0.25 * (double)c.ny * c.ny + (double)c.nz * c.nz);   // :97
    // The reference allocates spec(0:kmax) (:86-87) but bins with m = int(kmag / dk) (:100), which exceeds kmax
    // whenever dk < 1, i.e. for any box larger than pi (2, 2, 1): a latent out-of-bounds write there.  Here the
    // bins cover the largest reachable index, and the kernel checks it.
    const int nb = std::max(kmax, (int)((double)kmax * (1.0 / dk))) + 1;
    *nbins = nb; *dk_out = dk;
    if (!spec) return PS3D_OK;
    if (!num || nmax < nb) fail(PS3D_ERR_BAD_ARGUMENT, "genspec needs room for %d bins", nb);
    // kinetic energy of the current state (:62)
    field_reduce(c);
    ps_d2h(c.h_red, c.red.p, RQ_N * sizeof(double), c.stream);
    ps_sync(c.stream);
    double r1[RQ_N];
    for (int i = 0; i < RQ_N; ++i) r1[i] = c.h_red[i];
    allreduce_host(c, r1, RQ_N, RQ_OPMASK);
    const double ke = 0.5 * r1[RQ_SUMU2] / (double)c.ncell;
    // fully spectral velocity (:64-73) -> W[0..2]
    for (int i = 0; i < 3; ++i) {
        fft2d_fwd(c, c.vel[i].p, c.W[i].p);
        launch_zop(c, i < 2 ? ZOP_COSINE : ZOP_SINE, c.W[i].p, c.W[i].p);
    }
    DevBuf<double> bins;
    bins.alloc((size_t)2 * nb);
    ps_memset(bins.p, 0, (size_t)2 * nb * sizeof(double), c.stream);
    PS_LAUNCH((k_spec_bin), dim3(stream_blocks((size_t)c.nx * c.nyl * c.pz)), dim3(256), 0, c.stream,
              (const double*)c.W[0].p, (const double*)c.W[1].p, (const double*)c.W[2].p, (const double*)c.k2l2.p,
              (const double*)c.rkz.p, c.nx, c.nyl, c.nz, c.pz, 1.0 / dk, nb, bins.p, bins.p + nb);
    ++c.launches;
    std::vector<double> h((size_t)2 * nb);
    ps_d2h(h.data(), bins.p, h.size() * sizeof(double), c.stream);
    ps_sync(c.stream);
    bins.release();
    for (size_t o = 0; o < h.size(); o += 32) allreduce_host(c, h.data() + o, (int)std::min<size_t>(32, h.size() - o), 0u);   // :108-109
    const double pi = std::acos(-1.0);
    const double prefactor = 4.0 / 3.0 * pi * dk * dk * dk;                                     // :111
    double total = 0.0;
    for (int m = 0; m < nb; ++m) {
        const double cnt = h[(size_t)nb + m];
        num[m] = cnt;
        const double m0 = (double)m, m1 = (double)(m + 1);
        spec[m] = (cnt > 0.0) ? h[m] * prefactor * (m1 * m1 * m1 - m0 * m0 * m0) / cnt : h[m];  // :114-120
        total += spec[m] * dk;
    }
    const double snorm = ke / total;                                                            // :123
    for (int m = 0; m < nb; ++m) spec[m] *= snorm;
    PS_API_END
}

int ps3d_cuda_set_transport(ps3d_alltoall_fn alltoall, ps3d_allreduce_fn allreduce, void* user) {
    PS_API_BEGIN
    Ctx& c = ctx();
    c.tr.a2a_cb = alltoall; c.tr.ar_cb = allreduce; c.tr.user = user;
    PS_API_END
}

int ps3d_cuda_comm_stats(long long* n_alltoall, double* bytes_sent) {
    PS_API_BEGIN
    Ctx& c = ctx();
    if (n_alltoall) *n_alltoall = c.tr.n_alltoall;
    if (bytes_sent) *bytes_sent = c.tr.bytes_sent;
    PS_API_END
}

long long ps3d_cuda_kernel_launches(void) { return g_ctx ? g_ctx->launches : 0; }
long long ps3d_cuda_tma_launches(void) {
#ifndef PS3D_EMU
    return g_ctx ? g_ctx->tma_launches : 0;
#else
    return 0;
#endif
}
double ps3d_cuda_last_advance_ms(void) { return g_ctx ? g_ctx->last_advance_ms : 0.0; }

int ps3d_cuda_time_kernel(int which, int reps, double* ms_per_launch) {
    PS_API_BEGIN
    Ctx& c = ready();
    if (!ms_per_launch || reps < 1 || which < 0 || which > 9) fail(PS3D_ERR_BAD_ARGUMENT, "bad time_kernel arguments");
#ifndef PS3D_EMU
    if (which >= 7) {
        if (!(c.nranks > 1 && c.tr.p2p)) fail(PS3D_ERR_BAD_ARGUMENT, "time_kernel 7 (peer-memory scatter sweep) needs nranks > 1 with peer access");
        cross_rank_barrier(c, c.stream);          // every rank is here: the receive buffers are free
    }
#endif
#ifndef PS3D_EMU
    PS_CUDA_TRY(cudaEventRecord(c.ev0, c.stream));
#endif
    for (int r = 0; r < reps; ++r) {
        switch (which) {
            case 0: { Sweep s{1, false, PRO_PLAIN, {c.vor[0].p, nullptr, nullptr, nullptr}, 0, 0, c.W[3].p}; run_sweep(c, s); break; }
            case 1: { Sweep s{0, false, PRO_PLAIN, {c.W[3].p, nullptr, nullptr, nullptr}, 0, 0, c.W[4].p}; run_sweep(c, s); break; }
            case 2: { Sweep s{0, true, PRO_PLAIN, {c.svel[0].p, nullptr, nullptr, nullptr}, 0, 0, c.W[3].p}; run_sweep(c, s); break; }
            case 3: { Sweep s{1, true, PRO_PLAIN, {c.W[3].p, nullptr, nullptr, nullptr}, 0, 0, c.W[4].p}; run_sweep(c, s); break; }
            case 4: {
                V2VArgs a;
                // writes go to scratch so that the resident state is not disturbed
                a.svor0 = c.W[0].p; a.svor1 = c.W[1].p; a.svor2 = c.svor[2].p;
                a.wsem0 = c.W[2].p; a.wsem1 = c.W[3].p; a.wsem2 = c.W[4].p;
                a.svel0 = c.W[5].p; a.svel1 = c.W[5].p; a.svel2 = c.W[5].p;
                if (r == 0) { ps_d2d(c.W[0].p, c.svor[0].p, c.nint * sizeof(double), c.stream); ps_d2d(c.W[1].p, c.svor[1].p, c.nint * sizeof(double), c.stream); }
                launch_v2v(c, a);
                break;
            }
            case 6: {    // forward y sweep with the u x omega product in the load (inversion.f90:327)
                Sweep s{1, false, PRO_CROSS, {c.vel[0].p, c.vor[1].p, c.vel[1].p, c.vor[0].p}, 0, 0, c.W[3].p};
                run_sweep(c, s);
                break;
            }
            case 7: {    // forward y sweep that stores straight into the peers' receive buffers (the fused sweep + all-to-all)
                Sweep s{1, false, PRO_PLAIN, {c.vor[0].p, nullptr, nullptr, nullptr}, 0, 0, c.W[6].p};
                s.scatter = 0;
                s.max_ctas = (c.p2p_ctas_per_sm >= 0 ? c.p2p_ctas_per_sm : 2) * c.num_sms;
                run_sweep(c, s);
                break;
            }
#ifndef PS3D_EMU
            case 8: case 9: {    // NVLink ceiling of SM-issued stores: 8 = the whole field to the next rank, contiguous;
                                 // 9 = the exchange pattern (block d to rank d, this rank's block stays local), contiguous blocks
                PeerCopyDst dst;
                const long long nb = (long long)c.nxl * c.nyl * c.pz;
                if (which == 8) {
                    for (int d = 0; d < 8; ++d) dst.p[d] = c.tr.peer_t2[1][(c.rank + 1) % c.nranks];
                    PS_LAUNCH((k_peer_copy), dim3(c.num_sms * 8), dim3(256), 0, c.stream, (const double*)c.vor[0].p, dst, (long long)c.nint, 0LL, 1);
                } else {
                    for (int d = 0; d < 8; ++d) dst.p[d] = c.tr.peer_t2[1][d % c.nranks];
                    PS_LAUNCH((k_peer_copy), dim3(c.num_sms * 8), dim3(256), 0, c.stream, (const double*)c.vor[0].p, dst, nb, (long long)c.rank * nb, c.nranks);
                }
                ++c.launches;
                break;
            }
#endif
            case 5: {
                SrcArgs a;
                a.r = c.W[0].p; a.q = c.W[1].p; a.p = c.W[2].p;
                a.s0 = c.W[3].p; a.s1 = c.W[4].p; a.s2 = c.W[5].p;
                a.upd = -1; a.c1 = 0.0;
                for (int i = 0; i < 3; ++i) { a.svor[i] = nullptr; a.vortsm[i] = nullptr; }
                a.f2d = a.filtz = a.vd = nullptr;
                launch_src(c, a);
                break;
            }
        }
    }
#ifndef PS3D_EMU
    PS_CUDA_TRY(cudaEventRecord(c.ev1, c.stream));
    PS_CUDA_TRY(cudaEventSynchronize(c.ev1));
    float ms = 0.f;
    PS_CUDA_TRY(cudaEventElapsedTime(&ms, c.ev0, c.ev1));
    *ms_per_launch = (double)ms / reps;
    if (which >= 7) { cross_rank_barrier(c, c.stream); ps_sync(c.stream); }
#else
    ps_sync(c.stream);
    *ms_per_launch = 0.0;
#endif
    PS_API_END
}

}  // extern "C"
