// Slab exchange transport (replaces reference src/fft/fft_pencil.f90:283-330 `transpose_to_pencil`
// MPI_Alltoallv and the MPI_Allreduce calls of advance.f90:299-305 / field_diagnostics.f90:418-535).
//
// One all-to-all of P equal, contiguous blocks per 2-D FFT (pack/unpack are folded into the row
// addressing of the y sweeps, see RowMap in line_fft.cuh), plus small all-reduces of scalars.
// Two back ends:
//   * NCCL over NVLink/NVSwitch (grouped ncclSend/ncclRecv on the library's stream), resolved with
//     dlopen at run time so that a single-GPU build/run needs no NCCL at all and a process that already
//     carries an NCCL (e.g. through PyTorch) shares it;
//   * caller-supplied callbacks (ps3d_cuda_set_transport): lets the Fortran host plug its CUDA-aware
//     MPI_Alltoall / MPI_Allreduce (mpi_layout.f90 communicators), and lets the CPU block emulator run
//     multi-rank under torch.distributed/gloo in the tests.
#pragma once

#include <dlfcn.h>
#include "rt.h"

namespace ps3d {

typedef int (*alltoall_fn)(const void* send, void* recv, size_t bytes_per_rank, void* user);
typedef int (*allreduce_fn)(double* buf, int n, int op, void* user);   // op: 0 = sum, 1 = max; host buffer

struct NcclApi {
    typedef struct { char internal[128]; } UniqueId;
    typedef void* Comm;
    int (*CommInitRank)(Comm*, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(Comm) = nullptr;
    int (*Send)(const void*, size_t, int, int, Comm, void*) = nullptr;
    int (*Recv)(void*, size_t, int, int, Comm, void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, Comm, void*) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, Comm, void*) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    void* handle = nullptr;
    enum { kUint8 = 1, kFloat64 = 8, kSum = 0, kMax = 2 };

    bool load() {
        if (handle) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) { set_error("cannot dlopen libnccl.so.2: %s", dlerror()); return false; }
#define PS_NCCL_SYM(field, name)                                            \
        *(void**)(&field) = dlsym(handle, name);                            \
        if (!field) { set_error("NCCL symbol %s not found", name); return false; }
        PS_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        PS_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        PS_NCCL_SYM(Send, "ncclSend")
        PS_NCCL_SYM(Recv, "ncclRecv")
        PS_NCCL_SYM(AllReduce, "ncclAllReduce")
        PS_NCCL_SYM(AllGather, "ncclAllGather")
        PS_NCCL_SYM(GroupStart, "ncclGroupStart")
        PS_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        PS_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef PS_NCCL_SYM
        return true;
    }
};

struct Transport {
    int rank = 0, nranks = 1;
    NcclApi nccl;
    NcclApi::Comm comm = nullptr;
    alltoall_fn a2a_cb = nullptr;
    allreduce_fn ar_cb = nullptr;
    void* user = nullptr;
    bool p2p = false;                 // sweeps store straight into the peers' receive buffers (CUDA IPC)
    // [buffer][rank]: base of that rank's receive buffer.  Buffers 0, 1, 5, 6: the rotating receive fields of the 2-D
    // FFT pipeline (rot_bufs of them in use); 2..4: the kept x-transformed velocity of the last vor2vel (Ctx::velx)
    static constexpr int NPEERBUF = 7;
    int rot_bufs = 4;                           // PS3D_ROT_BUFS=2: two rotating receive buffers (r01 / r02k behaviour)
    static int rot_index(int j) { const int t[4] = {0, 1, 5, 6}; return t[j & 3]; }
    static bool is_kept(int b) { return b >= 2 && b <= 4; }
    double* peer_t2[NPEERBUF][8] = {};
    void* ipc_opened[NPEERBUF][8] = {};
    long long n_alltoall = 0;
    double bytes_sent = 0.0;
    // peer-memory mailbox (PeerMail below) behind the first receive buffer of every rank: epoch-flag barrier and
    // small all-reduce without a collective launch
    int peer_sync = 1;                          // PS3D_NO_PEER_SYNC=1: one-element NCCL all-reduce / NCCL all-reduces instead
    unsigned long long bar_epoch = 0, ar_epoch = 0;
    // split-phase exchange protocol (fft2d_batch): exchanges done so far on each peer receive buffer
    unsigned long long use_count[NPEERBUF] = {};
    int split_phase = 0;                        // PS3D_SPLIT_PHASE=1: signal / wait pairs instead of one full barrier per 2-D FFT

    bool have_nccl() const { return comm != nullptr; }
};

// Mailbox in every rank's peer-mapped memory.  Flags only ever grow (epochs), so a late reader never sees a stale "go".
struct PeerMail {
    unsigned long long bar_flag[8];             // [src rank]: epoch of src's last barrier arrival
    unsigned long long ar_flag[2][8];           // [parity][src rank]: epoch of src's last all-reduce contribution
    double vals[2][8][32];                      // [parity][src rank][value]
    // split-phase exchange: [buffer][src rank] = how many exchanges on receive buffer b rank src has
    //   arr : finished scattering into MY buffer b (its first sweep is complete and visible here)
    //   done: finished reading out of ITS OWN buffer b (its second sweep is complete: b may be overwritten there)
    unsigned long long arr_flag[7][8];
    unsigned long long done_flag[7][8];
};
struct PeerMailPtrs { PeerMail* m[8]; };

#ifndef PS3D_EMU
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// spin until *p >= e; gives up after ~4 s (a peer died): the caller's results are then garbage, but the GPU is not hung
__device__ __forceinline__ void spin_until(const unsigned long long* p, unsigned long long e) {
    const long long t0 = clock64();
    while (ld_acquire_sys(p) < e) {
        if (clock64() - t0 > 8000000000LL) break;
        __nanosleep(64);
    }
}

// Cross-rank barrier: all ranks' preceding work on this stream is complete (and its peer stores visible) before any
// rank's following work starts.  One block; thread p talks to rank p.
__global__ void k_peer_barrier(PeerMailPtrs mp, int rank, int nranks, unsigned long long epoch) {
    const int p = threadIdx.x;
    if (p < nranks) {
        __threadfence_system();
        st_release_sys(&mp.m[p]->bar_flag[rank], epoch);
        spin_until(&mp.m[rank]->bar_flag[p], epoch);
    }
}

// Split-phase form of the barrier between the two sweeps of a 2-D FFT.  `which` 0 = arr, 1 = done (PeerMail).
// k_peer_signal: this rank's preceding work on the stream is complete and visible -> every rank's flag [b][rank] = epoch
// (one thread per destination).  k_peer_wait: spin until every rank's flag [b][p] in MY mailbox has reached epoch.
// The NVLink-bound first sweeps (comm stream) signal and move on; only the second sweep (compute stream) waits.
__global__ void k_peer_signal(PeerMailPtrs mp, int rank, int nranks, int which, int b, unsigned long long epoch) {
    const int p = threadIdx.x;
    if (p < nranks) {
        __threadfence_system();
        PeerMail* m = mp.m[p];
        st_release_sys(which ? &m->done_flag[b][rank] : &m->arr_flag[b][rank], epoch);
    }
}
__global__ void k_peer_wait(PeerMailPtrs mp, int rank, int nranks, int which, int b, unsigned long long epoch) {
    const int p = threadIdx.x;
    if (p < nranks) {
        const PeerMail* m = mp.m[rank];
        spin_until(which ? &m->done_flag[b][p] : &m->arr_flag[b][p], epoch);
    }
    __threadfence_system();
}

// Plain streaming copy into peer memory (16-byte stores): the ceiling of SM-issued stores over NVLink, measured by
// ps3d_cuda_time_kernel(8 / 9) beside the scatter sweeps.  Block d of `src` (nb doubles) goes to dst[d] + off.
struct PeerCopyDst { double* p[8]; };
__global__ void k_peer_copy(const double* __restrict__ src, PeerCopyDst dst, long long nb, long long off, int nblk) {
    const long long n2 = nb / 2;
    for (int d = 0; d < nblk; ++d) {
        const double2* s2 = reinterpret_cast<const double2*>(src + (long long)d * nb);
        double2* d2 = reinterpret_cast<double2*>(dst.p[d] + off);
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x)
            d2[i] = s2[i];
    }
}

// All-reduce of red[0..n) over the ranks, sums or maxima per bit of opmask: every rank stores its vector into
// every rank's mailbox, then reduces the P vectors in rank order -- the same order on every rank, so the result
// is bitwise identical everywhere and independent of timing.
__global__ void k_peer_allreduce(PeerMailPtrs mp, int rank, int nranks, unsigned long long epoch, int n, unsigned opmask,
                                 double* red) {
    const int t = threadIdx.x, par = (int)(epoch & 1);
    if (t < n) {
        const double v = red[t];
        for (int p = 0; p < nranks; ++p) mp.m[p]->vals[par][rank][t] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (t < nranks) {
        st_release_sys(&mp.m[t]->ar_flag[par][rank], epoch);
        spin_until(&mp.m[rank]->ar_flag[par][t], epoch);
    }
    __syncthreads();
    if (t < n) {
        const PeerMail* me = mp.m[rank];
        const bool mx = (opmask >> t) & 1;
        double acc = *((volatile const double*)&me->vals[par][0][t]);
        for (int p = 1; p < nranks; ++p) {
            const double v = *((volatile const double*)&me->vals[par][p][t]);
            acc = mx ? fmax(acc, v) : acc + v;
        }
        red[t] = acc;
    }
}
#endif

}  // namespace ps3d
