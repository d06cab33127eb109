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
    double* peer_t2[2][8] = {};       // [buffer][rank]: base of that rank's receive buffer
    void* ipc_opened[2][8] = {};
    long long n_alltoall = 0;
    double bytes_sent = 0.0;

    bool have_nccl() const { return comm != nullptr; }
};

}  // namespace ps3d
