// Peer-memory all-gather of per-frame rows (score records, debug Euler sequences) over NVLink / NVSwitch.
//
// What it replaces: the reference keeps every frame's results in Python lists of ONE process
// (lib/core/base.py:144-151,168: `reba_results`, `rula_results`, `pose_logs`); with frames sharded over the GPUs of a
// box those lists become one exchange of 32-byte records (+ 24 B per debug joint) per frame.
//
// How: every rank owns a gather buffer (two slots, so that a slot is still readable while the next exchange is
// in flight) plus one 64-bit flag per slot and source rank.  The peers map that allocation through CUDA IPC -- or use
// the pointer as it is when they live in the same process -- and ONE kernel per exchange
//   1. copies this rank's rows into the current slot of EVERY rank (plain 16-byte stores to peer pointers),
//   2. fences at system scope; the last block to finish raises flag[slot][rank] = epoch on every rank
//      (st.release.sys), and
//   3. a one-warp kernel on the communicator's OWN stream then spins (ld.acquire.sys, time-bounded) until all `world`
//      flags of the local slot carry the epoch, i.e. every peer's rows have landed here (round 2c: the spin used to
//      sit in the push kernel, i.e. on the caller's stream -- inside prk_pipeline_host that is the scoring stream, and
//      the next call's scoring queued behind a wait for the slowest of 8 ranks: 202 instead of 150 us per step).
// prk_allgather_rows joins the caller's stream with that wait, so work queued behind it sees the complete gathered array;
// the pipeline entry points do not (prk_comm_wait joins where the rows are read): the next exchange of the communicator
// is the only thing that waits for the peers, which gives the ranks a full call of slack.  No NCCL kernel, no
// host round trip, a few microseconds per exchange, underneath the vertex kernel.
//
// Slot reuse: rank A writes slot e&1 of rank B during A's exchange e, which starts after A has seen B's flag of
// exchange e-1, i.e. after B's stream reached its exchange e-1.  B's readers of exchange e-2 (same slot) are safe if they
// were queued before B ISSUED its exchange e-1 (the exchange is ordered behind the work already on its stream) -- hence
// "the gathered rows of a call stay valid until this rank issues its next exchange on the communicator".
#include "prk_internal.h"

#include <cstdio>
#include <cstring>
#include <new>
#include <unistd.h>

namespace prk {
void comm_set_detail(const char* what, const char* msg);
}

constexpr int kMaxRanks = 8;
constexpr uint64_t kHandleMagic = 0x70726B636F6D6D32ull;   // "prkcomm2"

struct PeerPtrs { uint8_t* buf[kMaxRanks]; };

struct CommHandleBlob {
    uint64_t magic;
    int32_t pid;
    int32_t device;
    uint64_t ptr;
    uint64_t bytes;
    cudaIpcMemHandle_t ipc;
};

struct prk_comm {
    int rank = 0, world = 1, device = -1;
    size_t slot_bytes = 0, flags_off = 0, total_bytes = 0;
    uint8_t* d_buf = nullptr;          // [2][slot_bytes] | flags [2][kMaxRanks] u64 | done counter
    PeerPtrs peers = {};
    bool ipc_opened[kMaxRanks] = {};
    bool peers_open = false;
    uint64_t epoch = 0;
    int* h_status = nullptr;           // host-mapped: set to 1 by a wait that timed out
    int* d_status = nullptr;           // device alias of h_status
    cudaEvent_t ev_last = nullptr;     // end of the previous exchange's wait (exchanges of one comm are serialised)
    cudaEvent_t ev_push = nullptr;     // the previous exchange's rows and flags are out
    cudaStream_t s_wait = nullptr;     // the waits for the peers run here, not on the caller's stream
    bool ev_valid = false;
};

namespace {

__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// src: n16 units (16 or 8 bytes each) of this rank's rows; they belong at unit dst16 of the slot that starts slot_off
// bytes into every rank's buffer.  flags: [2][kMaxRanks] u64 at flags_off of every buffer.
template <typename U>
__global__ void __launch_bounds__(256)
allgather_push_kernel(const U* __restrict__ src, int64_t n16, int64_t dst16, PeerPtrs peers, int world, int rank,
                      size_t slot_off, size_t flags_off, int slot, uint64_t epoch, unsigned int* done) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // four loads in flight per thread (one at a time: 1.6 - 2.0 TB/s for the rank-local copy; the exchange of 128 MB of
    // records + Euler rows was 16 % of a 1M-frame joints-only job)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += 4 * stride) {
        U v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) if (i + k * stride < n16) v[k] = src[i + k * stride];
        for (int p = 0; p < world; ++p) {
            U* dst = reinterpret_cast<U*>(peers.buf[p] + slot_off) + dst16;
#pragma unroll
            for (int k = 0; k < 4; ++k) if (i + k * stride < n16) dst[i + k * stride] = v[k];
        }
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool s_last;
    if (threadIdx.x == 0) s_last = atomicAdd(done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    // last block: every block's rows are out (their system-scope fences precede the counter increments read here)
    if (threadIdx.x == 0) *done = 0;                      // for the next exchange (stream-ordered behind this one)
    __threadfence_system();
    if ((int)threadIdx.x < world)
        st_release_sys(reinterpret_cast<uint64_t*>(peers.buf[threadIdx.x] + flags_off) + slot * kMaxRanks + rank, epoch);
}

// one warp: lane p waits for rank p's flag of this slot
__global__ void __launch_bounds__(32)
allgather_wait_kernel(const uint8_t* own_buf, size_t flags_off, int world, int slot, uint64_t epoch, int* status,
                      uint64_t timeout_ns) {
    if ((int)threadIdx.x < world) {
        const uint64_t* f = reinterpret_cast<const uint64_t*>(own_buf + flags_off) + slot * kMaxRanks + threadIdx.x;
        const uint64_t t0 = global_ns();
        while (ld_acquire_sys(f) < epoch) {
            if (global_ns() - t0 > timeout_ns) { *status = 1; __threadfence_system(); break; }
            __nanosleep(200);
        }
    }
}

int fail(const char* what, cudaError_t e) {
    prk::comm_set_detail(what, cudaGetErrorString(e));
    return PRK_ERR_CUDA;
}
#define COMM_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return fail(#expr, _e); } while (0)

}  // namespace

extern "C" {

size_t prk_comm_handle_bytes(void) { return sizeof(CommHandleBlob); }

int prk_comm_create(prk_comm** out, int rank, int world, int device, size_t slot_bytes) {
    if (!out || world < 1 || world > kMaxRanks || rank < 0 || rank >= world || slot_bytes == 0) {
        prk::comm_set_detail("prk_comm_create", "invalid argument (1 <= world <= 8)");
        return PRK_ERR_INVALID_ARG;
    }
    *out = nullptr;
    COMM_CUDA(cudaSetDevice(device));
    prk_comm* c = new (std::nothrow) prk_comm();
    if (!c) return PRK_ERR_INVALID_ARG;
    c->rank = rank; c->world = world; c->device = device;
    c->slot_bytes = (slot_bytes + 1023) & ~(size_t)1023;
    c->flags_off = 2 * c->slot_bytes;
    c->total_bytes = c->flags_off + 1024;
    cudaError_t e = cudaMalloc(&c->d_buf, c->total_bytes);
    if (e == cudaSuccess) e = cudaMemset(c->d_buf + c->flags_off, 0, 1024);
    if (e == cudaSuccess) e = cudaHostAlloc(&c->h_status, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable);
    if (e == cudaSuccess) { *c->h_status = 0; e = cudaHostGetDevicePointer(&c->d_status, c->h_status, 0); }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_last, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_push, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_wait, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { prk_comm_destroy(c); return fail("prk_comm_create", e); }
    c->peers.buf[rank] = c->d_buf;
    if (world == 1) c->peers_open = true;
    *out = c;
    return PRK_OK;
}

void prk_comm_destroy(prk_comm* c) {
    if (!c) return;
    if (c->device >= 0) cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int p = 0; p < kMaxRanks; ++p)
        if (c->ipc_opened[p]) cudaIpcCloseMemHandle(c->peers.buf[p]);
    if (c->ev_last) cudaEventDestroy(c->ev_last);
    if (c->ev_push) cudaEventDestroy(c->ev_push);
    if (c->s_wait) cudaStreamDestroy(c->s_wait);
    if (c->h_status) cudaFreeHost(c->h_status);
    cudaFree(c->d_buf);
    delete c;
}

int prk_comm_get_handle(prk_comm* c, void* h_out) {
    if (!c || !h_out) { prk::comm_set_detail("prk_comm_get_handle", "null argument"); return PRK_ERR_INVALID_ARG; }
    CommHandleBlob b;
    memset(&b, 0, sizeof b);
    b.magic = kHandleMagic;
    b.pid = (int32_t)getpid();
    b.device = c->device;
    b.ptr = reinterpret_cast<uint64_t>(c->d_buf);
    b.bytes = c->total_bytes;
    COMM_CUDA(cudaSetDevice(c->device));
    // a failure here (IPC not permitted in this container) is reported when a peer in ANOTHER process needs it
    if (cudaIpcGetMemHandle(&b.ipc, c->d_buf) != cudaSuccess) { cudaGetLastError(); memset(&b.ipc, 0, sizeof b.ipc); }
    memcpy(h_out, &b, sizeof b);
    return PRK_OK;
}

int prk_comm_open_peers(prk_comm* c, const void* h_all) {
    if (!c || !h_all) { prk::comm_set_detail("prk_comm_open_peers", "null argument"); return PRK_ERR_INVALID_ARG; }
    COMM_CUDA(cudaSetDevice(c->device));
    const CommHandleBlob* all = static_cast<const CommHandleBlob*>(h_all);
    for (int p = 0; p < c->world; ++p) {
        if (p == c->rank) continue;
        CommHandleBlob b;
        memcpy(&b, &all[p], sizeof b);
        if (b.magic != kHandleMagic || b.bytes != c->total_bytes) {
            prk::comm_set_detail("prk_comm_open_peers", "handle of a peer is malformed or was made with another slot size");
            return PRK_ERR_PEER;
        }
        if (b.device != c->device) {
            int can = 0;
            COMM_CUDA(cudaDeviceCanAccessPeer(&can, c->device, b.device));
            if (!can) { prk::comm_set_detail("prk_comm_open_peers", "no peer access between the two devices"); return PRK_ERR_PEER; }
        }
        if (b.pid == (int32_t)getpid()) {          // same process: the pointer is valid here as it is
            if (b.device != c->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail("cudaDeviceEnablePeerAccess", e);
                cudaGetLastError();
            }
            c->peers.buf[p] = reinterpret_cast<uint8_t*>(b.ptr);
        } else {
            void* ptr = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&ptr, b.ipc, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                cudaGetLastError();
                prk::comm_set_detail("cudaIpcOpenMemHandle", cudaGetErrorString(e));
                return PRK_ERR_PEER;
            }
            c->peers.buf[p] = static_cast<uint8_t*>(ptr);
            c->ipc_opened[p] = true;
        }
    }
    c->peers_open = true;
    return PRK_OK;
}

// join: make `stream` wait for the peers' rows too (the public entry point); the pipeline calls pass false and leave that to
// prk_comm_wait / the communicator's next exchange
int prk_allgather_rows_impl(prk_comm* c, const void* d_local, int64_t n_local, int64_t row_offset, int64_t row_bytes,
                            void** d_gathered_out, void* stream, bool join) {
    if (!c || n_local < 0 || row_offset < 0 || row_bytes <= 0 || (row_bytes & 7) || (n_local > 0 && !d_local) ||
        (reinterpret_cast<uintptr_t>(d_local) & 7)) {
        prk::comm_set_detail("prk_allgather_rows", "invalid argument (rows must be 8-byte multiples, 8-byte aligned)");
        return PRK_ERR_INVALID_ARG;
    }
    if (!c->peers_open) { prk::comm_set_detail("prk_allgather_rows", "prk_comm_open_peers has not been called"); return PRK_ERR_PEER; }
    if ((size_t)(row_offset + n_local) * (size_t)row_bytes > c->slot_bytes) {
        prk::comm_set_detail("prk_allgather_rows", "rows do not fit in the gather slot");
        return PRK_ERR_WORKSPACE;
    }
    if (*c->h_status != 0) { prk::comm_set_detail("prk_allgather_rows", "an earlier exchange timed out"); return PRK_ERR_PEER; }
    COMM_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const uint64_t epoch = ++c->epoch;
    const int slot = (int)(epoch & 1);
    const size_t slot_off = (size_t)slot * c->slot_bytes;
    if (c->ev_valid) COMM_CUDA(cudaStreamWaitEvent(s, c->ev_last, 0));     // exchanges of one comm run one after the other
    const int64_t nbytes = n_local * row_bytes, obytes = row_offset * row_bytes;
    const bool wide = !((nbytes | obytes) & 15) && !(reinterpret_cast<uintptr_t>(d_local) & 15);   // 16-byte units
    const int unit = wide ? 16 : 8;
    const int64_t n16 = nbytes / unit, dst16 = obytes / unit;
    int grid = (int)((n16 + 256 * 4 - 1) / (256 * 4));
    if (grid < 1) grid = 1;
    if (grid > 148) grid = 148;
    unsigned int* done = reinterpret_cast<unsigned int*>(c->d_buf + c->flags_off + 2 * kMaxRanks * sizeof(uint64_t));
    const uint64_t timeout_ns = 20ull * 1000 * 1000 * 1000;
    if (wide)
        allgather_push_kernel<uint4><<<grid, 256, 0, s>>>(static_cast<const uint4*>(d_local), n16, dst16, c->peers, c->world,
                                                          c->rank, slot_off, c->flags_off, slot, epoch, done);
    else
        allgather_push_kernel<uint2><<<grid, 256, 0, s>>>(static_cast<const uint2*>(d_local), n16, dst16, c->peers, c->world,
                                                          c->rank, slot_off, c->flags_off, slot, epoch, done);
    prk::count_launch();
    COMM_CUDA(cudaGetLastError());
    COMM_CUDA(cudaEventRecord(c->ev_push, s));
    COMM_CUDA(cudaStreamWaitEvent(c->s_wait, c->ev_push, 0));
    allgather_wait_kernel<<<1, 32, 0, c->s_wait>>>(c->d_buf, c->flags_off, c->world, slot, epoch, c->d_status, timeout_ns);
    prk::count_launch();
    COMM_CUDA(cudaGetLastError());
    COMM_CUDA(cudaEventRecord(c->ev_last, c->s_wait));
    c->ev_valid = true;
    if (join) COMM_CUDA(cudaStreamWaitEvent(s, c->ev_last, 0));
    if (d_gathered_out) *d_gathered_out = c->d_buf + slot_off;
    return PRK_OK;
}

int prk_allgather_rows(prk_comm* c, const void* d_local, int64_t n_local, int64_t row_offset, int64_t row_bytes,
                       void** d_gathered_out, void* stream) {
    return prk_allgather_rows_impl(c, d_local, n_local, row_offset, row_bytes, d_gathered_out, stream, true);
}

int prk_allgather_scores(prk_comm* c, const prk_score_rec* d_local, int64_t n_local, int64_t frame_offset,
                         prk_score_rec** d_gathered_out, void* stream) {
    void* g = nullptr;
    const int rc = prk_allgather_rows(c, d_local, n_local, frame_offset, (int64_t)sizeof(prk_score_rec), &g, stream);
    if (rc == PRK_OK && d_gathered_out) *d_gathered_out = static_cast<prk_score_rec*>(g);
    return rc;
}

void* prk_comm_gathered(prk_comm* c) {
    if (!c || c->epoch == 0) return nullptr;
    return c->d_buf + (size_t)(c->epoch & 1) * c->slot_bytes;
}

int prk_comm_wait(prk_comm* c, void* stream) {
    if (!c) { prk::comm_set_detail("prk_comm_wait", "null argument"); return PRK_ERR_INVALID_ARG; }
    if (c->ev_valid) COMM_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), c->ev_last, 0));
    return PRK_OK;
}

int prk_comm_status(prk_comm* c) {
    if (!c) return PRK_ERR_INVALID_ARG;
    if (*c->h_status != 0) { prk::comm_set_detail("prk_comm_status", "a peer did not arrive within the time limit"); return PRK_ERR_PEER; }
    return PRK_OK;
}

}  // extern "C"
