// gather.cu -- the final gather of the ray buffer over NVLink (BASELINE north_star: "only a final NCCL gather of the ray
// buffer over NVLink"): gather-to-CONSUMER, tile by tile, overlapped with generation.
//
// One process per GPU.  Every rank contributes up to `tile` 32-byte ray records per ROUND; the consumer rank owns
// `slots` round buffers of world x tile records, and a consumer kernel (job.cu: checksum + counts, the renderer's
// stand-in) eats a round as soon as every rank's records are in.  Ingest is 1x the job (not (G-1)x as with an
// all-gather to everyone).  Three transports carry a producer's records into the consumer's buffer:
//
//   FUSED  the generate kernels store their records STRAIGHT INTO THE CONSUMER'S MEMORY: the round buffer is mapped into
//          every producer through CUDA IPC, and zoicb_generate's `d_rays` simply is a peer pointer, so each finished ray
//          leaves the SM as one 32-byte NVLink write -- compute and transfer are one kernel, nothing is staged.
//   PUSH   records are generated into a local staging tile and pushed by a copy engine (cudaMemcpyAsync on the mapped
//          peer pointer) on a second stream, while the next tile is generated.
//   NCCL   the same with ncclSend / a group of ncclRecv on the consumer (libnccl is dlopen'ed: the one torch already
//          loaded in-process, or the system's; no link-time dependency).
//
// FUSED and PUSH synchronise with 64-bit flags in IPC-mapped device memory, written and polled by one-thread kernels in
// stream order: arrive[r] (in the consumer's memory, written by rank r) = number of rounds rank r has delivered;
// freed (in every producer's memory, written by the consumer) = number of rounds the consumer has eaten.  A round k may
// be written into slot k % slots once freed >= k - slots + 1.  Rounds are numbered across jobs (every rank runs the same
// jobs, so every rank knows the number without talking), hence the flags only ever grow and are never cleared.  Every
// wait has a wall-clock bound and raises an error flag instead of hanging the GPU.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "capi_internal.h"
#include "gather.h"

using namespace zoicb;

namespace {

// ---- minimal NCCL surface, resolved at run time -----------------------------------------------------------------
typedef struct { char internal[128]; } NcclUniqueId;
typedef void* NcclComm;
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
NcclApi& nccl() {
    static NcclApi api = [] {
        NcclApi a;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { a.lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD); if (a.lib) break; }   // torch's copy, if loaded
        if (!a.lib) for (const char* n : names) { a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (a.lib) break; }
        if (!a.lib) return a;
        a.GetUniqueId = (int (*)(NcclUniqueId*))dlsym(a.lib, "ncclGetUniqueId");
        a.CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))dlsym(a.lib, "ncclCommInitRank");
        a.CommDestroy = (int (*)(NcclComm))dlsym(a.lib, "ncclCommDestroy");
        a.Send = (int (*)(const void*, size_t, int, int, NcclComm, cudaStream_t))dlsym(a.lib, "ncclSend");
        a.Recv = (int (*)(void*, size_t, int, int, NcclComm, cudaStream_t))dlsym(a.lib, "ncclRecv");
        a.GroupStart = (int (*)())dlsym(a.lib, "ncclGroupStart");
        a.GroupEnd = (int (*)())dlsym(a.lib, "ncclGroupEnd");
        a.GetErrorString = (const char* (*)(int))dlsym(a.lib, "ncclGetErrorString");
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.Send && a.Recv && a.GroupStart && a.GroupEnd;
        return a;
    }();
    return api;
}
constexpr int kNcclInt8 = 0;   // ncclInt8 / ncclChar

// ---- flags ----------------------------------------------------------------------------------------------------
constexpr int kMaxWorld = 64;
struct Flags {   // one per rank, in its own device memory, mapped into every peer
    unsigned long long arrive[kMaxWorld];   // consumer's copy: arrive[r] = rounds delivered by rank r
    unsigned long long freed;               // producer's copy: rounds the consumer has eaten
    unsigned long long error;               // local: a wait timed out
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// all earlier work of the stream (this GPU's stores into the peer's round buffer included) is complete when this kernel
// starts; the release store orders it before the flag for a reader that acquires the flag
__global__ void signal_kernel(unsigned long long* flag, unsigned long long value) {
    __threadfence_system();
    st_release_sys(flag, value);
}
// lanes 0..count-1 each watch one flag until it reaches `value`
__global__ void wait_kernel(const unsigned long long* flags, int count, int skip, unsigned long long value,
                            unsigned long long* error, unsigned long long timeout_ns) {
    const int i = threadIdx.x;
    if (i >= count || i == skip) return;
    if (ld_acquire_sys(error)) return;   // a wait already failed: do not add another time-out to the first
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(flags + i) < value) {
        if (global_ns() - t0 > timeout_ns) { atomicExch(error, 1ull); return; }
        __nanosleep(200);
    }
}
// the consumer tells every producer (through its mapped flags) that `value` rounds have been eaten
struct PeerFlags { unsigned long long* freed[kMaxWorld]; };
__global__ void release_kernel(const __grid_constant__ PeerFlags peers, int world, int skip, unsigned long long value) {
    const int i = threadIdx.x;
    if (i >= world || i == skip) return;
    __threadfence_system();
    st_release_sys(peers.freed[i], value);
}

struct Blob {   // what a rank publishes (zoicb_gather_export); ZOICB_GATHER_BLOB_BYTES bytes on the wire
    int32_t rank, world, consumer, has_data;
    uint64_t tile, slots;
    cudaIpcMemHandle_t flags, data;
};
static_assert(sizeof(Blob) <= ZOICB_GATHER_BLOB_BYTES, "blob fits its wire size");

}  // namespace

struct zoicb_gather {
    int device = 0, rank = 0, world = 1, consumer = 0, transport = ZOICB_GATHER_FUSED;
    uint64_t tile = 0;
    int slots = 2;
    bool connected = false;
    Flags* flags = nullptr;                 // own flags (device memory)
    RayRecord* data = nullptr;              // consumer: slots x world x tile records (own allocation)
    RayRecord* peer_data = nullptr;         // producer: the consumer's buffer, mapped
    Flags* peer_flags[kMaxWorld] = {};      // consumer: every producer's flags, mapped; producer: [consumer] only
    RayRecord* stage[2] = {nullptr, nullptr};   // producer, PUSH / NCCL: local staging tiles
    cudaStream_t s_copy = nullptr;          // producer: copies / sends; consumer: waits, receives, consumes
    cudaEvent_t ev_done[2] = {nullptr, nullptr};      // producer: staging tile shipped
    cudaEvent_t ev_gen = nullptr;                     // records of the current round complete on the caller's stream
    std::vector<cudaEvent_t> ev_eaten;                // consumer: slot consumed
    std::vector<uint64_t> counts;
    uint64_t base = 0, job_base = 0, job_rounds = 0;   // rounds of all earlier jobs; of the current / last job
    bool serial = false;                    // current job: ship / wait / consume on the caller's stream (no overlap; A/B)
    NcclComm comm = nullptr;
    bool own_comm = false;
    unsigned long long timeout_ns = 10ull * 1000 * 1000 * 1000;   // ZOICB_GATHER_TIMEOUT_S overrides

    bool is_consumer() const { return rank == consumer; }
    // `round` counts from the start of the current job; slots rotate on the global round number
    uint64_t global_round(uint64_t round) const { return job_base + round; }
    RayRecord* slot_base(RayRecord* buf, uint64_t round, int r) const {
        return buf + ((global_round(round) % (uint64_t)slots) * (uint64_t)world + (uint64_t)r) * tile;
    }
    uint64_t count_of(int r, uint64_t round) const {
        const uint64_t c = counts.empty() ? 0 : counts[r], b = round * tile;
        return c > b ? std::min<uint64_t>(tile, c - b) : 0;
    }
};

namespace zoicb {

int gather_device(const zoicb_gather* g) { return g->device; }
int gather_rank(const zoicb_gather* g) { return g->rank; }
uint64_t gather_tile_rays(const zoicb_gather* g) { return g->tile; }
uint64_t gather_rounds(const zoicb_gather* g, const uint64_t* counts) {
    uint64_t r = 0;
    for (int i = 0; i < g->world; ++i) r = std::max<uint64_t>(r, (counts[i] + g->tile - 1) / g->tile);
    return r;
}

cudaError_t gather_begin(zoicb_gather* g, const uint64_t* counts, cudaStream_t st, bool serial) {
    (void)st;
    if (!g->connected) return cudaErrorNotReady;
    g->serial = serial;
    g->counts.assign(counts, counts + g->world);
    g->job_base = g->base;
    g->job_rounds = gather_rounds(g, counts);
    g->base += g->job_rounds;
    return cudaSuccess;
}

cudaError_t gather_acquire(zoicb_gather* g, uint64_t round, cudaStream_t st, RayRecord** dst) {
    cudaError_t e;
    if (g->is_consumer()) {
        // own share goes straight into the round buffer; the slot is free once its previous round has been eaten
        const uint64_t G = g->global_round(round);
        if (G >= (uint64_t)g->slots && (e = cudaStreamWaitEvent(st, g->ev_eaten[G % g->slots], 0)) != cudaSuccess) return e;
        *dst = g->slot_base(g->data, round, g->rank);
        return cudaSuccess;
    }
    if (g->transport == ZOICB_GATHER_FUSED) {
        const uint64_t G = g->global_round(round);
        if (G >= (uint64_t)g->slots) {
            wait_kernel<<<1, 32, 0, st>>>(&g->flags->freed, 1, -1, G - g->slots + 1, &g->flags->error, g->timeout_ns);
            api_count_launches(1);
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
        }
        *dst = g->slot_base(g->peer_data, round, g->rank);
        return cudaSuccess;
    }
    // PUSH / NCCL: a local staging tile, free once the copy / send of two rounds ago has left it
    if (round >= 2 && (e = cudaStreamWaitEvent(st, g->ev_done[round & 1], 0)) != cudaSuccess) return e;
    *dst = g->stage[round & 1];
    return cudaSuccess;
}

cudaError_t gather_commit(zoicb_gather* g, uint64_t round, uint64_t m, cudaStream_t st, void* d_totals, int* launches) {
    cudaError_t e;
    NcclApi& N = nccl();
    const uint64_t G = g->global_round(round);
    cudaStream_t sc = g->serial ? st : g->s_copy;   // the stream that ships (producer) or waits and consumes (consumer)
    if (!g->is_consumer()) {
        Flags* cf = g->peer_flags[g->consumer];
        if (g->transport == ZOICB_GATHER_FUSED) {
            signal_kernel<<<1, 1, 0, st>>>(&cf->arrive[g->rank], G + 1);
            if (launches) *launches += 1;
            return cudaGetLastError();
        }
        if ((e = cudaEventRecord(g->ev_gen, st)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(sc, g->ev_gen, 0)) != cudaSuccess) return e;
        if (g->transport == ZOICB_GATHER_PUSH) {
            if (G >= (uint64_t)g->slots) {
                wait_kernel<<<1, 32, 0, sc>>>(&g->flags->freed, 1, -1, G - g->slots + 1, &g->flags->error, g->timeout_ns);
                if (launches) *launches += 1;
                if ((e = cudaGetLastError()) != cudaSuccess) return e;
            }
            if (m && (e = cudaMemcpyAsync(g->slot_base(g->peer_data, round, g->rank), g->stage[round & 1], m * sizeof(RayRecord),
                                          cudaMemcpyDefault, sc)) != cudaSuccess) return e;
            signal_kernel<<<1, 1, 0, sc>>>(&cf->arrive[g->rank], G + 1);
            if (launches) *launches += 1;
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
        } else {   // NCCL: the rendezvous with the consumer's receive is the back-pressure
            if (m && N.Send(g->stage[round & 1], m * sizeof(RayRecord), kNcclInt8, g->consumer, g->comm, sc) != 0) return cudaErrorUnknown;
        }
        return cudaEventRecord(g->ev_done[round & 1], sc);
    }
    // ---- consumer: everything below runs on its own stream, behind this round's own share
    if ((e = cudaEventRecord(g->ev_gen, st)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(sc, g->ev_gen, 0)) != cudaSuccess) return e;
    if (g->transport == ZOICB_GATHER_NCCL) {
        if (N.GroupStart() != 0) return cudaErrorUnknown;
        for (int r = 0; r < g->world; ++r) {
            const uint64_t mr = g->count_of(r, round);
            if (r == g->rank || !mr) continue;
            if (N.Recv(g->slot_base(g->data, round, r), mr * sizeof(RayRecord), kNcclInt8, r, g->comm, sc) != 0) { N.GroupEnd(); return cudaErrorUnknown; }
        }
        if (N.GroupEnd() != 0) return cudaErrorUnknown;
    } else {
        wait_kernel<<<1, 64, 0, sc>>>(g->flags->arrive, g->world, g->rank, G + 1, &g->flags->error, g->timeout_ns);
        if (launches) *launches += 1;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    // consume: one launch over the whole slot when every rank filled its segment, else segment by segment
    bool full = true;
    for (int r = 0; r < g->world; ++r) full = full && g->count_of(r, round) == g->tile;
    if (full) {
        if ((e = launch_consume(g->slot_base(g->data, round, 0), (uint64_t)g->world * g->tile, d_totals, sc, launches)) != cudaSuccess) return e;
    } else {
        for (int r = 0; r < g->world; ++r)
            if ((e = launch_consume(g->slot_base(g->data, round, r), g->count_of(r, round), d_totals, sc, launches)) != cudaSuccess) return e;
    }
    if (g->transport != ZOICB_GATHER_NCCL) {
        PeerFlags pf;
        for (int r = 0; r < g->world; ++r) pf.freed[r] = g->peer_flags[r] ? &g->peer_flags[r]->freed : nullptr;
        release_kernel<<<1, 64, 0, sc>>>(pf, g->world, g->rank, G + 1);
        if (launches) *launches += 1;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaEventRecord(g->ev_eaten[G % g->slots], sc);
}

cudaError_t gather_end(zoicb_gather* g, cudaStream_t st) {
    if (g->serial) return cudaSuccess;
    cudaError_t e = cudaEventRecord(g->ev_gen, g->s_copy);
    if (e != cudaSuccess) return e;
    return cudaStreamWaitEvent(st, g->ev_gen, 0);
}

bool gather_failed(zoicb_gather* g) {
    unsigned long long err = 0;
    if (cudaMemcpy(&err, &g->flags->error, sizeof err, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return true; }
    return err != 0;
}

}  // namespace zoicb

// ---------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------
extern "C" {

zoicb_status zoicb_gather_create(int device, int rank, int world, int consumer, uint64_t tile_rays, int slots, int transport,
                                 zoicb_gather** out) {
    if (!out) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_gather_create: null argument");
    *out = nullptr;
    if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world || consumer < 0 || consumer >= world || !tile_rays || slots < 1)
        return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_gather_create: bad rank / world / tile / slots");
    if (transport != ZOICB_GATHER_FUSED && transport != ZOICB_GATHER_PUSH && transport != ZOICB_GATHER_NCCL)
        return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_gather_create: unknown transport");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError();
        return api_fail(ZOICB_ERR_CUDA, "zoicb_gather_create: no such CUDA device (libzoicb has no CPU fallback)");
    }
    ZGUARD(device);
    zoicb_gather* g = new zoicb_gather();
    g->device = device; g->rank = rank; g->world = world; g->consumer = consumer; g->tile = tile_rays; g->slots = slots;
    g->transport = transport;
    if (const char* t = getenv("ZOICB_GATHER_TIMEOUT_S")) { const long v = atol(t); if (v > 0) g->timeout_ns = (unsigned long long)v * 1000000000ull; }
    cudaError_t e = cudaSuccess;
    do {
        if ((e = cudaMalloc(&g->flags, sizeof(Flags))) != cudaSuccess) break;
        if ((e = cudaMemset(g->flags, 0, sizeof(Flags))) != cudaSuccess) break;
        if ((e = cudaStreamCreateWithFlags(&g->s_copy, cudaStreamNonBlocking)) != cudaSuccess) break;
        if ((e = cudaEventCreateWithFlags(&g->ev_gen, cudaEventDisableTiming)) != cudaSuccess) break;
        for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&g->ev_done[i], cudaEventDisableTiming);
        if (e != cudaSuccess) break;
        if (g->is_consumer()) {
            if ((e = cudaMalloc(&g->data, (size_t)slots * world * tile_rays * sizeof(RayRecord))) != cudaSuccess) break;
            g->ev_eaten.resize(slots);
            for (int i = 0; i < slots && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&g->ev_eaten[i], cudaEventDisableTiming);
        } else if (transport != ZOICB_GATHER_FUSED) {
            for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaMalloc(&g->stage[i], tile_rays * sizeof(RayRecord));
        }
    } while (0);
    if (e != cudaSuccess) { zoicb_gather_destroy(g); return api_cuda_fail(e, "zoicb_gather_create"); }
    if (world == 1) g->connected = true;
    *out = g;
    return ZOICB_OK;
}

zoicb_status zoicb_gather_export(zoicb_gather* g, void* blob) {
    if (!g || !blob) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_gather_export: null argument");
    ZGUARD(g->device);
    Blob b;
    std::memset(&b, 0, sizeof b);
    b.rank = g->rank; b.world = g->world; b.consumer = g->consumer; b.has_data = g->data ? 1 : 0;
    b.tile = g->tile; b.slots = (uint64_t)g->slots;
    ZCUDA(cudaIpcGetMemHandle(&b.flags, g->flags), "cudaIpcGetMemHandle(flags)");
    if (g->data) ZCUDA(cudaIpcGetMemHandle(&b.data, g->data), "cudaIpcGetMemHandle(data)");
    std::memset(blob, 0, ZOICB_GATHER_BLOB_BYTES);
    std::memcpy(blob, &b, sizeof b);
    return ZOICB_OK;
}

zoicb_status zoicb_gather_connect(zoicb_gather* g, const void* blobs) {
    if (!g || !blobs) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_gather_connect: null argument");
    if (g->connected) return ZOICB_OK;
    ZGUARD(g->device);
    const char* p = static_cast<const char*>(blobs);
    for (int r = 0; r < g->world; ++r) {
        Blob b;
        std::memcpy(&b, p + (size_t)r * ZOICB_GATHER_BLOB_BYTES, sizeof b);
        if (b.rank != r || b.world != g->world || b.consumer != g->consumer || b.tile != g->tile || b.slots != (uint64_t)g->slots)
            return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_gather_connect: blob " + std::to_string(r) + " does not match this gather's shape");
        if (r == g->rank) continue;
        const bool want_flags = g->is_consumer() || r == g->consumer;
        if (want_flags && g->transport != ZOICB_GATHER_NCCL) {
            void* q = nullptr;
            ZCUDA(cudaIpcOpenMemHandle(&q, b.flags, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle(flags)");
            g->peer_flags[r] = static_cast<Flags*>(q);
        }
        if (r == g->consumer && g->transport != ZOICB_GATHER_NCCL) {
            if (!b.has_data) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_gather_connect: the consumer published no buffer");
            void* q = nullptr;
            ZCUDA(cudaIpcOpenMemHandle(&q, b.data, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle(data)");
            g->peer_data = static_cast<RayRecord*>(q);
        }
    }
    g->connected = true;
    return ZOICB_OK;
}

zoicb_status zoicb_nccl_unique_id(void* id128) {
    if (!id128) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_nccl_unique_id: null argument");
    NcclApi& N = nccl();
    if (!N.ok) return api_fail(ZOICB_ERR_UNSUPPORTED, "zoicb_nccl_unique_id: libnccl.so.2 could not be loaded");
    NcclUniqueId id;
    const int rc = N.GetUniqueId(&id);
    if (rc != 0) return api_fail(ZOICB_ERR_CUDA, std::string("ncclGetUniqueId: ") + (N.GetErrorString ? N.GetErrorString(rc) : "failed"));
    std::memcpy(id128, &id, sizeof id);
    return ZOICB_OK;
}

zoicb_status zoicb_gather_init_nccl(zoicb_gather* g, const void* id128) {
    if (!g || !id128) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_gather_init_nccl: null argument");
    NcclApi& N = nccl();
    if (!N.ok) return api_fail(ZOICB_ERR_UNSUPPORTED, "zoicb_gather_init_nccl: libnccl.so.2 could not be loaded");
    ZGUARD(g->device);
    NcclUniqueId id;
    std::memcpy(&id, id128, sizeof id);
    const int rc = N.CommInitRank(&g->comm, g->world, id, g->rank);
    if (rc != 0) return api_fail(ZOICB_ERR_CUDA, std::string("ncclCommInitRank: ") + (N.GetErrorString ? N.GetErrorString(rc) : "failed"));
    g->own_comm = true;
    if (g->transport == ZOICB_GATHER_NCCL) g->connected = true;
    return ZOICB_OK;
}

zoicb_status zoicb_gather_use_nccl_comm(zoicb_gather* g, void* nccl_comm) {
    if (!g || !nccl_comm) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_gather_use_nccl_comm: null argument");
    if (!nccl().ok) return api_fail(ZOICB_ERR_UNSUPPORTED, "zoicb_gather_use_nccl_comm: libnccl.so.2 could not be loaded");
    g->comm = nccl_comm;
    g->own_comm = false;
    if (g->transport == ZOICB_GATHER_NCCL) g->connected = true;
    return ZOICB_OK;
}

zoicb_status zoicb_gather_read(zoicb_gather* g, uint64_t round, int rank, uint64_t offset, uint64_t n, zoicb_ray* h_out) {
    if (!g || (n && !h_out)) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_gather_read: null argument");
    if (!g->is_consumer()) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_gather_read: only the consumer rank holds the round buffers");
    if (rank < 0 || rank >= g->world || offset > g->tile || n > g->tile - offset) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_gather_read: out of range");
    if (!n) return ZOICB_OK;
    ZGUARD(g->device);
    ZCUDA(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
    ZCUDA(cudaMemcpy(h_out, g->slot_base(g->data, round, rank) + offset, n * sizeof(RayRecord), cudaMemcpyDeviceToHost), "cudaMemcpy");
    return ZOICB_OK;
}

void zoicb_gather_destroy(zoicb_gather* g) {
    if (!g) return;
    DeviceGuard guard(g->device);
    cudaDeviceSynchronize();
    if (g->comm && g->own_comm && nccl().ok) nccl().CommDestroy(g->comm);
    for (int r = 0; r < kMaxWorld; ++r) if (g->peer_flags[r]) cudaIpcCloseMemHandle(g->peer_flags[r]);
    if (g->peer_data) cudaIpcCloseMemHandle(g->peer_data);
    cudaFree(g->flags); cudaFree(g->data);
    for (auto p : g->stage) cudaFree(p);
    if (g->s_copy) cudaStreamDestroy(g->s_copy);
    if (g->ev_gen) cudaEventDestroy(g->ev_gen);
    for (auto e : g->ev_done) if (e) cudaEventDestroy(e);
    for (auto e : g->ev_eaten) if (e) cudaEventDestroy(e);
    cudaGetLastError();
    delete g;
}

}  // extern "C"
