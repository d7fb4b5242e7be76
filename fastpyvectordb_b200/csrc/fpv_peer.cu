// Exchange of the row-sharded search over NVLink peer memory instead of a collective library call.
//
// The reference merges its chunks inside one process (_merge_top_k over a list, parallel_search.py:356-363); with one
// process per GPU the chunk results have to travel.  Each rank owns one IPC-exported region; a producer kernel stores
// its contribution straight into slot `rank` of EVERY peer's region (NVLink P2P stores, 16 bytes per thread per peer)
// and then publishes an epoch number in the peers' flag words; the consumer kernels (the phase-2 re-rank, the merge)
// spin on their LOCAL flag words at kernel start and read the gathered data from local memory.  No collective launch,
// no host round trip; the transfers of the different ranks overlap each other and the tail of the producing kernels.
//
// Memory model: a producer block fences (system scope) after its stores and bumps a device counter; the last block
// of the grid, having seen every other block's arrival, publishes the flags -- the threadfence-reduction pattern.
// Consumers read flags with volatile loads and the data with ld.global.cg (never through L1, which may hold lines of
// the previous epoch).  Two data buffers alternate by epoch parity: a rank that runs one step ahead cannot overwrite a
// buffer a slower peer is still reading (it can only start step n+2 after that peer has finished publishing n+1,
// which in stream order is after the peer's consumers of step n).
#include <cuda.h>

#include <algorithm>
#include <cstring>

#include "fpv_common.cuh"

namespace fpv {

__global__ void __launch_bounds__(256) peer_put_kernel(const uint4* __restrict__ src, size_t n16, void* const* __restrict__ peers,
                                                       int shards, int rank, size_t data_off, size_t slot_bytes, size_t flag_off,
                                                       uint32_t epoch, uint32_t* __restrict__ done_counter) {
    const size_t base = data_off + (size_t)rank * slot_bytes;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = src[i];
        for (int p = 0; p < shards; ++p)
            reinterpret_cast<uint4*>(static_cast<char*>(peers[p]) + base)[i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t prev = atomicAdd(done_counter, 1u);
        if (prev == gridDim.x - 1) {                     // every block's stores are visible system-wide: publish
            *done_counter = 0u;
            __threadfence_system();
            for (int p = 0; p < shards; ++p)
                *reinterpret_cast<volatile uint32_t*>(static_cast<char*>(peers[p]) + flag_off + 4 * (size_t)rank) = epoch;
        }
    }
}

}  // namespace fpv

using namespace fpv;

// A device region other processes of this node can map: cudaMalloc + zero fill + IPC handle (64 bytes).
extern "C" int fpv_peer_alloc(size_t bytes, void** out_ptr, unsigned char* handle64) {
    FPV_REQUIRE(bytes > 0 && out_ptr && handle64, "peer_alloc: bad argument");
    void* p = nullptr;
    FPV_CUDA(cudaMalloc(&p, bytes));
    FPV_CUDA(cudaMemset(p, 0, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "cudaIpcGetMemHandle"); }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handle64, &h, 64);
    *out_ptr = p;
    return FPV_OK;
}

extern "C" int fpv_peer_open(const unsigned char* handle64, void** out_ptr) {
    FPV_REQUIRE(handle64 && out_ptr, "peer_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    FPV_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *out_ptr = p;
    return FPV_OK;
}

extern "C" int fpv_peer_close(void* ptr) {
    if (ptr) FPV_CUDA(cudaIpcCloseMemHandle(ptr));
    return FPV_OK;
}

extern "C" int fpv_peer_free(void* ptr) {
    if (ptr) FPV_CUDA(cudaFree(ptr));
    return FPV_OK;
}

// All-gather by peer stores: src[0, nbytes) -> slot `rank` (slot stride slot_bytes, first slot at data_off) of every region
// in `peers` (a DEVICE array of `shards` region base pointers, our own included), then flag word `rank` at flag_off of
// every region := epoch.  nbytes must be a multiple of 16, src 16-byte aligned.  done_counter: a zeroed device uint32.
extern "C" int fpv_peer_put(const void* src, size_t nbytes, void* const* peers, int shards, int rank, size_t data_off,
                            size_t slot_bytes, size_t flag_off, uint32_t epoch, uint32_t* done_counter, void* stream) {
    FPV_REQUIRE(src && peers && done_counter && shards >= 1 && rank >= 0 && rank < shards, "peer_put: bad argument");
    FPV_REQUIRE(nbytes % 16 == 0 && nbytes <= slot_bytes && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && data_off % 16 == 0 &&
                    slot_bytes % 16 == 0 && flag_off % 4 == 0, "peer_put: sizes / alignment");
    const size_t n16 = nbytes / 16;
    int64_t blocks = (int64_t)((n16 + 255) / 256);
    blocks = std::max<int64_t>(1, std::min<int64_t>(blocks, (int64_t)sm_count() * 2));
    peer_put_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(static_cast<const uint4*>(src), n16, peers, shards, rank, data_off,
                                                                         slot_bytes, flag_off, epoch, done_counter);
    FPV_LAUNCH_CHECK();
    return FPV_OK;
}
