// One-shot all-reduce of small vectors over NVLink peer memory (the meta-gradient is 18-23 KB: pure latency).
//
// Every rank owns one IPC-exported buffer: two slots of `cap` floats + two slots of `world` arrival flags.  A call
//   1. copies the local vector into the rank's own slot (slot = epoch parity),
//   2. pushes the epoch number into its flag in EVERY rank's buffer (remote stores over NVLink),
//   3. spins on its own (local) flags until every rank has arrived,
//   4. sums all ranks' slots in RANK ORDER with peer loads -> every rank gets the bitwise identical result.
// It is a plain kernel launch, so - unlike an NCCL call in this PyTorch build - it can sit inside the CUDA graph
// of a meta-iteration; the epoch counter lives in device memory and advances on every replay.  Two slots are
// enough: a rank can only start epoch e+2 after every peer has signalled e+1, i.e. has finished reading epoch e.
#include <string.h>
#include "common.cuh"

namespace promp {

struct CommLayout {
    __host__ __device__ static int64_t data_off(int slot, int cap) { return (int64_t)slot * cap; }
    __host__ __device__ static int64_t flag_off(int slot, int cap, int world) { return 2 * (int64_t)cap + (int64_t)slot * world; }
};

__global__ void __launch_bounds__(1024) allreduce_p2p_kernel(int world, int rank, int n, int cap, const float* in, float* out,
                                                               float scale, float* const* peers, uint32_t* epoch_ptr,
                                                               uint32_t* error_flag) {
    const uint32_t epoch = *epoch_ptr + 1;
    const int slot = epoch & 1, tid = threadIdx.x;
    float* mine = peers[rank] + CommLayout::data_off(slot, cap);
    for (int p = tid; p < n; p += blockDim.x) mine[p] = in[p];
    __threadfence_system();
    __syncthreads();
    if (tid < world) {   // arrive: remote store of the epoch into rank `tid`'s flag array
        volatile uint32_t* f = reinterpret_cast<volatile uint32_t*>(peers[tid] + CommLayout::flag_off(slot, cap, world)) + rank;
        *f = epoch;
    }
    if (tid < world) {   // wait for rank `tid` (local polling), bounded so a mis-launch cannot hang the box
        volatile uint32_t* f = reinterpret_cast<volatile uint32_t*>(peers[rank] + CommLayout::flag_off(slot, cap, world)) + tid;
        const long long t0 = clock64();
        while (*f != epoch) {
            if (clock64() - t0 > 4000000000LL) {   // ~2 s
                *error_flag = 1;                   // sticky: every later call poisons its output too
                break;
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (*reinterpret_cast<volatile uint32_t*>(error_flag) != 0) {
        // A peer never arrived.  Summing whatever is in its slot would silently apply a partial / stale meta-gradient and
        // let the replicas diverge; poison the result instead so the failure is loud (NaN loss / parameters on this rank,
        // P2PComm.check() raises) and leave the epoch untouched.
        for (int p = tid; p < n; p += blockDim.x) out[p] = __int_as_float(0x7fc00000);
        return;
    }
    for (int p = tid; p < n; p += blockDim.x) {
        float s = 0.f;
        for (int r = 0; r < world; ++r) {
            const volatile float* src = peers[r] + CommLayout::data_off(slot, cap);
            s += src[p];       // volatile: peer data must not be served from a stale L1 line
        }
        out[p] = s * scale;
    }
    __syncthreads();
    if (tid == 0) *epoch_ptr = epoch;
}

}  // namespace promp

using namespace promp;

extern "C" int64_t promp_comm_buffer_bytes(int world, int capacity_floats) {
    return (2 * (int64_t)capacity_floats + 2 * (int64_t)world) * 4;
}

// The ONE place the library allocates: communication buffers must be whole cudaMalloc allocations to be IPC-exportable.
extern "C" int promp_comm_alloc(int64_t bytes, void** dev_ptr_host) {
    PROMP_REQUIRE(bytes > 0 && dev_ptr_host, "promp_comm_alloc: bad arguments");
    PROMP_CUDA(cudaMalloc(dev_ptr_host, (size_t)bytes));
    PROMP_CUDA(cudaMemset(*dev_ptr_host, 0, (size_t)bytes));
    return PROMP_OK;
}
extern "C" int promp_comm_free(void* dev_ptr) {
    PROMP_CUDA(cudaFree(dev_ptr));
    return PROMP_OK;
}
extern "C" int promp_ipc_get_handle(void* dev_ptr, void* handle64_host) {
    PROMP_REQUIRE(dev_ptr && handle64_host, "promp_ipc_get_handle: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    PROMP_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64_host), dev_ptr));
    return PROMP_OK;
}
extern "C" int promp_ipc_open_handle(const void* handle64_host, void** dev_ptr_host) {
    PROMP_REQUIRE(handle64_host && dev_ptr_host, "promp_ipc_open_handle: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64_host, sizeof(h));
    PROMP_CUDA(cudaIpcOpenMemHandle(dev_ptr_host, h, cudaIpcMemLazyEnablePeerAccess));
    return PROMP_OK;
}
extern "C" int promp_ipc_close_handle(void* dev_ptr) {
    PROMP_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return PROMP_OK;
}

extern "C" int promp_allreduce_p2p(int world, int rank, int n, int capacity_floats, const float* in, float* out, float scale,
                                   void* const* peers_dev, uint32_t* epoch_dev, uint32_t* error_flag_dev, void* stream) {
    PROMP_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world, "promp_allreduce_p2p: bad world/rank");
    PROMP_REQUIRE(n > 0 && n <= capacity_floats, "promp_allreduce_p2p: n=%d exceeds the buffer capacity %d", n, capacity_floats);
    PROMP_REQUIRE(in && out && peers_dev && epoch_dev && error_flag_dev, "promp_allreduce_p2p: null pointer argument");
    allreduce_p2p_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(world, rank, n, capacity_floats, in, out, scale,
                                                               reinterpret_cast<float* const*>(peers_dev), epoch_dev,
                                                               error_flag_dev);
    PROMP_LAUNCH_CHECK("allreduce_p2p_kernel");
    return PROMP_OK;
}
