// Shared device/host helpers for the promp_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/promp_b200.h"

namespace promp {

// ---------------------------------------------------------------- error plumbing (host)
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

#define PROMP_REQUIRE(cond, ...)                                   \
    do {                                                           \
        if (!(cond)) {                                             \
            promp::set_error(__VA_ARGS__);                         \
            return PROMP_ERR_INVALID_ARG;                          \
        }                                                          \
    } while (0)

#define PROMP_CUDA(call)                                           \
    do {                                                           \
        int _st = promp::check_cuda((call), #call);                \
        if (_st != PROMP_OK) return _st;                           \
    } while (0)

#define PROMP_LAUNCH_CHECK(name)                                   \
    do {                                                           \
        int _st = promp::check_cuda(cudaGetLastError(), name);     \
        if (_st != PROMP_OK) return _st;                           \
    } while (0)

// ---------------------------------------------------------------- parameter layout
// Flat parameter vector in the reference's creation order
// (ref: policies/gaussian_mlp_policy.py:55-80, policies/networks/mlp.py:100):
//   W0[Do,Hd] b0[Hd] W1[Hd,Hd] b1[Hd] W2[Hd,Da] b2[Da] log_std[Da]
template <int DO, int DA, int HID>
struct PLayout {
    static constexpr int W0 = 0;
    static constexpr int B0 = W0 + DO * HID;
    static constexpr int W1 = B0 + HID;
    static constexpr int B1 = W1 + HID * HID;
    static constexpr int W2 = B1 + HID;
    static constexpr int B2 = W2 + HID * DA;
    static constexpr int LS = B2 + DA;
    static constexpr int P = LS + DA;
};

__host__ __device__ inline int num_params(int Do, int Da, int Hd) {
    return Do * Hd + Hd + Hd * Hd + Hd + Hd * Da + Da + Da;
}

// ---------------------------------------------------------------- math
// tanh via one ex2.approx + one fast division: |abs error| <= ~2e-7 over the whole range (saturates to +-1
// exactly for |x| > 10), ~4x fewer instructions than tanhf.  MUFU.TANH (tanh.approx) is only 2^-11 accurate
// and would break the 1e-4 parity bar on gradients.
// Five instructions (FMUL, MUFU.EX2, FADD, MUFU.RCP, FFMA): the .ftz forms drop the denormal / huge-operand guard sequences of
// __expf / __fdividef (7-8 extra instructions per call, ~10 % of all instructions of the policy kernels), whose cases end in
// the same saturated values here (e -> 0: 1 - 2 = -1; e -> inf: 1 - 0 = 1).  2 * log2(e) is folded into one constant:
// (2 x) * c and x * (2 c) are the same real number, so the rounded product is bit-identical.
__device__ __forceinline__ float tanh_fast(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return fmaf(-2.0f, r, 1.0f);
}

// ---------------------------------------------------------------- warp helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------- TMA bulk copies (cp.async.bulk, 1-D) + mbarriers
// Contiguous global -> shared copies issued by ONE thread and completed on an mbarrier (SASS: UBLKCP).  Source, destination
// and size must be multiples of 16 bytes.
__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_addr_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tma_mbar_fence_init() {     // make the initialised barriers visible to the async proxy
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void tma_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}\n" ::"r"(smem_addr_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_addr_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_addr_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_addr_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------- Philox4x32-10
struct Philox {
    static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    __host__ __device__ static inline void round(uint32_t c[4], uint32_t k0, uint32_t k1) {
        uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    // counter = (c0,c1,c2,c3), key = seed
    __host__ __device__ static inline void gen(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint64_t seed,
                                               uint32_t out[4]) {
        uint32_t c[4] = {c0, c1, c2, c3};
        uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            round(c, k0, k1);
            k0 += W0;
            k1 += W1;
        }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    }
};
// uniform in (0,1]: never 0 so log() is finite
__host__ __device__ inline float u01(uint32_t x) { return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f); }
// two N(0,1) from two uint32 (Box-Muller)
__device__ inline void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
    float r = sqrtf(-2.0f * logf(u01(a)));
    float s, c;
    sincospif(2.0f * u01(b), &s, &c);
    z0 = r * c;
    z1 = r * s;
}

}  // namespace promp
