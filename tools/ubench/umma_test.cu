// Stand-alone validation of the tcgen05 (UMMA) building blocks planned for the policy kernels:
//   T1  D[128x64]  = A[128x64] . W[64x64]        A K-major (smem), B = W as MN-major
//   T2  D[128x64]  = A[128x64] . W^T             A K-major,        B = W as K-major
//   T3  G[64x64]   = A^T[64x128] . E[128x64]     A MN-major, B MN-major, M = 64, K = 128   (weight gradient)
// kind::tf32, fp32 accumulators in TMEM, single pass and the 3-term split (hi*hi + lo*hi + hi*lo).
// All tiles use the SWIZZLE_128B canonical layout: rows of 32 floats (128 B), 8-row groups of 1024 B,
// column blocks of 32 floats at `block stride`.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (r, c) inside a [R x 64] fp32 tile.
//  layout 0: two 32-column blocks, SWIZZLE_128B (K-major only)
//  layout 1: no swizzle, 8x4 core matrices (128 B each: row r%8 at 16 B stride), cores ordered column-of-cores major:
//            off = (c/4)*(R*16) + (r/8)*128 + (r%8)*16 + (c%4)*4.   The same bytes are a K-major tile (rows = M/N, cols = K)
//            and an MN-major tile (cols = M/N, rows = K).
__host__ __device__ inline int tile_off(int r, int c, int rows, int layout) {
    if (layout == 0) {
        const int blk = c >> 5, cc = c & 31;
        return blk * rows * 128 + (r >> 3) * 1024 + (r & 7) * 128 + ((((cc >> 2) ^ (r & 7))) << 4) + ((cc & 3) << 2);
    }
    return (c >> 2) * (rows * 16) + (r >> 3) * 128 + (r & 7) * 16 + ((c & 3) << 2);
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, int layout_type = 2) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;        // descriptor version (Blackwell)
    d |= (uint64_t)layout_type << 61;   // 2 = SWIZZLE_128B, 0 = no swizzle
    return d;
}
__host__ __device__ inline uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    uint32_t d = 0;
    d |= 1u << 4;                  // D format F32
    d |= 2u << 7;                  // A format TF32
    d |= 2u << 10;                 // B format TF32
    d |= (uint32_t)a_mn_major << 15;
    d |= (uint32_t)b_mn_major << 16;
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
        "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]),
          "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]),
          "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]),
          "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// mode 0: T1, mode 1: T2, mode 2: T3.  split: 0 single pass, 1 three-term.
__global__ void __launch_bounds__(128) umma_test_kernel(const float* A, const float* W, const float* E, float* out, int mode,
                                                          int split, int layout) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // layout: A_hi [128x64] 32K | A_lo 32K | W_hi [64x64] 16K | W_lo 16K | E_hi [128x64] 32K | E_lo 32K
    uint8_t* sA = smem;
    uint8_t* sAl = smem + 32768;
    uint8_t* sW = smem + 65536;
    uint8_t* sWl = smem + 81920;
    uint8_t* sE = smem + 98304;
    uint8_t* sEl = smem + 131072;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int i = tid; i < 128 * 64; i += 128) {
        const int r = i / 64, c = i % 64;
        const float a = A[i], e = E[i];
        *reinterpret_cast<float*>(sA + tile_off(r, c, 128, layout)) = a;
        *reinterpret_cast<float*>(sAl + tile_off(r, c, 128, layout)) = a - tf32_hi(a);
        *reinterpret_cast<float*>(sE + tile_off(r, c, 128, layout)) = e;
        *reinterpret_cast<float*>(sEl + tile_off(r, c, 128, layout)) = e - tf32_hi(e);
    }
    for (int i = tid; i < 64 * 64; i += 128) {
        const int r = i / 64, c = i % 64;
        const float w = W[i];
        *reinterpret_cast<float*>(sW + tile_off(r, c, 64, layout)) = w;
        *reinterpret_cast<float*>(sWl + tile_off(r, c, 64, layout)) = w - tf32_hi(w);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "n"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (tid == 0) mbar_init(&bar, 1);
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");      // generic-proxy smem writes -> async proxy (UMMA)
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (mode == 3) {       // A_lo operand from TMEM: lane = row, column = k
        for (int hb = 0; hb < 2; ++hb) {
            float v[32];
            for (int c = 0; c < 32; ++c) { const float a = A[tid * 64 + hb * 32 + c]; v[c] = a - tf32_hi(a); }
            tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + 128 + hb * 32, v);
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    }

    if (tid == 0) {
        const int nterm = split ? 3 : 1;
        uint32_t acc = 0;
        if (mode == 3) {
            // D = A.W^T, K-major smem operands (no-swizzle core tiles) except the A_lo term, which comes from TMEM
            const uint32_t ScA = 128 * 16, ScW = 64 * 16, Sr = 128;
            const uint32_t idesc = make_idesc(128, 64, 0, 0);
            for (int s = 0; s < 8; ++s) {     // lo(TMEM) * hi
                umma_tf32_ts(tmem, tmem + 128 + 8 * s, make_desc(smem_u32(sW) + 2 * s * ScW, ScW, Sr, 0), idesc, acc);
                acc = 1;
            }
            for (int term = 1; term < 3; ++term) {
                const uint8_t* w = (term == 1) ? sWl : sW;
                for (int s = 0; s < 8; ++s)
                    umma_tf32(tmem, make_desc(smem_u32(sA) + 2 * s * ScA, ScA, Sr, 0), make_desc(smem_u32(w) + 2 * s * ScW, ScW, Sr, 0), idesc, 1);
            }
        } else if (layout == 1) {
            // no-swizzle core-matrix tiles: S_r = 128 B (next 8 rows), S_c = rows*16 B (next 4 columns)
            const uint32_t ScA = 128 * 16, ScW = 64 * 16, Sr = 128;
            if (mode == 0 || mode == 1) {
                const uint32_t idesc = make_idesc(128, 64, 0, mode == 0 ? 1 : 0);
                for (int term = 0; term < nterm; ++term) {
                    const uint8_t* a = (term == 1) ? sAl : sA;
                    const uint8_t* w = (term == 2) ? sWl : sW;
                    for (int s = 0; s < 8; ++s) {
                        const uint64_t ad = make_desc(smem_u32(a) + 2 * s * ScA, ScA, Sr, 0);          // K-major: LBO = K-core stride, SBO = M-core stride
                        uint64_t bd;
                        if (mode == 0) bd = make_desc(smem_u32(w) + s * Sr, Sr, ScW, 0);               // MN-major: SBO = MN-core stride, LBO = K-group stride
                        else bd = make_desc(smem_u32(w) + 2 * s * ScW, ScW, Sr, 0);                    // K-major
                        umma_tf32(tmem, ad, bd, idesc, acc);
                        acc = 1;
                    }
                }
            } else {
                const uint32_t idesc = make_idesc(64, 64, 1, 1);
                for (int term = 0; term < nterm; ++term) {
                    const uint8_t* a = (term == 1) ? sAl : sA;
                    const uint8_t* e = (term == 2) ? sEl : sE;
                    for (int s = 0; s < 16; ++s) {
                        const uint64_t ad = make_desc(smem_u32(a) + s * Sr, Sr, ScA, 0);
                        const uint64_t bd = make_desc(smem_u32(e) + s * Sr, Sr, ScA, 0);
                        umma_tf32(tmem, ad, bd, idesc, acc);
                        acc = 1;
                    }
                }
            }
        } else
        if (mode == 0 || mode == 1) {
            const uint32_t idesc = make_idesc(128, 64, 0, mode == 0 ? 1 : 0);
            for (int term = 0; term < nterm; ++term) {
                const uint8_t* a = (term == 1) ? sAl : sA;      // hi*hi, lo*hi, hi*lo
                const uint8_t* w = (term == 2) ? sWl : sW;
                for (int s = 0; s < 8; ++s) {
                    const uint64_t ad = make_desc(smem_u32(a) + (s >> 2) * 16384 + (s & 3) * 32, 16, 1024);
                    uint64_t bd;
                    if (mode == 0) bd = make_desc(smem_u32(w) + s * 1024, 8192, 1024);                       // W rows k = K, MN-major
                    else bd = make_desc(smem_u32(w) + (s >> 2) * 8192 + (s & 3) * 32, 16, 1024);           // W rows = N, K-major
                    umma_tf32(tmem, ad, bd, idesc, acc);
                    acc = 1;
                }
            }
        } else {
            const uint32_t idesc = make_idesc(64, 64, 1, 1);
            for (int term = 0; term < nterm; ++term) {
                const uint8_t* a = (term == 1) ? sAl : sA;
                const uint8_t* e = (term == 2) ? sEl : sE;
                for (int s = 0; s < 16; ++s) {                 // K = 128 sample rows, 8 per MMA
                    const uint64_t ad = make_desc(smem_u32(a) + s * 1024, 16384, 1024);
                    const uint64_t bd = make_desc(smem_u32(e) + s * 1024, 16384, 1024);
                    umma_tf32(tmem, ad, bd, idesc, acc);
                    acc = 1;
                }
            }
        }
        umma_commit(&bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    float v[64];
    tmem_ld64(tmem + ((uint32_t)(warp * 32) << 16), v);
    for (int c = 0; c < 64; ++c) out[tid * 64 + c] = v[c];      // lane (= tid) major dump of all 128 lanes x 64 cols
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(256));
}

static float tf32h(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u &= 0xffffe000u;
    memcpy(&x, &u, 4);
    return x;
}

int main() {
    std::vector<float> A(128 * 64), W(64 * 64), E(128 * 64), out(128 * 64);
    srand(1);
    for (auto& x : A) x = (rand() / (float)RAND_MAX) * 2 - 1;
    for (auto& x : W) x = (rand() / (float)RAND_MAX) * 0.5f - 0.25f;
    for (auto& x : E) x = (rand() / (float)RAND_MAX) * 2 - 1;
    float *dA, *dW, *dE, *dO;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dE, E.size() * 4); cudaMalloc(&dO, out.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dE, E.data(), E.size() * 4, cudaMemcpyHostToDevice);
    const int smem = 163840 + 1024;
    cudaFuncSetAttribute(umma_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int layout = 1; layout < 2; ++layout)
    for (int mode = 1; mode < 4; mode += 2)
        for (int split = 1; split < 2; ++split) {
            cudaMemset(dO, 0, out.size() * 4);
            umma_test_kernel<<<1, 128, smem>>>(dA, dW, dE, dO, mode, split, layout);
            cudaError_t err = cudaDeviceSynchronize();
            if (err != cudaSuccess) { printf("mode %d split %d: CUDA error %s\n", mode, split, cudaGetErrorString(err)); return 1; }
            cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost);
            const int M = mode == 2 ? 64 : 128;
            // try the candidate lane mappings for M = 64 (rows -> lanes): identity, and 16 rows per 32-lane quadrant
            for (int mapping = 0; mapping < (mode == 2 ? 2 : 1); ++mapping) {
                double max_err_tf = 0, max_err_f32 = 0, max_ref = 0;
                for (int m = 0; m < M; ++m)
                    for (int n = 0; n < 64; ++n) {
                        double ref = 0, ref_tf = 0;
                        const int K = mode == 2 ? 128 : 64;
                        for (int k = 0; k < K; ++k) {
                            float a, b;
                            if (mode == 0) { a = A[m * 64 + k]; b = W[k * 64 + n]; }
                            else if (mode == 1 || mode == 3) { a = A[m * 64 + k]; b = W[n * 64 + k]; }
                            else { a = A[k * 64 + m]; b = E[k * 64 + n]; }
                            ref += (double)a * b;
                            ref_tf += (double)tf32h(a) * tf32h(b);
                        }
                        const int lane = (mode == 2 && mapping == 1) ? ((m / 16) * 32 + (m % 16)) : m;
                        const double got = out[lane * 64 + n];
                        max_err_tf = fmax(max_err_tf, fabs(got - ref_tf));
                        max_err_f32 = fmax(max_err_f32, fabs(got - ref));
                        max_ref = fmax(max_ref, fabs(ref));
                    }
                printf("layout %d mode %d split %d mapping %d: max|ref| %.3f  max err vs tf32-truncated ref %.3e  vs fp64 ref %.3e\n", layout, mode, split,
                       mapping, max_ref, max_err_tf, max_err_f32);
            }
        }
    return 0;
}
