// Tensor-core (tcgen05 / TMEM) variant of policy_grad_kernel for HID = 64.
//
// The two [128 x 64] x [64 x 64] layer GEMMs of a tile (forward H1.W1 and backward D2.W1^T) run on the 5th-gen tensor
// cores as tcgen05.mma.kind::tf32 with the 3-term split  a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo  (a_hi = the fp32
// value itself, which the tensor core truncates to TF32; a_lo = a - trunc_tf32(a)), i.e. fp32-level accuracy
// (1.5e-6 relative, tools/ubench/umma_test.cu) at 24 MMAs per GEMM.  Accumulators live in TMEM (2 x 64 columns) and
// are read back with tcgen05.ld (one sample row per thread, 64 / NQ columns each; NQ = 2 or 4 column groups = 256 or 512
// threads per CTA); operands are written by the epilogue
// threads straight into the UMMA canonical K-major layout WITHOUT swizzle: 8x4-float core matrices (128 contiguous
// bytes, rows 16 B apart), next 8 rows at +128 B (SBO), next 4 columns at +S_c (LBO).  S_c = rows*16 + 16 bytes: the
// extra 16 B make both the row-wise 16-byte stores of the epilogue and the column-wise scalar reads of the SIMT
// reductions bank-conflict free, and the same bytes stay readable as a plain fp32 tile by the CUDA-core code.
// The weight-gradient GEMM H1^T.D2 contracts over samples, i.e. needs MN-major operands, which for tf32 exist only
// in the 128B_BASE32B layout (a second copy of every tile): it runs on the warp-level tensor-core path (mma.sync
// m16n8k8 tf32, same 3-term split, fragments loaded straight from the K-major tiles: wgrad_mma_tile below) CONCURRENTLY
// with the backward tcgen05 MMA.  One elected thread of a warp-uniform branch issues the tcgen05 MMAs; completion is
// an mbarrier (tcgen05.commit).
#pragma once
#include "mlp_tile.cuh"

namespace promp {

constexpr int TBT = 128;                      // samples per tile = UMMA M
constexpr int TC_HID = 64;
constexpr int SCA = TBT * 16 + 16;            // bytes between 4-column chunks of a [128 x 64] activation tile
constexpr int SCW = TC_HID * 16 + 16;         // ... of a [64 x 64] weight tile
constexpr int TILE_A_BYTES = 16 * SCA;        // 33 024
constexpr int TILE_W_BYTES = 16 * SCW;        // 16 640

__device__ __forceinline__ int core_off(int r, int c, int sc) {      // byte offset of element (r, c)
    return (c >> 2) * sc + (r >> 3) * 128 + (r & 7) * 16 + ((c & 3) << 2);
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// K-major, no swizzle: LBO = byte distance between core matrices adjacent in K, SBO = ... adjacent in M/N
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;        // descriptor version (sm_100)
    return d;                      // layout type 0 = no swizzle
}
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);   // D f32, A/B tf32, K-major
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tMBAR_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra MBAR_DONE_%=;\n\tbra MBAR_WAIT_%=;\n\tMBAR_DONE_%=:\n\t}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
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
__device__ __forceinline__ void tmem_ldw(uint32_t taddr, float (&v)[32]) { tmem_ld32(taddr, v); }
__device__ __forceinline__ void tmem_ldw(uint32_t taddr, float (&v)[16]) { tmem_ld16(taddr, v); }
__device__ __forceinline__ void tmem_stw(uint32_t taddr, const float (&v)[32]) { tmem_st32(taddr, v); }
__device__ __forceinline__ void tmem_stw(uint32_t taddr, const float (&v)[16]) { tmem_st16(taddr, v); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void proxy_fence_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// D[128 x 64] (TMEM columns d_tmem..+63) = A[128 x 64] . B^T with B [64(N) x 64(K)], 3-term TF32 split, issued by ONE thread.
__device__ __forceinline__ void issue_gemm_3xtf32(uint32_t d_tmem, const unsigned char* a_hi, const unsigned char* a_lo,
                                                  const unsigned char* b_hi, const unsigned char* b_lo) {
    const uint32_t idesc = umma_idesc_tf32(TBT, TC_HID);
    uint32_t acc = 0;
#pragma unroll
    for (int term = 0; term < 3; ++term) {                 // lo*hi, hi*lo (small terms first), hi*hi
        const unsigned char* a = (term == 0) ? a_lo : a_hi;
        const unsigned char* b = (term == 1) ? b_lo : b_hi;
        const uint32_t a0 = smem_u32(a), b0 = smem_u32(b);
#pragma unroll
        for (int s = 0; s < TC_HID / 8; ++s) {             // K = 8 per MMA = two 4-column chunks
            umma_tf32(d_tmem, umma_desc(a0 + 2 * s * SCA, SCA, 128), umma_desc(b0 + 2 * s * SCW, SCW, 128), idesc, acc);
            acc = 1;
        }
    }
}

template <int DO, int DA>
struct SmallLayout {      // the parameters other than W1, compact
    static constexpr int W0 = 0;
    static constexpr int B0 = W0 + DO * TC_HID;
    static constexpr int B1 = B0 + TC_HID;
    static constexpr int W2 = B1 + TC_HID;
    static constexpr int B2 = W2 + TC_HID * DA;
    static constexpr int LS = B2 + DA;
    static constexpr int SIZE = (LS + DA + 3) / 4 * 4;
};

template <int DO, int DA, int NQ>
struct GradTcSmem {
    static constexpr int DOP = DOPad<DO>::V;
    alignas(16) unsigned char W1T_hi[TILE_W_BYTES];   // B of the forward GEMM: rows n = output unit, K = k
    alignas(16) unsigned char W1T_lo[TILE_W_BYTES];
    alignas(16) unsigned char W1_hi[TILE_W_BYTES];    // B of the backward GEMM: rows n = k', K = j
    alignas(16) unsigned char W1_lo[TILE_W_BYTES];
    alignas(16) unsigned char A0[TILE_A_BYTES];       // H1 -> D1
    alignas(16) unsigned char A1[TILE_A_BYTES];       // H2 -> D2 (hi)
    alignas(16) unsigned char LO[TILE_A_BYTES];       // H1_lo -> D2_lo ; flush scratch
    alignas(16) float Ps[SmallLayout<DO, DA>::SIZE];
    alignas(16) float X[TBT * DOP];
    float MUP[NQ * TBT * DA];
    float DMU[TBT * DA];
    float red[3 * 4 * NQ];
    alignas(8) uint64_t bar;
    uint32_t tmem_base;
    int last;
};


// -------------------------------------------------------------------------------------------------------------
// Weight-gradient GEMM  acc[k][j] += sum_b A[b][k] * (scale * D[b][j])  over the 128 rows of two fp32 tiles in the
// K-major core-matrix layout, on the warp-level tensor-core path (mma.sync m16n8k8 tf32, 3xTF32 split).  This GEMM
// contracts over SAMPLES, i.e. both tcgen05 operands would be MN-major, which kind::tf32 supports only in the
// 128B_BASE32B layout - incompatible with the K-major use of the same tiles by the layer MMAs.  The CUDA-core version
// was bound by shared-memory wavefronts (every LDS.128 costs 4, 8 per sample row per warp for 16 FFMA); the fragment
// loads below are 12 conflict-free LDS.32 per 8 sample rows and the math leaves the FP32 pipe.
//   warp w owns the 16 x 8NT block  k in [16 (w&3), +16), j in [8NT (w>>2), +8NT)  as NT n-tiles of m16n8
//   (NT = 4 with 8 warps, 2 with 16 warps).
//   fragment column t (t+4) <-> sample 8s+2t (8s+2t+1): with the 16-byte chunk padding this makes every fragment
//   load hit 32 distinct banks (bank = 4 (g>>2) + (g&3) + 8 t).
__device__ __forceinline__ void mma_tf32_16n8k8(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
// COLSUM: also accumulate the column sums of D (unscaled) from the B fragments: csum[nt] holds, per lane, the partial
// over this lane's sample rows for column 32 (warp>>2) + 8 nt + (lane>>2); reduce over lane&3 at flush time.
template <bool COLSUM, int NT>
__device__ __forceinline__ void wgrad_mma_tile(const unsigned char* __restrict__ At, const unsigned char* __restrict__ Dt,
                                               float scale, int warp, int lane, float (&acc)[NT][4], float (&csum)[NT]) {
    const int g = lane >> 2, t = lane & 3;
    const int k0 = 16 * (warp & 3) + g, j0 = 8 * NT * (warp >> 2) + g;
    const unsigned char* ap = At + (k0 >> 2) * SCA + (k0 & 3) * 4 + t * 32;     // rows k0 (and k0+8: two chunks further)
    const unsigned char* dp = Dt + (j0 >> 2) * SCA + (j0 & 3) * 4 + t * 32;     // n-tile nt: two chunks further each
#pragma unroll 2
    for (int s = 0; s < TBT / 8; ++s) {
        uint32_t ah[4], al[4], bh[NT][2], bl[NT][2];
        split_tf32(*reinterpret_cast<const float*>(ap + s * 128), ah[0], al[0]);
        split_tf32(*reinterpret_cast<const float*>(ap + s * 128 + 2 * SCA), ah[1], al[1]);
        split_tf32(*reinterpret_cast<const float*>(ap + s * 128 + 16), ah[2], al[2]);
        split_tf32(*reinterpret_cast<const float*>(ap + s * 128 + 2 * SCA + 16), ah[3], al[3]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const float d0 = *reinterpret_cast<const float*>(dp + nt * 2 * SCA + s * 128);
            const float d1 = *reinterpret_cast<const float*>(dp + nt * 2 * SCA + s * 128 + 16);
            if (COLSUM) csum[nt] += d0 + d1;
            split_tf32(scale * d0, bh[nt][0], bl[nt][0]);
            split_tf32(scale * d1, bh[nt][1], bl[nt][1]);
        }
        // term-major order: consecutive MMAs hit different accumulators (no back-to-back dependent issue)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32_16n8k8(acc[nt], al, bh[nt]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32_16n8k8(acc[nt], ah, bl[nt]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32_16n8k8(acc[nt], ah, bh[nt]);
    }
}
// accumulator element (nt, i) of wgrad_mma_tile -> flat index into the [HID, HID] weight (row k, column j)
template <int NT>
__device__ __forceinline__ int wgrad_mma_index(int warp, int lane, int nt, int i) {
    const int g = lane >> 2, t = lane & 3;
    return (16 * (warp & 3) + g + 8 * (i >> 1)) * TC_HID + 8 * NT * (warp >> 2) + 8 * nt + 2 * t + (i & 1);
}

#ifdef PROMP_EXP_CLOCKS
// experiment build only: per-phase clock64 totals of CTA 0 (tools/kernel_time.py --clocks)
__device__ unsigned long long g_phase_clk[16];
#define PCLK(i)                                                   \
    do {                                                          \
        if (blockIdx.x == 0 && threadIdx.x == 0) {                \
            const long long t_ = clock64();                       \
            s_clk[i] += (unsigned long long)(t_ - s_last);        \
            s_last = t_;                                          \
        }                                                         \
    } while (0)
#else
#define PCLK(i)
#endif
template <int DO, int DA, int NQ>
__global__ void __launch_bounds__(128 * NQ, 1) policy_grad_tc_kernel(PolicyArgs A) {
    constexpr int HID = TC_HID;
    using L = PLayout<DO, DA, HID>;
    using SL = SmallLayout<DO, DA>;
    using SM = GradTcSmem<DO, DA, NQ>;
    constexpr int TCT = 128 * NQ, CW = TC_HID / NQ, NW = TCT / 32, NT = 8 / NQ;   // threads, columns per thread, warps, wgrad n-tiles per warp
    constexpr int DOP = SM::DOP;
    constexpr int PSTRIDE = L::P + PSTAT;
    constexpr int NPART = TCT / HID, BPP = TBT / NPART;     // column role: 4 slices of 32 rows

    extern __shared__ __align__(16) unsigned char smem_raw[];
    SM& S = *reinterpret_cast<SM*>(smem_raw);
#ifdef PROMP_EXP_CLOCKS
    __shared__ unsigned long long s_clk[16];
    __shared__ long long s_last;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; ++i) s_clk[i] = 0;
        s_last = clock64();
    }
#endif

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);      // provably warp-uniform copy for the MMA-issue branch
    const int qd = warp & 3, cq = warp >> 2;               // TMEM lane quadrant, column group
    const int r = qd * 32 + lane, c0 = CW * cq;          // row / column-group role: sample row r, hidden units [c0, c0+32)
    const int cj = tid & (HID - 1), cp = tid / HID;        // column role
    if (grad_reuse_prologue<L::P, L::LS, DA>(A)) return;    // before any TMEM allocation / barrier initialisation
    const TileSched ts(A.M, A.N, A.q, TBT);
    const int N = A.N;
    float invN = 1.0f / (float)N;       // both re-set per task when A.n_valid is given (variable-length paths)
    int Nm = N;
    const bool want_grad = A.grad != nullptr;
    const float* th = nullptr;
    HeadIn<DA> hin;
    uint32_t phase = 0;

    // ---- TMEM + mbarrier setup
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&S.tmem_base)), "n"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (tid == 0) mbar_init(&S.bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;
    const uint32_t tmem_row = tmem + ((uint32_t)(qd * 32) << 16);

    float gW1[NT][4], gB1f[NT], gW0p[DO], gW2p[DA], gB0c, gB2w[DA], gLSw[DA];   // gB2w/gLSw: per-warp partials in lane 0
    float s_obj, s_kl, s_ratio;
    auto zero_acc = [&]() {
#pragma unroll
        for (int a = 0; a < NT; ++a) gW1[a][0] = gW1[a][1] = gW1[a][2] = gW1[a][3] = gB1f[a] = 0.f;
#pragma unroll
        for (int i = 0; i < DO; ++i) gW0p[i] = 0.f;
#pragma unroll
        for (int d = 0; d < DA; ++d) gW2p[d] = 0.f;
#pragma unroll
        for (int d = 0; d < DA; ++d) gB2w[d] = gLSw[d] = 0.f;
        gB0c = 0.f;
        s_obj = s_kl = s_ratio = 0.f;
    };
    auto load_task = [&](int m, bool first) {
        if (A.n_valid) { Nm = __ldg(A.n_valid + m); invN = 1.0f / (float)max(Nm, 1); }
        th = A.params + (int64_t)m * A.param_stride;
        if (!first && A.param_stride == 0) return;
        __syncthreads();
        for (int i = tid; i < DO * HID + HID; i += TCT) S.Ps[SL::W0 + i] = __ldg(th + L::W0 + i);          // W0, b0
        for (int i = tid; i < HID; i += TCT) S.Ps[SL::B1 + i] = __ldg(th + L::B1 + i);
        for (int i = tid; i < HID * DA + 2 * DA; i += TCT) S.Ps[SL::W2 + i] = __ldg(th + L::W2 + i);       // W2, b2, ls
        for (int i = tid; i < HID * HID; i += TCT) {
            const int k = i / HID, j = i % HID;
            const float w = __ldg(th + L::W1 + i), wl = w - tf32_trunc(w);
            *reinterpret_cast<float*>(S.W1_hi + core_off(k, j, SCW)) = w;
            *reinterpret_cast<float*>(S.W1_lo + core_off(k, j, SCW)) = wl;
            *reinterpret_cast<float*>(S.W1T_hi + core_off(j, k, SCW)) = w;
            *reinterpret_cast<float*>(S.W1T_lo + core_off(j, k, SCW)) = wl;
        }
        __syncthreads();
#pragma unroll
        for (int d = 0; d < DA; ++d) {
            const float raw = S.Ps[SL::LS + d];
            const bool clipped = A.clip_log_std && (raw < A.min_log_std);
            hin.ls[d] = clipped ? A.min_log_std : raw;
            hin.ls_mask[d] = clipped ? 0.f : 1.f;
            hin.sig[d] = expf(hin.ls[d]);
        }
    };
    auto flush = [&](int m) {
        float* part = A.partial + ((int64_t)blockIdx.x * A.kmax + (m - ts.first_task(blockIdx.x))) * PSTRIDE;
        float* scr = reinterpret_cast<float*>(S.A1);      // A1 + LO (contiguous, 2 tiles): free between tiles (all MMAs have completed)
        static_assert(NPART * DO * HID * 4 <= 2 * TILE_A_BYTES && NPART * HID * DA * 4 <= 2 * TILE_A_BYTES, "flush scratch");
        __syncthreads();
        if (want_grad) {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) part[L::W1 + wgrad_mma_index<NT>(warp, lane, nt, i)] = gW1[nt][i];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {            // b1 gradient: fragment column sums, reduced over the 4 lanes of a column
                float c = gB1f[nt];
                c += __shfl_xor_sync(0xffffffffu, c, 1);
                c += __shfl_xor_sync(0xffffffffu, c, 2);
                if ((lane & 3) == 0 && (warp & 3) == 0) part[L::B1 + 8 * NT * (warp >> 2) + 8 * nt + (lane >> 2)] = c;
            }
            scr[cp * HID + cj] = gB0c;
            __syncthreads();
            if (tid < HID) {
                float s = 0.f;
                for (int p = 0; p < NPART; ++p) s += scr[p * HID + tid];
                part[L::B0 + tid] = s;
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < DO; ++i) scr[(cp * DO + i) * HID + cj] = gW0p[i];
            __syncthreads();
            for (int idx = tid; idx < DO * HID; idx += TCT) {
                float s = 0.f;
                for (int p = 0; p < NPART; ++p) s += scr[p * DO * HID + idx];
                part[L::W0 + idx] = s;
            }
            __syncthreads();
#pragma unroll
            for (int d = 0; d < DA; ++d) scr[(cp * HID + cj) * DA + d] = gW2p[d];
            __syncthreads();
            for (int idx = tid; idx < HID * DA; idx += TCT) {
                float s = 0.f;
                for (int p = 0; p < NPART; ++p) s += scr[p * HID * DA + idx];
                part[L::W2 + idx] = s;
            }
            __syncthreads();
            if (cq == 0 && lane == 0) {
#pragma unroll
                for (int d = 0; d < DA; ++d) scr[qd * 2 * DA + d] = gB2w[d], scr[qd * 2 * DA + DA + d] = gLSw[d];
            }
            __syncthreads();
            if (tid < 2 * DA) part[L::B2 + tid] = scr[tid] + scr[2 * DA + tid] + scr[4 * DA + tid] + scr[6 * DA + tid];   // b2 then log_std
        }
        const float v0 = warp_sum(s_obj), v1 = warp_sum(s_kl), v2 = warp_sum(s_ratio);   // held by cq == 0 threads, 0 elsewhere
        __syncthreads();
        if (lane == 0) S.red[warp] = v0, S.red[NW + warp] = v1, S.red[2 * NW + warp] = v2;
        __syncthreads();
        if (tid < 3) {
            float s = 0.f;
            for (int w = 0; w < NW; ++w) s += S.red[tid * NW + w];
            part[L::P + tid] = s;
        }
        __threadfence();
        __syncthreads();
        PCLK(14);
        const int c_lo = ts.cta_lo(m), c_hi = ts.cta_hi(m);
        if (tid == 0) S.last = (atomicAdd(A.counters + m, 1) == c_hi - c_lo);
        __syncthreads();
        PCLK(15);
        if (S.last) {
            __threadfence();
            // the trailing float4 of every partial slot holds the objective / KL / ratio sums: reduced by the same loop
            static_assert(L::P % 4 == 0 && PSTAT == 4, "stats ride on the float4 reduction");
            if (!want_grad && tid == 0 && A.stats) {
                const float4 s = reduce_segments4(A.partial, ts, A.kmax, PSTRIDE, m, c_lo, c_hi, L::P);
                A.stats[(int64_t)m * 4 + 0] = s.x * invN, A.stats[(int64_t)m * 4 + 1] = s.y * invN, A.stats[(int64_t)m * 4 + 2] = s.z * invN;
            }
            if (want_grad) {
                for (int p = 4 * tid; p < L::P + 4; p += 4 * TCT) {
                    const float4 s = reduce_segments4(A.partial, ts, A.kmax, PSTRIDE, m, c_lo, c_hi, p);
                    if (p == L::P) {
                        if (A.stats)
                            A.stats[(int64_t)m * 4 + 0] = s.x * invN, A.stats[(int64_t)m * 4 + 1] = s.y * invN,
                                                    A.stats[(int64_t)m * 4 + 2] = s.z * invN;
                        continue;
                    }
                    *reinterpret_cast<float4*>(A.grad + (int64_t)m * L::P + p) = s;
                    if (A.out_params) {
                        const float4 t4 = __ldg(reinterpret_cast<const float4*>(th + p));
                        *reinterpret_cast<float4*>(A.out_params + (int64_t)m * L::P + p) =
                            make_float4(t4.x - A.sgd_lr * s.x, t4.y - A.sgd_lr * s.y, t4.z - A.sgd_lr * s.z, t4.w - A.sgd_lr * s.w);
                    }
                }
            }
            if (tid == 0) A.counters[m] = 0;
        }
        __syncthreads();
    };

    int cur_m = -1;
    for (int g = ts.g_lo; g < ts.g_hi; ++g) {
        const int m = g / ts.ntiles, tile = g - m * ts.ntiles;
        PCLK(10);
        if (m != cur_m) {
            if (cur_m >= 0) flush(cur_m);
            PCLK(11);
            load_task(m, cur_m < 0);
            zero_acc();
            cur_m = m;
            PCLK(12);
        }
        const int n0 = tile * TBT, nb = max(0, min(TBT, Nm - n0));
        const int64_t g0 = (int64_t)m * N + n0;
        __syncthreads();
        for (int i = tid; i < TBT * DOP; i += TCT) {
            const int b = i / DOP, c = i % DOP;
            S.X[i] = (b < nb && c < DO) ? __ldg(A.obs + (g0 + b) * DO + c) : 0.f;
        }
        __syncthreads();
        PCLK(0);
        // ---- layer 0 (CUDA cores, row / column-group role): H1 = tanh(X W0 + b0) -> A0 (fp32 = TF32 "hi" operand) and LO
        {
            float x[DO];
#pragma unroll
            for (int i = 0; i < DO; ++i) x[i] = S.X[r * DOP + i];
#pragma unroll
            for (int c4 = 0; c4 < CW / 4; ++c4) {
                float h[4], hl[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int c = c0 + 4 * c4 + e;
                    float z = S.Ps[SL::B0 + c];
#pragma unroll
                    for (int i = 0; i < DO; ++i) z = fmaf(x[i], S.Ps[SL::W0 + i * HID + c], z);
                    h[e] = tanh_fast(z);
                    hl[e] = h[e] - tf32_trunc(h[e]);
                }
                const int off = core_off(r, c0 + 4 * c4, SCA);
                *reinterpret_cast<float4*>(S.A0 + off) = make_float4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<float4*>(S.LO + off) = make_float4(hl[0], hl[1], hl[2], hl[3]);
            }
        }
        PCLK(1);
        // ---- layer 1 on the tensor cores: Z2 = H1 W1 -> TMEM columns [0, 64)
        proxy_fence_async();
        tc_fence_before();
        __syncthreads();
        if (warp_u == 0 && elect_one()) {      // warp-uniform branch + elect: descriptors go straight to uniform registers
            tc_fence_after();
            issue_gemm_3xtf32(tmem, S.A0, S.LO, S.W1T_hi, S.W1T_lo);
            umma_commit(&S.bar);
            PCLK(13);
        }
        mbar_wait(&S.bar, phase);
        phase ^= 1;
        tc_fence_after();
        PCLK(2);
        float h2[CW];
        tmem_ldw(tmem_row + c0, h2);
        float mup[DA];
#pragma unroll
        for (int d = 0; d < DA; ++d) mup[d] = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < CW / 4; ++c4) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = 4 * c4 + e;
                h2[c] = tanh_fast(h2[c] + S.Ps[SL::B1 + c0 + c]);
#pragma unroll
                for (int d = 0; d < DA; ++d) mup[d] = fmaf(h2[c], S.Ps[SL::W2 + (c0 + c) * DA + d], mup[d]);
            }
            *reinterpret_cast<float4*>(S.A1 + core_off(r, c0 + 4 * c4, SCA)) =
                make_float4(h2[4 * c4], h2[4 * c4 + 1], h2[4 * c4 + 2], h2[4 * c4 + 3]);
        }
#pragma unroll
        for (int d = 0; d < DA; ++d) S.MUP[(cq * TBT + r) * DA + d] = mup[d];
        tc_fence_before();
        __syncthreads();
        PCLK(3);
        // ---- Gaussian head: one thread per sample row (cq == 0)
        if (cq == 0) {
            float dmu[DA], dls[DA];
            if (r < nb) {
                const int64_t n = g0 + r;
                float mu[DA], a[DA], mo[DA], lso[DA];
#pragma unroll
                for (int d = 0; d < DA; ++d) {
                    float sm = S.Ps[SL::B2 + d];
#pragma unroll
                    for (int q = 0; q < NQ; ++q) sm += S.MUP[(q * TBT + r) * DA + d];
                    mu[d] = sm;
                    a[d] = __ldg(A.act + n * DA + d);
                    mo[d] = __ldg(A.old_mean + n * DA + d);
                    lso[d] = A.ls_per_sample ? __ldg(A.old_ls + n * DA + d) : __ldg(A.old_ls + (int64_t)m * DA + d);
                }
                const float adv = __ldg(A.adv + n);
                HeadOut<DA> o;
                gaussian_head<DA>(hin, mu, a, mo, lso, adv, A.obj_kind, A.clip_eps, o);
                const float wt = A.obj_scale * o.w * invN, kc = A.kl_coeff * invN;
#pragma unroll
                for (int d = 0; d < DA; ++d) {
                    dmu[d] = wt * o.zeta[d] / hin.sig[d] + kc * o.dkl_dmu[d];
                    dls[d] = (wt * (o.zeta[d] * o.zeta[d] - 1.f) + kc * o.dkl_dls[d]) * hin.ls_mask[d];
                }
                s_obj += o.obj;
                s_kl += o.kl;
                s_ratio += o.ratio;
            } else {
#pragma unroll
                for (int d = 0; d < DA; ++d) dmu[d] = dls[d] = 0.f;
            }
#pragma unroll
            for (int d = 0; d < DA; ++d) {
                S.DMU[r * DA + d] = dmu[d];
                if (want_grad) {                                               // gB2 / g_log_std column sums
                    const float s1 = warp_sum(dmu[d]), s2 = warp_sum(dls[d]);
                    if (lane == 0) gB2w[d] += s1, gLSw[d] += s2;
                }
            }
        }
        if (!want_grad) continue;
        __syncthreads();
        PCLK(4);
        // ---- output-layer gradients (column role) from the fp32 H2 tile
        {
            const int b0 = cp * BPP;
#pragma unroll 4
            for (int bb = 0; bb < BPP; ++bb) {
                const int b = b0 + bb;
                const float h = *reinterpret_cast<const float*>(S.A1 + core_off(b, cj, SCA));
#pragma unroll
                for (int d = 0; d < DA; ++d) gW2p[d] = fmaf(h, S.DMU[b * DA + d], gW2p[d]);
            }
        }
        __syncthreads();
        PCLK(5);
        // ---- D2 = (DMU W2^T) * (1 - H2^2) from the h2 registers -> A1 (hi) / LO (lo)
        {
            float dm[DA];
#pragma unroll
            for (int d = 0; d < DA; ++d) dm[d] = S.DMU[r * DA + d];
#pragma unroll
            for (int c4 = 0; c4 < CW / 4; ++c4) {
                float v[4], vl[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int c = 4 * c4 + e;
                    float dh = 0.f;
#pragma unroll
                    for (int d = 0; d < DA; ++d) dh = fmaf(dm[d], S.Ps[SL::W2 + (c0 + c) * DA + d], dh);
                    v[e] = dh * (1.f - h2[c] * h2[c]);
                    vl[e] = v[e] - tf32_trunc(v[e]);
                }
                const int off = core_off(r, c0 + 4 * c4, SCA);
                *reinterpret_cast<float4*>(S.A1 + off) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(S.LO + off) = make_float4(vl[0], vl[1], vl[2], vl[3]);
            }
        }
        // ---- backward GEMM on the tensor cores: dH1 = D2 W1^T -> TMEM columns [64, 128) ...
        proxy_fence_async();
        tc_fence_before();
        __syncthreads();
        if (warp_u == 0 && elect_one()) {      // warp-uniform branch + elect: descriptors go straight to uniform registers
            tc_fence_after();
            issue_gemm_3xtf32(tmem + 64, S.A1, S.LO, S.W1_hi, S.W1_lo);
            umma_commit(&S.bar);
        }
        PCLK(6);
        // ---- ... while the warps do the weight gradient gW1 += H1^T D2 (mma.sync 3xTF32) and the bias column sums
        {
            wgrad_mma_tile<true, NT>(S.A0, S.A1, 1.f, warp, lane, gW1, gB1f);
        }
        PCLK(7);
        mbar_wait(&S.bar, phase);
        phase ^= 1;
        tc_fence_after();
        float dh1[CW];
        tmem_ldw(tmem_row + 64 + c0, dh1);
        tc_fence_before();
        __syncthreads();       // every CUDA-core read of H1 (weight gradient) is done before A0 is overwritten
        PCLK(8);
        // ---- D1 = dH1 * (1 - H1^2) -> A0 in place
#pragma unroll
        for (int c4 = 0; c4 < CW / 4; ++c4) {
            float4* p = reinterpret_cast<float4*>(S.A0 + core_off(r, c0 + 4 * c4, SCA));
            const float4 h = *p;
            *p = make_float4(dh1[4 * c4] * (1.f - h.x * h.x), dh1[4 * c4 + 1] * (1.f - h.y * h.y),
                             dh1[4 * c4 + 2] * (1.f - h.z * h.z), dh1[4 * c4 + 3] * (1.f - h.w * h.w));
        }
        __syncthreads();
        PCLK(9);
        // ---- gW0 += X^T D1, gB0 += colsum(D1) (column role)
        {
            const int b0 = cp * BPP;
#pragma unroll 4
            for (int bb = 0; bb < BPP; ++bb) {
                const int b = b0 + bb;
                const float d1 = *reinterpret_cast<const float*>(S.A0 + core_off(b, cj, SCA));
                gB0c += d1;
#pragma unroll
                for (int i = 0; i < DO; ++i) gW0p[i] = fmaf(S.X[b * DOP + i], d1, gW0p[i]);
            }
        }
    }
    PCLK(10);
    if (cur_m >= 0) flush(cur_m);
    PCLK(11);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(128));
#ifdef PROMP_EXP_CLOCKS
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int i = 0; i < 16; ++i) g_phase_clk[i] += s_clk[i];
#endif
}


// =================================================================================================================
// Tensor-core variant of policy_hvp_kernel (HID = 64).  All six layer GEMMs of the exact Hessian-vector product
//   forward :  Z2 = H1 W1            RZ2 = R1 W1 + H1 V1
//   backward:  dH1 = D2 W1^T         CdH1 = C2 W1^T + D2 (ac V1)^T
// run as tcgen05.mma.kind::tf32 with the 3-term split.  Shared memory cannot hold hi+lo copies of four activation
// tiles next to hi+lo copies of four weight tiles, so
//   * the "lo" A operands live in TENSOR MEMORY (tcgen05.st by the thread that owns the row; MMA with A from TMEM),
//   * the B (weight) buffer is time-multiplexed: [W1^T, V1^T] for the forward MMAs, re-filled with [W1, ac V1] for the
//     backward MMAs (66 KB from L2 twice per 128-row tile: ~1 % of the tile time).
// The two weight-gradient GEMMs (H1^T C2, R1^T D2) contract over samples (MN-major operands) and stay on the CUDA
// cores, overlapped with the backward MMAs.  TMEM map (512 columns allocated): Z2 0-63, RZ2 64-127, dH1 128-191,
// CdH1 192-255, lo-A 256-319, lo-B 320-383.
template <int DO, int DA, int NQ>
struct HvpTcSmem {
    static constexpr int DOP = DOPad<DO>::V;
    alignas(16) unsigned char WB[4][TILE_W_BYTES];    // fwd: W1T_hi, W1T_lo, V1T_hi, V1T_lo ; bwd: W1_hi, W1_lo, aV1_hi, aV1_lo
    alignas(16) unsigned char H1[TILE_A_BYTES];       // H1 -> C1
    alignas(16) unsigned char R1[TILE_A_BYTES];
    alignas(16) unsigned char T2a[TILE_A_BYTES];      // H2 -> D2
    alignas(16) unsigned char T2b[TILE_A_BYTES];      // R2 -> C2 ; flush scratch
    alignas(16) float Ps[SmallLayout<DO, DA>::SIZE];
    alignas(16) float Vs[SmallLayout<DO, DA>::SIZE];
    // X (observations) aliases T2a (needed only before H2 is written and, re-read from L2, after D2 is dead);
    // MUP (per-column-group partial means) aliases WB[0..1] between the forward MMAs and the backward weight re-fill.
    float DMU[TBT * DA];
    float CMU[TBT * DA];
    float red[3 * 4 * NQ];
    alignas(8) uint64_t bar;
    uint32_t tmem_base;
    int last;
};

__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// D (+)= A.B^T over K = 64 with the A "lo" part in TMEM: a_lo(TMEM).b_hi + a_hi(smem).b_lo + a_hi.b_hi
__device__ __forceinline__ void issue_gemm_3xtf32_ts(uint32_t d_tmem, const unsigned char* a_hi, uint32_t a_lo_tmem,
                                                     const unsigned char* b_hi, const unsigned char* b_lo, uint32_t acc) {
    const uint32_t idesc = umma_idesc_tf32(TBT, TC_HID);
    const uint32_t a0 = smem_u32(a_hi), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
#pragma unroll
    for (int s = 0; s < TC_HID / 8; ++s) {
        umma_tf32_ts(d_tmem, a_lo_tmem + 8 * s, umma_desc(bh + 2 * s * SCW, SCW, 128), idesc, acc);
        acc = 1;
    }
#pragma unroll
    for (int s = 0; s < TC_HID / 8; ++s)
        umma_tf32(d_tmem, umma_desc(a0 + 2 * s * SCA, SCA, 128), umma_desc(bl + 2 * s * SCW, SCW, 128), idesc, 1);
#pragma unroll
    for (int s = 0; s < TC_HID / 8; ++s)
        umma_tf32(d_tmem, umma_desc(a0 + 2 * s * SCA, SCA, 128), umma_desc(bh + 2 * s * SCW, SCW, 128), idesc, 1);
}

template <int DO, int DA, int NQ>
__global__ void __launch_bounds__(128 * NQ, 1) policy_hvp_tc_kernel(PolicyArgs A) {
    constexpr int HID = TC_HID;
    using L = PLayout<DO, DA, HID>;
    using SL = SmallLayout<DO, DA>;
    using SM = HvpTcSmem<DO, DA, NQ>;
    constexpr int TCT = 128 * NQ, CW = TC_HID / NQ, NW = TCT / 32, NT = 8 / NQ;
    constexpr int DOP = SM::DOP;
    constexpr int PSTRIDE = L::P + PSTAT;
    constexpr int NPART = TCT / HID, BPP = TBT / NPART;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    SM& S = *reinterpret_cast<SM*>(smem_raw);
    float* const sX = reinterpret_cast<float*>(S.T2a);
    float* const sMUP = reinterpret_cast<float*>(S.WB[0]);
    static_assert(TBT * DOP * 4 <= TILE_A_BYTES && NQ * TBT * 2 * DA * 4 <= 2 * TILE_W_BYTES, "aliased buffers must fit");

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);      // provably warp-uniform copy for the MMA-issue branch
    const int qd = warp & 3, cq = warp >> 2;               // TMEM lane quadrant, column group
    const int r = qd * 32 + lane, c0 = CW * cq;
    const int cj = tid & (HID - 1), cp = tid / HID;
    const TileSched ts(A.M, A.N, A.q, TBT);
    const int N = A.N;
    float invN = 1.0f / (float)N;       // both re-set per task when A.n_valid is given (variable-length paths)
    int Nm = N;
    const float ac = -A.inner_lr;
    const float* th = nullptr;
    const float* vg = nullptr;
    HeadIn<DA> hin;
    float rls[DA];
    uint32_t phase = 0;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&S.tmem_base)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (tid == 0) mbar_init(&S.bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;
    const uint32_t tmem_row = tmem + ((uint32_t)(qd * 32) << 16);
    constexpr uint32_t C_Z2 = 0, C_RZ2 = 64, C_DH1 = 128, C_CH1 = 192, C_LOA = 256, C_LOB = 320;

    float gW1[NT][4], gB1f[NT], gW0p[DO], gW2p[DA], gB0c, gB2w[DA], gLSw[DA];
    float s_obj, s_kl, s_ratio;
    auto zero_acc = [&]() {
#pragma unroll
        for (int a = 0; a < NT; ++a) gW1[a][0] = gW1[a][1] = gW1[a][2] = gW1[a][3] = gB1f[a] = 0.f;
#pragma unroll
        for (int i = 0; i < DO; ++i) gW0p[i] = 0.f;
#pragma unroll
        for (int d = 0; d < DA; ++d) gW2p[d] = 0.f;
#pragma unroll
        for (int d = 0; d < DA; ++d) gB2w[d] = gLSw[d] = 0.f;
        gB0c = 0.f;
        s_obj = s_kl = s_ratio = 0.f;
    };
    auto load_task = [&](int m, bool first) {
        if (A.n_valid) { Nm = __ldg(A.n_valid + m); invN = 1.0f / (float)max(Nm, 1); }
        th = A.params + (int64_t)m * A.param_stride;
        vg = A.vec + (int64_t)m * L::P;
        __syncthreads();
        const bool reload_p = first || A.param_stride != 0;
        for (int i = tid; i < DO * HID + HID; i += TCT) {
            if (reload_p) S.Ps[SL::W0 + i] = __ldg(th + L::W0 + i);
            S.Vs[SL::W0 + i] = __ldcg(vg + L::W0 + i);
        }
        for (int i = tid; i < HID; i += TCT) {
            if (reload_p) S.Ps[SL::B1 + i] = __ldg(th + L::B1 + i);
            S.Vs[SL::B1 + i] = __ldcg(vg + L::B1 + i);
        }
        for (int i = tid; i < HID * DA + 2 * DA; i += TCT) {
            if (reload_p) S.Ps[SL::W2 + i] = __ldg(th + L::W2 + i);
            S.Vs[SL::W2 + i] = __ldcg(vg + L::W2 + i);
        }
        __syncthreads();
#pragma unroll
        for (int d = 0; d < DA; ++d) {
            const float raw = S.Ps[SL::LS + d];
            const bool clipped = A.clip_log_std && (raw < A.min_log_std);
            hin.ls[d] = clipped ? A.min_log_std : raw;
            hin.ls_mask[d] = clipped ? 0.f : 1.f;
            hin.sig[d] = expf(hin.ls[d]);
            rls[d] = S.Vs[SL::LS + d] * hin.ls_mask[d];
        }
    };
    // (re)fill the weight buffer from L2: forward = [W1^T, V1^T], backward = [W1, ac*V1], each as hi (fp32) + lo
    auto load_weights = [&](bool fwd) {
        for (int i = tid; i < HID * HID; i += TCT) {
            const int k = i / HID, j = i % HID;
            const float w = __ldg(th + L::W1 + i);
            const float v = (fwd ? 1.f : ac) * __ldcg(vg + L::W1 + i);
            const int off = fwd ? core_off(j, k, SCW) : core_off(k, j, SCW);
            *reinterpret_cast<float*>(S.WB[0] + off) = w;
            *reinterpret_cast<float*>(S.WB[1] + off) = w - tf32_trunc(w);
            *reinterpret_cast<float*>(S.WB[2] + off) = v;
            *reinterpret_cast<float*>(S.WB[3] + off) = v - tf32_trunc(v);
        }
    };
    auto flush = [&](int m) {
        float* part = A.partial + ((int64_t)blockIdx.x * A.kmax + (m - ts.first_task(blockIdx.x))) * PSTRIDE;
        float* scr = reinterpret_cast<float*>(S.T2a);     // T2a + T2b (contiguous, 2 tiles)
        static_assert(NPART * DO * HID * 4 <= 2 * TILE_A_BYTES && NPART * HID * DA * 4 <= 2 * TILE_A_BYTES, "flush scratch");
        __syncthreads();
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) part[L::W1 + wgrad_mma_index<NT>(warp, lane, nt, i)] = gW1[nt][i];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            float c = gB1f[nt];
            c += __shfl_xor_sync(0xffffffffu, c, 1);
            c += __shfl_xor_sync(0xffffffffu, c, 2);
            if ((lane & 3) == 0 && (warp & 3) == 0) part[L::B1 + 8 * NT * (warp >> 2) + 8 * nt + (lane >> 2)] = c;
        }
        scr[cp * HID + cj] = gB0c;
        __syncthreads();
        if (tid < HID) {
            float s = 0.f;
            for (int p = 0; p < NPART; ++p) s += scr[p * HID + tid];
            part[L::B0 + tid] = s;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < DO; ++i) scr[(cp * DO + i) * HID + cj] = gW0p[i];
        __syncthreads();
        for (int idx = tid; idx < DO * HID; idx += TCT) {
            float s = 0.f;
            for (int p = 0; p < NPART; ++p) s += scr[p * DO * HID + idx];
            part[L::W0 + idx] = s;
        }
        __syncthreads();
#pragma unroll
        for (int d = 0; d < DA; ++d) scr[(cp * HID + cj) * DA + d] = gW2p[d];
        __syncthreads();
        for (int idx = tid; idx < HID * DA; idx += TCT) {
            float s = 0.f;
            for (int p = 0; p < NPART; ++p) s += scr[p * HID * DA + idx];
            part[L::W2 + idx] = s;
        }
        __syncthreads();
        if (cq == 0 && lane == 0) {
#pragma unroll
            for (int d = 0; d < DA; ++d) scr[qd * 2 * DA + d] = gB2w[d], scr[qd * 2 * DA + DA + d] = gLSw[d];
        }
        __syncthreads();
        if (tid < 2 * DA) part[L::B2 + tid] = scr[tid] + scr[2 * DA + tid] + scr[4 * DA + tid] + scr[6 * DA + tid];
        const float v0 = warp_sum(s_obj), v1 = warp_sum(s_kl), v2 = warp_sum(s_ratio);
        __syncthreads();
        if (lane == 0) S.red[warp] = v0, S.red[NW + warp] = v1, S.red[2 * NW + warp] = v2;
        __syncthreads();
        if (tid < 3) {
            float s = 0.f;
            for (int w = 0; w < NW; ++w) s += S.red[tid * NW + w];
            part[L::P + tid] = s;
        }
        __threadfence();
        __syncthreads();
        const int c_lo = ts.cta_lo(m), c_hi = ts.cta_hi(m);
        if (tid == 0) S.last = (atomicAdd(A.counters + m, 1) == c_hi - c_lo);
        __syncthreads();
        if (S.last) {
            __threadfence();
            for (int p = 4 * tid; p < L::P + 4; p += 4 * TCT) {
                const float4 s = reduce_segments4(A.partial, ts, A.kmax, PSTRIDE, m, c_lo, c_hi, p);
                if (p == L::P) {                  // trailing float4 of the slot: objective / KL / ratio sums
                    if (A.stats)
                        A.stats[(int64_t)m * 4 + 0] = s.x * invN, A.stats[(int64_t)m * 4 + 1] = s.y * invN,
                                                A.stats[(int64_t)m * 4 + 2] = s.z * invN;
                    continue;
                }
                const float4 v4 = __ldcg(reinterpret_cast<const float4*>(vg + p));
                *reinterpret_cast<float4*>(A.out + (int64_t)m * L::P + p) = make_float4(v4.x + s.x, v4.y + s.y, v4.z + s.z, v4.w + s.w);
            }
            if (tid == 0) A.counters[m] = 0;
        }
        __syncthreads();
    };

    int cur_m = -1;
    for (int g = ts.g_lo; g < ts.g_hi; ++g) {
        const int m = g / ts.ntiles, tile = g - m * ts.ntiles;
        if (m != cur_m) {
            if (cur_m >= 0) flush(cur_m);
            load_task(m, cur_m < 0);
            zero_acc();
            cur_m = m;
        }
        const int n0 = tile * TBT, nb = max(0, min(TBT, Nm - n0));
        const int64_t g0 = (int64_t)m * N + n0;
        __syncthreads();
        auto load_x = [&]() {
            for (int i = tid; i < TBT * DOP; i += TCT) {
                const int b = i / DOP, c = i % DOP;
                sX[i] = (b < nb && c < DO) ? __ldg(A.obs + (g0 + b) * DO + c) : 0.f;
            }
        };
        load_x();
        load_weights(true);
        __syncthreads();
        // ---- layer 0 and its tangent (CUDA cores, row / column-group role); lo parts of H1 / R1 -> TMEM
        {
            float x[DO], hl[CW], rl[CW];
#pragma unroll
            for (int i = 0; i < DO; ++i) x[i] = sX[r * DOP + i];
#pragma unroll
            for (int c4 = 0; c4 < CW / 4; ++c4) {
                float h[4], r1[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int c = c0 + 4 * c4 + e;
                    float z = S.Ps[SL::B0 + c], rz = S.Vs[SL::B0 + c];
#pragma unroll
                    for (int i = 0; i < DO; ++i) {
                        z = fmaf(x[i], S.Ps[SL::W0 + i * HID + c], z);
                        rz = fmaf(x[i], S.Vs[SL::W0 + i * HID + c], rz);
                    }
                    h[e] = tanh_fast(z);
                    r1[e] = (1.f - h[e] * h[e]) * rz;
                    hl[4 * c4 + e] = h[e] - tf32_trunc(h[e]);
                    rl[4 * c4 + e] = r1[e] - tf32_trunc(r1[e]);
                }
                const int off = core_off(r, c0 + 4 * c4, SCA);
                *reinterpret_cast<float4*>(S.H1 + off) = make_float4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<float4*>(S.R1 + off) = make_float4(r1[0], r1[1], r1[2], r1[3]);
            }
            tmem_stw(tmem_row + C_LOA + c0, hl);
            tmem_stw(tmem_row + C_LOB + c0, rl);
        }
        // ---- forward MMAs: Z2 = H1 W1 ; RZ2 = R1 W1 + H1 V1
        proxy_fence_async();
        tc_fence_before();
        __syncthreads();
        if (warp_u == 0 && elect_one()) {      // warp-uniform branch + elect: descriptors go straight to uniform registers
            tc_fence_after();
            issue_gemm_3xtf32_ts(tmem + C_Z2, S.H1, tmem + C_LOA, S.WB[0], S.WB[1], 0);
            issue_gemm_3xtf32_ts(tmem + C_RZ2, S.R1, tmem + C_LOB, S.WB[0], S.WB[1], 0);
            issue_gemm_3xtf32_ts(tmem + C_RZ2, S.H1, tmem + C_LOA, S.WB[2], S.WB[3], 1);
            umma_commit(&S.bar);
        }
        mbar_wait(&S.bar, phase);
        phase ^= 1;
        tc_fence_after();
        float h2[CW], r2[CW];
        tmem_ldw(tmem_row + C_Z2 + c0, h2);
        tmem_ldw(tmem_row + C_RZ2 + c0, r2);
        {
            float mup[DA], rmup[DA];
#pragma unroll
            for (int d = 0; d < DA; ++d) mup[d] = rmup[d] = 0.f;
#pragma unroll
            for (int c4 = 0; c4 < CW / 4; ++c4) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int c = 4 * c4 + e;
                    h2[c] = tanh_fast(h2[c] + S.Ps[SL::B1 + c0 + c]);
                    r2[c] = (1.f - h2[c] * h2[c]) * (r2[c] + S.Vs[SL::B1 + c0 + c]);
#pragma unroll
                    for (int d = 0; d < DA; ++d) {
                        const float w2 = S.Ps[SL::W2 + (c0 + c) * DA + d];
                        mup[d] = fmaf(h2[c], w2, mup[d]);
                        rmup[d] = fmaf(r2[c], w2, fmaf(h2[c], S.Vs[SL::W2 + (c0 + c) * DA + d], rmup[d]));
                    }
                }
                const int off = core_off(r, c0 + 4 * c4, SCA);
                *reinterpret_cast<float4*>(S.T2a + off) = make_float4(h2[4 * c4], h2[4 * c4 + 1], h2[4 * c4 + 2], h2[4 * c4 + 3]);
                *reinterpret_cast<float4*>(S.T2b + off) = make_float4(r2[4 * c4], r2[4 * c4 + 1], r2[4 * c4 + 2], r2[4 * c4 + 3]);
            }
#pragma unroll
            for (int d = 0; d < DA; ++d) {
                sMUP[((cq * TBT + r) * 2 + 0) * DA + d] = mup[d];
                sMUP[((cq * TBT + r) * 2 + 1) * DA + d] = rmup[d];
            }
        }
        tc_fence_before();
        __syncthreads();
        // ---- Gaussian head and its tangent: one thread per sample row (cq == 0)
        if (cq == 0) {
            float dmu[DA], cmu[DA], cls[DA];
            if (r < nb) {
                const int64_t n = g0 + r;
                float mu[DA], rmu[DA], a[DA], mo[DA], lso[DA];
#pragma unroll
                for (int d = 0; d < DA; ++d) {
float sm = S.Ps[SL::B2 + d], sr = S.Vs[SL::B2 + d];
#pragma unroll
                    for (int q = 0; q < NQ; ++q) sm += sMUP[((q * TBT + r) * 2 + 0) * DA + d], sr += sMUP[((q * TBT + r) * 2 + 1) * DA + d];
                    mu[d] = sm;
                    rmu[d] = sr;
                    a[d] = __ldg(A.act + n * DA + d);
                    mo[d] = __ldg(A.old_mean + n * DA + d);
                    lso[d] = A.ls_per_sample ? __ldg(A.old_ls + n * DA + d) : __ldg(A.old_ls + (int64_t)m * DA + d);
                }
                const float adv = __ldg(A.adv + n);
                HeadOut<DA> o;
                gaussian_head<DA>(hin, mu, a, mo, lso, adv, A.obj_kind, A.clip_eps, o);
                const float wt = o.w * invN, kc = A.kl_coeff * invN;
                float rl_ = 0.f;
#pragma unroll
                for (int d = 0; d < DA; ++d)
                    rl_ += (o.zeta[d] / hin.sig[d]) * rmu[d] + (o.zeta[d] * o.zeta[d] - 1.f) * rls[d];
                const float rwt = (A.obj_kind == PROMP_OBJ_RATIO) ? wt * rl_ : 0.f;
#pragma unroll
                for (int d = 0; d < DA; ++d) {
                    const float is = 1.f / hin.sig[d], z = o.zeta[d];
                    const float rz = -rmu[d] * is - z * rls[d];
                    dmu[d] = wt * z * is;
                    const float rdmu = rwt * z * is + wt * (rz * is - z * rls[d] * is);
                    const float rdls = rwt * (z * z - 1.f) + wt * 2.f * z * rz;
                    cmu[d] = ac * rdmu + kc * o.dkl_dmu[d];
                    cls[d] = (ac * rdls + kc * o.dkl_dls[d]) * hin.ls_mask[d];
                }
                s_obj += o.obj;
                s_kl += o.kl;
                s_ratio += o.ratio;
            } else {
#pragma unroll
                for (int d = 0; d < DA; ++d) dmu[d] = cmu[d] = cls[d] = 0.f;
            }
#pragma unroll
            for (int d = 0; d < DA; ++d) {
                S.DMU[r * DA + d] = dmu[d], S.CMU[r * DA + d] = cmu[d];
                const float s1 = warp_sum(cmu[d]), s2 = warp_sum(cls[d]);
                if (lane == 0) gB2w[d] += s1, gLSw[d] += s2;
            }
        }
        __syncthreads();
        // the forward MMAs are complete and MUP is consumed: re-fill the weight buffer for the backward MMAs
        load_weights(false);
        // ---- output layer (column role): out_W2 += H2^T CMU + ac R2^T DMU ; out_b2 += colsum CMU ; out_ls += colsum CLS
        {
            const int b0 = cp * BPP;
#pragma unroll 4
            for (int bb = 0; bb < BPP; ++bb) {
                const int b = b0 + bb;
                const int off = core_off(b, cj, SCA);
                const float h = *reinterpret_cast<const float*>(S.T2a + off);
                const float rr = ac * *reinterpret_cast<const float*>(S.T2b + off);
#pragma unroll
                for (int d = 0; d < DA; ++d) gW2p[d] = fmaf(h, S.CMU[b * DA + d], fmaf(rr, S.DMU[b * DA + d], gW2p[d]));
            }
        }
        __syncthreads();
        // ---- D2 = dH2 g2 -> T2a ; C2 = CdH2 g2 + ac dH2 (-2 H2 R2) -> T2b ; lo parts -> TMEM
        {
            float dm[DA], cm[DA], dl[CW], cl[CW];
#pragma unroll
            for (int d = 0; d < DA; ++d) dm[d] = S.DMU[r * DA + d], cm[d] = S.CMU[r * DA + d];
#pragma unroll
            for (int c4 = 0; c4 < CW / 4; ++c4) {
                float d2[4], c2[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int c = 4 * c4 + e;
                    float dh = 0.f, ch = 0.f;
#pragma unroll
                    for (int d = 0; d < DA; ++d) {
                        const float w2 = S.Ps[SL::W2 + (c0 + c) * DA + d], v2 = S.Vs[SL::W2 + (c0 + c) * DA + d];
                        dh = fmaf(dm[d], w2, dh);
                        ch = fmaf(cm[d], w2, fmaf(ac * dm[d], v2, ch));
                    }
                    const float g2 = 1.f - h2[c] * h2[c];
                    d2[e] = dh * g2;
                    c2[e] = ch * g2 + ac * dh * (-2.f * h2[c] * r2[c]);
                    dl[c] = d2[e] - tf32_trunc(d2[e]);
                    cl[c] = c2[e] - tf32_trunc(c2[e]);
                }
                const int off = core_off(r, c0 + 4 * c4, SCA);
                *reinterpret_cast<float4*>(S.T2a + off) = make_float4(d2[0], d2[1], d2[2], d2[3]);
                *reinterpret_cast<float4*>(S.T2b + off) = make_float4(c2[0], c2[1], c2[2], c2[3]);
            }
            tmem_stw(tmem_row + C_LOA + c0, dl);
            tmem_stw(tmem_row + C_LOB + c0, cl);
        }
        // ---- backward MMAs: dH1 = D2 W1^T ; CdH1 = C2 W1^T + D2 (ac V1)^T ...
        proxy_fence_async();
        tc_fence_before();
        __syncthreads();
        if (warp_u == 0 && elect_one()) {      // warp-uniform branch + elect: descriptors go straight to uniform registers
            tc_fence_after();
            issue_gemm_3xtf32_ts(tmem + C_DH1, S.T2a, tmem + C_LOA, S.WB[0], S.WB[1], 0);
            issue_gemm_3xtf32_ts(tmem + C_CH1, S.T2b, tmem + C_LOB, S.WB[0], S.WB[1], 0);
            issue_gemm_3xtf32_ts(tmem + C_CH1, S.T2a, tmem + C_LOA, S.WB[2], S.WB[3], 1);
            umma_commit(&S.bar);
        }
        // ---- ... overlapped with the weight gradients out_W1 += H1^T C2 + R1^T (ac D2) (mma.sync 3xTF32) and colsum(C2)
        {
            float unused[NT] = {};
            wgrad_mma_tile<true, NT>(S.H1, S.T2b, 1.f, warp, lane, gW1, gB1f);
            wgrad_mma_tile<false, NT>(S.R1, S.T2a, ac, warp, lane, gW1, unused);
        }
        mbar_wait(&S.bar, phase);
        phase ^= 1;
        tc_fence_after();
        float dh1[CW], ch1[CW];
        tmem_ldw(tmem_row + C_DH1 + c0, dh1);
        tmem_ldw(tmem_row + C_CH1 + c0, ch1);
        tc_fence_before();
        __syncthreads();       // all CUDA-core reads of H1 / R1 are done before H1 is overwritten
        // ---- C1 = CdH1 g1 + ac dH1 (-2 H1 R1) -> H1 in place
#pragma unroll
        for (int c4 = 0; c4 < CW / 4; ++c4) {
            const int off = core_off(r, c0 + 4 * c4, SCA);
            float4* p = reinterpret_cast<float4*>(S.H1 + off);
            const float4 h = *p;
            const float4 rr = *reinterpret_cast<const float4*>(S.R1 + off);
            const float hv[4] = {h.x, h.y, h.z, h.w}, rv[4] = {rr.x, rr.y, rr.z, rr.w};
            float c1[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
                c1[e] = ch1[4 * c4 + e] * (1.f - hv[e] * hv[e]) + ac * dh1[4 * c4 + e] * (-2.f * hv[e] * rv[e]);
            *p = make_float4(c1[0], c1[1], c1[2], c1[3]);
        }
        load_x();              // D2 (T2a) is dead: bring the observations back for the input-layer gradient
        __syncthreads();
        // ---- out_W0 += X^T C1 ; out_b0 += colsum C1 (column role)
        {
            const int b0 = cp * BPP;
#pragma unroll 4
            for (int bb = 0; bb < BPP; ++bb) {
                const int b = b0 + bb;
                const float c1 = *reinterpret_cast<const float*>(S.H1 + core_off(b, cj, SCA));
                gB0c += c1;
#pragma unroll
                for (int i = 0; i < DO; ++i) gW0p[i] = fmaf(sX[b * DOP + i], c1, gW0p[i]);
            }
        }
    }
    if (cur_m >= 0) flush(cur_m);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512));
}

}  // namespace promp
