// Gaussian-MLP policy kernels: per-task objective gradient (inner adapt step, outer surrogates, KL
// terms) and the exact Hessian-vector product that carries the MAML meta-gradient through the inner
// SGD step.  See include/promp_b200.h for the interface and the reference functions replaced.
//
// Work decomposition: persistent CTAs, one wave.  The M*ceil(N/64) sample tiles form one task-major
// list; CTA c owns the contiguous range [c*q, (c+1)*q) (q = ceil(T/grid), grid = SMs x resident
// CTAs), so every SM gets the same number of tiles whatever M and N are.  A CTA keeps the current
// task's weights in shared memory and its weight-gradient accumulators in registers; when its range
// crosses a task boundary (or ends) it flushes them to a per-(CTA,task) partial slot, and the last
// segment of each task to arrive (atomic ticket) reduces that task's slots in CTA order -> bitwise
// run-to-run deterministic sums.  These kernels are fp32-FMA bound (AI ~ 10^2..10^3 FLOP/B).
#include <string.h>
#include "mlp_tile.cuh"

namespace promp {

template <int DO>
struct DOPad {
    static constexpr int V = (DO + 3) / 4 * 4;
};

constexpr int PSTAT = 4;   // per-partial trailing stats: sum obj, sum kl, sum ratio, unused

struct PolicyArgs {
    int M, N;
    const float* params;
    int64_t param_stride;
    const float *obs, *act, *adv, *old_mean, *old_ls;
    int ls_per_sample;
    int obj_kind;
    float obj_scale, clip_eps, kl_coeff;
    int clip_log_std;
    float min_log_std;
    // grad kernel
    float* grad;
    float* out_params;
    float sgd_lr;
    // hvp kernel
    const float* vec;
    float* out;
    float inner_lr;
    float* stats;
    float* partial;      // [grid][kmax][P + PSTAT] per-(CTA, task-segment) partial sums
    int* counters;       // [M], zero on entry, left zero on exit
    int q;               // tiles per CTA
    int kmax;            // max task segments per CTA
    const int32_t* n_valid;   // [M] valid samples per task (rows >= n_valid[m] are padding) or nullptr = N everywhere
    // Re-use of an identical earlier launch (grad kernels only): the inner pass of the first Adam epoch repeats MAMLAlgo._adapt
    // (same theta, same phase-0 data, same outputs) unless the reported-log_std clip of the step-0 graph is active.
    //   producer side: unclipped_out = 1 iff every log_std component >= min_log_std, theta_copy_out = the parameters it used
    //   consumer side: the whole grid returns at once if *skip_flag != 0 and params == skip_theta bit for bit (its outputs
    //                  alias the producer's, which are then already correct)
    const int* skip_flag;
    const float* skip_theta;
    int* unclipped_out;
    float* theta_copy_out;
    // optional device-resident multiplier of kl_coeff (ProMP's adaptive inner-KL coefficient lives on the device so that an
    // iteration has no host decision: promp_adapt_kl_coeff updates it between launches)
    const float* kl_coeff_ptr;
};
__device__ __forceinline__ float kl_coeff_eff(const PolicyArgs& A) {
    return A.kl_coeff_ptr ? A.kl_coeff * __ldcg(A.kl_coeff_ptr) : A.kl_coeff;
}

// consumer / producer halves of the launch re-use protocol above; returns true if the calling CTA must exit
template <int P, int LS, int DA>
__device__ __forceinline__ bool grad_reuse_prologue(const PolicyArgs& A) {
    if (A.skip_flag) {
        bool same = *reinterpret_cast<const volatile int*>(A.skip_flag) != 0;
        for (int i = threadIdx.x; i < P && same; i += blockDim.x)
            same = __float_as_uint(__ldcg(A.params + i)) == __float_as_uint(__ldcg(A.skip_theta + i));
        if (__syncthreads_and(same ? 1 : 0)) return true;
    }
    if (A.unclipped_out && blockIdx.x == 0) {
        if (threadIdx.x == 0) {
            int ok = 1;
            for (int d = 0; d < DA; ++d)
                if (!(__ldcg(A.params + LS + d) >= A.min_log_std)) ok = 0;
            *A.unclipped_out = ok;
        }
        for (int i = threadIdx.x; i < P; i += blockDim.x) A.theta_copy_out[i] = __ldcg(A.params + i);
    }
    return false;
}

// Tile-range bookkeeping shared by both kernels.
struct TileSched {
    int ntiles, T, q, g_lo, g_hi;
    __device__ __forceinline__ TileSched(int M, int N, int q_, int tb = TB) {
        ntiles = (N + tb - 1) / tb;
        T = M * ntiles;
        q = q_;
        g_lo = blockIdx.x * q;
        g_hi = min(g_lo + q, T);
    }
    __device__ __forceinline__ int first_task(int c) const { return (c * q) / ntiles; }
    __device__ __forceinline__ int cta_lo(int m) const { return (m * ntiles) / q; }
    __device__ __forceinline__ int cta_hi(int m) const { return ((m + 1) * ntiles - 1) / q; }
};


// Sum one float4 column of a task's partial slots over its CTA segments [c_lo, c_hi] in CTA order, with eight
// independent L2 loads in flight (the naive dependent loop costs ~50 us of serial latency at the kernel tail).
__device__ __forceinline__ float4 reduce_segments4(const float* partial, const TileSched& ts, int kmax, int pstride, int m,
                                                   int c_lo, int c_hi, int p) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c0 = c_lo; c0 <= c_hi; c0 += 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int c = c0 + u;
            v[u] = (c <= c_hi) ? __ldcg(reinterpret_cast<const float4*>(
                                     partial + ((int64_t)c * kmax + (m - ts.first_task(c))) * pstride + p))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) s.x += v[u].x, s.y += v[u].y, s.z += v[u].z, s.w += v[u].w;
    }
    return s;
}

template <int DO, int DA, int HID>
__device__ __forceinline__ void load_head_consts(const float* P, int clip, float min_ls, HeadIn<DA>& hin) {
    using L = PLayout<DO, DA, HID>;
#pragma unroll
    for (int d = 0; d < DA; ++d) {
        const float raw = P[L::LS + d];
        const bool clipped = clip && (raw < min_ls);      // tf.maximum: gradient goes to x when x >= y
        hin.ls[d] = clipped ? min_ls : raw;
        hin.ls_mask[d] = clipped ? 0.f : 1.f;
        hin.sig[d] = expf(hin.ls[d]);
    }
    head_in_finish<DA>(hin);
}

// -------------------------------------------------------------------------------------------------
// Thread roles inside a 256-thread CTA working on a 64-sample tile:
//   GEMM role      (tx, ty): 4 (HID=64) or 2 (HID=32) rows x 4 columns of every [64 x HID] product
//   row role       (rb, rq): 4 threads per sample row, each owns a quarter of the hidden units (layer 2 + head)
//   column role    (cj, cp): thread owns hidden unit cj for the sample slice cp (small weight gradients)
// Small reductions (bias / W0 / W2 gradients) are accumulated per thread in registers across tiles and
// only combined through shared memory when the CTA flushes a task segment.
template <int HID>
struct RoleCfg {
    static constexpr int NPART = PT_THREADS / HID;    // sample slices for the column role (4 | 8)
    static constexpr int BPP = TB / NPART;            // samples per slice (16 | 8)
    static constexpr int QW = HID / 4;                // hidden units per row-role thread (16 | 8)
};

template <int DO, int DA, int HID>
struct GradSmem {
    using L = PLayout<DO, DA, HID>;
    using C = TileCfg<HID>;
    static constexpr int DOP = DOPad<DO>::V;
    static constexpr int PP = (L::P + 3) / 4 * 4;
    float P[PP];
    float W1T[HID * HID];
    float X[TB * DOP];
    float H1[TB * C::LD];     // also: flush scratch
    float H2[TB * C::LD];
    float DMU[TB * DA];
    float DLS[TB * DA];
    float red[3 * (PT_THREADS / 32)];
    int last;
};

// Combine per-thread partial sums through shared memory: out[idx] = sum_p scratch[p][idx], idx < n.
// `mine(i)` yields this thread's partial for its i-th element; thread `tid` holds elements idx = base(tid) .. ; generic
// helper used only at flush time (once per task segment), so simplicity beats speed here.
template <int DO, int DA, int HID>
__global__ void __launch_bounds__(PT_THREADS, 2) policy_grad_kernel(PolicyArgs A) {
    using L = PLayout<DO, DA, HID>;
    using C = TileCfg<HID>;
    using R = RoleCfg<HID>;
    using SM = GradSmem<DO, DA, HID>;
    constexpr int LD = C::LD, RM = C::RM, RK = C::RK, DOP = SM::DOP;
    constexpr int NPART = R::NPART, BPP = R::BPP, QW = R::QW;
    constexpr int PSTRIDE = L::P + PSTAT;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    SM& S = *reinterpret_cast<SM*>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = tid % C::TX, ty = tid / C::TX;
    const int row0 = ty * RM, col0 = tx * 4;
    const int rb = tid >> 2, rq = tid & 3;           // row role
    const int cj = tid % HID, cp = tid / HID;        // column role
    if (grad_reuse_prologue<L::P, L::LS, DA>(A)) return;
    const TileSched ts(A.M, A.N, A.q);
    const int N = A.N;
    const float kl_eff = kl_coeff_eff(A);
    float invN = 1.0f / (float)N;       // both re-set per task when A.n_valid is given (variable-length paths)
    int Nm = N;
    const bool want_grad = A.grad != nullptr;
    const float* th = nullptr;
    HeadIn<DA> hin;

    // register accumulators (per task segment)
    float gW1[RK][4];        // GEMM role: H1^T D2 block
    float gB1p[4], gB0p[4];  // GEMM role: column sums over this thread's rows
    float gW0p[DO], gW2p[DA];   // column role: X^T D1 / H2^T DMU for unit cj over sample slice cp
    float gB2 = 0.f, gLS = 0.f; // thread tid < DA
    float s_obj, s_kl, s_ratio; // row role, rq == 0
    auto zero_acc = [&]() {
#pragma unroll
        for (int r = 0; r < RK; ++r) gW1[r][0] = gW1[r][1] = gW1[r][2] = gW1[r][3] = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) gB1p[c] = gB0p[c] = 0.f;
#pragma unroll
        for (int i = 0; i < DO; ++i) gW0p[i] = 0.f;
#pragma unroll
        for (int d = 0; d < DA; ++d) gW2p[d] = 0.f;
        gB2 = gLS = 0.f;
        s_obj = s_kl = s_ratio = 0.f;
    };
    auto load_task = [&](int m, bool first) {
        if (A.n_valid) { Nm = __ldg(A.n_valid + m); invN = 1.0f / (float)max(Nm, 1); }
        th = A.params + (int64_t)m * A.param_stride;
        if (!first && A.param_stride == 0) return;       // shared theta: weights already resident
        __syncthreads();
        for (int i = tid; i < L::P; i += PT_THREADS) S.P[i] = __ldg(th + i);
        __syncthreads();
        for (int i = tid; i < HID * HID; i += PT_THREADS) {
            const int j = i / HID, k = i % HID;      // consecutive threads -> consecutive W1T addresses
            S.W1T[j * HID + k] = S.P[L::W1 + k * HID + j];
        }
        load_head_consts<DO, DA, HID>(S.P, A.clip_log_std, A.min_log_std, hin);
    };
    // write this CTA's partial sums for task m; the last segment of the task reduces them in CTA order
    auto flush = [&](int m) {
        float* part = A.partial + ((int64_t)blockIdx.x * A.kmax + (m - ts.first_task(blockIdx.x))) * PSTRIDE;
        float* scr = S.H1;      // free between tiles
        __syncthreads();
        if (want_grad) {
#pragma unroll
            for (int r = 0; r < RK; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) part[L::W1 + (ty * RK + r) * HID + col0 + c] = gW1[r][c];
            // column sums: reduce the C::TY row groups
#pragma unroll
            for (int c = 0; c < 4; ++c) scr[ty * HID + col0 + c] = gB1p[c];
            __syncthreads();
            if (tid < HID) {
                float s = 0.f;
                for (int y = 0; y < C::TY; ++y) s += scr[y * HID + tid];
                part[L::B1 + tid] = s;
            }
            __syncthreads();
#pragma unroll
            for (int c = 0; c < 4; ++c) scr[ty * HID + col0 + c] = gB0p[c];
            __syncthreads();
            if (tid < HID) {
                float s = 0.f;
                for (int y = 0; y < C::TY; ++y) s += scr[y * HID + tid];
                part[L::B0 + tid] = s;
            }
            __syncthreads();
            // W0 / W2 gradients: reduce the NPART sample slices
#pragma unroll
            for (int i = 0; i < DO; ++i) scr[(cp * DO + i) * HID + cj] = gW0p[i];
            __syncthreads();
            for (int idx = tid; idx < DO * HID; idx += PT_THREADS) {
                float s = 0.f;
                for (int p = 0; p < NPART; ++p) s += scr[p * DO * HID + idx];
                part[L::W0 + idx] = s;
            }
            __syncthreads();
#pragma unroll
            for (int d = 0; d < DA; ++d) scr[(cp * HID + cj) * DA + d] = gW2p[d];
            __syncthreads();
            for (int idx = tid; idx < HID * DA; idx += PT_THREADS) {
                float s = 0.f;
                for (int p = 0; p < NPART; ++p) s += scr[p * HID * DA + idx];
                part[L::W2 + idx] = s;
            }
            if (tid < DA) part[L::B2 + tid] = gB2, part[L::LS + tid] = gLS;
        }
        // head statistics (held by the rq == 0 threads)
        const float v0 = warp_sum(rq == 0 ? s_obj : 0.f), v1 = warp_sum(rq == 0 ? s_kl : 0.f),
                    v2 = warp_sum(rq == 0 ? s_ratio : 0.f);
        __syncthreads();
        if (lane == 0) S.red[warp] = v0, S.red[8 + warp] = v1, S.red[16 + warp] = v2;
        __syncthreads();
        if (tid < 3) {
            float s = 0.f;
            for (int w = 0; w < PT_THREADS / 32; ++w) s += S.red[tid * 8 + w];
            part[L::P + tid] = s;
        }
        __threadfence();
        __syncthreads();
        const int c_lo = ts.cta_lo(m), c_hi = ts.cta_hi(m);
        if (tid == 0) S.last = (atomicAdd(A.counters + m, 1) == c_hi - c_lo);
        __syncthreads();
        if (S.last) {
            __threadfence();
            if (want_grad) {
                static_assert(L::P % 4 == 0 && PSTRIDE % 4 == 0, "float4 reduction needs P % 4 == 0");
                for (int p = 4 * tid; p < L::P; p += 4 * PT_THREADS) {
                    const float4 s = reduce_segments4(A.partial, ts, A.kmax, PSTRIDE, m, c_lo, c_hi, p);
                    *reinterpret_cast<float4*>(A.grad + (int64_t)m * L::P + p) = s;
                    if (A.out_params)      // meta_algos/base.py:209
                        *reinterpret_cast<float4*>(A.out_params + (int64_t)m * L::P + p) =
                            make_float4(S.P[p] - A.sgd_lr * s.x, S.P[p + 1] - A.sgd_lr * s.y, S.P[p + 2] - A.sgd_lr * s.z,
                                        S.P[p + 3] - A.sgd_lr * s.w);
                }
            }
            if (A.stats && tid < 3) {
                float s = 0.f;
                for (int c = c_lo; c <= c_hi; ++c)
                    s += __ldcg(A.partial + ((int64_t)c * A.kmax + (m - ts.first_task(c))) * PSTRIDE + L::P + tid);
                A.stats[(int64_t)m * 4 + tid] = s * invN;
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
        const int n0 = tile * TB, nb = max(0, min(TB, Nm - n0));
        const int64_t g0 = (int64_t)m * N + n0;
        __syncthreads();
        for (int i = tid; i < TB * DOP; i += PT_THREADS) {
            const int b = i / DOP, c = i % DOP;
            S.X[i] = (b < nb && c < DO) ? __ldg(A.obs + (g0 + b) * DO + c) : 0.f;
        }
        __syncthreads();
        // ---- layer 0: H1 = tanh(X W0 + b0)                      (policies/networks/mlp.py:96-117)
        {
            float acc[RM][4];
            const float4 bv = *reinterpret_cast<const float4*>(S.P + L::B0 + col0);
#pragma unroll
            for (int i = 0; i < RM; ++i) acc[i][0] = bv.x, acc[i][1] = bv.y, acc[i][2] = bv.z, acc[i][3] = bv.w;
            gemm_tile_smallk<DO, DOP, HID, RM>(S.X, S.P + L::W0, row0, col0, acc);
#pragma unroll
            for (int i = 0; i < RM; ++i)
                *reinterpret_cast<float4*>(S.H1 + (row0 + i) * LD + col0) =
                    make_float4(tanh_fast(acc[i][0]), tanh_fast(acc[i][1]), tanh_fast(acc[i][2]), tanh_fast(acc[i][3]));
        }
        __syncthreads();
        // ---- layer 1: H2 = tanh(H1 W1 + b1)
        {
            float acc[RM][4];
            const float4 bv = *reinterpret_cast<const float4*>(S.P + L::B1 + col0);
#pragma unroll
            for (int i = 0; i < RM; ++i) acc[i][0] = bv.x, acc[i][1] = bv.y, acc[i][2] = bv.z, acc[i][3] = bv.w;
            gemm_tile<HID, LD, HID, RM>(S.H1, S.P + L::W1, row0, col0, acc);
#pragma unroll
            for (int i = 0; i < RM; ++i)
                *reinterpret_cast<float4*>(S.H2 + (row0 + i) * LD + col0) =
                    make_float4(tanh_fast(acc[i][0]), tanh_fast(acc[i][1]), tanh_fast(acc[i][2]), tanh_fast(acc[i][3]));
        }
        __syncthreads();
        // ---- layer 2 + Gaussian head: 4 threads per sample row, quarter dot products + 2 shuffles
#ifndef PROMP_EXP_NO_HEAD
        {
            float mu[DA];
#pragma unroll
            for (int d = 0; d < DA; ++d) mu[d] = 0.f;
#pragma unroll
            for (int k4 = 0; k4 < QW / 4; ++k4) {
                const int j = rq * QW + 4 * k4;
                const float4 h = *reinterpret_cast<const float4*>(S.H2 + rb * LD + j);
#pragma unroll
                for (int d = 0; d < DA; ++d) {
                    mu[d] = fmaf(h.x, S.P[L::W2 + (j + 0) * DA + d], mu[d]);
                    mu[d] = fmaf(h.y, S.P[L::W2 + (j + 1) * DA + d], mu[d]);
                    mu[d] = fmaf(h.z, S.P[L::W2 + (j + 2) * DA + d], mu[d]);
                    mu[d] = fmaf(h.w, S.P[L::W2 + (j + 3) * DA + d], mu[d]);
                }
            }
#pragma unroll
            for (int d = 0; d < DA; ++d) {
                mu[d] += __shfl_xor_sync(0xffffffffu, mu[d], 1);
                mu[d] += __shfl_xor_sync(0xffffffffu, mu[d], 2);
                mu[d] += S.P[L::B2 + d];
            }
            if (rq == 0) {
                float dmu[DA], dls[DA];
                if (rb < nb) {
                    const int64_t n = g0 + rb;
                    float a[DA], mo[DA], lso[DA];
#pragma unroll
                    for (int d = 0; d < DA; ++d) {
                        a[d] = __ldg(A.act + n * DA + d);
                        mo[d] = __ldg(A.old_mean + n * DA + d);
                        lso[d] = A.ls_per_sample ? __ldg(A.old_ls + n * DA + d) : __ldg(A.old_ls + (int64_t)m * DA + d);
                    }
                    const float adv = __ldg(A.adv + n);
                    HeadOut<DA> o;
                    HeadOld<DA> ho;
                    head_old_from<DA>(lso, ho);
                    gaussian_head<DA>(hin, ho, mu, a, mo, adv, A.obj_kind, A.clip_eps, o);
                    const float wt = A.obj_scale * o.w * invN, kc = kl_eff * invN;
#pragma unroll
                    for (int d = 0; d < DA; ++d) {
                        dmu[d] = wt * o.zeta[d] * hin.inv_sig[d] + kc * o.dkl_dmu[d];
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
                for (int d = 0; d < DA; ++d) S.DMU[rb * DA + d] = dmu[d], S.DLS[rb * DA + d] = dls[d];
            }
        }
#endif
        if (!want_grad) continue;
        __syncthreads();
        // ---- output layer gradients (column role): gW2[cj][d] += sum_{b in slice} H2[b][cj] * DMU[b][d]
        {
            const int b0 = cp * BPP;
#pragma unroll 4
            for (int bb = 0; bb < BPP; ++bb) {
                const int b = b0 + bb;
                const float h = S.H2[b * LD + cj];
#pragma unroll
                for (int d = 0; d < DA; ++d) gW2p[d] = fmaf(h, S.DMU[b * DA + d], gW2p[d]);
            }
            if (tid < DA) {
                float s1 = 0.f, s2 = 0.f;
                for (int b = 0; b < nb; ++b) s1 += S.DMU[b * DA + tid], s2 += S.DLS[b * DA + tid];
                gB2 += s1;
                gLS += s2;
            }
        }
        __syncthreads();
        // ---- D2 = (DMU W2^T) * (1 - H2^2), in place over H2; bias gradient accumulates on the fly
#pragma unroll
        for (int i = 0; i < RM; ++i) {
            const int b = row0 + i;
            float4 h = *reinterpret_cast<float4*>(S.H2 + b * LD + col0);
            float hv[4] = {h.x, h.y, h.z, h.w}, o4[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float dh = 0.f;
#pragma unroll
                for (int d = 0; d < DA; ++d) dh = fmaf(S.DMU[b * DA + d], S.P[L::W2 + (col0 + c) * DA + d], dh);
                o4[c] = dh * (1.f - hv[c] * hv[c]);
                gB1p[c] += o4[c];
            }
            *reinterpret_cast<float4*>(S.H2 + b * LD + col0) = make_float4(o4[0], o4[1], o4[2], o4[3]);
        }
        __syncthreads();
        // ---- gW1 += H1^T D2 ; dH1 = D2 W1^T
        wgrad_tile<LD, RK>(S.H1, S.H2, ty * RK, col0, nb, gW1);
        float acc[RM][4];
#pragma unroll
        for (int i = 0; i < RM; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
        gemm_tile<HID, LD, HID, RM>(S.H2, S.W1T, row0, col0, acc);
        __syncthreads();   // every read of H1 (wgrad) is done before it is overwritten
#pragma unroll
        for (int i = 0; i < RM; ++i) {
            float4 h = *reinterpret_cast<float4*>(S.H1 + (row0 + i) * LD + col0);
            const float4 d1 = make_float4(acc[i][0] * (1.f - h.x * h.x), acc[i][1] * (1.f - h.y * h.y),
                                          acc[i][2] * (1.f - h.z * h.z), acc[i][3] * (1.f - h.w * h.w));
            gB0p[0] += d1.x; gB0p[1] += d1.y; gB0p[2] += d1.z; gB0p[3] += d1.w;
            *reinterpret_cast<float4*>(S.H1 + (row0 + i) * LD + col0) = d1;
        }
        __syncthreads();
        // ---- gW0 += X^T D1 (column role)
        {
            const int b0 = cp * BPP;
#pragma unroll 4
            for (int bb = 0; bb < BPP; ++bb) {
                const int b = b0 + bb;
                const float d1 = S.H1[b * LD + cj];
#pragma unroll
                for (int i = 0; i < DO; ++i) gW0p[i] = fmaf(S.X[b * DOP + i], d1, gW0p[i]);
            }
        }
    }
    if (cur_m >= 0) flush(cur_m);
}

// -------------------------------------------------------------------------------------------------
// Exact Hessian-vector product of the inner surrogate (R-operator: forward-mode tangent through the
// MLP + Gaussian log-likelihood, then the reverse sweep of the tangent), fused with the KL-penalty
// gradient:  out = vec - inner_lr * H vec + kl_coeff * grad KL.   Notation in the comments:
//   H1,H2 activations; R1,R2 their tangents; D* = backprop of the surrogate; C* = combined
//   (-inner_lr * tangent-of-backprop + kl-penalty backprop) signal that is accumulated into `out`.
template <int DO, int DA, int HID>
struct HvpSmem {
    using L = PLayout<DO, DA, HID>;
    using C = TileCfg<HID>;
    static constexpr int DOP = DOPad<DO>::V;
    static constexpr int PP = (L::P + 3) / 4 * 4;
    float P[PP];
    float V[PP];
    float W1T[HID * HID];
    float V1T[HID * HID];
    float X[TB * DOP];
    float H1[TB * C::LD];     // also: flush scratch
    float R1[TB * C::LD];
    float H2[TB * C::LD];
    float R2[TB * C::LD];
    float DMU[TB * DA];
    float CMU[TB * DA];
    float CLS[TB * DA];
    float red[3 * (PT_THREADS / 32)];
    int last;
};

template <int DO, int DA, int HID>
__global__ void __launch_bounds__(PT_THREADS) policy_hvp_kernel(PolicyArgs A) {
    using L = PLayout<DO, DA, HID>;
    using C = TileCfg<HID>;
    using R = RoleCfg<HID>;
    using SM = HvpSmem<DO, DA, HID>;
    constexpr int LD = C::LD, RM = C::RM, RK = C::RK, DOP = SM::DOP;
    constexpr int NPART = R::NPART, BPP = R::BPP, QW = R::QW;
    constexpr int PSTRIDE = L::P + PSTAT;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    SM& S = *reinterpret_cast<SM*>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = tid % C::TX, ty = tid / C::TX;
    const int row0 = ty * RM, col0 = tx * 4;
    const int rb = tid >> 2, rq = tid & 3;
    const int cj = tid % HID, cp = tid / HID;
    const TileSched ts(A.M, A.N, A.q);
    const int N = A.N;
    const float kl_eff = kl_coeff_eff(A);
    float invN = 1.0f / (float)N;       // both re-set per task when A.n_valid is given (variable-length paths)
    int Nm = N;
    const float ac = -A.inner_lr;                 // coefficient of H vec in `out`
    const float* th = nullptr;
    HeadIn<DA> hin;
    float rls[DA];     // tangent of the (clipped) log_std

    // accumulators: gW1c multiplies 1, gW1a multiplies `ac`
    float gW1c[RK][4], gW1a[RK][4], gB1p[4], gB0p[4], gW0p[DO], gW2p[DA];
    float gB2 = 0.f, gLS = 0.f;
    float s_obj, s_kl, s_ratio;
    auto zero_acc = [&]() {
#pragma unroll
        for (int r = 0; r < RK; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) gW1c[r][c] = gW1a[r][c] = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) gB1p[c] = gB0p[c] = 0.f;
#pragma unroll
        for (int i = 0; i < DO; ++i) gW0p[i] = 0.f;
#pragma unroll
        for (int d = 0; d < DA; ++d) gW2p[d] = 0.f;
        gB2 = gLS = 0.f;
        s_obj = s_kl = s_ratio = 0.f;
    };
    auto load_task = [&](int m, bool first) {
        if (A.n_valid) { Nm = __ldg(A.n_valid + m); invN = 1.0f / (float)max(Nm, 1); }
        th = A.params + (int64_t)m * A.param_stride;
        const float* vg = A.vec + (int64_t)m * L::P;
        const bool reload_p = first || A.param_stride != 0;      // shared theta stays resident across tasks
        __syncthreads();
        for (int i = tid; i < L::P; i += PT_THREADS) {
            if (reload_p) S.P[i] = __ldg(th + i);
            S.V[i] = __ldcg(vg + i);                             // written by the previous kernel on this stream
        }
        __syncthreads();
        for (int i = tid; i < HID * HID; i += PT_THREADS) {
            const int j = i / HID, k = i % HID;      // consecutive threads -> consecutive addresses of the transposes
            if (reload_p) S.W1T[j * HID + k] = S.P[L::W1 + k * HID + j];
            S.V1T[j * HID + k] = S.V[L::W1 + k * HID + j];
        }
        load_head_consts<DO, DA, HID>(S.P, A.clip_log_std, A.min_log_std, hin);
#pragma unroll
        for (int d = 0; d < DA; ++d) rls[d] = S.V[L::LS + d] * hin.ls_mask[d];
    };
    auto flush = [&](int m) {
        float* part = A.partial + ((int64_t)blockIdx.x * A.kmax + (m - ts.first_task(blockIdx.x))) * PSTRIDE;
        float* scr = S.H1;
        __syncthreads();
#pragma unroll
        for (int r = 0; r < RK; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) part[L::W1 + (ty * RK + r) * HID + col0 + c] = gW1c[r][c] + ac * gW1a[r][c];
#pragma unroll
        for (int c = 0; c < 4; ++c) scr[ty * HID + col0 + c] = gB1p[c];
        __syncthreads();
        if (tid < HID) {
            float s = 0.f;
            for (int y = 0; y < C::TY; ++y) s += scr[y * HID + tid];
            part[L::B1 + tid] = s;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 4; ++c) scr[ty * HID + col0 + c] = gB0p[c];
        __syncthreads();
        if (tid < HID) {
            float s = 0.f;
            for (int y = 0; y < C::TY; ++y) s += scr[y * HID + tid];
            part[L::B0 + tid] = s;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < DO; ++i) scr[(cp * DO + i) * HID + cj] = gW0p[i];
        __syncthreads();
        for (int idx = tid; idx < DO * HID; idx += PT_THREADS) {
            float s = 0.f;
            for (int p = 0; p < NPART; ++p) s += scr[p * DO * HID + idx];
            part[L::W0 + idx] = s;
        }
        __syncthreads();
#pragma unroll
        for (int d = 0; d < DA; ++d) scr[(cp * HID + cj) * DA + d] = gW2p[d];
        __syncthreads();
        for (int idx = tid; idx < HID * DA; idx += PT_THREADS) {
            float s = 0.f;
            for (int p = 0; p < NPART; ++p) s += scr[p * HID * DA + idx];
            part[L::W2 + idx] = s;
        }
        if (tid < DA) part[L::B2 + tid] = gB2, part[L::LS + tid] = gLS;
        const float v0 = warp_sum(rq == 0 ? s_obj : 0.f), v1 = warp_sum(rq == 0 ? s_kl : 0.f),
                    v2 = warp_sum(rq == 0 ? s_ratio : 0.f);
        __syncthreads();
        if (lane == 0) S.red[warp] = v0, S.red[8 + warp] = v1, S.red[16 + warp] = v2;
        __syncthreads();
        if (tid < 3) {
            float s = 0.f;
            for (int w = 0; w < PT_THREADS / 32; ++w) s += S.red[tid * 8 + w];
            part[L::P + tid] = s;
        }
        __threadfence();
        __syncthreads();
        const int c_lo = ts.cta_lo(m), c_hi = ts.cta_hi(m);
        if (tid == 0) S.last = (atomicAdd(A.counters + m, 1) == c_hi - c_lo);
        __syncthreads();
        if (S.last) {
            __threadfence();
            static_assert(L::P % 4 == 0 && PSTRIDE % 4 == 0, "float4 reduction needs P % 4 == 0");
            for (int p = 4 * tid; p < L::P; p += 4 * PT_THREADS) {
                const float4 s = reduce_segments4(A.partial, ts, A.kmax, PSTRIDE, m, c_lo, c_hi, p);
                *reinterpret_cast<float4*>(A.out + (int64_t)m * L::P + p) =
                    make_float4(S.V[p] + s.x, S.V[p + 1] + s.y, S.V[p + 2] + s.z, S.V[p + 3] + s.w);
            }
            if (A.stats && tid < 3) {
                float s = 0.f;
                for (int c = c_lo; c <= c_hi; ++c)
                    s += __ldcg(A.partial + ((int64_t)c * A.kmax + (m - ts.first_task(c))) * PSTRIDE + L::P + tid);
                A.stats[(int64_t)m * 4 + tid] = s * invN;
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
        const int n0 = tile * TB, nb = max(0, min(TB, Nm - n0));
        const int64_t g0 = (int64_t)m * N + n0;
        __syncthreads();
        for (int i = tid; i < TB * DOP; i += PT_THREADS) {
            const int b = i / DOP, c = i % DOP;
            S.X[i] = (b < nb && c < DO) ? __ldg(A.obs + (g0 + b) * DO + c) : 0.f;
        }
        __syncthreads();
        // ---- layer 0 and its tangent: H1 = tanh(X W0 + b0); R1 = (1-H1^2) * (X V0 + vb0)
        {
            float acc[RM][4], racc[RM][4];
            const float4 bv = *reinterpret_cast<const float4*>(S.P + L::B0 + col0);
            const float4 rv = *reinterpret_cast<const float4*>(S.V + L::B0 + col0);
#pragma unroll
            for (int i = 0; i < RM; ++i) {
                acc[i][0] = bv.x, acc[i][1] = bv.y, acc[i][2] = bv.z, acc[i][3] = bv.w;
                racc[i][0] = rv.x, racc[i][1] = rv.y, racc[i][2] = rv.z, racc[i][3] = rv.w;
            }
            gemm_tile_smallk<DO, DOP, HID, RM>(S.X, S.P + L::W0, row0, col0, acc);
            gemm_tile_smallk<DO, DOP, HID, RM>(S.X, S.V + L::W0, row0, col0, racc);
#pragma unroll
            for (int i = 0; i < RM; ++i) {
                float h[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) h[c] = tanh_fast(acc[i][c]);
                *reinterpret_cast<float4*>(S.H1 + (row0 + i) * LD + col0) = make_float4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<float4*>(S.R1 + (row0 + i) * LD + col0) =
                    make_float4((1.f - h[0] * h[0]) * racc[i][0], (1.f - h[1] * h[1]) * racc[i][1],
                                (1.f - h[2] * h[2]) * racc[i][2], (1.f - h[3] * h[3]) * racc[i][3]);
            }
        }
        __syncthreads();
        // ---- layer 1 and its tangent: R2 = (1-H2^2) * (R1 W1 + H1 V1 + vb1)
        {
            float acc[RM][4], racc[RM][4];
            const float4 bv = *reinterpret_cast<const float4*>(S.P + L::B1 + col0);
            const float4 rv = *reinterpret_cast<const float4*>(S.V + L::B1 + col0);
#pragma unroll
            for (int i = 0; i < RM; ++i) {
                acc[i][0] = bv.x, acc[i][1] = bv.y, acc[i][2] = bv.z, acc[i][3] = bv.w;
                racc[i][0] = rv.x, racc[i][1] = rv.y, racc[i][2] = rv.z, racc[i][3] = rv.w;
            }
            gemm_tile<HID, LD, HID, RM>(S.H1, S.P + L::W1, row0, col0, acc);
            gemm_tile<HID, LD, HID, RM>(S.R1, S.P + L::W1, row0, col0, racc);
            gemm_tile<HID, LD, HID, RM>(S.H1, S.V + L::W1, row0, col0, racc);
#pragma unroll
            for (int i = 0; i < RM; ++i) {
                float h[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) h[c] = tanh_fast(acc[i][c]);
                *reinterpret_cast<float4*>(S.H2 + (row0 + i) * LD + col0) = make_float4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<float4*>(S.R2 + (row0 + i) * LD + col0) =
                    make_float4((1.f - h[0] * h[0]) * racc[i][0], (1.f - h[1] * h[1]) * racc[i][1],
                                (1.f - h[2] * h[2]) * racc[i][2], (1.f - h[3] * h[3]) * racc[i][3]);
            }
        }
        __syncthreads();
        // ---- layer 2, its tangent, and the Gaussian head with its tangent (row role)
#ifndef PROMP_EXP_NO_HEAD
        {
            float mu[DA], rmu[DA];
#pragma unroll
            for (int d = 0; d < DA; ++d) mu[d] = rmu[d] = 0.f;
#pragma unroll
            for (int k4 = 0; k4 < QW / 4; ++k4) {
                const int j = rq * QW + 4 * k4;
                const float4 h4 = *reinterpret_cast<const float4*>(S.H2 + rb * LD + j);
                const float4 r4 = *reinterpret_cast<const float4*>(S.R2 + rb * LD + j);
                const float hv[4] = {h4.x, h4.y, h4.z, h4.w}, rv[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e)
#pragma unroll
                    for (int d = 0; d < DA; ++d) {
                        const float w2 = S.P[L::W2 + (j + e) * DA + d];
                        mu[d] = fmaf(hv[e], w2, mu[d]);
                        rmu[d] = fmaf(rv[e], w2, fmaf(hv[e], S.V[L::W2 + (j + e) * DA + d], rmu[d]));
                    }
            }
#pragma unroll
            for (int d = 0; d < DA; ++d) {
                mu[d] += __shfl_xor_sync(0xffffffffu, mu[d], 1);
                mu[d] += __shfl_xor_sync(0xffffffffu, mu[d], 2);
                rmu[d] += __shfl_xor_sync(0xffffffffu, rmu[d], 1);
                rmu[d] += __shfl_xor_sync(0xffffffffu, rmu[d], 2);
                mu[d] += S.P[L::B2 + d];
                rmu[d] += S.V[L::B2 + d];
            }
            if (rq == 0) {
                float dmu[DA], cmu[DA], cls[DA];
                if (rb < nb) {
                    const int64_t n = g0 + rb;
                    float a[DA], mo[DA], lso[DA];
#pragma unroll
                    for (int d = 0; d < DA; ++d) {
                        a[d] = __ldg(A.act + n * DA + d);
                        mo[d] = __ldg(A.old_mean + n * DA + d);
                        lso[d] = A.ls_per_sample ? __ldg(A.old_ls + n * DA + d) : __ldg(A.old_ls + (int64_t)m * DA + d);
                    }
                    const float adv = __ldg(A.adv + n);
                    HeadOut<DA> o;
                    HeadOld<DA> ho;
                    head_old_from<DA>(lso, ho);
                    gaussian_head<DA>(hin, ho, mu, a, mo, adv, A.obj_kind, A.clip_eps, o);
                    const float wt = o.w * invN, kc = kl_eff * invN;
                    // tangent of log p:  R l = sum_d (zeta/sig) R mu + (zeta^2 - 1) R ls
                    float rl = 0.f;
#pragma unroll
                    for (int d = 0; d < DA; ++d)
                        rl += (o.zeta[d] * hin.inv_sig[d]) * rmu[d] + (o.zeta[d] * o.zeta[d] - 1.f) * rls[d];
                    // d w / d logp: RATIO w = -A r -> R w = w R l ; LOGLIK w = -A -> 0
                    const float rwt = (A.obj_kind == PROMP_OBJ_RATIO) ? wt * rl : 0.f;
#pragma unroll
                    for (int d = 0; d < DA; ++d) {
                        const float is = hin.inv_sig[d], z = o.zeta[d];
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
                for (int d = 0; d < DA; ++d)
                    S.DMU[rb * DA + d] = dmu[d], S.CMU[rb * DA + d] = cmu[d], S.CLS[rb * DA + d] = cls[d];
            }
        }
#endif
        __syncthreads();
        // ---- output layer (column role): out_W2 += H2^T CMU + ac * R2^T DMU ; out_b2 += colsum CMU ; out_ls += colsum CLS
        {
            const int b0 = cp * BPP;
#pragma unroll 4
            for (int bb = 0; bb < BPP; ++bb) {
                const int b = b0 + bb;
                const float h = S.H2[b * LD + cj], r = ac * S.R2[b * LD + cj];
#pragma unroll
                for (int d = 0; d < DA; ++d) gW2p[d] = fmaf(h, S.CMU[b * DA + d], fmaf(r, S.DMU[b * DA + d], gW2p[d]));
            }
            if (tid < DA) {
                float s1 = 0.f, s2 = 0.f;
                for (int b = 0; b < nb; ++b) s1 += S.CMU[b * DA + tid], s2 += S.CLS[b * DA + tid];
                gB2 += s1;
                gLS += s2;
            }
        }
        __syncthreads();
        // ---- D2 = dH2 * g2 -> H2 ; C2 = CdH2 * g2 + ac * dH2 * (-2 H2 R2) -> R2 ; out_b1 accumulates C2 on the fly
#pragma unroll
        for (int i = 0; i < RM; ++i) {
            const int b = row0 + i;
            const float4 h4 = *reinterpret_cast<float4*>(S.H2 + b * LD + col0);
            const float4 r4 = *reinterpret_cast<float4*>(S.R2 + b * LD + col0);
            const float hv[4] = {h4.x, h4.y, h4.z, h4.w}, rv[4] = {r4.x, r4.y, r4.z, r4.w};
            float d2[4], c2[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float dh = 0.f, ch = 0.f;
#pragma unroll
                for (int d = 0; d < DA; ++d) {
                    const float w2 = S.P[L::W2 + (col0 + c) * DA + d], v2 = S.V[L::W2 + (col0 + c) * DA + d];
                    const float dm = S.DMU[b * DA + d];
                    dh = fmaf(dm, w2, dh);
                    ch = fmaf(S.CMU[b * DA + d], w2, fmaf(ac * dm, v2, ch));
                }
                const float g2 = 1.f - hv[c] * hv[c];
                d2[c] = dh * g2;
                c2[c] = ch * g2 + ac * dh * (-2.f * hv[c] * rv[c]);
                gB1p[c] += c2[c];
            }
            *reinterpret_cast<float4*>(S.H2 + b * LD + col0) = make_float4(d2[0], d2[1], d2[2], d2[3]);
            *reinterpret_cast<float4*>(S.R2 + b * LD + col0) = make_float4(c2[0], c2[1], c2[2], c2[3]);
        }
        __syncthreads();
        // ---- out_W1 += H1^T C2 + ac * R1^T D2
        wgrad_tile<LD, RK>(S.H1, S.R2, ty * RK, col0, nb, gW1c);
        wgrad_tile<LD, RK>(S.R1, S.H2, ty * RK, col0, nb, gW1a);
        // ---- dH1 = D2 W1^T ; CdH1 = C2 W1^T + ac * D2 V1^T
        float dh1[RM][4], ch1[RM][4];
#pragma unroll
        for (int i = 0; i < RM; ++i)
#pragma unroll
            for (int c = 0; c < 4; ++c) dh1[i][c] = ch1[i][c] = 0.f;
        gemm_tile<HID, LD, HID, RM>(S.H2, S.V1T, row0, col0, ch1);
#pragma unroll
        for (int i = 0; i < RM; ++i)
#pragma unroll
            for (int c = 0; c < 4; ++c) ch1[i][c] *= ac;
        gemm_tile<HID, LD, HID, RM>(S.R2, S.W1T, row0, col0, ch1);
        gemm_tile<HID, LD, HID, RM>(S.H2, S.W1T, row0, col0, dh1);
        __syncthreads();   // all reads of H1 / R1 by the weight-gradient loops are done
        // ---- C1 = CdH1 * g1 + ac * dH1 * (-2 H1 R1) -> H1 ; out_b0 accumulates C1 on the fly
#pragma unroll
        for (int i = 0; i < RM; ++i) {
            const float4 h4 = *reinterpret_cast<float4*>(S.H1 + (row0 + i) * LD + col0);
            const float4 r4 = *reinterpret_cast<float4*>(S.R1 + (row0 + i) * LD + col0);
            const float hv[4] = {h4.x, h4.y, h4.z, h4.w}, rv[4] = {r4.x, r4.y, r4.z, r4.w};
            float c1[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                c1[c] = ch1[i][c] * (1.f - hv[c] * hv[c]) + ac * dh1[i][c] * (-2.f * hv[c] * rv[c]);
                gB0p[c] += c1[c];
            }
            *reinterpret_cast<float4*>(S.H1 + (row0 + i) * LD + col0) = make_float4(c1[0], c1[1], c1[2], c1[3]);
        }
        __syncthreads();
        // ---- out_W0 += X^T C1 (column role)
        {
            const int b0 = cp * BPP;
#pragma unroll 4
            for (int bb = 0; bb < BPP; ++bb) {
                const int b = b0 + bb;
                const float c1 = S.H1[b * LD + cj];
#pragma unroll
                for (int i = 0; i < DO; ++i) gW0p[i] = fmaf(S.X[b * DOP + i], c1, gW0p[i]);
            }
        }
    }
    if (cur_m >= 0) flush(cur_m);
}

}  // namespace promp
#include "policy_tc.cuh"
namespace promp {

// -------------------------------------------------------------------------------------------------
// forward only: mean for arbitrary obs (distribution_info_sym / get_actions without sampling)
template <int DO, int DA, int HID>
__global__ void __launch_bounds__(128) policy_forward_kernel(int M, int N, const float* params, int64_t stride,
                                                              const float* obs, float* mean) {
    using L = PLayout<DO, DA, HID>;
    constexpr int NU = HID / 32;
    __shared__ float sP[L::P];
    __shared__ float sh[4][HID];
    const int m = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const float* th = params + (int64_t)m * stride;
    for (int i = threadIdx.x; i < L::P; i += blockDim.x) sP[i] = __ldg(th + i);
    __syncthreads();
    for (int n = blockIdx.x * 4 + w; n < N; n += gridDim.x * 4) {
        const float* o = obs + ((int64_t)m * N + n) * DO;
#pragma unroll
        for (int u = 0; u < NU; ++u) {
            const int j = lane + 32 * u;
            float z = sP[L::B0 + j];
            for (int i = 0; i < DO; ++i) z = fmaf(__ldg(o + i), sP[L::W0 + i * HID + j], z);
            sh[w][j] = tanh_fast(z);
        }
        __syncwarp();
        float mu[DA];
#pragma unroll
        for (int d = 0; d < DA; ++d) mu[d] = 0.f;
#pragma unroll
        for (int u = 0; u < NU; ++u) {
            const int j = lane + 32 * u;
            float z = sP[L::B1 + j];
            for (int k = 0; k < HID; ++k) z = fmaf(sh[w][k], sP[L::W1 + k * HID + j], z);
            const float h2 = tanh_fast(z);
#pragma unroll
            for (int d = 0; d < DA; ++d) mu[d] = fmaf(h2, sP[L::W2 + j * DA + d], mu[d]);
        }
#pragma unroll
        for (int d = 0; d < DA; ++d) {
            const float s = warp_sum(mu[d]) + sP[L::B2 + d];
            if (lane == d) mean[((int64_t)m * N + n) * DA + d] = s;
        }
        __syncwarp();
    }
}

__global__ void reduce_tasks_kernel(int M, int P, const float* in, float scale, float* out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    float s = 0.f;
    for (int m = 0; m < M; ++m) s += in[(int64_t)m * P + p];
    out[p] = s * scale;
}

__global__ void adam_tf1_kernel(int P, float* theta, const float* grad, float* mm, float* vv, int32_t* step, float lr,
                                float b1, float b2, float eps) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = *step + 1;
    // lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t)   (tf.train.AdamOptimizer)
    const float lr_t = lr * sqrtf(1.f - powf(b2, (float)t)) / (1.f - powf(b1, (float)t));
    if (p < P) {
        const float g = grad[p];
        const float mn = b1 * mm[p] + (1.f - b1) * g;
        const float vn = b2 * vv[p] + (1.f - b2) * g * g;
        mm[p] = mn;
        vv[p] = vn;
        theta[p] = theta[p] - lr_t * mn / (sqrtf(vn) + eps);
    }
}
__global__ void adam_step_inc_kernel(int32_t* step) { *step += 1; }

// [loss, inner_kl_0 .. inner_kl_{S-2}, outer_kl] of one meta-objective evaluation from the per-launch stats rows
// stats_all [S, M, 4] (row s < S-1: inner step s -> (surr, KL); row S-1: outer objective -> (surr, KL)):
//   loss = mean_m surr_{S-1,m} (+ mean_s coeff_s * inner_kl_s when coeff != NULL: pro_mp.py:151-155), means over M_global.
// One block, warp w reduces term w in a fixed order (deterministic).
__global__ void __launch_bounds__(256) meta_loss_terms_kernel(int S, int M, const float* __restrict__ stats_all, float inv_mg,
                                                               const float* __restrict__ coeff, int n_out, float* __restrict__ out) {
    __shared__ float terms[8];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (w < S + 1) {
        const int row = (w == 0 || w == S) ? S - 1 : w - 1;       // term 0: outer surr; 1..S-1: inner KLs; S: outer KL
        const int col = (w == 0) ? 0 : 1;
        float a = 0.f;
        for (int m = lane; m < M; m += 32) a += stats_all[((int64_t)row * M + m) * 4 + col];
        a = warp_sum(a) * inv_mg;
        if (lane == 0) terms[w] = a;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float loss = terms[0];
        if (coeff && S > 1) {
            float pen = 0.f;
            for (int s = 0; s < S - 1; ++s) pen += coeff[s] * terms[1 + s];
            loss += pen / (float)(S - 1);
        }
        out[0] = loss;
        for (int i = 1; i < n_out && i < S + 1; ++i) out[i] = terms[i];
    }
}

// Logged scalars of one sampling phase as float64, straight into the vector the Trainer reads back once per iteration:
//   out[0..5] = AverageDiscountedReturn, AverageReturn, NumTrajs, StdReturn, MaxReturn, MinReturn (samplers/base.py:135-149)
//               from the per-task sums promp_process_samples left in stats [M, 8],
//   out[6]    = AveragePolicyStd = mean exp(log_std) (policies/gaussian_mlp_policy.py:118-123).
// One launch instead of ~15 reduce / elementwise / cat kernels per phase.
__global__ void __launch_bounds__(256) phase_log_terms_kernel(int M, int Da, double n_paths, const double* __restrict__ stats,
                                                               const float* __restrict__ log_std, double* __restrict__ out) {
    __shared__ double red[8][6];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    double v[6] = {0.0, 0.0, 0.0, -1e300, 1e300, 0.0};      // sum R0, sum G, sum G^2, max G, min G, sum exp(log_std)
    for (int m = tid; m < M; m += 256) {
        const double* st = stats + (int64_t)m * 8;
        v[0] += st[0]; v[1] += st[1]; v[2] += st[2];
        v[3] = fmax(v[3], st[3]);
        v[4] = fmin(v[4], st[4]);
    }
    for (int i = tid; i < M * Da; i += 256) v[5] += (double)expf(log_std[i]);
#pragma unroll
    for (int k = 0; k < 6; ++k) v[k] = k == 3 ? warp_max(v[k]) : k == 4 ? warp_min(v[k]) : warp_sum(v[k]);
    if (lane == 0)
        for (int k = 0; k < 6; ++k) red[w][k] = v[k];
    __syncthreads();
    if (tid == 0) {
        for (int i = 1; i < 8; ++i) {
            v[0] += red[i][0]; v[1] += red[i][1]; v[2] += red[i][2]; v[5] += red[i][5];
            v[3] = fmax(v[3], red[i][3]);
            v[4] = fmin(v[4], red[i][4]);
        }
        const double mean_g = v[1] / n_paths;
        out[0] = v[0] / n_paths;
        out[1] = mean_g;
        out[2] = n_paths;
        out[3] = sqrt(fmax(v[2] / n_paths - mean_g * mean_g, 0.0));
        out[4] = v[3];
        out[5] = v[4];
        out[6] = v[5] / (double)(M * Da);
    }
}

// ProMP's logged scalars (pro_mp.py:193-198) from the optimizer's device vector [loss_before, loss_after, inner KLs.., outer KL]
__global__ void promp_log_terms_kernel(int S1, const float* __restrict__ final_terms, double* __restrict__ out) {
    if (threadIdx.x == 0) {
        out[0] = (double)final_terms[0];
        out[1] = (double)final_terms[1];
        float s = 0.f;
        for (int i = 0; i < S1; ++i) s += final_terms[2 + i];
        out[2] = S1 > 0 ? (double)(s / (float)S1) : 0.0;
    }
}

// ProMP._adapt_kl_coeff (pro_mp.py:201-214) on the device: coeff_s /= 2 if KL_s < target / 1.5, *= 2 if KL_s > target * 1.5
// (comparisons in double like the reference's Python floats; halving / doubling is exact in float32).  out4 (optional):
// ProMP's four logged scalars [LossBefore, LossAfter, KLInner, KLCoeffInner (after the update)] (pro_mp.py:193-198).
__global__ void adapt_kl_coeff_kernel(int S1, const float* __restrict__ final_terms, double target, int adapt, float* __restrict__ coeff,
                                      double* __restrict__ out4) {
    if (threadIdx.x == 0) {
        if (out4) {
            out4[0] = (double)final_terms[0];
            out4[1] = (double)final_terms[1];
            float s = 0.f;
            for (int i = 0; i < S1; ++i) s += final_terms[2 + i];
            out4[2] = S1 > 0 ? (double)(s / (float)S1) : 0.0;
        }
        double* out_mean = out4 ? out4 + 3 : nullptr;
        double sum = 0.0;
        for (int i = 0; i < S1; ++i) {
            float c = coeff[i];
            if (adapt) {
                const double kl = (double)final_terms[2 + i];
                if (kl < target / 1.5) c *= 0.5f;
                else if (kl > target * 1.5) c *= 2.0f;
                coeff[i] = c;
            }
            sum += (double)c;
        }
        if (out_mean) *out_mean = S1 > 0 ? sum / (double)S1 : 0.0;
    }
}

// -------------------------------------------------------------------------------------------------
static int g_use_tc = 1;     // promp_set_option("tensor_cores", 0|1): HID = 64 policy kernels on tcgen05 (default) or CUDA cores

static int g_tc_threads = 0;  // promp_set_option("tc_threads", 0|256|512): 0 = per-shape default
// column groups of the TC kernels' thread mapping: 2 -> 256 threads (32 hidden units per thread), 4 -> 512 threads (16)
static int tc_column_groups(int obs_dim) {
    if (g_tc_threads == 256) return 2;
    if (g_tc_threads == 512) return 4;
    return obs_dim <= 4 ? 4 : 2;
}

struct TilePlan {
    int grid, q, kmax;
    int64_t partial_floats;
};
// One-wave persistent plan: `slots` resident CTAs share the T tiles as evenly as possible.
static TilePlan plan_tiles(int M, int N, int slots, int P, int tb = TB) {
    const int ntiles = (N + tb - 1) / tb;
    const int64_t T = (int64_t)M * ntiles;
    TilePlan p;
    int g = (int)(T < slots ? T : slots);
    if (g < 1) g = 1;
    p.q = (int)((T + g - 1) / g);
    p.grid = (int)((T + p.q - 1) / p.q);
    p.kmax = (p.q + ntiles - 1) / ntiles + 1;
    p.partial_floats = (int64_t)p.grid * p.kmax * (P + PSTAT);
    return p;
}
static int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}
static int64_t counters_bytes(int M) { return (((int64_t)M * sizeof(int) + 15) / 16) * 16; }

template <typename Kernel>
static int launch_policy(Kernel kernel, int smem, int& occ_cache, PolicyArgs& A, int P, void* ws, int64_t ws_bytes,
                         cudaStream_t st, const char* name, int tb = TB, int threads = PT_THREADS) {
    if (occ_cache == 0) {
        PROMP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int occ = 0;
        PROMP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem));
        occ_cache = occ < 1 ? 1 : occ;
    }
    const TilePlan p = plan_tiles(A.M, A.N, sm_count() * occ_cache, P, tb);
    const int64_t need = counters_bytes(A.M) + p.partial_floats * (int64_t)sizeof(float);
    if (ws_bytes < need) {
        set_error("policy workspace too small (%lld < %lld bytes)", (long long)ws_bytes, (long long)need);
        return PROMP_ERR_WORKSPACE;
    }
    A.counters = (int*)ws;
    A.partial = (float*)((char*)ws + counters_bytes(A.M));
    A.q = p.q;
    A.kmax = p.kmax;
    kernel<<<p.grid, threads, smem, st>>>(A);
    PROMP_LAUNCH_CHECK(name);
    return PROMP_OK;
}

template <int DO, int DA, int HID>
static int launch_grad(PolicyArgs& A, void* ws, int64_t ws_bytes, cudaStream_t st) {
    static int occ = 0;
    return launch_policy(policy_grad_kernel<DO, DA, HID>, (int)sizeof(GradSmem<DO, DA, HID>), occ, A,
                         PLayout<DO, DA, HID>::P, ws, ws_bytes, st, "policy_grad_kernel");
}

template <int DO, int DA, int HID>
static int launch_grad_any(PolicyArgs& A, void* ws, int64_t ws_bytes, cudaStream_t st) {
    if constexpr (HID == TC_HID) {
        if (g_use_tc) {
            static int occ2 = 0, occ4 = 0;
            if (tc_column_groups(DO) == 4)
                return launch_policy(policy_grad_tc_kernel<DO, DA, 4>, (int)sizeof(GradTcSmem<DO, DA, 4>), occ4, A,
                                     PLayout<DO, DA, HID>::P, ws, ws_bytes, st, "policy_grad_tc_kernel", TBT, 512);
            return launch_policy(policy_grad_tc_kernel<DO, DA, 2>, (int)sizeof(GradTcSmem<DO, DA, 2>), occ2, A,
                                 PLayout<DO, DA, HID>::P, ws, ws_bytes, st, "policy_grad_tc_kernel", TBT, 256);
        }
    }
    return launch_grad<DO, DA, HID>(A, ws, ws_bytes, st);
}

template <int DO, int DA, int HID>
static int launch_hvp(PolicyArgs& A, void* ws, int64_t ws_bytes, cudaStream_t st) {
    if constexpr (HID == TC_HID && sizeof(HvpTcSmem<DO, DA, 2>) <= 227 * 1024) {     // fits the 227 KB of one SM
        if (g_use_tc) {
            static int occ2 = 0, occ4 = 0;
            if (tc_column_groups(DO) == 4)
                return launch_policy(policy_hvp_tc_kernel<DO, DA, 4>, (int)sizeof(HvpTcSmem<DO, DA, 4>), occ4, A,
                                     PLayout<DO, DA, HID>::P, ws, ws_bytes, st, "policy_hvp_tc_kernel", TBT, 512);
            return launch_policy(policy_hvp_tc_kernel<DO, DA, 2>, (int)sizeof(HvpTcSmem<DO, DA, 2>), occ2, A,
                                 PLayout<DO, DA, HID>::P, ws, ws_bytes, st, "policy_hvp_tc_kernel", TBT, 256);
        }
    }
    static int occ = 0;
    return launch_policy(policy_hvp_kernel<DO, DA, HID>, (int)sizeof(HvpSmem<DO, DA, HID>), occ, A,
                         PLayout<DO, DA, HID>::P, ws, ws_bytes, st, "policy_hvp_kernel");
}

// ---- dataflow chain (policy_chain_tc_kernel): work-item plan + launch ---------------------------------------------
static int g_chain = -1;         // promp_set_option("chain", -1|0|1): dataflow kernel always (1), never (0: one launch per stage), or
                                 // where it wins (-1, default; see chain_uses_dataflow)
static int g_chain_q = 0;        // promp_set_option("chain_q", q): tiles per work item (0 = automatic)
static int g_chain_taper = 1;    // promp_set_option("chain_taper", 0|1): last stage's items shrink to one tile towards the end

struct ChainPlan {
    ChainStageInfo info[CHAIN_MAX_STAGES];
    int n_items;
    int64_t ctrl_bytes, bytes;     // control words (zero on entry, left zero); whole workspace
};
static int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
// Items are (stage, task, q consecutive tiles).  q = the largest of 4, 2, 1 that still gives every stage at least one item
// per SM; the last stage (nothing left to fill its tail with) uses q, q/2 and 1 on the first half, third quarter and last
// quarter of its tasks, so the final imbalance over the SMs is one tile.
static ChainPlan plan_chain(int n_stages, const int* kinds, const int* Ns, int M, int P) {
    ChainPlan pl;
    memset(&pl, 0, sizeof(pl));
    const int sms = sm_count();
    int base = 0;
    for (int s = 0; s < n_stages; ++s) {
        ChainStageInfo& I = pl.info[s];
        I.kind = kinds[s];
        I.ntiles = (Ns[s] + TBT - 1) / TBT;
        int q = 4;
        if (g_chain_q > 0) q = g_chain_q;
        else while (q > 1 && (int64_t)M * ((I.ntiles + q - 1) / q) < sms) q >>= 1;
        if (q > I.ntiles) q = I.ntiles;
        I.item_base = base;
        const bool taper = g_chain_taper && s == n_stages - 1 && q > 1 && M >= 4;
        if (!taper) {
            I.n_regions = 1;
            I.reg_m0[0] = 0, I.reg_m0[1] = M;
            I.reg_q[0] = q;
            I.reg_item0[0] = 0;
            I.n_items = M * ((I.ntiles + q - 1) / q);
        } else {
            const int q2 = q > 2 ? q / 2 : q;              // q = 2: three quarters at 2, the last quarter at 1
            const int mB = M - M / 4, mA = q2 < q ? M / 2 : mB;
            I.n_regions = 0;
            int items = 0;
            const int m0[4] = {0, mA, mB, M}, qs[3] = {q, q2, 1};
            for (int r = 0; r < 3; ++r) {
                if (m0[r + 1] <= m0[r]) continue;
                I.reg_m0[I.n_regions] = m0[r];
                I.reg_q[I.n_regions] = qs[r];
                I.reg_item0[I.n_regions] = items;
                items += (m0[r + 1] - m0[r]) * ((I.ntiles + qs[r] - 1) / qs[r]);
                ++I.n_regions;
            }
            I.reg_m0[I.n_regions] = M;
            I.n_items = items;
        }
        base += I.n_items;
    }
    pl.n_items = base;
    // fixed layout whatever n_stages is: chains of different length share one workspace, and the words must stay zero between them
    pl.ctrl_bytes = align_up(16 + (int64_t)2 * CHAIN_MAX_STAGES * M * sizeof(int), 128);
    pl.bytes = pl.ctrl_bytes + (int64_t)pl.n_items * (P + PSTAT) * sizeof(float);
    return pl;
}

// the automatic choice: dataflow kernel for chains whose stages have one to three tiles per SM (where launch tails and the
// tile quantisation of every single launch dominate: measured wins of 2-10 %, profiles/r02_chain_time.txt), one launch per stage
// below (a stage is not even one wave: nothing to balance) and above (per-item flushes outweigh the gain)
static bool chain_uses_dataflow(const ChainPlan& pl, int n_stages, int M) {
    bool fits = true;
    for (int s = 0; s < n_stages; ++s) {
        const int64_t T = (int64_t)M * pl.info[s].ntiles;
        fits = fits && T >= sm_count() && T <= 3 * sm_count();
    }
    return g_use_tc && (g_chain == 1 || (g_chain < 0 && fits));
}

template <int DO, int DA, int HID>
static constexpr bool chain_tc_ok() {
    if constexpr (HID == TC_HID) return ChainSmem<DO, DA, 2>::SIZE <= 227 * 1024 && ChainSmem<DO, DA, 4>::SIZE <= 227 * 1024;
    return false;
}

template <int DO, int DA, int NQ, bool HAS_HVP>
static int launch_chain_nq(ChainArgs& C, cudaStream_t st) {
    static int configured = 0;
    constexpr int smem = ChainSmem<DO, DA, NQ>::SIZE;
    auto kernel = policy_chain_tc_kernel<DO, DA, NQ, HAS_HVP>;
    if (!configured) {
        PROMP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = 1;
    }
    int grid = sm_count();
    if (grid > C.n_items) grid = C.n_items;
    kernel<<<grid, 128 * NQ, smem, st>>>(C);
    PROMP_LAUNCH_CHECK("policy_chain_tc_kernel");
    return PROMP_OK;
}

// kinds / Ns / A: the stages in order.  Falls back to one launch per stage (same results up to summation order) for the
// shapes the tcgen05 kernels do not cover or when the "chain" option is off.
template <int DO, int DA, int HID>
static int launch_chain(int n_stages, const int* kinds, PolicyArgs* A, const int* skip_flag, const float* skip_theta, void* ws,
                        int64_t ws_bytes, cudaStream_t st) {
    constexpr int P = PLayout<DO, DA, HID>::P;
    int Ns[CHAIN_MAX_STAGES];
    for (int s = 0; s < n_stages; ++s) Ns[s] = A[s].N;
    const int M = A[0].M;
    const ChainPlan pl = plan_chain(n_stages, kinds, Ns, M, P);
    if constexpr (chain_tc_ok<DO, DA, HID>()) {
        if (chain_uses_dataflow(pl, n_stages, M)) {
            if (ws_bytes < pl.bytes) {
                set_error("policy chain workspace too small (%lld < %lld bytes)", (long long)ws_bytes, (long long)pl.bytes);
                return PROMP_ERR_WORKSPACE;
            }
            ChainArgs C;
            memset(&C, 0, sizeof(C));
            C.n_stages = n_stages, C.n_items = pl.n_items, C.M = M;
            C.ctrl = (int*)ws;
            C.ready = (int*)ws + 4;
            C.skip_flag = skip_flag, C.skip_theta = skip_theta;
            float* partial = (float*)((char*)ws + pl.ctrl_bytes);
            for (int s = 0; s < n_stages; ++s) {
                C.info[s] = pl.info[s];
                C.st[s] = A[s];
                C.st[s].counters = (int*)ws + 4 + (CHAIN_MAX_STAGES + s) * M;
                C.st[s].partial = partial;            // slots are numbered by global item id
            }
            bool has_hvp = false;
            for (int s = 0; s < n_stages; ++s) has_hvp = has_hvp || kinds[s] == 1;
            if (tc_column_groups(DO) == 4)
                return has_hvp ? launch_chain_nq<DO, DA, 4, true>(C, st) : launch_chain_nq<DO, DA, 4, false>(C, st);
            return has_hvp ? launch_chain_nq<DO, DA, 2, true>(C, st) : launch_chain_nq<DO, DA, 2, false>(C, st);
        }
    }
    // one launch per stage; their (counters + partial) workspace starts after the chain's control words
    void* ws1 = (char*)ws + pl.ctrl_bytes;
    const int64_t ws1_bytes = ws_bytes - pl.ctrl_bytes;
    for (int s = 0; s < n_stages; ++s) {
        PolicyArgs a = A[s];
        if (s == 0) a.skip_flag = skip_flag, a.skip_theta = skip_theta;
        const int rc = kinds[s] == 0 ? launch_grad_any<DO, DA, HID>(a, ws1, ws1_bytes, st) : launch_hvp<DO, DA, HID>(a, ws1, ws1_bytes, st);
        if (rc != PROMP_OK) return rc;
    }
    return PROMP_OK;
}

template <int DO, int DA, int HID>
static int chain_num_launches(int n_stages, const int* kinds, const int* Ns, int M) {
    if constexpr (chain_tc_ok<DO, DA, HID>()) {
        const ChainPlan pl = plan_chain(n_stages, kinds, Ns, M, PLayout<DO, DA, HID>::P);
        if (chain_uses_dataflow(pl, n_stages, M)) return 1;
    }
    return n_stages;
}

template <int DO, int DA, int HID>
static int64_t chain_ws_bytes(int n_stages, const int* kinds, const int* Ns, int M) {
    const ChainPlan pl = plan_chain(n_stages, kinds, Ns, M, PLayout<DO, DA, HID>::P);
    int nmax = 1;
    for (int s = 0; s < n_stages; ++s) nmax = Ns[s] > nmax ? Ns[s] : nmax;
    const int64_t single = pl.ctrl_bytes + promp_policy_workspace_bytes(M, nmax, DO, DA, HID);
    return pl.bytes > single ? pl.bytes : single;
}

template <int DO, int DA, int HID>
static int launch_forward(int M, int N, const float* params, int64_t stride, const float* obs, float* mean,
                          cudaStream_t st) {
    int gx = (N + 3) / 4;
    const int cap = (4 * 148 + M - 1) / M;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    policy_forward_kernel<DO, DA, HID><<<dim3(gx, M), 128, 0, st>>>(M, N, params, stride, obs, mean);
    PROMP_LAUNCH_CHECK("policy_forward_kernel");
    return PROMP_OK;
}

// supported (obs_dim, act_dim, hidden) instantiations
#define PROMP_DISPATCH_DIMS(FN, ...)                                                           \
    if (obs_dim == 2 && act_dim == 2 && hidden == 64) return FN<2, 2, 64>(__VA_ARGS__);        \
    if (obs_dim == 2 && act_dim == 2 && hidden == 32) return FN<2, 2, 32>(__VA_ARGS__);        \
    if (obs_dim == 17 && act_dim == 6 && hidden == 64) return FN<17, 6, 64>(__VA_ARGS__);      \
    if (obs_dim == 17 && act_dim == 6 && hidden == 32) return FN<17, 6, 32>(__VA_ARGS__);      \
    if (obs_dim == 4 && act_dim == 2 && hidden == 64) return FN<4, 2, 64>(__VA_ARGS__);        \
    if (obs_dim == 4 && act_dim == 2 && hidden == 32) return FN<4, 2, 32>(__VA_ARGS__);        \
    set_error("unsupported (obs_dim, act_dim, hidden) = (%d, %d, %d); built: (2,2,{32,64}), (4,2,{32,64}), (17,6,{32,64})", \
              obs_dim, act_dim, hidden);                                                       \
    return PROMP_ERR_INVALID_ARG;

}  // namespace promp

using namespace promp;

extern "C" int64_t promp_policy_workspace_bytes(int M, int N, int obs_dim, int act_dim, int hidden) {
    // upper bound over the occupancies the kernels can have (1..4 CTAs per SM on 148..160 SMs)
    const int P = promp::num_params(obs_dim, act_dim, hidden);
    int64_t worst = 0;
    for (int occ = 1; occ <= 4; ++occ)
        for (int tb : {TB, TBT}) {            // CUDA-core kernels tile by 64 samples, the tensor-core kernels by 128
            const TilePlan p = plan_tiles(M, N, 160 * occ, P, tb);
            if (p.partial_floats > worst) worst = p.partial_floats;
            const TilePlan p2 = plan_tiles(M, N, 148 * occ, P, tb);
            if (p2.partial_floats > worst) worst = p2.partial_floats;
        }
    return counters_bytes(M) + worst * (int64_t)sizeof(float) + 16;
}

static int check_policy_args(const char* who, int M, int N, const void* params, const void* obs, const void* act,
                             const void* adv, const void* old_mean, const void* old_ls, const void* ws) {
    PROMP_REQUIRE(M > 0 && N > 0, "%s: M and N must be positive (got %d, %d)", who, M, N);
        PROMP_REQUIRE(params && obs && act && adv && old_mean && old_ls && ws, "%s: null pointer argument", who);
    return PROMP_OK;
}

extern "C" int promp_policy_grad_ex(int obs_dim, int act_dim, int hidden, int M, int N, const int32_t* n_valid,
                                    const float* params, int64_t param_stride, const float* obs, const float* act,
                                    const float* adv, const float* old_mean, const float* old_log_std, int ls_per_sample,
                                    int obj_kind, float obj_scale, float clip_eps, float kl_coeff, int clip_log_std,
                                    float min_log_std, float* grad, float* out_params, float sgd_lr, float* stats,
                                    const int32_t* skip_flag, const float* skip_theta, int32_t* unclipped_out, float* theta_copy_out,
                                    void* workspace, int64_t workspace_bytes, void* stream) {
    int st = check_policy_args(n_valid ? "promp_policy_grad_ragged" : "promp_policy_grad", M, N, params, obs, act, adv, old_mean, old_log_std, workspace);
    if (st != PROMP_OK) return st;
    PROMP_REQUIRE(obj_kind >= 0 && obj_kind <= 3, "promp_policy_grad: bad obj_kind %d", obj_kind);
    PROMP_REQUIRE(!(out_params && !grad), "promp_policy_grad: out_params needs grad");
    PROMP_REQUIRE((skip_flag == nullptr) == (skip_theta == nullptr) && (unclipped_out == nullptr) == (theta_copy_out == nullptr),
                  "promp_policy_grad_ex: skip_flag / skip_theta and unclipped_out / theta_copy_out come in pairs");
    PROMP_REQUIRE(!(skip_flag || unclipped_out) || param_stride == 0,
                  "promp_policy_grad_ex: launch re-use is defined for the shared pre-update parameters (param_stride 0)");
    PolicyArgs A{};
    A.M = M; A.N = N; A.params = params; A.param_stride = param_stride;
    A.obs = obs; A.act = act; A.adv = adv; A.old_mean = old_mean; A.old_ls = old_log_std;
    A.ls_per_sample = ls_per_sample; A.obj_kind = obj_kind; A.obj_scale = obj_scale; A.clip_eps = clip_eps;
    A.kl_coeff = kl_coeff; A.clip_log_std = clip_log_std; A.min_log_std = min_log_std;
    A.grad = grad; A.out_params = out_params; A.sgd_lr = sgd_lr; A.stats = stats; A.n_valid = n_valid;
    A.skip_flag = skip_flag; A.skip_theta = skip_theta; A.unclipped_out = unclipped_out; A.theta_copy_out = theta_copy_out;
    cudaStream_t s = (cudaStream_t)stream;
    PROMP_DISPATCH_DIMS(launch_grad_any, A, workspace, workspace_bytes, s)
}

extern "C" int promp_policy_grad_ragged(int obs_dim, int act_dim, int hidden, int M, int N, const int32_t* n_valid,
                                        const float* params, int64_t param_stride, const float* obs, const float* act,
                                        const float* adv, const float* old_mean, const float* old_log_std, int ls_per_sample,
                                        int obj_kind, float obj_scale, float clip_eps, float kl_coeff, int clip_log_std,
                                        float min_log_std, float* grad, float* out_params, float sgd_lr, float* stats,
                                        void* workspace, int64_t workspace_bytes, void* stream) {
    return promp_policy_grad_ex(obs_dim, act_dim, hidden, M, N, n_valid, params, param_stride, obs, act, adv, old_mean, old_log_std,
                                ls_per_sample, obj_kind, obj_scale, clip_eps, kl_coeff, clip_log_std, min_log_std, grad, out_params,
                                sgd_lr, stats, nullptr, nullptr, nullptr, nullptr, workspace, workspace_bytes, stream);
}

extern "C" int promp_policy_grad(int obs_dim, int act_dim, int hidden, int M, int N, const float* params,
                                 int64_t param_stride, const float* obs, const float* act, const float* adv,
                                 const float* old_mean, const float* old_log_std, int ls_per_sample, int obj_kind,
                                 float obj_scale, float clip_eps, float kl_coeff, int clip_log_std, float min_log_std,
                                 float* grad, float* out_params, float sgd_lr, float* stats, void* workspace,
                                 int64_t workspace_bytes, void* stream) {
    return promp_policy_grad_ragged(obs_dim, act_dim, hidden, M, N, nullptr, params, param_stride, obs, act, adv, old_mean,
                                    old_log_std, ls_per_sample, obj_kind, obj_scale, clip_eps, kl_coeff, clip_log_std, min_log_std,
                                    grad, out_params, sgd_lr, stats, workspace, workspace_bytes, stream);
}

extern "C" int promp_policy_hvp_ragged(int obs_dim, int act_dim, int hidden, int M, int N, const int32_t* n_valid,
                                       const float* params, int64_t param_stride, const float* obs, const float* act,
                                       const float* adv, const float* old_mean, const float* old_log_std, int ls_per_sample,
                                       int obj_kind, float inner_lr, float kl_coeff, int clip_log_std, float min_log_std,
                                       const float* vec, float* out, float* stats, void* workspace, int64_t workspace_bytes,
                                       void* stream) {
    int st = check_policy_args(n_valid ? "promp_policy_hvp_ragged" : "promp_policy_hvp", M, N, params, obs, act, adv, old_mean, old_log_std, workspace);
    if (st != PROMP_OK) return st;
    PROMP_REQUIRE(obj_kind == PROMP_OBJ_RATIO || obj_kind == PROMP_OBJ_LOGLIK,
                  "promp_policy_hvp: inner objective must be RATIO or LOGLIK (got %d)", obj_kind);
    PROMP_REQUIRE(vec && out, "promp_policy_hvp: vec/out must not be null");
    PolicyArgs A{};
    A.M = M; A.N = N; A.params = params; A.param_stride = param_stride;
    A.obs = obs; A.act = act; A.adv = adv; A.old_mean = old_mean; A.old_ls = old_log_std;
    A.ls_per_sample = ls_per_sample; A.obj_kind = obj_kind; A.obj_scale = 1.f; A.clip_eps = 0.f;
    A.kl_coeff = kl_coeff; A.clip_log_std = clip_log_std; A.min_log_std = min_log_std;
    A.vec = vec; A.out = out; A.inner_lr = inner_lr; A.stats = stats; A.n_valid = n_valid;
    cudaStream_t s = (cudaStream_t)stream;
    PROMP_DISPATCH_DIMS(launch_hvp, A, workspace, workspace_bytes, s)
}

extern "C" int promp_policy_hvp(int obs_dim, int act_dim, int hidden, int M, int N, const float* params,
                                int64_t param_stride, const float* obs, const float* act, const float* adv,
                                const float* old_mean, const float* old_log_std, int ls_per_sample, int obj_kind,
                                float inner_lr, float kl_coeff, int clip_log_std, float min_log_std, const float* vec,
                                float* out, float* stats, void* workspace, int64_t workspace_bytes, void* stream) {
    return promp_policy_hvp_ragged(obs_dim, act_dim, hidden, M, N, nullptr, params, param_stride, obs, act, adv, old_mean,
                                   old_log_std, ls_per_sample, obj_kind, inner_lr, kl_coeff, clip_log_std, min_log_std, vec, out,
                                   stats, workspace, workspace_bytes, stream);
}

static int chain_stage_args(const promp_policy_stage* stages, int n_stages, int M, float min_log_std, PolicyArgs* A, int* kinds, int* Ns) {
    PROMP_REQUIRE(stages != nullptr && n_stages >= 1 && n_stages <= CHAIN_MAX_STAGES,
                  "promp_policy_chain: 1..%d stages (got %d)", CHAIN_MAX_STAGES, n_stages);
    for (int s = 0; s < n_stages; ++s) {
        const promp_policy_stage& g = stages[s];
        PROMP_REQUIRE(g.kind == 0 || g.kind == 1, "promp_policy_chain: stage %d has kind %d (0 = gradient, 1 = HVP)", s, g.kind);
        PROMP_REQUIRE(g.N > 0 && g.params && g.obs && g.act && g.adv && g.old_mean && g.old_log_std,
                      "promp_policy_chain: stage %d: null pointer / non-positive N", s);
        PolicyArgs& a = A[s];
        memset(&a, 0, sizeof(a));
        a.M = M; a.N = g.N; a.params = g.params; a.param_stride = g.param_stride;
        a.obs = g.obs; a.act = g.act; a.adv = g.adv; a.old_mean = g.old_mean; a.old_ls = g.old_log_std;
        a.ls_per_sample = g.ls_per_sample; a.obj_kind = g.obj_kind; a.kl_coeff = g.kl_coeff;
        a.clip_log_std = g.clip_log_std; a.min_log_std = min_log_std; a.stats = g.stats; a.n_valid = g.n_valid;
        a.kl_coeff_ptr = g.kl_coeff_dev;
        if (g.kind == 0) {
            PROMP_REQUIRE(g.obj_kind >= 0 && g.obj_kind <= 3, "promp_policy_chain: stage %d: bad obj_kind %d", s, g.obj_kind);
            PROMP_REQUIRE(!(g.out_params && !g.grad), "promp_policy_chain: stage %d: out_params needs grad", s);
            a.obj_scale = g.obj_scale; a.clip_eps = g.clip_eps; a.grad = g.grad; a.out_params = g.out_params; a.sgd_lr = g.sgd_lr;
        } else {
            PROMP_REQUIRE(g.obj_kind == PROMP_OBJ_RATIO || g.obj_kind == PROMP_OBJ_LOGLIK,
                          "promp_policy_chain: stage %d: HVP inner objective must be RATIO or LOGLIK (got %d)", s, g.obj_kind);
            PROMP_REQUIRE(g.vec && g.out, "promp_policy_chain: stage %d: vec / out must not be null", s);
            a.obj_scale = 1.f; a.vec = g.vec; a.out = g.out; a.inner_lr = g.inner_lr;
        }
        kinds[s] = g.kind;
        Ns[s] = g.N;
    }
    return PROMP_OK;
}

extern "C" int64_t promp_policy_chain_workspace_bytes(int obs_dim, int act_dim, int hidden, int M, int n_stages,
                                                      const promp_policy_stage* stages) {
    if (stages == nullptr || n_stages < 1 || n_stages > CHAIN_MAX_STAGES || M < 1) return -1;
    int kinds[CHAIN_MAX_STAGES], Ns[CHAIN_MAX_STAGES];
    for (int s = 0; s < n_stages; ++s) kinds[s] = stages[s].kind, Ns[s] = stages[s].N > 0 ? stages[s].N : 1;
    PROMP_DISPATCH_DIMS(chain_ws_bytes, n_stages, kinds, Ns, M)
}

extern "C" int promp_policy_chain_num_launches(int obs_dim, int act_dim, int hidden, int M, int n_stages,
                                               const promp_policy_stage* stages) {
    if (stages == nullptr || n_stages < 1 || n_stages > CHAIN_MAX_STAGES || M < 1) return -1;
    int kinds[CHAIN_MAX_STAGES], Ns[CHAIN_MAX_STAGES];
    for (int s = 0; s < n_stages; ++s) kinds[s] = stages[s].kind, Ns[s] = stages[s].N > 0 ? stages[s].N : 1;
    PROMP_DISPATCH_DIMS(chain_num_launches, n_stages, kinds, Ns, M)
}

extern "C" int promp_policy_chain(int obs_dim, int act_dim, int hidden, int M, float min_log_std, int n_stages,
                                  const promp_policy_stage* stages, const int32_t* skip_flag, const float* skip_theta,
                                  void* workspace, int64_t workspace_bytes, void* stream) {
    PROMP_REQUIRE(M > 0 && workspace != nullptr, "promp_policy_chain: M must be positive and workspace non-null");
    PROMP_REQUIRE((skip_flag == nullptr) == (skip_theta == nullptr), "promp_policy_chain: skip_flag / skip_theta come as a pair");
    PolicyArgs A[CHAIN_MAX_STAGES];
    int kinds[CHAIN_MAX_STAGES], Ns[CHAIN_MAX_STAGES];
    const int rc = chain_stage_args(stages, n_stages, M, min_log_std, A, kinds, Ns);
    if (rc != PROMP_OK) return rc;
    PROMP_REQUIRE(!skip_flag || (kinds[0] == 0 && A[0].param_stride == 0),
                  "promp_policy_chain: launch re-use is defined for a gradient stage 0 on shared parameters (param_stride 0)");
    cudaStream_t s = (cudaStream_t)stream;
    PROMP_DISPATCH_DIMS(launch_chain, n_stages, kinds, A, skip_flag, skip_theta, workspace, workspace_bytes, s)
}

extern "C" int promp_set_option(const char* name, int value) {
    PROMP_REQUIRE(name != nullptr, "promp_set_option: null name");
    if (strcmp(name, "tensor_cores") == 0) {
        g_use_tc = value ? 1 : 0;
        return PROMP_OK;
    }
    if (strcmp(name, "chain") == 0) {          // promp_policy_chain: dataflow kernel (1), one launch per stage (0), automatic (-1)
        g_chain = value < 0 ? -1 : (value ? 1 : 0);
        return PROMP_OK;
    }
    if (strcmp(name, "chain_q") == 0) {        // tiles per work item of the dataflow kernel (0 = automatic)
        PROMP_REQUIRE(value >= 0 && value <= 64, "promp_set_option: chain_q must be in [0, 64]");
        g_chain_q = value;
        return PROMP_OK;
    }
    if (strcmp(name, "chain_taper") == 0) {    // shrinking items at the end of the last stage (1, default) or uniform (0)
        g_chain_taper = value ? 1 : 0;
        return PROMP_OK;
    }
    if (strcmp(name, "tc_threads") == 0) {
        PROMP_REQUIRE(value == 0 || value == 256 || value == 512, "promp_set_option: tc_threads must be 0, 256 or 512");
        g_tc_threads = value;
        return PROMP_OK;
    }
    set_error("promp_set_option: unknown option '%s'", name);
    return PROMP_ERR_INVALID_ARG;
}

extern "C" int promp_policy_forward(int obs_dim, int act_dim, int hidden, int M, int N, const float* params,
                                    int64_t param_stride, const float* obs, float* mean, void* stream) {
    PROMP_REQUIRE(M > 0 && N > 0 && params && obs && mean, "promp_policy_forward: bad arguments");
    PROMP_REQUIRE(M <= 65535, "promp_policy_forward: M=%d exceeds the grid.y limit", M);
    cudaStream_t s = (cudaStream_t)stream;
    PROMP_DISPATCH_DIMS(launch_forward, M, N, params, param_stride, obs, mean, s)
}

extern "C" int promp_reduce_tasks(int M, int P, const float* in, float scale, float* out, void* stream) {
    PROMP_REQUIRE(M > 0 && P > 0 && in && out, "promp_reduce_tasks: bad arguments");
    reduce_tasks_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(M, P, in, scale, out);
    PROMP_LAUNCH_CHECK("reduce_tasks_kernel");
    return PROMP_OK;
}

extern "C" int promp_meta_loss_terms(int S, int M, const float* stats_all, float inv_m_global, const float* coeff, int n_out,
                                     float* out, void* stream) {
    PROMP_REQUIRE(S >= 1 && S <= 7 && M > 0 && stats_all && out && n_out >= 1 && n_out <= S + 1,
                  "promp_meta_loss_terms: bad arguments (1 <= S <= 7 sampling phases, 1 <= n_out <= S+1)");
    meta_loss_terms_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(S, M, stats_all, inv_m_global, coeff, n_out, out);
    PROMP_LAUNCH_CHECK("meta_loss_terms_kernel");
    return PROMP_OK;
}

extern "C" int promp_phase_log_terms(int M, int act_dim, double n_paths, const double* stats, const float* log_std, double* out7,
                                     void* stream) {
    PROMP_REQUIRE(M > 0 && act_dim > 0 && n_paths > 0 && stats && log_std && out7, "promp_phase_log_terms: bad arguments");
    phase_log_terms_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(M, act_dim, n_paths, stats, log_std, out7);
    PROMP_LAUNCH_CHECK("phase_log_terms_kernel");
    return PROMP_OK;
}

extern "C" int promp_promp_log_terms(int num_inner_steps, const float* final_terms, double* out3, void* stream) {
    PROMP_REQUIRE(num_inner_steps >= 0 && final_terms && out3, "promp_promp_log_terms: bad arguments");
    promp_log_terms_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(num_inner_steps, final_terms, out3);
    PROMP_LAUNCH_CHECK("promp_log_terms_kernel");
    return PROMP_OK;
}

extern "C" int promp_adapt_kl_coeff(int num_inner_steps, const float* final_terms, double kl_target, int adapt, float* coeff_dev,
                                    double* out4, void* stream) {
    PROMP_REQUIRE(num_inner_steps >= 0 && num_inner_steps <= 7 && final_terms && (coeff_dev || num_inner_steps == 0),
                  "promp_adapt_kl_coeff: bad arguments");
    adapt_kl_coeff_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(num_inner_steps, final_terms, kl_target, adapt, coeff_dev, out4);
    PROMP_LAUNCH_CHECK("adapt_kl_coeff_kernel");
    return PROMP_OK;
}

extern "C" int promp_adam_tf1(int P, float* theta, const float* grad, float* m, float* v, int32_t* step, float lr,
                              float beta1, float beta2, float eps, void* stream) {
    PROMP_REQUIRE(P > 0 && theta && grad && m && v && step, "promp_adam_tf1: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    adam_tf1_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, theta, grad, m, v, step, lr, beta1, beta2, eps);
    adam_step_inc_kernel<<<1, 1, 0, s>>>(step);
    PROMP_LAUNCH_CHECK("adam_tf1_kernel");
    return PROMP_OK;
}

#ifdef PROMP_EXP_CLOCKS
extern "C" int promp_debug_chain_clocks(unsigned long long* out16, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out16, promp::g_chain_clk, 16 * sizeof(unsigned long long));
    if (reset) {
        unsigned long long z[16] = {0};
        cudaMemcpyToSymbol(promp::g_chain_clk, z, sizeof(z));
    }
    return 0;
}
extern "C" int promp_debug_phase_clocks(unsigned long long* out16, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out16, promp::g_phase_clk, 16 * sizeof(unsigned long long));
    if (reset) {
        unsigned long long z[16] = {0};
        cudaMemcpyToSymbol(promp::g_phase_clk, z, sizeof(z));
    }
    return 0;
}
#endif
