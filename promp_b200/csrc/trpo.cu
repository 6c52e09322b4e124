// Device-resident pieces of ConjugateGradientOptimizer (ref: optimizers/conjugate_gradient_optimizer.py:239-354) so that a
// TRPO-MAML outer step needs no host round trip until the line-search verdict:
//   * candidate parameter vectors theta +- eps*p / theta - ratio*step                       (vec_axpy)
//   * the CG iteration with the finite-difference Hx(p) = (dKL(theta+eps p) - dKL(theta-eps p)) / (2 eps) fused into the
//     update (:59-89, :325-354): z, v = rdotr / p.z, x += v p, r -= v z, mu, p = r + mu p       (cg_init / cg_step)
//   * the initial step length beta = sqrt(2 delta / (x.Hx(x) + 1e-8)) and step = beta x (:262-268)   (trpo_step)
//   * the backtracking verdict over a group of speculatively evaluated candidates (:274-300)         (trpo_select)
// All vectors are the flat float32 parameter layout (P = 4.5-5.7 k): one CTA, dot products accumulated in float64 in a
// fixed order (deterministic, identical on every rank), scalars kept in a small device array.
#include "common.cuh"

namespace promp {

constexpr int TV_THREADS = 1024;

__device__ __forceinline__ double cta_sum(double v, double* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < TV_THREADS / 32; ++i) t += red[i];
    return t;
}

// out = y + a*x with numpy's two roundings (float32 multiply, then float32 add)
__global__ void vec_axpy_kernel(int n, float a, const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __fadd_rn(y[i], __fmul_rn(a, x[i]));
}

// scal: [0] rdotr  [1] done (residual below tolerance)  [2] beta  [3] step is NaN  (float32 values in a float array)
__global__ void __launch_bounds__(TV_THREADS) cg_init_kernel(int n, const float* __restrict__ g, float* __restrict__ p,
                                                             float* __restrict__ r, float* __restrict__ x, float* scal) {
    __shared__ double red[TV_THREADS / 32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += TV_THREADS) {
        const float gi = g[i];
        p[i] = gi;
        r[i] = gi;
        x[i] = 0.f;
        acc += (double)gi * (double)gi;
    }
    const double rr = cta_sum(acc, red);
    if (threadIdx.x == 0) { scal[0] = (float)rr; scal[1] = 0.f; scal[2] = 0.f; scal[3] = 0.f; }
}

__global__ void __launch_bounds__(TV_THREADS) cg_step_kernel(int n, const float* __restrict__ gp, const float* __restrict__ gm,
                                                             float two_eps, float reg, float* __restrict__ p,
                                                             float* __restrict__ r, float* __restrict__ x, float* scal,
                                                             float residual_tol) {
    __shared__ double red[TV_THREADS / 32];
    if (scal[1] != 0.f) return;                       // converged earlier (the reference's `break`): later steps are no-ops
    const float rdotr = scal[0];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += TV_THREADS) {
        const float z = __fadd_rn(__fdiv_rn(__fsub_rn(gp[i], gm[i]), two_eps), __fmul_rn(reg, p[i]));
        acc += (double)p[i] * (double)z;
    }
    const float pz = (float)cta_sum(acc, red);
    const float v = __fdiv_rn(rdotr, pz);
    acc = 0.0;
    for (int i = threadIdx.x; i < n; i += TV_THREADS) {
        const float z = __fadd_rn(__fdiv_rn(__fsub_rn(gp[i], gm[i]), two_eps), __fmul_rn(reg, p[i]));
        x[i] = __fadd_rn(x[i], __fmul_rn(v, p[i]));
        const float ri = __fsub_rn(r[i], __fmul_rn(v, z));
        r[i] = ri;
        acc += (double)ri * (double)ri;
    }
    const float newrdotr = (float)cta_sum(acc, red);
    const float mu = __fdiv_rn(newrdotr, rdotr);
    for (int i = threadIdx.x; i < n; i += TV_THREADS) p[i] = __fadd_rn(r[i], __fmul_rn(mu, p[i]));
    if (threadIdx.x == 0) {
        scal[0] = newrdotr;
        if (newrdotr < residual_tol) scal[1] = 1.f;
    }
}

__global__ void __launch_bounds__(TV_THREADS) trpo_step_kernel(int n, const float* __restrict__ gp, const float* __restrict__ gm,
                                                               float two_eps, float reg, const float* __restrict__ x,
                                                               float delta, float* __restrict__ step, float* scal) {
    __shared__ double red[TV_THREADS / 32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += TV_THREADS) {
        const float z = __fadd_rn(__fdiv_rn(__fsub_rn(gp[i], gm[i]), two_eps), __fmul_rn(reg, x[i]));
        acc += (double)x[i] * (double)z;
    }
    const double xHx = (double)(float)cta_sum(acc, red);
    // np.sqrt(2.0 * delta * (1. / (x.dot(Hx(x)) + 1e-8))) - float64 arithmetic on a float32 dot product (:264-265)
    const double beta = sqrt(2.0 * (double)delta * (1.0 / (xHx + 1e-8)));
    const float b32 = (float)beta;
    for (int i = threadIdx.x; i < n; i += TV_THREADS) step[i] = __fmul_rn(b32, x[i]);
    if (threadIdx.x == 0) { scal[2] = b32; scal[3] = (beta != beta) ? 1.f : 0.f; }
}

// Verdict over candidates k0..k0+K-1 (cands [K][n] = theta_prev - ratio^k * step), whose [loss, ..., kl] rows sit in
// `terms` [K][T].
// result: [0] loss_before [1] kl_before [2] loss_after [3] kl_after [4] accepted k (-1: none) [5] rejected [6] need_more
//         [7] beta.  base: [loss_before, ..., kl_before] row (T floats).
__global__ void __launch_bounds__(TV_THREADS) trpo_select_kernel(int n, int K, int T, int k0, int k_max, const float* terms,
                                                                 const float* base, float delta,
                                                                 const float* __restrict__ theta_prev,
                                                                 const float* __restrict__ cands, const float* scal,
                                                                 float* __restrict__ theta_out, float* result) {
    __shared__ int s_k, s_reject, s_more;
    if (threadIdx.x == 0) {
        const float loss_before = base[0], kl_before = base[T - 1];
        int acc = -1, reject = 0, more = 0;
        float la = loss_before, ka = kl_before;
        if (scal[3] != 0.f) {
            reject = 1;                                    // "Initial step size is NaN! Rejecting the step!" (:266-268)
        } else {
            for (int k = 0; k < K; ++k) {
                const float l = terms[k * T], c = terms[k * T + T - 1];
                if (l < loss_before && c <= delta) { acc = k0 + k; la = l; ka = c; break; }       // (:283-285)
            }
            if (acc >= 0) {
                if (!(la < loss_before) || !(ka < delta) || la != la || ka != ka) reject = 1;       // violated (:286-297)
            } else if (k0 + K >= k_max) {
                reject = 1;                                // budget exhausted: line search violated, parameters restored
            } else {
                more = 1;
            }
        }
        if (!(acc >= 0 && !reject) && !more) {             // rejected: parameters restored, loss / KL are the old ones
            la = loss_before;
            ka = kl_before;
        }
        s_k = acc; s_reject = reject; s_more = more;
        result[0] = loss_before; result[1] = kl_before; result[2] = la; result[3] = ka;
        result[4] = (float)acc; result[5] = (float)reject; result[6] = (float)more; result[7] = scal[2];
    }
    __syncthreads();
    if (s_more) return;                                    // undecided: theta_out untouched
    const bool take = s_k >= 0 && !s_reject;
    const float* src = take ? cands + (int64_t)(s_k - k0) * n : theta_prev;      // the very vector that was evaluated
    for (int i = threadIdx.x; i < n; i += TV_THREADS) theta_out[i] = src[i];
}

}  // namespace promp

using namespace promp;

extern "C" int promp_vec_axpy(int n, float a, const float* x, const float* y, float* out, void* stream) {
    PROMP_REQUIRE(n > 0 && x && y && out, "promp_vec_axpy: bad arguments");
    vec_axpy_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, a, x, y, out);
    PROMP_LAUNCH_CHECK("vec_axpy_kernel");
    return PROMP_OK;
}

extern "C" int promp_cg_init(int n, const float* g, float* p, float* r, float* x, float* scal, void* stream) {
    PROMP_REQUIRE(n > 0 && g && p && r && x && scal, "promp_cg_init: bad arguments");
    cg_init_kernel<<<1, TV_THREADS, 0, (cudaStream_t)stream>>>(n, g, p, r, x, scal);
    PROMP_LAUNCH_CHECK("cg_init_kernel");
    return PROMP_OK;
}

extern "C" int promp_cg_step(int n, const float* grad_plus, const float* grad_minus, float two_eps, float reg_coeff, float* p,
                             float* r, float* x, float* scal, float residual_tol, void* stream) {
    PROMP_REQUIRE(n > 0 && grad_plus && grad_minus && p && r && x && scal, "promp_cg_step: bad arguments");
    PROMP_REQUIRE(two_eps != 0.f, "promp_cg_step: eps must be non-zero");
    cg_step_kernel<<<1, TV_THREADS, 0, (cudaStream_t)stream>>>(n, grad_plus, grad_minus, two_eps, reg_coeff, p, r, x, scal,
                                                               residual_tol);
    PROMP_LAUNCH_CHECK("cg_step_kernel");
    return PROMP_OK;
}

extern "C" int promp_trpo_step(int n, const float* grad_plus, const float* grad_minus, float two_eps, float reg_coeff,
                               const float* x, float max_constraint, float* step, float* scal, void* stream) {
    PROMP_REQUIRE(n > 0 && grad_plus && grad_minus && x && step && scal, "promp_trpo_step: bad arguments");
    PROMP_REQUIRE(two_eps != 0.f, "promp_trpo_step: eps must be non-zero");
    trpo_step_kernel<<<1, TV_THREADS, 0, (cudaStream_t)stream>>>(n, grad_plus, grad_minus, two_eps, reg_coeff, x, max_constraint,
                                                                 step, scal);
    PROMP_LAUNCH_CHECK("trpo_step_kernel");
    return PROMP_OK;
}

extern "C" int promp_trpo_select(int n, int n_candidates, int n_terms, int k0, int max_backtracks, const float* terms,
                                 const float* base_terms, float max_constraint, const float* theta_prev,
                                 const float* candidates, const float* scal, float* theta_out, float* result, void* stream) {
    PROMP_REQUIRE(n > 0 && n_candidates > 0 && n_terms >= 2 && k0 >= 0, "promp_trpo_select: bad sizes");
    PROMP_REQUIRE(terms && base_terms && theta_prev && candidates && scal && theta_out && result, "promp_trpo_select: null pointer argument");
    trpo_select_kernel<<<1, TV_THREADS, 0, (cudaStream_t)stream>>>(n, n_candidates, n_terms, k0, max_backtracks, terms, base_terms,
                                                                   max_constraint, theta_prev, candidates, scal, theta_out, result);
    PROMP_LAUNCH_CHECK("trpo_select_kernel");
    return PROMP_OK;
}
