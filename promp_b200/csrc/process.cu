// Sample processing: returns -> linear-feature baseline fit/predict -> GAE -> normalisation -> stats.
// One CTA per task (tasks are independent: the reference re-fits the shared baseline inside its
// task loop, samplers/meta_sample_processor.py:31-34).  The scans, Gram matrix, Cholesky solve and
// moments run in float64 like the reference's numpy/LAPACK path; inputs/outputs are float32.
//
// HBM-bound stage (AI < 1 FLOP/B): per env-step it reads obs (4*Do B) twice + rew (4 B) twice and
// writes returns + advantages (8 B); fp64 intermediates live in an L2-resident workspace.
#include "common.cuh"

namespace promp {

constexpr int PS_THREADS = 256;
constexpr int PS_TS = 32;          // samples per Gram tile
constexpr int PS_MAXCOL = 44;      // F+1 <= 44  (obs_dim <= 19)
constexpr int PS_MAXPAIR = PS_MAXCOL * (PS_MAXCOL + 1) / 2;   // 990
constexpr int PS_MAXITEM = 4;      // pair-items per thread

__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < PS_THREADS / 32; ++i) t += red[i];
    return t;
}
__device__ __forceinline__ double block_max(double v, double* red) {
    v = warp_max(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double t = red[0];
#pragma unroll
    for (int i = 1; i < PS_THREADS / 32; ++i) t = fmax(t, red[i]);
    return t;
}
__device__ __forceinline__ double block_min(double v, double* red) {
    v = warp_min(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double t = red[0];
#pragma unroll
    for (int i = 1; i < PS_THREADS / 32; ++i) t = fmin(t, red[i]);
    return t;
}

struct ProcArgs {
    int M, E, H, Do;
    const float* obs;
    const float* rew;
    double discount, gae_lambda, reg_coeff;
    int baseline_kind, normalize_adv, positive_adv;
    float* returns;
    float* adv;
    double* coeffs;
    double* stats;
    double* ws;   // [M][2][N] float64: returns, baseline->advantages
};

// LinearFeatureBaseline._features (baselines/linear_baseline.py:101-106) for one sample, float64:
//   [clip(o,-10,10), clip(o)^2, t, t^2, t^3, 1] with t = step/100
__device__ __forceinline__ void features(const float* o, int Do, int step, double* f) {
    for (int i = 0; i < Do; ++i) {
        double c = fmin(fmax((double)o[i], -10.0), 10.0);
        f[i] = c;
        f[Do + i] = c * c;
    }
    const double tt = (double)step / 100.0;
    f[2 * Do] = tt;
    f[2 * Do + 1] = tt * tt;
    f[2 * Do + 2] = tt * tt * tt;
    f[2 * Do + 3] = 1.0;
}

__global__ void __launch_bounds__(PS_THREADS) process_kernel(ProcArgs A) {
    const int m = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int E = A.E, H = A.H, Do = A.Do, N = E * H;
    const int F = 2 * Do + 4, NC = F + 1;            // NC columns: features + target
    const int n_pairs = NC * (NC + 1) / 2;
    const float* obs = A.obs + (int64_t)m * N * Do;
    const float* rew = A.rew + (int64_t)m * N;
    double* ret64 = A.ws + (int64_t)m * 2 * N;
    double* adv64 = ret64 + N;

    __shared__ double red[PS_THREADS / 32];
    __shared__ double tile[PS_TS * PS_MAXCOL];
    __shared__ double gram[PS_MAXPAIR];                 // packed upper triangle over NC columns
    __shared__ double Lm[(PS_MAXCOL - 1) * (PS_MAXCOL - 1)];
    __shared__ double wv[PS_MAXCOL];
    __shared__ unsigned char pair_i[PS_MAXPAIR], pair_j[PS_MAXPAIR];
    __shared__ int s_flag;
    __shared__ double s_reg;

    // ---- (1) discounted returns R_t = r_t + g R_{t+1}  (utils/utils.py:74-81) + path statistics
    double sR0 = 0, sG = 0, sG2 = 0, mxG = -1e300, mnG = 1e300, sr = 0, sr2 = 0;
    for (int e = tid; e < E; e += PS_THREADS) {
        double R = 0.0, G = 0.0;
        for (int t = H - 1; t >= 0; --t) {
            const double r = (double)rew[e * H + t];
            R = r + A.discount * R;
            G += r;
            sr += r;
            sr2 += r * r;
            ret64[e * H + t] = R;
        }
        sR0 += R;
        sG += G;
        sG2 += G * G;
        mxG = fmax(mxG, G);
        mnG = fmin(mnG, G);
    }
    sR0 = block_sum(sR0, red); sG = block_sum(sG, red); sG2 = block_sum(sG2, red);
    sr = block_sum(sr, red); sr2 = block_sum(sr2, red);
    mxG = block_max(mxG, red); mnG = block_min(mnG, red);
    __syncthreads();   // ret64 visible to the whole CTA
    for (int n = tid; n < N; n += PS_THREADS) A.returns[(int64_t)m * N + n] = (float)ret64[n];

    double reg_used = 0.0;
    if (A.baseline_kind == PROMP_BASELINE_LINEAR_FEATURE) {
        // ---- (2) Gram matrix Phi^T [Phi | y] in float64 (baselines/linear_baseline.py:66-73)
        for (int p = tid; p < n_pairs; p += PS_THREADS) {   // packed (i<=j) index tables
            int i = 0, rem = p;
            while (rem >= NC - i) { rem -= NC - i; ++i; }
            pair_i[p] = (unsigned char)i;
            pair_j[p] = (unsigned char)(i + rem);
        }
        const int G = max(1, PS_THREADS / n_pairs);        // sample groups per pair
        const int n_items = n_pairs * G;
        double acc[PS_MAXITEM] = {0, 0, 0, 0};
        for (int n0 = 0; n0 < N; n0 += PS_TS) {
            const int ns = min(PS_TS, N - n0);
            __syncthreads();
            for (int s = tid; s < ns; s += PS_THREADS) {     // one thread builds one sample's feature row
                const int n = n0 + s;
                features(obs + (int64_t)n * Do, Do, n % H, &tile[s * NC]);
                tile[s * NC + F] = ret64[n];
            }
            __syncthreads();
#pragma unroll
            for (int it = 0; it < PS_MAXITEM; ++it) {
                const int item = tid + it * PS_THREADS;
                if (item < n_items) {
                    const int p = item % n_pairs, g = item / n_pairs;
                    const int i = pair_i[p], j = pair_j[p];
                    double a = acc[it];
                    for (int s = g; s < ns; s += G) a = fma(tile[s * NC + i], tile[s * NC + j], a);
                    acc[it] = a;
                }
            }
        }
        __syncthreads();
        for (int p = tid; p < n_pairs; p += PS_THREADS) gram[p] = 0.0;
        __syncthreads();
        // deterministic group reduction: groups added in order g = 0..G-1
        for (int g = 0; g < G; ++g) {
#pragma unroll
            for (int it = 0; it < PS_MAXITEM; ++it) {
                const int item = tid + it * PS_THREADS;
                if (item < n_items && item / n_pairs == g) gram[item % n_pairs] += acc[it];
            }
            __syncthreads();
        }

        // ---- (3) solve (Phi^T Phi + reg I) w = Phi^T y; retry with 10x reg on NaN, up to 5 tries (:68-77)
        if (tid == 0) { s_reg = A.reg_coeff; s_flag = 0; }
        __syncthreads();
        for (int attempt = 0; attempt < 5; ++attempt) {
            const double reg = s_reg;
            // unpack A into Lm (lower triangle incl. diagonal)
            for (int p = tid; p < n_pairs; p += PS_THREADS) {
                const int i = pair_i[p], j = pair_j[p];
                if (j < F) Lm[j * F + i] = gram[p] + (i == j ? reg : 0.0);
            }
            __syncthreads();
            if (tid < 32) {
                bool bad = false;
                // Cholesky A = L L^T, column by column; lanes own rows
                for (int j = 0; j < F; ++j) {
                    double d = 0.0;
                    for (int k = lane; k < j; k += 32) d += Lm[j * F + k] * Lm[j * F + k];
                    d = Lm[j * F + j] - warp_sum(d);
                    const double ljj = sqrt(d);     // d <= 0 -> NaN -> retry with a larger ridge
                    if (!(ljj > 0.0)) bad = true;
                    __syncwarp();
                    if (lane == 0) Lm[j * F + j] = ljj;
                    for (int i = j + 1 + lane; i < F; i += 32) {
                        double s = Lm[i * F + j];
                        for (int k = 0; k < j; ++k) s -= Lm[i * F + k] * Lm[j * F + k];
                        Lm[i * F + j] = s / ljj;
                    }
                    __syncwarp();
                }
                // forward: L z = b ; backward: L^T w = z   (b = Phi^T y = gram column F)
                if (lane == 0) {
                    for (int i = 0; i < F; ++i) {
                        // packed index of (i, F): row i starts at i*NC - i(i-1)/2, column offset F-i
                        double s = gram[i * NC - i * (i - 1) / 2 + (F - i)];
                        for (int k = 0; k < i; ++k) s -= Lm[i * F + k] * wv[k];
                        wv[i] = s / Lm[i * F + i];
                    }
                    for (int i = F - 1; i >= 0; --i) {
                        double s = wv[i];
                        for (int k = i + 1; k < F; ++k) s -= Lm[k * F + i] * wv[k];
                        wv[i] = s / Lm[i * F + i];
                    }
                    for (int i = 0; i < F; ++i)
                        if (isnan(wv[i]) || isinf(wv[i])) bad = true;
                }
                bad = __any_sync(0xffffffffu, bad);
                if (lane == 0) {
                    s_flag = bad ? 0 : 1;
                    if (bad) s_reg = reg * 10.0;
                }
            }
            __syncthreads();
            reg_used = reg;
            if (s_flag) break;
        }
        if (A.coeffs)
            for (int i = tid; i < F; i += PS_THREADS) A.coeffs[(int64_t)m * F + i] = wv[i];
        // ---- (4) predict b_n = phi_n . w (baselines/linear_baseline.py:17-33)
        for (int n = tid; n < N; n += PS_THREADS) {
            double f[PS_MAXCOL];
            features(obs + (int64_t)n * Do, Do, n % H, f);
            double b = 0.0;
            for (int i = 0; i < F; ++i) b = fma(f[i], wv[i], b);
            adv64[n] = b;
        }
    } else {
        for (int n = tid; n < N; n += PS_THREADS) adv64[n] = 0.0;   // ZeroBaseline.predict
    }
    __syncthreads();

    // ---- (5) GAE: delta_t = r_t + g b_{t+1} - b_t (b_H = 0); A_t = delta_t + g*lam A_{t+1}  (samplers/base.py:151-162)
    const double gl = A.discount * A.gae_lambda;
    double s1 = 0.0;
    for (int e = tid; e < E; e += PS_THREADS) {
        double b_next = 0.0, a_next = 0.0;
        for (int t = H - 1; t >= 0; --t) {
            const double b = adv64[e * H + t];
            const double delta = (double)rew[e * H + t] + A.discount * b_next - b;
            const double a = delta + gl * a_next;
            adv64[e * H + t] = a;
            s1 += a;
            a_next = a;
            b_next = b;
        }
    }
    __syncthreads();

    // ---- (6) per-task normalisation / positive shift (utils/utils.py:59-71; population std)
    double mean = 0.0, inv = 1.0;
    if (A.normalize_adv) {
        mean = block_sum(s1, red) / (double)N;
        double s2 = 0.0;
        for (int n = tid; n < N; n += PS_THREADS) {
            const double d = adv64[n] - mean;
            s2 += d * d;
        }
        const double var = block_sum(s2, red) / (double)N;
        inv = 1.0 / (sqrt(var) + 1e-8);
    }
    double mn = 0.0;
    if (A.positive_adv) {
        double lm = 1e300;
        for (int n = tid; n < N; n += PS_THREADS) lm = fmin(lm, (adv64[n] - mean) * inv);
        mn = block_min(lm, red);
    }
    for (int n = tid; n < N; n += PS_THREADS) {
        double a = (adv64[n] - mean) * inv;
        if (A.positive_adv) a = (a - mn) + 1e-8;
        A.adv[(int64_t)m * N + n] = (float)a;
    }

    if (A.stats && tid == 0) {
        double* st = A.stats + (int64_t)m * 8;
        st[0] = sR0; st[1] = sG; st[2] = sG2; st[3] = mxG; st[4] = mnG; st[5] = sr; st[6] = sr2; st[7] = reg_used;
    }
}

__global__ void adj_avg_rewards_kernel(int64_t n, const float* rew, double mean, double inv, float* out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)(((double)rew[i] - mean) * inv);
}

}  // namespace promp

using namespace promp;

extern "C" int64_t promp_process_workspace_bytes(int M, int E, int H, int obs_dim) {
    (void)obs_dim;
    return (int64_t)M * 2 * E * H * (int64_t)sizeof(double);
}

extern "C" int promp_process_samples(int M, int E, int H, int obs_dim, const float* obs, const float* rew,
                                     double discount, double gae_lambda, double reg_coeff, int baseline_kind,
                                     int normalize_adv, int positive_adv, float* returns, float* adv, double* coeffs,
                                     double* stats, void* workspace, int64_t workspace_bytes, void* stream) {
    PROMP_REQUIRE(M > 0 && E > 0 && H > 0 && obs_dim > 0, "promp_process_samples: dimensions must be positive");
    PROMP_REQUIRE(2 * obs_dim + 5 <= PS_MAXCOL, "promp_process_samples: obs_dim %d too large (max %d)", obs_dim,
                  (PS_MAXCOL - 5) / 2);
    PROMP_REQUIRE(obs && rew && returns && adv && workspace, "promp_process_samples: null pointer argument");
    PROMP_REQUIRE(baseline_kind == PROMP_BASELINE_ZERO || baseline_kind == PROMP_BASELINE_LINEAR_FEATURE,
                  "promp_process_samples: unknown baseline kind %d", baseline_kind);
    PROMP_REQUIRE(discount >= 0.0 && discount <= 1.0 && gae_lambda >= 0.0 && gae_lambda <= 1.0,
                  "promp_process_samples: discount and gae_lambda must be in [0,1]");   // samplers/base.py:56-57
    if (workspace_bytes < promp_process_workspace_bytes(M, E, H, obs_dim)) {
        set_error("promp_process_samples: workspace too small (%lld < %lld bytes)", (long long)workspace_bytes,
                  (long long)promp_process_workspace_bytes(M, E, H, obs_dim));
        return PROMP_ERR_WORKSPACE;
    }
    ProcArgs A{M, E, H, obs_dim, obs, rew, discount, gae_lambda, reg_coeff, baseline_kind, normalize_adv, positive_adv,
               returns, adv, coeffs, stats, (double*)workspace};
    process_kernel<<<M, PS_THREADS, 0, (cudaStream_t)stream>>>(A);
    PROMP_LAUNCH_CHECK("process_kernel");
    return PROMP_OK;
}

extern "C" int promp_adj_avg_rewards(int64_t n, const float* rew, double mean, double std, float* out, void* stream) {
    PROMP_REQUIRE(n > 0 && rew && out, "promp_adj_avg_rewards: bad arguments");
    const int bs = 256;
    adj_avg_rewards_kernel<<<(unsigned)((n + bs - 1) / bs), bs, 0, (cudaStream_t)stream>>>(n, rew, mean,
                                                                                         1.0 / (std + 1e-8), out);
    PROMP_LAUNCH_CHECK("adj_avg_rewards_kernel");
    return PROMP_OK;
}
