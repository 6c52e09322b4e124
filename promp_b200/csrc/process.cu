// Sample processing: returns -> linear-feature baseline fit/predict -> GAE -> normalisation -> stats.
// Tasks are independent (the reference re-fits the shared baseline inside its task loop,
// samplers/meta_sample_processor.py:31-34); a task is split over C CTAs for the returns + Gram stage and
// finished by one CTA.  The scans, Gram matrix, Cholesky solve and
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
    double* ws;      // [M][2][N] float64 (returns, baseline->advantages), then partials
    double* gram_p;  // [M][C][PS_MAXPAIR] partial Gram matrices
    double* stat_p;  // [M][C][8] partial path statistics
    int C, EPC;      // trajectory chunks per task, trajectories per chunk
    // variable-length paths (early termination, meta_sampler.py:116-125); path_off == nullptr: E paths of H steps each
    const int32_t* path_off;   // [M][Pmax+1] sample offset of every path inside its task (prefix sums), E := Pmax
    const int32_t* n_paths;    // [M] number of paths of each task (<= Pmax)
    int NS;                    // sample stride between tasks (E*H, or Nmax for variable-length paths)
    int32_t* tpos;             // [M][NS] workspace: time index of every sample inside its path (variable-length only)
};
// path table helpers: number of paths of task m, offset of path e
__device__ __forceinline__ int n_paths_of(const ProcArgs& A, int m) { return A.path_off ? __ldg(A.n_paths + m) : A.E; }
__device__ __forceinline__ int path_begin(const ProcArgs& A, int m, int e) {
    return A.path_off ? __ldg(A.path_off + (int64_t)m * (A.E + 1) + e) : e * A.H;
}

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

__device__ __forceinline__ void pair_tables(int NC, int n_pairs, unsigned char* pair_i, unsigned char* pair_j) {
    for (int p = threadIdx.x; p < n_pairs; p += PS_THREADS) {   // packed (i<=j) index tables
        int i = 0, rem = p;
        while (rem >= NC - i) { rem -= NC - i; ++i; }
        pair_i[p] = (unsigned char)i;
        pair_j[p] = (unsigned char)(i + rem);
    }
}

// Stage 1, grid (C, M): CTA (c, m) owns trajectories [c*EPC, (c+1)*EPC) of task m: discounted returns, path
// statistics, and its slice of the Gram matrix Phi^T [Phi | y] in float64.  Splitting a task over C CTAs puts
// ~2 CTAs on every SM (a task-per-CTA launch uses only M of the 148 SMs and is fp64-FMA bound for F = 38).
__global__ void __launch_bounds__(PS_THREADS) process_gram_kernel(ProcArgs A) {
    const int c = blockIdx.x, m = blockIdx.y, tid = threadIdx.x;
    const int E = n_paths_of(A, m), H = A.H, Do = A.Do, NS = A.NS;
    const int F = 2 * Do + 4, NC = F + 1;
    const int n_pairs = NC * (NC + 1) / 2;
    const int e_lo = min(E, c * A.EPC), e_hi = min(E, e_lo + A.EPC);
    const float* obs = A.obs + (int64_t)m * NS * Do;
    const float* rew = A.rew + (int64_t)m * NS;
    double* ret64 = A.ws + (int64_t)m * 2 * NS;
    int32_t* tpos = A.tpos ? A.tpos + (int64_t)m * NS : nullptr;

    __shared__ double red[PS_THREADS / 32];
    __shared__ double tile[PS_TS * PS_MAXCOL];
    __shared__ double gram[PS_MAXPAIR];
    __shared__ unsigned char pair_i[PS_MAXPAIR], pair_j[PS_MAXPAIR];

    // ---- discounted returns R_t = r_t + g R_{t+1}  (utils/utils.py:74-81) + path statistics
    double sR0 = 0, sG = 0, sG2 = 0, mxG = -1e300, mnG = 1e300, sr = 0, sr2 = 0;
    for (int e = e_lo + tid; e < e_hi; e += PS_THREADS) {
        double R = 0.0, G = 0.0;
        const int o = path_begin(A, m, e), L = path_begin(A, m, e + 1) - o;
        for (int t = L - 1; t >= 0; --t) {
            const double r = (double)rew[o + t];
            R = r + A.discount * R;
            G += r;
            sr += r;
            sr2 += r * r;
            ret64[o + t] = R;
            if (tpos) tpos[o + t] = t;
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
    if (tid == 0) {
        double* sp = A.stat_p + ((int64_t)m * A.C + c) * 8;
        sp[0] = sR0; sp[1] = sG; sp[2] = sG2; sp[3] = mxG; sp[4] = mnG; sp[5] = sr; sp[6] = sr2; sp[7] = 0.0;
    }
    __syncthreads();   // ret64 of this CTA's trajectories visible to the whole CTA
    const int n_lo = path_begin(A, m, e_lo), n_hi = path_begin(A, m, e_hi);
    for (int n = n_lo + tid; n < n_hi; n += PS_THREADS) A.returns[(int64_t)m * NS + n] = (float)ret64[n];
    if (A.baseline_kind != PROMP_BASELINE_LINEAR_FEATURE) return;

    // ---- partial Gram matrix over this CTA's samples (baselines/linear_baseline.py:66-73)
    pair_tables(NC, n_pairs, pair_i, pair_j);
    const int G = max(1, PS_THREADS / n_pairs);        // sample groups per pair
    const int n_items = n_pairs * G;
    double acc[PS_MAXITEM] = {0, 0, 0, 0};
    for (int n0 = n_lo; n0 < n_hi; n0 += PS_TS) {
        const int ns = min(PS_TS, n_hi - n0);
        __syncthreads();
        for (int s = tid; s < ns; s += PS_THREADS) {     // one thread builds one sample's feature row
            const int n = n0 + s;
            features(obs + (int64_t)n * Do, Do, tpos ? tpos[n] : n % H, &tile[s * NC]);
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
    for (int g = 0; g < G; ++g) {      // deterministic group reduction: groups added in order g = 0..G-1
#pragma unroll
        for (int it = 0; it < PS_MAXITEM; ++it) {
            const int item = tid + it * PS_THREADS;
            if (item < n_items && item / n_pairs == g) gram[item % n_pairs] += acc[it];
        }
        __syncthreads();
    }
    double* gp = A.gram_p + ((int64_t)m * A.C + c) * PS_MAXPAIR;
    for (int p = tid; p < n_pairs; p += PS_THREADS) gp[p] = gram[p];
}

// Stage 2, one CTA per task: reduce the partials (fixed chunk order), Cholesky solve with the reference's
// ridge / NaN-retry rule, predict, GAE scan, per-task moments, advantages.
__global__ void __launch_bounds__(PS_THREADS) process_finish_kernel(ProcArgs A) {
    const int m = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int E = n_paths_of(A, m), H = A.H, Do = A.Do, NS = A.NS;
    const int N = path_begin(A, m, E);                 // valid samples of this task
    const int F = 2 * Do + 4, NC = F + 1;
    const int n_pairs = NC * (NC + 1) / 2;
    const float* obs = A.obs + (int64_t)m * NS * Do;
    const float* rew = A.rew + (int64_t)m * NS;
    double* adv64 = A.ws + (int64_t)m * 2 * NS + NS;
    const int32_t* tpos = A.tpos ? A.tpos + (int64_t)m * NS : nullptr;

    __shared__ double red[PS_THREADS / 32];
    __shared__ double gram[PS_MAXPAIR];                 // packed upper triangle over NC columns
    __shared__ double Lm[(PS_MAXCOL - 1) * (PS_MAXCOL - 1)];
    __shared__ double wv[PS_MAXCOL];
    __shared__ unsigned char pair_i[PS_MAXPAIR], pair_j[PS_MAXPAIR];
    __shared__ int s_flag;
    __shared__ double s_reg;

    if (tid < 8) {   // path statistics: sums for 0,1,2,5,6; max for 3; min for 4
        const double* sp = A.stat_p + (int64_t)m * A.C * 8 + tid;
        double v = sp[0];
        for (int c = 1; c < A.C; ++c) {
            const double x = sp[c * 8];
            v = (tid == 3) ? fmax(v, x) : (tid == 4) ? fmin(v, x) : v + x;
        }
        if (A.stats && tid < 7) A.stats[(int64_t)m * 8 + tid] = v;
    }

    double reg_used = 0.0;
    if (A.baseline_kind == PROMP_BASELINE_LINEAR_FEATURE) {
        pair_tables(NC, n_pairs, pair_i, pair_j);
        for (int p = tid; p < n_pairs; p += PS_THREADS) {
            const double* gp = A.gram_p + (int64_t)m * A.C * PS_MAXPAIR + p;
            double v = 0.0;
            for (int c = 0; c < A.C; ++c) v += gp[(int64_t)c * PS_MAXPAIR];
            gram[p] = v;
        }
        __syncthreads();

        // ---- solve (Phi^T Phi + reg I) w = Phi^T y; retry with 10x reg on NaN, up to 5 tries (:68-77)
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
        // ---- predict b_n = phi_n . w (baselines/linear_baseline.py:17-33)
        for (int n = tid; n < N; n += PS_THREADS) {
            double f[PS_MAXCOL];
            features(obs + (int64_t)n * Do, Do, tpos ? tpos[n] : n % H, f);
            double b = 0.0;
            for (int i = 0; i < F; ++i) b = fma(f[i], wv[i], b);
            adv64[n] = b;
        }
    } else {
        for (int n = tid; n < N; n += PS_THREADS) adv64[n] = 0.0;   // ZeroBaseline.predict
    }
    __syncthreads();

    // ---- GAE: delta_t = r_t + g b_{t+1} - b_t (b_H = 0); A_t = delta_t + g*lam A_{t+1}  (samplers/base.py:151-162)
    const double gl = A.discount * A.gae_lambda;
    double s1 = 0.0;
    for (int e = tid; e < E; e += PS_THREADS) {
        double b_next = 0.0, a_next = 0.0;
        const int o = path_begin(A, m, e), L = path_begin(A, m, e + 1) - o;
        for (int t = L - 1; t >= 0; --t) {
            const double b = adv64[o + t];
            const double delta = (double)rew[o + t] + A.discount * b_next - b;
            const double a = delta + gl * a_next;
            adv64[o + t] = a;
            s1 += a;
            a_next = a;
            b_next = b;
        }
    }
    __syncthreads();

    // ---- per-task normalisation / positive shift (utils/utils.py:59-71; population std)
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
        A.adv[(int64_t)m * NS + n] = (float)a;
    }
    for (int n = N + tid; n < NS; n += PS_THREADS) A.adv[(int64_t)m * NS + n] = 0.f;      // padding rows (variable-length paths)
    if (A.stats && tid == 0) A.stats[(int64_t)m * 8 + 7] = reg_used;
}

__global__ void adj_avg_rewards_kernel(int64_t n, const float* rew, double mean, double inv, float* out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)(((double)rew[i] - mean) * inv);
}

}  // namespace promp

using namespace promp;

static void proc_chunks(int M, int E, int* C, int* EPC) {
    int target = (2 * 148 + M - 1) / M;          // ~2 CTAs per SM
    if (target > E) target = E;
    if (target < 1) target = 1;
    *EPC = (E + target - 1) / target;
    *C = (E + *EPC - 1) / *EPC;
}

extern "C" int64_t promp_process_workspace_bytes(int M, int E, int H, int obs_dim) {
    (void)obs_dim;
    int C, EPC;
    proc_chunks(M, E, &C, &EPC);
    return ((int64_t)M * 2 * E * H + (int64_t)M * C * (PS_MAXPAIR + 8)) * (int64_t)sizeof(double);
}

extern "C" int promp_process_samples(int M, int E, int H, int obs_dim, const float* obs, const float* rew,
                                     double discount, double gae_lambda, double reg_coeff, int baseline_kind,
                                     int normalize_adv, int positive_adv, float* returns, float* adv, double* coeffs,
                                     double* stats, void* workspace, int64_t workspace_bytes, void* stream) {
    PROMP_REQUIRE(M > 0 && E > 0 && H > 0 && obs_dim > 0, "promp_process_samples: dimensions must be positive");
    PROMP_REQUIRE(M <= 65535, "promp_process_samples: M=%d exceeds the grid.y limit", M);
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
    int C, EPC;
    proc_chunks(M, E, &C, &EPC);
    double* ws = (double*)workspace;
    double* gram_p = ws + (int64_t)M * 2 * E * H;
    double* stat_p = gram_p + (int64_t)M * C * PS_MAXPAIR;
    ProcArgs A{M, E, H, obs_dim, obs, rew, discount, gae_lambda, reg_coeff, baseline_kind, normalize_adv, positive_adv,
               returns, adv, coeffs, stats, ws, gram_p, stat_p, C, EPC, nullptr, nullptr, E * H, nullptr};
    process_gram_kernel<<<dim3(C, M), PS_THREADS, 0, (cudaStream_t)stream>>>(A);
    PROMP_LAUNCH_CHECK("process_gram_kernel");
    process_finish_kernel<<<M, PS_THREADS, 0, (cudaStream_t)stream>>>(A);
    PROMP_LAUNCH_CHECK("process_finish_kernel");
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

// ---- variable-length paths: same kernels driven by a per-task path table ------------------------------------------
extern "C" int64_t promp_process_workspace_bytes_ragged(int M, int max_paths, int max_samples, int obs_dim) {
    (void)obs_dim;
    int C, EPC;
    proc_chunks(M, max_paths, &C, &EPC);
    return ((int64_t)M * 2 * max_samples + (int64_t)M * C * (PS_MAXPAIR + 8)) * (int64_t)sizeof(double) +
           (int64_t)M * max_samples * (int64_t)sizeof(int32_t);
}

extern "C" int promp_process_samples_ragged(int M, int max_paths, int max_samples, int obs_dim, const float* obs,
                                            const float* rew, const int32_t* path_off, const int32_t* n_paths, double discount,
                                            double gae_lambda, double reg_coeff, int baseline_kind, int normalize_adv,
                                            int positive_adv, float* returns, float* adv, double* coeffs, double* stats,
                                            void* workspace, int64_t workspace_bytes, void* stream) {
    PROMP_REQUIRE(M > 0 && max_paths > 0 && max_samples > 0 && obs_dim > 0, "promp_process_samples_ragged: dimensions must be positive");
    PROMP_REQUIRE(M <= 65535, "promp_process_samples_ragged: M=%d exceeds the grid.y limit", M);
    PROMP_REQUIRE(2 * obs_dim + 5 <= PS_MAXCOL, "promp_process_samples_ragged: obs_dim %d too large (max %d)", obs_dim,
                  (PS_MAXCOL - 5) / 2);
    PROMP_REQUIRE(obs && rew && returns && adv && workspace && path_off && n_paths, "promp_process_samples_ragged: null pointer argument");
    PROMP_REQUIRE(baseline_kind == PROMP_BASELINE_ZERO || baseline_kind == PROMP_BASELINE_LINEAR_FEATURE,
                  "promp_process_samples_ragged: unknown baseline kind %d", baseline_kind);
    PROMP_REQUIRE(discount >= 0.0 && discount <= 1.0 && gae_lambda >= 0.0 && gae_lambda <= 1.0,
                  "promp_process_samples_ragged: discount and gae_lambda must be in [0,1]");
    const int64_t need = promp_process_workspace_bytes_ragged(M, max_paths, max_samples, obs_dim);
    if (workspace_bytes < need) {
        set_error("promp_process_samples_ragged: workspace too small (%lld < %lld bytes)", (long long)workspace_bytes, (long long)need);
        return PROMP_ERR_WORKSPACE;
    }
    int C, EPC;
    proc_chunks(M, max_paths, &C, &EPC);
    double* ws = (double*)workspace;
    double* gram_p = ws + (int64_t)M * 2 * max_samples;
    double* stat_p = gram_p + (int64_t)M * C * PS_MAXPAIR;
    int32_t* tpos = (int32_t*)(stat_p + (int64_t)M * C * 8);
    ProcArgs A{M, max_paths, 0, obs_dim, obs, rew, discount, gae_lambda, reg_coeff, baseline_kind, normalize_adv, positive_adv,
               returns, adv, coeffs, stats, ws, gram_p, stat_p, C, EPC, path_off, n_paths, max_samples, tpos};
    process_gram_kernel<<<dim3(C, M), PS_THREADS, 0, (cudaStream_t)stream>>>(A);
    PROMP_LAUNCH_CHECK("process_gram_kernel");
    process_finish_kernel<<<M, PS_THREADS, 0, (cudaStream_t)stream>>>(A);
    PROMP_LAUNCH_CHECK("process_finish_kernel");
    return PROMP_OK;
}
