// k_analyse.cu -- analysis front-end: the envelope half of gf.extract_features (the step BEFORE the render path,
// SURVEY.md section 8f row 1).  f0 and formants come from Praat in the reference (third party, out of scope);
// everything numeric about the spectral envelope is reproduced here:
//
//   stft(y)                          GOOFER.py:944 (355-370)        gf_stft_kernel (k_stage.cu)
//   mag = |S| + 1e-8                 GOOFER.py:945                  gf_an_env_kernel
//   env = gaussian(mag, sigma 2)     GOOFER.py:946 (241-261)              "
//   compress_env_to_knots            GOOFER.py:97-147: sigma 0.5 blur, log, mel knots K = 32, 48, .. 192 until the
//                                    2-tap log-lerp reconstruction is within 1e-2 on <= 256 probe frames
//                                                                   gf_an_search_kernel + gf_an_pack_kernel
// Output per signal: K, hz_knots (K,) f32, knot_vals_log (K, T) f16 -- the arrays gf.save_features stores.
#include <cuda_fp16.h>
#include "gf_frame.cuh"

#define GF_AN_KMAX 192
#define GF_AN_NK 11                 // K = 32 + 16 c, c = 0..10
#define GF_AN_FR 8                  // frames per CTA of the envelope kernel
#define GF_AN_LD 544                // 513 + halo 8 + slack

__device__ __forceinline__ int gf_refl513(int q) { return q < 0 ? -q : (q > 512 ? 1024 - q : q); }

// S (513, T) c64 -> log_env (513, T) f32 and env_s (513, T) f32 (the sigma-0.5 smoothed envelope the search compares with)
__global__ void __launch_bounds__(256)
gf_an_env_kernel(const float2 *__restrict__ S, int T, float *__restrict__ log_env, float *__restrict__ env_s)
{
    __shared__ float mag[GF_AN_FR][GF_AN_LD];
    __shared__ float e32[GF_AN_FR][GF_AN_LD];
    __shared__ double g2[17], g05[5];
    const float2 *Ss = S + (size_t)blockIdx.y * GF_NBINS * T;
    float *le = log_env + (size_t)blockIdx.y * GF_NBINS * T, *es = env_s + (size_t)blockIdx.y * GF_NBINS * T;
    const int t0 = blockIdx.x * GF_AN_FR;
    if (threadIdx.x < 17) {
        double norm = 0.0;
        for (int j = 0; j < 17; ++j) { const double t = (double)(j - 8) / 2.0; norm += exp(-0.5 * t * t); }
        const double t = (double)((int)threadIdx.x - 8) / 2.0;
        g2[threadIdx.x] = exp(-0.5 * t * t) / norm;
    } else if (threadIdx.x >= 32 && threadIdx.x < 37) {
        double norm = 0.0;
        for (int j = 0; j < 5; ++j) { const double t = (double)(j - 2) / 0.5; norm += exp(-0.5 * t * t); }
        const double t = (double)((int)threadIdx.x - 34) / 0.5;
        g05[threadIdx.x - 32] = exp(-0.5 * t * t) / norm;
    }
    for (int idx = threadIdx.x; idx < GF_NBINS * GF_AN_FR; idx += blockDim.x) {
        const int b = idx >> 3, f = idx & 7, t = t0 + f;
        float m = 0.0f;
        if (t < T) { const float2 z = Ss[(size_t)b * T + t]; m = hypotf(z.x, z.y) + 1e-8f; }      // GOOFER.py:945
        mag[f][b] = m;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < GF_NBINS * GF_AN_FR; idx += blockDim.x) {
        const int b = idx >> 3, f = idx & 7;
        double a = 0.0;
#pragma unroll
        for (int j = 0; j < 17; ++j) a += g2[j] * (double)mag[f][gf_refl513(b + j - 8)];           // GOOFER.py:946
        e32[f][b] = (float)a;                                                                     // to_compute (:98)
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < GF_NBINS * GF_AN_FR; idx += blockDim.x) {
        const int b = idx >> 3, f = idx & 7, t = t0 + f;
        if (t >= T) continue;
        double a = 0.0;
#pragma unroll
        for (int j = 0; j < 5; ++j) a += g05[j] * (double)e32[f][gf_refl513(b + j - 2)];           // :99-100
        le[(size_t)b * T + t] = (float)log(fmax(a, 1e-8));                                        // :101
        es[(size_t)b * T + t] = (float)a;
    }
}

// knot bins and the 2-tap lerp table of precompute_interp_matrix for one K (shared memory)
struct GfAnKnots { int bk[GF_AN_KMAX]; int idx[GF_NBINS]; float w0[GF_NBINS], w1[GF_NBINS]; };

__device__ __forceinline__ void gf_an_tables(GfAnKnots &kt, const float *__restrict__ hz, int K, int sr)
{
    const float res = (float)((double)sr / 1024.0);
    const double fstep = 1024.0 * (1.0 / (double)sr);
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        int b = (int)rintf(hz[k] / res);                                                          // :113 np.round: half to even
        kt.bk[k] = b < 0 ? 0 : (b > 512 ? 512 : b);
    }
    for (int b = threadIdx.x; b < GF_NBINS; b += blockDim.x) {
        const float f = (float)((double)b / fstep);
        int lo = 0, hi = K;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (hz[mid] <= f) lo = mid + 1; else hi = mid; }
        int i = lo - 1;
        i = i < 0 ? 0 : (i > K - 2 ? K - 2 : i);
        const float x0 = hz[i], x1 = hz[i + 1];
        const float w1 = (f - x0) / fmaxf(x1 - x0, 1e-12f);
        kt.idx[b] = i; kt.w1[b] = w1; kt.w0[b] = 1.0f - w1;
    }
}

// errs[sig][c] = max relative reconstruction error of K = 32 + 16 c on the probe frames   (GOOFER.py:116-119)
__global__ void __launch_bounds__(256)
gf_an_search_kernel(const float *__restrict__ log_env, const float *__restrict__ env_s, int T, int sr,
                    const float *__restrict__ hz_tab /* (11, 192) */, float *__restrict__ errs)
{
    __shared__ GfAnKnots kt;
    __shared__ float red[8];
    const int c = blockIdx.x, K = 32 + 16 * c;
    const float *le = log_env + (size_t)blockIdx.y * GF_NBINS * T, *es = env_s + (size_t)blockIdx.y * GF_NBINS * T;
    gf_an_tables(kt, hz_tab + (size_t)c * GF_AN_KMAX, K, sr);
    __syncthreads();
    const int np = min(256, T);
    float mx = 0.0f;
    for (int item = threadIdx.x; item < np * GF_NBINS; item += blockDim.x) {
        const int p = item % np, b = item / np;
        // check_idx = np.linspace(0, T - 1, np, dtype=int)
        const int t = (np <= 1) ? 0 : ((p == np - 1) ? T - 1 : (int)((double)p * ((double)(T - 1) / (double)(np - 1))));
        const int i = kt.idx[b];
        const float rec = kt.w0[b] * le[(size_t)kt.bk[i] * T + t] + kt.w1[b] * le[(size_t)kt.bk[i + 1] * T + t];
        const float ref = es[(size_t)b * T + t];
        mx = fmaxf(mx, fabsf(expf(rec) - ref) / (ref + 1e-8f));
    }
    mx = gf_warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = 0.0f;
        for (int w = 0; w < 8; ++w) m = fmaxf(m, red[w]);
        errs[(size_t)blockIdx.y * GF_AN_NK + c] = m;
    }
}

// smallest K whose error is below eps (else K_max); write hz_knots and the f16 log-knots
__global__ void __launch_bounds__(256)
gf_an_pack_kernel(const float *__restrict__ log_env, int T, int sr, const float *__restrict__ hz_tab, const float *__restrict__ errs,
                  float eps, uint16_t *__restrict__ knots_out, float *__restrict__ hz_out, int *__restrict__ K_out)
{
    __shared__ int s_bk[GF_AN_KMAX];
    const int sig = blockIdx.x;
    int c = GF_AN_NK - 1;
    for (int q = 0; q < GF_AN_NK; ++q)
        if (errs[(size_t)sig * GF_AN_NK + q] < eps) { c = q; break; }
    const int K = 32 + 16 * c;
    const float *hz = hz_tab + (size_t)c * GF_AN_KMAX;
    const float res = (float)((double)sr / 1024.0);
    for (int k = threadIdx.x; k < GF_AN_KMAX; k += blockDim.x) {
        int b = 0;
        if (k < K) { b = (int)rintf(hz[k] / res); b = b < 0 ? 0 : (b > 512 ? 512 : b); }
        s_bk[k] = b;
        hz_out[(size_t)sig * GF_AN_KMAX + k] = k < K ? hz[k] : 0.0f;
    }
    if (threadIdx.x == 0) K_out[sig] = K;
    __syncthreads();
    const float *le = log_env + (size_t)sig * GF_NBINS * T;
    uint16_t *ko = knots_out + (size_t)sig * GF_AN_KMAX * T;
    for (int item = threadIdx.x; item < K * T; item += blockDim.x) {
        const int k = item / T, t = item - k * T;
        ko[(size_t)k * T + t] = __half_as_ushort(__float2half_rn(le[(size_t)s_bk[k] * T + t]));   // astype(float16)
    }
}

// mel-uniform knot frequencies for every candidate K, with numpy's float32 arithmetic (GOOFER.py:74-82)
static void gf_an_hz_tables(int sr, float *tab /* (11, 192) */)
{
    const double mel_max = 2595.0 * std::log10(1.0 + ((double)sr / 2.0) / 700.0);
    for (int c = 0; c < GF_AN_NK; ++c) {
        const int K = 32 + 16 * c;
        const double step = mel_max / (double)(K - 1);
        for (int k = 0; k < GF_AN_KMAX; ++k) {
            float hz = 0.0f;
            if (k < K) {
                const float mel = (k == K - 1) ? (float)mel_max : (float)((double)k * step);
                const float e = mel / 2595.0f;
                hz = 700.0f * (powf(10.0f, e) - 1.0f);
            }
            tab[c * GF_AN_KMAX + k] = hz;
        }
    }
}

extern "C" size_t goofer_analyse_work_bytes(int32_t n_sig, int32_t n)
{
    if (n_sig < 0 || n < 2) return 0;
    const size_t T = 1 + (size_t)n / GF_HOP;
    const size_t per = GF_NBINS * T * (sizeof(float2) + 2 * sizeof(float)) + GF_AN_NK * sizeof(float) + 1024;
    return (size_t)n_sig * per + GF_AN_NK * GF_AN_KMAX * sizeof(float) + 4096;
}

extern "C" int goofer_analyse_batch(const float *y, int32_t n_sig, int32_t n, int32_t sr, uint16_t *knots_out, float *hz_out,
                                    int32_t *K_out, void *work, void *stream)
{
    if (!y || !knots_out || !hz_out || !K_out || !work || n_sig < 0 || n < 2 || sr <= 0) {
        gf_set_error("goofer_analyse_batch: invalid arguments");
        return GOOFER_ERR_INVALID;
    }
    if (n_sig == 0) return GOOFER_OK;
    int rc = gf_tables_init(sr);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int T = 1 + n / GF_HOP;
    Bump bp{(char *)work, goofer_analyse_work_bytes(n_sig, n), 0};
    float2 *S = bp.arr<float2>((size_t)n_sig * GF_NBINS * T);
    float *le = bp.arr<float>((size_t)n_sig * GF_NBINS * T);
    float *es = bp.arr<float>((size_t)n_sig * GF_NBINS * T);
    float *errs = bp.arr<float>((size_t)n_sig * GF_AN_NK);
    float *d_hz = bp.arr<float>(GF_AN_NK * GF_AN_KMAX);
    float tab[GF_AN_NK * GF_AN_KMAX];
    gf_an_hz_tables(sr, tab);
    void *stage = gf_pin_take(sizeof(tab));
    if (!stage) { gf_set_error("cudaMallocHost failed for the analysis tables"); return GOOFER_ERR_CUDA; }
    std::memcpy(stage, tab, sizeof(tab));
    if ((rc = gf_meta_copy(d_hz, stage, sizeof(tab), st)) != GOOFER_OK) return rc;
    gf_launch_stft(y, n_sig, n, S, st);
    gf_an_env_kernel<<<dim3((T + GF_AN_FR - 1) / GF_AN_FR, n_sig), 256, 0, st>>>(S, T, le, es);
    gf_an_search_kernel<<<dim3(GF_AN_NK, n_sig), 256, 0, st>>>(le, es, T, sr, d_hz, errs);
    gf_an_pack_kernel<<<n_sig, 256, 0, st>>>(le, T, sr, d_hz, errs, 1e-2f, knots_out, hz_out, K_out);
    GF_CUDA(cudaGetLastError());
    return GOOFER_OK;
}
