// k_conv.cu -- gaussian_filter1d along time as an overlap-save convolution for LONG kernels
//   gf_fftconv_kernel<float, 8192>   f32 jobs: sigma 441 on the pitch-deviation curve and on the voicing mask
//                                    (3,529 taps, /root/reference/SillySampler.py:865-866, 880)
//   gf_fftconv_kernel<double, 4096>  fp64 jobs: the jitter curves, sigma 73.5 / 49 (589 / 393 taps,
//                                    /root/reference/GOOFER.py:654, 667); their result scales f0, so the transform,
//                                    the spectrum product and the twiddles stay fp64
// Short kernels (sigma 20 / 25) stay on the direct sliding window of k_prep.cu.  Same definition as there:
// numpy 'reflect' padding (GOOFER.py:249-250), taps exp(-t^2 / 2 sigma^2) / sum, radius int(4 sigma + 0.5).
#include "gf_device.cuh"
#include "gf_conv.cuh"
#include "gf_kernels.h"

// in: v[r] = x[j + r N / 8];  out: v[r] = X[j + r N / 8] (forward transform, natural order).  buf is scratch.
template <typename T, int N>
__device__ __forceinline__ void gf_conv_fft(GfC<T> *buf, const GfC<T> *__restrict__ tw, int j, GfC<T> *v)
{
    constexpr int NS0 = GfConvShape<N>::NS0;
    __syncthreads();                                   // the last pass of the previous transform has read buf
    if (GfConvShape<N>::R2) {
        gf_conv_r2_store<T, N>(j, buf, v);             __syncthreads();
        gf_conv_pass_load<T, N, NS0>(j, buf, tw, v);   __syncthreads();
    } else gf_cdft8(v);                                // pass NS = 1: no twiddles, inputs already in registers
    gf_conv_pass_store<T, N, NS0>(j, buf, v);          __syncthreads();
    gf_conv_pass_load<T, N, NS0 * 8>(j, buf, tw, v);   __syncthreads();
    gf_conv_pass_store<T, N, NS0 * 8>(j, buf, v);      __syncthreads();
    gf_conv_pass_load<T, N, NS0 * 64>(j, buf, tw, v);  __syncthreads();
    gf_conv_pass_store<T, N, NS0 * 64>(j, buf, v);     __syncthreads();
    gf_conv_pass_load<T, N, NS0 * 512>(j, buf, tw, v);
}

template <typename T> __device__ __forceinline__ const GfC<T> *gf_conv_tw();
template <> __device__ __forceinline__ const GfC<float> *gf_conv_tw<float>() { return d_conv.tw32; }
template <> __device__ __forceinline__ const GfC<double> *gf_conv_tw<double>() { return d_conv.tw64; }

template <typename T>
__device__ __forceinline__ T gf_conv_fetch(const GfFirJob &jb, int pos, int n, int radius)
{
    if (pos >= n + radius || pos < -radius) return (T)0;          // never contributes to an output below n
    // numpy 'reflect': one fold suffices when the signal is longer than the halo (the general form costs an integer
    // modulo per sample: 9 % of the fp64 kernel's instructions in the first version)
    int q = pos;
    if (pos < 0 || pos >= n) q = (n > radius) ? (pos < 0 ? -pos : 2 * (n - 1) - pos) : gf_reflect(pos, n);
    if (sizeof(T) == 4) return (T)((const float *)jb.in)[(size_t)q * jb.in_stride];
    double v = jb.in_f64 ? ((const double *)jb.in)[(size_t)q * jb.in_stride] : (double)((const float *)jb.in)[(size_t)q * jb.in_stride];
    if (jb.in_cast_f32) v = (double)(float)v;
    return (T)v;
}

#ifndef GF_CONV64_MINB
#define GF_CONV64_MINB 2        // resident CTAs per SM the fp64 kernel is compiled for: 2 caps it at 64 registers (412 bytes of
                                // spills) and hides the barrier / twiddle latency better than 126 registers and one CTA: 0.64 -> 0.57 ms (c3)
#endif
template <typename T, int N>
__global__ void __launch_bounds__(N / 8, sizeof(T) == 8 ? GF_CONV64_MINB : 1) gf_fftconv_kernel(const GfFirJob *__restrict__ jobs, int pairs_per_cta)
{
    constexpr int THREADS = N / 8;
    extern __shared__ __align__(16) unsigned char conv_raw[];
    __shared__ double red[THREADS / 32];
    GfC<T> *buf = reinterpret_cast<GfC<T> *>(conv_raw);                // GfConvBuf<T, N>::LEN
    const GfFirJob jb = jobs[blockIdx.y];
    const bool f32 = gf_fir_f32(jb.in_f64, jb.maxabs, jb.in_cast_f32);
    if (f32 != (sizeof(T) == 4) || !gf_fir_wants_fft(jb.sigma, f32)) return;
    const int n = jb.n;
    const int radius = gf_fir_radius(jb.sigma);
    const int V = N - 2 * radius;                                     // valid outputs per block
    const int n_pairs = ((n + V - 1) / V + 1) / 2;
    const int p0 = blockIdx.x * pairs_per_cta;
    if (p0 >= n_pairs) return;
    const int tid = threadIdx.x;
    const GfC<T> *tw = gf_conv_tw<T>();

    // taps: exp(-t^2 / 2 sigma^2) / sum over the 2 radius + 1 taps (fp64, like the direct kernels)
    double part = 0.0;
    for (int j = tid; j <= 2 * radius; j += THREADS) { const double t = (double)(j - radius) / jb.sigma; part += exp(-0.5 * t * t); }
    part = gf_warp_sum(part);
    if ((tid & 31) == 0) red[tid >> 5] = part;
    __syncthreads();
    double norm = 0.0;
    for (int w = 0; w < THREADS / 32; ++w) norm += red[w];
    // spectrum of the kernel centred on index 0 of the circular block (h[m] = tap[radius + m], m = -radius..radius):
    // real and even; thread tid keeps H[tid + r N / 8] / N in registers for every block it filters
    GfC<T> v[8];
    T h[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int i = tid + r * THREADS;
        const int d = min(i, N - i);
        v[r] = gf_c<T>(d <= radius ? (T)gf_gauss_tap(radius + d, radius, jb.sigma, norm) : (T)0, (T)0);
    }
    gf_conv_fft<T, N>(buf, tw, tid, v);
#pragma unroll
    for (int r = 0; r < 8; ++r) h[r] = v[r].x * (T)(1.0 / N);

    double mx = 0.0;
    const int p1 = min(p0 + pairs_per_cta, n_pairs);
    for (int p = p0; p < p1; ++p) {
        const int o1 = 2 * p * V;                                     // first output of block 1; block 2 starts V later
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int i = tid + r * THREADS;
            v[r] = gf_c<T>(gf_conv_fetch<T>(jb, o1 - radius + i, n, radius), gf_conv_fetch<T>(jb, o1 + V - radius + i, n, radius));
        }
        gf_conv_fft<T, N>(buf, tw, tid, v);
        // Y = Z H / N; the inverse transform is the forward one on conj(Y), conjugated again
#pragma unroll
        for (int r = 0; r < 8; ++r) v[r] = gf_c<T>(v[r].x * h[r], -v[r].y * h[r]);
        gf_conv_fft<T, N>(buf, tw, tid, v);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int i = tid + r * THREADS - radius;                 // output index inside the block
            if (i < 0 || i >= V) continue;
            const int t1 = o1 + i, t2 = o1 + V + i;
            if (t1 < n) {
                const T a = v[r].x;
                if (jb.out_f64) ((double *)jb.out)[t1] = (double)a; else ((float *)jb.out)[t1] = (float)a;
                mx = fmax(mx, fabs((double)a) + 1e-6);
            }
            if (t2 < n) {
                const T a = -v[r].y;
                if (jb.out_f64) ((double *)jb.out)[t2] = (double)a; else ((float *)jb.out)[t2] = (float)a;
                mx = fmax(mx, fabs((double)a) + 1e-6);
            }
        }
    }
    if (jb.maxabs) {
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((tid & 31) == 0) atomicMax((unsigned long long *)jb.maxabs, (unsigned long long)__double_as_longlong(mx));
    }
}

template <typename T, int N> static constexpr size_t gf_fftconv_smem() { return sizeof(GfC<T>) * GfConvBuf<T, N>::LEN; }

// launches the overlap-save kernels for the jobs that want them; returns the number of launches
int gf_launch_fftconv(const GfFirJob *h_jobs, const GfFirJob *d_jobs, int n_jobs, cudaStream_t st)
{
    int pairs32 = 0, pairs64 = 0, jobs32 = 0, jobs64 = 0;
    for (int k = 0; k < n_jobs; ++k) {
        const GfFirJob &j = h_jobs[k];
        const bool f32 = gf_fir_f32(j.in_f64, j.maxabs, j.in_cast_f32);
        if (j.n <= 0 || !gf_fir_wants_fft(j.sigma, f32)) continue;
        const int V = (f32 ? GF_CONV_N32 : GF_CONV_N64) - 2 * gf_fir_radius(j.sigma);
        const int pairs = ((j.n + V - 1) / V + 1) / 2;
        if (f32) { pairs32 = std::max(pairs32, pairs); ++jobs32; } else { pairs64 = std::max(pairs64, pairs); ++jobs64; }
    }
    // every CTA rebuilds the spectrum of its job's taps (one transform): give it several block pairs to amortise
    // that, as long as the grid still covers the SMs about twice
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    auto ppc_of = [sms](int pairs, int jobs) { return std::max(1, std::min(8, (int)((long long)pairs * jobs / (2 * sms)))); };
    int launches = 0;
    if (jobs32) {
        static GfSmemLimit memo;
        gf_smem_limit(gf_fftconv_kernel<float, GF_CONV_N32>, gf_fftconv_smem<float, GF_CONV_N32>(), memo);
        const int ppc = ppc_of(pairs32, jobs32);
        gf_fftconv_kernel<float, GF_CONV_N32><<<dim3((pairs32 + ppc - 1) / ppc, n_jobs), GF_CONV_N32 / 8, gf_fftconv_smem<float, GF_CONV_N32>(), st>>>(d_jobs, ppc);
        ++launches;
    }
    if (jobs64) {
        static GfSmemLimit memo;
        gf_smem_limit(gf_fftconv_kernel<double, GF_CONV_N64>, gf_fftconv_smem<double, GF_CONV_N64>(), memo);
        const int ppc = ppc_of(pairs64, jobs64);
        gf_fftconv_kernel<double, GF_CONV_N64><<<dim3((pairs64 + ppc - 1) / ppc, n_jobs), GF_CONV_N64 / 8, gf_fftconv_smem<double, GF_CONV_N64>(), st>>>(d_jobs, ppc);
        ++launches;
    }
    return launches;
}
