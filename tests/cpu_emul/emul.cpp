// TEST-ONLY: drives the __host__ __device__ index arithmetic of goofer_b200/csrc (FFT passes, index
// maps) with serial loops over "threads" so that it can be checked against the oracle on a machine
// without a GPU.  Never linked into libgoofer_b200.so, never on the product path.
#include <vector>
#include <cstring>
#include <cmath>
#include "../../goofer_b200/csrc/gf_fft.cuh"
#include "../../goofer_b200/csrc/gf_maps.cuh"
#include "../../goofer_b200/csrc/gf_conv.cuh"

static void tables(std::vector<float2> &tw512, std::vector<float2> &tw1024)
{
    const double PI = 3.141592653589793238462643383279502884;
    tw512.resize(GF_TWL_N); tw1024.resize(513);
    gf_twl_fill(tw512.data());
    for (int k = 0; k <= 512; ++k) tw1024[k] = make_float2((float)std::cos(-2.0 * PI * k / 1024.0), (float)std::sin(-2.0 * PI * k / 1024.0));
}

template <bool INV> static void fft512(float2 *buf, const float2 *tw512)
{
    float2 v[64][8];
    for (int j = 0; j < 64; ++j) gf_fft_pass_load<INV, 1>(j, buf, tw512, v[j]);
    for (int j = 0; j < 64; ++j) gf_fft_pass_store<1>(j, buf, v[j]);
    for (int j = 0; j < 64; ++j) gf_fft_pass_load<INV, 8>(j, buf, tw512, v[j]);
    for (int j = 0; j < 64; ++j) gf_fft_pass_store<8>(j, buf, v[j]);
    for (int j = 0; j < 64; ++j) gf_fft_pass_load<INV, 64>(j, buf, tw512, v[j]);
    for (int j = 0; j < 64; ++j) gf_fft_pass_store<64>(j, buf, v[j]);
}

extern "C" void emul_rfft1024(const float *x, float *X /* 513 complex interleaved */)
{
    std::vector<float2> tw512, tw1024;
    tables(tw512, tw1024);
    std::vector<float2> buf(GF_FFT_BUF);
    for (int m = 0; m < 512; ++m) buf[gf_fpad(m)] = make_float2(x[2 * m], x[2 * m + 1]);
    fft512<false>(buf.data(), tw512.data());
    for (int k = 0; k <= 256; ++k) {
        float2 Xk, Xm;
        gf_rfft_split(buf[gf_fpad(k)], buf[gf_fpad((512 - k) & 511)], tw1024[k], Xk, Xm);
        X[2 * k] = Xk.x; X[2 * k + 1] = Xk.y;
        X[2 * (512 - k)] = Xm.x; X[2 * (512 - k) + 1] = Xm.y;
    }
}

extern "C" void emul_irfft1024(const float *X, float *x)
{
    std::vector<float2> tw512, tw1024;
    tables(tw512, tw1024);
    std::vector<float2> buf(GF_FFT_BUF);
    for (int k = 0; k <= 256; ++k) {
        float2 Xk = make_float2(X[2 * k], X[2 * k + 1]), Xm = make_float2(X[2 * (512 - k)], X[2 * (512 - k) + 1]);
        if (k == 0) { Xk.y = 0.f; Xm.y = 0.f; }
        float2 Zk, Zm;
        gf_irfft_merge(Xk, Xm, tw1024[k], Zk, Zm);
        buf[gf_fpad(k)] = Zk;
        if (k != 0 && k != 256) buf[gf_fpad(512 - k)] = Zm;
    }
    fft512<true>(buf.data(), tw512.data());
    for (int m = 0; m < 512; ++m) { x[2 * m] = buf[gf_fpad(m)].x * (1.0f / 512.0f); x[2 * m + 1] = buf[gf_fpad(m)].y * (1.0f / 512.0f); }
}

extern "C" int emul_plan_size() { return (int)sizeof(GfNotePlan); }

// env_new[:, t] as a mix of stored source frames: returns n, fills f[4], w[4]
extern "C" int emul_env_mix(const void *plan, int t, int *f, double *w)
{
    GfNotePlan p;
    std::memcpy(&p, plan, sizeof(p));
    GfMix m;
    gf_env_mix(p, t, m);
    for (int k = 0; k < m.n; ++k) { f[k] = gf_src_frame(p, m.f[k]); w[k] = m.w[k]; }
    return m.n;
}

extern "C" void emul_mask_new(const void *plan, const float *mask_src, double *out)
{
    GfNotePlan p;
    std::memcpy(&p, plan, sizeof(p));
    for (int i = 0; i < p.n_total; ++i) out[i] = gf_mask_new(p, mask_src, i);
}

extern "C" void emul_track_canon(const void *plan, const double *trk, int k, float *out)
{
    GfNotePlan p;
    std::memcpy(&p, plan, sizeof(p));
    const GfTrackSlices s = gf_track_slices(p, k);
    for (int t = 0; t < p.T_env; ++t) out[t] = gf_track_canon(p, s, trk, k, t);
}

// ---- overlap-save Gaussian smoothing (k_conv.cu): the same passes, block layout and spectrum product, run serially ----
// v[j * 8 + r]: the registers of "thread" j
template <typename T, int N> static void conv_fft(GfC<T> *buf, const GfC<T> *tw, GfC<T> *v)
{
    const int TH = N / 8;
    constexpr int NS0 = GfConvShape<N>::NS0;
    if (GfConvShape<N>::R2) {
        for (int j = 0; j < TH; ++j) gf_conv_r2_store<T, N>(j, buf, &v[8 * j]);
        for (int j = 0; j < TH; ++j) gf_conv_pass_load<T, N, NS0>(j, buf, tw, &v[8 * j]);
    } else
        for (int j = 0; j < TH; ++j) gf_cdft8(&v[8 * j]);
    for (int j = 0; j < TH; ++j) gf_conv_pass_store<T, N, NS0>(j, buf, &v[8 * j]);
    for (int j = 0; j < TH; ++j) gf_conv_pass_load<T, N, NS0 * 8>(j, buf, tw, &v[8 * j]);
    for (int j = 0; j < TH; ++j) gf_conv_pass_store<T, N, NS0 * 8>(j, buf, &v[8 * j]);
    for (int j = 0; j < TH; ++j) gf_conv_pass_load<T, N, NS0 * 64>(j, buf, tw, &v[8 * j]);
    for (int j = 0; j < TH; ++j) gf_conv_pass_store<T, N, NS0 * 64>(j, buf, &v[8 * j]);
    for (int j = 0; j < TH; ++j) gf_conv_pass_load<T, N, NS0 * 512>(j, buf, tw, &v[8 * j]);
}

static int reflect_idx(int q, int n)
{
    if (n <= 1) return 0;
    const int per = 2 * (n - 1);
    q %= per;
    if (q < 0) q += per;
    return q < n ? q : per - q;
}

template <typename T, int N> static int conv_run(const T *x, int n, double sigma, T *out)
{
    if (!gf_fir_wants_fft(sigma, sizeof(T) == 4)) return 0;
    const int TH = N / 8;
    std::vector<GfC<T>> tw(N), buf(GfConvBuf<T, N>::LEN), v((size_t)N);
    gf_conv_tw_fill<T, N>(tw.data());
    const int radius = gf_fir_radius(sigma), V = N - 2 * radius;
    double norm = 0.0;
    for (int j = 0; j <= 2 * radius; ++j) { const double t = (double)(j - radius) / sigma; norm += std::exp(-0.5 * t * t); }
    for (int j = 0; j < TH; ++j)
        for (int r = 0; r < 8; ++r) {
            const int i = j + r * TH;
            const int d = i < N - i ? i : N - i;
            const double t = (double)d / sigma;
            v[8 * j + r] = gf_c<T>(d <= radius ? (T)(std::exp(-0.5 * t * t) / norm) : (T)0, (T)0);
        }
    conv_fft<T, N>(buf.data(), tw.data(), v.data());
    std::vector<T> h(N);
    for (int q = 0; q < N; ++q) h[q] = v[q].x * (T)(1.0 / N);
    auto fetch = [&](int pos) -> T {
        if (pos >= n + radius || pos < -radius) return (T)0;
        return x[(pos >= 0 && pos < n) ? pos : reflect_idx(pos, n)];
    };
    const int n_pairs = ((n + V - 1) / V + 1) / 2;
    for (int p = 0; p < n_pairs; ++p) {
        const int o1 = 2 * p * V;
        for (int j = 0; j < TH; ++j)
            for (int r = 0; r < 8; ++r) {
                const int i = j + r * TH;
                v[8 * j + r] = gf_c<T>(fetch(o1 - radius + i), fetch(o1 + V - radius + i));
            }
        conv_fft<T, N>(buf.data(), tw.data(), v.data());
        for (int q = 0; q < N; ++q) v[q] = gf_c<T>(v[q].x * h[q], -v[q].y * h[q]);
        conv_fft<T, N>(buf.data(), tw.data(), v.data());
        for (int j = 0; j < TH; ++j)
            for (int r = 0; r < 8; ++r) {
                const int i = j + r * TH - radius;
                if (i < 0 || i >= V) continue;
                if (o1 + i < n) out[o1 + i] = v[8 * j + r].x;
                if (o1 + V + i < n) out[o1 + V + i] = -v[8 * j + r].y;
            }
    }
    return 1;
}

extern "C" int emul_fftconv_f32(const float *x, int n, double sigma, float *out) { return conv_run<float, GF_CONV_N32>(x, n, sigma, out); }
extern "C" int emul_fftconv_f64(const double *x, int n, double sigma, double *out) { return conv_run<double, GF_CONV_N64>(x, n, sigma, out); }
