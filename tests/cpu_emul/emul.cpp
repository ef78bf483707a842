// TEST-ONLY: drives the __host__ __device__ index arithmetic of goofer_b200/csrc (FFT passes, index
// maps) with serial loops over "threads" so that it can be checked against the oracle on a machine
// without a GPU.  Never linked into libgoofer_b200.so, never on the product path.
#include <vector>
#include <cstring>
#include <cmath>
#include "../../goofer_b200/csrc/gf_fft.cuh"
#include "../../goofer_b200/csrc/gf_maps.cuh"

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
