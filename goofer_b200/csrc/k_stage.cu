// k_stage.cu -- stage-level kernels behind goofer_stft_batch / goofer_istft_batch: the same
// shared-memory FFT, framing and overlap-add blocks as the fused frame kernel, one stage at a time.
//   gf.stft  GOOFER.py:355-370      gf.istft  GOOFER.py:392-413 (+ _overlap_add :372-390)
#include "gf_frame.cuh"

#define GF_STAGE_R 4
#define GF_STAGE_THREADS (64 * GF_STAGE_R)

struct GfStageSmem {
    GfFrameTables tab;
    float2 z[GF_STAGE_R][GF_FFT_BUF];
    float ring[GF_RING];
};

// S layout: (n_sig, 513, T) complex64
__global__ void __launch_bounds__(GF_STAGE_THREADS)
gf_stft_kernel(const float *__restrict__ x, int n, int T, float2 *__restrict__ S)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GfStageSmem &sm = *reinterpret_cast<GfStageSmem *>(smem_raw);
    const float *xs = x + (size_t)blockIdx.y * n;
    float2 *Ss = S + (size_t)blockIdx.y * GF_NBINS * T;
    const int t0 = blockIdx.x * GF_STAGE_R;
    if (t0 >= T) return;
    const int nf = min(GF_STAGE_R, T - t0);
    gf_stage_tables(&sm.tab);
    __syncthreads();
    gf_load_frames(&sm.z[0][0], t0, nf, n, sm.tab.win, [&](int i) { return xs[i]; });
    __syncthreads();
    gf_cta_fft512<false>(&sm.z[0][0], nf, sm.tab.twl);
    for (int idx = threadIdx.x; idx < nf * 257; idx += blockDim.x) {
        const int k = idx / nf, f = idx - k * nf;
        const float2 Zk = sm.z[f][gf_fpad(k)], Zm = sm.z[f][gf_fpad((512 - k) & 511)];
        float2 Xk, Xm;
        gf_rfft_split(Zk, Zm, sm.tab.tw1024[k], Xk, Xm);
        Ss[(size_t)k * T + t0 + f] = Xk;
        Ss[(size_t)(512 - k) * T + t0 + f] = Xm;
    }
}

__global__ void __launch_bounds__(GF_STAGE_THREADS)
gf_istft_kernel(const float2 *__restrict__ S, int T, int length, float *__restrict__ y, int blocks_per_cta)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GfStageSmem &sm = *reinterpret_cast<GfStageSmem *>(smem_raw);
    const float2 *Ss = S + (size_t)blockIdx.y * GF_NBINS * T;
    float *ys = y + (size_t)blockIdx.y * length;
    const int b0 = 2 + blockIdx.x * blocks_per_cta;
    const int nb = blocks_per_cta;
    if (b0 > max(T, 2)) return;
    gf_stage_tables(&sm.tab);
    for (int i = threadIdx.x; i < GF_RING; i += blockDim.x) sm.ring[i] = 0.0f;
    __syncthreads();
    const int t_begin = max(0, b0 - 3), t_end = min(T - 1, b0 + nb - 1);
    for (int t0 = t_begin; t0 <= t_end; t0 += GF_STAGE_R) {
        const int nf = min(GF_STAGE_R, t_end - t0 + 1);
        for (int idx = threadIdx.x; idx < nf * 257; idx += blockDim.x) {
            const int k = idx / nf, f = idx - k * nf;
            float2 Xk = Ss[(size_t)k * T + t0 + f], Xm = Ss[(size_t)(512 - k) * T + t0 + f];
            if (k == 0) { Xk.y = 0.f; Xm.y = 0.f; }
            float2 Zk, Zm;
            gf_irfft_merge(Xk, Xm, sm.tab.tw1024[k], Zk, Zm);
            sm.z[f][gf_fpad(k)] = Zk;
            if (k != 0 && k != 256) sm.z[f][gf_fpad(512 - k)] = Zm;
        }
        __syncthreads();
        gf_cta_fft512<true>(&sm.z[0][0], nf, sm.tab.twl);
        gf_ola_add(sm.ring, &sm.z[0][0], t0, nf, sm.tab.win);
        __syncthreads();
        const int last_blk = (t0 + nf - 1 == T - 1) ? T : (t0 + nf - 1);
        for (int b = t0; b <= last_blk; ++b)
            gf_ola_emit(sm.ring, b, T, length, ys, b >= b0 && b < b0 + nb && b >= 2);
        __syncthreads();
    }
    if (b0 + nb > T)
        for (int i = GF_HOP * (T - 1) + threadIdx.x; i < length; i += blockDim.x) ys[i] = 0.0f;
}

static void gf_stage_smem_limits()
{
    static GfSmemLimit memo_f, memo_i;
    gf_smem_limit(gf_stft_kernel, sizeof(GfStageSmem), memo_f);
    gf_smem_limit(gf_istft_kernel, sizeof(GfStageSmem), memo_i);
}

void gf_launch_stft(const float *x, int n_sig, int n, float2 *S, cudaStream_t st)
{
    const int T = 1 + n / GF_HOP;
    gf_stage_smem_limits();
    dim3 grid((T + GF_STAGE_R - 1) / GF_STAGE_R, n_sig);
    gf_stft_kernel<<<grid, GF_STAGE_THREADS, sizeof(GfStageSmem), st>>>(x, n, T, S);
}

void gf_launch_istft(const float2 *S, int n_sig, int T, int length, float *y, cudaStream_t st)
{
    gf_stage_smem_limits();
    const int bpc = 32;
    const int n_blocks = max(T, 2) - 2 + 1;            // hop blocks 2 .. max(T, 2)
    dim3 grid((n_blocks + bpc - 1) / bpc, n_sig);
    gf_istft_kernel<<<grid, GF_STAGE_THREADS, sizeof(GfStageSmem), st>>>(S, T, length, y, bpc);
}
