// gf_frame.cuh -- CTA-level building blocks of the frame pipeline: 64 threads per 512-point
// complex FFT through shared memory, sqrt-Hann framing with numpy 'reflect' padding, and the
// windowed overlap-add with the win^2 normalisation of GOOFER.py:372-413.
#pragma once
#include "gf_device.cuh"
#include "gf_fft.cuh"

// shared-memory tables every frame kernel stages once per CTA
struct __align__(16) GfFrameTables {
    float2 twl[GF_TWL_N];                     // per-thread FFT twiddles (gf_fft.cuh)
    float2 tw1024[513];
    float win[1024];
};

__device__ __forceinline__ void gf_stage_tables(GfFrameTables *st)
{
    for (int i = threadIdx.x; i < GF_TWL_N; i += blockDim.x) st->twl[i] = d_tab.twl[i];
    for (int i = threadIdx.x; i < 513; i += blockDim.x) st->tw1024[i] = d_tab.tw1024[i];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) st->win[i] = d_tab.win[i];
}

// barrier among the 64 threads (two warps) that share one transform: named barrier 1 + lane.  The Stockham passes of
// a transform only exchange data inside its own 64-thread lane, so the CTA-wide barrier is only needed where the
// thread-to-data mapping changes (before the first pass -- the caller's barrier -- and after the last store).
__device__ __forceinline__ void gf_lane_sync(int lane)
{
    asm volatile("bar.sync %0, 64;" ::"r"(lane + 1) : "memory");
}

// All threads of the CTA call this together, after a CTA-wide barrier that made the input visible.  `n_xf` transforms
// live at bufs + q * GF_FFT_BUF (q < n_xf); lane = threadIdx.x / 64 works on transforms lane, lane + n_lanes, ...
template <bool INV>
__device__ __forceinline__ void gf_cta_fft512(float2 *bufs, int n_xf, const float2 *twl)
{
    const int lane = threadIdx.x >> 6, j = threadIdx.x & 63, n_lanes = blockDim.x >> 6;
    for (int q0 = 0; q0 < n_xf; q0 += n_lanes) {
        const int q = q0 + lane;
        const bool on = q < n_xf;                            // uniform over the 64-thread lane
        float2 *buf = bufs + (size_t)q * GF_FFT_BUF;
        float2 v[8];
        if (on) {
            gf_fft_pass_load<INV, 1>(j, buf, twl, v);
            gf_lane_sync(lane);
            gf_fft_pass_store<1>(j, buf, v);
            gf_lane_sync(lane);
            gf_fft_pass_load<INV, 8>(j, buf, twl, v);
            gf_lane_sync(lane);
            gf_fft_pass_store<8>(j, buf, v);
            gf_lane_sync(lane);
            gf_fft_pass_load<INV, 64>(j, buf, twl, v);
            gf_lane_sync(lane);
            gf_fft_pass_store<64>(j, buf, v);
        }
    }
    __syncthreads();
}

// NQ transforms per 64-thread lane (n_xf = NQ * lanes), all advanced through a pass before the lane synchronises:
// five two-warp barriers and one CTA-wide barrier for the whole batch
template <bool INV, int NQ>
__device__ __forceinline__ void gf_cta_fft512_multi(float2 *bufs, const float2 *twl)
{
    const int lane = threadIdx.x >> 6, j = threadIdx.x & 63, n_lanes = blockDim.x >> 6;
    float2 v[NQ][8];
#pragma unroll
    for (int q = 0; q < NQ; ++q) gf_fft_pass_load<INV, 1>(j, bufs + (size_t)(lane + q * n_lanes) * GF_FFT_BUF, twl, v[q]);
    gf_lane_sync(lane);
#pragma unroll
    for (int q = 0; q < NQ; ++q) gf_fft_pass_store<1>(j, bufs + (size_t)(lane + q * n_lanes) * GF_FFT_BUF, v[q]);
    gf_lane_sync(lane);
#pragma unroll
    for (int q = 0; q < NQ; ++q) gf_fft_pass_load<INV, 8>(j, bufs + (size_t)(lane + q * n_lanes) * GF_FFT_BUF, twl, v[q]);
    gf_lane_sync(lane);
#pragma unroll
    for (int q = 0; q < NQ; ++q) gf_fft_pass_store<8>(j, bufs + (size_t)(lane + q * n_lanes) * GF_FFT_BUF, v[q]);
    gf_lane_sync(lane);
#pragma unroll
    for (int q = 0; q < NQ; ++q) gf_fft_pass_load<INV, 64>(j, bufs + (size_t)(lane + q * n_lanes) * GF_FFT_BUF, twl, v[q]);
    gf_lane_sync(lane);
#pragma unroll
    for (int q = 0; q < NQ; ++q) gf_fft_pass_store<64>(j, bufs + (size_t)(lane + q * n_lanes) * GF_FFT_BUF, v[q]);
    __syncthreads();
}

// NQ inverse transforms owned by ONE 64-thread group (buffers `stride` float2 apart), advanced through each pass
// together so that the group meets five times for the whole batch; the caller provides the barrier that publishes the
// result to the rest of the CTA.
template <int NQ>
__device__ __forceinline__ void gf_group_ifft512(float2 *buf0, int stride, const float2 *twl, int group, int j)
{
    float2 v[NQ][8];
#pragma unroll
    for (int q = 0; q < NQ; ++q) gf_fft_pass_load<true, 1>(j, buf0 + (size_t)q * stride, twl, v[q]);
    gf_lane_sync(group);
#pragma unroll
    for (int q = 0; q < NQ; ++q) gf_fft_pass_store<1>(j, buf0 + (size_t)q * stride, v[q]);
    gf_lane_sync(group);
#pragma unroll
    for (int q = 0; q < NQ; ++q) gf_fft_pass_load<true, 8>(j, buf0 + (size_t)q * stride, twl, v[q]);
    gf_lane_sync(group);
#pragma unroll
    for (int q = 0; q < NQ; ++q) gf_fft_pass_store<8>(j, buf0 + (size_t)q * stride, v[q]);
    gf_lane_sync(group);
#pragma unroll
    for (int q = 0; q < NQ; ++q) gf_fft_pass_load<true, 64>(j, buf0 + (size_t)q * stride, twl, v[q]);
    gf_lane_sync(group);
#pragma unroll
    for (int q = 0; q < NQ; ++q) gf_fft_pass_store<64>(j, buf0 + (size_t)q * stride, v[q]);
}

// frame `t` of the reflect-padded signal x (length n), sqrt-Hann windowed, packed as
// z[m] = x[2m] + i x[2m+1] into a padded FFT buffer.  GOOFER.py:355-369
template <typename LoadFn>
__device__ __forceinline__ void gf_load_frames(float2 *bufs, int t0, int nf, int n, const float *win, LoadFn load)
{
    for (int idx = threadIdx.x; idx < nf * 512; idx += blockDim.x) {
        const int f = idx >> 9, m = idx & 511;
        const int p = GF_HOP * (t0 + f) + 2 * m - GF_NFFT / 2;
        const bool inside = (p >= 0) && (p + 1 < n);         // reflect only near the two ends
        const float x0 = load(inside ? p : gf_reflect(p, n)) * win[2 * m];
        const float x1 = load(inside ? p + 1 : gf_reflect(p + 1, n)) * win[2 * m + 1];
        bufs[(size_t)f * GF_FFT_BUF + gf_fpad(m)] = make_float2(x0, x1);
    }
}

// ---- overlap-add ring -----------------------------------------------------------------------
// Padded-signal hop block b holds samples [256 b, 256 b + 256); frame t adds into blocks t..t+3.
// The ring keeps 8 block slots (slot = b & 7) per stream.
#define GF_RING_BLOCKS 8
#define GF_RING (GF_RING_BLOCKS * GF_HOP)

// add frames t0 .. t0+nf-1 of stream buffers (time samples = the floats of the inverse FFT buffer,
// unnormalised => scale 1/512) in ascending frame order, like _overlap_add.   GOOFER.py:380-385
__device__ __forceinline__ void gf_ola_add(float *ring, const float2 *bufs, int t0, int nf, const float *win)
{
    // positions relative to block t0: a in [0, (nf + 3) * 256)
    for (int a = threadIdx.x; a < (nf + 3) * GF_HOP; a += blockDim.x) {
        const int slot = ((t0 + (a >> 8)) & (GF_RING_BLOCKS - 1)) * GF_HOP + (a & 255);
        float acc = ring[slot];
        for (int f = 0; f < nf; ++f) {
            const int j = a - GF_HOP * f;
            if (j >= 0 && j < GF_NFFT) {
                const float2 z = bufs[(size_t)f * GF_FFT_BUF + gf_fpad(j >> 1)];
                const float v = ((j & 1) ? z.y : z.x) * (1.0f / 512.0f);
                acc = __fadd_rn(acc, __fmul_rn(v, win[j]));
            }
        }
        ring[slot] = acc;
    }
}

// finished block b -> output samples 256 (b - 2) + r, normalised by the win^2 sum over the frames
// that cover it (ascending frame order, f32), then the slot is cleared.   GOOFER.py:386-390, :402-411
__device__ __forceinline__ void gf_ola_emit(float *ring, int b, int T, int n_out, float *out, bool write)
{
    for (int r = threadIdx.x; r < GF_HOP; r += blockDim.x) {
        const int slot = (b & (GF_RING_BLOCKS - 1)) * GF_HOP + r;
        if (write) {
            float ws = 0.0f;
            for (int t = b - 3; t <= b; ++t)
                if (t >= 0 && t < T) ws = __fadd_rn(ws, d_tab.win2[GF_HOP * (b - t) + r]);
            float y = ring[slot];
            if ((double)ws > 1e-9) y = y / ws;
            const int n = GF_HOP * (b - 2) + r;
            if (n >= 0 && n < n_out) out[n] = y;
        }
        ring[slot] = 0.0f;
    }
}
