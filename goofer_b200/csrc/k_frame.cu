// k_frame.cu -- the fused frame kernel: STFT of the excitation -> spectral shaping -> three inverse
// FFTs -> windowed overlap-add, without the spectrogram ever leaving the SM.
//
// Replaces, per gf.synthesize call (/root/reference/GOOFER.py):
//   :1099        stft(pulse)                                   (:355-370)
//   :1101-1114   sigmoid high-pass at the frame f0
//   :1115-1129   env frame match, S / max|S| * env * boost     (1/max|S| is a scalar: applied later)
//   :1131-1144   voiced frames: brightness curve + 5-tap Gaussian along frequency
//   :1146        istft -> harmonic                             (:392-413, :372-390)
//   :1148-1176   U = exp(i phi); S_uv = U * env4breath; S_breath = S_uv * hp; brightness; 2 x istft
// One CTA owns a run of output hop blocks of one (note, pass); it recomputes the 3 frames of halo.
#include "gf_frame.cuh"

#define GF_RND 4                      // frames per round (= FFT lanes)
#define GF_FRAME_THREADS (64 * GF_RND)
#ifndef GF_FRAME_TMA
#define GF_FRAME_TMA 0
#endif
#ifndef GF_FRAME_CTAS
#define GF_FRAME_CTAS (GF_FRAME_TMA ? 2 : 3)   // resident CTAs per SM the register allocation aims at (67 KB of shared memory each; 92 KB with the TMA stage)
#endif
#define GF_FRAME_META 96              // frames a CTA may touch: GF_BLOCKS_PER_CTA (api.cu, static_assert there) + 3 of halo
#define GF_BLUR_K 12                  // reach of the edge correction of the time-domain blur (see gf_blur_edges)

// GF_FRAME_TMA=1: the three per-frame operand rows of the shaping stage (envF, envN, phi: 3 x 2,080 bytes, contiguous
// because all three are frame-major) are staged in shared memory by the TMA engine -- one cp.async.bulk per row, issued
// by the leader thread of the 64-thread group that owns the frame, completion signalled on the group's mbarrier -- a
// whole round ahead of their use.  25 KB more shared memory per CTA: two CTAs per SM instead of three.
#ifndef GF_FRAME_TMA
#define GF_FRAME_TMA 0
#endif

__device__ __forceinline__ unsigned gf_smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gf_mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gf_smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void gf_mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gf_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gf_mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "GF_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra GF_MBAR_DONE;\n"
        "bra GF_MBAR_WAIT;\n"
        "GF_MBAR_DONE:\n"
        "}\n" ::"r"(gf_smem_addr(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared (TMA, no tensor map): 16-byte aligned on both sides, size a multiple of 16
__device__ __forceinline__ void gf_bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(gf_smem_addr(dst)), "l"(src), "r"(bytes), "r"(gf_smem_addr(bar)) : "memory");
}

struct GfFrameSmem {
    float2 twl[GF_TWL_N];                     // per-thread FFT twiddles (gf_fft.cuh): every butterfly reads them; window and split twiddles
                                              // come from d_tab through L1 (keeps the CTA at <= 113 KB: two per SM)
    float2 z[3][GF_RND][GF_FFT_BUF];          // [0] harmonic, [1] breath, [2] unvoiced (also the forward buffer)
    float edge[2][GF_RND][4];                 // voiced frames, harmonic / breath: Im X[0], Im X[1], Im X[512], Im X[511] before the blur
    float carry[3][3][GF_HOP];                // per stream: the three hop blocks still waiting for later frames
    // per frame of the CTA (filled once in the prologue: the dependent global loads leave the per-round critical path)
    float f0fr[GF_FRAME_META];
    int voiced[GF_FRAME_META];
    int uvskip[GF_FRAME_META];                // frame lies where the smoothed mask is exactly 1: aper_uv * (1 - mask) == 0
    unsigned char dead[GF_FRAME_META];        // per owned hop block (b - b0): nobody reads the unvoiced stream there
    float red[GF_FRAME_THREADS / 32];
#if GF_FRAME_TMA
    __align__(16) float stage[GF_RND][3][GF_ENVS_LD];   // per group: envF row, envN row, phi row of its next frame (bulk copies)
    __align__(8) unsigned long long full[GF_RND];       // per group: mbarrier the three copies complete on
#endif
};

size_t gf_frame_smem_bytes() { return sizeof(GfFrameSmem); }
int gf_frame_threads() { return GF_FRAME_THREADS; }

__device__ __forceinline__ float gf_hp_sigmoid(float f, float f0)
{
    // GOOFER.py:1111  1 / (1 + exp(-clip((f - f0) / 5, -60, 60)))  (all f32)
    // (f - f0) / 5 correctly rounded in three instructions (Markstein: q = RN(a y), r = a - 5 q exact, RN(q + r y) with
    // y = RN(1 / 5); |a| <= 22,050 or 0, nothing under- or overflows)
    const float d = f - f0;
    const float q = d * 0.2f;
    float a = fmaf(fmaf(-q, 5.0f, d), 0.2f, q);
    a = fminf(fmaxf(a, -60.0f), 60.0f);
    return __fdividef(1.0f, 1.0f + __expf(-a));
}

struct GfShapeIn { float ef[2], en[2], ph[2]; };        // envF, envN, phi at bins (k, 512 - k) of one frame

// spectral shaping of one bin of one frame (GOOFER.py:1101-1173): high-pass sigmoid, envelope, boost, noise
// phases, brightness curves of voiced frames.  Returns harmonic / breath / unvoiced bins.
__device__ __forceinline__ void gf_shape_bin(int bq, float2 S, float f0f, bool vo, float ef, float en, float phi,
                                             float2 &h, float2 &b, float2 &v, float &local_max)
{
    const float hp = gf_hp_sigmoid(d_tab.freq32[bq], f0f);
    const float2 s = make_float2(S.x * hp, S.y * hp);
    local_max = fmaxf(local_max, fmaf(s.x, s.x, s.y * s.y));      // max |S|^2: sqrt taken once at the end
    const float bo = d_tab.boost[bq];
    h = make_float2(s.x * ef * bo, s.y * ef * bo);
    // U = cos(phi) + i sin(phi), phi in [0, 2 pi): evaluated at phi - pi where the fast path is accurate
    float sn, cs;
    __sincosf(phi - 3.14159274f, &sn, &cs);
    v = make_float2(-cs * en, -sn * en);
    b = make_float2(v.x * hp, v.y * hp);
    if (vo) {
        const float bh = d_tab.bright_h[bq], bb = d_tab.bright_b[bq];
        h.x *= bh; h.y *= bh; b.x *= bb; b.y *= bb;
    }
}

__device__ __forceinline__ void gf_shape_pair(GfFrameSmem &sm, int k, int f, float f0f, bool vo, bool uv_on, const GfShapeIn &in,
                                              const float2 *__restrict__ tw1024, float &local_max)
{
    const int km = 512 - k;
    float2 *zf = &sm.z[2][f][0];
    const float2 w = tw1024[k];
    float2 S0, S1;
    gf_rfft_split(zf[gf_fpad(k)], zf[gf_fpad(km & 511)], w, S0, S1);
    float2 H0, H1, B0, B1, V0, V1;
    gf_shape_bin(k, S0, f0f, vo, in.ef[0], in.en[0], in.ph[0], H0, B0, V0, local_max);
    gf_shape_bin(km, S1, f0f, vo, in.ef[1], in.en[1], in.ph[1], H1, B1, V1, local_max);
    if (vo && k <= 1) {
        // inputs of the blur's edge terms (gf_blur_edges): taken before the DC / Nyquist imaginary parts are dropped
        sm.edge[0][f][k] = H0.y; sm.edge[0][f][2 + k] = H1.y;
        sm.edge[1][f][k] = B0.y; sm.edge[1][f][2 + k] = B1.y;
    }
    // pocketfft c2r ignores the imaginary parts of DC and Nyquist
    if (k == 0) { V0.y = 0.f; V1.y = 0.f; H0.y = 0.f; H1.y = 0.f; B0.y = 0.f; B1.y = 0.f; }
    float2 Zk, Zm;
    if (uv_on) {
        gf_irfft_merge(V0, V1, w, Zk, Zm);
        zf[gf_fpad(k)] = Zk;
        if (k != 0) zf[gf_fpad(km)] = Zm;
    }
    gf_irfft_merge(H0, H1, w, Zk, Zm);
    sm.z[0][f][gf_fpad(k)] = Zk;
    if (k != 0) sm.z[0][f][gf_fpad(km)] = Zm;
    gf_irfft_merge(B0, B1, w, Zk, Zm);
    sm.z[1][f][gf_fpad(k)] = Zk;
    if (k != 0) sm.z[1][f][gf_fpad(km)] = Zm;
}

// bin 256 (pairs with itself)
__device__ __forceinline__ void gf_shape_mid(GfFrameSmem &sm, int f, float f0f, bool vo, bool uv_on, float ef, float en, float phi,
                                             const float2 *__restrict__ tw1024, float &local_max)
{
    float2 *zf = &sm.z[2][f][0];
    const float2 w = tw1024[256];
    const float2 Zq = zf[gf_fpad(256)];
    float2 S, dummy, H, B, V, Zk, Zm;
    gf_rfft_split(Zq, Zq, w, S, dummy);
    gf_shape_bin(256, S, f0f, vo, ef, en, phi, H, B, V, local_max);
    if (uv_on) {
        gf_irfft_merge(V, V, w, Zk, Zm);
        zf[gf_fpad(256)] = Zk;
    }
    gf_irfft_merge(H, H, w, Zk, Zm);
    sm.z[0][f][gf_fpad(256)] = Zk;
    gf_irfft_merge(B, B, w, Zk, Zm);
    sm.z[1][f][gf_fpad(256)] = Zk;
}

// Brightness blur of the voiced frames (GOOFER.py:1143, 1171: 5-tap Gaussian, sigma 0.5, along the 513 bins of the
// complex spectrum, numpy 'reflect' ends) WITHOUT a pass over the spectrum.  A symmetric convolution along frequency
// of a Hermitian spectrum is a multiplication in time:  irfft(g * X) = G . irfft(X),
//     G[n] = g[2] + 2 g[1] cos(2 pi n / 1024) + 2 g[0] cos(4 pi n / 1024)  in [0.57, 1],
// which folds into the synthesis window of the overlap-add (d_tab.winG = win . G).  The reference's blur differs from
// the Hermitian one only at the two ends: (i) 'reflect' continues the half spectrum with X[1], X[2] where the
// Hermitian extension has their conjugates, (ii) it smears the imaginary parts of X[0] and X[512] (non-zero for the
// noise spectra; irfft itself ignores them) into bins 1, 2, 510, 511.  Both are purely imaginary terms
//     dY[1] = i (2 g0 Im X[1] + g1 Im X[0]),  dY[2] = i g0 Im X[0]     (mirrored at the top with X[512], X[511]);
// pushed back through the inverse of the blur (taps q1, q2 = differences of IDFT(1 / G), decaying 7.4 x per bin,
// < 3e-9 at 12 bins) they become imaginary additions to bins 1..12 and 500..511 of the pre-blur spectrum.  In exact
// arithmetic the result equals the reference's (checked in numpy to 2e-14); rounding differs at the 1e-7 level.
// Called by the 64 threads that own frame f (thread j of the group): j = 2 (k - 1) + s for bin k = 1..12, stream s.
__device__ __forceinline__ void gf_blur_edges(GfFrameSmem &sm, int f, int j, const float2 *__restrict__ tw1024)
{
    if (j >= 2 * GF_BLUR_K) return;
    const int k = 1 + (j >> 1), s = j & 1;
    const float *e = sm.edge[s][f];
    const float g0 = (float)d_tab.g05[0], g1 = (float)d_tab.g05[1];
    const float p1 = 2.0f * g0 * e[1] + g1 * e[0], p2 = g0 * e[0];
    const float p1t = 2.0f * g0 * e[3] + g1 * e[2], p2t = g0 * e[2];
    const float q1 = d_tab.bq1[k], q2 = d_tab.bq2[k];
    const float2 dXk = make_float2(0.0f, p1 * q1 + p2 * q2), dXm = make_float2(0.0f, p1t * q1 + p2t * q2);
    float2 Zk, Zm;
    gf_irfft_merge(dXk, dXm, tw1024[k], Zk, Zm);
    float2 *z = &sm.z[s][f][0];
    float2 &a = z[gf_fpad(k)], &b = z[gf_fpad(512 - k)];
    a = gf_cadd(a, Zk);
    b = gf_cadd(b, Zm);
}

// One round of the windowed overlap-add for one stream (GOOFER.py:372-390, 402-411).  Frames t0 .. t0+nf-1 (nf <= 4)
// sit in `bufs` as unnormalised inverse-FFT output (scale 1/512).  Thread `tid` owns sample column tid of every hop
// block: it adds the nf x 4 windowed contributions in ascending frame order (like _overlap_add), divides the
// finished blocks by their win^2 sum, writes them, and carries the three unfinished blocks to the next round.
// One body for every stream and every round length (the stream loop of the caller is NOT unrolled, a short last
// round predicates its missing frames off): round 1 instantiated this 12 times, 3,600 of the kernel's 8,000
// instructions, and the hot loop no longer fitted the instruction cache.
__device__ __forceinline__ void gf_ola_round(float (*carry)[GF_HOP], const float2 *bufs, const float (&w)[4], const float (&wg)[4],
                                             int t0, int nf, int T, int n_out, float *__restrict__ out, int b0, int nb, bool last,
                                             bool no_input, const unsigned char *blk_dead /* indexed by b - b0 */, bool all_dead, float ws_full,
                                             const int *voiced /* per frame: use the blur-carrying window winG; NULL: never */,
                                             float rcp_full /* RN(1 / ws_full), or 0 when the exact short division does not apply */)
{
    const int r = threadIdx.x;
    float acc[GF_RND + 3];
#pragma unroll
    for (int m = 0; m < 3; ++m) acc[m] = carry[m][r];
#pragma unroll
    for (int m = 3; m < GF_RND + 3; ++m) acc[m] = 0.0f;
    if (!no_input) {                                       // skipped stream: only flush what earlier rounds carried
#pragma unroll
        for (int f = 0; f < GF_RND; ++f) {
            if (f < nf) {
                const float *zf = reinterpret_cast<const float *>(bufs + (size_t)f * GF_FFT_BUF);
                const bool blur = voiced && voiced[f];         // CTA-uniform: a branch, not four selects
                float v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int j = GF_HOP * q + r;
                    v[q] = zf[2 * gf_fpad(j >> 1) + (j & 1)];
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[f + q] = __fadd_rn(acc[f + q], __fmul_rn(v[q], blur ? wg[q] : w[q]));
            }
        }
    }
    const int n_emit = last ? nf + 1 : nf;                 // the final frame also finishes block T
#pragma unroll
    for (int m = 0; m < GF_RND + 1; ++m) {
        const int b = t0 + m;
        if (m < n_emit && b >= b0 && b < b0 + nb && b >= 2) {
            float ws = ws_full;                            // all four covering frames exist: the sum is the same for every block
            if (b < 3 || b >= T) {
                ws = 0.0f;
#pragma unroll 1
                for (int q = 3; q >= 0; --q) {             // frames b-3 .. b ascending
                    const int t = b - q;
                    if (t >= 0 && t < T) ws = __fadd_rn(ws, d_tab.win2[GF_HOP * q + r]);
                }
            }
            float y = acc[m];
            if (ws == ws_full && rcp_full != 0.0f) {
                // y / ws correctly rounded in three instructions (Markstein: q = RN(y r), rem = y - ws q exact in an FMA,
                // RN(q + rem r) with r = RN(1 / ws); ws ~ 1.998 here, significand not all ones: checked in the kernel prologue)
                const float q = y * rcp_full;
                y = fmaf(fmaf(-q, ws, y), rcp_full, q);
            } else if (ws > 1e-9f) y = y / ws;             // (double) ws > 1e-9: RN_f32(1e-9) < 1e-9, so the f32 compare selects the same floats
            const int i = GF_HOP * (b - 2) + r;
            // blocks of the unvoiced stream that nobody reads (gain (1 - mask) * 0.75 == 0 on the whole block) are not stored
            const bool dead = all_dead || (blk_dead && blk_dead[b - b0]);
            if (i < n_out && !dead) out[i] = y;
        }
    }
    // a short round is the CTA's last one: nothing reads the carry afterwards
#pragma unroll
    for (int m = 0; m < 3; ++m) carry[m][r] = acc[GF_RND + m];
}

// Excitation sample i of the (note, pass): pulse train, plus the growl layer when the note has one
// (GOOFER.py:724-736: sub *= mask; sub /= max; sub *= weight; pulse (f32) += sub (f64))
struct GfExcite {
    const float *pulse, *sub, *vm;
    double sub_scale;
    __device__ __forceinline__ float at(int i) const
    {
        return sub ? (float)((double)pulse[i] + ((double)sub[i] * (double)vm[i]) * sub_scale) : pulse[i];
    }
};

// framing of a frame that touches the reflected ends of the signal or carries the growl layer (rare: kept out of line so
// that the sixteen reflect computations do not sit in the instruction stream of the hot loop)
__device__ __noinline__ void gf_frame_edge(const GfExcite &ex, int p0, int n, float2 (&v)[8])
{
#pragma unroll 1
    for (int r = 0; r < 8; ++r) {
        const int p = p0 + 128 * r;
        const bool inside = (p >= 0) && (p + 1 < n);
        float2 x;
        x.x = ex.at(inside ? p : gf_reflect(p, n));
        x.y = ex.at(inside ? p + 1 : gf_reflect(p + 1, n));
        // v[r] with a run-time r: the array lives in registers, so select
#pragma unroll
        for (int q = 0; q < 8; ++q) if (q == r) v[q] = x;
    }
}

// work item: x = pass index (into the wave's pass arrays), y = first owned block, z = block count
//
// A round = four consecutive frames.  The 64 threads of group g = tid / 64 own frame t0 + g from the excitation samples
// to its inverse transforms -- framing straight into the registers of the first radix-8 pass, forward FFT, spectral
// shaping of ITS frame (bins j, j + 64, j + 128, j + 192 and their mirrors: coalesced envelope rows, conflict-free
// shared-memory runs), blur edge terms, two or three inverse FFTs -- synchronising only with each other (named
// barriers of 64 threads).  The CTA meets twice per round, around the overlap-add, where thread r sums column r of
// the four frames.  (Round 1 framed through shared memory and shaped (bin, frame)-interleaved: six CTA-wide barriers
// per round, a framing store + load per sample, and the groups in lock step.)
__global__ void __launch_bounds__(GF_FRAME_THREADS, GF_FRAME_CTAS)
gf_frame_kernel(const int4 *__restrict__ work, const GfPassDev *__restrict__ passes, GfPassScal *scal,
                const GfNoteDev *__restrict__ notes, const GfNotePlan *__restrict__ plans)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GfFrameSmem &sm = *reinterpret_cast<GfFrameSmem *>(smem_raw);
    const int4 wk = work[blockIdx.x];
    const GfPassDev ps = passes[wk.x];
    const GfNoteDev nd = notes[ps.note];
    const GfNotePlan &pl = plans[ps.note];
    const int n = ps.n_total, T = ps.T_out;
    const int b0 = wk.y, nb = wk.z;
    const int tid = threadIdx.x;
    const int g = tid >> 6, j = tid & 63;

    for (int i = tid; i < GF_TWL_N; i += blockDim.x) sm.twl[i] = d_tab.twl[i];
    const float *__restrict__ win = d_tab.win;
    const float2 *__restrict__ tw1024 = d_tab.tw1024;
    for (int i = tid; i < 9 * GF_HOP; i += blockDim.x) (&sm.carry[0][0][0])[i] = 0.0f;

    const int t_begin = max(0, b0 - 3), t_end = min(T - 1, b0 + nb - 1);
    const int n_f0 = (n + GF_HOP - 1) / GF_HOP;          // len(f0[::256])
    GfExcite ex;
    ex.pulse = ps.pulse; ex.sub = ps.sub; ex.vm = nd.vm; ex.sub_scale = 0.0;
    if (ex.sub) {
        const float mx = __uint_as_float(scal[wk.x].submax_bits);
        ex.sub_scale = ((double)mx > 1e-6) ? pl.subharm_weight / (double)mx : pl.subharm_weight;
    }
    const float *__restrict__ phi = ps.phi;                 // frame-major (T, GF_ENVS_LD): drawn by gf_phi_kernel or transposed by gf_phi_fm_kernel
    // global operands of a group's next frame -> L2, no registers held (each is read exactly once): the two envelope
    // rows (17 lines each), the phase row (17 lines frame-major; one line per bin otherwise), the new excitation samples
    auto prefetch_frame = [&](int t) {
        auto pf = [](const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); };
        if (!GF_FRAME_TMA && j < 17) {
            pf(reinterpret_cast<const char *>(nd.envF + (size_t)t * GF_ENVS_LD) + 128 * j);
            pf(reinterpret_cast<const char *>(nd.envN + (size_t)t * GF_ENVS_LD) + 128 * j);
            pf(reinterpret_cast<const char *>(phi + (size_t)t * GF_ENVS_LD) + 128 * j);
        } else if (j >= 17 && j < 25) {
            const int i = min(n - 1, GF_HOP * t + GF_NFFT / 2 - GF_HOP + 32 * (j - 17));
            if (i >= 0) pf(ex.pulse + i);
        }
    };
#ifndef GF_NO_PREFETCH
    if (t_begin + g <= t_end) prefetch_frame(t_begin + g);    // the first round's operands travel while the tables are staged
#endif
    for (int q = tid; q <= t_end - t_begin; q += blockDim.x) {
        const int t = t_begin + q;
        const int fi = min(t, n_f0 - 1) * GF_HOP;             // f0[::hop] edge-padded   GOOFER.py:1104-1106
        sm.f0fr[q] = ps.f0[fi];
        sm.voiced[q] = ps.mask_ones ? 1 : (nd.vm[fi] > 0.0f);    // GOOFER.py:1132-1136
        // samples of frame t live in hop blocks t-2 .. t+1 of the output
        int one = 1;
        const int nblk = (n + GF_HOP - 1) / GF_HOP;
        for (int bb = t - 2; bb <= t + 1; ++bb)
            if (bb >= 0 && bb < nblk) one &= (int)nd.ms_one[bb];
        sm.uvskip[q] = ps.mask_ones ? 1 : one;                // sa pass: uv is multiplied by 0 (SillySampler.py:1156-1170)
    }
    for (int q = tid; q < nb; q += blockDim.x) sm.dead[q] = nd.ms_one[b0 + q - 2];
    float local_max = 0.0f;
    // win^2 sum of an interior hop block (frames b-3 .. b all exist), summed in ascending frame order like the general path
    float ws_full = 0.0f;
#pragma unroll
    for (int q = 3; q >= 0; --q) ws_full = __fadd_rn(ws_full, d_tab.win2[GF_HOP * q + tid]);
    const float rcp_full = (ws_full > 1e-9f && (__float_as_uint(ws_full) & 0x7fffffu) != 0x7fffffu) ? __frcp_rn(ws_full) : 0.0f;
#if GF_FRAME_TMA
    if (j == 0) gf_mbar_init(&sm.full[g], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    unsigned tma_parity = 0;
    // the leader of a group requests the three operand rows of frame t; they land in the group's stage while it works on
    auto stage_frame = [&](int t) {
        constexpr unsigned ROW = GF_ENVS_LD * sizeof(float);
        gf_mbar_expect_tx(&sm.full[g], 3 * ROW);
        gf_bulk_g2s(&sm.stage[g][0][0], nd.envF + (size_t)t * GF_ENVS_LD, ROW, &sm.full[g]);
        gf_bulk_g2s(&sm.stage[g][1][0], nd.envN + (size_t)t * GF_ENVS_LD, ROW, &sm.full[g]);
        gf_bulk_g2s(&sm.stage[g][2][0], phi + (size_t)t * GF_ENVS_LD, ROW, &sm.full[g]);
    };
#endif
    __syncthreads();
#if GF_FRAME_TMA
    if (j == 0 && t_begin + g <= t_end) stage_frame(t_begin + g);
#endif

    for (int t0 = t_begin; t0 <= t_end; t0 += GF_RND) {
        const int nf = min(GF_RND, t_end - t0 + 1);
        const int m0r = t0 - t_begin;                         // this round's first entry of the per-frame tables
        bool uv_on = false;                                   // uniform: the round computes the unvoiced stream unless every frame may skip it
        for (int q = 0; q < nf; ++q) uv_on = uv_on || (sm.uvskip[m0r + q] == 0);
        if (g < nf) {
            const int f = g, t = t0 + g;
            float2 *zf = &sm.z[2][f][0];
            // ---- 1. frame the excitation into the registers of the first radix-8 pass (element m = j + 64 r of the packed
            //         sequence z[m] = x[2m] + i x[2m+1]; numpy 'reflect' padding; sqrt-Hann)   GOOFER.py:355-369 ----
            float2 v[8];
            {
                const int p0 = GF_HOP * t + 2 * j - GF_NFFT / 2;               // even
                if (!ex.sub && p0 >= 0 && p0 + 2 * 64 * 7 + 1 < n) {
#pragma unroll
                    for (int r = 0; r < 8; ++r) v[r] = *reinterpret_cast<const float2 *>(ex.pulse + p0 + 128 * r);
                } else {
                    gf_frame_edge(ex, p0, n, v);                  // note ends / growl layer: out of line, off the hot path
                }
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const float2 w = *reinterpret_cast<const float2 *>(win + 2 * j + 128 * r);
                    v[r].x *= w.x; v[r].y *= w.y;
                }
            }
#ifndef GF_NO_PREFETCH
            if (t + GF_RND <= t_end) prefetch_frame(t + GF_RND);
#endif
            // ---- 2. forward FFT of the group's frame ----
            gf_dft8<false>(v);
            gf_fft_pass_store<1>(j, zf, v);
            gf_lane_sync(g);
            gf_fft_pass_load<false, 8>(j, zf, sm.twl, v);
            gf_lane_sync(g);
            gf_fft_pass_store<8>(j, zf, v);
            gf_lane_sync(g);
            gf_fft_pass_load<false, 64>(j, zf, sm.twl, v);
            gf_lane_sync(g);
            gf_fft_pass_store<64>(j, zf, v);
            gf_lane_sync(g);
            // ---- 3. shaping, per bin pair (k, 512 - k), k = j + 64 m; bin 256 pairs with itself (thread 0 of the group) ----
            {
#if GF_FRAME_TMA
                gf_mbar_wait(&sm.full[g], tma_parity);        // the rows requested a round ago have landed
                tma_parity ^= 1u;
                const float *eF = &sm.stage[g][0][0], *eN = &sm.stage[g][1][0], *ph = &sm.stage[g][2][0];
#else
                const float *eF = nd.envF + (size_t)t * GF_ENVS_LD, *eN = nd.envN + (size_t)t * GF_ENVS_LD;
                const float *ph = phi + (size_t)t * GF_ENVS_LD;
#endif
                const float f0f = sm.f0fr[m0r + f];
                const bool vo = sm.voiced[m0r + f] != 0;
#ifndef GF_SHAPE_BATCH
#define GF_SHAPE_BATCH 2              // bin pairs whose global operands are requested together (6 loads each)
#endif
#ifdef GF_SHAPE_UNROLL
#pragma unroll
#else
#pragma unroll 1
#endif
                for (int mb = 0; mb < 4; mb += GF_SHAPE_BATCH) {
                    GfShapeIn in[GF_SHAPE_BATCH];
#pragma unroll
                    for (int m = 0; m < GF_SHAPE_BATCH; ++m) {
                        const int k = j + 64 * (mb + m), km = 512 - k;
                        in[m].ef[0] = eF[k];  in[m].ef[1] = eF[km];
                        in[m].en[0] = eN[k];  in[m].en[1] = eN[km];
                        in[m].ph[0] = ph[k];  in[m].ph[1] = ph[km];
                    }
#pragma unroll
                    for (int m = 0; m < GF_SHAPE_BATCH; ++m) gf_shape_pair(sm, j + 64 * (mb + m), f, f0f, vo, uv_on, in[m], tw1024, local_max);
                }
                if (j == 0) gf_shape_mid(sm, f, f0f, vo, uv_on, eF[256], eN[256], ph[256], tw1024, local_max);
                gf_lane_sync(g);
#if GF_FRAME_TMA
                if (j == 0 && t + GF_RND <= t_end) stage_frame(t + GF_RND);       // every read of the stage is behind the barrier above
#endif
                // ---- 4. voiced frames: edge terms of the brightness blur (the blur itself rides on the synthesis window) ----
                if (vo) gf_blur_edges(sm, f, j, tw1024);      // uniform over the group
                gf_lane_sync(g);
            }
            // ---- 5. inverse FFTs of the group's frame: harmonic, breath (, unvoiced), advanced pass by pass together ----
            // (the unvoiced stream is rare -- only where the smoothed mask is below 1 -- and goes through a second, single
            // transform call instead of a third instantiation of the batch: code size)
            gf_group_ifft512<2>(&sm.z[0][f][0], GF_RND * GF_FFT_BUF, sm.twl, g, j);
            if (uv_on) gf_group_ifft512<1>(&sm.z[2][f][0], GF_RND * GF_FFT_BUF, sm.twl, g, j);
        }
        __syncthreads();
        // ---- 6. overlap-add + emit: thread `tid` owns column tid of every hop block ----
        {
            const bool last = (t0 + nf - 1 == T - 1);
            float w[4], wg[4];
#pragma unroll
            // the 1/512 of the unnormalised inverse FFT is folded into the window: scaling by a power of two commutes with rounding
            for (int q = 0; q < 4; ++q) { w[q] = win[GF_HOP * q + tid] * (1.0f / 512.0f); wg[q] = d_tab.winG[GF_HOP * q + tid] * (1.0f / 512.0f); }
#pragma unroll 1
            for (int s = 0; s < 3; ++s) {
                float *out = s == 0 ? ps.harm : (s == 1 ? ps.bre : ps.uv);
                gf_ola_round(sm.carry[s], &sm.z[s][0][0], w, wg, t0, nf, T, n, out, b0, nb, last, s == 2 && !uv_on, s == 2 ? sm.dead : nullptr,
                             s == 2 && ps.mask_ones, ws_full, s < 2 ? sm.voiced + m0r : nullptr, rcp_full);
            }
        }
        __syncthreads();
    }
    // zero tail of istft: samples 256 (T - 1) .. n - 1     GOOFER.py:407-409
    if (b0 + nb > T) {
        for (int i = GF_HOP * (T - 1) + tid; i < n; i += blockDim.x) { ps.harm[i] = 0.f; ps.bre[i] = 0.f; ps.uv[i] = 0.f; }   // (the 68-sample tail: always stored)
    }
    // max(|S| + 1e-8) over the whole (note, pass)          GOOFER.py:1121
    local_max = gf_warp_max(local_max);
    if ((tid & 31) == 0) sm.red[tid >> 5] = local_max;
    __syncthreads();
    if (tid == 0) {
        float m = 0.f;
        for (int w = 0; w < GF_FRAME_THREADS / 32; ++w) m = fmaxf(m, sm.red[w]);
        gf_atomic_max_pos(&scal[wk.x].mag_bits, sqrtf(m) + 1e-8f);
    }
}

void gf_launch_frame(const int4 *work, int n_work, const GfPassDev *passes, GfPassScal *scal, const GfNoteDev *notes,
                     const GfNotePlan *plans, cudaStream_t st)
{
    if (n_work <= 0) return;
    static GfSmemLimit memo;
    gf_smem_limit(gf_frame_kernel, sizeof(GfFrameSmem), memo);
    gf_frame_kernel<<<n_work, GF_FRAME_THREADS, sizeof(GfFrameSmem), st>>>(work, passes, scal, notes, plans);
}
