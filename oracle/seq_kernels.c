/* ORACLE (test infrastructure, never shipped, never on the product path).
 *
 * Plain-C restatement of the reference's sequential numba / pure-Python loops, used by
 * oracle/goofer_oracle.py.  Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: no FMA
 * contraction, no fast-math, so every operation rounds exactly once like the reference's
 * scalar code).
 *
 *   orc_pulse_train      <- GOOFER.py:473-554  pulse_train_numba
 *   orc_sub_events       <- GOOFER.py:672-698  _detect_pulse_events (one ratio)
 *   orc_onepole_cascade  <- SillySampler.py:154-174 (_dynamic_butter_filter_core recurrences)
 *   orc_overlap_add      <- GOOFER.py:372-390  _overlap_add
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>

#define ORC_PI 3.141592653589793

/* LF pulse table for one period length T0 (GOOFER.py:507-528). T = 1/f0 in seconds. */
static void orc_lf_table(double T, long T0, double Ra, double Rg, double Rk, float *buf)
{
    double Ta = Ra * T, Te = T, Tp = Ta, Tc = Tp + Rk * (Te - Tp);
    for (long j = 0; j < T0; ++j) {
        double ti = ((double)j * T) / (double)T0;
        double v;
        if (ti < Tp) {
            double s = sin(ORC_PI * ti / (2.0 * Tp + 1e-12));
            v = s * s;
        } else if (ti < Tc) {
            double tau = (ti - Tp) / (Tc - Tp + 1e-12);
            v = exp(-Rg * tau) * cos(ORC_PI * tau / 2.0);
        } else {
            v = 0.0;
        }
        buf[j] = (float)v;
    }
    double m = 0.0;
    for (long j = 0; j < T0; ++j) {
        float a = fabsf(buf[j]);
        if ((double)a > m) m = (double)a;
    }
    if (m > 0.0)
        for (long j = 0; j < T0; ++j) buf[j] = (float)((double)buf[j] / m);
}

/* GOOFER.py:473-554.  out must hold n floats (overwritten).
 * onset_idx/onset_T0 (optional, may be NULL) receive up to max_onsets onsets; returns onset count. */
long orc_pulse_train(const float *f0, long n, double sr, double Ra, double Rg, double Rk,
                     float *out, int32_t *onset_idx, int32_t *onset_T0, long max_onsets)
{
    memset(out, 0, sizeof(float) * (size_t)n);
    double total_phase = 0.0, next_k = 1.0, last_valid_f0 = 160.0;
    long cache_T0[5] = {0, 0, 0, 0, 0};
    int cache_len = 0;
    float *bank = (float *)calloc(5 * 8192, sizeof(float));
    float *tmp = (float *)malloc(8192 * sizeof(float));
    long n_on = 0;
    for (long i = 0; i < n; ++i) {
        float f0i = f0[i];
        if ((double)f0i > 1e-6) last_valid_f0 = (double)f0i;   /* numba compares f32 against an f64 constant */
        total_phase += (double)f0i / sr;
        while (total_phase >= next_k) {
            double lv = last_valid_f0 > 1e-6 ? last_valid_f0 : 1e-6;
            double T = 1.0 / lv;
            long T0 = (long)nearbyint(sr * T);           /* round-half-even like Python round() */
            if (T0 < 3) T0 = 3;
            if (T0 > 8192) T0 = 8192;
            int found = -1;
            for (int c = 0; c < cache_len; ++c)
                if (cache_T0[c] == T0) { found = c; break; }
            if (found < 0) {
                orc_lf_table(T, T0, Ra, Rg, Rk, tmp);
                if (cache_len < 5) { found = cache_len++; } else { found = 0; }
                cache_T0[found] = T0;
                memcpy(bank + (size_t)found * 8192, tmp, sizeof(float) * (size_t)T0);
            }
            long end = i + T0; if (end > n) end = n;
            const float *src = bank + (size_t)found * 8192;
            for (long j = i, k = 0; j < end; ++j, ++k) out[j] += src[k];
            if (onset_idx && n_on < max_onsets) { onset_idx[n_on] = (int32_t)i; onset_T0[n_on] = (int32_t)T0; }
            ++n_on;
            next_k += 1.0;
        }
    }
    free(bank); free(tmp);
    return n_on;
}

/* GOOFER.py:672-698 for a single ratio. Returns number of events written (<= max_ev). */
long orc_sub_events(const double *f0, const double *mask, long n, double sr, double ratio,
                    int32_t *ev_idx, double *ev_f0, long max_ev)
{
    double last_f0 = 160.0, phase = 0.0;
    long ne = 0;
    for (long i = 0; i < n; ++i) {
        double f = f0[i];
        if (mask[i] <= 0.0 || f <= 0.0) continue;
        last_f0 = f;
        double sub = last_f0 * ratio;
        if (sub < 1e-2) continue;
        phase += sub / sr;
        if (phase >= 1.0) {
            if (ne < max_ev) { ev_idx[ne] = (int32_t)i; ev_f0[ne] = sub; }
            ++ne;
            phase -= 1.0;
        }
    }
    return ne;
}

/* SillySampler.py:154-174.  y is filtered in place, `order` passes, float32 arithmetic. */
void orc_onepole_cascade(float *y, const float *alpha, long n, int order, int highpass)
{
    if (order < 1) order = 1;
    for (int p = 0; p < order; ++p) {
        if (!highpass) {
            float yp = 0.0f;
            for (long i = 0; i < n; ++i) {
                float a = alpha[i], xp = y[i];
                float d = xp - yp;
                yp = fmaf(a, d, yp);     /* numba fastmath contracts yp + a*(xp-yp) into one FMA (probe) */
                y[i] = yp;
            }
        } else {
            float yp = 0.0f;
            float prev_x = n > 0 ? y[0] : 0.0f;
            for (long i = 0; i < n; ++i) {
                float a = alpha[i], xp = y[i];
                float s = yp - prev_x;   /* numba fastmath (LLVM reassociate) evaluates (yp - prev_x) + xp (probe: */
                s = s + xp;              /* bit-identical to the reference on 44,100/44,100 samples)            */
                yp = a * s;
                y[i] = yp;
                prev_x = xp;
            }
        }
    }
}

/* GOOFER.py:372-390.  frames is (n_fft, n_frames) row-major float32. */
void orc_overlap_add(const float *frames, const float *window, long n_fft, long n_frames, long hop,
                     float *y, long expected_len)
{
    float *ws = (float *)calloc((size_t)expected_len, sizeof(float));
    memset(y, 0, sizeof(float) * (size_t)expected_len);
    for (long i = 0; i < n_frames; ++i) {
        long start = i * hop;
        for (long j = 0; j < n_fft; ++j) {
            long idx = start + j;
            float val = frames[j * n_frames + i] * window[j];
            y[idx] += val;
            float w2 = window[j] * window[j];
            ws[idx] += w2;
        }
    }
    for (long i = 0; i < expected_len; ++i)
        if ((double)ws[i] > 1e-9) y[i] /= ws[i];
    free(ws);
}
