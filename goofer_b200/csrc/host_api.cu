// host_api.cu -- goofer_render_batch_host: the same render with HOST buffers (what a ctypes / cgo
// caller that owns numpy arrays passes).  Copies sources, noise and bends in, renders, copies the
// output back; device and pinned staging buffers are cached per host thread.
#include <cstdlib>
#include <cstdint>
#include <climits>
#include <ctime>

struct GfHostCache {
    int device = -1;                                                   // the CUDA device the buffers, streams and events belong to
    void *dev = nullptr;  size_t dev_cap = 0;
    void *ws = nullptr;   size_t ws_cap = 0;
    cudaStream_t st = nullptr, st_in = nullptr, st_out = nullptr;     // compute, H2D, D2H
    std::vector<cudaEvent_t> ev;
};
static thread_local GfHostCache g_hc;

static int gf_hc_reserve(void **p, size_t *cap, size_t need)
{
    if (*cap >= need) return GOOFER_OK;
    if (*p) { cudaFree(*p); *p = nullptr; *cap = 0; }
    need = need + need / 8 + (1 << 20);
    GF_CUDA(cudaMalloc(p, need));
    *cap = need;
    return GOOFER_OK;
}

// ------------------------------------------------------------------------------------------------
// Gather copy of many medium-sized page-locked host arrays (a voicebank's knot packs and voicing masks: two arrays of
// 60-180 KB per source) into device memory by ONE kernel that reads the host memory directly over PCIe, instead of one
// cudaMemcpyAsync per array (~5 us of host time each: 130 calls held the preparation kernels back by 0.7 ms).
// One CTA per 16 KB block of one array.
// ------------------------------------------------------------------------------------------------
struct GfPullJob { const char *src; char *dst; unsigned long long bytes; unsigned int first_block; unsigned int pad; };
#define GF_PULL_BLOCK 16384

__global__ void __launch_bounds__(256) gf_pull_kernel(const GfPullJob *__restrict__ jobs, int n_jobs)
{
    // job of this CTA: last one whose first_block <= blockIdx.x
    int lo = 0, hi = n_jobs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (jobs[mid].first_block <= blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const GfPullJob jb = jobs[lo];
    const unsigned long long off = (unsigned long long)(blockIdx.x - jb.first_block) * GF_PULL_BLOCK;
    const unsigned long long cnt = min((unsigned long long)GF_PULL_BLOCK, jb.bytes - off);
    const char *src = jb.src + off;
    char *dst = jb.dst + off;
    if ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
        const int n16 = (int)(cnt >> 4);
        const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
        uint4 *d4 = reinterpret_cast<uint4 *>(dst);
        uint4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { const int i = threadIdx.x + 256 * k; if (i < n16) v[k] = s4[i]; }
#pragma unroll
        for (int k = 0; k < 4; ++k) { const int i = threadIdx.x + 256 * k; if (i < n16) d4[i] = v[k]; }
        for (int i = (n16 << 4) + threadIdx.x; i < (int)cnt; i += 256) dst[i] = src[i];
    } else {
        for (int i = threadIdx.x; i < (int)cnt; i += 256) dst[i] = src[i];
    }
}

// frees what this thread cached ON THE DEVICE IT WAS CREATED FOR (the caller may have moved to another GPU since)
static void gf_hc_free()
{
    int cur = -1;
    cudaGetDevice(&cur);
    if (g_hc.device >= 0 && g_hc.device != cur) cudaSetDevice(g_hc.device);
    if (g_hc.st) cudaStreamSynchronize(g_hc.st);
    if (g_hc.st_in) cudaStreamSynchronize(g_hc.st_in);
    if (g_hc.st_out) cudaStreamSynchronize(g_hc.st_out);
    if (g_hc.dev) cudaFree(g_hc.dev);
    if (g_hc.ws) cudaFree(g_hc.ws);
    for (cudaEvent_t e : g_hc.ev) cudaEventDestroy(e);
    if (g_hc.st) cudaStreamDestroy(g_hc.st);
    if (g_hc.st_in) cudaStreamDestroy(g_hc.st_in);
    if (g_hc.st_out) cudaStreamDestroy(g_hc.st_out);
    g_hc = GfHostCache();
    if (cur >= 0) cudaSetDevice(cur);
}

// The thread-local caches have no destructor (CUDA may be gone by the time a thread's statics are torn down): a host
// thread that used goofer_render_batch_host calls this before it exits.
extern "C" void goofer_host_release(void)
{
    gf_hc_free();
    if (g_side.sx) {
        int cur = -1;
        cudaGetDevice(&cur);
        if (g_side.dev >= 0 && g_side.dev != cur) cudaSetDevice(g_side.dev);
        cudaStreamSynchronize(g_side.sx);
        cudaStreamDestroy(g_side.sx); cudaEventDestroy(g_side.fork); cudaEventDestroy(g_side.join);
        if (g_side.alt) { cudaStreamSynchronize(g_side.alt); cudaStreamDestroy(g_side.alt); cudaEventDestroy(g_side.prep); cudaEventDestroy(g_side.alt_done); }
        g_side = GfSide();
        if (cur >= 0) cudaSetDevice(cur);
    }
    if (g_pin.base) {
        if (g_pin.dev >= 0) { int cur = -1; cudaGetDevice(&cur); cudaSetDevice(g_pin.dev); cudaDeviceSynchronize(); if (cur >= 0) cudaSetDevice(cur); }
        cudaFreeHost(g_pin.base);
        g_pin = GfPinned();
    }
}

// Only the frame kernel reads the noise phases (2/3 of the input bytes).  The batch is rendered ONCE: the
// preparation kernels (tracks, f0, walk, pulse, env) of all notes run at full width while the phases stream in on
// st_in; frame / peak / mix then go part by part (a part = a run of notes whose phases arrived together), and a
// part's output leaves on st_out while the next part computes (PCIe is full duplex).
static int gf_render_batch_host_impl(const GooferBatch *b);

extern "C" int goofer_render_batch_host(const GooferBatch *b)
{
    const int rc = gf_render_batch_host_impl(b);
    if (rc != GOOFER_OK) {
        // copies from / into the caller's host buffers may still be in flight: the caller is free to release or
        // overwrite them once this call returns
        if (g_hc.st_in) cudaStreamSynchronize(g_hc.st_in);
        if (g_hc.st) cudaStreamSynchronize(g_hc.st);
        if (g_hc.st_out) cudaStreamSynchronize(g_hc.st_out);
        cudaGetLastError();
    }
    return rc;
}

static int gf_render_batch_host_impl(const GooferBatch *b)
{
    int rc = gf_validate(b);
    if (rc != GOOFER_OK) return rc;
    g_stats.h2d_bytes = 0; g_stats.d2h_bytes = 0; g_stats.kernel_launches = 0; g_stats.waves = 0;
    if (b->n_notes == 0) return GOOFER_OK;
    if ((!b->out && !b->out_pcm16) || !b->bend_cents) { gf_set_error("NULL out (and out_pcm16) / bend_cents"); return GOOFER_ERR_INVALID; }
    {
        int cur = -1;
        GF_CUDA(cudaGetDevice(&cur));
        if (g_hc.device != cur) { gf_hc_free(); g_hc.device = cur; }      // the thread moved to another GPU: its cache goes with the old one
    }
    if (!g_hc.st) GF_CUDA(cudaStreamCreateWithFlags(&g_hc.st, cudaStreamNonBlocking));
    if (!g_hc.st_in) GF_CUDA(cudaStreamCreateWithFlags(&g_hc.st_in, cudaStreamNonBlocking));
    if (!g_hc.st_out) GF_CUDA(cudaStreamCreateWithFlags(&g_hc.st_out, cudaStreamNonBlocking));
    cudaStream_t st = g_hc.st, st_in = g_hc.st_in, st_out = g_hc.st_out;
    const bool trace = getenv("GOOFER_HOST_TRACE") != nullptr;
    auto now_ms = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return 1e3 * (double)ts.tv_sec + 1e-6 * (double)ts.tv_nsec; };
    const double h0 = trace ? now_ms() : 0.0;
    g_htrace.on = trace;
    g_htrace.marks.clear();
    double h_first_copy = 0.0, h_enqueued = 0.0, h_uploads = 0.0, h_sized = 0.0;
    static thread_local cudaEvent_t ev_t0 = nullptr, ev_d2h[64] = {nullptr};
    if (trace) {
        if (!ev_t0) cudaEventCreate(&ev_t0);
        cudaEventRecord(ev_t0, st_in);
    }

    // ---- device image of every input array (same element offsets as on the host) ----
    Bump sz{nullptr, 0, 0};
    const void *small_dev = nullptr;
    GfPullJob *pull_tab = nullptr;
    auto carve = [&](Bump &bp, GooferBatch &db, std::vector<GooferSource> &ds) {
        for (int s = 0; s < b->n_sources; ++s) ds[s] = b->sources[s];
        // small arrays back to back, each padded to 256 bytes (Bump aligns every array to 256)
        bool first = true;
        for (int s = 0; s < b->n_sources; ++s) {
            const GooferSource &g = b->sources[s];
            if (g.hz_knots) { ds[s].hz_knots = bp.arr<float>((size_t)g.K); if (first) { small_dev = ds[s].hz_knots; first = false; } }
            for (int k = 0; k < 4; ++k)
                if (g.formants[k]) { ds[s].formants[k] = bp.arr<double>((size_t)g.formant_len[k]); if (first) { small_dev = ds[s].formants[k]; first = false; } }
        }
        for (int s = 0; s < b->n_sources; ++s) {
            const GooferSource &g = b->sources[s];
            if (g.knots_log_f16) ds[s].knots_log_f16 = bp.arr<uint16_t>((size_t)g.K * g.T);
            if (g.env_dense) ds[s].env_dense = bp.arr<float>((size_t)GF_NBINS * g.T);
            if (g.mask) ds[s].mask = bp.arr<float>((size_t)g.N);
        }
        db.bend_cents = bp.arr<float>((size_t)b->bend_total);
        db.phi = b->phi ? bp.arr<float>((size_t)std::max<int64_t>(b->phi_total, 1)) : nullptr;
        db.normals = b->normals ? bp.arr<double>((size_t)b->nrm_total) : nullptr;
        db.f0_curves = b->f0_curves ? bp.arr<float>((size_t)b->f0_total) : nullptr;
        db.out = b->out ? bp.arr<float>((size_t)b->out_total) : nullptr;
        db.out_pcm16 = b->out_pcm16 ? bp.arr<int16_t>((size_t)b->out_total) : nullptr;
        db.tap_harm = b->tap_harm ? bp.arr<float>((size_t)b->out_total) : nullptr;
        db.tap_uv = b->tap_uv ? bp.arr<float>((size_t)b->out_total) : nullptr;
        db.tap_bre = b->tap_bre ? bp.arr<float>((size_t)b->out_total) : nullptr;
        pull_tab = bp.arr<GfPullJob>((size_t)3 * b->n_sources + 8);
    };
    GooferBatch db = *b;
    std::vector<GooferSource> ds(b->n_sources);
    carve(sz, db, ds);
    if ((rc = gf_hc_reserve(&g_hc.dev, &g_hc.dev_cap, sz.off + 4096)) != GOOFER_OK) return rc;
    Bump bp{(char *)g_hc.dev, g_hc.dev_cap, 0};
    carve(bp, db, ds);
    db.sources = ds.data();

    auto h2d = [&](const void *dst, const void *src, size_t bytes) -> int {
        if (!bytes) return GOOFER_OK;
        GF_CUDA(cudaMemcpyAsync(const_cast<void *>(dst), src, bytes, cudaMemcpyHostToDevice, st_in));
        g_stats.h2d_bytes += (int64_t)bytes;
        return GOOFER_OK;
    };
    auto d2h = [&](void *dst, const void *src, size_t bytes) -> int {
        if (!bytes) return GOOFER_OK;
        GF_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st_out));
        g_stats.d2h_bytes += (int64_t)bytes;
        return GOOFER_OK;
    };
    // The sources and bends: their places in the device image depend on the descriptor alone, so they need no planning
    // and go up FIRST, before the notes are planned: the 0.3 ms they spend on the wire (plus 0.1 ms of host time to issue
    // them) hide behind the planning / carving / list building the host does next.  (Alternative, kept as a knob: issue
    // them right after the first wave's phase generator has been launched -- gf_render_wave calls `deferred`.)
    auto upload_sources = [&]() -> int {
        int rc = GOOFER_OK;
    if (trace) h_first_copy = now_ms();
    // The small per-source arrays (mel-knot frequencies, four formant tracks: a few KB each) are gathered in the
    // pinned staging arena and go up as ONE copy -- hundreds of tiny cudaMemcpyAsync calls cost more host time
    // than the transfer itself.  They were carved back to back (see `carve`), so one device range covers them.
    {
        size_t small_bytes = 0;
        for (int s = 0; s < b->n_sources; ++s) {
            const GooferSource &g = b->sources[s];
            if (g.hz_knots) small_bytes += (sizeof(float) * (size_t)g.K + 255) & ~(size_t)255;
            for (int k = 0; k < 4; ++k) if (g.formants[k]) small_bytes += (sizeof(double) * (size_t)g.formant_len[k] + 255) & ~(size_t)255;
        }
        if (small_bytes) {
            char *stage = (char *)gf_pin_take(small_bytes);
            if (!stage) { gf_set_error("cudaMallocHost failed for the source staging arena"); return GOOFER_ERR_CUDA; }
            size_t off = 0;
            for (int s = 0; s < b->n_sources; ++s) {
                const GooferSource &g = b->sources[s];
                if (g.hz_knots) { std::memcpy(stage + off, g.hz_knots, sizeof(float) * (size_t)g.K); off += (sizeof(float) * (size_t)g.K + 255) & ~(size_t)255; }
                for (int k = 0; k < 4; ++k)
                    if (g.formants[k]) { std::memcpy(stage + off, g.formants[k], sizeof(double) * (size_t)g.formant_len[k]); off += (sizeof(double) * (size_t)g.formant_len[k] + 255) & ~(size_t)255; }
            }
            if ((rc = h2d(small_dev, stage, small_bytes))) return rc;
        }
    }
    // source envelopes, voicing masks, bends: one gather kernel when every array is page-locked (see gf_pull_kernel),
    // else one copy-engine transfer per array
    {
        struct Cp { const void *dst, *src; size_t bytes; };
        std::vector<Cp> cps;
        for (int s = 0; s < b->n_sources; ++s) {
            const GooferSource &g = b->sources[s];
            const GooferSource &d = ds[s];
            if (g.knots_log_f16) cps.push_back({d.knots_log_f16, g.knots_log_f16, sizeof(uint16_t) * (size_t)g.K * g.T});
            if (g.env_dense) cps.push_back({d.env_dense, g.env_dense, sizeof(float) * (size_t)GF_NBINS * g.T});
            if (g.mask) cps.push_back({d.mask, g.mask, sizeof(float) * (size_t)g.N});
        }
        cps.push_back({db.bend_cents, b->bend_cents, sizeof(float) * (size_t)b->bend_total});
        bool pull = cps.size() > 8 && !getenv("GOOFER_HOST_NO_PULL");
        std::vector<GfPullJob> jobs;
        unsigned int blocks = 0;
        for (size_t i = 0; pull && i < cps.size(); ++i) {
            if (!cps[i].bytes) continue;
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, cps[i].src) != cudaSuccess || at.type != cudaMemoryTypeHost || !at.devicePointer) {
                cudaGetLastError();
                pull = false;
                break;
            }
            jobs.push_back({(const char *)at.devicePointer, (char *)const_cast<void *>(cps[i].dst), (unsigned long long)cps[i].bytes, blocks, 0u});
            blocks += (unsigned int)((cps[i].bytes + GF_PULL_BLOCK - 1) / GF_PULL_BLOCK);
        }
        if (pull && !jobs.empty()) {
            GfPullJob *stage = (GfPullJob *)gf_pin_take(jobs.size() * sizeof(GfPullJob));
            if (!stage) { gf_set_error("cudaMallocHost failed for the gather table"); return GOOFER_ERR_CUDA; }
            std::memcpy(stage, jobs.data(), jobs.size() * sizeof(GfPullJob));
            // on the upload stream: the first kernels that need the sources wait for its event (src_ready below), the
            // phase generator of the first wave runs beside it
            if ((rc = gf_meta_copy(pull_tab, stage, jobs.size() * sizeof(GfPullJob), st_in)) != GOOFER_OK) return rc;
            gf_pull_kernel<<<blocks, 256, 0, st_in>>>(pull_tab, (int)jobs.size());
            GF_CUDA(cudaGetLastError());
            for (const Cp &c : cps) g_stats.h2d_bytes += (int64_t)c.bytes;
        } else {
            for (const Cp &c : cps)
                if ((rc = h2d(c.dst, c.src, c.bytes))) return rc;
        }
    }


        return GOOFER_OK;
    };
    // Measured on B200 (c2, device-drawn phases, PCM16 out): issued first 5.48-5.50 ms per call, deferred behind the generator
    // 5.65-5.67 ms (the sources then reach HBM at 0.91 ms instead of 0.46 ms and the preparation kernels wait for them):
    // the deferral is an opt-in knob (GOOFER_HOST_DEFER=1), off by default.
    const bool defer_sources = (!b->phi || b->phi_total <= 1) && getenv("GOOFER_HOST_DEFER") != nullptr;
    if (!defer_sources && (rc = upload_sources()) != GOOFER_OK) return rc;

    std::vector<GfNotePlan> plans;
    if ((rc = gf_make_plans(b, plans)) != GOOFER_OK) return rc;
    gf_htrace("host: notes planned");

    // ---- parts: runs of consecutive notes whose phases travel together ----
    // Measured on B200 (c2, 1,024 notes, 5.7 ms of kernels): the phase upload takes 7.4 ms (380 MB at 52 GB/s) and the
    // download 4.5 ms (180 MB at ~40 GB/s while the upload runs); the first output exists only after the preparation
    // kernels of the whole batch (~4 ms).  Small equal parts start the download early and leave little to do after the
    // last phase byte: 8 parts 8.7 ms, 4 parts 9.3 ms, graded 50/25/15/10 10.0 ms.  Parts are sized by phase bytes, so 96
    // sixteen-second notes go in 12 parts (18.3 -> 12.0 ms).  GOOFER_HOST_CHUNK = uniform parts of that many notes;
    // GOOFER_HOST_PARTS = comma-separated cumulative fractions.
    std::vector<int> ends;                                 // note_end of every part, ascending, last == n_notes
    {
        const int nn = b->n_notes;
        const char *e = getenv("GOOFER_HOST_CHUNK"), *f = getenv("GOOFER_HOST_PARTS");
        if (e && atoi(e) > 0) {
            for (int i = atoi(e); i < nn; i += atoi(e)) ends.push_back(i);
        } else if (f && f[0]) {
            for (const char *q = f; *q;) {
                char *nx = nullptr;
                const double fr = strtod(q, &nx);
                if (nx == q) break;
                const int v = (int)(fr * nn);
                if (v > (ends.empty() ? 0 : ends.back()) && v < nn) ends.push_back(v);
                q = (*nx == ',') ? nx + 1 : nx;
            }
        } else {
            // parts of about 45 MB of noise phases each (1,024 one-second notes: 8 parts; 96 sixteen-second notes: 12),
            // cut where the cumulative phase bytes cross k / parts of the total, at most 16, at least 2 above 8 MB
            std::vector<int64_t> cum(nn + 1, 0);
            int64_t up_total = 0;                              // phase bytes that really cross PCIe (not the device-drawn slots)
            for (int i = 0; i < nn; ++i) {
                const GfNotePlan &p = plans[i];
                for (int k = 0; k < p.n_passes; ++k)
                    if (!((p.phi_rng_mask >> p.pass_kind[k]) & 1u)) up_total += (int64_t)GF_NBINS * p.T_out * 4;
            }
            const bool upload_bound = up_total >= (8 << 20);
            // cut by uploaded phase bytes when phases are uploaded, else by the bytes that go back (the parts then only
            // exist to start the download early: with device-drawn phases and PCM16 output a 1,024-note batch returns
            // 90 MB, four parts; every part's frame launch should still fill the GPU -- 128-note parts ran the
            // frame / peak / mix tail 45 % slower than one launch, measured on B200)
            const int64_t per_sample_down = (b->out ? 4 : 0) + (b->out_pcm16 ? 2 : 0) + (b->tap_harm ? 4 : 0) + (b->tap_uv ? 4 : 0) + (b->tap_bre ? 4 : 0);
            for (int i = 0; i < nn; ++i)
                cum[i + 1] = cum[i] + (upload_bound ? (int64_t)plans[i].n_passes * GF_NBINS * plans[i].T_out * 4 : (int64_t)plans[i].n_total * per_sample_down);
            const int64_t total = cum[nn];
            int np = upload_bound ? (int)std::min<int64_t>(16, (total + (22 << 20)) / (45 << 20))
                                  : (int)std::min<int64_t>(8, (total + (12 << 20)) / (24 << 20));
            if (np < 2 && total >= (8 << 20)) np = 2;
            // notes with post-FX / pitch dynamics run a dozen small kernels and uploads per part: two parts at most
            // (256 all-flag notes: 19.6 ms in 2 parts, 25.5 ms in 8)
            bool any_fx = false;
            for (int i = 0; i < nn && !any_fx; ++i) any_fx = note_needs_fx(plans[i]) || plans[i].pd != 0.0;
            if (any_fx) np = std::min(np, 2);
            np = std::max(1, std::min(np, nn));
            // Download-bound parts taper: the last part's download is the tail of the call (nothing is left to hide it), the
            // first part's kernels run while nothing is being downloaded yet.  One part more, sizes falling from 30 % to 6 %
            // (five parts: .30 .27 .22 .15 .06): 5.21 ms per call against 5.36 with four equal parts (c2, B200).
            std::vector<double> cutf;
            if (!upload_bound && !any_fx && np >= 3 && np + 1 <= nn && !getenv("GOOFER_HOST_EQUAL_PARTS")) {
                ++np;
                double sum = 0.0;
                std::vector<double> w(np);
                for (int k = 0; k < np; ++k) { w[k] = 1.0 - 0.8 * std::pow((double)k / (np - 1), 1.5); sum += w[k]; }
                double acc = 0.0;
                for (int k = 0; k + 1 < np; ++k) { acc += w[k] / sum; cutf.push_back(acc); }
            }
            for (int k = 1; k < np; ++k) {
                const int64_t target = cutf.empty() ? total * k / np : (int64_t)((double)total * cutf[k - 1]);
                int i = (int)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
                i = std::max(i, (ends.empty() ? 0 : ends.back()) + 1);
                if (i < nn) ends.push_back(i);
            }
        }
        ends.push_back(nn);
    }
    const int n_chunks = (int)ends.size();
    if (trace) h_uploads = now_ms();
    const size_t want = gf_workspace_bytes_planned(&db, plans, 0);
    if (trace) h_sized = now_ms();
    if ((rc = gf_hc_reserve(&g_hc.ws, &g_hc.ws_cap, want)) != GOOFER_OK) return rc;
    while ((int)g_hc.ev.size() < 2 * n_chunks + 1) {
        cudaEvent_t e;
        GF_CUDA(cudaEventCreate(&e));
        g_hc.ev.push_back(e);
    }
    // element ranges of every part's noise and output inside the concatenated buffers
    struct Range { int64_t plo, phi, nlo, nhi, olo, ohi; };
    std::vector<Range> rg(n_chunks);
    for (int c = 0; c < n_chunks; ++c) {
        const int i0 = c ? ends[c - 1] : 0, i1 = ends[c];
        Range r = {INT64_MAX, 0, INT64_MAX, 0, INT64_MAX, 0};
        for (int i = i0; i < i1; ++i) {
            const GfNotePlan &p = plans[i];
            for (int k = 0; k < p.n_passes; ++k) {
                if ((p.phi_rng_mask >> p.pass_kind[k]) & 1u) continue;           // drawn on the device: nothing to upload
                const int64_t o = p.phi_off[p.pass_kind[k]];
                if (o < 0 || !b->phi) { gf_set_error("note %d: phi slot %d not supplied", i, p.pass_kind[k]); return GOOFER_ERR_INVALID; }
                r.plo = std::min(r.plo, o); r.phi = std::max(r.phi, o + (int64_t)GF_NBINS * p.T_out);
            }
            const int need[4] = {p.f0_jitter, p.vol_jitter, p.vol_jitter, p.sj > 0.0};
            for (int k = 0; k < 4; ++k)
                if (need[k]) {
                    if (p.nrm_off[k] < 0) { gf_set_error("note %d: normal slot %d not supplied", i, k); return GOOFER_ERR_INVALID; }
                    r.nlo = std::min(r.nlo, p.nrm_off[k]); r.nhi = std::max(r.nhi, p.nrm_off[k] + (int64_t)p.n_total);
                }
            r.olo = std::min(r.olo, p.out_off); r.ohi = std::max(r.ohi, p.out_off + (int64_t)p.n_total);
        }
        if (r.phi == 0) r.plo = 0;                                           // no uploaded phases in this part
        if (r.plo < 0 || r.phi > b->phi_total || (r.nhi > 0 && (!b->normals || r.nlo < 0 || r.nhi > b->nrm_total)) || r.olo < 0 || r.ohi > b->out_total) {
            gf_set_error("part %d: noise / output offsets outside their buffers", c);
            return GOOFER_ERR_INVALID;
        }
        rg[c] = r;
    }
    std::vector<GfPart> parts(n_chunks);
    for (int c = 0; c < n_chunks; ++c) {
        parts[c].note_end = ends[c];
        parts[c].phi_ready = g_hc.ev[2 * c];
        parts[c].done = g_hc.ev[2 * c + 1];
    }
    // uploads: normals of every part (the preparation kernels of the whole batch need them), then the phases part by part
    for (int c = 0; c < n_chunks; ++c)
        if (rg[c].nhi > rg[c].nlo && (rc = h2d(db.normals + rg[c].nlo, b->normals + rg[c].nlo, sizeof(double) * (size_t)(rg[c].nhi - rg[c].nlo)))) return rc;
    if (b->f0_curves && b->f0_total > 0 && (rc = h2d(db.f0_curves, b->f0_curves, sizeof(float) * (size_t)b->f0_total))) return rc;
    if (!defer_sources) GF_CUDA(cudaEventRecord(g_hc.ev[2 * n_chunks], st_in));          // sources, bends, normals, f0 curves
    std::function<int()> deferred = [&]() -> int {
        const int r = upload_sources();
        if (r != GOOFER_OK) return r;
        GF_CUDA(cudaEventRecord(g_hc.ev[2 * n_chunks], st_in));      // recorded before anything waits for it (gf_render_wave)
        return GOOFER_OK;
    };
    for (int c = 0; c < n_chunks; ++c) {
        if (rg[c].phi > rg[c].plo && (rc = h2d(db.phi + rg[c].plo, b->phi + rg[c].plo, sizeof(float) * (size_t)(rg[c].phi - rg[c].plo)))) return rc;
        GF_CUDA(cudaEventRecord(g_hc.ev[2 * c], st_in));             // this part's phases
    }
    // one render of the whole batch: preparation kernels run once at full width; frame / peak / mix go part by part
    // GOOFER_HOST_SUBBATCHES=S (default 1): the batch is rendered as S sub-batches back to back on the compute stream, each a
    // run of whole parts, so that the results of sub-batch k cross PCIe while sub-batch k+1 is PREPARED (not only while its
    // frame / peak / mix tail runs).  One rank alone loses a little (smaller launches: 1,024 notes take 5.5 ms in one piece);
    // eight ranks sharing one host memory path gain a lot: in one piece all of them download at the same time, at the end
    // of the step, and the burst saturates the path (bench.py sets it for multi-rank runs).
    int n_sub = 1;
    { const char *e = getenv("GOOFER_HOST_SUBBATCHES"); if (e && atoi(e) > 1) n_sub = std::min(atoi(e), n_chunks); }
    int64_t launches = 0;
    int waves = 0;
    for (int k = 0; k < n_sub; ++k) {
        GfSubBatch sb;
        const int c0 = (int)((long)k * n_chunks / n_sub), c1 = (int)((long)(k + 1) * n_chunks / n_sub);
        sb.n0 = c0 ? ends[c0 - 1] : 0;
        sb.n1 = ends[c1 - 1];
        sb.first = k == 0;
        if ((rc = gf_render_batch_ex(&db, g_hc.ws, g_hc.ws_cap, st, parts.data(), n_chunks, &plans, g_hc.ev[2 * n_chunks],
                                     (defer_sources && k == 0) ? &deferred : nullptr, sb)) != GOOFER_OK) return rc;
    }
    launches = g_stats.kernel_launches;
    waves = g_stats.waves;
    if (defer_sources && deferred) { gf_set_error("internal: the deferred source uploads were never issued"); return GOOFER_ERR_CUDA; }
    int *status_host = (int *)gf_pin_take(256);              // render status word (overflowed pulse lists), read back with the results
    if (!status_host) { gf_set_error("cudaMallocHost failed for the status word"); return GOOFER_ERR_CUDA; }
    status_host[0] = 0; status_host[1] = -1;
    GF_CUDA(cudaMemcpyAsync(status_host, g_hc.ws, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (trace) h_enqueued = now_ms();
    for (int c = 0; c < n_chunks; ++c) {
        GF_CUDA(cudaStreamWaitEvent(st_out, g_hc.ev[2 * c + 1], 0));
        const int64_t olo = rg[c].olo;
        const size_t ob = sizeof(float) * (size_t)(rg[c].ohi - olo);
        if (b->out && (rc = d2h(b->out + olo, db.out + olo, ob))) return rc;
        if (b->out_pcm16 && (rc = d2h(b->out_pcm16 + olo, db.out_pcm16 + olo, ob / 2))) return rc;
        if (b->tap_harm && (rc = d2h(b->tap_harm + olo, db.tap_harm + olo, ob))) return rc;
        if (b->tap_uv && (rc = d2h(b->tap_uv + olo, db.tap_uv + olo, ob))) return rc;
        if (b->tap_bre && (rc = d2h(b->tap_bre + olo, db.tap_bre + olo, ob))) return rc;
        if (trace && c < 64) {
            if (!ev_d2h[c]) cudaEventCreate(&ev_d2h[c]);
            cudaEventRecord(ev_d2h[c], st_out);
        }
    }
    g_stats.kernel_launches = launches;
    g_stats.waves = waves;
    GF_CUDA(cudaStreamSynchronize(st_out));
    GF_CUDA(cudaStreamSynchronize(st));
    if (status_host[0]) {
        gf_set_error("note %d: %s list overflowed (mean f0 above sr / 8); the pulse train of that note is truncated", status_host[1],
                     (status_host[0] & 1) ? "pulse-onset" : "growl-event");
        return GOOFER_ERR_NOTE;
    }
    if (trace) {
        const double h_done = now_ms();
        float small = 0;
        cudaEventElapsedTime(&small, ev_t0, g_hc.ev[2 * n_chunks]);
        fprintf(stderr, "[host trace] host: first copy issued at %.3f ms, source uploads issued at %.3f ms, workspace sized at %.3f ms, everything "
                        "enqueued at %.3f ms, synchronised at %.3f ms; device: small inputs in at %.3f ms after the H2D stream started\n",
                h_first_copy - h0, h_uploads - h0, h_sized - h0, h_enqueued - h0, h_done - h0, small);
        for (const auto &m : g_htrace.marks) fprintf(stderr, "[host trace] host +%.3f ms: %s\n", m.second - h0, m.first);
        for (int c = 0; c < n_chunks; ++c) {
            float a = 0, bq = 0, dq = 0;
            cudaEventElapsedTime(&a, ev_t0, g_hc.ev[2 * c]);
            cudaEventElapsedTime(&bq, ev_t0, g_hc.ev[2 * c + 1]);
            if (c < 64) cudaEventElapsedTime(&dq, ev_t0, ev_d2h[c]);
            fprintf(stderr, "[host trace] part %d: phases in at %.3f ms, mixed at %.3f ms, downloaded at %.3f ms\n", c, a, bq, dq);
        }
    }
    return GOOFER_OK;
}
