// host_api.cu -- goofer_render_batch_host: the same render with HOST buffers (what a ctypes / cgo
// caller that owns numpy arrays passes).  Copies sources, noise and bends in, renders, copies the
// output back; device and pinned staging buffers are cached per host thread.
#include <cstdlib>

struct GfHostCache {
    void *dev = nullptr;  size_t dev_cap = 0;
    void *ws = nullptr;   size_t ws_cap = 0;
    cudaStream_t st = nullptr;
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

extern "C" void goofer_host_release(void)
{
    if (g_hc.dev) cudaFree(g_hc.dev);
    if (g_hc.ws) cudaFree(g_hc.ws);
    if (g_hc.st) cudaStreamDestroy(g_hc.st);
    g_hc = GfHostCache();
}

extern "C" int goofer_render_batch_host(const GooferBatch *b)
{
    int rc = gf_validate(b);
    if (rc != GOOFER_OK) return rc;
    g_stats.h2d_bytes = 0; g_stats.d2h_bytes = 0;
    if (b->n_notes == 0) return GOOFER_OK;
    if (!b->out || !b->phi || !b->bend_cents) { gf_set_error("NULL out / phi / bend_cents"); return GOOFER_ERR_INVALID; }
    if (!g_hc.st) GF_CUDA(cudaStreamCreateWithFlags(&g_hc.st, cudaStreamNonBlocking));
    cudaStream_t st = g_hc.st;

    // ---- device image of every input array ----
    Bump sz{nullptr, 0, 0};
    auto carve = [&](Bump &bp, GooferBatch &db, std::vector<GooferSource> &ds) {
        for (int s = 0; s < b->n_sources; ++s) {
            const GooferSource &g = b->sources[s];
            GooferSource d = g;
            if (g.knots_log_f16) d.knots_log_f16 = bp.arr<uint16_t>((size_t)g.K * g.T);
            if (g.hz_knots) d.hz_knots = bp.arr<float>((size_t)g.K);
            if (g.env_dense) d.env_dense = bp.arr<float>((size_t)GF_NBINS * g.T);
            if (g.mask) d.mask = bp.arr<float>((size_t)g.N);
            for (int k = 0; k < 4; ++k) if (g.formants[k]) d.formants[k] = bp.arr<double>((size_t)g.formant_len[k]);
            ds[s] = d;
        }
        db.bend_cents = bp.arr<float>((size_t)b->bend_total);
        db.phi = bp.arr<float>((size_t)b->phi_total);
        db.normals = b->normals ? bp.arr<double>((size_t)b->nrm_total) : nullptr;
        db.out = bp.arr<float>((size_t)b->out_total);
        db.tap_harm = b->tap_harm ? bp.arr<float>((size_t)b->out_total) : nullptr;
        db.tap_uv = b->tap_uv ? bp.arr<float>((size_t)b->out_total) : nullptr;
        db.tap_bre = b->tap_bre ? bp.arr<float>((size_t)b->out_total) : nullptr;
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
        GF_CUDA(cudaMemcpyAsync(const_cast<void *>(dst), src, bytes, cudaMemcpyHostToDevice, st));
        g_stats.h2d_bytes += (int64_t)bytes;
        return GOOFER_OK;
    };
    for (int s = 0; s < b->n_sources; ++s) {
        const GooferSource &g = b->sources[s];
        const GooferSource &d = ds[s];
        if (g.knots_log_f16 && (rc = h2d(d.knots_log_f16, g.knots_log_f16, sizeof(uint16_t) * (size_t)g.K * g.T))) return rc;
        if (g.hz_knots && (rc = h2d(d.hz_knots, g.hz_knots, sizeof(float) * (size_t)g.K))) return rc;
        if (g.env_dense && (rc = h2d(d.env_dense, g.env_dense, sizeof(float) * (size_t)GF_NBINS * g.T))) return rc;
        if (g.mask && (rc = h2d(d.mask, g.mask, sizeof(float) * (size_t)g.N))) return rc;
        for (int k = 0; k < 4; ++k)
            if (g.formants[k] && (rc = h2d(d.formants[k], g.formants[k], sizeof(double) * (size_t)g.formant_len[k]))) return rc;
    }
    if ((rc = h2d(db.bend_cents, b->bend_cents, sizeof(float) * (size_t)b->bend_total))) return rc;
    if ((rc = h2d(db.phi, b->phi, sizeof(float) * (size_t)b->phi_total))) return rc;
    if (b->normals && (rc = h2d(db.normals, b->normals, sizeof(double) * (size_t)b->nrm_total))) return rc;

    const size_t want = goofer_workspace_bytes(&db, 0);
    if (want == 0) return GOOFER_ERR_NOTE;
    if ((rc = gf_hc_reserve(&g_hc.ws, &g_hc.ws_cap, want)) != GOOFER_OK) return rc;
    if ((rc = goofer_render_batch(&db, g_hc.ws, g_hc.ws_cap, st)) != GOOFER_OK) return rc;

    auto d2h = [&](void *dst, const void *src, size_t bytes) -> int {
        GF_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
        g_stats.d2h_bytes += (int64_t)bytes;
        return GOOFER_OK;
    };
    if ((rc = d2h(b->out, db.out, sizeof(float) * (size_t)b->out_total))) return rc;
    if (b->tap_harm && (rc = d2h(b->tap_harm, db.tap_harm, sizeof(float) * (size_t)b->out_total))) return rc;
    if (b->tap_uv && (rc = d2h(b->tap_uv, db.tap_uv, sizeof(float) * (size_t)b->out_total))) return rc;
    if (b->tap_bre && (rc = d2h(b->tap_bre, db.tap_bre, sizeof(float) * (size_t)b->out_total))) return rc;
    GF_CUDA(cudaStreamSynchronize(st));
    return GOOFER_OK;
}
