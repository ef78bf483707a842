// k_fx.cu -- growl (sg), pitch dynamics (pd) and the sequential post-FX of GooferResampler.resample
// (su / sj layers, vocal-fry high-pass blend, sd tremolo, st tension).   SillySampler.py:1038-1140
#include "gf_device.cuh"
#include "gf_maps.cuh"

int gf_growl(const WaveHost &wh, const GfNotePlan *, const GfNoteDev *, const GfPassDev *, GfPassScal *, int, cudaStream_t, int64_t *)
{
    for (const GfNotePlan &p : wh.plans)
        if (p.add_subharm) { gf_set_error("sg (growl) is not implemented yet"); return GOOFER_ERR_INVALID; }
    return GOOFER_OK;
}

int gf_pitch_dyn(const WaveHost &wh, const GfNotePlan *, const GfNoteDev *, const float *, Bump &, int, cudaStream_t, int64_t *)
{
    for (const GfNotePlan &p : wh.plans)
        if (p.pd != 0.0) { gf_set_error("pd (pitch dynamics) is not implemented yet"); return GOOFER_ERR_INVALID; }
    return GOOFER_OK;
}

int gf_post_fx(const WaveHost &wh, const GfNotePlan *, const GfNoteDev *, const GfPassDev *, GfPassScal *, Bump &, int, int,
               cudaStream_t, int64_t *)
{
    for (const GfNotePlan &p : wh.plans)
        if (p.su > 0.0 || p.sj > 0.0 || p.fry_mask_on || p.sd > 0 || p.tension != 0.0) {
            gf_set_error("su / sj / vf / sd / st post-FX are not implemented yet");
            return GOOFER_ERR_INVALID;
        }
    return GOOFER_OK;
}
