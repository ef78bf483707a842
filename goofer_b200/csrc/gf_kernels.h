// gf_kernels.h -- host-side launchers of the kernels (one per .cu file), used by api.cu
#pragma once
#include "gf_device.cuh"

void gf_launch_src_env(const GfSourceDev *srcs, int n_src, int max_T, cudaStream_t st);
void gf_launch_tracks(const GfNotePlan *plans, const GfNoteDev *notes, const GfSourceDev *srcs, int n_notes, cudaStream_t st);
void gf_launch_env(const int2 *work, int n_work, const GfNotePlan *plans, const GfNoteDev *notes, const GfSourceDev *srcs, cudaStream_t st);
void gf_launch_mask(const GfNotePlan *plans, const GfNoteDev *notes, const GfSourceDev *srcs, int n_notes, int max_n, cudaStream_t st);
// Gaussian smoothing jobs: long kernels go to the overlap-save kernels (k_conv.cu), the rest to the direct sliding
// window (k_prep.cu); h_jobs is the host copy of d_jobs.  Both return the number of kernels launched.
int gf_launch_fir(const GfFirJob *h_jobs, const GfFirJob *d_jobs, int n_jobs, cudaStream_t st);
int gf_launch_fftconv(const GfFirJob *h_jobs, const GfFirJob *d_jobs, int n_jobs, cudaStream_t st);
void gf_launch_f0(const GfNotePlan *plans, const GfNoteDev *notes, const GfPassDev *passes, const GfSourceDev *srcs,
                  const float *bend, const double *normals, const float *f0_curves, int n_notes, int max_n, cudaStream_t st);
// flag_count: a zeroed device int the onset scan counts its flagged passes in (or NULL: the host picks the walk's shape)
int gf_launch_walk(const GfPassDev *passes, GfPassScal *scal, int n_pass, int max_n, int sr, cudaStream_t st, int *flag_count = nullptr);
void gf_launch_pulse(const GfPassDev *passes, const GfPassScal *scal, int n_pass, int max_n, cudaStream_t st);
void gf_launch_frame(const int4 *work, int n_work, const GfPassDev *passes, GfPassScal *scal, const GfNoteDev *notes,
                     const GfNotePlan *plans, cudaStream_t st);
void gf_launch_peak(const GfNotePlan *plans, const GfNoteDev *notes, const GfPassDev *passes, GfPassScal *scal, int pass0,
                    int n_pass, int max_n, bool any_simple, bool any_general, cudaStream_t st);
void gf_launch_mix(const GfNotePlan *plans, const GfNoteDev *notes, const GfPassDev *passes, const GfPassScal *scal,
                   int note0, int n_notes, int max_n, bool any_simple, bool any_general, cudaStream_t st);
struct GfPhiJob;
void gf_launch_phi(const GfPhiJob *jobs, int n_jobs, int max_T, cudaStream_t st);
void gf_launch_phi_fm(const GfPassDev *passes, int n_pass, int max_T, cudaStream_t st);
void gf_launch_onepole(const GfOnepoleJob *jobs, int n_jobs, cudaStream_t st);
size_t gf_frame_smem_bytes();

// stage-level kernels (k_stage.cu)
void gf_launch_stft(const float *x, int n_sig, int n, float2 *S, cudaStream_t st);
void gf_launch_istft(const float2 *S, int n_sig, int T, int length, float *y, cudaStream_t st);

int gf_tables_init(int sr);      // uploads d_tab for the current device (idempotent); 0 on success
