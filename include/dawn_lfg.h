/* dawn_lfg.h — C-ABI of the B200-native LFG flow decoder: the stage of DAWN that turns the sampled latent
 * flow / occlusion maps into video frames.
 *
 * Reference seam: `Generator.compute_fea` and `Generator.forward_with_flow` (LFG/modules/generator.py:132-171), called by
 * `FlowDiffusion.sample_one_video` once per clip and once per FRAME respectively, batch 1, in a Python loop
 * (DM_3/modules/video_flow_diffusion_model_multiGPU_v0_crema_vgg_floss_plus_faceemb_flow_fast_init_cond_test.py:327, 375-383).
 * Here the source-image encoder (first + down blocks, generator.py:140-146) runs once per clip and all frames are decoded as
 * one batch.  Plain pointers and sizes; one handle per GPU, not thread-safe, stream-ordered, no host synchronisation inside
 * set_source / decode.  All tensors fp32.  Return: 0 ok, -1 bad argument / order / unsupported configuration, -2 CUDA error;
 * text through dawn_last_error, declared in include/dawn_unet.h.
 */
#ifndef DAWN_LFG_H_
#define DAWN_LFG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dawn_lfg dawn_lfg;

/* generator_params of config/hdtf128.yaml:82-93 (Generator.__init__, generator.py:25-57) */
typedef struct {
  int num_channels;          /* 3 */
  int block_expansion;       /* 64 */
  int max_features;          /* 512 */
  int num_down_blocks;       /* 2 */
  int num_bottleneck_blocks; /* 6 */
  int skips;                 /* 1 */
} dawn_lfg_cfg;

int dawn_lfg_create(const dawn_lfg_cfg* cfg, dawn_lfg** out);
void dawn_lfg_destroy(dawn_lfg* h);

/* replaces generator.load_state_dict(checkpoint['generator']) (FD:120): `name` is the reference state_dict key
 * (first.conv.weight, bottleneck.r0.norm1.running_var, ...); entries under pixelwise_flow_predictor.* and
 * *.num_batches_tracked are accepted and ignored (never read by the decode path).  host: fp32 values, row-major in `shape`. */
int dawn_lfg_set_param(dawn_lfg* h, const char* name, const float* host, const int64_t* shape, int ndim);
/* fold the eval-mode BatchNorms into the convolutions where a conv precedes them, repack and upload */
int dawn_lfg_commit_params(dawn_lfg* h);

/* frames per decode call, image size (H, W: multiples of 2^num_down_blocks * 16 / 8 so every level tiles), flow size (h, w) */
int dawn_lfg_set_geometry(dawn_lfg* h, int frames, int H, int W, int flow_h, int flow_w);

/* per clip: source image (3, H, W) in [0, 1] on the device -> skip features of every level (generator.py:140-146) */
int dawn_lfg_set_source(dawn_lfg* h, const float* source, void* stream);
/* replaces Generator.compute_fea (generator.py:132-136): fea (C_bottleneck, H/2^n, W/2^n) of the current source, device */
int dawn_lfg_get_fea(dawn_lfg* h, float* fea, void* stream);

/* replaces the per-frame loop over Generator.forward_with_flow (FD:375-383) for `frames` frames at once:
 *   flow (frames, h, w, 2) sampling grid in [-1, 1] (x, y), occ (frames, 1, h, w)  ->  prediction (frames, 3, H, W),
 *   deformed (frames, 3, H, W) or NULL.  All device pointers. */
int dawn_lfg_decode(dawn_lfg* h, const float* flow, const float* occ, float* prediction, float* deformed, void* stream);
/* same from the sampler's output: sample (3, frames, h, w) = [grid_x, grid_y, conf], occlusion = (conf + 1) / 2 (FD:366-369) */
int dawn_lfg_decode_sample(dawn_lfg* h, const float* sample, float* prediction, float* deformed, void* stream);

/* debugging / sub-module parity: copy of an internal activation of the last decode as (C, frames, Hl, Wl):
 * "bottleneck", "up0", "up1" (names as in oracle/lfg_oracle.py taps).  Writes C, Hl, Wl; dst may be NULL to query the shape. */
int dawn_lfg_read_tap(dawn_lfg* h, const char* name, float* dst, int* C, int* Hl, int* Wl, void* stream);

/* one layer of Face_loc_Encoder (FD:39-50), the per-clip face-box embedding fed to the UNet next to the source features:
 * out (Co, ceil(H/2), ceil(W/2)) = relu(conv3x3 stride 2 pad 1 of x (Ci, H, W) + bias).  All device pointers; weight (Co, Ci, 3, 3). */
int dawn_conv3x3_s2_relu(const float* x, int Ci, int H, int W, const float* weight, const float* bias, int Co, float* out, void* stream);

int64_t dawn_lfg_last_launch_count(dawn_lfg* h);
int64_t dawn_lfg_workspace_bytes(dawn_lfg* h);

#ifdef __cplusplus
}
#endif
#endif /* DAWN_LFG_H_ */
