/*
 * adanerf_b200 -- C ABI of the B200-native AdaNeRF inference renderer (libadanerf_b200.so).
 *
 * One data-parallel hot path, hand-written for sm_100a:
 *   rays -> SpherePosDir features -> sampling MLP (tcgen05, bf16x3 split precision)
 *        -> threshold / top-K / scan compaction -> positional encoding
 *        -> shading MLP (tcgen05, bf16) -> per-ray transmittance scan + alpha composite.
 *
 * Each entry point cites the reference interface (relative to thomasneff/AdaNeRF) it replaces.
 * Conventions: plain pointers and sizes only; integer status codes (0 = ok), never exceptions;
 * "d_" pointers are device memory owned by the caller, "h_" pointers are host memory; the context
 * owns packed weights and scratch; work is stream ordered (`stream` is a cudaStream_t passed as
 * void*, NULL = legacy default stream); no host synchronisation inside unless stated; one context
 * per device; a context is not thread safe.  There is NO CPU fallback: every call fails with
 * ADN_ERR_CUDA / ADN_ERR_NO_DEVICE when no sm_100 device is usable.
 */
#ifndef ADANERF_B200_H
#define ADANERF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct adn_ctx adn_ctx;
typedef int adn_status;

enum {
  ADN_OK = 0,
  ADN_ERR_INVALID = 1,      /* bad argument / unsupported shape */
  ADN_ERR_CUDA = 2,         /* CUDA runtime error (see adn_last_error) */
  ADN_ERR_NO_DEVICE = 3,    /* no CUDA device / not compute capability 10.x */
  ADN_ERR_NO_WEIGHTS = 4,   /* render called before both nets were set */
  ADN_ERR_IO = 5,           /* export directory / file problems */
  ADN_ERR_KERNEL = 6        /* device-side watchdog tripped (mbarrier timeout) */
};

/* Scene constants.  Source: the export's dataset_info.txt written by src/export.py:47-54
 * (view_cell_center, view_cell_size, depth_range = WARPED range, fov, max_depth) and the feature
 * set constructors src/features.py:286-287 (zNear/zFar), :338-339 (posEncArgs "10-4"). */
typedef struct adn_scene {
  float view_cell_center[3];
  float view_cell_size[3];
  float depth_range[2];
  float max_depth;
  float fov;            /* radians; focal = 0.5*W/tan(fov/2)  (src/datasets.py:181-182) */
  float z_near, z_far;  /* 0.001, 1.0 */
  int32_t n_freq_pos;   /* 10: shading-net encoding, posEncArgs[1] = "10-4" */
  int32_t n_freq_dir;   /* 4 */
  /* Sampling-net encoding, posEncArgs[0]: 0/0 = same as above ("10-4", 90 features) or 2/2 ("2-2", 30 features,
   * configs/fine_training_ndc.ini:8). */
  int32_t n_freq_pos0, n_freq_dir0;
  /* NDC / LLFF variant (configs/fine_training_ndc.ini: useNDC, FromClassifiedDepthAdaptiveNoDepthRange,
   * rayMarchNormalization[1] = None): rays go through ndc_rays(H, W, focal, near = 1)
   * (src/nerf_raymarch_common.py:71-88, src/features.py:429-431), sample depths are the cell centres in [0,1]
   * (no depth-range warp), positions are not normalised, NeRFOutputDepth is the depth map itself.
   * ndc_w / ndc_h: the dataset's image size (features.py:350-351); ndc_focal <= 0 -> 0.5 * ndc_w / tan(fov / 2). */
  int32_t use_ndc;
  int32_t ndc_w, ndc_h;
  float ndc_focal;
} adn_scene;

/* One named parameter tensor, fp32, row-major [rows, cols] ([out, in] for weights, [out, 1] or
 * [1, out] for biases), with the reference's state_dict names: `layers.{i}.weight|bias` for the
 * sampling net (src/models.py:71-76), `pts_linears.{i}`, `views_linears.0`, `feature_linear`,
 * `alpha_linear`, `rgb_linear` for the shading net (src/models.py:226-244). */
typedef struct adn_tensor_desc {
  const char* name;
  const float* data;    /* host pointer */
  int64_t rows, cols;
} adn_tensor_desc;

typedef struct adn_stats {
  int64_t n_rays;          /* rays of the last render call */
  int64_t n_samples;       /* M = surviving samples of the last render call (valid after a sync) */
  float ms_stage[6];       /* device ms of stages 0..5 of the last *profiled* render (adn_set_option "profile"=1) */
  int64_t kernel_launches; /* kernels launched by this context so far */
} adn_stats;

/* ---- lifetime --------------------------------------------------------------------------- */
/* Replaces ImageGenerator::ImageGenerator + FeatureSet::create (imagegenerator.cpp:84-201,
 * featureset.cpp:67-137) and TrainConfig.initialize's feature setup (src/train_data.py:63-239). */
adn_status adn_create(adn_ctx** out, const adn_scene* scene, int device);
void adn_destroy(adn_ctx* ctx);
const char* adn_strerror(adn_status s);
const char* adn_last_error(const adn_ctx* ctx);
const char* adn_version(void);

/* net_id 0 = sampling net (BaseNet, src/models.py:18-195), 1 = shading net (NeRF, :199-277).
 * Packs the fp32 parameters into the kernels' layout (bf16 hi/lo split, swizzled K-major tiles).
 * Replaces load_state_dict / ImageGenerator::initEngine's ONNX->TensorRT build. */
adn_status adn_set_weights(adn_ctx* ctx, int net_id, const adn_tensor_desc* tensors, int n_tensors);

/* Loads {config.ini, dataset_info.txt, model0.onnx, model1.onnx} as written by src/export.py:28-93
 * (the directory the C++ viewer takes with -mp, adanerf_real_time_viewer/README.md:38-43).
 * Creates the context with the scene from dataset_info.txt; threshold / K from config.ini are
 * returned through thr_out / k_out (may be NULL). */
adn_status adn_create_from_export_dir(adn_ctx** out, const char* dir, int device, float* thr_out, int* k_out);

/* Host-only (no GPU needed): parses the export directory like adn_create_from_export_dir and returns the
 * scene, threshold, K and the number of fp32 tensors found in model0.onnx / model1.onnx.  Replaces
 * Config::load + Config::loadDatasetInfo (adanerf_real_time_viewer/src/config.cpp:270-344). */
adn_status adn_probe_export_dir(const char* dir, adn_scene* scene_out, float* thr_out, int* k_out, int* n_tensors_out /*[2]*/);

/* name: "chunk_rays" (rays per internal batch, 0 = auto), "profile" (0/1 per-stage event timing),
 * "mlp0_terms" (3 = bf16x3 split precision [default], 1 = plain bf16; parity experiments only),
 * "cta_group" (2 = CTA-pair MMAs [default], 1 = single-CTA MMAs; A/B runs),
 * "fuse_encoder" (1 = positional encoding of the samples inside the shading kernel: no [M,90]-sized tile buffer, ~5 %
 *   slower; 0 = separate kernel [default]),
 * "trace" (debug: net id whose MLP kernel records an in-kernel timeline, -1 = off; profiles/trace_mlp.py). */
adn_status adn_set_option(adn_ctx* ctx, const char* name, int64_t value);
adn_status adn_get_stats(adn_ctx* ctx, adn_stats* out);   /* synchronises the context's stream */

/* ---- the hot path ----------------------------------------------------------------------- */
/* render(rays, sampling_net, shading_net, adaptiveSamplingThreshold): one call = what
 * TrainConfig.inference(batch, gradient=False, is_inference=True) computes (src/train_data.py:278-299):
 *   d_dirs  [N,3] camera-space pixel directions (DatasetKeyConstants.ray_directions_samples)
 *   pose[3], rot[9] (row-major 3x3) HOST pointers (ImagePose / ImageRotation)
 *   thr = adaptiveSamplingThreshold (0 = dense K samples, K must then be 128), K = numRaymarchSamples[1]
 *   d_rgb [N,3] fp32 (outs[-1][:, :3]); d_nsamples [N] int32 or NULL (AdaptiveSamplePositions * K);
 *   d_oracle_weights [N,128] fp32 or NULL (dicts[1]["OracleWeights"], i.e. raw sampling-net output). */
adn_status adn_render_rays(adn_ctx* ctx, const float* pose, const float* rot, const float* d_dirs, int64_t n_rays,
                           float thr, int K, float* d_rgb, int32_t* d_nsamples, float* d_oracle_weights,
                           void* stream);

/* Auxiliary outputs of the same call: the other return values of adaptive_raw2outputs
 * (src/nerf_raymarch_common.py:137-144) and the tensors RayMarchFromPoses.postprocess stores in the inference dict
 * (src/features.py:536-577), read by plots.render_all_imgs / the depth export (src/plots.py:272-306).
 * Every pointer is a device pointer and may be NULL; [N,K] tensors are padded like the reference's. */
typedef struct adn_aux_outputs {
  float* d_weights;    /* [N,K] "NeRFWeightsOutput": alpha * transmittance, 0 in unused slots */
  float* d_alpha;      /* [N,K] "NeRFAlphaOutput": sigmoid(raw alpha) * sampling-net value, 0 in unused slots */
  float* d_z_vals;     /* [N,K] "NeRFInputFeatureZVals": world depth of the samples, NaN in unused slots */
  float* d_depth_map;  /* [N] sum_k w z (world depth) */
  float* d_acc_map;    /* [N] sum_k w */
  float* d_disp_map;   /* [N] 1 / max(1e-10, depth_map / acc_map) */
  float* d_depth_est;  /* [N] "NeRFOutputDepth": LogTransform.from_world(depth_map, depth_range) */
} adn_aux_outputs;
adn_status adn_render_rays_aux(adn_ctx* ctx, const float* pose, const float* rot, const float* d_dirs, int64_t n_rays,
                               float thr, int K, float* d_rgb, int32_t* d_nsamples, float* d_oracle_weights,
                               const adn_aux_outputs* aux, void* stream);

/* Same, generating the pinhole rays of image rows [row0, row0+rows) of a WxH frame on the device
 * (src/util/raygeneration.py:10-26; ray id = y*W + x).  Replaces Camera::UpdateFeaturesBatch +
 * ImageGenerator::inference's batch loop (camera.cpp:160-201, imagegenerator.cpp:247-478).
 * d_rgb [rows*W,3].  This is the multi-GPU tile entry: each rank renders its row band. */
adn_status adn_render_camera(adn_ctx* ctx, const float* pose, const float* rot, int W, int H, int row0, int rows,
                             float thr, int K, float* d_rgb, int32_t* d_nsamples, void* stream);

/* Viewer output format: RGBA8 (adaptive_cuda_kernels.cu:846-851 writes uchar4 through surf2Dwrite);
 * d_rgba8 is a linear [rows*W] uchar4 buffer the caller copies/maps to its GL resource. */
adn_status adn_render_camera_rgba8(adn_ctx* ctx, const float* pose, const float* rot, int W, int H, int row0, int rows,
                                   float thr, int K, uint8_t* d_rgba8, void* stream);

/* The viewer's own output target: a cudaSurfaceObject_t bound to the GL-registered cudaArray of the current render buffer
 * (adanerf_real_time_viewer/src/interoprenderbuffer.cpp:53-83); uchar4 pixels written with surf2Dwrite at (x, row0 + y)
 * exactly like adaptive_cuda_kernels.cu:846-851.  `surface` is the cudaSurfaceObject_t value (an unsigned 64-bit handle;
 * the header stays free of CUDA types). */
adn_status adn_render_camera_surface(adn_ctx* ctx, const float* pose, const float* rot, int W, int H, int row0, int rows,
                                     float thr, int K, unsigned long long surface, void* stream);

/* End-to-end variants with HOST buffers: H2D of the inputs and D2H of the results happen inside the
 * call and the call returns after the results are on the host.  Buffers the caller registered (below) or allocated
 * page-locked itself are the source / target of the DMA; anything else goes through pinned staging owned by the
 * context (one extra host copy each way). */
adn_status adn_render_rays_host(adn_ctx* ctx, const float* pose, const float* rot, const float* h_dirs, int64_t n_rays,
                                float thr, int K, float* h_rgb, int32_t* h_nsamples);
adn_status adn_render_camera_host(adn_ctx* ctx, const float* pose, const float* rot, int W, int H, int row0, int rows,
                                  float thr, int K, float* h_rgb, int32_t* h_nsamples);

/* Page-locks [p, p + bytes) in place (cudaHostRegister) for the *_host entry points.  The caller owns the memory and must
 * keep it allocated until adn_unregister_host_buffer / adn_destroy: the library never registers memory on its own (a
 * buffer freed and re-allocated at the same address would keep a stale registration).  The viewer's equivalent is the
 * GL-registered cudaArray of InteropRenderbuffer (adanerf_real_time_viewer/src/interoprenderbuffer.cpp:53-83). */
adn_status adn_register_host_buffer(adn_ctx* ctx, const void* p, size_t bytes);
adn_status adn_unregister_host_buffer(adn_ctx* ctx, const void* p);

/* Input / output width of a network set with adn_set_weights (sampling net: n_out is 128 in a render, the stage-level
 * adn_mlp0_forward also accepts test networks with 256 outputs and writes [N, n_out]). */
adn_status adn_net_dims(adn_ctx* ctx, int net_id, int* n_in, int* n_out);

/* ---- stage-level entry points (parity tests drive each kernel in isolation) ------------- */
/* stage 0: SpherePosDir.batch (src/features.py:845-899). d_x0 [N,90] fp32 (dir block first), d_ray_o/d [N,3]. */
adn_status adn_stage0_features(adn_ctx* ctx, const float* pose, const float* rot, const float* d_dirs, int64_t n_rays,
                               float* d_x0, float* d_ray_o, float* d_ray_d, void* stream);
/* stage 0a: generate_ray_directions (src/util/raygeneration.py:10-26). d_dirs [rows*W,3]. */
adn_status adn_generate_ray_directions(adn_ctx* ctx, int W, int H, int row0, int rows, float* d_dirs, void* stream);
/* stage 1: BaseNet.forward (src/models.py:183-195). d_x0 [N,n_in] fp32 -> d_raw0 [N,n_out] fp32. */
adn_status adn_mlp0_forward(adn_ctx* ctx, const float* d_x0, int64_t n_rays, float* d_raw0, void* stream);
/* stage 2: FromClassifiedDepthAdaptive.generate (src/nerf_raymarch_common.py:699-757) + the mask
 * compaction of RayMarchFromPoses.batch (src/features.py:445-446,481-484).  Packed order is ray-major,
 * depth-ascending (= torch boolean-mask order).  Outputs: d_count/d_offset [N] int32 (exclusive scan),
 * d_cell/d_ray [cap] int32, d_z (world depth) / d_zp [cap] fp32, d_total (int64, device). cap >= N*K. */
adn_status adn_stage2_sample(adn_ctx* ctx, const float* d_raw0, int64_t n_rays, float thr, int K,
                             int32_t* d_count, int32_t* d_offset, int32_t* d_cell, int32_t* d_ray,
                             float* d_z, float* d_zp, int64_t* d_total, void* stream);
/* stage 3: RayMarchFromPoses.batch encode (src/features.py:458-479). d_x1 [M,90] fp32 (pos block first). */
adn_status adn_stage3_encode(adn_ctx* ctx, const float* d_ray_o, const float* d_ray_d, const int32_t* d_ray,
                             const float* d_z, int64_t n_samples, float* d_x1, void* stream);
/* stage 4: NeRF.forward (src/models.py:254-277). d_x1 [M,90] fp32 -> d_raw1 [M,4] fp32 = [rgb, alpha]. */
adn_status adn_mlp1_forward(adn_ctx* ctx, const float* d_x1, int64_t n_samples, float* d_raw1, void* stream);
/* stage 5: adaptive_raw2outputs (src/nerf_raymarch_common.py:91-144, accumulation_mult "alpha").
 * d_weights / d_depth_map may be NULL; d_weights is [N,K] zero padded like the reference's. */
adn_status adn_stage5_composite(adn_ctx* ctx, const float* d_raw1, const float* d_zp, const float* d_z,
                                const int32_t* d_offset, const int32_t* d_count, int64_t n_rays, int K,
                                float* d_rgb, float* d_weights, float* d_depth_map, void* stream);

/* ---- evaluation metric on the device ------------------------------------------------------ */
/* calculate_mse / calculate_psnr (src/evaluate.py:49-54) of two device images of n_values floats each:
 * mse = sum((a - b)^2) / n_values (double accumulation, deterministic), psnr = 10 log10(1 / mse).
 * clamp01 != 0 clips d_image to [0,1] first (src/evaluate.py:257-258).  Synchronises the stream; results on the host. */
adn_status adn_image_metrics(adn_ctx* ctx, const float* d_image, const float* d_reference, int64_t n_values, int clamp01,
                             double* mse_out, double* psnr_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ADANERF_B200_H */
