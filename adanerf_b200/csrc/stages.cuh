// SIMT stages of the AdaNeRF hot path (HBM-bound byte/index work): ray generation, SpherePosDir
// features (stage 0), threshold / top-K / scan compaction (stage 2), positional encoding (stage 3),
// per-ray transmittance scan + composite (stage 5).  Launchers only; kernels in stages.cu.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace adn {

struct SceneDev {
  float c[3];          // view_cell_center (fp32, features.py:759 / :345)
  float r2;            // float(view_cell_radius**2) (features.py:761,786)
  float sqrt_max_depth;  // float(math.sqrt(max_depth)) (nerf_raymarch_common.py:229)
  int n_freq_pos, n_freq_dir;
  int n_freq_pos0, n_freq_dir0;   // sampling-net encoding (10/4 or 2/2)
  int ndc;                        // NDC variant: ndc_rays + no position normalisation (features.py:429-431)
  float ndc_cw, ndc_ch;           // float(-1 / (W / (2 focal))), float(-1 / (H / (2 focal)))
};

struct CameraRays {  // src/util/raygeneration.py:10-26 in double precision
  double start_x, start_y, focal, x_pp, y_pp;
  int W, H, row0;
};

struct PoseDev {
  float pose[3];
  float rot[9];
};

// Packed-tile destinations (nullptr = skip): MLP0 input tiles (hi/lo) and MLP1 input tiles.
cudaError_t launch_gen_dirs(const CameraRays& cam, long long n_rays, float* d_dirs, cudaStream_t s);
cudaError_t launch_stage0(const SceneDev& sc, const PoseDev& pd, const float* d_dirs, const CameraRays* cam,
                          long long n_rays, float* d_x0, float* d_ray_o, float* d_ray_d, uint8_t* d_tiles0,
                          cudaStream_t s);

// Stage 2.  tile_state: [n_ctas + 2] uint64 scratch zeroed by the launcher (memsetAsync).
size_t stage2_scratch_bytes(long long n_rays);
// Host-side bookkeeping of a stage-2 scratch buffer (owned by whoever owns the buffer): launch epoch and ticket base, so
// the tile states / ticket counter need no per-launch memset.
struct Stage2Sync {
  void* scratch = nullptr;
  size_t cleared_bytes = 0;
  uint32_t epoch = 0, ticket_base = 0;
};
cudaError_t launch_stage2(const float* d_raw0, long long n_rays, float thr, int K, const float* d_zlut, int32_t* d_count,
                          int32_t* d_offset, int32_t* d_cell, int32_t* d_ray, float* d_z, float* d_zp, long long* d_total,
                          void* d_scratch, Stage2Sync* sync, cudaStream_t s);
// Dense (thr == 0): count = K, offset = ray*K, total = N*K; no index arrays are materialised.
cudaError_t launch_stage2_dense(long long n_rays, int K, int32_t* d_count, int32_t* d_offset, long long* d_total,
                                cudaStream_t s);

// Stage 3.  Adaptive: sample s -> (d_ray[s], d_z[s]).  Dense (d_ray == nullptr): ray = s / K,
// z = d_zlut_dense[s % K].  n_samples read from d_total when non-null.
cudaError_t launch_stage3(const SceneDev& sc, const float* d_ray_o, const float* d_ray_d, const int32_t* d_ray,
                          const float* d_z, const float* d_zlut_dense, int K, long long n_samples, const long long* d_total,
                          float* d_x1, uint8_t* d_tiles1, cudaStream_t s);

// Optional per-ray / per-slot outputs of the composite (adaptive_raw2outputs' other return values and the tensors
// RayMarchFromPoses.postprocess puts into the inference dict, src/features.py:536-577).  Any pointer may be null.
struct Stage5Aux {
  float* weights = nullptr;     // [N,K] zero padded (NeRFWeightsOutput)
  float* alpha = nullptr;       // [N,K] zero padded, sigmoid(a) * zp (NeRFAlphaOutput)
  float* z_vals = nullptr;      // [N,K] world depth, NaN padded (NeRFInputFeatureZVals)
  float* depth_map = nullptr;   // [N] sum w z
  float* acc_map = nullptr;     // [N] sum w
  float* disp_map = nullptr;    // [N] 1 / max(1e-10, depth_map / acc_map)
  float* depth_est = nullptr;   // [N] LogTransform.from_world(depth_map, depth_range) (NeRFOutputDepth)
  int linear_depth = 0;         // NDC: depth_est = depth_map (features.py:573-574)
  float dr_min = 0.0f;          // depth_range[0]
  float log_range = 1.0f;       // float(log(depth_range[1] - depth_range[0] + 1))
  bool any() const { return weights || alpha || z_vals || depth_map || acc_map || disp_map || depth_est; }
};

// Stage 5.  zp: adaptive -> packed [M]; dense (dense_zp_stride > 0) -> raw0 [N, stride].
cudaError_t launch_stage5(const float* d_raw1, const float* d_zp, const float* d_z, const float* d_zlut_dense,
                          const int32_t* d_offset, const int32_t* d_count, long long n_rays, int K, int dense,
                          float* d_rgb, uint8_t* d_rgba8, const Stage5Aux& aux, cudaStream_t s);

// Linear RGBA8 pixels [rows * W] -> surf2Dwrite(uchar4, surface, 4 x, row0 + y) (adaptive_cuda_kernels.cu:846-851).
cudaError_t launch_rgba_to_surface(const uint8_t* d_rgba8, int W, int row0, int rows, unsigned long long surface, cudaStream_t s);

// Image metric (src/evaluate.py:49-54): sum over all values of (a - b)^2 in double, deterministic two-stage
// reduction.  d_partials: kMetricBlocks doubles of scratch; d_sum receives the total.  clamp01: clip `a` to [0,1] first
// (what the reference does to an image before it is written / compared as 8-bit, src/evaluate.py:257-258).
constexpr int kMetricBlocks = 592;
cudaError_t launch_image_sqdiff(const float* d_a, const float* d_b, long long n_values, int clamp01, double* d_partials,
                                double* d_sum, cudaStream_t s);

}  // namespace adn
