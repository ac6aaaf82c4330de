// C++ host side for the viewer: the public surface of the reference's `ImageGenerator`
// (adanerf_real_time_viewer/include/imagegenerator.h:61-62 `inference(...)`, :58 `load(...)`) and the pieces
// of `Config` / `Camera` it needs (config.h:31-59, camera.h), re-implemented over the C ABI
// (include/adanerf_b200.h).  No TensorRT, no OpenGL: the frame is produced into a linear RGBA8 / fp32 buffer
// that the caller maps to its display resource (the reference writes the same uchar4 through surf2Dwrite,
// adaptive_cuda_kernels.cu:846-851).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../../include/adanerf_b200.h"

// The viewer's own types that appear in ImageGenerator::inference's parameter list (include/featureset.h:25,
// include/encoding.h:10).  The replacement never looks inside them -- the features and encodings of the hot path live in
// the CUDA kernels behind the C ABI -- so forward declarations are all it needs; in the viewer build they resolve to the
// viewer's classes and NeuralRenderer::render (neuralrenderer.cpp:158-160) compiles unchanged.
class FeatureSet;
class Encoding;

namespace adn_host {

// Counterpart of Config (config.cpp:270-344): what the export directory says.
struct Config {
  adn_scene scene{};
  float adaptiveSamplingThreshold = 0.f;
  int numRaymarchSamples = 0;
  std::string model_dir;
  bool load(const std::string& dir);   // parses config.ini + dataset_info.txt + model{0,1}.onnx headers
};

// Counterpart of Camera (camera.cpp:26-94): position + yaw/pitch fly camera producing the 3x3 rotation
// the feature kernels consume (row-major, columns = right / up / -forward in world space).
struct Camera {
  float pos[3] = {0, 0, 0};
  float yaw = 0.f, pitch = 0.f;
  int width = 800, height = 800;
  void rotation(float rot[9]) const;
};

class ImageGenerator {
 public:
  ImageGenerator() = default;
  ~ImageGenerator();
  ImageGenerator(const ImageGenerator&) = delete;
  ImageGenerator& operator=(const ImageGenerator&) = delete;

  // ImageGenerator::load (imagegenerator.cpp:203-245): builds the device context and packs both networks.
  bool load(const Config& config, int device = 0);

  // ImageGenerator::inference (imagegenerator.cpp:247-478): renders the camera's frame.  `batch_size` is
  // the rays-per-batch knob of the viewer (-bs); `num_samples` = K.  d_rgba8: device buffer [W*H] uchar4.
  bool inference(const Camera& camera, uint8_t* d_rgba8, int batch_size, int num_samples, void* stream = nullptr);
  // The reference's parameter list, verbatim (include/imagegenerator.h:61-62; called from neuralrenderer.cpp:158-160):
  // `output_surf` is the cudaSurfaceObject_t of the GL-registered render buffer (buffermanager.h:34) and receives uchar4
  // pixels through surf2Dwrite (adaptive_cuda_kernels.cu:846-851); the feature-set / encoding vectors are accepted and
  // ignored.
  bool inference(Camera& camera, unsigned long long /*cudaSurfaceObject_t*/ output_surf, int batch_size, int num_samples,
                 std::vector<::FeatureSet*>& feature_sets, std::vector<::Encoding>& encodings);
  // Same into a host fp32 buffer [W*H*3] (copies inside).
  bool inference_host(const Camera& camera, float* h_rgb, int batch_size, int num_samples, int32_t* h_nsamples = nullptr);

  const char* last_error() const;
  bool stats(adn_stats* out);

 private:
  adn_ctx* ctx_ = nullptr;
  float thr_ = 0.f;
  std::string err_;
};

}  // namespace adn_host
