// Headless counterpart of the viewer's main (adanerf_real_time_viewer/src/main.cpp:15-98): same positional
// model-path argument and -s / -bs options, renders `-f` frames along a small orbit inside the view cell,
// prints the per-frame time the way ImageGenerator::inference logs its 100-frame averages
// (imagegenerator.cpp:370-393) and optionally writes the last frame as a binary PPM (-w).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime_api.h>

#include "../../../include/adanerf_b200_multi.h"
#include "image_generator.h"

class FeatureSet {};   // stand-ins for the viewer's classes: only their names appear in ImageGenerator::inference
class Encoding {};

int main(int argc, char** argv) {
  std::string model = "sample/";
  int W = 800, H = 800, batch = -1, frames = 20, device = 0, gpus = 1;
  bool write = false, surface = false;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if ((a == "-s" || a == "--size") && i + 2 < argc) { W = std::atoi(argv[++i]); H = std::atoi(argv[++i]); }
    else if ((a == "-bs" || a == "--batchSize") && i + 1 < argc) batch = std::atoi(argv[++i]);
    else if ((a == "-f" || a == "--frames") && i + 1 < argc) frames = std::atoi(argv[++i]);
    else if ((a == "-dev" || a == "--device") && i + 1 < argc) device = std::atoi(argv[++i]);
    else if (a == "-w" || a == "--writeImages") write = true;
    else if (a == "--surface") surface = true;   // one frame through ImageGenerator::inference(camera, cudaSurfaceObject_t, ...)
    else if ((a == "-g" || a == "--gpus") && i + 1 < argc) gpus = std::atoi(argv[++i]);
    else if (a[0] != '-') model = a;
    else { std::fprintf(stderr, "usage: %s modelPath [-s W H] [-bs raysPerBatch] [-f frames] [-dev id] [-g gpus] [--surface] [-w]\n", argv[0]); return 2; }
  }
  adn_host::Config config;
  if (!config.load(model)) { std::fprintf(stderr, "couldn't read export directory %s\n", model.c_str()); return 1; }
  std::printf("model %s: K = %d, adaptiveSamplingThreshold = %g\n", model.c_str(), config.numRaymarchSamples, config.adaptiveSamplingThreshold);
  if (gpus > 1) {
    // Row bands over `gpus` devices of this node + one NCCL gather per frame (include/adanerf_b200_multi.h); two frames
    // in flight, so the gather of a frame overlaps the next frame's sampling MLP.
    adn_multi* m = nullptr;
    float thr = 0.f;
    int K = 0;
    if (adn_multi_create_from_export_dir(&m, model.c_str(), nullptr, gpus, &thr, &K) != ADN_OK) { std::fprintf(stderr, "multi-GPU load failed\n"); return 1; }
    adn_host::Camera cam;
    cam.width = W;
    cam.height = H;
    std::vector<float> rgb(size_t(W) * H * 3);
    float rot[9];
    auto pose_at = [&](int f) {
      const float t = 6.2831853f * float(f) / float(frames > 0 ? frames : 1);
      for (int a = 0; a < 3; ++a) cam.pos[a] = config.scene.view_cell_center[a];
      cam.pos[0] += 0.3f * config.scene.view_cell_size[0] * std::cos(t);
      cam.pos[1] += 0.3f * config.scene.view_cell_size[1] * std::sin(t);
      cam.yaw = t;
      cam.rotation(rot);
    };
    std::chrono::steady_clock::time_point t0;
    for (int f = 0; f < frames + 2; ++f) {
      if (f == 2) {   // two warm-up frames (allocation, NCCL channel setup): drain, then time a full pipeline
        if (adn_multi_wait_frame(m, nullptr, nullptr) != ADN_OK) { std::fprintf(stderr, "%s\n", adn_multi_last_error(m)); return 1; }
        t0 = std::chrono::steady_clock::now();
      }
      pose_at(f);
      if (adn_multi_render_camera(m, cam.pos, rot, W, H, thr, K) != ADN_OK) { std::fprintf(stderr, "%s\n", adn_multi_last_error(m)); return 1; }
      if (f >= 1 && f != 2 && adn_multi_wait_frame(m, nullptr, nullptr) != ADN_OK) { std::fprintf(stderr, "%s\n", adn_multi_last_error(m)); return 1; }
    }
    if (adn_multi_wait_frame(m, nullptr, rgb.data()) != ADN_OK) { std::fprintf(stderr, "%s\n", adn_multi_last_error(m)); return 1; }
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    const size_t ng = size_t(gpus);
    std::vector<float> r_ms(ng, 0.f), g_ms(ng, 0.f);
    adn_multi_last_times(m, r_ms.data(), g_ms.data());
    float r_max = 0.f, g_max = 0.f;
    for (int i = 0; i < gpus; ++i) { r_max = std::max(r_max, r_ms[size_t(i)]); g_max = std::max(g_max, g_ms[size_t(i)]); }
    std::printf("%d frames %dx%d on %d GPUs: %.3f ms/frame (%.1f fps); last frame: slowest band %.3f ms, gather %.3f ms\n", frames, W, H, gpus,
                ms / frames, 1000.0 * frames / ms, r_max, g_max);
    double sum = 0;
    for (float v : rgb) sum += v;
    std::printf("checksum %.6f\n", sum);
    adn_multi_destroy(m);
    return 0;
  }
  adn_host::ImageGenerator gen;
  if (!gen.load(config, device)) { std::fprintf(stderr, "load failed: %s\n", gen.last_error()); return 1; }
  adn_host::Camera cam;
  cam.width = W;
  cam.height = H;
  std::vector<float> rgb(size_t(W) * H * 3);
  double total_ms = 0;
  for (int f = 0; f < frames + 2; ++f) {
    const float t = 6.2831853f * float(f) / float(frames > 0 ? frames : 1);
    for (int a = 0; a < 3; ++a) cam.pos[a] = config.scene.view_cell_center[a];
    cam.pos[0] += 0.3f * config.scene.view_cell_size[0] * std::cos(t);
    cam.pos[1] += 0.3f * config.scene.view_cell_size[1] * std::sin(t);
    cam.yaw = t;
    const auto t0 = std::chrono::steady_clock::now();
    if (!gen.inference_host(cam, rgb.data(), batch, config.numRaymarchSamples)) { std::fprintf(stderr, "inference failed: %s\n", gen.last_error()); return 1; }
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (f >= 2) total_ms += ms;   // two warm-up frames (allocation)
  }
  if (surface) {
    // The viewer's frame path: a cudaArray with surface load / store (what cudaGraphicsGLRegisterImage hands out,
    // interoprenderbuffer.cpp:53-83), a surface object on it, ImageGenerator::inference with the reference's signature.
    cudaArray_t arr = nullptr;
    const cudaChannelFormatDesc fmt = cudaCreateChannelDesc(8, 8, 8, 8, cudaChannelFormatKindUnsigned);   // uchar4
    cudaResourceDesc res{};
    cudaSurfaceObject_t surf = 0;
    if (cudaMallocArray(&arr, &fmt, size_t(W), size_t(H), cudaArraySurfaceLoadStore) != cudaSuccess) { std::fprintf(stderr, "cudaMallocArray failed\n"); return 1; }
    res.resType = cudaResourceTypeArray;
    res.res.array.array = arr;
    if (cudaCreateSurfaceObject(&surf, &res) != cudaSuccess) { std::fprintf(stderr, "cudaCreateSurfaceObject failed\n"); return 1; }
    std::vector<FeatureSet*> fs;
    std::vector<Encoding> enc;
    if (!gen.inference(cam, (unsigned long long)surf, batch, config.numRaymarchSamples, fs, enc)) { std::fprintf(stderr, "inference failed: %s\n", gen.last_error()); return 1; }
    cudaDeviceSynchronize();
    std::vector<unsigned char> px(size_t(W) * H * 4);
    if (cudaMemcpy2DFromArray(px.data(), size_t(W) * 4, arr, 0, 0, size_t(W) * 4, size_t(H), cudaMemcpyDeviceToHost) != cudaSuccess) { std::fprintf(stderr, "readback failed\n"); return 1; }
    size_t bad = 0;   // rgb holds the same camera's fp32 frame (last loop iteration): clamp * 255, alpha 255
    for (size_t i = 0; i < size_t(W) * H; ++i) {
      for (int c = 0; c < 3; ++c) {
        const float v = rgb[3 * i + c] < 0.f ? 0.f : (rgb[3 * i + c] > 1.f ? 1.f : rgb[3 * i + c]);
        if (px[4 * i + c] != (unsigned char)(v * 255.0f)) ++bad;
      }
      if (px[4 * i + 3] != 255) ++bad;
    }
    std::printf("surface frame %dx%d: %zu mismatching bytes against the fp32 frame\n", W, H, bad);
    cudaDestroySurfaceObject(surf);
    cudaFreeArray(arr);
    if (bad) return 1;
  }
  adn_stats st{};
  gen.stats(&st);
  std::printf("%d frames %dx%d: %.3f ms/frame (%.1f fps), last frame %lld samples (%.2f per ray), %lld kernel launches\n", frames, W, H,
              total_ms / frames, 1000.0 * frames / total_ms, (long long)st.n_samples, double(st.n_samples) / (double(W) * H),
              (long long)st.kernel_launches);
  if (write) {
    const std::string out = model + "/adn_frame.ppm";
    FILE* fp = std::fopen(out.c_str(), "wb");
    if (!fp) { std::fprintf(stderr, "cannot write %s\n", out.c_str()); return 1; }
    std::fprintf(fp, "P6\n%d %d\n255\n", W, H);
    for (size_t i = 0; i < rgb.size(); ++i) {
      const float v = rgb[i] < 0.f ? 0.f : (rgb[i] > 1.f ? 1.f : rgb[i]);
      std::fputc(int(v * 255.0f), fp);
    }
    std::fclose(fp);
    std::printf("wrote %s\n", out.c_str());
  }
  return 0;
}
