// Headless counterpart of the viewer's main (adanerf_real_time_viewer/src/main.cpp:15-98): same positional
// model-path argument and -s / -bs options, renders `-f` frames along a small orbit inside the view cell,
// prints the per-frame time the way ImageGenerator::inference logs its 100-frame averages
// (imagegenerator.cpp:370-393) and optionally writes the last frame as a binary PPM (-w).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "image_generator.h"

int main(int argc, char** argv) {
  std::string model = "sample/";
  int W = 800, H = 800, batch = -1, frames = 20, device = 0;
  bool write = false;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if ((a == "-s" || a == "--size") && i + 2 < argc) { W = std::atoi(argv[++i]); H = std::atoi(argv[++i]); }
    else if ((a == "-bs" || a == "--batchSize") && i + 1 < argc) batch = std::atoi(argv[++i]);
    else if ((a == "-f" || a == "--frames") && i + 1 < argc) frames = std::atoi(argv[++i]);
    else if ((a == "-dev" || a == "--device") && i + 1 < argc) device = std::atoi(argv[++i]);
    else if (a == "-w" || a == "--writeImages") write = true;
    else if (a[0] != '-') model = a;
    else { std::fprintf(stderr, "usage: %s modelPath [-s W H] [-bs raysPerBatch] [-f frames] [-dev id] [-w]\n", argv[0]); return 2; }
  }
  adn_host::Config config;
  if (!config.load(model)) { std::fprintf(stderr, "couldn't read export directory %s\n", model.c_str()); return 1; }
  std::printf("model %s: K = %d, adaptiveSamplingThreshold = %g\n", model.c_str(), config.numRaymarchSamples, config.adaptiveSamplingThreshold);
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
