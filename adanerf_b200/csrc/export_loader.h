// Loader for the reference's export directory {config.ini, dataset_info.txt, model0.onnx, model1.onnx}
// (writer: src/export.py:28-93, src/train_data.py:180-195; reader in the viewer: config.cpp:200-344,
// imagegenerator.cpp:92-147).  Pure C++: a key=value parser and a protobuf wire-format walker that
// pulls the fp32 initialisers out of the ONNX files by name (no TensorRT / onnx libraries needed).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/adanerf_b200.h"

namespace adn {

struct NamedTensor {
  std::string name;
  std::vector<float> data;
  int64_t rows = 0, cols = 0;
};

struct ExportDir {
  adn_scene scene{};
  float threshold = 0.0f;  // adaptiveSamplingThreshold
  int num_samples = 0;     // numRaymarchSamples[1]
  std::vector<NamedTensor> nets[2];
};

bool read_onnx_initializers(const std::string& path, std::vector<NamedTensor>& out, std::string& err);
bool load_export_dir(const std::string& dir, ExportDir& out, std::string& err);

}  // namespace adn
