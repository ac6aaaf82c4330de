#include "export_loader.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>

namespace adn {
namespace {

// ---- protobuf wire format ------------------------------------------------------------------
struct Cursor {
  const uint8_t* p;
  const uint8_t* end;
  bool ok = true;
  uint64_t varint() {
    uint64_t r = 0;
    int shift = 0;
    while (p < end) {
      const uint8_t b = *p++;
      r |= uint64_t(b & 0x7F) << shift;
      if (!(b & 0x80)) return r;
      shift += 7;
      if (shift > 63) break;
    }
    ok = false;
    return 0;
  }
};

struct Field {
  uint32_t number;
  uint32_t wire;
  uint64_t value;          // wire 0
  const uint8_t* data;     // wire 1, 2, 5
  size_t size;
};

bool next_field(Cursor& c, Field& f) {
  if (c.p >= c.end || !c.ok) return false;
  const uint64_t key = c.varint();
  if (!c.ok) return false;
  f.number = uint32_t(key >> 3);
  f.wire = uint32_t(key & 7);
  f.value = 0;
  f.data = nullptr;
  f.size = 0;
  switch (f.wire) {
    case 0: f.value = c.varint(); break;
    case 1:
      if (size_t(c.end - c.p) < 8) { c.ok = false; return false; }   // truncated file: an error, not "end of message"
      f.data = c.p; f.size = 8; c.p += 8;
      break;
    case 2: {
      const uint64_t n = c.varint();
      if (!c.ok || n > uint64_t(c.end - c.p)) { c.ok = false; return false; }
      f.data = c.p; f.size = size_t(n); c.p += n;
      break;
    }
    case 5:
      if (size_t(c.end - c.p) < 4) { c.ok = false; return false; }
      f.data = c.p; f.size = 4; c.p += 4;
      break;
    default: c.ok = false; return false;
  }
  return c.ok && c.p <= c.end;
}

// TensorProto: dims(1) data_type(2) float_data(4) name(8) raw_data(9)
bool parse_tensor(const uint8_t* p, size_t n, NamedTensor& t, bool& is_float) {
  Cursor c{p, p + n};
  Field f;
  std::vector<int64_t> dims;
  const uint8_t* raw = nullptr;
  size_t raw_n = 0;
  std::vector<float> floats;
  int dtype = 0;
  while (next_field(c, f)) {
    if (f.number == 1) {
      if (f.wire == 0) dims.push_back(int64_t(f.value));
      else if (f.wire == 2) { Cursor d{f.data, f.data + f.size}; while (d.p < d.end && d.ok) dims.push_back(int64_t(d.varint())); }
    } else if (f.number == 2 && f.wire == 0) {
      dtype = int(f.value);
    } else if (f.number == 4) {
      if (f.wire == 5) { float v; std::memcpy(&v, f.data, 4); floats.push_back(v); }
      else if (f.wire == 2) { const size_t k = f.size / 4; const size_t o = floats.size(); floats.resize(o + k); std::memcpy(floats.data() + o, f.data, k * 4); }
    } else if (f.number == 8 && f.wire == 2) {
      t.name.assign(reinterpret_cast<const char*>(f.data), f.size);
    } else if (f.number == 9 && f.wire == 2) {
      raw = f.data; raw_n = f.size;
    }
  }
  if (!c.ok) return false;
  is_float = (dtype == 1);
  if (!is_float) return true;
  if (raw) { t.data.resize(raw_n / 4); std::memcpy(t.data.data(), raw, raw_n / 4 * 4); }   // little-endian fp32
  else t.data = std::move(floats);
  if (dims.size() == 2) { t.rows = dims[0]; t.cols = dims[1]; }
  else if (dims.size() == 1) { t.rows = dims[0]; t.cols = 1; }
  else { t.rows = int64_t(t.data.size()); t.cols = 1; }
  return int64_t(t.data.size()) == t.rows * t.cols;
}

std::string strip(const std::string& s) {
  std::string r;
  for (char ch : s) if (!std::isspace(static_cast<unsigned char>(ch))) r.push_back(ch);
  return r;
}

// "[a, b]" / "a,b" -> tokens (whitespace stripped like Config::store, config.cpp:200-204)
std::vector<std::string> list_items(const std::string& v) {
  std::string s = strip(v);
  if (!s.empty() && s.front() == '[') s.erase(0, 1);
  if (!s.empty() && s.back() == ']') s.pop_back();
  std::vector<std::string> out;
  std::stringstream ss(s);
  std::string item;
  while (std::getline(ss, item, ',')) out.push_back(item);
  return out;
}

bool read_kv(const std::string& path, std::map<std::string, std::string>& kv) {
  std::ifstream f(path);
  if (!f) return false;
  std::string line;
  while (std::getline(f, line)) {
    const size_t eq = line.find('=');
    if (eq == std::string::npos) continue;
    kv[strip(line.substr(0, eq))] = line.substr(eq + 1);
  }
  return true;
}

bool floats_of(const std::map<std::string, std::string>& kv, const char* key, float* dst, size_t n) {
  auto it = kv.find(key);
  if (it == kv.end()) return false;
  const auto items = list_items(it->second);
  if (items.size() < n) return false;
  for (size_t i = 0; i < n; ++i) dst[i] = float(std::atof(items[i].c_str()));
  return true;
}

}  // namespace

bool read_onnx_initializers(const std::string& path, std::vector<NamedTensor>& out, std::string& err) {
  std::ifstream f(path, std::ios::binary);
  if (!f) { err = "cannot open " + path; return false; }
  std::vector<uint8_t> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  Cursor c{buf.data(), buf.data() + buf.size()};
  Field fld;
  while (next_field(c, fld)) {
    if (fld.number == 7 && fld.wire == 2) {  // ModelProto.graph
      Cursor g{fld.data, fld.data + fld.size};
      Field gf;
      while (next_field(g, gf)) {
        if (gf.number == 5 && gf.wire == 2) {  // GraphProto.initializer
          NamedTensor t;
          bool is_float = false;
          if (!parse_tensor(gf.data, gf.size, t, is_float)) { err = "malformed initializer in " + path; return false; }
          if (is_float) out.push_back(std::move(t));
        }
      }
      if (!g.ok) { err = "malformed graph in " + path; return false; }
    }
  }
  if (!c.ok) { err = "malformed protobuf " + path; return false; }
  if (out.empty()) { err = "no fp32 initializers in " + path; return false; }
  return true;
}

bool load_export_dir(const std::string& dir_in, ExportDir& out, std::string& err) {
  std::string dir = dir_in;
  if (!dir.empty() && dir.back() != '/') dir.push_back('/');
  std::map<std::string, std::string> cfg, info;
  if (!read_kv(dir + "config.ini", cfg)) { err = "cannot read " + dir + "config.ini"; return false; }
  if (!read_kv(dir + "dataset_info.txt", info)) { err = "cannot read " + dir + "dataset_info.txt"; return false; }
  adn_scene& s = out.scene;
  float fov = 0, md = 0;
  if (!floats_of(info, "view_cell_center", s.view_cell_center, 3) || !floats_of(info, "view_cell_size", s.view_cell_size, 3) ||
      !floats_of(info, "depth_range", s.depth_range, 2) || !floats_of(info, "fov", &fov, 1) || !floats_of(info, "max_depth", &md, 1)) {
    err = "dataset_info.txt: missing view_cell_center / view_cell_size / depth_range / fov / max_depth";
    return false;
  }
  s.fov = fov;
  s.max_depth = md;
  float zn[2] = {0.001f, 0.001f}, zf[2] = {1.0f, 1.0f};
  floats_of(cfg, "zNear", zn, 2);
  floats_of(cfg, "zFar", zf, 2);
  s.z_near = zn[1];
  s.z_far = zf[1];
  s.n_freq_pos = 10;
  s.n_freq_dir = 4;
  auto pe = cfg.find("posEncArgs");
  if (pe != cfg.end()) {
    const auto items = list_items(pe->second);   // [sampling net, shading net], e.g. [10-4, 10-4] or [2-2, 10-4]
    auto parse = [](const std::string& it, int32_t& pos, int32_t& dir) {
      const size_t dash = it.find('-');
      if (dash == std::string::npos) return;
      pos = std::atoi(it.substr(0, dash).c_str());
      dir = std::atoi(it.substr(dash + 1).c_str());
    };
    if (!items.empty()) parse(items.back(), s.n_freq_pos, s.n_freq_dir);
    if (items.size() >= 2) parse(items.front(), s.n_freq_pos0, s.n_freq_dir0);
  }
  auto ndc = cfg.find("useNDC");
  s.use_ndc = (ndc != cfg.end() && strip(ndc->second) == "True") ? 1 : 0;   // config.cpp:265-266
  if (s.use_ndc) {
    // the reference's export does not record the image size (src/export.py:47-54); accept optional w / h / focal keys,
    // otherwise ndc_rays uses the size of the frame being rendered (featureset.cpp:83-84)
    float v = 0;
    if (floats_of(info, "w", &v, 1)) s.ndc_w = int32_t(v);
    if (floats_of(info, "h", &v, 1)) s.ndc_h = int32_t(v);
  }
  auto th = cfg.find("adaptiveSamplingThreshold");
  out.threshold = th == cfg.end() ? 0.0f : float(std::atof(strip(th->second).c_str()));
  auto ns = cfg.find("numRaymarchSamples");
  if (ns == cfg.end()) { err = "config.ini: numRaymarchSamples missing"; return false; }
  const auto nitems = list_items(ns->second);
  if (nitems.empty()) { err = "config.ini: numRaymarchSamples empty"; return false; }
  out.num_samples = std::atoi(nitems.back().c_str());
  auto check = [&](const char* key, const char* want) {
    auto it = cfg.find(key);
    if (it == cfg.end()) return true;
    const auto items = list_items(it->second);
    return items.empty() || items.back() == want;
  };
  if (s.use_ndc) {
    if (!check("rayMarchSampler", "FromClassifiedDepthAdaptiveNoDepthRange") || !check("rayMarchNormalization", "None") ||
        !check("accumulationMult", "alpha")) {
      err = "config.ini: NDC exports must use FromClassifiedDepthAdaptiveNoDepthRange / rayMarchNormalization None / alpha";
      return false;
    }
  } else if (!check("rayMarchSampler", "FromClassifiedDepthAdaptive") || !check("rayMarchNormalization", "InverseSqrtDistCentered") ||
             !check("depthTransform", "log") || !check("accumulationMult", "alpha")) {
    err = "config.ini: only FromClassifiedDepthAdaptive / InverseSqrtDistCentered / log / alpha exports are supported";
    return false;
  }
  for (int i = 0; i < 2; ++i)
    if (!read_onnx_initializers(dir + "model" + std::to_string(i) + ".onnx", out.nets[i], err)) return false;
  return true;
}

}  // namespace adn
