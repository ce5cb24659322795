// Weight and configuration I/O around the handles (host code only):
//   * safetensors reader (mmap + header JSON) -> dsm_weights table, what candle's
//     VarBuilder::from_mmaped_safetensors hands the reference's constructors
//     (moshi-server/src/batched_asr.rs:738-758), and a writer for tests / tooling;
//   * reader of the reference's server TOML (configs/stt/*.toml: [modules.asr], [modules.asr.model],
//     [modules.asr.model.transformer], [modules.asr.model.extra_heads]) -> dsm_asr_config.
// The load-time folds (weight-norm conv.rs:27-45, codebook normalisation quantization.rs:86-107, depthwise
// upsample conv.rs:144-150) happen in dsm_mimi_create on the table this file produces.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "common.cuh"
#include "runtime.cuh"

namespace dsm {

// ---------------------------------------------------------------------------------------------
// minimal JSON (the safetensors header: objects, arrays, strings, integers)
// ---------------------------------------------------------------------------------------------
struct JsonValue {
  enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
  bool b = false;
  double num = 0.0;
  std::string str;
  std::vector<JsonValue> arr;
  std::vector<std::pair<std::string, JsonValue>> obj;   // insertion order kept
};

class JsonParser {
 public:
  JsonParser(const char* p, size_t n) : p_(p), end_(p + n) {}
  JsonValue parse() {
    JsonValue v = value(0);
    ws();
    DSM_CHECK(p_ == end_, DSM_ERR_INVALID_ARG, "json: trailing characters");
    return v;
  }

 private:
  const char* p_;
  const char* end_;
  void ws() {
    while (p_ < end_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r')) ++p_;
  }
  char peek() {
    ws();
    DSM_CHECK(p_ < end_, DSM_ERR_INVALID_ARG, "json: unexpected end");
    return *p_;
  }
  void expect(char c) {
    DSM_CHECK(peek() == c, DSM_ERR_INVALID_ARG, "json: expected '%c' at offset %ld", c, (long)(end_ - p_));
    ++p_;
  }
  static void utf8(std::string& s, uint32_t cp) {
    if (cp < 0x80) s += (char)cp;
    else if (cp < 0x800) s += (char)(0xc0 | (cp >> 6)), s += (char)(0x80 | (cp & 0x3f));
    else if (cp < 0x10000) s += (char)(0xe0 | (cp >> 12)), s += (char)(0x80 | ((cp >> 6) & 0x3f)), s += (char)(0x80 | (cp & 0x3f));
    else s += (char)(0xf0 | (cp >> 18)), s += (char)(0x80 | ((cp >> 12) & 0x3f)), s += (char)(0x80 | ((cp >> 6) & 0x3f)), s += (char)(0x80 | (cp & 0x3f));
  }
  uint32_t hex4() {
    DSM_CHECK(end_ - p_ >= 4, DSM_ERR_INVALID_ARG, "json: short \\u escape");
    uint32_t v = 0;
    for (int i = 0; i < 4; ++i) {
      const char c = *p_++;
      v <<= 4;
      if (c >= '0' && c <= '9') v |= c - '0';
      else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
      else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
      else fail(DSM_ERR_INVALID_ARG, "json: bad \\u escape");
    }
    return v;
  }
  std::string string() {
    expect('"');
    std::string s;
    while (true) {
      DSM_CHECK(p_ < end_, DSM_ERR_INVALID_ARG, "json: unterminated string");
      char c = *p_++;
      if (c == '"') break;
      if (c != '\\') {
        s += c;
        continue;
      }
      DSM_CHECK(p_ < end_, DSM_ERR_INVALID_ARG, "json: unterminated escape");
      c = *p_++;
      switch (c) {
        case 'n': s += '\n'; break;
        case 't': s += '\t'; break;
        case 'r': s += '\r'; break;
        case 'b': s += '\b'; break;
        case 'f': s += '\f'; break;
        case 'u': {
          uint32_t cp = hex4();
          if (cp >= 0xd800 && cp < 0xdc00 && end_ - p_ >= 6 && p_[0] == '\\' && p_[1] == 'u') {
            p_ += 2;
            cp = 0x10000 + ((cp - 0xd800) << 10) + (hex4() - 0xdc00);
          }
          utf8(s, cp);
          break;
        }
        default: s += c;
      }
    }
    return s;
  }
  JsonValue value(int depth) {
    DSM_CHECK(depth < 64, DSM_ERR_INVALID_ARG, "json: nesting too deep");
    JsonValue v;
    const char c = peek();
    if (c == '{') {
      ++p_;
      v.kind = JsonValue::Object;
      if (peek() == '}') {
        ++p_;
        return v;
      }
      while (true) {
        std::string k = string();
        expect(':');
        v.obj.emplace_back(std::move(k), value(depth + 1));
        if (peek() == ',') {
          ++p_;
          continue;
        }
        expect('}');
        break;
      }
    } else if (c == '[') {
      ++p_;
      v.kind = JsonValue::Array;
      if (peek() == ']') {
        ++p_;
        return v;
      }
      while (true) {
        v.arr.push_back(value(depth + 1));
        if (peek() == ',') {
          ++p_;
          continue;
        }
        expect(']');
        break;
      }
    } else if (c == '"') {
      v.kind = JsonValue::String;
      v.str = string();
    } else if (c == 't' || c == 'f' || c == 'n') {
      const char* words[3] = {"true", "false", "null"};
      for (const char* w : words) {
        const size_t n = strlen(w);
        if ((size_t)(end_ - p_) >= n && memcmp(p_, w, n) == 0) {
          p_ += n;
          v.kind = w[0] == 'n' ? JsonValue::Null : JsonValue::Bool;
          v.b = w[0] == 't';
          return v;
        }
      }
      fail(DSM_ERR_INVALID_ARG, "json: bad literal");
    } else {
      char* e = nullptr;
      std::string tmp(p_, std::min<size_t>(end_ - p_, 64));
      v.kind = JsonValue::Number;
      v.num = strtod(tmp.c_str(), &e);
      DSM_CHECK(e != tmp.c_str(), DSM_ERR_INVALID_ARG, "json: bad number");
      p_ += e - tmp.c_str();
    }
    return v;
  }
};

// ---------------------------------------------------------------------------------------------
// safetensors
// ---------------------------------------------------------------------------------------------
struct WeightFile {
  int fd = -1;
  void* map = nullptr;
  size_t size = 0;
  std::vector<std::string> names;
  std::vector<dsm_tensor> tensors;
  dsm_weights table{};
  std::vector<std::pair<std::string, std::string>> metadata;
  ~WeightFile() {
    if (map && map != MAP_FAILED) munmap(map, size);
    if (fd >= 0) close(fd);
  }
};

static int dtype_from_name(const std::string& s, size_t* esz) {
  if (s == "F32") return *esz = 4, DSM_F32;
  if (s == "BF16") return *esz = 2, DSM_BF16;
  if (s == "F16") return *esz = 2, DSM_F16;
  if (s == "U32") return *esz = 4, DSM_U32;
  if (s == "I64") return *esz = 8, DSM_I64;
  return -1;
}
static const char* dtype_name(int dt, size_t* esz) {
  switch (dt) {
    case DSM_F32: *esz = 4; return "F32";
    case DSM_BF16: *esz = 2; return "BF16";
    case DSM_F16: *esz = 2; return "F16";
    case DSM_U32: *esz = 4; return "U32";
    case DSM_I64: *esz = 8; return "I64";
  }
  fail(DSM_ERR_INVALID_ARG, "unknown dtype %d", dt);
}

WeightFile* open_safetensors(const char* path) {
  std::unique_ptr<WeightFile> f(new WeightFile);
  f->fd = open(path, O_RDONLY);
  DSM_CHECK(f->fd >= 0, DSM_ERR_INVALID_ARG, "cannot open %s: %s", path, strerror(errno));
  struct stat st;
  DSM_CHECK(fstat(f->fd, &st) == 0, DSM_ERR_INVALID_ARG, "cannot stat %s", path);
  f->size = (size_t)st.st_size;
  DSM_CHECK(f->size >= 8, DSM_ERR_INVALID_ARG, "%s: too short for a safetensors file", path);
  f->map = mmap(nullptr, f->size, PROT_READ, MAP_PRIVATE, f->fd, 0);
  DSM_CHECK(f->map != MAP_FAILED, DSM_ERR_OOM, "mmap of %s failed: %s", path, strerror(errno));
  const uint8_t* base = static_cast<const uint8_t*>(f->map);
  uint64_t hlen = 0;
  for (int i = 7; i >= 0; --i) hlen = (hlen << 8) | base[i];   // little-endian u64
  DSM_CHECK(hlen <= f->size - 8, DSM_ERR_INVALID_ARG, "%s: header length %llu exceeds the file", path, (unsigned long long)hlen);
  JsonValue hdr = JsonParser(reinterpret_cast<const char*>(base + 8), (size_t)hlen).parse();
  DSM_CHECK(hdr.kind == JsonValue::Object, DSM_ERR_INVALID_ARG, "%s: header is not a JSON object", path);
  const uint8_t* data = base + 8 + hlen;
  const size_t data_size = f->size - 8 - (size_t)hlen;
  f->names.reserve(hdr.obj.size());
  for (auto& kv : hdr.obj) {
    if (kv.first == "__metadata__") {
      for (auto& m : kv.second.obj) f->metadata.emplace_back(m.first, m.second.str);
      continue;
    }
    const JsonValue& t = kv.second;
    DSM_CHECK(t.kind == JsonValue::Object, DSM_ERR_INVALID_ARG, "%s: entry '%s' is not an object", path, kv.first.c_str());
    std::string dt;
    std::vector<int64_t> shape;
    uint64_t b = 0, e = 0;
    bool have_off = false;
    for (auto& fld : t.obj) {
      if (fld.first == "dtype") dt = fld.second.str;
      else if (fld.first == "shape")
        for (auto& d : fld.second.arr) shape.push_back((int64_t)d.num);
      else if (fld.first == "data_offsets" && fld.second.arr.size() == 2)
        b = (uint64_t)fld.second.arr[0].num, e = (uint64_t)fld.second.arr[1].num, have_off = true;
    }
    size_t esz = 0;
    const int dtype = dtype_from_name(dt, &esz);
    DSM_CHECK(dtype >= 0, DSM_ERR_UNSUPPORTED, "%s: tensor '%s' has unsupported dtype %s", path, kv.first.c_str(), dt.c_str());
    DSM_CHECK(have_off && b <= e && e <= data_size, DSM_ERR_INVALID_ARG, "%s: tensor '%s' has bad data_offsets", path, kv.first.c_str());
    size_t numel = 1;
    for (int64_t d : shape) numel *= (size_t)d;
    DSM_CHECK(numel * esz == e - b, DSM_ERR_SHAPE, "%s: tensor '%s' shape does not match its %llu bytes", path, kv.first.c_str(),
              (unsigned long long)(e - b));
    f->names.push_back(kv.first);
    dsm_tensor d{};
    d.dtype = dtype;
    // dsm_tensor carries up to 4 dims; higher ranks are flattened (only the element count is checked downstream)
    if (shape.size() <= 4) {
      d.ndim = (int32_t)shape.size();
      for (size_t i = 0; i < shape.size(); ++i) d.shape[i] = shape[i];
    } else {
      d.ndim = 1, d.shape[0] = (int64_t)numel;
    }
    d.data = data + b;
    f->tensors.push_back(d);
  }
  for (size_t i = 0; i < f->tensors.size(); ++i) f->tensors[i].name = f->names[i].c_str();
  f->table.tensors = f->tensors.data();
  f->table.count = (int64_t)f->tensors.size();
  return f.release();
}

static void json_escape(std::ostringstream& o, const std::string& s) {
  o << '"';
  for (unsigned char c : s) {
    if (c == '"' || c == '\\') o << '\\' << c;
    else if (c < 0x20) {
      char b[8];
      snprintf(b, sizeof b, "\\u%04x", c);
      o << b;
    } else o << c;
  }
  o << '"';
}

void save_safetensors(const char* path, const dsm_weights* w) {
  DSM_CHECK(w && w->count >= 0, DSM_ERR_INVALID_ARG, "weights table is required");
  std::ostringstream h;
  h << '{';
  uint64_t off = 0;
  std::vector<uint64_t> sizes((size_t)w->count);
  for (int64_t i = 0; i < w->count; ++i) {
    const dsm_tensor& t = w->tensors[i];
    size_t esz = 0;
    const char* dn = dtype_name(t.dtype, &esz);
    size_t numel = 1;
    for (int d = 0; d < t.ndim; ++d) numel *= (size_t)t.shape[d];
    sizes[i] = numel * esz;
    if (i) h << ',';
    json_escape(h, t.name);
    h << ":{\"dtype\":\"" << dn << "\",\"shape\":[";
    for (int d = 0; d < t.ndim; ++d) h << (d ? "," : "") << t.shape[d];
    h << "],\"data_offsets\":[" << off << ',' << off + sizes[i] << "]}";
    off += sizes[i];
  }
  h << '}';
  std::string hs = h.str();
  while (hs.size() % 8) hs += ' ';           // the format pads the header so the data starts aligned
  std::ofstream f(path, std::ios::binary);
  DSM_CHECK(f.good(), DSM_ERR_INVALID_ARG, "cannot create %s", path);
  uint64_t n = hs.size();
  uint8_t le[8];
  for (int i = 0; i < 8; ++i) le[i] = (uint8_t)(n >> (8 * i));
  f.write(reinterpret_cast<const char*>(le), 8);
  f.write(hs.data(), (std::streamsize)hs.size());
  for (int64_t i = 0; i < w->count; ++i) f.write(static_cast<const char*>(w->tensors[i].data), (std::streamsize)sizes[i]);
  DSM_CHECK(f.good(), DSM_ERR_INVALID_ARG, "write to %s failed", path);
}

// ---------------------------------------------------------------------------------------------
// TOML subset: [dotted.section] headers, key = string | integer | float | bool | [array] (arrays are skipped)
// ---------------------------------------------------------------------------------------------
static std::string trim(const std::string& s) {
  size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
  return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

std::map<std::string, std::string> parse_toml(const std::string& text) {
  std::map<std::string, std::string> kv;   // "section.key" -> raw value (strings unquoted)
  std::istringstream in(text);
  std::string line, section;
  int lineno = 0;
  while (std::getline(in, line)) {
    ++lineno;
    // strip comments outside of strings
    bool in_str = false;
    for (size_t i = 0; i < line.size(); ++i) {
      if (line[i] == '"' && (i == 0 || line[i - 1] != '\\')) in_str = !in_str;
      if (line[i] == '#' && !in_str) {
        line.resize(i);
        break;
      }
    }
    line = trim(line);
    if (line.empty()) continue;
    if (line[0] == '[') {
      const size_t e = line.find(']');
      DSM_CHECK(e != std::string::npos, DSM_ERR_INVALID_ARG, "toml line %d: unterminated section header", lineno);
      section = trim(line.substr(1, e - 1));
      continue;
    }
    const size_t eq = line.find('=');
    DSM_CHECK(eq != std::string::npos, DSM_ERR_INVALID_ARG, "toml line %d: expected key = value", lineno);
    std::string key = trim(line.substr(0, eq)), val = trim(line.substr(eq + 1));
    if (!val.empty() && val[0] == '"') {
      std::string s;
      size_t i = 1;
      for (; i < val.size() && val[i] != '"'; ++i) {
        if (val[i] == '\\' && i + 1 < val.size()) {
          ++i;
          s += val[i] == 'n' ? '\n' : val[i] == 't' ? '\t' : val[i];
        } else s += val[i];
      }
      DSM_CHECK(i < val.size(), DSM_ERR_INVALID_ARG, "toml line %d: unterminated string", lineno);
      val = s;
    }
    kv[section.empty() ? key : section + "." + key] = val;
  }
  return kv;
}

static bool toml_int(const std::map<std::string, std::string>& kv, const std::string& k, int32_t* out) {
  auto it = kv.find(k);
  if (it == kv.end()) return false;
  std::string v = it->second;
  v.erase(std::remove(v.begin(), v.end(), '_'), v.end());
  char* e = nullptr;
  const double d = strtod(v.c_str(), &e);
  DSM_CHECK(e != v.c_str() && *e == 0, DSM_ERR_INVALID_ARG, "toml: %s = '%s' is not a number", k.c_str(), it->second.c_str());
  *out = (int32_t)d;
  return true;
}
static bool toml_float(const std::map<std::string, std::string>& kv, const std::string& k, float* out) {
  auto it = kv.find(k);
  if (it == kv.end()) return false;
  std::string v = it->second;
  v.erase(std::remove(v.begin(), v.end(), '_'), v.end());
  char* e = nullptr;
  const double d = strtod(v.c_str(), &e);
  DSM_CHECK(e != v.c_str() && *e == 0, DSM_ERR_INVALID_ARG, "toml: %s = '%s' is not a number", k.c_str(), it->second.c_str());
  *out = (float)d;
  return true;
}
static void toml_str(const std::map<std::string, std::string>& kv, const std::string& k, char* out, size_t cap) {
  auto it = kv.find(k);
  out[0] = 0;
  if (it == kv.end()) return;
  snprintf(out, cap, "%s", it->second.c_str());
}

// lm::Config as the server deserialises it ([modules.<module>.model.*]); AsrConfig fields beside it
// (moshi-server/src/lib.rs AsrConfig; batched_asr.rs:738-771).
void asr_config_from_toml(const std::string& text, const char* module, dsm_asr_config* c) {
  memset(c, 0, sizeof *c);
  const auto kv = parse_toml(text);
  const std::string m = std::string("modules.") + (module && module[0] ? module : "asr");
  const std::string mm = m + ".model", tr = mm + ".transformer", eh = mm + ".extra_heads";
  auto need = [&](const std::string& k, int32_t* v) {
    DSM_CHECK(toml_int(kv, k, v), DSM_ERR_INVALID_ARG, "toml: missing key %s", k.c_str());
  };
  dsm_lm_spec& s = c->lm;
  need(tr + ".d_model", &s.d_model);
  need(tr + ".num_heads", &s.num_heads);
  need(tr + ".num_layers", &s.num_layers);
  need(tr + ".dim_feedforward", &s.dim_feedforward);
  need(tr + ".context", &s.context);
  s.max_period = 10000.f;
  toml_float(kv, tr + ".max_period", &s.max_period);
  need(mm + ".text_in_vocab_size", &s.text_in_vocab_size);
  need(mm + ".text_out_vocab_size", &s.text_out_vocab_size);
  need(mm + ".audio_vocab_size", &s.audio_vocab_size);
  need(mm + ".audio_codebooks", &s.audio_codebooks);
  toml_int(kv, eh + ".num_heads", &s.extra_heads_num);
  toml_int(kv, eh + ".dim", &s.extra_heads_dim);
  s.rms_eps = 1e-8f;   // batched_transformer.rs:246-248: fixed, not a config key
  // what this library implements of lm::Config; anything else is refused rather than silently ignored
  auto expect_str = [&](const std::string& k, const char* want) {
    auto it = kv.find(k);
    DSM_CHECK(it == kv.end() || it->second == want, DSM_ERR_UNSUPPORTED, "toml: %s = \"%s\" (only \"%s\" is implemented)", k.c_str(),
              it == kv.end() ? "" : it->second.c_str(), want);
  };
  expect_str(tr + ".gating", "silu");
  expect_str(tr + ".norm", "RmsNorm");
  expect_str(tr + ".positional_embedding", "Rope");
  expect_str(tr + ".causal", "true");
  expect_str(tr + ".norm_first", "true");
  expect_str(tr + ".bias_ff", "false");
  expect_str(tr + ".bias_attn", "false");
  expect_str(tr + ".use_conv_block", "false");
  int32_t kv_repeat = 1;
  toml_int(kv, tr + ".kv_repeat", &kv_repeat);
  DSM_CHECK(kv_repeat == 1, DSM_ERR_UNSUPPORTED, "toml: kv_repeat = %d (only 1 is implemented)", kv_repeat);
  // AsrConfig
  toml_int(kv, m + ".asr_delay_in_tokens", &c->asr_delay_in_tokens);
  float cd = 0.f;
  if (toml_float(kv, m + ".conditioning_delay", &cd)) {
    c->has_conditioning_delay = 1, c->conditioning_delay = cd;
    c->asr_delay_in_tokens = (int32_t)(cd * 12.5f) + 1;   // batched_asr.rs:769-770
  }
  c->temperature = 0.f;
  toml_float(kv, m + ".temperature", &c->temperature);
  toml_int(kv, m + ".batch_size", &c->batch_size);
  toml_str(kv, m + ".lm_model_file", c->lm_model_file, sizeof c->lm_model_file);
  toml_str(kv, m + ".audio_tokenizer_file", c->audio_tokenizer_file, sizeof c->audio_tokenizer_file);
  toml_str(kv, m + ".text_tokenizer_file", c->text_tokenizer_file, sizeof c->text_tokenizer_file);
  toml_str(kv, m + ".type", c->module_type, sizeof c->module_type);
}

}  // namespace dsm

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
using namespace dsm;

struct dsm_weight_file {
  WeightFile* impl;
};

template <typename F>
static dsm_status io_guarded(F&& f) {
  try {
    f();
    return DSM_OK;
  } catch (const Error& e) {
    set_last_error(e.what());
    return e.code;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return DSM_ERR_INVALID_ARG;
  }
}

extern "C" {

dsm_status dsm_weights_open(const char* path, dsm_weight_file** out) {
  return io_guarded([&] {
    DSM_CHECK(path && out, DSM_ERR_INVALID_ARG, "path and out are required");
    *out = nullptr;
    *out = new dsm_weight_file{open_safetensors(path)};
  });
}
const dsm_weights* dsm_weight_file_table(const dsm_weight_file* f) { return f ? &f->impl->table : nullptr; }
int64_t dsm_weight_file_metadata(const dsm_weight_file* f, const char* key, char* buf, size_t cap) {
  if (!f || !key) return -1;
  for (auto& kv : f->impl->metadata)
    if (kv.first == key) {
      if (buf && cap) snprintf(buf, cap, "%s", kv.second.c_str());
      return (int64_t)kv.second.size();
    }
  return -1;
}
void dsm_weight_file_close(dsm_weight_file* f) {
  if (!f) return;
  delete f->impl;
  delete f;
}
dsm_status dsm_weights_save(const char* path, const dsm_weights* weights) {
  return io_guarded([&] {
    DSM_CHECK(path, DSM_ERR_INVALID_ARG, "path is required");
    save_safetensors(path, weights);
  });
}
dsm_status dsm_asr_config_from_toml(const char* path, const char* module, dsm_asr_config* out) {
  return io_guarded([&] {
    DSM_CHECK(path && out, DSM_ERR_INVALID_ARG, "path and out are required");
    std::ifstream f(path);
    DSM_CHECK(f.good(), DSM_ERR_INVALID_ARG, "cannot open %s", path);
    std::stringstream ss;
    ss << f.rdbuf();
    asr_config_from_toml(ss.str(), module, out);
  });
}
dsm_status dsm_asr_config_from_toml_text(const char* text, const char* module, dsm_asr_config* out) {
  return io_guarded([&] {
    DSM_CHECK(text && out, DSM_ERR_INVALID_ARG, "text and out are required");
    asr_config_from_toml(text, module, out);
  });
}

}  // extern "C"
