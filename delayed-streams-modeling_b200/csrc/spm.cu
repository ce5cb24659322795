// SentencePiece detokenisation for the Word messages (host code only): the reference turns a word's token ids into
// text with SentencePieceProcessor::decode_piece_ids (moshi-server/src/batched_asr.rs:667, asr.rs:275) -- the
// `sentencepiece` crate over Google's C++ library.  That library is not vendored; this file restates its published
// DecodeIds algorithm for the model kinds the DSM tokenisers use (unigram / BPE with byte fallback) on top of a
// minimal protobuf reader for sentencepiece_model.proto:
//   ModelProto      { repeated SentencePiece pieces = 1; TrainerSpec trainer_spec = 2; NormalizerSpec normalizer_spec = 3 }
//   SentencePiece   { string piece = 1; float score = 2; Type type = 3 }   NORMAL 1, UNKNOWN 2, CONTROL 3,
//                                                                          USER_DEFINED 4, UNUSED 5, BYTE 6
//   TrainerSpec     { bool treat_whitespace_as_suffix = 24; string unk_surface = 44 [default " ⁇ "] }
//   NormalizerSpec  { bool add_dummy_prefix = 3 [true]; bool remove_extra_whitespaces = 4 [true] }
// Pinned against the library itself: tests/golden/tiny_spm_decodes.json (tests/test_spm_cpu.py).
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "common.cuh"
#include "runtime.cuh"

namespace dsm {

namespace {

struct PbReader {
  const uint8_t* p;
  const uint8_t* end;
  bool done() const { return p >= end; }
  uint64_t varint() {
    uint64_t v = 0;
    for (int shift = 0; shift < 64; shift += 7) {
      DSM_CHECK(p < end, DSM_ERR_INVALID_ARG, "sentencepiece model: truncated varint");
      const uint8_t b = *p++;
      v |= (uint64_t)(b & 0x7f) << shift;
      if (!(b & 0x80)) return v;
    }
    fail(DSM_ERR_INVALID_ARG, "sentencepiece model: varint too long");
  }
  // reads one field header; returns false at the end
  bool field(uint32_t* num, uint32_t* wire) {
    if (done()) return false;
    const uint64_t k = varint();
    *num = (uint32_t)(k >> 3), *wire = (uint32_t)(k & 7);
    return true;
  }
  PbReader bytes() {
    const uint64_t n = varint();
    DSM_CHECK((uint64_t)(end - p) >= n, DSM_ERR_INVALID_ARG, "sentencepiece model: truncated field");
    PbReader r{p, p + n};
    p += n;
    return r;
  }
  void skip(uint32_t wire) {
    switch (wire) {
      case 0: varint(); break;
      case 1: DSM_CHECK(end - p >= 8, DSM_ERR_INVALID_ARG, "sentencepiece model: truncated fixed64"); p += 8; break;
      case 2: bytes(); break;
      case 5: DSM_CHECK(end - p >= 4, DSM_ERR_INVALID_ARG, "sentencepiece model: truncated fixed32"); p += 4; break;
      default: fail(DSM_ERR_INVALID_ARG, "sentencepiece model: unsupported wire type %u", wire);
    }
  }
};

const char kSpace[] = "\xe2\x96\x81";   // U+2581 LOWER ONE EIGHTH BLOCK, the whitespace marker

}  // namespace

struct SpmModel {
  enum Type { NORMAL = 1, UNKNOWN = 2, CONTROL = 3, USER_DEFINED = 4, UNUSED = 5, BYTE = 6 };
  std::vector<std::string> pieces;
  std::vector<float> scores;
  std::vector<uint8_t> types;
  std::vector<int16_t> byte_value;   // BYTE pieces "<0xNN>": NN, else -1
  std::string unk_surface = " \xe2\x81\x87 ";
  bool add_dummy_prefix = true, remove_extra_whitespaces = true, treat_whitespace_as_suffix = false;

  void parse(const uint8_t* data, size_t n) {
    PbReader r{data, data + n};
    uint32_t num, wire;
    while (r.field(&num, &wire)) {
      if (num == 1 && wire == 2) {
        PbReader sp = r.bytes();
        std::string piece;
        float score = 0.f;
        uint32_t type = NORMAL;
        uint32_t n2, w2;
        while (sp.field(&n2, &w2)) {
          if (n2 == 1 && w2 == 2) {
            PbReader s = sp.bytes();
            piece.assign(reinterpret_cast<const char*>(s.p), s.end - s.p);
          } else if (n2 == 2 && w2 == 5) {
            DSM_CHECK(sp.end - sp.p >= 4, DSM_ERR_INVALID_ARG, "sentencepiece model: truncated score");
            memcpy(&score, sp.p, 4);
            sp.p += 4;
          } else if (n2 == 3 && w2 == 0) {
            type = (uint32_t)sp.varint();
          } else {
            sp.skip(w2);
          }
        }
        int16_t bv = -1;
        if (type == BYTE && piece.size() == 6 && piece.compare(0, 3, "<0x") == 0 && piece[5] == '>')
          bv = (int16_t)strtol(piece.substr(3, 2).c_str(), nullptr, 16);
        pieces.push_back(std::move(piece));
        scores.push_back(score);
        types.push_back((uint8_t)type);
        byte_value.push_back(bv);
      } else if (num == 2 && wire == 2) {
        PbReader t = r.bytes();
        uint32_t n2, w2;
        while (t.field(&n2, &w2)) {
          if (n2 == 44 && w2 == 2) {
            PbReader s = t.bytes();
            unk_surface.assign(reinterpret_cast<const char*>(s.p), s.end - s.p);
          } else if (n2 == 24 && w2 == 0) {
            treat_whitespace_as_suffix = t.varint() != 0;
          } else {
            t.skip(w2);
          }
        }
      } else if (num == 3 && wire == 2) {
        PbReader t = r.bytes();
        uint32_t n2, w2;
        while (t.field(&n2, &w2)) {
          if (n2 == 3 && w2 == 0) add_dummy_prefix = t.varint() != 0;
          else if (n2 == 4 && w2 == 0) remove_extra_whitespaces = t.varint() != 0;
          else t.skip(w2);
        }
      } else {
        r.skip(wire);
      }
    }
    DSM_CHECK(!pieces.empty(), DSM_ERR_INVALID_ARG, "sentencepiece model: no pieces");
  }

  // One well-formed UTF-8 character at s[0..n)?  Returns its length, or 0 (the library then emits U+FFFD for one byte).
  static int utf8_len(const uint8_t* s, size_t n) {
    const uint8_t c = s[0];
    if (c < 0x80) return 1;
    int len;
    uint32_t cp, min;
    if ((c & 0xe0) == 0xc0) len = 2, cp = c & 0x1f, min = 0x80;
    else if ((c & 0xf0) == 0xe0) len = 3, cp = c & 0x0f, min = 0x800;
    else if ((c & 0xf8) == 0xf0) len = 4, cp = c & 0x07, min = 0x10000;
    else return 0;
    if ((size_t)len > n) return 0;
    for (int i = 1; i < len; ++i) {
      if ((s[i] & 0xc0) != 0x80) return 0;
      cp = (cp << 6) | (s[i] & 0x3f);
    }
    if (cp < min || cp > 0x10ffff || (cp >= 0xd800 && cp < 0xe000)) return 0;
    return len;
  }

  std::string decode(const uint32_t* ids, size_t n) const {
    std::string text, bytes;
    bool is_bos_ws = true;   // still at the beginning of the text: a leading whitespace marker is consumed
    auto flush_bytes = [&] {
      size_t i = 0;
      while (i < bytes.size()) {
        const int l = utf8_len(reinterpret_cast<const uint8_t*>(bytes.data()) + i, bytes.size() - i);
        if (l == 0) text += "\xef\xbf\xbd", i += 1;
        else text.append(bytes, i, l), i += l;
      }
      if (!bytes.empty()) is_bos_ws = false;
      bytes.clear();
    };
    for (size_t k = 0; k < n; ++k) {
      DSM_CHECK(ids[k] < pieces.size(), DSM_ERR_INVALID_ARG, "token id %u outside the vocabulary of %zu pieces", ids[k], pieces.size());
      const uint32_t id = ids[k];
      if (types[id] == BYTE && byte_value[id] >= 0) {
        bytes += (char)byte_value[id];
        continue;
      }
      flush_bytes();
      if (types[id] == CONTROL) continue;      // <s>, </s>, ...: invisible
      if (types[id] == UNKNOWN) {
        text += unk_surface;
        is_bos_ws = false;
        continue;
      }
      const std::string& piece = pieces[id];
      size_t start = 0;
      bool consumed = false;
      if (is_bos_ws && !treat_whitespace_as_suffix && (add_dummy_prefix || remove_extra_whitespaces) &&
          piece.compare(0, 3, kSpace) == 0)
        start = 3, consumed = true;
      const size_t before = text.size();
      for (size_t i = start; i < piece.size();) {
        if (piece.compare(i, 3, kSpace) == 0) text += ' ', i += 3;
        else text += piece[i], i += 1;
      }
      // with remove_extra_whitespaces every leading marker goes; otherwise only the dummy prefix
      if (text.size() > before || (consumed && !remove_extra_whitespaces)) is_bos_ws = false;
    }
    flush_bytes();
    return text;
  }
};

}  // namespace dsm

using namespace dsm;

struct dsm_spm {
  SpmModel model;
  std::string last;
};

template <typename F>
static dsm_status spm_guarded(F&& f) {
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

dsm_status dsm_spm_open(const char* path, dsm_spm** out) {
  return spm_guarded([&] {
    DSM_CHECK(path && out, DSM_ERR_INVALID_ARG, "path and out are required");
    *out = nullptr;
    std::ifstream f(path, std::ios::binary);
    DSM_CHECK(f.good(), DSM_ERR_INVALID_ARG, "cannot open %s", path);
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string blob = ss.str();
    std::unique_ptr<dsm_spm> h(new dsm_spm);
    h->model.parse(reinterpret_cast<const uint8_t*>(blob.data()), blob.size());
    *out = h.release();
  });
}
void dsm_spm_close(dsm_spm* h) { delete h; }
int32_t dsm_spm_vocab_size(const dsm_spm* h) { return h ? (int32_t)h->model.pieces.size() : 0; }
const char* dsm_spm_piece(const dsm_spm* h, int32_t id) {
  return (h && id >= 0 && id < (int32_t)h->model.pieces.size()) ? h->model.pieces[id].c_str() : nullptr;
}
// SentencePieceProcessor::decode_piece_ids: UTF-8 text into buf (NUL-terminated when it fits); *len = its length in bytes.
dsm_status dsm_spm_decode(dsm_spm* h, const uint32_t* ids, int32_t n, char* buf, size_t cap, size_t* len) {
  return spm_guarded([&] {
    DSM_CHECK(h && (ids || n == 0) && n >= 0, DSM_ERR_INVALID_ARG, "handle and ids are required");
    const std::string s = h->model.decode(ids, (size_t)n);
    if (len) *len = s.size();
    DSM_CHECK(!buf || s.size() < cap, DSM_ERR_CAPACITY, "decoded text needs %zu bytes, buffer has %zu", s.size() + 1, cap);
    if (buf) memcpy(buf, s.c_str(), s.size() + 1);
  });
}

}  // extern "C"
