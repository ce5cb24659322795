// Wire messages either side of the batched step scheduler (host code only): MessagePack encodings of the
// server's InMsg / OutMsg (moshi-server/src/asr.rs:15-34) exactly as rmp_serde writes them with
// `.with_struct_map()` (asr.rs:311-313, batched_asr.rs:877-879): an internally tagged enum is a map whose first
// entry is "type" -> variant name, followed by the variant's fields in declaration order; integers take the
// smallest encoding, f32 is 0xca, f64 is 0xcb, Vec<f32> an array of f32, Vec<u8> an array of integers (bin is
// accepted on input too).  Known-answer bytes: client/rust/kyutai-client/src/stt/protocol.rs:81-96.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "runtime.cuh"

namespace dsm {

// ---------------------------------------------------------------------------------------------
// writer
// ---------------------------------------------------------------------------------------------
class MpWriter {
 public:
  MpWriter(uint8_t* buf, size_t cap) : buf_(buf), cap_(cap) {}
  size_t size() const { return n_; }      // bytes needed (may exceed the capacity: nothing is written past it)
  void u8(uint8_t v) {
    if (n_ < cap_) buf_[n_] = v;
    ++n_;
  }
  void be(uint64_t v, int bytes) {
    for (int i = bytes - 1; i >= 0; --i) u8((uint8_t)(v >> (8 * i)));
  }
  void map(uint32_t n) {
    if (n < 16) u8(0x80 | n);
    else if (n < 65536) u8(0xde), be(n, 2);
    else u8(0xdf), be(n, 4);
  }
  void array(uint32_t n) {
    if (n < 16) u8(0x90 | n);
    else if (n < 65536) u8(0xdc), be(n, 2);
    else u8(0xdd), be(n, 4);
  }
  void str(const char* s, size_t n) {
    if (n < 32) u8(0xa0 | (uint8_t)n);
    else if (n < 256) u8(0xd9), u8((uint8_t)n);
    else if (n < 65536) u8(0xda), be(n, 2);
    else u8(0xdb), be(n, 4);
    for (size_t i = 0; i < n; ++i) u8((uint8_t)s[i]);
  }
  void str(const char* s) { str(s, strlen(s)); }
  void uint(uint64_t v) {
    if (v < 128) u8((uint8_t)v);
    else if (v < 256) u8(0xcc), u8((uint8_t)v);
    else if (v < 65536) u8(0xcd), be(v, 2);
    else if (v < (1ULL << 32)) u8(0xce), be(v, 4);
    else u8(0xcf), be(v, 8);
  }
  void sint(int64_t v) {
    if (v >= 0) return uint((uint64_t)v);
    if (v >= -32) u8((uint8_t)(int8_t)v);
    else if (v >= -128) u8(0xd0), u8((uint8_t)(int8_t)v);
    else if (v >= -32768) u8(0xd1), be((uint16_t)(int16_t)v, 2);
    else if (v >= -2147483648LL) u8(0xd2), be((uint32_t)(int32_t)v, 4);
    else u8(0xd3), be((uint64_t)v, 8);
  }
  void f32(float v) {
    uint32_t u;
    memcpy(&u, &v, 4);
    u8(0xca), be(u, 4);
  }
  void f64(double v) {
    uint64_t u;
    memcpy(&u, &v, 8);
    u8(0xcb), be(u, 8);
  }

 private:
  uint8_t* buf_;
  size_t cap_, n_ = 0;
};

// ---------------------------------------------------------------------------------------------
// reader
// ---------------------------------------------------------------------------------------------
class MpReader {
 public:
  MpReader(const uint8_t* p, size_t n) : p_(p), end_(p + n) {}
  bool done() const { return p_ == end_; }
  uint8_t peek() {
    need(1);
    return *p_;
  }
  uint32_t map_len() {
    const uint8_t t = take();
    if ((t & 0xf0) == 0x80) return t & 0x0f;
    if (t == 0xde) return (uint32_t)be(2);
    if (t == 0xdf) return (uint32_t)be(4);
    fail(DSM_ERR_INVALID_ARG, "msgpack: expected a map, found 0x%02x", t);
  }
  uint32_t array_len() {
    const uint8_t t = take();
    if ((t & 0xf0) == 0x90) return t & 0x0f;
    if (t == 0xdc) return (uint32_t)be(2);
    if (t == 0xdd) return (uint32_t)be(4);
    fail(DSM_ERR_INVALID_ARG, "msgpack: expected an array, found 0x%02x", t);
  }
  // a string's bytes inside the input buffer
  const char* str_view(size_t* len) {
    const uint8_t t = take();
    size_t n;
    if ((t & 0xe0) == 0xa0) n = t & 0x1f;
    else if (t == 0xd9) n = be(1);
    else if (t == 0xda) n = be(2);
    else if (t == 0xdb) n = be(4);
    else fail(DSM_ERR_INVALID_ARG, "msgpack: expected a string, found 0x%02x", t);
    need(n);
    const char* s = reinterpret_cast<const char*>(p_);
    p_ += n;
    *len = n;
    return s;
  }
  std::string str() {
    size_t n = 0;
    const char* s = str_view(&n);
    return std::string(s, n);
  }
  int64_t integer() {
    const uint8_t t = take();
    if (t < 0x80) return t;
    if (t >= 0xe0) return (int8_t)t;
    switch (t) {
      case 0xcc: return (int64_t)be(1);
      case 0xcd: return (int64_t)be(2);
      case 0xce: return (int64_t)be(4);
      case 0xcf: return (int64_t)be(8);
      case 0xd0: return (int8_t)be(1);
      case 0xd1: return (int16_t)be(2);
      case 0xd2: return (int32_t)be(4);
      case 0xd3: return (int64_t)be(8);
    }
    fail(DSM_ERR_INVALID_ARG, "msgpack: expected an integer, found 0x%02x", t);
  }
  double number() {
    const uint8_t t = peek();
    if (t == 0xca) {
      take();
      const uint32_t u = (uint32_t)be(4);
      float f;
      memcpy(&f, &u, 4);
      return f;
    }
    if (t == 0xcb) {
      take();
      const uint64_t u = be(8);
      double d;
      memcpy(&d, &u, 8);
      return d;
    }
    return (double)integer();
  }
  // bin 8/16/32: returns the payload length and leaves the cursor at its first byte; 0xffffffff if not a bin
  bool bin(const uint8_t** data, size_t* n) {
    const uint8_t t = peek();
    if (t != 0xc4 && t != 0xc5 && t != 0xc6) return false;
    take();
    *n = be(t == 0xc4 ? 1 : t == 0xc5 ? 2 : 4);
    need(*n);
    *data = p_;
    p_ += *n;
    return true;
  }
  void skip() {
    const uint8_t t = peek();
    if (t < 0x80 || t >= 0xe0 || t == 0xc0 || t == 0xc2 || t == 0xc3) {
      take();
    } else if ((t & 0xf0) == 0x80 || t == 0xde || t == 0xdf) {
      const uint32_t n = map_len();
      for (uint32_t i = 0; i < 2 * n; ++i) skip();
    } else if ((t & 0xf0) == 0x90 || t == 0xdc || t == 0xdd) {
      const uint32_t n = array_len();
      for (uint32_t i = 0; i < n; ++i) skip();
    } else if ((t & 0xe0) == 0xa0 || t == 0xd9 || t == 0xda || t == 0xdb) {
      str();
    } else if (t == 0xc4 || t == 0xc5 || t == 0xc6) {
      const uint8_t* d;
      size_t n;
      bin(&d, &n);
    } else if (t == 0xca || t == 0xcb) {
      number();
    } else if (t >= 0xcc && t <= 0xd3) {
      integer();
    } else {
      fail(DSM_ERR_UNSUPPORTED, "msgpack: unsupported type 0x%02x", t);
    }
  }

 private:
  const uint8_t* p_;
  const uint8_t* end_;
  void need(size_t n) { DSM_CHECK((size_t)(end_ - p_) >= n, DSM_ERR_INVALID_ARG, "msgpack: truncated message"); }
  uint8_t take() {
    need(1);
    return *p_++;
  }
  uint64_t be(int bytes) {
    need(bytes);
    uint64_t v = 0;
    for (int i = 0; i < bytes; ++i) v = (v << 8) | *p_++;
    return v;
  }
};

}  // namespace dsm

using namespace dsm;

template <typename F>
static dsm_status wire_guarded(F&& f) {
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

// OutMsg -> bytes.  *len receives the encoded size; DSM_ERR_CAPACITY when it exceeds cap (nothing useful written).
dsm_status dsm_wire_encode_out(const dsm_wire_out* m, uint8_t* buf, size_t cap, size_t* len) {
  return wire_guarded([&] {
    DSM_CHECK(m && len && (buf || cap == 0), DSM_ERR_INVALID_ARG, "message and len are required");
    MpWriter w(buf, cap);
    switch (m->kind) {
      case DSM_WIRE_OUT_WORD:
        w.map(3), w.str("type"), w.str("Word"), w.str("text"), w.str(m->text ? m->text : "", m->text ? (size_t)m->text_len : 0);
        w.str("start_time"), w.f64(m->time);
        break;
      case DSM_WIRE_OUT_END_WORD:
        w.map(2), w.str("type"), w.str("EndWord"), w.str("stop_time"), w.f64(m->time);
        break;
      case DSM_WIRE_OUT_MARKER:
        w.map(2), w.str("type"), w.str("Marker"), w.str("id"), w.sint(m->id);
        break;
      case DSM_WIRE_OUT_STEP:
        w.map(4), w.str("type"), w.str("Step"), w.str("step_idx"), w.uint((uint64_t)m->step_idx), w.str("prs");
        w.array((uint32_t)m->n_prs);
        for (int i = 0; i < m->n_prs; ++i) w.f32(m->prs[i]);
        w.str("buffered_pcm"), w.uint((uint64_t)m->buffered_pcm);
        break;
      case DSM_WIRE_OUT_ERROR:
        w.map(2), w.str("type"), w.str("Error"), w.str("message"), w.str(m->text ? m->text : "", m->text ? (size_t)m->text_len : 0);
        break;
      case DSM_WIRE_OUT_READY:
        w.map(1), w.str("type"), w.str("Ready");
        break;
      default:
        fail(DSM_ERR_INVALID_ARG, "unknown OutMsg kind %d", m->kind);
    }
    *len = w.size();
    DSM_CHECK(w.size() <= cap, DSM_ERR_CAPACITY, "OutMsg needs %zu bytes, buffer has %zu", w.size(), cap);
  });
}

// bytes -> OutMsg (client side).  `text` points into `bytes`; prs are copied into prs_buf (capacity prs_cap).
dsm_status dsm_wire_decode_out(const uint8_t* bytes, size_t len, dsm_wire_out* m, float* prs_buf, int32_t prs_cap) {
  return wire_guarded([&] {
    DSM_CHECK(bytes && m, DSM_ERR_INVALID_ARG, "bytes and message are required");
    memset(m, 0, sizeof *m);
    m->kind = -1;
    MpReader r(bytes, len);
    const uint32_t n = r.map_len();
    for (uint32_t i = 0; i < n; ++i) {
      const std::string k = r.str();
      if (k == "type") {
        const std::string t = r.str();
        m->kind = t == "Word" ? DSM_WIRE_OUT_WORD : t == "EndWord" ? DSM_WIRE_OUT_END_WORD : t == "Marker" ? DSM_WIRE_OUT_MARKER
                  : t == "Step" ? DSM_WIRE_OUT_STEP : t == "Error" ? DSM_WIRE_OUT_ERROR : t == "Ready" ? DSM_WIRE_OUT_READY : -1;
        DSM_CHECK(m->kind >= 0, DSM_ERR_INVALID_ARG, "unknown OutMsg type '%s'", t.c_str());
      } else if (k == "text" || k == "message") {
        size_t tl = 0;
        m->text = r.str_view(&tl);     // points into `bytes`
        m->text_len = (int32_t)tl;
      } else if (k == "start_time" || k == "stop_time") {
        m->time = r.number();
      } else if (k == "id") {
        m->id = r.integer();
      } else if (k == "step_idx") {
        m->step_idx = r.integer();
      } else if (k == "buffered_pcm") {
        m->buffered_pcm = r.integer();
      } else if (k == "prs") {
        const uint32_t np = r.array_len();
        DSM_CHECK((int32_t)np <= prs_cap || !prs_buf, DSM_ERR_CAPACITY, "prs has %u entries, buffer holds %d", np, prs_cap);
        for (uint32_t j = 0; j < np; ++j) {
          const float v = (float)r.number();
          if (prs_buf) prs_buf[j] = v;
        }
        m->n_prs = (int32_t)np;
        m->prs = prs_buf;
      } else {
        r.skip();
      }
    }
    DSM_CHECK(m->kind >= 0, DSM_ERR_INVALID_ARG, "OutMsg without a type tag");
  });
}

// InMsg -> bytes (client side).
dsm_status dsm_wire_encode_in(const dsm_wire_in* m, uint8_t* buf, size_t cap, size_t* len) {
  return wire_guarded([&] {
    DSM_CHECK(m && len && (buf || cap == 0), DSM_ERR_INVALID_ARG, "message and len are required");
    MpWriter w(buf, cap);
    switch (m->kind) {
      case DSM_WIRE_IN_INIT: w.map(1), w.str("type"), w.str("Init"); break;
      case DSM_WIRE_IN_PING: w.map(1), w.str("type"), w.str("Ping"); break;
      case DSM_WIRE_IN_MARKER: w.map(2), w.str("type"), w.str("Marker"), w.str("id"), w.sint(m->id); break;
      case DSM_WIRE_IN_AUDIO:
        w.map(2), w.str("type"), w.str("Audio"), w.str("pcm"), w.array((uint32_t)m->n);
        for (int64_t i = 0; i < m->n; ++i) w.f32(m->pcm[i]);
        break;
      case DSM_WIRE_IN_OGG_OPUS:
        w.map(2), w.str("type"), w.str("OggOpus"), w.str("data"), w.array((uint32_t)m->n);
        for (int64_t i = 0; i < m->n; ++i) w.uint(m->data[i]);
        break;
      default: fail(DSM_ERR_INVALID_ARG, "unknown InMsg kind %d", m->kind);
    }
    *len = w.size();
    DSM_CHECK(w.size() <= cap, DSM_ERR_CAPACITY, "InMsg needs %zu bytes, buffer has %zu", w.size(), cap);
  });
}

// bytes -> InMsg (server side, batched_asr.rs:927).  Audio samples are written to pcm_buf (capacity pcm_cap floats),
// OggOpus bytes to data_buf (capacity data_cap); m->n is the element count, m->pcm / m->data point at those buffers.
dsm_status dsm_wire_decode_in(const uint8_t* bytes, size_t len, dsm_wire_in* m, float* pcm_buf, int64_t pcm_cap,
                              uint8_t* data_buf, int64_t data_cap) {
  return wire_guarded([&] {
    DSM_CHECK(bytes && m, DSM_ERR_INVALID_ARG, "bytes and message are required");
    memset(m, 0, sizeof *m);
    m->kind = -1;
    MpReader r(bytes, len);
    const uint32_t n = r.map_len();
    for (uint32_t i = 0; i < n; ++i) {
      const std::string k = r.str();
      if (k == "type") {
        const std::string t = r.str();
        m->kind = t == "Init" ? DSM_WIRE_IN_INIT : t == "Marker" ? DSM_WIRE_IN_MARKER : t == "Audio" ? DSM_WIRE_IN_AUDIO
                  : t == "OggOpus" ? DSM_WIRE_IN_OGG_OPUS : t == "Ping" ? DSM_WIRE_IN_PING : -1;
        DSM_CHECK(m->kind >= 0, DSM_ERR_INVALID_ARG, "unknown InMsg type '%s'", t.c_str());
      } else if (k == "id") {
        m->id = r.integer();
      } else if (k == "pcm") {
        const uint32_t np = r.array_len();
        DSM_CHECK((int64_t)np <= pcm_cap && pcm_buf, DSM_ERR_CAPACITY, "Audio has %u samples, buffer holds %lld", np, (long long)pcm_cap);
        for (uint32_t j = 0; j < np; ++j) pcm_buf[j] = (float)r.number();
        m->n = np, m->pcm = pcm_buf;
      } else if (k == "data") {
        const uint8_t* d = nullptr;
        size_t nb = 0;
        if (r.bin(&d, &nb)) {
          DSM_CHECK((int64_t)nb <= data_cap && data_buf, DSM_ERR_CAPACITY, "OggOpus has %zu bytes, buffer holds %lld", nb, (long long)data_cap);
          memcpy(data_buf, d, nb);
        } else {
          nb = r.array_len();
          DSM_CHECK((int64_t)nb <= data_cap && data_buf, DSM_ERR_CAPACITY, "OggOpus has %zu bytes, buffer holds %lld", nb, (long long)data_cap);
          for (size_t j = 0; j < nb; ++j) data_buf[j] = (uint8_t)r.integer();
        }
        m->n = (int64_t)nb, m->data = data_buf;
      } else {
        r.skip();
      }
    }
    DSM_CHECK(m->kind >= 0, DSM_ERR_INVALID_ARG, "InMsg without a type tag");
  });
}

}  // extern "C"
