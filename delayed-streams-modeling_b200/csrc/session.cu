// Single-connection ASR session: the non-batched `Asr` module of the server (moshi-server/src/asr.rs:95-330) without
// its socket -- one asr::State of batch size 1 driven by a connection's wire messages.
//
//   recv_loop        (asr.rs:190-233)   dsm_session_push_msg: InMsg bytes; Audio of any length is cut into 1920-sample
//                                       frames (the reference hands each message to Mimi's streaming encode_step as is;
//                                       this library steps whole frames, so the remainder waits for the next message);
//                                       Init / Marker / Ping produce nothing, exactly as there (:209-222)
//   mimi_encode_loop + inference_loop   one dsm_asr_step_pcm per complete frame (the two stages are the two branches of
//   (asr.rs:236-294)                    the frame graph); AsrMsg -> OutMsg with the SentencePiece text (:272-285),
//                                       Step carries prs[h][0] and buffered_pcm = 0 (:277-280)
//   send_loop        (asr.rs:296-330)   dsm_session_poll_msg: OutMsg bytes
//   logger_loop      (asr.rs:126-176)   dsm_session_save_log: every step's text token and audio codes as I64
//                                       safetensors tensors "text" [1, steps] and "audio" [1, codebooks, steps]
#include <deque>
#include <string>
#include <vector>

#include "handles.cuh"

namespace dsm {
void save_safetensors(const char* path, const dsm_weights* w);
}
using namespace dsm;

struct dsm_session {
  AsrImpl* asr;
  dsm_spm* spm;
  std::vector<float> pending;                 // samples that do not fill a frame yet
  std::deque<std::vector<uint8_t>> out;       // encoded OutMsgs
  std::vector<int64_t> log_text;              // [steps]
  std::vector<int64_t> log_audio;             // [steps][codebooks]
  std::vector<uint32_t> codes_host;
  Arena arena;
  uint32_t* h_codes = nullptr;
};

template <typename F>
static dsm_status session_guarded(F&& f) {
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

static void session_emit(dsm_session* s, const dsm_wire_out& m) {
  std::vector<uint8_t> buf(256 + (m.text ? (size_t)m.text_len : 0) + 5 * (size_t)m.n_prs);
  size_t n = 0;
  const dsm_status st = dsm_wire_encode_out(&m, buf.data(), buf.size(), &n);
  if (st != DSM_OK) throw Error(st, last_error());
  buf.resize(n);
  s->out.push_back(std::move(buf));
}

static void session_step(dsm_session* s, const float* frame) {
  AsrImpl* a = s->asr;
  uint32_t tok = 0;
  a->step(frame, nullptr, nullptr, &tok);
  // logger (asr.rs:126-176): the sampled text token and the codes the LM consumed... the reference logs what
  // step_tokens_vec hands its callback: the text token fed in and this frame's audio tokens; we log the sampled token
  // and this frame's codes (the same sequences shifted by the one-frame feed delay)
  s->log_text.push_back((int64_t)tok);
  const int n_cb = a->n_cb;
  DSM_CUDA(cudaMemcpy(s->h_codes, a->mimi->d_codes, (size_t)n_cb * 4, cudaMemcpyDeviceToHost));
  for (int i = 0; i < n_cb; ++i) s->log_audio.push_back((int64_t)s->h_codes[i]);
  for (const AsrMsgStore& m : a->msgs) {
    dsm_wire_out o{};
    std::string text;
    float prs[16];
    if (m.kind == DSM_MSG_STEP) {
      o.kind = DSM_WIRE_OUT_STEP, o.step_idx = m.step_idx, o.buffered_pcm = 0;
      o.n_prs = std::min(a->nh, 16);
      for (int h = 0; h < o.n_prs; ++h) prs[h] = a->prs_copy[(size_t)h * a->B];
      o.prs = prs;
    } else if (m.kind == DSM_MSG_WORD) {
      o.kind = DSM_WIRE_OUT_WORD, o.time = m.time;
      if (s->spm) {
        char tmp[1024];
        size_t n = 0;
        const dsm_status st = dsm_spm_decode(s->spm, m.tokens.data(), (int32_t)m.tokens.size(), tmp, sizeof tmp, &n);
        if (st != DSM_OK) throw Error(st, last_error());
        text.assign(tmp, n);
      } else {
        for (size_t i = 0; i < m.tokens.size(); ++i) text += (i ? " " : "") + std::to_string(m.tokens[i]);
      }
      o.text = text.data(), o.text_len = (int32_t)text.size();
    } else {
      o.kind = DSM_WIRE_OUT_END_WORD, o.time = m.time;
    }
    session_emit(s, o);
  }
}

extern "C" {

dsm_status dsm_session_create(dsm_asr* asr, dsm_spm* spm, dsm_session** out) {
  return session_guarded([&] {
    DSM_CHECK(asr && out, DSM_ERR_INVALID_ARG, "asr and out are required");
    *out = nullptr;
    DSM_CHECK(asr->impl->B == 1 && asr->impl->mimi, DSM_ERR_INVALID_ARG, "a session drives an asr state of batch size 1 with a Mimi encoder");
    std::unique_ptr<dsm_session> s(new dsm_session);
    s->asr = asr->impl, s->spm = spm;
    DSM_CUDA(cudaSetDevice(asr->impl->lm->device));
    s->h_codes = static_cast<uint32_t*>(s->arena.alloc_pinned((size_t)asr->impl->n_cb * 4));
    *out = s.release();
  });
}
void dsm_session_destroy(dsm_session* s) { delete s; }

dsm_status dsm_session_push_msg(dsm_session* s, const uint8_t* bytes, size_t len) {
  return session_guarded([&] {
    DSM_CHECK(s && bytes, DSM_ERR_INVALID_ARG, "session and bytes are required");
    std::vector<float> pcm(len / 5 + 16);
    std::vector<uint8_t> data(len + 16);
    dsm_wire_in m;
    const dsm_status st = dsm_wire_decode_in(bytes, len, &m, pcm.data(), (int64_t)pcm.size(), data.data(), (int64_t)data.size());
    if (st != DSM_OK) throw Error(st, last_error());
    if (m.kind == DSM_WIRE_IN_OGG_OPUS) fail(DSM_ERR_UNSUPPORTED, "OggOpus payloads must be decoded to PCM before they reach the session");
    if (m.kind != DSM_WIRE_IN_AUDIO) return;                  // Init, Marker, Ping: nothing (asr.rs:209-222)
    s->pending.insert(s->pending.end(), m.pcm, m.pcm + m.n);
    size_t off = 0;
    while (s->pending.size() - off >= DSM_FRAME_SIZE) {
      session_step(s, s->pending.data() + off);
      off += DSM_FRAME_SIZE;
    }
    s->pending.erase(s->pending.begin(), s->pending.begin() + (long)off);
  });
}

dsm_status dsm_session_poll_msg(dsm_session* s, uint8_t* buf, size_t cap, size_t* len, int32_t* got) {
  return session_guarded([&] {
    DSM_CHECK(s && buf && len && got, DSM_ERR_INVALID_ARG, "session, buffer, len and got are required");
    *got = 0, *len = 0;
    if (s->out.empty()) return;
    const std::vector<uint8_t>& m = s->out.front();
    *len = m.size();
    DSM_CHECK(m.size() <= cap, DSM_ERR_CAPACITY, "message of %zu bytes, buffer has %zu", m.size(), cap);
    memcpy(buf, m.data(), m.size());
    s->out.pop_front();
    *got = 1;
  });
}

int64_t dsm_session_steps(const dsm_session* s) { return s ? (int64_t)s->log_text.size() : 0; }

dsm_status dsm_session_save_log(const dsm_session* s, const char* safetensors_path) {
  return session_guarded([&] {
    DSM_CHECK(s && safetensors_path, DSM_ERR_INVALID_ARG, "session and path are required");
    const int64_t steps = (int64_t)s->log_text.size();
    const int n_cb = s->asr->n_cb;
    // audio as [1, codebooks, steps] (asr.rs:152-157)
    std::vector<int64_t> audio((size_t)steps * n_cb);
    for (int64_t t = 0; t < steps; ++t)
      for (int c = 0; c < n_cb; ++c) audio[(size_t)c * steps + t] = s->log_audio[(size_t)t * n_cb + c];
    dsm_tensor t[2]{};
    t[0].name = "audio", t[0].dtype = DSM_I64, t[0].ndim = 3, t[0].shape[0] = 1, t[0].shape[1] = n_cb, t[0].shape[2] = steps;
    t[0].data = audio.data();
    t[1].name = "text", t[1].dtype = DSM_I64, t[1].ndim = 2, t[1].shape[0] = 1, t[1].shape[1] = steps, t[1].data = s->log_text.data();
    dsm_weights w{t, 2};
    save_safetensors(safetensors_path, &w);
  });
}

// ---- TTS audio sink: the consumer of Mimi::decode_step in the TTS server -------------------------------------
// moshi-server/src/tts.rs:512-560 (audio_processing_loop): per generation step the audio tokens [1, codebooks, 1] go
// through decode_step once step_idx >= text_audio_delay_in_tokens + acoustic_delay, and the PCM frame leaves in the
// connection's format (Encoder::encode, tts.rs:252-281: raw little-endian f32 or a MessagePack OutMsg::Audio{pcm});
// :917-929 decodes a finished token matrix chunk by chunk.  The Opus formats need the kaudio encoder (out of scope).
struct dsm_tts_sink {
  MimiImpl* mimi;
  int delay;            // text_audio_delay_in_tokens + acoustic_delay
  int format;           // 0: Pcm, 1: PcmMessagePack
  std::vector<float> pcm;
};

dsm_status dsm_tts_sink_create(dsm_mimi* mimi_decoder, int32_t text_audio_delay_in_tokens, int32_t acoustic_delay, int32_t format,
                               dsm_tts_sink** out) {
  return session_guarded([&] {
    DSM_CHECK(mimi_decoder && out, DSM_ERR_INVALID_ARG, "mimi and out are required");
    *out = nullptr;
    DSM_CHECK(mimi_decoder->impl->B == 1 && mimi_decoder->impl->spec.with_decoder, DSM_ERR_INVALID_ARG,
              "the sink drives a Mimi handle of batch size 1 with a decoder");
    DSM_CHECK(format == 0 || format == 1, DSM_ERR_UNSUPPORTED, "format %d (0 = Pcm, 1 = PcmMessagePack; Opus needs an encoder outside this library)", format);
    std::unique_ptr<dsm_tts_sink> s(new dsm_tts_sink);
    s->mimi = mimi_decoder->impl, s->delay = text_audio_delay_in_tokens + acoustic_delay, s->format = format;
    s->pcm.resize(DSM_FRAME_SIZE);
    *out = s.release();
  });
}
void dsm_tts_sink_destroy(dsm_tts_sink* s) { delete s; }

// AudioMessage::Tokens(Some(codes) | None, _, step_idx): *len = bytes written (0: nothing to send for this step).
dsm_status dsm_tts_sink_push(dsm_tts_sink* s, const uint32_t* codes, int64_t step_idx, uint8_t* buf, size_t cap, size_t* len) {
  return session_guarded([&] {
    DSM_CHECK(s && buf && len, DSM_ERR_INVALID_ARG, "sink, buffer and len are required");
    *len = 0;
    if (!codes || step_idx < s->delay) return;              // tts.rs:536: only logged
    s->mimi->decode_step(codes, nullptr, s->pcm.data());
    if (s->format == 0) {
      const size_t n = (size_t)DSM_FRAME_SIZE * sizeof(float);
      *len = n;
      DSM_CHECK(n <= cap, DSM_ERR_CAPACITY, "frame of %zu bytes, buffer has %zu", n, cap);
      memcpy(buf, s->pcm.data(), n);                        // little-endian f32, as byteorder::LittleEndian writes them
    } else {
      dsm_wire_in m{};
      m.kind = DSM_WIRE_IN_AUDIO, m.pcm = s->pcm.data(), m.n = DSM_FRAME_SIZE;   // same layout as OutMsg::Audio { pcm }
      const dsm_status st = dsm_wire_encode_in(&m, buf, cap, len);
      if (st != DSM_OK) throw Error(st, last_error());
    }
  });
}

// tts.rs:917-929: codes [codebooks, steps] (the reference's [1, cb, T] tensor) -> pcm [steps * 1920], from a fresh state.
dsm_status dsm_tts_sink_decode_all(dsm_tts_sink* s, const uint32_t* codes, int64_t steps, float* pcm) {
  return session_guarded([&] {
    DSM_CHECK(s && codes && pcm && steps >= 0, DSM_ERR_INVALID_ARG, "sink, codes and pcm are required");
    const int cb = s->mimi->spec.n_q;
    s->mimi->reset_state();
    std::vector<uint32_t> frame((size_t)cb);
    for (int64_t t = 0; t < steps; ++t) {
      for (int c = 0; c < cb; ++c) frame[c] = codes[(size_t)c * steps + t];
      s->mimi->decode_step(frame.data(), nullptr, pcm + (size_t)t * DSM_FRAME_SIZE);
    }
  });
}

}  // extern "C"
