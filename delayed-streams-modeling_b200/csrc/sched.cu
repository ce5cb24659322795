// Batched step scheduler: the host side of moshi-server's BatchedAsr (moshi-server/src/batched_asr.rs)
// minus the network front end -- a fixed pool of B slots, a PCM queue per slot, one tick that gathers at
// most one 1920-sample frame per active slot into a pinned [B, 1920] buffer, builds the mask, applies the
// pending per-slot resets, runs the fused ASR step and fans the results back out per slot.
//
//   reference                                   here
//   channels() slot pop / capacity refusal      dsm_sched_open            (batched_asr.rs:795-808, 864-890)
//   InMsg::Init -> resets.push(bid)             open marks the slot for reset_batch_idx (:567-573, :467-471)
//   InMsg::Audio -> Channel::extend_data        dsm_sched_push_audio + the gather in tick (:90-113, :594-600)
//   InMsg::Marker                               dsm_sched_push_marker     (:574-586, fired :703-715)
//   encoder_loop + model_loop + post_process    dsm_sched_tick            (:314-403, 432-522, 655-717)
//   closed out_tx -> slot freed                 dsm_sched_close, reclaimed at the next tick (:633-652)
//
// The reference needs three OS threads because candle serialises everything on one CUDA stream; here the
// encode of frame t already overlaps the LM step on frame t-1 inside dsm_asr_step_pcm's CUDA graph, so a
// tick is one call.  push_* may be called from other threads while a tick runs (per-slot locks).
#include <algorithm>
#include <deque>
#include <mutex>
#include <queue>

#include "handles.cuh"

namespace dsm {

struct SchedSlot {
  std::mutex mu;
  bool open = false, closing = false, init_pending = false;
  uint64_t channel_id = 0;
  std::vector<float> data;      // buffered samples, consumed from `head`
  size_t head = 0;
  size_t steps = 0;
  std::vector<int64_t> new_markers;
  size_t buffered() const { return data.size() - head; }
};

struct SchedMarker {
  uint64_t channel_id;
  int slot;
  long long step_idx;
  int64_t id;
  bool operator<(const SchedMarker& o) const { return step_idx > o.step_idx; }   // min-heap on step_idx
};

struct SchedOut {
  int kind, slot;
  uint64_t channel_id;
  int step_idx;
  double time;
  std::vector<uint32_t> tokens;
  std::vector<float> prs;
  int64_t marker_id;
  long long buffered_pcm;
};

struct SchedImpl {
  AsrImpl* asr;
  int B, delay, nh;
  std::vector<std::unique_ptr<SchedSlot>> slots;
  std::mutex pool_mu;
  std::deque<int> free_idx, active_idx;
  uint64_t next_channel = 1;
  float* batch_pcm = nullptr;      // pinned [B, 1920] (batched_asr.rs:301-306)
  bool pinned = false;
  std::vector<uint8_t> mask;
  std::vector<uint64_t> tick_channel;
  std::priority_queue<SchedMarker> markers;
  std::vector<SchedOut> out;
  long long step_idx = 0;          // model steps run so far
  std::vector<uint32_t> tokens;

  SchedImpl(AsrImpl* a, int batch) : asr(a), B(a ? a->B : batch), delay(a ? a->delay : 0), nh(a ? a->nh : 0) {
    DSM_CHECK(B > 0, DSM_ERR_INVALID_ARG, "scheduler needs a positive batch size");
    for (int i = 0; i < B; ++i) {
      slots.emplace_back(new SchedSlot());
      free_idx.push_back(i);
    }
    const size_t bytes = (size_t)B * DSM_FRAME_SIZE * sizeof(float);
    if (asr && cudaMallocHost(reinterpret_cast<void**>(&batch_pcm), bytes) == cudaSuccess) pinned = true;
    else {
      cudaGetLastError();
      batch_pcm = static_cast<float*>(malloc(bytes));
      DSM_CHECK(batch_pcm != nullptr, DSM_ERR_OOM, "host allocation of %zu bytes failed", bytes);
    }
    memset(batch_pcm, 0, bytes);
    mask.assign(B, 0);
    tick_channel.assign(B, 0);
    tokens.assign(B, 0);
  }
  ~SchedImpl() {
    if (pinned) cudaFreeHost(batch_pcm);
    else free(batch_pcm);
  }

  int open(uint64_t* channel_id) {
    std::lock_guard<std::mutex> g(pool_mu);
    if (free_idx.empty()) fail(DSM_ERR_CAPACITY, "no free slot: all %d streams are in use", B);
    const int s = free_idx.front();
    free_idx.pop_front();
    SchedSlot& sl = *slots[s];
    {
      std::lock_guard<std::mutex> gs(sl.mu);
      sl.open = true, sl.closing = false, sl.init_pending = true;
      sl.channel_id = next_channel++;
      sl.data.clear(), sl.head = 0, sl.steps = 0, sl.new_markers.clear();
      if (channel_id) *channel_id = sl.channel_id;
    }
    active_idx.push_back(s);
    return s;
  }

  SchedSlot& checked(int slot) {
    DSM_CHECK(slot >= 0 && slot < B, DSM_ERR_INVALID_ARG, "slot %d out of range [0, %d)", slot, B);
    return *slots[slot];
  }

  void tick(int* stepped) {
    out.clear();
    if (stepped) *stepped = 0;
    std::vector<int> active;
    {
      std::lock_guard<std::mutex> g(pool_mu);
      active.assign(active_idx.begin(), active_idx.end());
    }
    std::fill(mask.begin(), mask.end(), 0);
    std::fill(tick_channel.begin(), tick_channel.end(), 0);
    std::vector<int> resets;
    bool with_data = false;
    // ---- pre_process (batched_asr.rs:527-653) ----
    for (int s : active) {
      SchedSlot& sl = *slots[s];
      std::lock_guard<std::mutex> g(sl.mu);
      if (!sl.open || sl.closing) continue;
      tick_channel[s] = sl.channel_id;
      if (sl.init_pending) {
        sl.init_pending = false;
        resets.push_back(s);
        out.push_back(SchedOut{DSM_SCHED_READY, s, sl.channel_id, (int)step_idx, 0.0, {}, {}, 0, (long long)sl.buffered()});
      }
      for (int64_t id : sl.new_markers) {
        // fires once the model has consumed what is buffered now plus the text delay (:577-585)
        const long long at = step_idx + delay + (long long)(sl.buffered() / DSM_FRAME_SIZE);
        markers.push(SchedMarker{sl.channel_id, s, at, id});
      }
      sl.new_markers.clear();
      float* dst = batch_pcm + (size_t)s * DSM_FRAME_SIZE;
      if (sl.buffered() >= DSM_FRAME_SIZE) {
        memcpy(dst, sl.data.data() + sl.head, DSM_FRAME_SIZE * sizeof(float));
        sl.head += DSM_FRAME_SIZE;
        if (sl.head >= sl.data.size()) sl.data.clear(), sl.head = 0;
        else if (sl.head > (1u << 20)) {      // compact occasionally
          sl.data.erase(sl.data.begin(), sl.data.begin() + (long)sl.head);
          sl.head = 0;
        }
        sl.steps += 1;
        mask[s] = 1;
        with_data = true;
      } else {
        memset(dst, 0, DSM_FRAME_SIZE * sizeof(float));
      }
    }
    // ---- model_loop (batched_asr.rs:467-506) ----
    for (int s : resets)
      if (asr) asr->reset_batch_idx(s);
    if (with_data) {
      if (asr) asr->step(batch_pcm, nullptr, mask.data(), tokens.data());
      step_idx += 1;
      if (stepped) *stepped = 1;
      // ---- post_process (batched_asr.rs:655-717) ----
      if (asr) {
        for (const AsrMsgStore& m : asr->msgs) {
          if (m.kind == DSM_MSG_STEP) {
            for (int s = 0; s < B; ++s) {
              if (!mask[s]) continue;
              SchedOut o{DSM_SCHED_STEP, s, tick_channel[s], m.step_idx, 0.0, {}, {}, 0, 0};
              o.prs.resize(nh);
              for (int h = 0; h < nh; ++h) o.prs[h] = asr->prs_copy[(size_t)h * B + s];
              {
                std::lock_guard<std::mutex> g(slots[s]->mu);
                o.buffered_pcm = (long long)slots[s]->buffered();
              }
              out.push_back(std::move(o));
            }
          } else {
            SchedOut o{m.kind == DSM_MSG_WORD ? DSM_SCHED_WORD : DSM_SCHED_END_WORD, m.batch_idx, tick_channel[m.batch_idx],
                       (int)step_idx, m.time, m.tokens, {}, 0, 0};
            out.push_back(std::move(o));
          }
        }
      }
    }
    while (!markers.empty() && markers.top().step_idx <= step_idx) {
      const SchedMarker m = markers.top();
      markers.pop();
      // a slot re-used by another connection must not receive the old connection's marker (:110-118)
      if (slots[m.slot]->channel_id == m.channel_id && slots[m.slot]->open)
        out.push_back(SchedOut{DSM_SCHED_MARKER, m.slot, m.channel_id, (int)step_idx, 0.0, {}, {}, m.id, 0});
    }
    // ---- reclaim closed slots (batched_asr.rs:633-652) ----
    {
      std::lock_guard<std::mutex> g(pool_mu);
      for (auto it = active_idx.begin(); it != active_idx.end();) {
        SchedSlot& sl = *slots[*it];
        std::lock_guard<std::mutex> gs(sl.mu);
        if (sl.closing || !sl.open) {
          sl.open = false, sl.closing = false;
          sl.data.clear(), sl.head = 0;
          free_idx.push_back(*it);
          it = active_idx.erase(it);
        } else {
          ++it;
        }
      }
    }
  }
};

}  // namespace dsm

using namespace dsm;

struct dsm_sched { SchedImpl* impl; };

template <typename F>
static dsm_status guarded_s(F&& f) {
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

dsm_status dsm_sched_create(dsm_asr* asr, int32_t batch_size_if_no_asr, dsm_sched** out) {
  return guarded_s([&] {
    DSM_CHECK(out != nullptr, DSM_ERR_INVALID_ARG, "out is required");
    *out = nullptr;
    std::unique_ptr<SchedImpl> impl(new SchedImpl(asr ? asr->impl : nullptr, batch_size_if_no_asr));
    *out = new dsm_sched{impl.release()};
  });
}
void dsm_sched_destroy(dsm_sched* h) {
  if (!h) return;
  delete h->impl;
  delete h;
}
dsm_status dsm_sched_open(dsm_sched* h, int32_t* slot, uint64_t* channel_id) {
  return guarded_s([&] {
    DSM_CHECK(h && slot, DSM_ERR_INVALID_ARG, "handle and slot are required");
    *slot = h->impl->open(channel_id);
  });
}
dsm_status dsm_sched_close(dsm_sched* h, int32_t slot) {
  return guarded_s([&] {
    DSM_CHECK(h, DSM_ERR_INVALID_ARG, "handle is required");
    SchedSlot& sl = h->impl->checked(slot);
    std::lock_guard<std::mutex> g(sl.mu);
    DSM_CHECK(sl.open, DSM_ERR_INVALID_ARG, "slot %d is not open", slot);
    sl.closing = true;
  });
}
dsm_status dsm_sched_push_audio(dsm_sched* h, int32_t slot, const float* pcm, int64_t n) {
  return guarded_s([&] {
    DSM_CHECK(h && (pcm || n == 0) && n >= 0, DSM_ERR_INVALID_ARG, "handle and pcm are required");
    SchedSlot& sl = h->impl->checked(slot);
    std::lock_guard<std::mutex> g(sl.mu);
    DSM_CHECK(sl.open && !sl.closing, DSM_ERR_INVALID_ARG, "slot %d is not open", slot);
    sl.data.insert(sl.data.end(), pcm, pcm + n);
  });
}
dsm_status dsm_sched_push_marker(dsm_sched* h, int32_t slot, int64_t id) {
  return guarded_s([&] {
    DSM_CHECK(h, DSM_ERR_INVALID_ARG, "handle is required");
    SchedSlot& sl = h->impl->checked(slot);
    std::lock_guard<std::mutex> g(sl.mu);
    DSM_CHECK(sl.open && !sl.closing, DSM_ERR_INVALID_ARG, "slot %d is not open", slot);
    sl.new_markers.push_back(id);
  });
}
dsm_status dsm_sched_tick(dsm_sched* h, int32_t* stepped) {
  return guarded_s([&] {
    DSM_CHECK(h, DSM_ERR_INVALID_ARG, "handle is required");
    h->impl->tick(stepped);
  });
}
int32_t dsm_sched_out_count(const dsm_sched* h) { return h ? (int32_t)h->impl->out.size() : 0; }
dsm_status dsm_sched_out_get(const dsm_sched* h, int32_t i, dsm_sched_out* o) {
  return guarded_s([&] {
    DSM_CHECK(h && o, DSM_ERR_INVALID_ARG, "handle and out are required");
    DSM_CHECK(i >= 0 && i < (int32_t)h->impl->out.size(), DSM_ERR_INVALID_ARG, "output index %d out of range", i);
    const SchedOut& s = h->impl->out[i];
    memset(o, 0, sizeof *o);
    o->kind = s.kind, o->slot = s.slot, o->channel_id = s.channel_id, o->step_idx = s.step_idx, o->time = s.time;
    o->n_tokens = (int32_t)s.tokens.size(), o->tokens = s.tokens.empty() ? nullptr : s.tokens.data();
    o->n_prs = (int32_t)s.prs.size(), o->prs = s.prs.empty() ? nullptr : s.prs.data();
    o->marker_id = s.marker_id, o->buffered_pcm = s.buffered_pcm;
  });
}
int32_t dsm_sched_free_slots(const dsm_sched* h) {
  if (!h) return 0;
  std::lock_guard<std::mutex> g(h->impl->pool_mu);
  return (int32_t)h->impl->free_idx.size();
}
int64_t dsm_sched_step_idx(const dsm_sched* h) { return h ? h->impl->step_idx : 0; }
const uint8_t* dsm_sched_last_mask(const dsm_sched* h) { return h ? h->impl->mask.data() : nullptr; }
const float* dsm_sched_batch_pcm(const dsm_sched* h) { return h ? h->impl->batch_pcm : nullptr; }

}  // extern "C"
