// Batched step scheduler / slot router: the host side of moshi-server's BatchedAsr
// (moshi-server/src/batched_asr.rs) minus the network front end, for ONE OR SEVERAL GPUs.
//
//   reference                                   here
//   channels() slot pop / capacity refusal      dsm_sched_open            (batched_asr.rs:795-808, 864-890)
//   InMsg::Init -> resets.push(bid)             open marks the slot for reset_batch_idx (:567-573, :467-471)
//   InMsg::Audio -> Channel::extend_data        dsm_sched_push_audio + the gather of a tick (:90-113, :594-600)
//   InMsg::Marker                               dsm_sched_push_marker     (:574-586, fired :703-715)
//   encoder_loop + model_loop + post_process    dsm_sched_tick / the per-device loop of dsm_sched_start (:314-522, 655-717)
//   closed out_tx -> slot freed                 dsm_sched_close, reclaimed by the next tick (:633-652)
//
// Layout.  Device g owns slots [g*B, (g+1)*B) with its own asr::State (weights replica, KV rings, frame graph):
// streams never interact, so routing is `device = slot / B` and no step-path exchange exists.  The
// reference funnels every connection through per-slot mutexes and one rayon gather (16,000 locks per tick at
// 8 x 2,000 slots); here the per-tick path takes no lock at all:
//   * each slot has a single-producer / single-consumer PCM ring (producer: the connection's thread, consumer: its
//     device's tick), a small SPSC marker queue and, in threaded mode, an SPSC queue of outgoing events;
//   * slot state is one atomic word; only open() / reclaim touch the per-device free list (a mutex, off the tick path);
//   * a device's loop (dsm_sched_start: one thread per device) is pipelined: while the GPU runs the step of frame t the
//     thread gathers frame t+1 into the OTHER half of a double-buffered pinned batch, then collects step t
//     (AsrImpl::step_begin / step_end), so H2D, gather and fan-out overlap the device work.
// One deviation, on purpose: the reference's gather overwrites the batch row when a connection has several complete
// frames queued in one tick and drops the earlier ones (extend_data called once per queued Audio message,
// batched_asr.rs:594-600 with :90-113); here a slot contributes exactly one frame per tick and nothing is dropped.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <deque>
#include <memory>
#include <mutex>
#include <queue>
#include <thread>

#include "handles.cuh"

namespace dsm {

// Lock-free single-producer / single-consumer ring (power-of-two capacity).
template <typename T>
class SpscRing {
 public:
  explicit SpscRing(size_t capacity_pow2) : buf_(capacity_pow2), mask_(capacity_pow2 - 1) {}
  size_t capacity() const { return buf_.size(); }
  size_t size() const { return tail_.load(std::memory_order_acquire) - head_.load(std::memory_order_acquire); }
  // producer
  bool push(const T* src, size_t n) {
    const size_t t = tail_.load(std::memory_order_relaxed), h = head_.load(std::memory_order_acquire);
    if (buf_.size() - (t - h) < n) return false;
    for (size_t i = 0; i < n; ++i) buf_[(t + i) & mask_] = src[i];
    tail_.store(t + n, std::memory_order_release);
    return true;
  }
  bool push(const T& v) { return push(&v, 1); }
  // consumer
  bool pop(T* dst, size_t n) {
    const size_t h = head_.load(std::memory_order_relaxed), t = tail_.load(std::memory_order_acquire);
    if (t - h < n) return false;
    const size_t first = std::min(n, buf_.size() - (h & mask_));
    std::copy(buf_.begin() + (h & mask_), buf_.begin() + (h & mask_) + first, dst);
    std::copy(buf_.begin(), buf_.begin() + (n - first), dst + first);
    head_.store(h + n, std::memory_order_release);
    return true;
  }
  bool pop(T& v) { return pop(&v, 1); }
  void clear_consumer() { head_.store(tail_.load(std::memory_order_acquire), std::memory_order_release); }
  uint64_t total_pushed() const { return tail_.load(std::memory_order_acquire); }
  uint64_t total_popped() const { return head_.load(std::memory_order_acquire); }

 private:
  std::vector<T> buf_;
  size_t mask_;
  alignas(64) std::atomic<size_t> head_{0};
  alignas(64) std::atomic<size_t> tail_{0};
};

enum SlotState : uint32_t { SLOT_FREE = 0, SLOT_OPEN = 1, SLOT_CLOSING = 2 };

struct MarkerIn {
  int64_t id;
  uint64_t pushed_at;   // samples pushed on the connection before the marker (its place in the message order)
};

constexpr int kEventTokens = 64, kEventPrs = 8;
struct Event {          // one outgoing message of a connection (threaded mode)
  int32_t kind, slot, step_idx, n_tokens, n_prs;
  uint64_t channel_id;
  double time;
  int64_t marker_id, buffered_pcm;
  uint32_t tokens[kEventTokens];
  float prs[kEventPrs];
};

struct SchedSlot {
  std::atomic<uint32_t> state{SLOT_FREE};
  std::atomic<uint64_t> channel_id{0};
  std::atomic<uint32_t> init_pending{0};
  SpscRing<float> pcm;
  SpscRing<MarkerIn> markers;
  SpscRing<Event> events;
  size_t steps = 0;          // tick thread only
  SchedSlot(size_t pcm_cap, size_t ev_cap) : pcm(pcm_cap), markers(64), events(ev_cap) {}
  size_t buffered() const { return pcm.size(); }
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

struct DeviceStats {
  std::atomic<long long> ticks{0}, steps{0}, idle_sleeps{0}, frames{0};
  std::atomic<long long> gather_ns{0}, step_wait_ns{0}, fanout_ns{0};
};

static long long now_ns() {
  return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Everything one device's tick touches.
struct SchedDevice {
  AsrImpl* asr = nullptr;
  int index = 0, B = 0, delay = 0, nh = 0;
  std::mutex pool_mu;                  // open / reclaim only
  std::deque<int> free_local;          // local slot indices
  float* batch[2] = {nullptr, nullptr};   // pinned [B, 1920] x 2 (batched_asr.rs:301-306), double-buffered
  bool pinned = false;
  int cur = 0;
  std::vector<uint8_t> mask[2];
  std::vector<uint64_t> tick_channel[2];
  std::priority_queue<SchedMarker> markers;
  std::vector<uint32_t> tokens;
  long long step_idx = 0;
  DeviceStats stats;
  std::thread worker;
};

struct Gathered {
  bool with_data = false;
  int buf = 0;
  std::vector<int> resets;
};

struct SchedImpl {
  int n_dev, B;
  std::vector<std::unique_ptr<SchedDevice>> dev;
  std::vector<std::unique_ptr<SchedSlot>> slots;   // global index = device * B + local
  std::atomic<uint64_t> next_channel{1};
  std::vector<SchedOut> out;                       // outputs of the last dsm_sched_tick (all devices)
  std::atomic<bool> running{false};
  int dry_step_us = 0;                             // dry mode: pretend a step takes this long

  SchedImpl(AsrImpl* const* asrs, int n, int batch_if_none) : n_dev(n) {
    DSM_CHECK(n >= 1 && n <= 64, DSM_ERR_INVALID_ARG, "scheduler over %d devices", n);
    B = (asrs && asrs[0]) ? asrs[0]->B : batch_if_none;
    DSM_CHECK(B > 0, DSM_ERR_INVALID_ARG, "scheduler needs a positive batch size");
    for (int g = 0; g < n; ++g) {
      std::unique_ptr<SchedDevice> d(new SchedDevice);
      d->asr = asrs ? asrs[g] : nullptr;
      DSM_CHECK(!asrs || (asrs[g] != nullptr) == (asrs[0] != nullptr), DSM_ERR_INVALID_ARG, "either every device has an asr state or none");
      if (d->asr) DSM_CHECK(d->asr->B == B, DSM_ERR_SHAPE, "device %d has batch %d, device 0 has %d", g, d->asr->B, B);
      d->index = g, d->B = B, d->delay = d->asr ? d->asr->delay : 0, d->nh = d->asr ? d->asr->nh : 0;
      DSM_CHECK(d->nh <= kEventPrs, DSM_ERR_UNSUPPORTED, "%d extra heads (events carry up to %d)", d->nh, kEventPrs);
      const size_t bytes = (size_t)B * DSM_FRAME_SIZE * sizeof(float);
      if (d->asr) DSM_CUDA(cudaSetDevice(d->asr->lm->device));
      for (int k = 0; k < 2; ++k) {
        if (d->asr && cudaMallocHost(reinterpret_cast<void**>(&d->batch[k]), bytes) == cudaSuccess) d->pinned = true;
        else {
          cudaGetLastError();
          d->batch[k] = static_cast<float*>(malloc(bytes));
          DSM_CHECK(d->batch[k] != nullptr, DSM_ERR_OOM, "host allocation of %zu bytes failed", bytes);
        }
        memset(d->batch[k], 0, bytes);
        d->mask[k].assign(B, 0);
        d->tick_channel[k].assign(B, 0);
      }
      d->tokens.assign(B, 0);
      for (int i = 0; i < B; ++i) d->free_local.push_back(i);
      dev.push_back(std::move(d));
    }
    // 2^17 samples = 5.5 s of audio per connection; 256 queued events
    for (int i = 0; i < n * B; ++i) slots.emplace_back(new SchedSlot(1u << 17, 256));
  }
  ~SchedImpl() {
    stop();
    for (auto& d : dev)
      for (int k = 0; k < 2; ++k) {
        if (!d->batch[k]) continue;
        if (d->pinned) cudaFreeHost(d->batch[k]);
        else free(d->batch[k]);
      }
  }

  // BatchedAsr::channels: a free slot on the least-loaded device, or DSM_ERR_CAPACITY
  int open(uint64_t* channel_id) {
    int best = -1;
    size_t best_free = 0;
    for (int g = 0; g < n_dev; ++g) {
      std::lock_guard<std::mutex> l(dev[g]->pool_mu);
      if (dev[g]->free_local.size() > best_free) best_free = dev[g]->free_local.size(), best = g;
    }
    if (best < 0) fail(DSM_ERR_CAPACITY, "no free slot: all %d streams are in use", n_dev * B);
    int local;
    {
      std::lock_guard<std::mutex> l(dev[best]->pool_mu);
      if (dev[best]->free_local.empty()) fail(DSM_ERR_CAPACITY, "no free slot: all %d streams are in use", n_dev * B);
      local = dev[best]->free_local.front();
      dev[best]->free_local.pop_front();
    }
    const int s = best * B + local;
    SchedSlot& sl = *slots[s];
    // the slot is FREE: no producer, and the tick ignores it, so its rings can be emptied from here
    sl.pcm.clear_consumer(), sl.markers.clear_consumer(), sl.events.clear_consumer();
    const uint64_t cid = next_channel.fetch_add(1);
    sl.channel_id.store(cid, std::memory_order_relaxed);
    sl.init_pending.store(1, std::memory_order_relaxed);
    sl.steps = 0;
    sl.state.store(SLOT_OPEN, std::memory_order_release);
    if (channel_id) *channel_id = cid;
    return s;
  }

  SchedSlot& checked(int slot) {
    DSM_CHECK(slot >= 0 && slot < n_dev * B, DSM_ERR_INVALID_ARG, "slot %d out of range [0, %d)", slot, n_dev * B);
    return *slots[slot];
  }
  // a stale client (its connection closed, the slot re-opened by someone else) must not reach the new connection
  SchedSlot& checked_open(int slot, uint64_t channel_id) {
    SchedSlot& sl = checked(slot);
    DSM_CHECK(sl.state.load(std::memory_order_acquire) == SLOT_OPEN, DSM_ERR_INVALID_ARG, "slot %d is not open", slot);
    DSM_CHECK(channel_id == 0 || sl.channel_id.load(std::memory_order_relaxed) == channel_id, DSM_ERR_INVALID_ARG,
              "slot %d now belongs to another connection", slot);
    return sl;
  }

  // ---- pre_process (batched_asr.rs:527-653): one frame per slot with data into batch[buf], mask, resets, markers ----
  // in_flight: steps begun but not yet counted in d.step_idx (the pipelined loop gathers frame t+1 before it collects
  // step t); frames they consumed are already in total_popped(), so a marker's due step must count them as well.
  Gathered gather(SchedDevice& d, std::vector<SchedOut>* outs, bool to_queues, int in_flight = 0) {
    const long long t0 = now_ns();
    Gathered g;
    g.buf = d.cur;
    float* batch = d.batch[g.buf];
    std::vector<uint8_t>& mask = d.mask[g.buf];
    std::vector<uint64_t>& chan = d.tick_channel[g.buf];
    long long frames = 0;
    for (int i = 0; i < d.B; ++i) {
      SchedSlot& sl = *slots[d.index * d.B + i];
      mask[i] = 0, chan[i] = 0;
      if (sl.state.load(std::memory_order_acquire) != SLOT_OPEN) continue;
      const uint64_t cid = sl.channel_id.load(std::memory_order_relaxed);
      chan[i] = cid;
      if (sl.init_pending.exchange(0, std::memory_order_acq_rel)) {
        g.resets.push_back(i);
        emit(d, SchedOut{DSM_SCHED_READY, d.index * d.B + i, cid, (int)d.step_idx, 0.0, {}, {}, 0, (long long)sl.buffered()}, outs,
             to_queues);
      }
      MarkerIn m;
      while (sl.markers.pop(m)) {
        // fires once the model has consumed what was queued ahead of the marker plus the text delay (:577-585)
        const uint64_t consumed = sl.pcm.total_popped();
        const long long ahead = m.pushed_at > consumed ? (long long)((m.pushed_at - consumed) / DSM_FRAME_SIZE) : 0;
        d.markers.push(SchedMarker{cid, d.index * d.B + i, d.step_idx + in_flight + d.delay + ahead, m.id});
      }
      float* dst = batch + (size_t)i * DSM_FRAME_SIZE;
      if (sl.pcm.pop(dst, DSM_FRAME_SIZE)) {
        sl.steps += 1;
        mask[i] = 1;
        g.with_data = true;
        ++frames;
      } else {
        memset(dst, 0, DSM_FRAME_SIZE * sizeof(float));
      }
    }
    d.stats.gather_ns += now_ns() - t0;
    d.stats.frames += frames;
    return g;
  }

  void emit(SchedDevice& d, SchedOut&& o, std::vector<SchedOut>* outs, bool to_queues) {
    if (to_queues) {
      SchedSlot& sl = *slots[o.slot];
      Event e{};
      e.kind = o.kind, e.slot = o.slot, e.step_idx = o.step_idx, e.channel_id = o.channel_id, e.time = o.time;
      e.marker_id = o.marker_id, e.buffered_pcm = o.buffered_pcm;
      e.n_tokens = (int32_t)std::min<size_t>(o.tokens.size(), kEventTokens);
      std::copy(o.tokens.begin(), o.tokens.begin() + e.n_tokens, e.tokens);
      e.n_prs = (int32_t)std::min<size_t>(o.prs.size(), kEventPrs);
      std::copy(o.prs.begin(), o.prs.begin() + e.n_prs, e.prs);
      // the reference's channel send fails when the client is gone; a full queue means the client stopped reading
      if (!sl.events.push(e)) sl.state.store(SLOT_CLOSING, std::memory_order_release);
    }
    if (outs) outs->push_back(std::move(o));
    (void)d;
  }

  // ---- post_process (batched_asr.rs:655-717) for the step that used batch[buf] ----
  void fanout(SchedDevice& d, int buf, bool stepped, std::vector<SchedOut>* outs, bool to_queues) {
    const long long t0 = now_ns();
    const std::vector<uint8_t>& mask = d.mask[buf];
    const std::vector<uint64_t>& chan = d.tick_channel[buf];
    if (stepped && d.asr) {
      for (const AsrMsgStore& m : d.asr->msgs) {
        if (m.kind == DSM_MSG_STEP) {
          for (int i = 0; i < d.B; ++i) {
            if (!mask[i]) continue;
            SchedOut o{DSM_SCHED_STEP, d.index * d.B + i, chan[i], m.step_idx, 0.0, {}, {}, 0, 0};
            o.prs.resize(d.nh);
            for (int h = 0; h < d.nh; ++h) o.prs[h] = d.asr->prs_copy[(size_t)h * d.B + i];
            o.buffered_pcm = (long long)slots[d.index * d.B + i]->buffered();
            emit(d, std::move(o), outs, to_queues);
          }
        } else {
          emit(d, SchedOut{m.kind == DSM_MSG_WORD ? DSM_SCHED_WORD : DSM_SCHED_END_WORD, d.index * d.B + m.batch_idx, chan[m.batch_idx],
                           (int)d.step_idx, m.time, m.tokens, {}, 0, 0}, outs, to_queues);
        }
      }
    }
    while (!d.markers.empty() && d.markers.top().step_idx <= d.step_idx) {
      const SchedMarker m = d.markers.top();
      d.markers.pop();
      // a slot re-used by another connection must not receive the old connection's marker (:110-118)
      SchedSlot& sl = *slots[m.slot];
      if (sl.channel_id.load(std::memory_order_relaxed) == m.channel_id && sl.state.load(std::memory_order_acquire) == SLOT_OPEN)
        emit(d, SchedOut{DSM_SCHED_MARKER, m.slot, m.channel_id, (int)d.step_idx, 0.0, {}, {}, m.id, 0}, outs, to_queues);
    }
    // ---- reclaim closed slots (batched_asr.rs:633-652) ----
    for (int i = 0; i < d.B; ++i) {
      SchedSlot& sl = *slots[d.index * d.B + i];
      if (sl.state.load(std::memory_order_acquire) != SLOT_CLOSING) continue;
      sl.state.store(SLOT_FREE, std::memory_order_release);
      std::lock_guard<std::mutex> l(d.pool_mu);
      d.free_local.push_back(i);
    }
    d.stats.fanout_ns += now_ns() - t0;
  }

  void apply_resets(SchedDevice& d, const Gathered& g) {
    for (int i : g.resets)
      if (d.asr) d.asr->reset_batch_idx(i);
  }

  // One synchronous pass over a device (the form the tests drive): gather, reset, step, fan out.
  bool tick_device(SchedDevice& d, std::vector<SchedOut>* outs) {
    d.stats.ticks += 1;
    Gathered g = gather(d, outs, false);
    apply_resets(d, g);
    if (g.with_data) {
      const long long t0 = now_ns();
      if (d.asr) d.asr->step(d.batch[g.buf], nullptr, d.mask[g.buf].data(), d.tokens.data());
      d.stats.step_wait_ns += now_ns() - t0;
      d.step_idx += 1;
      d.stats.steps += 1;
    }
    fanout(d, g.buf, g.with_data, outs, false);
    return g.with_data;
  }

  // The pipelined loop of one device's thread: the gather of frame t+1 overlaps the device step of frame t.
  void run_device(SchedDevice& d) {
    if (d.asr) cudaSetDevice(d.asr->lm->device);
    bool pending = false;
    int pending_buf = 0;
    try {
      while (running.load(std::memory_order_acquire)) {
        d.stats.ticks += 1;
        Gathered g = gather(d, nullptr, true, pending ? 1 : 0);   // into batch[d.cur], which no step in flight reads
        if (pending) {
          const long long t0 = now_ns();
          if (d.asr) d.asr->step_end(d.tokens.data());
          else if (dry_step_us > 0) std::this_thread::sleep_for(std::chrono::microseconds(dry_step_us));
          d.stats.step_wait_ns += now_ns() - t0;
          d.step_idx += 1;
          d.stats.steps += 1;
          fanout(d, pending_buf, true, nullptr, true);
          pending = false;
        } else {
          fanout(d, g.buf, false, nullptr, true);       // markers / reclaim still progress on idle ticks
        }
        apply_resets(d, g);
        if (g.with_data) {
          if (d.asr) d.asr->step_begin(d.batch[g.buf], nullptr, d.mask[g.buf].data());
          pending = true, pending_buf = g.buf;
          d.cur ^= 1;
        } else {
          d.stats.idle_sleeps += 1;
          std::this_thread::sleep_for(std::chrono::milliseconds(2));   // batched_asr.rs:398-400
        }
      }
      if (pending) {
        if (d.asr) d.asr->step_end(d.tokens.data());
        d.step_idx += 1;
        fanout(d, pending_buf, true, nullptr, true);
      }
    } catch (const std::exception& e) {
      set_last_error(e.what());
      fprintf(stderr, "dsm_sched: device %d loop stopped: %s\n", d.index, e.what());
    }
  }

  void start() {
    DSM_CHECK(!running.load(), DSM_ERR_INVALID_ARG, "scheduler threads are already running");
    running.store(true, std::memory_order_release);
    for (auto& d : dev) {
      SchedDevice* dp = d.get();
      d->worker = std::thread([this, dp] { run_device(*dp); });
    }
  }
  void stop() {
    if (!running.exchange(false)) return;
    for (auto& d : dev)
      if (d->worker.joinable()) d->worker.join();
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
    AsrImpl* a = asr ? asr->impl : nullptr;
    std::unique_ptr<SchedImpl> impl(new SchedImpl(&a, 1, batch_size_if_no_asr));
    *out = new dsm_sched{impl.release()};
  });
}
dsm_status dsm_sched_create_multi(dsm_asr* const* asrs, int32_t n_devices, int32_t batch_size_if_no_asr, dsm_sched** out) {
  return guarded_s([&] {
    DSM_CHECK(out != nullptr && n_devices >= 1, DSM_ERR_INVALID_ARG, "out and n_devices >= 1 are required");
    *out = nullptr;
    std::vector<AsrImpl*> v((size_t)n_devices, nullptr);
    if (asrs)
      for (int i = 0; i < n_devices; ++i) v[i] = asrs[i] ? asrs[i]->impl : nullptr;
    std::unique_ptr<SchedImpl> impl(new SchedImpl(v.data(), n_devices, batch_size_if_no_asr));
    *out = new dsm_sched{impl.release()};
  });
}
void dsm_sched_destroy(dsm_sched* h) {
  if (!h) return;
  delete h->impl;
  delete h;
}
int32_t dsm_sched_num_devices(const dsm_sched* h) { return h ? h->impl->n_dev : 0; }
int32_t dsm_sched_slots_per_device(const dsm_sched* h) { return h ? h->impl->B : 0; }
dsm_status dsm_sched_open(dsm_sched* h, int32_t* slot, uint64_t* channel_id) {
  return guarded_s([&] {
    DSM_CHECK(h && slot, DSM_ERR_INVALID_ARG, "handle and slot are required");
    *slot = h->impl->open(channel_id);
  });
}
dsm_status dsm_sched_close_ch(dsm_sched* h, int32_t slot, uint64_t channel_id) {
  return guarded_s([&] {
    DSM_CHECK(h, DSM_ERR_INVALID_ARG, "handle is required");
    SchedSlot& sl = h->impl->checked_open(slot, channel_id);
    sl.state.store(SLOT_CLOSING, std::memory_order_release);
  });
}
dsm_status dsm_sched_close(dsm_sched* h, int32_t slot) { return dsm_sched_close_ch(h, slot, 0); }
dsm_status dsm_sched_push_audio_ch(dsm_sched* h, int32_t slot, uint64_t channel_id, const float* pcm, int64_t n) {
  return guarded_s([&] {
    DSM_CHECK(h && (pcm || n == 0) && n >= 0, DSM_ERR_INVALID_ARG, "handle and pcm are required");
    SchedSlot& sl = h->impl->checked_open(slot, channel_id);
    DSM_CHECK(sl.pcm.push(pcm, (size_t)n), DSM_ERR_CAPACITY, "slot %d: %lld samples do not fit its %zu-sample queue (%zu buffered)", slot,
              (long long)n, sl.pcm.capacity(), sl.pcm.size());
  });
}
dsm_status dsm_sched_push_audio(dsm_sched* h, int32_t slot, const float* pcm, int64_t n) {
  return dsm_sched_push_audio_ch(h, slot, 0, pcm, n);
}
dsm_status dsm_sched_push_marker_ch(dsm_sched* h, int32_t slot, uint64_t channel_id, int64_t id) {
  return guarded_s([&] {
    DSM_CHECK(h, DSM_ERR_INVALID_ARG, "handle is required");
    SchedSlot& sl = h->impl->checked_open(slot, channel_id);
    DSM_CHECK(sl.markers.push(MarkerIn{id, sl.pcm.total_pushed()}), DSM_ERR_CAPACITY, "slot %d: too many pending markers", slot);
  });
}
dsm_status dsm_sched_push_marker(dsm_sched* h, int32_t slot, int64_t id) { return dsm_sched_push_marker_ch(h, slot, 0, id); }
dsm_status dsm_sched_tick(dsm_sched* h, int32_t* stepped) {
  return guarded_s([&] {
    DSM_CHECK(h, DSM_ERR_INVALID_ARG, "handle is required");
    DSM_CHECK(!h->impl->running.load(), DSM_ERR_INVALID_ARG, "dsm_sched_tick while the scheduler threads run");
    h->impl->out.clear();
    bool any = false;
    for (auto& d : h->impl->dev) any |= h->impl->tick_device(*d, &h->impl->out);
    if (stepped) *stepped = any ? 1 : 0;
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
dsm_status dsm_sched_start(dsm_sched* h) {
  return guarded_s([&] {
    DSM_CHECK(h, DSM_ERR_INVALID_ARG, "handle is required");
    h->impl->start();
  });
}
dsm_status dsm_sched_stop(dsm_sched* h) {
  return guarded_s([&] {
    DSM_CHECK(h, DSM_ERR_INVALID_ARG, "handle is required");
    h->impl->stop();
  });
}
dsm_status dsm_sched_poll(dsm_sched* h, int32_t slot, dsm_sched_event* ev, int32_t* got) {
  return guarded_s([&] {
    DSM_CHECK(h && ev && got, DSM_ERR_INVALID_ARG, "handle, event and got are required");
    SchedSlot& sl = h->impl->checked(slot);
    Event e;
    *got = sl.events.pop(e) ? 1 : 0;
    if (!*got) return;
    static_assert(sizeof(dsm_sched_event) == sizeof(Event), "dsm_sched_event must mirror Event");
    memcpy(ev, &e, sizeof e);
  });
}
// One connection's wire traffic (batched_asr.rs:894-949 recv loop, :950-985 send loop): MessagePack InMsg bytes in,
// MessagePack OutMsg bytes out.  Init re-arms the slot reset + Ready (the server itself sends one at open, :894);
// OggOpus is refused: the Ogg/Opus decoder (kaudio) is outside this library -- decode to PCM and send Audio.
dsm_status dsm_sched_push_msg(dsm_sched* h, int32_t slot, uint64_t channel_id, const uint8_t* bytes, size_t len) {
  return guarded_s([&] {
    DSM_CHECK(h && bytes, DSM_ERR_INVALID_ARG, "handle and bytes are required");
    SchedSlot& sl = h->impl->checked_open(slot, channel_id);
    thread_local std::vector<float> pcm;
    pcm.resize(len / 5 + 16);                       // an f32 sample takes 5 bytes on the wire
    thread_local std::vector<uint8_t> data;
    data.resize(len + 16);
    dsm_wire_in m;
    const dsm_status st = dsm_wire_decode_in(bytes, len, &m, pcm.data(), (int64_t)pcm.size(), data.data(), (int64_t)data.size());
    if (st != DSM_OK) throw Error(st, last_error());
    switch (m.kind) {
      case DSM_WIRE_IN_INIT: sl.init_pending.store(1, std::memory_order_release); break;
      case DSM_WIRE_IN_PING: break;
      case DSM_WIRE_IN_MARKER:
        DSM_CHECK(sl.markers.push(MarkerIn{m.id, sl.pcm.total_pushed()}), DSM_ERR_CAPACITY, "slot %d: too many pending markers", slot);
        break;
      case DSM_WIRE_IN_AUDIO:
        DSM_CHECK(sl.pcm.push(m.pcm, (size_t)m.n), DSM_ERR_CAPACITY, "slot %d: %lld samples do not fit its queue", slot, (long long)m.n);
        break;
      default: fail(DSM_ERR_UNSUPPORTED, "OggOpus payloads must be decoded to PCM before they reach the scheduler");
    }
  });
}
// Pops the connection's next event as OutMsg bytes; Word text comes from the SentencePiece model (batched_asr.rs:667)
// or, without one, is the decimal token ids joined by spaces.  *got = 0 when nothing is queued.
dsm_status dsm_sched_poll_msg(dsm_sched* h, int32_t slot, dsm_spm* spm, uint8_t* buf, size_t cap, size_t* len, int32_t* got) {
  return guarded_s([&] {
    DSM_CHECK(h && buf && len && got, DSM_ERR_INVALID_ARG, "handle, buffer, len and got are required");
    SchedSlot& sl = h->impl->checked(slot);
    Event e;
    *got = sl.events.pop(e) ? 1 : 0;
    *len = 0;
    if (!*got) return;
    dsm_wire_out m{};
    std::string text;
    switch (e.kind) {
      case DSM_SCHED_READY: m.kind = DSM_WIRE_OUT_READY; break;
      case DSM_SCHED_MARKER: m.kind = DSM_WIRE_OUT_MARKER, m.id = e.marker_id; break;
      case DSM_SCHED_END_WORD: m.kind = DSM_WIRE_OUT_END_WORD, m.time = e.time; break;
      case DSM_SCHED_STEP:
        m.kind = DSM_WIRE_OUT_STEP, m.step_idx = e.step_idx, m.prs = e.prs, m.n_prs = e.n_prs, m.buffered_pcm = e.buffered_pcm;
        break;
      case DSM_SCHED_WORD: {
        m.kind = DSM_WIRE_OUT_WORD, m.time = e.time;
        if (spm) {
          char tmp[1024];
          size_t n = 0;
          const dsm_status st = dsm_spm_decode(spm, e.tokens, e.n_tokens, tmp, sizeof tmp, &n);
          if (st != DSM_OK) throw Error(st, last_error());
          text.assign(tmp, n);
        } else {
          for (int i = 0; i < e.n_tokens; ++i) text += (i ? " " : "") + std::to_string(e.tokens[i]);
        }
        m.text = text.data(), m.text_len = (int32_t)text.size();
        break;
      }
      default: fail(DSM_ERR_INVALID_ARG, "unknown event kind %d", e.kind);
    }
    const dsm_status st = dsm_wire_encode_out(&m, buf, cap, len);
    if (st != DSM_OK) throw Error(st, last_error());
  });
}
dsm_status dsm_sched_get_stats(const dsm_sched* h, int32_t device, dsm_sched_stats* st) {
  return guarded_s([&] {
    DSM_CHECK(h && st && device >= 0 && device < h->impl->n_dev, DSM_ERR_INVALID_ARG, "handle, stats and a valid device are required");
    const DeviceStats& s = h->impl->dev[device]->stats;
    st->ticks = s.ticks, st->steps = s.steps, st->idle_sleeps = s.idle_sleeps, st->frames = s.frames;
    st->gather_ns = s.gather_ns, st->step_wait_ns = s.step_wait_ns, st->fanout_ns = s.fanout_ns;
  });
}
void dsm_sched_set_dry_step_us(dsm_sched* h, int32_t us) {
  if (h) h->impl->dry_step_us = us;
}
int32_t dsm_sched_free_slots(const dsm_sched* h) {
  if (!h) return 0;
  size_t n = 0;
  for (auto& d : h->impl->dev) {
    std::lock_guard<std::mutex> g(d->pool_mu);
    n += d->free_local.size();
  }
  return (int32_t)n;
}
int64_t dsm_sched_step_idx(const dsm_sched* h) { return h ? h->impl->dev[0]->step_idx : 0; }
int64_t dsm_sched_device_step_idx(const dsm_sched* h, int32_t device) {
  return (h && device >= 0 && device < h->impl->n_dev) ? h->impl->dev[device]->step_idx : -1;
}
// The mask [B] and the gathered batch [B,1920] of device 0's last tick (host memory; for tests).
const uint8_t* dsm_sched_last_mask(const dsm_sched* h) { return h ? h->impl->dev[0]->mask[h->impl->dev[0]->cur].data() : nullptr; }
const float* dsm_sched_batch_pcm(const dsm_sched* h) { return h ? h->impl->dev[0]->batch[h->impl->dev[0]->cur] : nullptr; }

}  // extern "C"
