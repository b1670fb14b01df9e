// grl_experience.hpp — C++ mirror of the reference's `internal/experience` collection path over the device's read-outs.
//
//   experience::Experience        the fields of experiencepb.Experience (proto/experience/v1/experience.proto:25-60)
//   experience::Buffer            internal/experience/buffer.go:21-296 (ring buffer: the oldest record is dropped when full)
//   experience::SimpleCollector   internal/experience/collector.go:13-140 (one record per player that submitted a move)
//
// The reference serialises both states, computes the reward and the action mask on the host, six board scans per
// player (serializer.go, rewards.go).  Here every one of those tensors is a read-out of the turn kernel
// (game::Engine::LastTransition); the collector only frames them.  Header-only.
#pragma once

#include <cstdint>
#include <map>
#include <mutex>
#include <random>
#include <string>
#include <vector>

#include "grl_engine.hpp"

namespace grl {
namespace experience {

inline const core::Sentinel ErrBufferFull{"experience buffer is full", 0};      // buffer.go:13-14
inline const core::Sentinel ErrBufferClosed{"experience buffer is closed", 0};  // buffer.go:15-16

constexpr int NumChannels = 9;  // serializer.go:9-20

struct TensorState {
  std::vector<int32_t> Shape;  // {NumChannels, H, W}
  std::vector<float> Data;
};

struct Experience {
  std::string ExperienceId, GameId;
  int32_t PlayerId = 0, Turn = 0;
  TensorState State, NextState;
  int32_t Action = -1;
  float Reward = 0.f;
  bool Done = false;
  std::vector<bool> ActionMask;
  std::map<std::string, std::string> Metadata;
};

struct BufferStats {  // buffer.go:289-296
  int CurrentSize = 0, Capacity = 0;
  int64_t TotalAdded = 0, TotalDropped = 0;
  double UtilizationPct = 0;
};

class Buffer {
 public:
  explicit Buffer(int capacity) : capacity_(capacity <= 0 ? 10000 : capacity), buf_(size_t(capacity <= 0 ? 10000 : capacity)) {}

  core::Error Add(const Experience &exp) {  // buffer.go:58-93
    std::lock_guard<std::mutex> g(mu_);
    if (closed_) return core::Error(ErrBufferClosed);
    if (size_ >= capacity_) {  // circular: the oldest record makes room
      tail_ = (tail_ + 1) % capacity_;
      dropped_++;
    } else {
      size_++;
    }
    buf_[head_] = exp;
    head_ = (head_ + 1) % capacity_;
    added_++;
    return core::Error();
  }
  core::Error AddBatch(const std::vector<Experience> &batch) {  // buffer.go:96-134
    for (const Experience &e : batch)
      if (core::Error err = Add(e)) return err;
    return core::Error();
  }
  std::vector<Experience> Get(int n) {  // buffer.go:137-153, FIFO, removes
    std::lock_guard<std::mutex> g(mu_);
    if (n > size_) n = size_;
    std::vector<Experience> out;
    out.reserve(n > 0 ? n : 0);
    for (int i = 0; i < n; i++) {
      out.push_back(buf_[tail_]);
      tail_ = (tail_ + 1) % capacity_;
      size_--;
    }
    return out;
  }
  std::vector<Experience> GetAll() { return Get(Size()); }  // buffer.go:156-172
  std::vector<Experience> Sample(int n) const {             // buffer.go:175-192: the oldest n, not removed
    std::lock_guard<std::mutex> g(mu_);
    if (n > size_) n = size_;
    std::vector<Experience> out;
    for (int i = 0; i < n; i++) out.push_back(buf_[(tail_ + i) % capacity_]);
    return out;
  }
  std::vector<Experience> GetLatest(int n) const {  // buffer.go:195-212
    std::lock_guard<std::mutex> g(mu_);
    if (n > size_) n = size_;
    std::vector<Experience> out;
    for (int i = 0; i < n; i++) out.push_back(buf_[(head_ - n + i + capacity_) % capacity_]);
    return out;
  }
  int Size() const {
    std::lock_guard<std::mutex> g(mu_);
    return size_;
  }
  int Capacity() const { return capacity_; }
  bool IsFull() const { return Size() >= capacity_; }
  void Clear() {  // buffer.go:239-249
    std::lock_guard<std::mutex> g(mu_);
    size_ = head_ = tail_ = 0;
  }
  core::Error Close() {  // buffer.go:252-271
    std::lock_guard<std::mutex> g(mu_);
    if (closed_) return core::Error(ErrBufferClosed);
    closed_ = true;
    return core::Error();
  }
  BufferStats Stats() const {  // buffer.go:274-286
    std::lock_guard<std::mutex> g(mu_);
    BufferStats s;
    s.CurrentSize = size_;
    s.Capacity = capacity_;
    s.TotalAdded = added_;
    s.TotalDropped = dropped_;
    s.UtilizationPct = double(size_) / double(capacity_) * 100.0;
    return s;
  }

 private:
  mutable std::mutex mu_;
  int capacity_;
  std::vector<Experience> buf_;
  int size_ = 0, head_ = 0, tail_ = 0;
  bool closed_ = false;
  int64_t added_ = 0, dropped_ = 0;
};

// collector.go:13-140.  Attach() names the engine whose transitions are framed (the Go collector reads the states it is
// handed; this one reads the device's tensors of the same transition from the engine that calls it).
class SimpleCollector : public game::ExperienceCollector {
 public:
  SimpleCollector(int maxSize, std::string gameID) : buffer_(maxSize), gameID_(std::move(gameID)) {}
  void Attach(game::Engine *engine) { engine_ = engine; }

  void OnStateTransition(const game::GameState *prevState, const game::GameState *currState,
                         const std::map<int, game::Action> &actions) override {
    if (!engine_) return;
    for (const auto &kv : actions) {  // collector.go:33-36: every player that took an action
      const Transition *t = engine_->LastTransition(kv.first);
      if (!t) continue;
      Experience exp;
      exp.ExperienceId = NewId();  // uuid.New() in the reference: an opaque unique id
      exp.GameId = gameID_;
      exp.PlayerId = kv.first;
      exp.Turn = currState->Turn;
      exp.State.Shape = {NumChannels, prevState->Board->H, prevState->Board->W};
      exp.State.Data = t->State;
      exp.Action = t->Action;
      exp.Reward = t->Reward;
      exp.NextState.Shape = {NumChannels, currState->Board->H, currState->Board->W};
      exp.NextState.Data = t->NextState;
      exp.Done = currState->IsGameOver();
      exp.ActionMask = t->ActionMask;
      exp.Metadata["collector_version"] = "1.0.0";
      buffer_.Add(exp);
    }
  }
  void OnGameEnd(const game::GameState *) override { gamesEnded_++; }

  std::vector<Experience> GetExperiences() { return buffer_.GetAll(); }                    // collector.go:114-117
  int GetExperienceCount() const { return buffer_.Size(); }                                // collector.go:119-122
  std::vector<Experience> GetLatestExperiences(int n) const { return buffer_.GetLatest(n); }
  void Clear() { buffer_.Clear(); }
  Buffer &GetBuffer() { return buffer_; }
  const std::string &GameID() const { return gameID_; }
  int GamesEnded() const { return gamesEnded_; }

 private:
  static std::string NewId() {
    static std::mt19937_64 gen{std::random_device{}()};
    static std::mutex mu;
    std::lock_guard<std::mutex> g(mu);
    char out[37];
    const uint64_t a = gen(), b = gen();
    std::snprintf(out, sizeof out, "%08x-%04x-4%03x-%04x-%012llx", unsigned(a >> 32), unsigned((a >> 16) & 0xffff), unsigned(a & 0xfff),
                  unsigned(0x8000 | ((b >> 48) & 0x3fff)), (unsigned long long)(b & 0xffffffffffffULL));
    return out;
  }
  Buffer buffer_;
  std::string gameID_;
  game::Engine *engine_ = nullptr;
  int gamesEnded_ = 0;
};

}  // namespace experience
}  // namespace grl
