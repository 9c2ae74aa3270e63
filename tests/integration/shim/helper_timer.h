// Minimal stand-in for the CUDA-samples header <helper_timer.h> that the reference drivers include
// (src/main.cu:48, test/SpMV_test.cu:48) and that is not installed in this image.  Only what they use:
// StopWatchInterface with start/stop/reset/getTime (milliseconds) and sdkCreateTimer.  TEST INFRASTRUCTURE.
#pragma once
#include <chrono>
class StopWatchInterface {
  public:
    void start() { t0_ = clock::now(); running_ = true; }
    void stop() { if (running_) { total_ += std::chrono::duration<float, std::milli>(clock::now() - t0_).count(); running_ = false; } }
    void reset() { total_ = 0; running_ = false; }
    float getTime() { return running_ ? total_ + std::chrono::duration<float, std::milli>(clock::now() - t0_).count() : total_; }
  private:
    using clock = std::chrono::steady_clock;
    clock::time_point t0_;
    float total_ = 0;
    bool running_ = false;
};
inline bool sdkCreateTimer(StopWatchInterface** t) { *t = new StopWatchInterface(); return *t != nullptr; }
inline bool sdkDeleteTimer(StopWatchInterface** t) { delete *t; *t = nullptr; return true; }
