// fft_config.h - per-size kernel configuration (shared by the launchers and the test emulator).
#pragma once

namespace pdsp {

constexpr int kMaxLog2M = 13;  // largest in-CTA complex length 8192 (real frames up to 16384)

// Points per thread / radix / CTA size for a complex length 2^LOG2M.
template <typename T, int LOG2M>
struct KCfg {
  static constexpr int LOG2P = LOG2M < 3 ? LOG2M : (LOG2M >= 9 ? 4 : 3);
  static constexpr int MAXRB = 3;
  static constexpr int TF = (1 << LOG2M) >> LOG2P;
  static constexpr int THREADS = TF > 128 ? TF : 128;
};

}  // namespace pdsp
