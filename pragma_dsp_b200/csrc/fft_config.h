// fft_config.h - per-size kernel configuration (shared by the launchers and the test emulator).
#pragma once

namespace pdsp {

constexpr int kMaxLog2M = 13;      // largest in-CTA complex length 8192 (real frames up to 16384)
constexpr int kMinSpecLog2M = 5;   // sizes below this only get the generic (runtime-flag) kernel
constexpr int kNumVariants = 18;   // tuning variants compiled for kVariantLog2M (see inst_var.cu)
constexpr int kVariantLog2M = 9;   // N = 1024, the headline size

// Points per thread / radix / CTA size / occupancy target for a complex length 2^LOG2M.
// VAR = 0 is what ships for every size; VAR > 0 are alternative mappings of the headline size,
// selectable at run time with PDSP_VARIANT=<n> so one GPU session can rank them.
template <typename T, int LOG2M, int VAR = 0>
struct KCfg {
  // doubles: 16 points per thread from M = 512 up (radix 16 passes: 512 = 16 x 16 x 2, one warp per frame);
  // floats: 32 points per thread (radix 32: 512 = 32 x 16, a single exchange) - registers allow it and the
  // fp32 kernels are issue-bound, so fewer exchange instructions win (profiles/r1 sweep).  Below 512: 8.
#ifndef PDSP_F32_P32_FROM
#define PDSP_F32_P32_FROM 9  // smallest log2(M) at which fp32 kernels hold 32 points per thread
#endif
#ifndef PDSP_F64_P32_AT
#define PDSP_F64_P32_AT -1  // experiment: log2(M) at which fp64 kernels hold 32 points per thread (one warp per frame)
#endif
#ifndef PDSP_F32_SMALL_LOG2P
#define PDSP_F32_SMALL_LOG2P 4  // fp32, M = 128 / 256: 16 points per thread (N = 512: 0.55 -> 0.70 of roofline)
#endif
// fp32, small frames: log2(points per thread) for M = 128 (N = 256), 64, 32, 16.  A frame's lanes read 8 bytes each, so
// fewer points per thread = more lanes per frame = longer contiguous segments per load (TF = 16 lanes -> 128 bytes).
#ifndef PDSP_F32_LOG2P_M7
#define PDSP_F32_LOG2P_M7 3  // N = 256: 8 points, 16 lanes per frame (amp 0.60 -> 0.65, all-bins forward 0.55 -> 0.77)
#endif
#ifndef PDSP_F32_LOG2P_M6
#define PDSP_F32_LOG2P_M6 3  // N = 128: 4 points (16 lanes) helps the all-bins forward (0.53 -> 0.83) but costs amplitude modes 10 %
#endif
#ifndef PDSP_F32_LOG2P_M5
#define PDSP_F32_LOG2P_M5 3
#endif
#ifndef PDSP_F32_LOG2P_M4
#define PDSP_F32_LOG2P_M4 3
#endif
  static constexpr int LOG2P = LOG2M < 3 ? LOG2M
                               : sizeof(T) == 4 ? (LOG2M >= PDSP_F32_P32_FROM ? 5 : LOG2M == 8 ? PDSP_F32_SMALL_LOG2P : LOG2M == 7 ? PDSP_F32_LOG2P_M7
                                  : LOG2M == 6 ? PDSP_F32_LOG2P_M6 : LOG2M == 5 ? PDSP_F32_LOG2P_M5 : LOG2M == 4 ? PDSP_F32_LOG2P_M4 : 3)
                                                : (LOG2M == PDSP_F64_P32_AT ? 5 : (LOG2M >= 9 ? 4 : 3));
  static constexpr int MAXRB = LOG2P >= 4 ? LOG2P : 3;
  static constexpr int TF = (1 << LOG2M) >> LOG2P;
  static constexpr int THREADS = LOG2P == 5 ? (TF > 32 ? TF : 32) : (TF > 128 ? TF : 128);
  // occupancy target via __launch_bounds__(THREADS, MINB): 128 regs/thread, except 96 for floats holding
  // 8 or 16 points.  MAXREG > 0 selects a hard __maxnreg__ cap instead (tuning variants).
#ifndef PDSP_F64_MINB_M9
#define PDSP_F64_MINB_M9 4  // experiment: CTAs per SM of the fp64 N = 1024 kernels (4 = 128 registers, 5 = 102)
#endif
  static constexpr int MINB = (sizeof(T) == 8 && LOG2M == 9 && LOG2P == 4) ? PDSP_F64_MINB_M9 : (sizeof(T) == 8 && LOG2P == 5) ? (256 / THREADS > 0 ? 256 / THREADS : 1)
                              : (sizeof(T) == 4 && LOG2P < 5 && THREADS <= 128) ? 5 : (512 / THREADS > 0 ? 512 / THREADS : 1);
  static constexpr int MAXREG = 0;
};

#define PDSP_VARIANT(VAR, LP, RB, THR, MB64, MB32, REG64, REG32)             \
  template <typename T>                                                     \
  struct KCfg<T, kVariantLog2M, VAR> {                                      \
    static constexpr int LOG2P = LP;                                        \
    static constexpr int MAXRB = RB;                                        \
    static constexpr int TF = (1 << kVariantLog2M) >> LOG2P;                \
    static constexpr int THREADS = THR;                                     \
    static constexpr int MINB = sizeof(T) == 8 ? MB64 : MB32;               \
    static constexpr int MAXREG = sizeof(T) == 8 ? REG64 : REG32;           \
  };
//           var log2P radix-bits threads minB(f64,f32) maxnreg(f64,f32; 0 = use minB)
PDSP_VARIANT(1, 4, 3, 128, 4, 5, 0, 0)      // 8 x 8 x 8, one warp per frame (the default is 16 x 16 x 2)
PDSP_VARIANT(2, 5, 5, 64, 4, 8, 0, 0)       // 32 x 16, 16 threads per frame, one exchange
PDSP_VARIANT(3, 3, 3, 128, 6, 8, 0, 0)      // 8 x 8 x 8, two warps per frame (named barriers)
PDSP_VARIANT(4, 4, 3, 128, 3, 4, 0, 0)      // baseline mapping with a looser register cap
PDSP_VARIANT(5, 5, 5, 128, 2, 4, 0, 0)      // 32 x 16, 8 frames per CTA
PDSP_VARIANT(6, 4, 3, 256, 2, 2, 0, 0)      // baseline mapping, 8 frames per CTA
PDSP_VARIANT(7, 5, 5, 32, 8, 16, 0, 0)      // 32 x 16, one warp (2 frames) per CTA (= the fp32 default)

PDSP_VARIANT(8, 4, 4, 32, 18, 24, 0, 0)     // 16 x 16 x 2, single-warp CTAs: 18 (f64) / 24 (f32) per SM
PDSP_VARIANT(9, 4, 4, 32, 20, 28, 0, 0)     // same, 20 / 28 per SM
PDSP_VARIANT(10, 4, 4, 32, 1, 1, 112, 80)   // same, hard caps 112 / 80 registers
PDSP_VARIANT(11, 4, 3, 32, 18, 24, 0, 0)    // 8 x 8 x 8, single-warp CTAs
PDSP_VARIANT(12, 4, 4, 32, 1, 1, 120, 88)   // 16 x 16 x 2, hard caps 120 / 88 registers
PDSP_VARIANT(13, 4, 4, 32, 16, 21, 0, 0)    // 16 x 16 x 2, single-warp CTAs at the baseline occupancy
PDSP_VARIANT(14, 5, 5, 32, 8, 20, 0, 0)     // fp32 default mapping (32 x 16) at 20 warps/SM (96 regs)
PDSP_VARIANT(15, 5, 5, 32, 8, 24, 0, 0)     // ... 24 warps/SM (80 regs)
PDSP_VARIANT(16, 5, 5, 64, 4, 10, 0, 0)     // ... two-warp CTAs, 20 warps/SM
PDSP_VARIANT(17, 5, 5, 32, 8, 1, 0, 112)    // ... hard cap 112 registers (18 warps/SM)
#undef PDSP_VARIANT

}  // namespace pdsp
