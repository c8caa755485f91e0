// inst_groups.h - how the kernel instantiations are split over translation units.
// X(kind, typetag, ctype, lo, hi): kind 0 = r2c (complex length 2^lo..2^hi, real N = 2M), 1 = c2c.
// build.py parses these lines; pragma_b200.cu expands them into the dispatch table.
#pragma once
#ifdef PDSP_EMU
// emulated test library (tests/simt_emu): a few sizes are enough to exercise the host pipeline
#define PDSP_GROUPS(X)        \
  X(0, f64, double, 0, 5)     \
  X(0, f64, double, 9, 9)     \
  X(0, f64, double, 11, 11)   \
  X(0, f32, float, 9, 9)      \
  X(1, f64, double, 0, 7)     \
  X(1, f64, double, 8, 10)
#else
#define PDSP_GROUPS(X)        \
  X(0, f64, double, 0, 5)     \
  X(0, f64, double, 6, 7)     \
  X(0, f64, double, 8, 8)     \
  X(0, f64, double, 9, 9)     \
  X(0, f64, double, 10, 10)   \
  X(0, f64, double, 11, 11)   \
  X(0, f64, double, 12, 12)   \
  X(0, f64, double, 13, 13)   \
  X(0, f32, float, 0, 5)      \
  X(0, f32, float, 6, 7)      \
  X(0, f32, float, 8, 8)      \
  X(0, f32, float, 9, 9)      \
  X(0, f32, float, 10, 10)    \
  X(0, f32, float, 11, 11)    \
  X(0, f32, float, 12, 12)    \
  X(0, f32, float, 13, 13)    \
  X(1, f64, double, 0, 7)     \
  X(1, f64, double, 8, 10)    \
  X(1, f64, double, 11, 13)   \
  X(1, f32, float, 0, 7)      \
  X(1, f32, float, 8, 10)     \
  X(1, f32, float, 11, 13)
#endif
