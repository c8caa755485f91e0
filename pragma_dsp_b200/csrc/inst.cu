// inst.cu - compiled once per (kind, type, size range); the build passes
//   -DPDSP_INST_KIND=0|1 (r2c|c2c) -DPDSP_INST_T=float|double -DPDSP_INST_LO=a -DPDSP_INST_HI=b
//   -DPDSP_INST_NAME=symbol
// so the unrolled kernels compile in parallel.  The ranges are listed in inst_groups.h.
#include "fft_launch.cuh"

namespace pdsp {
#define PDSP_CASES(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13)

#if PDSP_INST_KIND == 0
cudaError_t PDSP_INST_NAME(int log2m, int mode, const R2CParams& p, const LaunchCtx& lc) {
  switch (log2m) {
#define X(L) \
  case L:    \
    return r2c_case<PDSP_INST_T, L, PDSP_INST_LO, PDSP_INST_HI>(mode, p, lc);
    PDSP_CASES(X)
#undef X
    default:
      return cudaErrorInvalidValue;
  }
}
#else
cudaError_t PDSP_INST_NAME(int log2m, const C2CParams& p, const LaunchCtx& lc) {
  switch (log2m) {
#define X(L) \
  case L:    \
    return c2c_case<PDSP_INST_T, L, PDSP_INST_LO, PDSP_INST_HI>(p, lc);
    PDSP_CASES(X)
#undef X
    default:
      return cudaErrorInvalidValue;
  }
}
#endif
}  // namespace pdsp
