// inst_var.cu - tuning variants of the headline size (N = 1024): alternative thread/radix mappings
// (fft_config.h PDSP_VARIANT) for the three hot output modes.  Selected at run time with
// PDSP_VARIANT=<n>; compiled per (type, variant) with -DPDSP_VAR_T=float|double -DPDSP_VAR=<n>
// -DPDSP_INST_NAME=symbol.  Results are identical across variants up to rounding.
#include "fft_launch.cuh"

namespace pdsp {
cudaError_t PDSP_INST_NAME(int mode, const R2CParams& p, const LaunchCtx& lc) {
  switch (mode) {
    case MD_AMP:
      return launch_r2c_t<PDSP_VAR_T, kVariantLog2M, MD_AMP, PDSP_VAR>(p, lc);
    case MD_AMP | MD_PEAK:
      return launch_r2c_t<PDSP_VAR_T, kVariantLog2M, MD_AMP | MD_PEAK, PDSP_VAR>(p, lc);
    case MD_PEAK:
      return launch_r2c_t<PDSP_VAR_T, kVariantLog2M, MD_PEAK, PDSP_VAR>(p, lc);
    default:
      return cudaErrorInvalidValue;
  }
}
}  // namespace pdsp
