// pragma-dsp/xform (reference src/xform/index.ts:1)
export * from "./fourier.js";
