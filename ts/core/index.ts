// pragma-dsp/core (reference src/core/index.ts:3)
export * from "./fft.js";
