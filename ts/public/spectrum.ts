// Drop-in for the root export (reference src/public/spectrum.ts): spectrum(samples, options).
// One addon call = one fused kernel launch (frame build, window, FFT, magnitude, phase, scaling, peak).
import { nextPowerOfTwo } from "../core/fft.js";
import { F64, PEAK_BYTES_F64, SIDES_CODE, WINDOW_CODE, asSamples, native, plan } from "../native.js";
import { type FftSides, type WindowType, binFrequencies } from "../xform/fourier.js";

export type SpectrumPeak = { index: number; frequency: number; amplitude: number; phase: number };
export type SpectrumResult = { frequencies: Float64Array; amplitude: Float64Array; phase: Float64Array; peak: SpectrumPeak };
export type SpectrumOptions = { sampleRate?: number; fftSize?: number; window?: WindowType; sides?: FftSides };

export const readPeak = (bytes: Uint8Array, frame: number): SpectrumPeak => {
  const view = new DataView(bytes.buffer, bytes.byteOffset + frame * PEAK_BYTES_F64, PEAK_BYTES_F64);
  return {
    index: view.getInt32(0, true),
    frequency: view.getFloat64(8, true),
    amplitude: view.getFloat64(16, true),
    phase: view.getFloat64(24, true)
  };
};

export const spectrum = (samples: ArrayLike<number>, options: SpectrumOptions = {}): SpectrumResult => {
  const sampleRate = options.sampleRate ?? 1;
  const sides: FftSides = options.sides ?? "one";
  const size = options.fftSize ?? nextPowerOfTwo(samples.length);
  const window: WindowType = options.window ?? "rect";
  if (WINDOW_CODE[window] === undefined) throw new Error(`Unsupported window type: ${window}`);
  const handle = plan(size, F64); // throws "FFT size must be power of two" through Radix2Fft's check upstream
  const bins = sides === "one" ? Math.floor(size / 2) + 1 : size;
  const amplitude = new Float64Array(bins);
  const phase = new Float64Array(bins);
  const peaks = new Uint8Array(PEAK_BYTES_F64);
  native().spectrum(handle, asSamples(samples),
    { frameLen: samples.length, hop: samples.length, batch: 1, window: WINDOW_CODE[window], sides: SIDES_CODE[sides], sampleRate },
    amplitude, phase, peaks);
  return { frequencies: binFrequencies(size, sampleRate, sides), amplitude, phase, peak: readPeak(peaks, 0) };
};
