// Drop-in for pragma-dsp/xform/fourier (reference src/xform/fourier.ts): every export - FFT, createWindow, applyWindow,
// magnitude, phase, fftShift, fftShiftComplex, binFrequencies, WindowType, FftSides - keeps its signature and messages.
import { type ComplexArray, Radix2Fft, createComplexArray, isPowerOfTwo } from "../core/fft.js";
import { SIDES_CODE, WINDOW_CODE, ctx, native } from "../native.js";

export type WindowType = "rect" | "hann" | "hamming" | "blackman";
export type FftSides = "one" | "two";

export const createWindow = (type: WindowType, size: number): Float64Array => {
  if (size <= 0) throw new Error(`Window size must be positive, got ${size}`);
  const code = WINDOW_CODE[type];
  if (code === undefined) throw new Error(`Unsupported window type: ${type}`);
  const out = new Float64Array(size);
  native().createWindow(code, size, out);
  return out;
};

// applyWindow(input, window, out?) - reference src/xform/fourier.ts:54-67 (the fused spectrum() path never materialises it)
export const applyWindow = (input: ArrayLike<number>, window: ArrayLike<number>, out?: Float64Array): Float64Array => {
  if (input.length !== window.length) throw new Error("Window length must match input length.");
  const result = out ?? new Float64Array(input.length);
  if (result.length < input.length) throw new Error("Window length must match input length.");
  native().applyWindow(ctx(), toF64(input), toF64(window), result);
  return result;
};

const toF64 = (a: ArrayLike<number>): Float64Array => {
  if (a instanceof Float64Array) return a;
  const out = new Float64Array(a.length);
  for (let i = 0; i < a.length; i += 1) out[i] = a[i] ?? 0;
  return out;
};

export class FFT {
  readonly size: number;
  private readonly kernel: Radix2Fft;
  constructor(size: number) {
    if (!isPowerOfTwo(size)) throw new Error(`FFT size must be power of two, got ${size}`);
    this.size = size;
    this.kernel = new Radix2Fft(size);
  }
  forward(input: ArrayLike<number>, out?: ComplexArray): ComplexArray { return this.kernel.forward(input, out); }
  forwardComplex(input: ComplexArray, out?: ComplexArray): ComplexArray { return this.kernel.forwardComplex(input, out); }
  inverse(input: ComplexArray, out?: ComplexArray): ComplexArray { return this.kernel.inverse(input, out); }
  createComplexArray(fill = 0): ComplexArray { return createComplexArray(this.size, fill); }
}

export const magnitude = (input: ComplexArray, out?: Float64Array): Float64Array => {
  const result = out ?? new Float64Array(input.real.length);
  native().magnitude(ctx(), input.real, input.imag, result);
  return result;
};

export const phase = (input: ComplexArray, out?: Float64Array): Float64Array => {
  const result = out ?? new Float64Array(input.real.length);
  native().phase(ctx(), input.real, input.imag, result);
  return result;
};

// fftShift(input, out?) - reference src/xform/fourier.ts:122-134: rotate by floor(n/2) so that DC sits in the middle
export const fftShift = (input: ArrayLike<number>, out?: Float64Array): Float64Array => {
  const result = out ?? new Float64Array(input.length);
  if (input.length > 0) native().fftShift(ctx(), toF64(input), result);
  return result;
};

// fftShiftComplex(input, out?) - reference src/xform/fourier.ts:136-145: fftShift applied to each plane
export const fftShiftComplex = (input: ComplexArray, out?: ComplexArray): ComplexArray => {
  const result = out ?? createComplexArray(input.real.length);
  fftShift(input.real, result.real);
  fftShift(input.imag, result.imag);
  return result;
};

export const binFrequencies = (size: number, sampleRate: number, sides: FftSides = "one"): Float64Array => {
  if (size <= 0) throw new Error(`FFT size must be positive, got ${size}`);
  if (sampleRate <= 0) throw new Error(`Sample rate must be positive, got ${sampleRate}`);
  const out = new Float64Array(sides === "one" ? Math.floor(size / 2) + 1 : size);
  native().binFrequencies(size, sampleRate, SIDES_CODE[sides], out);
  return out;
};
