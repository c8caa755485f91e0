// Drop-in for pragma-dsp/xform/fourier (reference src/xform/fourier.ts): FFT, createWindow,
// magnitude, phase, binFrequencies keep their signatures and messages.
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

export const binFrequencies = (size: number, sampleRate: number, sides: FftSides = "one"): Float64Array => {
  if (size <= 0) throw new Error(`FFT size must be positive, got ${size}`);
  if (sampleRate <= 0) throw new Error(`Sample rate must be positive, got ${sampleRate}`);
  const out = new Float64Array(sides === "one" ? Math.floor(size / 2) + 1 : size);
  native().binFrequencies(size, sampleRate, SIDES_CODE[sides], out);
  return out;
};
