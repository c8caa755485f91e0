// Drop-in for pragma-dsp/core (reference src/core/fft.ts): same exports, same signatures, same
// error messages; the transform itself is one call into the B200 library.
import { F32, F64, asSamples, native, plan, type Handle } from "../native.js";

export type ComplexArray = { real: Float64Array; imag: Float64Array };

export const createComplexArray = (size: number, fill = 0): ComplexArray => {
  const out: ComplexArray = { real: new Float64Array(size), imag: new Float64Array(size) };
  if (fill !== 0) {
    out.real.fill(fill);
    out.imag.fill(fill);
  }
  return out;
};

export const isPowerOfTwo = (n: number): boolean => n > 0 && (n & (n - 1)) === 0;

export const nextPowerOfTwo = (n: number): number => {
  let p = 1;
  while (p < n) p *= 2;
  return p;
};

export class Radix2Fft {
  readonly size: number;
  private readonly handle: Handle;

  constructor(size: number) {
    if (!isPowerOfTwo(size)) throw new Error(`FFT size must be power of two, got ${size}`);
    this.size = size;
    this.handle = plan(size, F64);
  }

  private check(len: number): void {
    if (len !== this.size) throw new Error(`FFT input length ${len} != size ${this.size}`);
  }

  forward(input: ArrayLike<number>, out?: ComplexArray): ComplexArray {
    this.check(input.length);
    const result = out ?? createComplexArray(this.size);
    native().fftForwardReal(this.handle, asSamples(input), result.real, result.imag);
    return result;
  }

  forwardComplex(input: ComplexArray, out?: ComplexArray): ComplexArray {
    this.check(input.real.length);
    this.check(input.imag.length);
    const result = out ?? createComplexArray(this.size);
    native().fftForwardComplex(this.handle, input.real, input.imag, result.real, result.imag);
    return result;
  }

  inverse(input: ComplexArray, out?: ComplexArray): ComplexArray {
    this.check(input.real.length);
    this.check(input.imag.length);
    const result = out ?? createComplexArray(this.size);
    native().fftInverse(this.handle, input.real, input.imag, result.real, result.imag);
    return result;
  }
}

export { F32, F64 };
