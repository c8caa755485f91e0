// Drop-in for pragma-dsp/effect (reference src/effect/index.ts): Fourier Tag, FourierLive, spectrumFx,
// spectrumStream.  The service caches FFT objects (device plan handles) and windows exactly like the
// reference's Maps; spectrumStream keeps the 1:1 ordered contract but maps whole chunks of the stream
// through one batched kernel launch instead of one transform per element.
import { Chunk, Context, Effect, Layer, Stream } from "effect";
import { nextPowerOfTwo } from "../core/fft.js";
import { F64, PEAK_BYTES_F64, SIDES_CODE, WINDOW_CODE, native, plan } from "../native.js";
import { readPeak, spectrum, type SpectrumResult } from "../public/spectrum.js";
import { FFT, type FftSides, type WindowType, binFrequencies, createWindow } from "../xform/fourier.js";

export interface FourierService {
  fft: (size: number) => FFT;
  window: (type: WindowType, size: number) => Float64Array;
}
export class Fourier extends Context.Tag("pragma-dsp/Fourier")<Fourier, FourierService>() {}

export const FourierLive = Layer.effect(Fourier, Effect.sync(() => {
  const ffts = new Map<number, FFT>();
  const windows = new Map<string, Float64Array>();
  const service: FourierService = {
    fft: (size) => ffts.get(size) ?? (ffts.set(size, new FFT(size)), ffts.get(size)!),
    window: (type, size) => {
      const key = `${type}:${size}`;
      return windows.get(key) ?? (windows.set(key, createWindow(type, size)), windows.get(key)!);
    }
  };
  return service;
}));

export type SpectrumFxOptions = { sampleRate?: number; fftSize?: number; window?: WindowType; sides?: FftSides };
export type SpectrumFxResult = SpectrumResult;

export const spectrumFx = (samples: ArrayLike<number>, options: SpectrumFxOptions = {}) =>
  Effect.gen(function* () {
    const service = yield* Fourier;
    service.fft(options.fftSize ?? nextPowerOfTwo(samples.length)); // plan stays cached with the Layer
    return spectrum(samples, options);
  });

const batch = (frames: ReadonlyArray<Float32Array>, options: SpectrumFxOptions): SpectrumFxResult[] => {
  const len = frames[0]!.length;
  const size = options.fftSize ?? nextPowerOfTwo(len);
  const sides: FftSides = options.sides ?? "one";
  const bins = sides === "one" ? Math.floor(size / 2) + 1 : size;
  const packed = new Float32Array(len * frames.length);
  frames.forEach((f, i) => packed.set(f, i * len));
  const amplitude = new Float64Array(bins * frames.length);
  const phase = new Float64Array(bins * frames.length);
  const peaks = new Uint8Array(PEAK_BYTES_F64 * frames.length);
  native().spectrum(plan(size, F64), packed,
    { frameLen: len, hop: len, batch: frames.length, window: WINDOW_CODE[options.window ?? "rect"], sides: SIDES_CODE[sides],
      sampleRate: options.sampleRate ?? 1 }, amplitude, phase, peaks);
  const frequencies = binFrequencies(size, options.sampleRate ?? 1, sides);
  return frames.map((_, i) => ({
    frequencies,
    amplitude: amplitude.subarray(i * bins, (i + 1) * bins),
    phase: phase.subarray(i * bins, (i + 1) * bins),
    peak: readPeak(peaks, i)
  }));
};

export const spectrumStream = (frames: Stream.Stream<Float32Array>, options: SpectrumFxOptions = {}) =>
  Stream.mapChunksEffect(frames, (chunk) =>
    Effect.gen(function* () {
      yield* Fourier;
      const arr = Chunk.toReadonlyArray(chunk);
      const out: SpectrumFxResult[] = [];
      // runs of equal-length frames share a launch; order is preserved
      for (let i = 0; i < arr.length;) {
        let j = i + 1;
        while (j < arr.length && arr[j]!.length === arr[i]!.length) j += 1;
        out.push(...batch(arr.slice(i, j), options));
        i = j;
      }
      return Chunk.fromIterable(out);
    }));
