// native.ts - loads the Node-API addon (napi/pragma_napi.cc over include/pragma_b200.h) and keeps the
// per-process context + plan handles.  Everything numeric happens on the GPU behind these calls;
// there is no JavaScript fallback - a missing addon or device throws on first use.
import { createRequire } from "node:module";

export type Handle = unknown;
export interface Native {
  abiVersion(): number;
  contextCreate(device: number): Handle;
  planGet(ctx: Handle, size: number, precision: number): Handle;
  fftForwardReal(plan: Handle, input: Float32Array | Float64Array, outReal: Float64Array, outImag: Float64Array): void;
  fftForwardComplex(plan: Handle, inRe: Float64Array, inIm: Float64Array, outRe: Float64Array, outIm: Float64Array): void;
  fftInverse(plan: Handle, inRe: Float64Array, inIm: Float64Array, outRe: Float64Array, outIm: Float64Array): void;
  magnitude(ctx: Handle, re: Float64Array, im: Float64Array, out: Float64Array): void;
  phase(ctx: Handle, re: Float64Array, im: Float64Array, out: Float64Array): void;
  applyWindow(ctx: Handle, input: Float64Array, window: Float64Array, out: Float64Array): void;
  fftShift(ctx: Handle, input: Float64Array, out: Float64Array): void;
  spectrum(plan: Handle, samples: Float32Array | Float64Array, opts: Record<string, number>,
           amplitude: Float32Array | Float64Array | null, phase: Float32Array | Float64Array | null,
           peaks: Uint8Array | null): void;
  createWindow(type: number, size: number, out: Float64Array): void;
  binFrequencies(size: number, sampleRate: number, sides: number, out: Float64Array): void;
  hostAlloc(ctx: Handle, bytes: number): ArrayBuffer;
  // ingestion ring (pdsp_ingest_*): frames in as they arrive, results out in arrival order; a long-lived source
  // (microphone, websocket) pushes each frame and pops whatever is finished instead of assembling batches
  ingestOpen(plan: Handle, desc: { frameLen: number; window: number; sides: number; sampleRate: number; sampleDtype: number },
             wantAmplitude: number, wantPhase: number, wantPeaks: number, framesPerChunk: number, depth: number): Handle;
  ingestPush(ring: Handle, frames: Float32Array | Float64Array, count: number): number;
  ingestFlush(ring: Handle): void;
  ingestPop(ring: Handle, amplitude: Float64Array | Float32Array | null, phase: Float64Array | Float32Array | null,
            peaks: Uint8Array | null, maxFrames: number): number;
}

export const F32 = 0, F64 = 1;
export const WINDOW_CODE = { rect: 0, hann: 1, hamming: 2, blackman: 3 } as const;
export const SIDES_CODE = { one: 0, two: 1 } as const;
export const PEAK_BYTES_F64 = 32;

let addon: Native | undefined;
let context: Handle | undefined;
const plans = new Map<string, Handle>();

export const native = (): Native => {
  if (!addon) {
    const require = createRequire(import.meta.url);
    addon = require(process.env.PRAGMA_B200_ADDON ?? "../build/pragma_b200.node") as Native;
    if (addon.abiVersion() !== 1) throw new Error("pragma-dsp/b200: addon ABI mismatch");
  }
  return addon;
};

export const ctx = (): Handle => {
  if (context === undefined) context = native().contextCreate(Number(process.env.PDSP_DEVICE ?? 0));
  return context;
};

export const plan = (size: number, precision: number = F64): Handle => {
  const key = `${size}:${precision}`;
  let p = plans.get(key);
  if (p === undefined) {
    p = native().planGet(ctx(), size, precision);
    plans.set(key, p);
  }
  return p;
};

// ArrayLike<number> -> typed array the addon can borrow; undefined elements read as 0 like `?? 0`.
export const asSamples = (input: ArrayLike<number>): Float32Array | Float64Array => {
  if (input instanceof Float64Array || input instanceof Float32Array) return input;
  const out = new Float64Array(input.length);
  for (let i = 0; i < input.length; i += 1) out[i] = input[i] ?? 0;
  return out;
};
