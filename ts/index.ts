// Root export of the drop-in (reference src/index.ts:1): spectrum() and its types.
export * from "./public/spectrum.js";
