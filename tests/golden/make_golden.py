#!/usr/bin/env python3
"""Build the compact golden fixtures under tests/golden/ from the reference checkout.

Run in the BUILD container only (needs /root/reference; the GPU box has no copy):

    python tests/golden/make_golden.py [--reference /root/reference]

What it produces (all float64, lossless: JSON decimal -> binary64 round-trips exactly):

* ``reallife.npz``  - the six committed NumPy/SciPy golden files of the reference,
  ``test/reallife/references/{pure_sine,cosine,multi_tone,chirp,special,windows_dsp}.json``
  (loaded by the reference at ``test/reallife/helpers.ts:62-87``), repacked as arrays
  ``signal/fftRe/fftIm/magnitude/phase`` of shape (35, 1024) plus the 16 window cases.
* ``reallife_meta.json`` - per-case name/kind/params/peakBin (small, human-readable).
* ``fixtures_v0_1.npz`` + ``fixtures_v0_1_meta.json`` - the fixture file the reference's
  ``test/fft.test.ts``, ``test/spectrum.test.ts``, ``test/window.test.ts`` and ``bench/run.ts``
  load through ``test/fixtures.ts:44-51``. It is MISSING from the reference checkout, so it is
  regenerated here by running the reference's own generator ``scripts/gen_fixtures.py``
  (seed 1337) unmodified, as a subprocess, into a temp dir.

Nothing from the reference's *source* is copied; only its test vectors are repacked.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SIGNAL_FILES = ["pure_sine", "cosine", "multi_tone", "chirp", "special"]
ARRS = ["signal", "fftRe", "fftIm", "magnitude", "phase"]


def pack_reallife(ref_root: str) -> None:
    refs = os.path.join(ref_root, "test", "reallife", "references")
    arrays = {a: [] for a in ARRS}
    meta = {"cases": [], "windows": [], "source": "test/reallife/references/*.json"}
    for fname in SIGNAL_FILES:
        with open(os.path.join(refs, fname + ".json")) as fh:
            doc = json.load(fh)
        meta.setdefault("generator", {k: doc[k] for k in ("generatedAt", "generator", "python", "numpy", "scipy")})
        for case in doc["cases"]:
            for a in ARRS:
                arrays[a].append(np.asarray(case[a], dtype=np.float64))
            meta["cases"].append(
                {
                    "file": fname,
                    "name": case["name"],
                    "kind": case["kind"],
                    "n": case["n"],
                    "sampleRate": case["sampleRate"],
                    "peakBin": case["peakBin"],
                    "peakMagnitude": case["peakMagnitude"],
                    "peakPhase": case["peakPhase"],
                    "params": case["params"],
                }
            )
    out = {a: np.stack(arrays[a]) for a in ARRS}
    with open(os.path.join(refs, "windows_dsp.json")) as fh:
        wdoc = json.load(fh)
    for i, w in enumerate(wdoc["cases"]):
        out[f"window_{i}"] = np.asarray(w["values"], dtype=np.float64)
        meta["windows"].append({"type": w["type"], "n": w["n"], "coherentGain": w["coherentGain"], "enbw": w["enbw"]})
    np.savez_compressed(os.path.join(HERE, "reallife.npz"), **out)
    with open(os.path.join(HERE, "reallife_meta.json"), "w") as fh:
        json.dump(meta, fh, indent=1)
    print(f"reallife: {len(meta['cases'])} signal cases, {len(meta['windows'])} windows")


def pack_fixtures(ref_root: str) -> None:
    gen = os.path.join(ref_root, "scripts", "gen_fixtures.py")
    with tempfile.TemporaryDirectory() as tmp:
        out_json = os.path.join(tmp, "pragma-dsp.v0.1.json")
        subprocess.run([sys.executable, gen, "--out", out_json], check=True, cwd=tmp)
        with open(out_json) as fh:
            doc = json.load(fh)
    out = {}
    meta = {
        "source": "scripts/gen_fixtures.py --out <tmp> (seed 1337), regenerated: file absent from checkout",
        "generator": doc["generator"],
        "convention": doc["convention"],
        "fftCases": [],
        "windows": [],
    }
    for i, c in enumerate(doc["fftCases"]):
        out[f"case_{i}_input"] = np.asarray(c["input"], dtype=np.float64)
        out[f"case_{i}_fftRe"] = np.asarray(c["fftRe"], dtype=np.float64)
        out[f"case_{i}_fftIm"] = np.asarray(c["fftIm"], dtype=np.float64)
        meta["fftCases"].append(
            {"name": c["name"], "kind": c["kind"], "n": c["n"], "sampleRate": c["sampleRate"], "meta": c["meta"]}
        )
    for i, w in enumerate(doc["windows"]):
        out[f"window_{i}"] = np.asarray(w["values"], dtype=np.float64)
        meta["windows"].append({"type": w["type"], "n": w["n"], "sym": w["sym"]})
    np.savez_compressed(os.path.join(HERE, "fixtures_v0_1.npz"), **out)
    with open(os.path.join(HERE, "fixtures_v0_1_meta.json"), "w") as fh:
        json.dump(meta, fh, indent=1)
    print(f"fixtures v0.1: {len(meta['fftCases'])} fft cases, {len(meta['windows'])} windows")


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    pack_reallife(args.reference)
    pack_fixtures(args.reference)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
