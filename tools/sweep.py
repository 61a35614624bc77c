#!/usr/bin/env python
"""Batch-size sweep (BASELINE.json config 5) + the 500-phoneme stress shape (config 4): runs bench.py per point
and prints a markdown table. Usage: python tools/sweep.py > profiles/rNN_sweep.md"""
import json, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
points = [("S", b, False) for b in (1, 8, 64, 256, 1024, 4096)] + [("T", b, False) for b in (1, 32, 256, 1024)] + \
         [("S", 64, True), ("T", 16, True)]
print("# batch-size sweep, one B200, 16-bit tensor-core path (bench.py --steps 5 --warmup 3; device-resident / end-to-end)\n")
print("| model | batch | workload | frames/step | ms/step | M frames/s | e2e M frames/s | decoder ms | decoder frac of the burst tensor peak |")
print("|---|---:|---|---:|---:|---:|---:|---:|---:|")
for model, batch, stress in points:
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--model", model, "--batch", str(batch), "--steps", "5",
           "--warmup", "3", "--no-cpu-baseline", "--no-extra"] + (["--stress"] if stress else [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    try:
        j = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        print(f"| {model} | {batch} | FAILED: {r.stderr.strip().splitlines()[-1] if r.stderr else ''} |")
        continue
    print(f"| {model} | {batch} | {'500-phoneme stress' if stress else 'LJSpeech-shaped'} | {int(j['frames_per_step'])} | "
          f"{j['ms_per_step']:.3f} | {j['value'] / 1e6:.2f} | {j['e2e']['value'] / 1e6:.2f} | "
          f"{j['stage_ms_per_step']['decoder_loop']:.3f} | {j['roofline']['frac']:.3f} |", flush=True)
