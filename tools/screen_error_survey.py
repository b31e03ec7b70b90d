#!/usr/bin/env python
"""fp32 screening error survey (run on a GPU box): for every FFT length the screening kernels serve, the worst observed
distance between the fp32 second stage's |cc| maximum and the exact fp64 score, and the smallest margin the bounds left,
over the adversarial generator of tests/test_gpu_screen.py (offsets of 1e9, 1e-12 / 1e12 amplitudes, spikes, trends, ...).
Writes gpurun_out/screen_error_survey.json (committed as profiles/screen_error_survey.json)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "go-muse_b200"), os.path.join(ROOT, "tests")]
import muse_b200 as mb  # noqa: E402
from test_gpu_screen import _adversarial  # noqa: E402

SLACK = 1e-4
ctx = mb.default_context(0)
rows = []
for N in (66, 100, 128, 200, 255, 256, 300, 480, 512, 1000, 1024, 1026, 1440, 1441, 2047, 2048, 2500, 4096, 5000, 8192, 10080, 10081, 16384):
    rng = np.random.default_rng(1000 + N)
    S = 12000 if N <= 2048 else 3000
    Y = _adversarial(rng, S, N)
    ref = np.zeros(N)
    ref[N // 2 - 5:N // 2 + 5] = 1.5
    ref += 0.1 * (rng.random(N) - 0.5)
    store = mb.DeviceStore(ctx, N, 0, S)
    store.append(Y)
    b = mb.DeviceBatch(ctx, store, ref)
    sc, lg = b.score_all()
    u = b.screen_bounds().astype(np.float64)
    up, lo = b.screen_bounds(refine=True, max_lag=N)          # window = every lag: upper and lower bracket the score itself
    up, lo = up.astype(np.float64), lo.astype(np.float64)
    dec = u <= 1.5
    ref_dec = (up <= 1.5) & (up >= 0)
    s32 = np.minimum((up[ref_dec] - SLACK) / 1.00001, 1.0)    # the fp32 |cc| maximum / std the bound was built from (clamped like the score)
    row = {"N": N, "fft_len": int(2 ** int(np.ceil(np.log2(N)))), "series": S, "decided_frac": float(dec.mean()),
           "spectral_bound_min_margin": float((u[dec] - sc[dec]).min()),
           "refined_upper_min_margin": float((up[ref_dec] - sc[ref_dec]).min()),
           "refined_lower_min_margin": float((sc[lo >= 0] - lo[lo >= 0]).min()) if (lo >= 0).any() else None,
           "fp32_vs_fp64_worst_abs_error": float(np.abs(s32 - sc[ref_dec]).max()),
           "slack": SLACK}
    rows.append(row)
    print(row, flush=True)
    b.close()
    store.close()
out = {"what": "worst fp32-vs-fp64 score error and smallest bound margins per series length, adversarial generator; every margin must stay >= slack/4 "
               "and every error <= slack/4 (tests/test_abi.py checks this file, tests/test_gpu_screen.py the live kernels)",
       "slack": SLACK, "rows": rows}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "screen_error_survey.json"), "w") as f:
    json.dump(out, f, indent=1)
