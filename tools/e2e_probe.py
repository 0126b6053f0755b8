#!/usr/bin/env python
"""Where does the end-to-end step lose time against the device-resident step?  Variants of bench.py's e2e loop
(config 2, one GPU): number of steps, lanes in flight, and which of upload / download is in the loop."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import smk_b200 as smk

dev = torch.device("cuda:0")
G = 128
I = smk.Input(source_2D_regions=5000, segments=100_000_000, egroups=G, seg_per_thread=100, seed=42).finalize()
R, F = I.source_3D_regions, I.fine_axial_intervals
rows = R * F
src = smk.alloc_pinned((rows, G)); flux0 = smk.alloc_pinned((rows, G)); sig = smk.alloc_pinned((R, G))
rng = np.random.default_rng(42)
src[...] = rng.random(src.shape, dtype=np.float32); flux0[...] = rng.random(flux0.shape, dtype=np.float32)
sig[...] = rng.random(sig.shape, dtype=np.float32)


def run(steps, nlanes, up=True, down=True, label=""):
    outs = [smk.alloc_pinned((rows, G)) for _ in range(nlanes)]
    lanes = []
    for _ in range(nlanes):
        st = torch.cuda.Stream(device=dev)
        ctx = smk.Context(I)
        ctx.set_stream(st.cuda_stream)
        ctx.upload(src, flux0, sig)
        lanes.append((st, ctx))

    def enqueue(k):
        st, ctx = lanes[k % nlanes]
        with torch.cuda.stream(st):
            if up:
                ctx.upload_async(src, flux0, sig)
            else:
                ctx.reset_tallies()
            ctx.run_async(0, I.n_tracks)
            if down:
                ctx.download_flux_rows_async(0, rows, outs[k % nlanes])

    for k in range(nlanes):
        enqueue(k); lanes[k % nlanes][0].synchronize()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for k in range(steps):
        enqueue(k)
        if k >= nlanes - 1:
            lanes[(k - (nlanes - 1)) % nlanes][0].synchronize()
    for k in range(max(0, steps - (nlanes - 1)), steps):
        lanes[k % nlanes][0].synchronize()
    torch.cuda.synchronize(dev)
    sec = time.perf_counter() - t0
    print(f"{label or ''} steps={steps} lanes={nlanes} up={up} down={down}: {sec / steps * 1e3:.3f} ms/step "
          f"{I.segments * G * steps / sec:.4e} int/s", flush=True)
    for _, c in lanes:
        c.close()


for rep in range(2):
    run(10, 1, False, False, "resident, 1 lane")
    run(20, 2, False, False, "resident")
    run(10, 2)
    run(20, 2)
    run(40, 2)
    run(20, 3)
    run(20, 2, True, False)
    run(20, 2, False, True)
