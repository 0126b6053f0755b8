#!/usr/bin/env python
"""Timeline of bench.py's end-to-end loop (config 2, one GPU, two contexts in flight): CUDA events around the
upload, the sweep and the download of every step, printed as offsets from the first sweep's start."""
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
nlanes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
WAIT = len(sys.argv) > 2 and sys.argv[2] == 'wait'      # order the other lane's finalize in front of the sweep
steps = 12
outs = [smk.alloc_pinned((rows, G)) for _ in range(nlanes)]
lanes = []
for _ in range(nlanes):
    st = torch.cuda.Stream(device=dev)
    ctx = smk.Context(I)
    ctx.set_stream(st.cuda_stream)
    ctx.upload(src, flux0, sig)
    lanes.append((st, ctx))
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
host = []


def enqueue(k):
    st, ctx = lanes[k % nlanes]
    with torch.cuda.stream(st):
        ev[k][0].record(st)
        ctx.upload_async(src, flux0, sig)
        if WAIT:
            ctx.wait_finalized(lanes[(k - 1) % nlanes][1])
        ev[k][1].record(st)
        ctx.run_async(0, I.n_tracks)
        ev[k][2].record(st)
        ctx.download_flux_rows_async(0, rows, outs[k % nlanes])
        ev[k][3].record(st)


torch.cuda.synchronize(dev)
t0 = time.perf_counter()
for k in range(steps):
    host.append(time.perf_counter() - t0)
    enqueue(k)
    if k >= nlanes - 1:
        lanes[(k - (nlanes - 1)) % nlanes][0].synchronize()
torch.cuda.synchronize(dev)
total = time.perf_counter() - t0
base = ev[0][0]
print(f"lanes={nlanes}: {total / steps * 1e3:.3f} ms/step")
print("step  host_enqueue  upload_start  sweep_start(after upload)  sweep_end  download_end   [ms]   gap_to_prev_sweep_end")
prev_end = None
for k in range(steps):
    t = [base.elapsed_time(e) for e in ev[k]]
    gap = (t[1] - prev_end) if prev_end is not None else float('nan')
    print(f"{k:3d}  {host[k]*1e3:9.3f}  {t[0]:9.3f}  {t[1]:9.3f}  {t[2]:9.3f}  {t[3]:9.3f}   sweep {t[2]-t[1]:7.3f}  upload {t[1]-t[0]:6.3f}  dl {t[3]-t[2]:6.3f}  gap {gap:6.3f}")
    prev_end = t[2]
