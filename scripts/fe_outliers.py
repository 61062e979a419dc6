"""Diagnostic: which quantity does the float32 transform's error scale with?  For the pixels of a few clips with the largest
dB error (refinement switched off), prints the LINEAR error of |X| relative to (a) the white-spectrum level of the chain's
largest frame energy, (b) the rectangular-window DFT magnitude |R| at the pixel and its neighbours (what the Hann
combination cancels), (c) the largest |R[k]| along the chain from its anchor."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["NBM_REFINE_REL_DB"] = "-200"
from birdsoundclassif_b200 import frontend, synth  # noqa: E402
from oracle import frontend_oracle as fo  # noqa: E402

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 3
plan = frontend.FrontendPlan()
p = fo.derive_params()
N, hop, lo = 1324, 132, 16
stats = []
for i in range(n_clips):
    pcm = synth.synth_pcm(20.0, 300 + i, calls_per_s=3.0 + i)
    x = fo.to_float(pcm).astype(np.float64)
    ref = fo.db_spectrogram(x.astype(np.float32), p)[0]
    plan.run(torch.from_numpy(pcm).cuda())
    torch.cuda.synchronize()
    db = plan.spectrogram_view(0).cpu().numpy().astype(np.float64)
    T = ref.shape[1]
    xp = np.pad(x, N // 2)
    frames = np.lib.stride_tricks.as_strided(xp, shape=(T, N), strides=(hop * 8, 8))
    R = np.abs(np.fft.rfft(frames, axis=1))[:, lo - 1:lo + 376].T          # rect window, bins lo-1 .. lo+375
    W = (frames ** 2).sum(axis=1)
    white = np.sqrt(0.375 * W)
    xa, xr = 10 ** (db / 20), 10 ** (ref / 20)
    d = np.abs(xa - xr)
    rel_db_err = np.abs(db - ref)
    Rloc = np.maximum(np.maximum(R[:-2], R[1:-1]), R[2:])                  # max |R| over k-1, k, k+1
    chain = np.arange(T) // 32
    nch = chain.max() + 1
    wmax = np.array([white[max(0, c * 32 - 1):c * 32 + 34].max() for c in range(nch)])[chain]
    Rch = np.stack([Rloc[:, max(0, c * 32 - 1):c * 32 + 34].max(axis=1) for c in range(nch)], axis=1)[:, chain]
    a, b, c_ = d / wmax[None], d / Rloc, d / Rch
    print(f"clip {i}: linear error / chain white level: rms {np.sqrt((a ** 2).mean()):.2e} max {a.max():.2e} | / local |R|: rms "
          f"{np.sqrt((b ** 2).mean()):.2e} max {b.max():.2e} | / chain max |R[k]|: rms {np.sqrt((c_ ** 2).mean()):.2e} max {c_.max():.2e}")
    idx = np.argsort(-rel_db_err, axis=None)[:10]
    for j in idx:
        bb, t = np.unravel_index(j, d.shape)
        print(f"   bin {bb:3d} frame {t:5d} pos {t % 32:2d}: ref {ref[bb, t]:7.2f} dB, dB err {rel_db_err[bb, t]:.4f} | d/white_chain {a[bb, t]:.2e}  "
              f"d/Rloc {b[bb, t]:.2e}  d/Rchain {c_[bb, t]:.2e} | Rloc/white {Rloc[bb, t] / wmax[t]:.1f} Rchain/white {Rch[bb, t] / wmax[t]:.1f}")
    # error growth along the chain: rms of d / Rch by position in the chain
    pos = np.arange(T) % 64
    g = [np.sqrt((c_[:, pos == q] ** 2).mean()) for q in range(64)]
    print("   rms(d / chain max |R|) by frame position in the 64-frame group:", " ".join(f"{v:.1e}" for v in g[::4]))
