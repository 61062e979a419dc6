"""Quick front-end accuracy check on the GPU against the oracle (diagnostic, not a test)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from birdsoundclassif_b200 import frontend, synth
from oracle import frontend_oracle as fo

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 5.0
pcm = synth.synth_pcm(secs, 77)
plan = frontend.get_plan()
print("impl:", plan.impl)
tiles, mm = plan.run(torch.from_numpy(pcm).cuda())
torch.cuda.synchronize()
db = plan.spectrogram_view(0).cpu().numpy().astype(np.float64)
ref = fo.db_spectrogram(fo.to_float(pcm), fo.derive_params())[0]
e = np.abs(db - ref)
print("shape", db.shape, "finite", np.isfinite(db).all())
print(f"dB err: max {e.max():.5f} p99.99 {np.quantile(e, 0.9999):.2e} p99 {np.quantile(e, 0.99):.2e} median {np.median(e):.2e}")
bad = np.argwhere(e > 0.5)
print("n>0.5dB:", len(bad), "first:", bad[:8].tolist())
if len(bad):
    cols = np.unique(bad[:, 1]); rows = np.unique(bad[:, 0])
    print("bad cols (first 40):", cols[:40].tolist(), "n", len(cols)); print("bad rows (first 40):", rows[:40].tolist(), "n", len(rows))
    print("sample gpu/ref:", db[bad[0][0], bad[0][1]], ref[bad[0][0], bad[0][1]])
r = fo.process(pcm)
t = tiles[:, 0].cpu().numpy().astype(np.float64)
err = np.abs(t - np.stack(r.tiles))
print(f"tiles: max {err.max():.3e} frac>1e-4 {(err > 1e-4).mean():.2e} rms {np.sqrt((err**2).mean()):.2e}; smin {mm[0].item():.4f} vs {r.s_min:.4f}; smax {mm[1].item():.4f} vs {r.s_max:.4f}")
