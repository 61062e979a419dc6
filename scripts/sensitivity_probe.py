"""How much does the stand-in detector amplify a tolerance-level change of its input?  Reference CPU tiles vs GPU tiles of the
same clip through the same (unpatched) reference model: input difference, difference of the RPN scores / FPN features, and
the box-level match; with a control (reference tiles + uniform noise of the same size)."""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

ref_shims.install()
from birdsoundclassif_b200 import frontend, run_detection as rd, synth  # noqa: E402
from oracle import frontend_oracle as fo  # noqa: E402

d = tempfile.mkdtemp()
synth.write_standin_checkpoint(d, seed=0, sharpen=float(sys.argv[1]) if len(sys.argv) > 1 else 400.0)
m, a = ref_shims.ref("nbm_model.run_detection").load_model(d)
pcm = synth.synth_pcm(12.5, 52, calls_per_s=6.0)
ref = torch.from_numpy(np.stack(fo.process(pcm).tiles)).float().cuda()
fp = frontend.File_Processor("x.wav")
gpu, _ = fp.process_pcm(torch.from_numpy(pcm).cuda())
print("input: max |gpu - ref| %.2e rms %.2e" % ((gpu - ref).abs().max().item(), (gpu - ref).pow(2).mean().sqrt().item()))
g = torch.Generator(device="cuda").manual_seed(1)
variants = {"gpu tiles": gpu, "ref + 1 ulp noise (6e-8)": ref + (torch.rand(ref.shape, device="cuda", generator=g) - 0.5) * 1.2e-7,
            "ref + 1e-6 noise": ref + (torch.rand(ref.shape, device="cuda", generator=g) - 0.5) * 2e-6}
with torch.no_grad():
    o0 = m.forward_first_stage(ref[:4, None])
    for name, t in variants.items():
        o1 = m.forward_first_stage(t[:4, None])
        f0, f1 = o0["fpn_out"][2], o1["fpn_out"][2]
        s0, s1 = o0["rpn_cls_scores"], o1["rpn_cls_scores"]
        same_rois = (o0["rois"] == o1["rois"]).all(dim=-1).float().mean().item()
        print(f"{name:28s} fpn[2] rel diff {((f1 - f0).norm() / f0.norm()).item():.2e} (|f| max {f0.abs().max().item():.2e}) | rpn score max diff "
              f"{(s1 - s0).abs().max().item():.2e} | identical rois {same_rois:.2f}")
