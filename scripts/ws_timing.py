"""Diagnostic: per-phase cycle split of slide_ws_kernel (CTA 0), from a -DNBM_WS_TIMING build of the library.
    python scripts/ws_timing.py build      # here (nvcc): writes gpurun_out/libnbm_b200_dbg.so ... no: scripts/_dbg/
    NBM_B200_LIB=scripts/_dbg/libnbm_b200_dbg.so python scripts/ws_timing.py run   # on the GPU box
"""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DBG = os.environ.get("NBM_WS_DBG", os.path.join(ROOT, "scripts", "_dbg", "libnbm_b200_dbg.so"))
if sys.argv[1] == "build":
    sys.path.insert(0, ROOT)
    from birdsoundclassif_b200 import build as B
    os.makedirs(os.path.dirname(DBG), exist_ok=True)
    cmd = [B._nvcc(), *B.NVCC_FLAGS, "-DNBM_WS_TIMING", "-o", DBG, *[os.path.join(B.CSRC, s) for s in B.SOURCES]]
    subprocess.run(cmd, check=True)
    print(DBG)
else:
    os.environ["NBM_B200_LIB"] = DBG
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    from birdsoundclassif_b200 import frontend, synth, _lib
    n = 2646000
    pcm = np.concatenate([synth.synth_pcm(60.0, 100 + i)[:n] for i in range(4)] * 16)
    offs = (np.arange(65) * n).tolist()
    plan = frontend.get_plan()
    flat = torch.from_numpy(pcm).cuda()
    plan.run_batch(flat, offs); torch.cuda.synchronize()
    h = _lib.lib()
    buf = (ctypes.c_ulonglong * 32)()
    h.nbm_debug_ws_timing(buf, 1)
    plan.run_batch(flat, offs); torch.cuda.synchronize()
    h.nbm_debug_ws_timing(buf, 0)
    v = np.array(list(buf), dtype=np.float64)
    names = ["seek/anchor", "wait acc_full", "tmem ld + recur", "barrier A", "build (+wait samples)", "emit", "barrier B"]
    tot = v[:8].sum()
    print("worker warps of CTA 0 (12 warps summed): total %.0f cycles = %.0f per warp" % (tot, tot / 12))
    for nme, x in zip(names, v[:8]):
        print("  %-22s %6.1f %%   %.0f cycles per warp" % (nme, 100 * x / tot, x / 12))
    print("mma warp: wait b_ready %.0f, issue %.0f cycles" % (v[8], v[9]))
    print("fill warps (3): issue loads %.0f, wait s_free %.0f, store %.0f cycles per warp" % (v[16] / 3, v[17] / 3, v[18] / 3))
