"""Diagnostic (NBM_WS_TIMING build): does the tiling kernel really run beside the slide kernel in fused mode?"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["NBM_B200_LIB"] = os.path.join(ROOT, "scripts", "_dbg", "libnbm_b200_dbg.so")
os.environ["NBM_FRONTEND_FUSED"] = "1"
sys.path.insert(0, ROOT)
import numpy as np, torch
from birdsoundclassif_b200 import frontend, _lib
n, clips = 2646000, 256
plan = frontend.get_plan()
pcm = torch.randint(-3000, 3000, (clips * n,), dtype=torch.int16, device='cuda')
offs = [i * n for i in range(clips + 1)]
_, tile_off, _ = plan.query_batch([n] * clips)
tiles = torch.empty((tile_off[-1], 1, 375, 1024), dtype=torch.float32, device='cuda')
for _ in range(3):
    plan.run_batch(pcm, offs, out=tiles)
torch.cuda.synchronize()
h = _lib.lib()
a = (ctypes.c_ulonglong * 32)(); b = (ctypes.c_ulonglong * 8)()
h.nbm_debug_ws_timing(a, 0); h.nbm_debug_follow_timing(b)
s0, s1 = a[24], a[25]
print("slide: 0 .. %.2f ms" % ((s1 - s0) / 1e6))
print("follow first block start %.2f ms, file 0 ready %.2f ms, middle file ready %.2f ms, last block end %.2f ms" %
      tuple((x - s0) / 1e6 for x in b[:4]))
ff = (ctypes.c_ulonglong * 4096)(); h.nbm_debug_follow_files(ff)
print("file: min/max block scheduled ms | file complete ms")
for f in list(range(0, clips, max(1, clips // 32))) + [clips - 1]:
    print("%4d  %8.3f  %8.3f" % (f, (ff[f] - s0) / 1e6, (ff[2048 + f] - s0) / 1e6))
