"""Role timeline (clock64 stamps of CTA 0) of one tcgen05 conv launch: where do the epilogue / MMA / producer wait?
Needs the instrumented build:  MDB_TIMELINE=1 python -m monodetr_b200.build --force   (the product library has no stamps)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from monodetr_b200 import _lib, tc  # noqa: E402

L = _lib.lib()
L.mdb_debug_set_timeline.argtypes = [ctypes.c_void_p]
L.mdb_debug_set_timeline.restype = None
B, H, W, Cin, Cout = 8, 96, 320, int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 256
tc.set_precision(sys.argv[3] if len(sys.argv) > 3 else "tf32x3")
x = torch.randn(B, H, W, Cin, device="cuda")
w = torch.randn(Cout, Cin, 1, 1, device="cuda") / Cin ** 0.5
wp = tc.pack_weight(w)
for _ in range(3):
    tc.conv2d_forward(x, wp, None, None, 1, 1, 1, 0)
dbg = torch.zeros(512, dtype=torch.int64, device="cuda")
L.mdb_debug_set_timeline(dbg.data_ptr())
tc.conv2d_forward(x, wp, None, None, 1, 1, 1, 0)
torch.cuda.synchronize()
L.mdb_debug_set_timeline(None)
d = dbg.cpu().tolist()
t0 = min(v for v in d if v > 0)
print("EPILOGUE (warp 0 of 4) per tile: start, got tmem_full, then per 32-col chunk [tmem ld done, transposed, stored]")
for lt in range(8):
    r = d[lt * 16:lt * 16 + 14]
    print(lt, [v - t0 if v else None for v in r])
print("MMA per tile: [start wait tmem_empty, got it, main loop issued+committed]")
for lt in range(8):
    r = d[256 + lt * 4:256 + lt * 4 + 3]
    print(lt, [v - t0 if v else None for v in r])
print("PRODUCER: time of TMA issue per stage")
print([v - t0 if v else None for v in d[384:384 + 24]])
