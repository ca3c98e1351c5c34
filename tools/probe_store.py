"""How fast can W warps per SM push the conv epilogue's store pattern?  (per-SM write throughput vs storing warps)"""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monodetr_b200 import _lib
L = _lib.lib()
L.mdb_debug_probe_store.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
rows, ld = 245760, 256
out = torch.empty(rows, ld, device="cuda")
tiles = 26
for ctas_per_sm in (1, 2):
    for warps in (1, 2, 4, 8, 16):
        ctas = 148 * ctas_per_sm
        for _ in range(3):
            L.mdb_debug_probe_store(out.data_ptr(), ld, rows, ctas, warps, tiles, None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            L.mdb_debug_probe_store(out.data_ptr(), ld, rows, ctas, warps, tiles, None)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 10 * 1e3
        byt = ctas * tiles * 128 * 128 * 4
        print(f"ctas/SM {ctas_per_sm} warps/CTA {warps:2d}: {us:8.1f} us  {byt / us / 1e3:8.0f} GB/s  ({byt / us / 1e3 / 148 / 1.965:5.1f} B/clk/SM)")
