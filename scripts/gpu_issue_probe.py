#!/usr/bin/env python3
"""Issue-rate probe: IPC per SM sub-partition (SMSP) of dependent-free FFMA / MUFU / Philox loops vs warps per SMSP."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rwm_pt_pytorch_b200 import _lib
lib = _lib.load()
clk = 1.965e9
for kind, name in ((0, "FFMA"), (1, "MUFU.EX2"), (2, "Philox(IMAD+LOP3)")):
    for blocks, threads in ((148 * 4, 32), (148 * 7, 32), (148 * 8, 32), (148 * 4, 128), (148 * 8, 128), (148 * 8, 256)):
        r = C.c_double()
        _lib.check(lib.rwmpt_probe_issue(kind, blocks, threads, 2048 if kind != 1 else 512, C.byref(r)))
        wps = blocks * threads / 32 / (148 * 4)
        print(f"{name:18s} warps/SMSP {wps:5.2f}: {r.value / (148 * 4) / clk:.3f} warp-instr/clk/SMSP (at {clk/1e9} GHz)")
