#!/usr/bin/env python
"""Per-tile timeline of the backward tile kernels' pipeline hand-offs (tuning tool; needs the KGEB_TRACE build:
tools/build_variant.sh trace tc_bwd -DKGEB_TRACE, KGEB200_LIB=kge-1_b200/variants/libkgeb200_trace.so).
Runs the dTable and the dQ kernel of the bench shape once each and prints, for block 0, the clock of every event:
  MMA1 issuer: 1 streamed tile landed, 2 S buffer free, 3 MMA1 issued+committed     MMA2 issuer: 4 G ready, 5 MMA2 issued
  epilogue warps (lane 0): 8 tile start, 9 S ready, 10 S in registers, 11 G computed, 12 G buffer free, 13 G stored."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kgeb200 as kb  # noqa: E402

B, E, d = 4096, 14541, 128
dev = torch.device("cuda")
torch.manual_seed(0)
Q = (torch.randn(B, d, device=dev) * 0.3)
ent = (torch.randn(E, d, device=dev) * 0.1)
mirror = ent.bfloat16()
L = kb.lib
lib = L.load()
ws = torch.empty(lib.kgeb_fused_workspace_bytes(B, d, E, 1), dtype=torch.uint8, device=dev)
off0 = torch.zeros(B + 1, dtype=torch.int64, device=dev)
col0 = torch.zeros(1, dtype=torch.int64, device=dev)
lse = torch.zeros(B, device=dev)
dQ = torch.empty(B, d, device=dev)
g = torch.zeros(E, d, device=dev)
rowstat = torch.empty(B, 4, device=dev)
loss_kind = L.LOSS_BCE if len(sys.argv) < 2 or sys.argv[1] == "bce" else L.LOSS_KL


def bwd(dq, dt, stats=False):
    L.call("kgeb_fused_bwd", loss_kind, L.MATH_BF16, Q.data_ptr(), B, d, ent.data_ptr(), 0, E, E, off0.data_ptr(),
           col0.data_ptr(), 0, None, 0.0, 0.0, lse.data_ptr(), 1.0 / B, None, mirror.data_ptr(),
           dQ.data_ptr() if dq else None, g.data_ptr() if dt else None, rowstat.data_ptr() if stats else None,
           ws.data_ptr(), ws.numel(), L.stream_ptr(ent))


lib.kgeb_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
lib.kgeb_debug_trace.restype = ctypes.c_int
buf = np.zeros(32 << 11, dtype=np.uint64)
NAMES = {1: "str_full", 2: "s_empty", 3: "mma1_issued", 4: "g_full", 5: "mma2_issued", 8: "tile_start", 9: "s_ready",
         10: "s_loaded", 11: "g_computed", 12: "g_buf_free", 13: "g_stored"}
for name, args in (("dTable", (False, True)), ("dQ", (True, False)), ("dQ+stats", (True, False, True))):
    if name == "dQ+stats" and loss_kind != L.LOSS_BCE:
        continue
    for _ in range(2):
        bwd(*args)
        n = lib.kgeb_debug_trace(buf.ctypes.data, len(buf))
    used = buf != 0
    warp = (np.arange(len(buf)) >> 11)[used]
    ev = buf[used]
    n = len(ev)
    code, tile, clk = ((ev >> 52) & 63).astype(int), ((ev >> 40) & 0xfff).astype(int), (ev & 0xffffffffff).astype(np.int64)
    t0 = clk.min()
    print(f"===== {name}: {n} events, span {clk.max() - t0} clk, tiles {tile.max() + 1}")
    # per-tile period seen by the MMA1 issuer
    m = code == 3
    order = np.argsort(clk[m])
    tt = clk[m][order]
    print("MMA1 issue period (clk), median / p90:", int(np.median(np.diff(tt))), int(np.percentile(np.diff(tt), 90)))
    for lo in (20, 21, 22, 23):
        sel = tile == lo
        rows = sorted(zip(clk[sel] - t0, warp[sel], code[sel]))
        print(f"--- tile {lo}")
        for c, w, k in rows:
            if w in (1, 3) or w in (4, 8, 12, 16):     # issuers + the first warp of each epilogue column part
                print(f"   {c:8d}  warp {w:2d}  {NAMES.get(k, k)}")
    # average phase durations of epilogue warp 4 / 12 (one of each group)
    for w in (4, 12):
        d_ = {}
        for u in range(8, int(tile.max()) - 4):
            sel = (warp == w) & (tile == u)
            if sel.sum() < 6:
                continue
            ck = dict(zip(code[sel], clk[sel]))
            if all(k in ck for k in (8, 9, 10, 11, 12, 13)):
                for a, b, nm in ((8, 9, "wait S"), (9, 10, "load S"), (10, 11, "compute G"), (11, 12, "wait G buffer"), (12, 13, "store G")):
                    d_.setdefault(nm, []).append(ck[b] - ck[a])
        print(f"epilogue warp {w}: mean clk per phase:", {k: int(np.mean(v)) for k, v in d_.items()}, "tiles", len(next(iter(d_.values()), [])))
