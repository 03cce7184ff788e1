"""The tile kernels alone at the two bench shapes (eager launches, for ncu): python tools/prof_tiles.py [wd5m|fb237]"""
import sys, torch
sys.path.insert(0, ".")
import kgeb200 as kb
which = sys.argv[1] if len(sys.argv) > 1 else "wd5m"
torch.manual_seed(0)
if which == "wd5m":
    e, b, d, loss = 4_600_000, 2048, 128, kb.lib.LOSS_KL
else:
    e, b, d, loss = 14541, 4096, 128, kb.lib.LOSS_BCE
w = (torch.randn(e, d, device="cuda") * 0.1)
q = (torch.randn(b, d, device="cuda") * 0.1)
lab_off = torch.arange(b + 1, dtype=torch.int64, device="cuda")
lab_col = torch.randint(0, e, (b,), device="cuda")
sh = kb.fused.Shard.full(e)
for it in range(3):
    if loss == kb.lib.LOSS_KL:
        loc, o_sum, _ = kb.fused.flash_forward(q, w, lab_off, lab_col, sh)
        rows, lse = kb.fused.rows_loss(loc, lab_off, loss, 0.0, e)
        dw = torch.empty_like(w)
        kb.fused.fused_backward(q, w, lab_off, lab_col, loss, 0.0, 0.0, lse, 1.0 / b, None, kb.lib.MATH_BF16, sh, dw, want_dq=False, overwrite=True)
    else:
        dw = torch.zeros_like(w)
        kb.fused.fused_backward(q, w, lab_off, lab_col, loss, 0.0, 0.0, None, 1.0 / b, None, kb.lib.MATH_BF16, sh, dw)
    torch.cuda.synchronize()
print("ok", float(dw.abs().sum()))
