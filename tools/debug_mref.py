"""mref of the flash pass on the problem of test_flash_reports_scores_outside_its_window (debug)"""
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import kgeb200 as kb
import test_gpu_shapes as T
e, r, d, b = 14541, 50, 128, 256
m, q, lab_off, lab_col = T._problem(kb, "distmult", e, r, d, b, 2, seed=35, empty_row=False)
w = m.get_s_embedder().weight.detach().clone()
hot = torch.tensor([11, 7001, 14001], device="cuda")
w[hot] *= 1000.0
shard = kb.fused.Shard.full(e)
loc, o_sum, status = kb.fused.flash_forward(q, w, lab_off, lab_col, shard)
x = q @ w.t()
print("status", status.tolist(), "mref min/max", loc[:, 0].min().item(), loc[:, 0].max().item(), "finite", torch.isfinite(loc).all().item())
print("row max min/max", x.max(1).values.min().item(), x.max(1).values.max().item())
print("gap max", (x.max(1).values - loc[:, 0]).max().item(), "l min/max", loc[:, 1].min().item(), loc[:, 1].max().item())
print("o_sum finite", torch.isfinite(o_sum).all().item(), o_sum.abs().max().item())
