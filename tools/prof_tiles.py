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
if which == "wd5m-upd":
    # the dense table-gradient kernel with Adagrad in its flush (kgeb_fused_bwd_update) against the storing kernel, eager
    import os
    lib = kb.lib
    e, b, d, loss = 4_600_000, 2048, 128, kb.lib.LOSS_KL
    w = (torch.randn(e, d, device="cuda") * 0.1)
    q = (torch.randn(b, d, device="cuda") * 0.1)
    lab_off = torch.arange(b + 1, dtype=torch.int64, device="cuda")
    lab_col = torch.randint(0, e, (b,), device="cuda")
    sh = kb.fused.Shard.full(e)
    loc, o_sum, _ = kb.fused.flash_forward(q, w, lab_off, lab_col, sh)
    rows, lse = kb.fused.rows_loss(loc, lab_off, loss, 0.0, e)
    mirror = w.bfloat16()
    state = torch.full_like(w, 0.1)
    slot_of = torch.full((e,), -1, dtype=torch.int32, device="cuda")
    gbuf = torch.zeros(8192, d, device="cuda")
    ws = torch.empty(lib.load().kgeb_fused_workspace_bytes(b, d, e, 1), dtype=torch.uint8, device="cuda")
    st = lib.stream_ptr(w)
    def upd():
        lib.call("kgeb_fused_bwd_update", loss, q.data_ptr(), b, d, w.data_ptr(), 0, e, e, lab_off.data_ptr(), 0.0, 0.0,
                 lse.data_ptr(), 1.0 / b, None, mirror.data_ptr(), state.data_ptr(), 1e-3, 1e-10, slot_of.data_ptr(),
                 gbuf.data_ptr(), None, ws.data_ptr(), ws.numel(), st)
    dw = torch.empty_like(w)
    def plain():
        kb.fused.fused_backward(q, w, lab_off, lab_col, loss, 0.0, 0.0, lse, 1.0 / b, None, kb.lib.MATH_BF16, sh, dw,
                                want_dq=False, overwrite=True)
    def timed(fn, n=3):
        fn(); torch.cuda.synchronize()
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        z.record(); torch.cuda.synchronize()
        return a.elapsed_time(z) / n
    modes = [int(x) for x in os.environ.get("PROF_UPD_MODES", "0").split(",")]
    print("storing kernel  %.3f ms" % timed(plain), flush=True)
    w0 = w.clone()
    for m in modes:      # (clr = 0: the table stays what the log-sum-exp was computed for; mode 32 adds into it: restored)
        os.environ["KGEB_UPD_DEBUG"] = str(m)
        t = timed(upd)
        w.copy_(w0); state.fill_(0.1); mirror.copy_(w0)
        print("update kernel, debug=%d  %.3f ms" % (m, t), flush=True)
    print("storing kernel  %.3f ms" % timed(plain), flush=True)
    sys.exit(0)
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
