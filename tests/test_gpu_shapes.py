"""Parity at the BASELINE.json shapes (the golden vectors are E=53; these are the shapes the bench runs):

* FB15k-237 shape, E = 14,541 / 4,096 query rows / d = 128 (114 entity tiles with a ragged last one, 128 dQ jobs):
  fused forward statistics, dQ and dTable of the tensor-tile paths against the fp32 CUDA-core path on ALL rows, and the
  fp32 path against float64 on a row / entity sample (the oracle's arithmetic in double precision).
* Wikidata5M shape, E = 4,600,000 (int64 offsets, 35,938 tiles, multi-wave persistence, TMA reduce-add flush of a 2.4 GB
  gradient): row-sampled scores / log-sum-exp / dQ / dTable rows against float64 (`oracle.kge_oracle.score_emb` for the
  scores) and rank counts on dyadic tables, which must be bit-exact in fp32 AND on the TF32 tiles.
Bounds (stated per assert): fp32 2e-5, TF32 3e-3, BF16 1e-2 (statistics) / 1.5e-2 (gradients), relative to max |ref|."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kb():
    import kgeb200
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return kgeb200


def close(got, ref, rtol, what):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    assert got.shape == ref.shape, what
    diff = (got - ref).abs()
    err = diff.max().item()
    where = np.unravel_index(int(diff.argmax()), tuple(diff.shape)) if diff.numel() else ()
    assert err <= rtol * max(ref.abs().max().item(), 1e-30), \
        f"{what}: max err {err:.3e} at {where} vs bound {rtol * ref.abs().max().item():.3e}"


def _problem(kb, model, e, r, d, b, nlab_max, seed, empty_row=True):
    """Queries of a real model (normal(0, 0.1) tables), rows sp_ then _po, and a ragged label CSR."""
    torch.manual_seed(seed)
    m = kb.KgeModel(model, e, r, d).cuda()
    gen = torch.Generator().manual_seed(seed + 1)
    a = torch.randint(0, e, (b,), generator=gen).cuda()
    p = torch.randint(0, r, (b,), generator=gen).cuda()
    with torch.no_grad():
        q = torch.cat((m.queries(kb.lib.SP_, a[: b // 2], p[: b // 2]), m.queries(kb.lib._PO, a[b // 2:], p[b // 2:]))).contiguous()
        q *= 8.0          # scores of a partly trained model (O(1)), so that sigmoid / softmax are not flat
    n = torch.randint(0 if empty_row else 1, nlab_max + 1, (b,), generator=gen)
    # empty rows (BCE only: a KvsAll / 1vsAll query always has a label, and the reference's KL of an all-zero label row
    # is 0 by F.normalize's eps, loss.py:211-213 -- not a case its jobs can produce) and a full one
    n[0], n[1] = (0 if empty_row else 1), nlab_max
    lab_off = torch.zeros(b + 1, dtype=torch.int64)
    lab_off[1:] = torch.cumsum(n, 0)
    cols = [torch.randperm(min(e, 100000), generator=gen)[: int(k)].sort().values * (e // min(e, 100000)) for k in n.tolist()]
    cols[1][-1] = e - 1                            # the last entity (ragged tile)
    lab_col = torch.cat(cols).long()
    return m, q, lab_off.cuda(), lab_col.cuda()


def _float64_rows(q, w, lab_off, lab_col, rows, loss, offset, inv_batch):
    """Reference arithmetic in double precision for a sample of query rows: per-row loss, lse, dQ rows."""
    out_loss, out_lse, out_dq = [], [], []
    wd = w.double()
    for r in rows:
        qd = q[r].double().requires_grad_(True)
        x = wd @ qd
        cols = lab_col[lab_off[r]:lab_off[r + 1]]
        if loss == "kl":
            lse = torch.logsumexp(x, 0)
            val = (lse - x[cols].sum() / max(len(cols), 1) - np.log(max(len(cols), 1))) if len(cols) else x.sum() * 0
            out_lse.append(lse.item())
        else:
            val = torch.nn.functional.softplus(x + offset).sum() - (x[cols] + offset).sum()
        (val * inv_batch).backward()
        out_loss.append(val.item() * inv_batch)
        out_dq.append(qd.grad.float())
    return torch.tensor(out_loss), torch.tensor(out_lse), torch.stack(out_dq)


def _float64_table_rows(q, w, lab_off, lab_col, ents, loss, offset, inv_batch, lse):
    """dTable rows of a sample of entities in double precision: sum_q G[q, e] * Q[q]."""
    qd = q.double()
    x = qd @ w[ents].double().t()                                 # [B, n_sample]
    if loss == "kl":
        nnz = (lab_off[1:] - lab_off[:-1]).double()
        g = torch.exp(x - lse.double()[:, None]) * (nnz > 0).double()[:, None]
        tval = torch.where(nnz > 0, 1.0 / nnz.clamp(min=1), torch.zeros_like(nnz))
    else:
        g = torch.sigmoid(x + offset)
        tval = torch.ones(q.shape[0], dtype=torch.float64, device=q.device)
    rows = torch.repeat_interleave(torch.arange(q.shape[0], device=q.device), lab_off[1:] - lab_off[:-1])
    for j, e in enumerate(ents.tolist()):
        hit = rows[lab_col == e]
        if len(hit):
            g[:, j].index_add_(0, hit, -tval[hit])
    return (g.t() @ qd * inv_batch).float()


@pytest.mark.parametrize("loss", ["bce", "kl"])
def test_fb15k237_shape_fused_kernels(kb, loss):
    e, r, d, b = 14541, 237, 128, 4096
    m, q, lab_off, lab_col = _problem(kb, "complex", e, r, d, b, 6, seed=3, empty_row=(loss == "bce"))
    w = m.get_s_embedder().weight.detach()
    kind = kb.lib.LOSS_KL if loss == "kl" else kb.lib.LOSS_BCE
    offset = 0.1 if loss == "bce" else 0.0
    shard = kb.fused.Shard.full(e)
    res = {}
    for name, mode in (("fp32", kb.lib.MATH_FP32), ("tf32", kb.lib.MATH_TF32), ("bf16", kb.lib.MATH_BF16)):
        st = kb.fused.fused_rowstats(q, w, lab_off, lab_col, kind, 0.0, offset, mode, shard)
        rows, lse = kb.fused.rows_loss(st, lab_off, kind, 0.0, e)
        dw = torch.zeros_like(w)
        dq = kb.fused.fused_backward(q, w, lab_off, lab_col, kind, 0.0, offset, lse, 1.0 / b, None, mode, shard, dw)
        res[name] = (rows / b, lse, dq, dw)
    # fp32 CUDA-core path against float64 on a sample of rows / entities
    sample = [0, 1, 2, 127, 128, 2047, 2048, 4095]
    l64, lse64, dq64 = _float64_rows(q, w, lab_off, lab_col, sample, loss, offset, 1.0 / b)
    close(res["fp32"][0][sample], l64, 2e-5, f"{loss} fp32 row losses vs float64")
    close(res["fp32"][2][sample], dq64, 2e-5, f"{loss} fp32 dQ rows vs float64")
    if loss == "kl":
        close(res["fp32"][1][sample], lse64, 2e-6, "kl fp32 lse vs float64")
    ents = torch.tensor([0, 1, 63, 64, 127, 128, 14463, 14464, 14539, 14540, int(lab_col[3])], device="cuda")
    dt64 = _float64_table_rows(q, w, lab_off, lab_col, ents, loss, offset, 1.0 / b, res["fp32"][1] if loss == "kl" else None)
    close(res["fp32"][3][ents], dt64, 2e-5, f"{loss} fp32 dTable rows vs float64")
    # tensor-tile paths against the fp32 path on ALL rows
    for name, st_tol, g_tol in (("tf32", 3e-3, 3e-3), ("bf16", 1e-2, 1.5e-2)):
        close(res[name][0], res["fp32"][0], st_tol, f"{loss} {name} row losses (all {b} rows)")
        close(res[name][2], res["fp32"][2], g_tol, f"{loss} {name} dQ (all rows)")
        close(res[name][3], res["fp32"][3], g_tol, f"{loss} {name} dTable (all {e} rows)")


def test_fb15k237_shape_captured_step_matches_autograd_fp32(kb):
    """The bench's configuration (ComplEx KvsAll + BCE, B = 4096, bf16 tiles, CUDA graph) against the fp32 autograd flow
    on the same batch: loss 1e-3 relative, post-Adagrad tables 2e-3 of the learning rate scale."""
    g = kb.graph.synthetic_graph("fb15k-237", seed=0)
    e, r, d, b = g["num_entities"], g["num_relations"], 128, 4096
    idx = [kb.index.KvsAllIndex(g["train"], "sp"), kb.index.KvsAllIndex(g["train"], "po")]
    ids = torch.from_numpy(np.random.default_rng(5).choice(len(idx[0]) + len(idx[1]), b, replace=False)).cuda()
    a, p, rc, lab_off, lab_col, _ = kb.index.kvsall_batch(idx[0], idx[1], ids)
    rows = torch.repeat_interleave(torch.arange(b, device="cuda"), lab_off[1:] - lab_off[:-1])
    batch = {"queries": torch.where(rc[:, None] == 0, torch.stack((a, p), 1), torch.stack((p, a), 1)),
             "label_coords": torch.stack((rows, lab_col), 1).int(), "query_type_indexes": rc.long().cpu()}
    torch.manual_seed(0)
    ref = kb.KgeModel("complex", e, r, d).cuda()
    new = kb.KgeModel("complex", e, r, d).cuda()
    new.load_state_dict(ref.state_dict())
    mk = lambda m, mode: kb.TrainingJobKvsAll(m, kb.optim.create("Adagrad", m.parameters(), lr=0.2, initial_accumulator_value=0.1),  # noqa: E731
                                              kb.KgeLoss.create("bce"), e, r, math_mode=mode)
    jr, jn = mk(ref, kb.lib.MATH_FP32), mk(new, kb.lib.MATH_BF16)
    jn.enable_graph_step(b, int(lab_col.numel()))
    ra, rn = jr.step(0, batch), jn.step(0, batch)
    assert rn.total_loss == pytest.approx(ra.total_loss, rel=1e-3)
    assert rn.avg_loss == pytest.approx(ra.avg_loss, rel=1e-3)
    for x, y in ((new.get_s_embedder().weight, ref.get_s_embedder().weight), (new.get_p_embedder().weight, ref.get_p_embedder().weight)):
        assert (x - y).abs().max().item() <= 0.2 * 2e-2, (x - y).abs().max().item()


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_wikidata5m_shape_kl_rows_against_float64(kb, mode):
    """DistMult 1vsAll + KL at E = 4.6 M: log-sum-exp / loss / dQ of sampled query rows and dTable rows of sampled entities
    (first / last tile, a label entity) against float64.  fp32 CUDA-core tiles with a small batch (they are 20x slower),
    bf16 tensor tiles with 512 query rows."""
    e, r, d = 4_600_000, 822, 128
    b = 512 if mode == "bf16" else 128
    m, q, lab_off, lab_col = _problem(kb, "distmult", e, r, d, b, 1, seed=11)
    lab_off = torch.arange(b + 1, dtype=torch.int64, device="cuda")            # 1vsAll: exactly one label per row
    lab_col = torch.randint(0, e, (b,), generator=torch.Generator().manual_seed(2)).cuda()
    lab_col[0], lab_col[1] = e - 1, 0
    w = m.get_s_embedder().weight.detach()
    math_mode = kb.lib.MATH_BF16 if mode == "bf16" else kb.lib.MATH_FP32
    st_tol, g_tol = (1e-2, 1.5e-2) if mode == "bf16" else (2e-5, 2e-5)
    shard = kb.fused.Shard.full(e)
    st = kb.fused.fused_rowstats(q, w, lab_off, lab_col, kb.lib.LOSS_KL, 0.0, 0.0, math_mode, shard)
    rows, lse = kb.fused.rows_loss(st, lab_off, kb.lib.LOSS_KL, 0.0, e)
    dw = torch.zeros_like(w)
    dq = kb.fused.fused_backward(q, w, lab_off, lab_col, kb.lib.LOSS_KL, 0.0, 0.0, lse, 1.0 / b, None, math_mode, shard, dw)
    sample = [0, 1, b // 2 - 1, b // 2, b - 1]
    l64, lse64, dq64 = _float64_rows(q, w, lab_off, lab_col, sample, "kl", 0.0, 1.0 / b)
    # lse ~ log(4.6e6) + O(1) = 15.3: an absolute bound in nats is the meaningful one for the tensor tiles
    assert (lse[sample].double().cpu() - lse64.double()).abs().max().item() <= (2e-3 if mode == "bf16" else 2e-5)
    close(rows[sample] / b, l64, st_tol if mode == "bf16" else 2e-5, f"wd5m {mode} row losses vs float64")
    close(dq[sample], dq64, g_tol, f"wd5m {mode} dQ rows vs float64")
    ents = torch.tensor([0, 1, 127, 128, 2_300_000, e - 129, e - 128, e - 2, e - 1, int(lab_col[5])], device="cuda")
    # float64 reference with the float64 lse of ALL rows is too slow; the kernel's own lse was just checked on the sample
    dt64 = _float64_table_rows(q, w, lab_off, lab_col, ents, "kl", 0.0, 1.0 / b, lse)
    close(dw[ents], dt64, g_tol, f"wd5m {mode} dTable rows vs float64")
    # rows of entities that are nobody's label and far from any query stay tiny but non-zero: the TMA reduce-add flush
    # must have touched every tile exactly once
    assert torch.isfinite(dw).all() and (dw[e - 1].abs().sum() > 0)


def test_wikidata5m_shape_scores_match_oracle_rows(kb):
    """score_sp rows at E = 4.6 M against the CPU oracle (kge_oracle.score_emb) for sampled queries, fp32 1e-5."""
    from oracle import kge_oracle as ko
    e, r, d = 4_600_000, 822, 128
    torch.manual_seed(4)
    m = kb.KgeModel("distmult", e, r, d, math_mode=kb.lib.MATH_FP32).cuda()
    s = torch.tensor([0, e - 1, 12345], device="cuda")
    p = torch.tensor([0, 821, 7], device="cuda")
    with torch.no_grad():
        got = m.score_sp(s, p)
        ent, rel = m.get_s_embedder().weight.detach().cpu(), m.get_p_embedder().weight.detach().cpu()
        want = ko.score_emb("distmult", ent[s.cpu()], rel[p.cpu()], ent, "sp_")
    close(got, want, 1e-5, "score_sp rows at E=4.6M vs oracle")


@pytest.mark.parametrize("model,math", [("complex", "tf32"), ("complex", "fp32"), ("transe", "fp32")])
def test_wikidata5m_shape_rank_counts_bit_exact_on_dyadic_tables(kb, model, math):
    """Filtered rank / tie counts at E = 4.6 M on dyadic tables (entries k/8, |k| <= 4: every product and partial sum is
    exact in fp32 and in TF32), against exact integer counting in float64 -- bit-exact, ties included."""
    e, r, d, b = 4_600_000, 822, 128, 64
    gen = torch.Generator().manual_seed(9)
    math_mode = kb.lib.MATH_TF32 if math == "tf32" else kb.lib.MATH_FP32
    m = kb.KgeModel(model, e, r, d, math_mode=math_mode).cuda()
    with torch.no_grad():
        m.get_s_embedder().weight.copy_((torch.randint(-4, 5, (e, d), generator=gen, dtype=torch.int8).float() / 8).cuda())
        m.get_p_embedder().weight.copy_((torch.randint(-4, 5, tuple(m.get_p_embedder().weight.shape), generator=gen).float() / 8).cuda())
    rng = np.random.default_rng(1)
    train = np.stack([rng.integers(0, e, 200000), rng.integers(0, 20, 200000), rng.integers(0, e, 200000)], 1).astype(np.int32)
    train[:, 0] = train[:, 0] % 5000                     # dense (s, p) keys: filters with many entries per row
    batch = torch.from_numpy(train[:b].copy())
    job = kb.EntityRankingJob(m, e, [train], None, batch_size=b, math_mode=math_mode, hits_at_k_s=(1, 10))
    got = job.rank_batch(batch)
    # exact reference for the sp_ direction of a few rows
    ent, rel = m.get_s_embedder().weight.detach(), m.get_p_embedder().weight.detach()
    sp = kb.index.KvsAllIndex(train, "sp")
    for i in (0, 1, 17, 63):
        s_i, p_i, o_i = (int(x) for x in train[i])
        with torch.no_grad():
            row = m.score_sp(torch.tensor([s_i], device="cuda"), torch.tensor([p_i], device="cuda")).double().view(-1)
            t = m.score_spo(torch.tensor([s_i], device="cuda"), torch.tensor([p_i], device="cuda"),
                            torch.tensor([o_i], device="cuda"), "o").double().view(-1)
        row[o_i] = t
        raw = int((row > t).sum()) + int((row == t).sum()) // 2
        known = sp.get((s_i, p_i)).cuda()
        row[known[known != o_i]] = float("-inf")
        filt = int((row > t).sum()) + int((row == t).sum()) // 2
        assert int(got["o_raw"][i]) == raw, (model, math, i, int(got["o_raw"][i]), raw)
        assert int(got["o_filt"][i]) == filt, (model, math, i, int(got["o_filt"][i]), filt)


@pytest.mark.parametrize("loss", ["kl", "bce"])
def test_large_table_captured_step_one_gradient_buffer(kb, loss):
    """Tables of >= 64 MB take the no-gradient-buffer flow of FusedAllEntityStepper (Adagrad applied by the flush of the dense
    table-gradient kernel, kgeb_fused_bwd_update; label and query-side rows summed per touched row and finished by a row
    kernel): three 1vsAll steps must equal the autograd flow on the same bf16 tiles to fp32 rounding."""
    e, r, d, b = 140_000, 50, 128, 256
    torch.manual_seed(0)
    ref = kb.KgeModel("distmult", e, r, d).cuda()
    new = kb.KgeModel("distmult", e, r, d).cuda()
    new.load_state_dict(ref.state_dict())
    mk = lambda m: kb.optim.create("Adagrad", m.parameters(), lr=0.2, initial_accumulator_value=0.1)   # noqa: E731
    jr = kb.TrainingJob1vsAll(ref, mk(ref), kb.KgeLoss.create(loss), math_mode=kb.lib.MATH_BF16)
    kind = kb.lib.LOSS_KL if loss == "kl" else kb.lib.LOSS_BCE
    st = kb.trainer.FusedAllEntityStepper(new, mk(new), 2 * b, 2 * b, kind, b, math_mode=kb.lib.MATH_BF16)
    assert st.seq and st.fuse_update, "the no-gradient-buffer flow (Adagrad in the tile kernel) was not selected"
    gen = torch.Generator().manual_seed(1)
    for step in range(3):
        t = torch.stack((torch.randint(0, e, (b,), generator=gen), torch.randint(0, r, (b,), generator=gen),
                         torch.randint(0, e, (b,), generator=gen)), 1).cuda()
        t[:8, 2] = t[0, 2]                                  # a hub object: several label rows on one entity
        a = jr.step(step, {"triples": t})
        z = torch.zeros(b, dtype=torch.int32, device="cuda")
        st.set_inputs(torch.cat((t[:, 0], t[:, 2])), torch.cat((t[:, 1], t[:, 1])), torch.cat((z, z + 1)),
                      torch.arange(2 * b + 1, device="cuda"), torch.cat((t[:, 2], t[:, 0])))
        got = st.step().item()
        assert got == pytest.approx(a.avg_loss, rel=2e-5), (loss, step)
        for x, y in ((new.get_s_embedder().weight, ref.get_s_embedder().weight), (new.get_p_embedder().weight, ref.get_p_embedder().weight)):
            assert (x - y).abs().max().item() <= 0.2 * 1e-3, (loss, step, (x - y).abs().max().item())


def test_touched_rows_numbering(kb):
    """kgeb_touched_build against numpy: distinct in-shard ids numbered in ascending order, ids of other shards on the dummy
    slot, padding of the second list on slot 0 and not touched."""
    rng = np.random.default_rng(5)
    e_lo, e_hi, n_a, n_b, real = 1000, 9000, 700, 900, 640
    a = rng.integers(0, 12000, n_a)
    b = rng.integers(0, 12000, n_b)
    b[real:] = 1234                                        # stale padding
    t = kb.trainer._TouchedRows(e_hi - e_lo, n_a + n_b, 16, n_a, n_b, torch.device("cuda"))
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    t.build(ta, tb, torch.tensor([real], device="cuda"), e_lo, e_hi, kb.lib.stream_ptr(ta))
    live = np.concatenate((a, b[:real]))
    uniq = np.unique(live[(live >= e_lo) & (live < e_hi)])
    assert int(t.num.item()) == len(uniq)
    assert np.array_equal(t.uniq[:len(uniq)].cpu().numpy(), uniq - e_lo)
    slot_of = np.full(e_hi - e_lo, -1, np.int32)
    slot_of[uniq - e_lo] = np.arange(len(uniq))
    assert np.array_equal(t.slot_of.cpu().numpy(), slot_of)
    want = lambda ids: np.where((ids >= e_lo) & (ids < e_hi), slot_of[np.clip(ids - e_lo, 0, e_hi - e_lo - 1)], n_a + n_b)  # noqa: E731
    assert np.array_equal(t.slot_a.cpu().numpy(), want(a))
    wb = want(b)
    wb[real:] = 0
    assert np.array_equal(t.slot_b.cpu().numpy(), wb)


@pytest.mark.parametrize("loss,d", [("kl", 128), ("bce", 128), ("kl", 64), ("bce", 256)])
def test_adagrad_in_the_table_gradient_kernel_equals_the_separate_pass(kb, loss, d, monkeypatch):
    """kgeb_fused_bwd_update + kgeb_touched_update (Adagrad applied by the flush of the dense table-gradient kernel; rows
    that also get label / query-side rows finished by a row kernel) against the same step with a stored gradient buffer
    and kgeb_adagrad_dense: rows outside the touched set get the same gradient bits and an update term within a few ulp
    (MUFU sqrt / rcp in the tile kernel), touched rows agree to rounding (dense + (labels + query side) instead of
    (dense + labels) + query side; IEEE arithmetic in the row kernel).  d = 64: two of the four column parts have no box;
    d = 256: two boxes per part and no A operand in tensor memory."""
    e, r, b = 140_000 * 128 // d, 50, 256
    torch.manual_seed(0)
    base = kb.KgeModel("distmult", e, r, d).cuda()
    mk = lambda m: kb.optim.create("Adagrad", m.parameters(), lr=0.2, initial_accumulator_value=0.1)   # noqa: E731
    kind = kb.lib.LOSS_KL if loss == "kl" else kb.lib.LOSS_BCE
    steppers = []
    for fused_update in (False, True):
        m = kb.KgeModel("distmult", e, r, d).cuda()
        m.load_state_dict(base.state_dict())
        monkeypatch.setenv("KGEB_NO_FUSED_UPDATE", "0" if fused_update else "1")
        st = kb.trainer.FusedAllEntityStepper(m, mk(m), 2 * b, 2 * b, kind, b, math_mode=kb.lib.MATH_BF16)
        assert st.seq and st.fuse_update == fused_update
        steppers.append((m, st))
    gen = torch.Generator().manual_seed(1)
    touched = torch.zeros(e, dtype=torch.bool, device="cuda")
    for step in range(3):
        t = torch.stack((torch.randint(0, e, (b,), generator=gen), torch.randint(0, r, (b,), generator=gen),
                         torch.randint(0, e, (b,), generator=gen)), 1).cuda()
        t[:8, 2] = t[0, 2]
        touched[t[:, 0]] = True
        touched[t[:, 2]] = True
        z = torch.zeros(b, dtype=torch.int32, device="cuda")
        losses = []
        for m, st in steppers:
            st.set_inputs(torch.cat((t[:, 0], t[:, 2])), torch.cat((t[:, 1], t[:, 1])), torch.cat((z, z + 1)),
                          torch.arange(2 * b + 1, device="cuda"), torch.cat((t[:, 2], t[:, 0])))
            losses.append(st.step().item())
            assert st.flash_fallbacks == 0
        assert losses[0] == pytest.approx(losses[1], rel=1e-6), (loss, step)
    (m0, s0), (m1, s1) = steppers
    torch.cuda.synchronize()
    w0, w1 = m0.get_s_embedder().weight.detach(), m1.get_s_embedder().weight.detach()
    a0, a1 = s0.opt.state[s0.ent]["sum"], s1.opt.state[s1.ent]["sum"]
    assert not torch.equal(w0, base.get_s_embedder().weight.detach()), "nothing was trained"
    # rows outside the touched set: the same gradient arithmetic; the update term uses the MUFU square root and reciprocal in
    # the tile kernel (within a few ulp of the IEEE one), and a weight that differs in its last bits can round to another
    # bf16 operand in the next step: everything agrees to rounding, nothing bit for bit after the first step
    wb = base.get_s_embedder().weight.detach()
    d0, d1 = (w0 - wb)[~touched], (w1 - wb)[~touched]
    # (2e-6: one element of G = sigmoid / batch ~ 2e-3 rounded the other way in bf16, times |Q| ~ 0.05, times lr / sqrt(state):
    # measured on 5 of 140,000 rows, ~6 columns each -- a weight that differs in its last bit moves a score across a tie)
    bad = (d0 - d1).abs() > 1e-4 * d0.abs() + 2e-6
    if bool(bad.any()):
        i = int((d0 - d1).abs().argmax())
        r_, c_ = divmod(i, d)
        rows_idx = torch.nonzero(~touched).flatten()
        raise AssertionError(f"{int(bad.sum())} elements off; worst: row {int(rows_idx[r_])} col {c_} d0 {float(d0.flatten()[i]):.6e} "
                             f"d1 {float(d1.flatten()[i]):.6e} state {float(a0[~touched].flatten()[i]):.6e} / "
                             f"{float(a1[~touched].flatten()[i]):.6e}; rows with a bad element: "
                             f"{rows_idx[bad.any(1)][:8].tolist()} of {int(bad.any(1).sum())}")
    assert ((a0 - a1).abs() <= 1e-4 * a0.abs()).all()
    assert (w0 - w1).abs().max().item() <= 1e-5 and (a0 - a1).abs().max().item() <= 1e-6 * a0.abs().max().item()
    assert torch.equal(s1.mirror, w1.bfloat16()), "bf16 mirror out of step with the table"
    # (an entity used again in a later step carries its rounding difference into that step's queries)
    assert (m0.get_p_embedder().weight - m1.get_p_embedder().weight).abs().max().item() <= 1e-5
    assert int((s1.touched.slot_of != -1).sum().item()) == 0 and float(s1.touched.g_sparse.abs().max().item()) == 0.0


@pytest.mark.parametrize("e,b,scale", [(14541, 4096, 8.0), (14541, 512, 1500.0), (4_600_000, 256, 8.0), (1000, 130, 8.0)])
def test_flash_kl_forward_and_dq_match_float64(kb, e, b, scale):
    """kgeb_fused_flash_fwd / _dq (KL, bf16 tiles: log-sum-exp and dQ from one table pass against a sampled per-row
    reference score) vs float64 on sampled rows, and vs the two-pass bf16 kernels on all rows.  scale = 1500 gives scores
    with a standard deviation of ~17 nats, the row maximum ~20 nats above the sampled reference and a softmax dominated by
    a few entities: the fixed-reference exponentials must hold (compared with the two-pass bf16 kernels there: the bf16
    rounding of the operands alone moves such a sharp softmax by several percent against float64)."""
    r, d = 50, 128
    m, q, lab_off, lab_col = _problem(kb, "distmult", e, r, d, b, 3, seed=21, empty_row=False)
    q = (q * (scale / 8.0)).contiguous()
    w = m.get_s_embedder().weight.detach()
    shard = kb.fused.Shard.full(e)
    local, o_sum, status = kb.fused.flash_forward(q, w, lab_off, lab_col, shard)
    assert int(status.item()) == 0, "the flash kernel reported a score outside its window"
    rows, lse = kb.fused.rows_loss(local, lab_off, kb.lib.LOSS_KL, 0.0, e)
    assert torch.isfinite(lse).all() and torch.isfinite(o_sum).all()
    dq = kb.fused.flash_dq(q, w, lab_off, lab_col, local, lse, 1.0 / b, None, o_sum, shard)
    sample = [0, 1, b // 2, b - 1]
    l64, lse64, dq64 = _float64_rows(q, w, lab_off, lab_col, sample, "kl", 0.0, 1.0 / b)
    if scale <= 8.0:
        assert (lse[sample].double().cpu() - lse64.double()).abs().max().item() <= 2e-3
        close(rows[sample] / b, l64, 1e-2, "flash row losses vs float64")
        close(dq[sample], dq64, 1.5e-2, "flash dQ rows vs float64")
    # the two-pass bf16 kernels (same operand rounding) on all rows
    st = kb.fused.fused_rowstats(q, w, lab_off, lab_col, kb.lib.LOSS_KL, 0.0, 0.0, kb.lib.MATH_BF16, shard)
    rows2, lse2 = kb.fused.rows_loss(st, lab_off, kb.lib.LOSS_KL, 0.0, e)
    dq2 = kb.fused.fused_backward(q, w, lab_off, lab_col, kb.lib.LOSS_KL, 0.0, 0.0, lse2, 1.0 / b, None, kb.lib.MATH_BF16,
                                  shard, None)
    assert (lse - lse2).abs().max().item() <= 2e-3 * max(1.0, scale / 40)
    close(dq, dq2, 1.5e-2, "flash dQ vs two-pass bf16 dQ (all rows)")


def test_flash_reports_scores_outside_its_window_and_autograd_falls_back(kb):
    """Entities that rows score hundreds of nats above anything the reference sampling sees: kgeb_fused_flash_fwd must
    set its status flag (its sums overflowed), and fused.all_entity_loss must then produce the two-pass result -- loss and
    gradients equal to the fp32 CUDA-core path within the bf16 bounds, everything finite."""
    e, r, d, b = 14541, 50, 128, 256
    m, q, lab_off, lab_col = _problem(kb, "distmult", e, r, d, b, 2, seed=35, empty_row=False)
    w = m.get_s_embedder().weight.detach().clone()
    hot = torch.tensor([11, 7001, 14001], device="cuda")        # no row's label, not on the sampling stride
    assert not torch.isin(hot, lab_col).any()
    w[hot] *= 1000.0           # scores of +-500 nats
    shard = kb.fused.Shard.full(e)
    _, _, status = kb.fused.flash_forward(q, w, lab_off, lab_col, shard)
    assert int(status.item()) == 1
    res = {}
    for name, mode in (("fp32", kb.lib.MATH_FP32), ("bf16", kb.lib.MATH_BF16)):
        qq, ww = q.clone().requires_grad_(True), w.clone().requires_grad_(True)
        rows = kb.fused.all_entity_loss(qq, ww, lab_off, lab_col, kb.lib.LOSS_KL, b, 0.0, 0.0, mode)
        rows.sum().backward()
        res[name] = (rows.detach(), qq.grad, ww.grad)
        assert all(torch.isfinite(t).all() for t in res[name])
    # (scores of +-500 carry +-2 nats of bf16 operand rounding: the row losses agree to that, and a softmax peaked on
    # one of three near-tied hot entities makes dQ incomparable between roundings -- finiteness was asserted above)
    close(res["bf16"][0], res["fp32"][0], 1e-2, "fallback row losses")


def test_failed_flash_pass_in_the_captured_step_is_a_no_op_and_the_two_pass_graph_repeats_it(kb):
    """The captured large-table step (flash forward, Adagrad inside the table-gradient kernel) on a table where three
    entities score hundreds of nats above the flash reference: the flash graph must change NOTHING (status word -> the tile
    kernel's update warps and the row kernel skip, the relation gradient is cleared), step() must replay the two-pass graph,
    and the result must equal the autograd flow on the same bf16 tiles."""
    e, r, d, b = 140_000, 20, 128, 128
    torch.manual_seed(0)
    ref = kb.KgeModel("distmult", e, r, d).cuda()
    hot = torch.tensor([11, 70_001, 139_001], device="cuda")
    with torch.no_grad():
        ref.get_s_embedder().weight[hot] *= 3.0e4       # scores of +-300 nats against queries of ~0.01 per dimension
    new = kb.KgeModel("distmult", e, r, d).cuda()
    new.load_state_dict(ref.state_dict())
    mk = lambda m: kb.optim.create("Adagrad", m.parameters(), lr=0.2, initial_accumulator_value=0.1)   # noqa: E731
    jr = kb.TrainingJob1vsAll(ref, mk(ref), kb.KgeLoss.create("kl"), math_mode=kb.lib.MATH_BF16)
    st = kb.trainer.FusedAllEntityStepper(new, mk(new), 2 * b, 2 * b, kb.lib.LOSS_KL, b, math_mode=kb.lib.MATH_BF16)
    assert st.seq and st.fuse_update and st.flash
    gen = torch.Generator().manual_seed(2)
    t = torch.stack((torch.randint(0, e, (b,), generator=gen), torch.randint(0, r, (b,), generator=gen),
                     torch.randint(0, e, (b,), generator=gen)), 1).cuda()
    assert not torch.isin(hot, t[:, [0, 2]].flatten()).any()
    z = torch.zeros(b, dtype=torch.int32, device="cuda")
    st.set_inputs(torch.cat((t[:, 0], t[:, 2])), torch.cat((t[:, 1], t[:, 1])), torch.cat((z, z + 1)),
                  torch.arange(2 * b + 1, device="cuda"), torch.cat((t[:, 2], t[:, 0])))
    # (1) the flash graph alone: a failed pass leaves tables, optimizer state and mirror untouched
    before = [x.detach().clone() for x in (st.ent, st.rel, st.opt.state[st.ent]["sum"], st.opt.state[st.rel]["sum"], st.mirror)]
    st.graph.replay()
    torch.cuda.synchronize()
    assert int(st.flash_status[0].item()) != 0, "the flash pass did not report the overflow"
    after = (st.ent, st.rel, st.opt.state[st.ent]["sum"], st.opt.state[st.rel]["sum"], st.mirror)
    assert all(torch.equal(x, y.detach()) for x, y in zip(before, after)), "a failed flash pass changed the model"
    assert int((st.touched.slot_of != -1).sum().item()) == 0 and float(st.touched.g_sparse.abs().max().item()) == 0.0
    # (2) the public step: falls back, and equals the autograd flow
    got = st.step().item()
    assert st.flash_fallbacks == 1
    a = jr.step(0, {"triples": t})
    assert got == pytest.approx(a.avg_loss, rel=2e-5)
    for x, y in ((new.get_s_embedder().weight, ref.get_s_embedder().weight), (new.get_p_embedder().weight, ref.get_p_embedder().weight)):
        assert torch.isfinite(x).all()
        assert bool(((x - y).abs() <= 0.2 * 1e-3 + 1e-6 * y.abs()).all()), (x - y).abs().max().item()


def test_bf16_mirror_cache_belongs_to_a_storage_not_to_an_address(kb):
    """The caching allocator hands the address of a freed table to the next tensor of the same shape, and both have version
    counter 0: the cached bf16 mirror of the first must not be served for the second."""
    a = torch.full((4096, 128), 1.0, device="cuda")
    ptr = a.data_ptr()
    assert float(kb.ops.bf16_mirror(a).float().max()) == 1.0
    del a
    b = torch.full((4096, 128), 2.0, device="cuda")
    if b.data_ptr() != ptr:
        pytest.skip("the allocator did not reuse the address")
    assert float(kb.ops.bf16_mirror(b).float().min()) == 2.0
    b.mul_(2.0)                                   # in-place change: version counter
    assert float(kb.ops.bf16_mirror(b).float().min()) == 4.0


@pytest.mark.parametrize("b", [1, 37, 1024, 4096])
def test_onevsall_batch_build_equals_the_host_side_construction(kb, b):
    """kgeb_onevsall_batch_build (first kernel of the captured 1vsAll step) against the torch construction it replaces:
    queries, label CSR, row types and the three STABLE argsorts, bit for bit (hub entities / few relations: many ties)."""
    e, r = 4_600_000, 7
    gen = torch.Generator().manual_seed(b)
    t = torch.stack((torch.randint(0, e, (b,), generator=gen), torch.randint(0, r, (b,), generator=gen),
                     torch.randint(0, e, (b,), generator=gen)), 1).cuda()
    t[: b // 3, 2] = t[0, 2]
    n = 2 * b
    i64 = lambda k: torch.full((k,), -7, dtype=torch.int64, device="cuda")     # noqa: E731
    i32 = lambda k: torch.full((k,), -7, dtype=torch.int32, device="cuda")     # noqa: E731
    a, p, lo, lc, rc, ap, pp, lp = i64(n), i64(n), i64(n + 1), i64(n), i32(n), i32(n), i32(n), i32(n)
    kb.lib.call("kgeb_onevsall_batch_build", t.data_ptr(), b, e, r, a.data_ptr(), p.data_ptr(), rc.data_ptr(), lo.data_ptr(),
                lc.data_ptr(), ap.data_ptr(), pp.data_ptr(), lp.data_ptr(), kb.lib.stream_ptr(t))
    s_, p_, o_ = t[:, 0], t[:, 1], t[:, 2]
    want_a, want_p, want_l = torch.cat((s_, o_)), torch.cat((p_, p_)), torch.cat((o_, s_))
    z = torch.zeros(b, dtype=torch.int32, device="cuda")
    sp = kb.trainer.FusedAllEntityStepper.sort_perm
    assert torch.equal(a, want_a) and torch.equal(p, want_p) and torch.equal(lc, want_l)
    assert torch.equal(rc, torch.cat((z, z + 1))) and torch.equal(lo, torch.arange(n + 1, device="cuda"))
    assert torch.equal(ap, sp(want_a)) and torch.equal(pp, sp(want_p)) and torch.equal(lp, sp(want_l, n))
