"""Parity of the CUDA path (through the C-ABI, kge-1_b200/) against golden vectors produced by the
unmodified reference and against the oracle on seeded inputs.  Needs a B200: `pytest -m gpu`.

Tolerances (BASELINE.json north_star): fp32 paths |a-b| <= 1e-5 * max|ref| ; TF32 tensor tiles
|a-b| <= 2e-3 * max|ref|, BF16 tensor tiles <= 1e-2 * max|ref| (documented in DESIGN.md) ; indices / rank counts bit-exact."""
import ast

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import kge_oracle as ko  # noqa: E402

T = torch.from_numpy
TAGS = ["distmult", "complex", "cp", "simple", "rescal", "transe_l1", "transe_l2", "rotate_l1", "rotate_l2"]
E, R = 53, 7


@pytest.fixture(scope="module")
def kb():
    import kgeb200
    assert torch.cuda.is_available(), "these tests need a GPU"
    kgeb200.lib.load()
    return kgeb200


def close(got, ref, rtol=1e-5, what=""):
    got = got.detach().cpu().double().numpy() if isinstance(got, torch.Tensor) else np.asarray(got, dtype=np.float64)
    ref = ref.detach().cpu().double().numpy() if isinstance(ref, torch.Tensor) else np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    scale = max(np.abs(ref).max(), 1e-30) if ref.size else 1.0
    err = np.abs(got - ref).max() if ref.size else 0.0
    assert err <= rtol * scale + 1e-30, f"{what}: max abs err {err:.3e} > {rtol:g} * {scale:.3e}"


def split_tag(tag):
    m = tag.split("_l")[0]
    return m, (float(tag.split("_l")[1]) if "_l" in tag else 1.0)


def make_model(kb, model, ent, rel, l_norm=1.0, **kw):
    m = kb.KgeModel(model, ent.shape[0], rel.shape[0], ent.shape[1], l_norm=l_norm, relation_dim=rel.shape[1], **kw).cuda()
    with torch.no_grad():
        m.get_s_embedder().weight.copy_(torch.as_tensor(ent))
        m.get_p_embedder().weight.copy_(torch.as_tensor(rel))
    return m


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", TAGS)
def test_scores_match_reference_golden(kb, golden, tag):
    g = golden("scores")
    model, ln = split_tag(tag)
    m = make_model(kb, model, g[f"{tag}.ent"], g[f"{tag}.rel"], ln)
    s, p, o, sub = (T(g[k]).cuda() for k in ("idx_s", "idx_p", "idx_o", "subset"))
    with torch.no_grad():
        close(m.score_spo(s, p, o), g[f"{tag}.spo"], what="spo")
        close(m.score_spo(s.int(), p.int(), o.int()), g[f"{tag}.spo_i32"], what="spo int32")
        close(m.score_sp(s, p), g[f"{tag}.sp"], what="sp_")
        close(m.score_po(p, o), g[f"{tag}.po"], what="_po")
        close(m.score_sp(s, p, sub), g[f"{tag}.sp_sub"], what="sp_ subset")
        close(m.score_sp_po(s, p, o, sub), g[f"{tag}.sp_po_sub"], what="sp_po subset")
        close(m.score_so(s, o), g[f"{tag}.so"], what="s_o")
        # embedder API
        close(m.get_s_embedder().embed(s.int()), g[f"{tag}.ent"][g["idx_s"]], rtol=0, what="embed")
        assert m.get_s_embedder().embed_all().shape == (E, g[f"{tag}.ent"].shape[1])


@pytest.mark.parametrize("tag", TAGS)
def test_score_gradients_match_oracle(kb, golden, tag):
    g = golden("scores")
    model, ln = split_tag(tag)
    ent, rel = T(g[f"{tag}.ent"]), T(g[f"{tag}.rel"])
    s, p, o = (T(g[k]) for k in ("idx_s", "idx_p", "idx_o"))
    gen = torch.Generator().manual_seed(3)
    for form in ("spo", "sp", "po"):
        prm = ko.Params(ent, rel)
        if form == "spo":
            ref = ko.score_spo(model, prm.ent, prm.rel, s, p, o, ln)
        elif form == "sp":
            ref = ko.score_sp(model, prm.ent, prm.rel, s, p, l_norm=ln)
        else:
            ref = ko.score_po(model, prm.ent, prm.rel, p, o, l_norm=ln)
        up = torch.randn(ref.shape, generator=gen)
        (ref * up).sum().backward()
        m = make_model(kb, model, ent, rel, ln)
        sc, pc, oc = s.cuda(), p.cuda(), o.cuda()
        got = m.score_spo(sc, pc, oc) if form == "spo" else (m.score_sp(sc, pc) if form == "sp" else m.score_po(pc, oc))
        (got * up.cuda()).sum().backward()
        ge, gr = prm.grads()
        close(m.get_s_embedder().weight.grad, ge, rtol=2e-5, what=f"{tag} {form} d entity")
        close(m.get_p_embedder().weight.grad, gr, rtol=2e-5, what=f"{tag} {form} d relation")


def test_error_behaviour(kb):
    with pytest.raises(ValueError):
        kb.KgeModel("complex", 10, 3, 7)          # odd dim (complex.py / cp.py:45-49 style)
    m = kb.KgeModel("distmult", 10, 3, 8).cuda()
    x = torch.zeros(2, 8, device="cuda")
    with pytest.raises(ValueError):
        m.get_scorer().score_emb(x, x, x, "xyz")   # kge_model.py:180
    with pytest.raises(ValueError):
        kb.ops.score_all(kb.lib.DOT, torch.zeros(2, 8), x)  # CPU tensor: no CPU path
    with pytest.raises(NotImplementedError):
        kb.KgeModel("transe", 10, 3, 8, l_norm=3.0)


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(300, 1000, 128), (128, 128, 32), (77, 333, 64), (513, 2049, 100)])
def test_tf32_tensor_tiles_scores(kb, shape):
    b, e, d = shape
    gen = torch.Generator().manual_seed(b + e)
    q = torch.randn(b, d, generator=gen).cuda() * 0.3
    w = torch.randn(e, d, generator=gen).cuda() * 0.3
    ref = (q.double() @ w.double().t()).float()
    fp32 = kb.ops.score_all(kb.lib.DOT, q, w, kb.lib.MATH_FP32)
    close(fp32, ref, rtol=1e-5, what="fp32 tiles")
    tf32 = kb.ops.score_all(kb.lib.DOT, q, w, kb.lib.MATH_TF32)
    close(tf32, ref, rtol=2e-3, what="tf32 tcgen05 tiles")
    # dyadic inputs are exact in TF32: the tensor path must then be bit-identical to exact arithmetic
    qd = torch.round(q * 16) / 16
    wd = torch.round(w * 16) / 16
    exact = (qd.double() @ wd.double().t()).float()
    assert torch.equal(kb.ops.score_all(kb.lib.DOT, qd, wd, kb.lib.MATH_TF32), exact)


# ---------------------------------------------------------------------------------------------
def _run_case(kb, golden, tag, fused_path, math_mode=None):
    g = golden("train")
    _, ttype, model, loss = tag.split(".")
    opts = dict(ast.literal_eval(str(g[tag + ".options"]))) if tag + ".options" in g else {}
    ln = float(opts.get(model + ".l_norm", 1.0))
    off = float(opts.get("train.loss_arg", 0.0))
    ls = float(opts.get("KvsAll.label_smoothing", 0.0))
    math_mode = kb.lib.MATH_FP32 if math_mode is None else math_mode
    m = make_model(kb, model, g[tag + ".ent0"], g[tag + ".rel0"], ln)
    opt = kb.optim.create("Adagrad", m.parameters(), lr=0.2)
    lossf = kb.KgeLoss.create(loss, off)
    kw = dict(fused_path=fused_path, math_mode=math_mode)
    if ttype == "1vsAll":
        job = kb.TrainingJob1vsAll(m, opt, lossf, **kw)
    elif ttype == "KvsAll":
        job = kb.TrainingJobKvsAll(m, opt, lossf, E, R, label_smoothing=ls, **kw)
    else:
        job = kb.TrainingJobNegativeSampling(m, opt, lossf, **kw)
    rtol = {kb.lib.MATH_FP32: 1e-5, kb.lib.MATH_TF32: 2e-3, kb.lib.MATH_BF16: 1e-2}[math_mode]
    for step in range(2):
        pre = f"{tag}.b{step}"
        if ttype == "KvsAll":
            batch = {"queries": T(g[pre + ".queries"]), "label_coords": T(g[pre + ".label_coords"]),
                     "query_type_indexes": T(g[pre + ".query_type"])}
        else:
            batch = {"triples": T(g[pre + ".triples"])}
            if ttype == "negative_sampling":
                batch["negative_samples"] = [T(g[f"{pre}.neg{slot}"]) for slot in range(3)]
        res = job.step(step, batch)
        assert res.avg_loss == pytest.approx(float(g[pre + ".loss"]), rel=max(rtol, 2e-5)), f"{tag} loss step {step}"
        # the golden graphs use LibKGE's default unit-variance initialisation, so scores are O(10): the
        # softmax / sigmoid turn the tensor tiles' score error (rtol * max|score|) into a gradient error that is
        # |score| times larger than on realistically scaled tables (see test_fused_backward_matches_float64)
        gtol = 2 * rtol if math_mode == kb.lib.MATH_FP32 else 12 * rtol
        close(m.get_s_embedder().weight.grad, g[pre + ".grad_ent"], rtol=gtol, what=f"{tag} grad entity {step}")
        close(m.get_p_embedder().weight.grad, g[pre + ".grad_rel"], rtol=gtol, what=f"{tag} grad relation {step}")
        if math_mode != kb.lib.MATH_FP32:
            # Adagrad's first steps move every touched weight by ~lr * sign(g): entries with g ~ 0 flip sign under
            # any rounding change, so post-step parameters are only compared on the fp32 path
            with torch.no_grad():
                m.get_s_embedder().weight.copy_(T(g[pre + ".ent"]))
                m.get_p_embedder().weight.copy_(T(g[pre + ".rel"]))
            torch.autograd.graph.increment_version(m.get_s_embedder().weight)
            continue
        # Adagrad divides by sqrt(sum g^2): where g ~ 0 the update direction is ill-conditioned, so the
        # post-step parameters are compared with an absolute bound in units of the learning rate
        for got, ref in ((m.get_s_embedder().weight, g[pre + ".ent"]), (m.get_p_embedder().weight, g[pre + ".rel"])):
            err = (got.detach().cpu() - T(ref)).abs().max().item()
            assert err <= 0.2 * (2e-3 if math_mode == kb.lib.MATH_FP32 else 0.25), f"{tag} params step {step}: {err}"


def _cases(golden):
    return [str(x) for x in golden("train")["train.cases"]]


def test_training_steps_reference_flow(kb, golden):
    """The reference's own batch bodies (score_* -> loss -> backward -> Adagrad) on the CUDA path."""
    for tag in _cases(golden):
        _run_case(kb, golden, tag, fused_path=False)


def test_training_steps_fused_flow(kb, golden):
    """Fused score+loss kernels (scores never materialised), fp32 CUDA-core tiles."""
    for tag in _cases(golden):
        _run_case(kb, golden, tag, fused_path=True)


# ---------------------------------------------------------------------------------------------
def test_fused_forward_stats_tf32_vs_fp32(kb):
    gen = torch.Generator().manual_seed(5)
    b, e, d = 200, 1500, 128
    q = (torch.randn(b, d, generator=gen) * 0.2).cuda()
    w = (torch.randn(e, d, generator=gen) * 0.2).cuda()
    cols = torch.randint(0, e, (b, 3), generator=gen).sort(dim=1).values.cuda()
    lab_off = torch.arange(0, 3 * b + 1, 3, dtype=torch.int64).cuda()
    lab_col = cols.reshape(-1).contiguous()
    x = (q.double() @ w.double().t())
    for loss, name in ((kb.lib.LOSS_KL, "kl"), (kb.lib.LOSS_BCE, "bce")):
        shard = kb.fused.Shard.full(e)
        ref_rows = None
        for math_mode, rtol in ((kb.lib.MATH_FP32, 2e-5), (kb.lib.MATH_TF32, 2e-3), (kb.lib.MATH_BF16, 1e-2)):
            st = kb.fused.fused_rowstats(q, w, lab_off, lab_col, loss, 0.0, 0.25 if loss else 0.0, math_mode, shard)
            rows, lse = kb.fused.rows_loss(st, lab_off, loss, 0.0, e)
            if loss == kb.lib.LOSS_KL:
                want = torch.logsumexp(x, 1) - x.gather(1, cols).sum(1) / 3 - np.log(3.0)
            else:
                xo = x + 0.25
                want = torch.nn.functional.softplus(xo).sum(1) - xo.gather(1, cols).sum(1)
            close(rows, want.float(), rtol=rtol, what=f"{name} rows math={math_mode}")


def test_entity_ranking_matches_reference_golden(kb, golden):
    g = golden("eval")
    graph = {k: g[f"eval.graph.{k}"] for k in ("train", "valid", "test")}
    for tag in [str(x) for x in g["eval.cases"]]:
        _, model, chunk, ties = tag.split(".")
        for math_mode in (kb.lib.MATH_FP32, kb.lib.MATH_TF32):
            m = make_model(kb, model, g[tag + ".ent"], g[tag + ".rel"], 1.0, math_mode=math_mode)
            job = kb.EntityRankingJob(m, E, [graph["train"], graph["valid"]], graph["test"], batch_size=16,
                                      tie_handling=ties, hits_at_k_s=(1, 3, 10, 50), math_mode=math_mode)
            res = job.run(graph["valid"])
            exact = model != "rotate"   # dyadic weights: every sum is exact, so ties must match bit-for-bit
            for nm in ("raw", "filt", "filt_test"):
                for d in "so":
                    got = res["ranks"][f"{d}_{nm}"].cpu().numpy()
                    want = g[f"{tag}.{d}_ranks_{nm}"]
                    if exact:
                        np.testing.assert_array_equal(got, want, err_msg=f"{tag} {d}_{nm} math={math_mode}")
                    else:
                        assert np.mean(got != want) < 0.1
            mm = res["metrics"]
            for key in ("mean_reciprocal_rank", "mean_reciprocal_rank_filtered",
                        "mean_reciprocal_rank_filtered_with_test", "hits_at_10_filtered", "hits_at_1_filtered"):
                assert mm[key] == pytest.approx(float(g[f"{tag}.metric.{key}"]), abs=1e-6 if exact else 1e-3), (tag, key)


def test_scatter_and_segment_reduce_are_deterministic(kb):
    gen = torch.Generator().manual_seed(0)
    n, d, v = 5000, 48, 97
    idx = torch.randint(0, v, (n,), generator=gen).cuda()
    rows = torch.randn(n, d, generator=gen).cuda()
    a = torch.zeros(v, d, device="cuda")
    b = torch.zeros(v, d, device="cuda")
    kb.ops.scatter_add_rows_(a, idx, rows)
    kb.ops.scatter_add_rows_(b, idx.int(), rows)
    assert torch.equal(a, b)
    ref = torch.zeros(v, d, dtype=torch.float64).index_add_(0, idx.cpu(), rows.cpu().double())
    close(a, ref.float(), rtol=1e-6, what="scatter")
    ids, red, cnt = kb.ops.segment_reduce_rows(idx, rows)
    k = int(cnt.item())
    assert torch.equal(ids[:k].cpu(), torch.unique(idx.cpu()))
    assert torch.equal(red[:k], a[ids[:k]])


@pytest.mark.parametrize("n,v", [(5000, 97), (20000, 14541), (300, 3)])
def test_scatter_with_collate_permutation_is_bit_identical(kb, n, v):
    """kgeb_scatter_add_rows_perm (sort hoisted into the collate) == kgeb_scatter_add_rows (device sort)."""
    gen = torch.Generator().manual_seed(n)
    idx = torch.randint(0, v, (n,), generator=gen)
    rows = torch.randn(n, 40, generator=gen).cuda()
    perm = torch.sort(idx, stable=True).indices.to(torch.int32).cuda()
    a = torch.zeros(v, 40, device="cuda")
    b = torch.zeros(v, 40, device="cuda")
    kb.ops.scatter_add_rows_(a, idx.cuda(), rows)
    kb.ops.scatter_add_rows_(b, idx.cuda(), rows, perm=perm)
    assert torch.equal(a, b)


def test_fused_backward_label_permutation_is_bit_identical(kb):
    b, e, d = 300, 1000, 64
    gen = torch.Generator().manual_seed(5)
    q = (torch.randn(b, d, generator=gen) * 0.3).cuda()
    w = (torch.randn(e, d, generator=gen) * 0.3).cuda()
    cols = torch.stack([torch.randperm(e, generator=gen)[:3].sort().values for _ in range(b)])
    lab_off = torch.arange(0, 3 * b + 1, 3, dtype=torch.int64).cuda()
    lab_col = cols.reshape(-1).contiguous().cuda()
    perm = torch.sort(lab_col, stable=True).indices.to(torch.int32)
    shard = kb.fused.Shard.full(e)
    out = []
    for lp in (None, perm):
        dw = torch.zeros_like(w)
        dq = kb.fused.fused_backward(q, w, lab_off, lab_col, kb.lib.LOSS_BCE, 0.0, 0.0, None, 1.0 / b, None,
                                     kb.lib.MATH_BF16, shard, dw, lab_perm=lp)
        out.append((dq, dw))
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])


def test_graph_stepper_bf16_tiles_match_autograd_bf16_flow(kb):
    """The graph-captured step on the BF16 tensor tiles (statistics from the dQ kernel, label permutation from the
    collate, forked label part) against the autograd flow on the same tiles."""
    g = kb.graph.synthetic_graph("toy", seed=3)
    e, r, d, b = g["num_entities"], g["num_relations"], 32, 64
    idx = [ko.kvsall_index(g["train"], "sp"), ko.kvsall_index(g["train"], "po")]
    rng = np.random.default_rng(4)
    batches = []
    for _ in range(3):
        ids = rng.choice(len(idx[0][0]) + len(idx[1][0]), b, replace=False)
        q, c, qt = ko.kvsall_collate(ids.tolist(), idx)
        batches.append({"queries": T(q), "label_coords": T(c), "query_type_indexes": T(qt)})
    nnz_max = max(len(x["label_coords"]) for x in batches) + 7
    torch.manual_seed(0)
    ref = kb.KgeModel("complex", e, r, d).cuda()
    with torch.no_grad():
        for p_ in ref.parameters():
            p_.mul_(0.2)
    new = kb.KgeModel("complex", e, r, d).cuda()
    new.load_state_dict(ref.state_dict())
    mk = lambda m: kb.TrainingJobKvsAll(m, kb.optim.create("Adagrad", m.parameters(), lr=0.05), kb.KgeLoss.create("bce"),
                                        e, r, math_mode=kb.lib.MATH_BF16)
    jr, jn = mk(ref), mk(new)
    jn.enable_graph_step(b, nnz_max)
    for i, batch in enumerate(batches):
        a = jr.step(i, batch)
        c = jn.step(i, jn.collate_packed(batch) if i % 2 else batch)
        assert c.total_loss == pytest.approx(a.total_loss, rel=1e-3)
        assert c.avg_loss == pytest.approx(a.avg_loss, rel=1e-3)   # value of the last query type (train.py:747)
        close(new.get_s_embedder().weight, ref.get_s_embedder().weight, rtol=2e-3, what=f"entity table step {i}")
        close(new.get_p_embedder().weight, ref.get_p_embedder().weight, rtol=2e-3, what=f"relation table step {i}")


def test_device_filter_csr_matches_reference_coordinates(kb, golden):
    """kgeb_filter_csr_* against get_sp_po_coords_from_spo_batch (job/util.py:5-38) restated in the oracle, for one
    split and for the union of two splits: per row the same multiset of (column) coordinates, ascending."""
    g = kb.graph.synthetic_graph("toy", seed=5)
    e = g["num_entities"]
    splits = [g["train"], g["valid"]]
    batch = np.concatenate([g["valid"][:40], g["train"][:25], np.array([[0, 0, 0], [e - 1, 0, e - 1]])]).astype(np.int64)
    s, p, o = (torch.from_numpy(batch[:, i].copy()).cuda() for i in range(3))
    for use in ([0], [0, 1]):
        sp = [kb.index.KvsAllIndex(splits[i], "sp") for i in use]
        po = [kb.index.KvsAllIndex(splits[i], "po") for i in use]
        off, col = kb.index.filter_csr(sp, po, s, p, o)
        off, col = off.cpu().numpy(), col.cpu().numpy()
        b = len(batch)
        want = [[] for _ in range(2 * b)]
        for i in use:
            coords = ko.sp_po_coords(batch, e, ko.CsrIndex(splits[i], "sp"), ko.CsrIndex(splits[i], "po"))
            for r, c in coords.tolist():
                want[r if c < e else b + r].append(c if c < e else c - e)
        assert off[0] == 0 and off[-1] == len(col) == sum(len(w) for w in want)
        for r in range(2 * b):
            assert col[off[r]:off[r + 1]].tolist() == sorted(want[r]), f"row {r} splits {use}"
    # int32 triples (the evaluation dtype of the reference, dataset.py:178) give the same CSR
    off32, col32 = kb.index.filter_csr(sp, po, s.int(), p.int(), o.int())
    assert torch.equal(off32.cpu(), torch.from_numpy(off)) and torch.equal(col32.cpu(), torch.from_numpy(col))


def test_device_kvsall_batch_matches_reference_collate(kb):
    """kgeb_kvsall_batch_* against TrainingJobKvsAll's collate (train.py:590-677) restated in the oracle."""
    g = kb.graph.synthetic_graph("toy", seed=3)
    idx = [ko.kvsall_index(g["train"], "sp"), ko.kvsall_index(g["train"], "po")]
    sp, po = kb.index.KvsAllIndex(g["train"], "sp"), kb.index.KvsAllIndex(g["train"], "po")
    ids = np.random.default_rng(0).choice(len(idx[0][0]) + len(idx[1][0]), 97, replace=False)
    q, coords, qt = ko.kvsall_collate(ids.tolist(), idx)
    a, p_, rc, lab_off, lab_col, ovf = kb.index.kvsall_batch(sp, po, torch.from_numpy(ids).cuda())
    wa, wp, wrc = kb.trainer.kvsall_rows(T(q), T(qt))
    assert torch.equal(a.cpu(), wa) and torch.equal(p_.cpu(), wp) and torch.equal(rc.cpu(), wrc) and ovf.item() == 0
    w_off, w_col = kb.fused.csr_from_coords(T(coords), len(ids))
    assert torch.equal(lab_off.cpu(), w_off) and torch.equal(lab_col.cpu(), w_col)
    # static buffers: no allocation, overflow flag when the label buffer is too small
    out = (torch.empty_like(a), torch.empty_like(p_), torch.empty_like(rc), torch.empty_like(lab_off),
           torch.zeros(len(w_col) - 1, dtype=torch.int64, device="cuda"))
    *_, ovf = kb.index.kvsall_batch(sp, po, torch.from_numpy(ids).cuda(), out=out, capacity=len(w_col) - 1)
    assert ovf.item() == 1 and torch.equal(out[3].cpu(), w_off)


def test_device_collated_steps_equal_host_collated_steps(kb):
    """TrainingJobKvsAll.step_ids (batches built on the device from example ids, double-buffered) is bit-identical
    to step() on the host-collated batches of the same ids."""
    g = kb.graph.synthetic_graph("toy", seed=3)
    e, r, d, b = g["num_entities"], g["num_relations"], 32, 64
    oidx = [ko.kvsall_index(g["train"], "sp"), ko.kvsall_index(g["train"], "po")]
    sp, po = kb.index.KvsAllIndex(g["train"], "sp"), kb.index.KvsAllIndex(g["train"], "po")
    rng = np.random.default_rng(9)
    ids = [rng.choice(len(oidx[0][0]) + len(oidx[1][0]), b, replace=False) for _ in range(4)]
    host = []
    for x in ids:
        q, c, qt = ko.kvsall_collate(x.tolist(), oidx)
        host.append({"queries": T(q), "label_coords": T(c), "query_type_indexes": T(qt)})
    nnz_max = max(len(x["label_coords"]) for x in host) + 3
    res = []
    for device_collate in (False, True):
        torch.manual_seed(0)
        m = kb.KgeModel("complex", e, r, d).cuda()
        job = kb.TrainingJobKvsAll(m, kb.optim.create("Adagrad", m.parameters(), lr=0.2), kb.KgeLoss.create("bce"), e, r,
                                   math_mode=kb.lib.MATH_BF16)
        job.enable_graph_step(b, nnz_max)
        losses = []
        if device_collate:
            job.enable_device_collate(sp, po)
            pinned = [torch.from_numpy(x.astype(np.int64)).pin_memory() for x in ids]
            job.prefetch_ids(pinned[0])
            for i in range(len(ids)):
                if i + 1 < len(ids):
                    job.prefetch_ids(pinned[i + 1])
                losses.append(job.step_ids().avg_loss)
        else:
            for i, batch in enumerate(host):
                losses.append(job.step(i, job.collate_packed(batch)).avg_loss)
        res.append((losses, m.get_s_embedder().weight.detach().clone(), m.get_p_embedder().weight.detach().clone()))
    assert res[0][0] == res[1][0]
    assert torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])


def test_kvsall_index_device_lookup_bit_exact(kb, golden):
    g = golden("index")
    tr = g["index.train.triples"]
    for key in ("sp", "po"):
        ix = kb.index.KvsAllIndex(tr, key)
        np.testing.assert_array_equal(ix._keys.numpy(), g[f"index.train.{key}.keys"])
        np.testing.assert_array_equal(ix._values_offset.numpy(), g[f"index.train.{key}.offsets"])
        np.testing.assert_array_equal(ix._values.numpy(), g[f"index.train.{key}.values"])
    # device-side filter coordinates == get_sp_po_coords_from_spo_batch (as a set per row)
    batch = T(g["index.coords_batch"]).long().cuda()
    sp, po = kb.index.KvsAllIndex(tr, "sp"), kb.index.KvsAllIndex(tr, "po")
    _, off, val = sp.device_arrays("cuda")
    o_off, o_col = kb.index.gather_csr_rows(off, val, sp.lookup(batch[:, [0, 1]].contiguous()))
    _, off, val = po.device_arrays("cuda")
    s_off, s_col = kb.index.gather_csr_rows(off, val, po.lookup(batch[:, [1, 2]].contiguous()), add=E)
    got = set()
    for r in range(len(batch)):
        got |= {(r, int(c)) for c in o_col[o_off[r]:o_off[r + 1]].tolist()}
        got |= {(r, int(c)) for c in s_col[s_off[r]:s_off[r + 1]].tolist()}
    assert got == {(int(a), int(b)) for a, b in g["index.coords"]}


@pytest.mark.parametrize("mode", ["tf32", "bf16"])
def test_training_steps_fused_tensor_tiles(kb, golden, mode):
    """Fused flow on the tcgen05 tiles: TF32 = tensor-tile forward statistics + fp32 backward tiles;
    BF16 = tensor tiles for the forward statistics and both backward GEMMs."""
    math_mode = kb.lib.MATH_TF32 if mode == "tf32" else kb.lib.MATH_BF16
    for tag in _cases(golden):
        if tag.split(".")[2] in ("distmult", "complex", "cp", "simple", "rescal") and "negative_sampling" not in tag:
            _run_case(kb, golden, tag, fused_path=True, math_mode=math_mode)


@pytest.mark.parametrize("shape", [(300, 1111, 128), (64, 4000, 64), (1000, 500, 32), (129, 257, 16)])
def test_fused_backward_matches_float64(kb, shape):
    b, e, d = shape
    gen = torch.Generator().manual_seed(e)
    q = (torch.randn(b, d, generator=gen) * 0.3).cuda()
    w = (torch.randn(e, d, generator=gen) * 0.3).cuda()
    nlab = 2
    cols = torch.stack([torch.randperm(e, generator=gen)[:nlab].sort().values for _ in range(b)]).cuda()
    lab_off = torch.arange(0, nlab * b + 1, nlab, dtype=torch.int64).cuda()
    lab_col = cols.reshape(-1).contiguous()
    rscale = (torch.rand(b, generator=gen) + 0.5).cuda()
    shard = kb.fused.Shard.full(e)
    for loss in (kb.lib.LOSS_KL, kb.lib.LOSS_BCE):
        qd, wd = q.double().requires_grad_(True), w.double().requires_grad_(True)
        x = qd @ wd.t()
        if loss == kb.lib.LOSS_KL:
            rows = torch.logsumexp(x, 1) - x.gather(1, cols).sum(1) / nlab
        else:
            rows = torch.nn.functional.softplus(x + 0.1).sum(1) - (x + 0.1).gather(1, cols).sum(1)
        ((rows * rscale.double()).sum() / b).backward()
        for math_mode, rtol in ((kb.lib.MATH_FP32, 2e-5), (kb.lib.MATH_TF32, 3e-3), (kb.lib.MATH_BF16, 1.5e-2)):
            st = kb.fused.fused_rowstats(q, w, lab_off, lab_col, loss, 0.0, 0.1 if loss else 0.0, math_mode, shard)
            _, lse = kb.fused.rows_loss(st, lab_off, loss, 0.0, e)
            dw = torch.zeros_like(w)
            dq = kb.fused.fused_backward(q, w, lab_off, lab_col, loss, 0.0, 0.1 if loss else 0.0, lse, 1.0 / b, rscale,
                                         math_mode, shard, dw)
            close(dq, qd.grad.float(), rtol=rtol, what=f"dQ loss={loss} math={math_mode} {shape}")
            close(dw, wd.grad.float(), rtol=rtol, what=f"dTable loss={loss} math={math_mode} {shape}")


@pytest.mark.parametrize("use_graph", [False, True])
def test_graph_captured_stepper_matches_autograd_flow(kb, use_graph):
    """FusedAllEntityStepper (no autograd, CUDA graph) == TrainingJobKvsAll's fused autograd flow, fp32 tiles."""
    g = kb.graph.synthetic_graph("toy", seed=3)
    e, r, d, b = g["num_entities"], g["num_relations"], 32, 64
    idx = [ko.kvsall_index(g["train"], "sp"), ko.kvsall_index(g["train"], "po")]
    rng = np.random.default_rng(1)
    batches = []
    for _ in range(3):
        ids = rng.choice(len(idx[0][0]) + len(idx[1][0]), b, replace=False)
        q, c, qt = ko.kvsall_collate(ids.tolist(), idx)
        batches.append({"queries": T(q), "label_coords": T(c), "query_type_indexes": T(qt)})
    nnz_max = max(len(x["label_coords"]) for x in batches)
    torch.manual_seed(0)
    ref = kb.KgeModel("complex", e, r, d).cuda()
    new = kb.KgeModel("complex", e, r, d).cuda()
    new.load_state_dict(ref.state_dict())
    jr = kb.TrainingJobKvsAll(ref, kb.optim.create("Adagrad", ref.parameters(), lr=0.2), kb.KgeLoss.create("bce"), e, r)
    jn = kb.TrainingJobKvsAll(new, kb.optim.create("Adagrad", new.parameters(), lr=0.2), kb.KgeLoss.create("bce"), e, r)
    jn.enable_graph_step(b, nnz_max, use_graph=use_graph)
    for i, batch in enumerate(batches):
        a, c = jr.step(i, batch), jn.step(i, batch)
        assert c.total_loss == pytest.approx(a.total_loss, rel=1e-5)
        assert c.avg_loss == pytest.approx(a.avg_loss, rel=1e-5)   # value of the last query type (train.py:747)
        close(new.get_s_embedder().weight, ref.get_s_embedder().weight, rtol=1e-5, what=f"entity table step {i}")
        close(new.get_p_embedder().weight, ref.get_p_embedder().weight, rtol=1e-5, what=f"relation table step {i}")


def test_packed_host_batches_equal_dict_batches(kb):
    g = kb.graph.synthetic_graph("toy", seed=3)
    e, r, d, b = g["num_entities"], g["num_relations"], 32, 64
    idx = [ko.kvsall_index(g["train"], "sp"), ko.kvsall_index(g["train"], "po")]
    ids = np.random.default_rng(2).choice(len(idx[0][0]) + len(idx[1][0]), b, replace=False)
    q, c, qt = ko.kvsall_collate(ids.tolist(), idx)
    batch = {"queries": T(q), "label_coords": T(c), "query_type_indexes": T(qt)}
    res = []
    for packed in (False, True):
        torch.manual_seed(0)
        m = kb.KgeModel("distmult", e, r, d).cuda()
        job = kb.TrainingJobKvsAll(m, kb.optim.create("Adagrad", m.parameters(), lr=0.2), kb.KgeLoss.create("kl"), e, r)
        job.enable_graph_step(b, len(c) + 5)
        out = job.step(0, job.collate_packed(batch) if packed else batch)
        res.append((out.avg_loss, m.get_s_embedder().weight.detach().clone()))
    assert res[0][0] == res[1][0] and torch.equal(res[0][1], res[1][1])


@pytest.mark.parametrize("case", ["train.negative_sampling.rotate.kl", "train.negative_sampling.transe.bce",
                                  "train.negative_sampling.complex.bce", "train.negative_sampling.distmult.kl"])
def test_graph_captured_negative_sampling_step_matches_reference_golden(kb, golden, case):
    g = golden("train")
    _, _, model, loss = case.split(".")
    opts = dict(ast.literal_eval(str(g[case + ".options"]))) if case + ".options" in g else {}
    ln = float(opts.get(model + ".l_norm", 1.0))
    m = make_model(kb, model, g[case + ".ent0"], g[case + ".rel0"], ln)
    opt = kb.optim.create("Adagrad", m.parameters(), lr=0.2)
    job = kb.TrainingJobNegativeSampling(m, opt, kb.KgeLoss.create(loss, float(opts.get("train.loss_arg", 0.0))))
    negs0 = [T(g[f"{case}.b0.neg{slot}"]) for slot in range(3)]
    job.enable_graph_step(len(g[case + ".b0.triples"]), negs0[0].shape[1], negs0[2].shape[1])
    for step in range(2):
        pre = f"{case}.b{step}"
        batch = {"triples": T(g[pre + ".triples"]), "negative_samples": [T(g[f"{pre}.neg{slot}"]) for slot in range(3)]}
        res = job.step(step, batch)
        assert res.avg_loss == pytest.approx(float(g[pre + ".loss"]), rel=2e-5), f"{case} loss step {step}"
        for got, ref in ((m.get_s_embedder().weight, g[pre + ".ent"]), (m.get_p_embedder().weight, g[pre + ".rel"])):
            err = (got.detach().cpu() - T(ref)).abs().max().item()
            assert err <= 0.2 * 2e-3, f"{case} params step {step}: {err}"
