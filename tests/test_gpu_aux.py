"""GPU parity of the rows around the hot path, through the C-ABI:
on-device negative sampling (bit-exact against the numpy restatement of the device generator, plus the reference's
sampling contract), Lp penalties fused with their gradient / the Adagrad update (against golden vectors of the
unmodified reference), rank histograms + metrics incl. the head/tail and relation-type drill-downs (golden)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import kge_oracle as ko  # noqa: E402
from oracle import sampler_oracle as so  # noqa: E402

T = torch.from_numpy


@pytest.fixture(scope="module")
def kb():
    import kgeb200
    assert torch.cuda.is_available(), "these tests need a GPU"
    kgeb200.lib.load()
    return kgeb200


def _state(sampler):
    seed, off = sampler.state.tolist()
    return int(seed), int(off)


# ---------------------------------------------------------------------------------------------
# 8f-2 sampling
# ---------------------------------------------------------------------------------------------
def test_device_philox_matches_known_answers_and_oracle(kb):
    n = 1000
    out = torch.empty(4 * n, dtype=torch.int32, device="cuda")
    kb.lib.call("kgeb_philox_words", 0, 0, 0, n, out.data_ptr(), kb.lib.stream_ptr(out))
    got = out.cpu().numpy().view(np.uint32).reshape(n, 4)
    assert tuple(int(x) for x in got[0]) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)   # Random123 KAT
    seed, off = 0x123456789abcdef, 0x1_0000_0005
    kb.lib.call("kgeb_philox_words", seed, off, 77, n, out.data_ptr(), kb.lib.stream_ptr(out))
    got = out.cpu().numpy().view(np.uint32).reshape(n, 4)
    elem = np.arange(77, 77 + n, dtype=np.uint64)
    ctr = np.stack([np.full(n, off & 0xFFFFFFFF, np.uint64), np.full(n, off >> 32, np.uint64), elem & so.MASK,
                    elem >> np.uint64(32)], 1)
    want = so.philox4x32_10(ctr, np.array([seed & 0xFFFFFFFF, seed >> 32], dtype=np.uint64))
    assert np.array_equal(got, want)


def test_uniform_sampling_bit_exact_and_uniform(kb):
    e, r = 40943, 11
    smp = kb.KgeUniformSampler(e, r, num_samples=(256, 3, -1), seed=1234)
    assert smp.num_samples == [256, 3, 256]                       # sampler.py:59-65
    triples = torch.randint(0, 11, (512, 3), device="cuda")
    for slot in (0, 1, 2):
        seed, off = _state(smp)
        neg = smp.sample(triples, slot)
        vocab = r if slot == 1 else e
        assert neg.shape == (512, smp.num_samples[slot]) and neg.dtype == torch.int64
        want = so.sample_uniform(seed, off, vocab, neg.numel()).reshape(neg.shape)
        assert np.array_equal(neg.cpu().numpy(), want)
        assert int(neg.min()) >= 0 and int(neg.max()) < vocab
    # successive calls draw different numbers; the same state reproduces them
    a = smp.sample(triples, 0)
    b = smp.sample(triples, 0)
    assert not torch.equal(a, b)
    big = kb.KgeUniformSampler(97, 5, num_samples=(1, 0, 1), seed=5).sample(torch.zeros(200000, 3, dtype=torch.long, device="cuda"), 0)
    cnt = torch.bincount(big.view(-1), minlength=97).cpu().numpy()
    expect = 200000 / 97
    assert ((cnt - expect) ** 2 / expect).sum() < 160          # chi-square, 96 dof: p ~ 1e-5


def test_filtered_sampling_removes_known_positives(kb):
    g = kb.graph.synthetic_graph("toy", seed=2)
    e, r = g["num_entities"], g["num_relations"]
    train = g["train"]
    smp = kb.KgeUniformSampler(e, r, num_samples=(40, 0, 40), filter_positives=(True, False, True), train_triples=train,
                               seed=99)
    triples = T(train[:300].astype(np.int64)).cuda()
    for slot, pair, cols in ((0, "po", (1, 2)), (2, "sp", (0, 1))):
        seed, off = _state(smp)
        neg = smp.sample(triples, slot).cpu().numpy()
        smp.check_status()
        keys, offs, vals = ko.kvsall_index(train, pair)
        index = {(int(k[0]), int(k[1])): vals[offs[i]:offs[i + 1]] for i, k in enumerate(keys)}
        tr = triples.cpu().numpy()
        raw = so.sample_uniform(seed, off, e, neg.size).reshape(neg.shape)
        want = so.sample_filter(seed, off, e, index, tr[:, cols[0]], tr[:, cols[1]], raw)
        assert np.array_equal(neg, want)                                        # bit-exact redraws
        changed = 0
        for i in range(len(tr)):
            known = set(int(x) for x in index[(int(tr[i, cols[0]]), int(tr[i, cols[1]]))])
            assert not (set(neg[i].tolist()) & known)                           # sampler.py:148-176
            changed += int((neg[i] != raw[i]).sum())
        assert changed > 0


@pytest.mark.parametrize("wr", [False, True])
def test_shared_sampling_contract_and_bit_exactness(kb, wr):
    e, r, b, n = 2000, 7, 300, 64
    smp = kb.KgeUniformSampler(e, r, num_samples=(n, 0, n), shared=True, with_replacement=wr, seed=31)
    triples = torch.randint(0, e, (b, 3), device="cuda")
    triples[:, 1] %= r
    for slot in (0, 2):
        seed, off = _state(smp)
        neg = smp.sample(triples, slot).cpu().numpy()
        smp.check_status()
        pos = triples[:, slot].cpu().numpy()
        want, nd, shared = so.sample_shared(seed, off, e, pos, n, wr)
        assert int(smp.meta[0].item()) == nd
        assert np.array_equal(neg, want)
        assert nd == n if not wr else nd <= n
        for i in range(b):
            assert pos[i] not in neg[i]                                          # own positive dropped (sampler.py:234-247)
            assert len(set(neg[i, :nd].tolist())) == nd                          # WOR part is distinct
            assert (neg[i, :nd] != shared[:nd]).sum() <= 1                       # at most one position replaced
            assert set(neg[i, nd:].tolist()) <= set(neg[i, :nd].tolist())        # WR upsample copies columns
    with pytest.raises(ValueError):
        kb.KgeUniformSampler(e, r, shared=False, with_replacement=False)         # sampler.py:27-31
    with pytest.raises(ValueError):
        kb.KgeUniformSampler(e, r, shared=True, filter_positives=(True, False, False), train_triples=np.zeros((1, 3)))
    # a tiny vocabulary: the positive is almost always among the shared samples
    small = kb.KgeUniformSampler(12, 3, num_samples=(10, 0, 10), shared=True, with_replacement=False, seed=3)
    tr = torch.randint(0, 3, (50, 3), device="cuda")
    out = small.sample(tr, 2).cpu().numpy()
    for i in range(50):
        assert len(set(out[i].tolist())) == 10 and int(tr[i, 2]) not in out[i]


def test_negative_sampling_step_with_device_sampler(kb):
    """step_triples(): positives in, negatives drawn on the device into the captured step's inputs; equals the same
    step fed with those negatives through the reference-style batch."""
    g = kb.graph.synthetic_graph("toy", seed=1)
    e, r, d, b, n = g["num_entities"], g["num_relations"], 32, 64, 16
    torch.manual_seed(0)
    models = [kb.KgeModel("transe", e, r, d).cuda() for _ in range(2)]
    models[1].load_state_dict(models[0].state_dict())
    jobs = []
    for m in models:
        job = kb.TrainingJobNegativeSampling(m, kb.optim.create("Adagrad", m.parameters(), lr=0.1), kb.KgeLoss.create("kl"))
        job.enable_graph_step(b, n, n, use_graph=True)
        jobs.append(job)
    smp = jobs[0].enable_device_sampling(kb.KgeUniformSampler(e, r, num_samples=(n, 0, n), seed=8))
    twin = kb.KgeUniformSampler(e, r, num_samples=(n, 0, n), seed=8)
    for step in range(3):
        triples = T(g["train"][step * b:(step + 1) * b].astype(np.int64))
        res = jobs[0].step_triples(triples)
        tc = triples.cuda()
        negs = [twin.sample(tc, 0), torch.zeros(b, 0, dtype=torch.long, device="cuda"), twin.sample(tc, 2)]
        assert torch.equal(negs[0], jobs[0].stepper.neg[0]) and torch.equal(negs[2], jobs[0].stepper.neg[2])
        ref = jobs[1].step(step, {"triples": triples, "negative_samples": negs})
        assert np.isfinite(res.avg_loss) and res.avg_loss == ref.avg_loss
    assert torch.equal(models[0].get_s_embedder().weight, models[1].get_s_embedder().weight)


# ---------------------------------------------------------------------------------------------
# 8f-3 penalties
# ---------------------------------------------------------------------------------------------
def _penalty_model(kb, g, tag, p, weighted):
    ent, rel = g[tag + ".ent"], g[tag + ".rel"]
    m = kb.KgeModel("complex", ent.shape[0], rel.shape[0], ent.shape[1],
                    entity_embedder=dict(regularize_weight=0.05, regularize_p=p, regularize_weighted=weighted),
                    relation_embedder=dict(regularize_weight=0.02, regularize_p=p, regularize_weighted=weighted)).cuda()
    with torch.no_grad():
        m.get_s_embedder().weight.copy_(T(ent))
        m.get_p_embedder().weight.copy_(T(rel))
    return m


def test_penalties_match_reference_golden(kb, golden):
    g = golden("aux")
    triples = T(g["penalty.triples"])
    for tag in g["penalty.cases"]:
        tag = str(tag)
        p, weighted = int(tag.split(".")[1][1:]), tag.endswith(".weighted")
        m = _penalty_model(kb, g, tag, p, weighted)
        terms = m.penalty(batch={"triples": triples})
        assert [k for k, _ in terms] == [str(k) for k in g[tag + ".keys"]]
        vals = []
        for _, v in terms:
            v.backward()
            vals.append(v.item())
        np.testing.assert_allclose(vals, g[tag + ".values"], rtol=2e-6, err_msg=tag)
        np.testing.assert_allclose(m.get_s_embedder().weight.grad.cpu().numpy(), g[tag + ".grad_ent"], rtol=1e-5,
                                   atol=1e-9, err_msg=tag)
        np.testing.assert_allclose(m.get_p_embedder().weight.grad.cpu().numpy(), g[tag + ".grad_rel"], rtol=1e-5,
                                   atol=1e-9, err_msg=tag)


@pytest.mark.parametrize("p", [1, 2, 3])
def test_adagrad_with_fused_penalty_equals_penalty_then_adagrad(kb, p):
    """kgeb_adagrad_dense_lp == back-propagated unweighted penalty (lookup_embedder.py:126-134) + torch Adagrad."""
    gen = torch.Generator().manual_seed(p)
    n, d = 1031, 24
    w0 = torch.randn(n, d, generator=gen) * 0.3
    grad = torch.randn(n, d, generator=gen) * 0.01
    lam = 0.03
    wref = w0.clone().requires_grad_()
    opt = torch.optim.Adagrad([wref], lr=0.1)
    value = ko.lp_penalty(wref, p, lam)
    value.backward()
    wref.grad += grad
    opt.step()
    w, state = w0.clone().cuda(), torch.zeros(n, d, device="cuda")
    mirror = torch.empty(n, d, dtype=torch.bfloat16, device="cuda")
    out = torch.empty(1, device="cuda")
    L = kb.lib.load()
    ws = torch.empty(L.kgeb_penalty_workspace_bytes(0, n * d), dtype=torch.uint8, device="cuda")
    kb.lib.call("kgeb_adagrad_dense_lp", w.data_ptr(), state.data_ptr(), grad.cuda().data_ptr(), None, n * d, 0.1, 1e-10, 0.0,
                p, lam, mirror.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(), kb.lib.stream_ptr(w))
    assert out.item() == pytest.approx(value.item(), rel=1e-5)
    np.testing.assert_allclose(state.cpu().numpy(), opt.state[wref]["sum"].numpy(), rtol=1e-5, atol=1e-12)
    np.testing.assert_allclose(w.cpu().numpy(), wref.detach().numpy(), rtol=0, atol=2e-6)
    assert torch.equal(mirror, w.bfloat16())
    # deterministic value
    out2 = torch.empty(1, device="cuda")
    kb.lib.call("kgeb_lp_penalty_dense", w.data_ptr(), n * d, p, lam, None, out2.data_ptr(), ws.data_ptr(), ws.numel(),
                kb.lib.stream_ptr(w))
    out3 = torch.empty(1, device="cuda")
    kb.lib.call("kgeb_lp_penalty_dense", w.data_ptr(), n * d, p, lam, None, out3.data_ptr(), ws.data_ptr(), ws.numel(),
                kb.lib.stream_ptr(w))
    assert out2.item() == out3.item()


# ---------------------------------------------------------------------------------------------
# 8f-4 histograms and metrics
# ---------------------------------------------------------------------------------------------
def test_grouped_metrics_match_reference_golden(kb, golden):
    g = golden("aux")
    graph = {k: g[f"metrics.graph.{k}"] for k in ("train", "valid", "test")}
    ent, rel = g["metrics.ent"], g["metrics.rel"]
    m = kb.KgeModel("complex", ent.shape[0], rel.shape[0], ent.shape[1]).cuda()
    with torch.no_grad():
        m.get_s_embedder().weight.copy_(T(ent))
        m.get_p_embedder().weight.copy_(T(rel))
    rpt = {str(t): g[f"metrics.relations_of.{t}"].tolist() for t in g["metrics.relation_types"]}
    ev = kb.EntityRankingJob(m, ent.shape[0], [graph["train"], graph["valid"]], graph["test"], batch_size=16,
                             hits_at_k_s=(1, 3, 10))
    got = ev.run(graph["valid"], head_and_tail=True, relations_per_type=rpt)["metrics"]
    keys = [str(k) for k in g["metrics.keys"]]
    assert len(keys) >= 60
    for k in keys:
        assert k in got, k
        assert abs(got[k] - float(g[f"metrics.value.{k}"])) <= 1e-6, (k, got[k], float(g[f"metrics.value.{k}"]))


def test_rank_hist_and_metrics_kernels(kb):
    from importlib import import_module
    dm = import_module("kge-1_b200.metrics")
    e, n = 4_600_000, 200_000
    gen = torch.Generator().manual_seed(0)
    ranks = (torch.rand(n, generator=gen) ** 4 * e).long().clamp_(0, e - 1)
    mask = (torch.rand(n, generator=gen) < 0.3).to(torch.uint8)
    hist = dm.rank_hist(ranks.cuda(), e)
    assert torch.equal(hist.cpu(), torch.bincount(ranks, minlength=e).float())              # exact integer counts
    hm = dm.rank_hist(ranks.cuda(), e, mask=mask.cuda())
    assert torch.equal(hm.cpu(), torch.bincount(ranks[mask.bool()], minlength=e).float())
    got = dm.rank_metrics(hist, (1, 3, 10, 100))
    r = ranks.double() + 1
    assert got["mean_rank"] == pytest.approx(r.mean().item(), rel=1e-12)
    assert got["mean_reciprocal_rank"] == pytest.approx((1 / r).mean().item(), rel=1e-12)
    for k in (1, 3, 10, 100):
        assert got[f"hits_at_{k}"] == pytest.approx((ranks < k).double().mean().item(), rel=1e-12)
    want = ko.metrics_from_hist(hist.cpu(), (1, 3, 10, 100))                                  # reference fp32 arithmetic
    for k, v in want.items():
        assert abs(got[k] - v) <= 1e-3 * max(1.0, abs(v)), k
    empty = dm.rank_metrics(torch.zeros(e, device="cuda"), (1, 10))
    assert empty == {"mean_rank": 0.0, "mean_reciprocal_rank": 0.0, "hits_at_1": 0.0, "hits_at_10": 0.0}
    vals = torch.randint(0, 1000, (5000,), generator=gen)
    sset = torch.unique(torch.randint(0, 1000, (100,), generator=gen))
    for v in (vals, vals.int()):
        assert torch.equal(dm.isin_sorted(v.cuda(), sset.cuda()).cpu().bool(), torch.isin(vals, sset))
    st = torch.zeros(1, dtype=torch.int32, device="cuda")
    dm.rank_hist(torch.tensor([5, e], device="cuda"), e, status=st)
    assert st.item() == 1


@pytest.mark.parametrize("p", [2, 3])
def test_captured_kvsall_step_with_fused_penalty_matches_autograd_flow(kb, p):
    """The toy config (examples/toy-complex-train.yaml: ComplEx KvsAll + lookup_embedder.regularize_weight) through
    the captured step with the penalty inside the Adagrad kernels == the reference's order of operations
    (loss backward, each penalty term backward, optimizer.step; train.py:309-376) on the autograd path."""
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
    reg = dict(entity_embedder=dict(regularize_weight=0.01, regularize_p=p),
               relation_embedder=dict(regularize_weight=0.003, regularize_p=p))
    ref = kb.KgeModel("complex", e, r, d, **reg).cuda()
    new = kb.KgeModel("complex", e, r, d, **reg).cuda()
    new.load_state_dict(ref.state_dict())
    # a non-zero initial accumulator: with sum = 0 the first Adagrad step is lr * g / (|g| + 1e-10), which turns the
    # ~1e-8 summation-order differences between the two loss-gradient paths into O(lr) differences wherever g ~ 0
    mk = lambda m: kb.optim.create("Adagrad", m.parameters(), lr=0.2, initial_accumulator_value=0.1)  # noqa: E731
    jr = kb.TrainingJobKvsAll(ref, mk(ref), kb.KgeLoss.create("kl"), e, r)
    jn = kb.TrainingJobKvsAll(new, mk(new), kb.KgeLoss.create("kl"), e, r)
    jn.enable_graph_step(b, nnz_max, use_graph=True)
    for i, batch in enumerate(batches):
        a, c = jr.step(i, batch), jn.step(i, batch)
        assert c.total_loss == pytest.approx(a.total_loss, rel=1e-5)
        assert c.avg_loss == pytest.approx(a.avg_loss, rel=1e-5)   # value of the last query type (train.py:747)
        assert c.penalty == pytest.approx(a.penalty, rel=2e-5) and a.penalty > 0
        for got, want in ((new.get_s_embedder().weight, ref.get_s_embedder().weight),
                          (new.get_p_embedder().weight, ref.get_p_embedder().weight)):
            err = (got - want).abs().max().item()
            assert err <= 1e-5 * want.abs().max().item() + 1e-6, (i, err)


def test_prefetched_packed_batches_equal_direct_copies(kb):
    """prefetch_packed(): the double-buffered H2D staging feeds step() the same bytes as the direct copy."""
    g = kb.graph.synthetic_graph("toy", seed=3)
    e, r, d, b = g["num_entities"], g["num_relations"], 32, 64
    idx = [ko.kvsall_index(g["train"], "sp"), ko.kvsall_index(g["train"], "po")]
    rng = np.random.default_rng(4)
    batches = []
    for _ in range(5):
        ids = rng.choice(len(idx[0][0]) + len(idx[1][0]), b, replace=False)
        q, c, qt = ko.kvsall_collate(ids.tolist(), idx)
        batches.append({"queries": T(q), "label_coords": T(c), "query_type_indexes": T(qt)})
    nnz_max = max(len(x["label_coords"]) for x in batches)
    res = []
    for prefetch in (False, True):
        torch.manual_seed(0)
        m = kb.KgeModel("complex", e, r, d).cuda()
        job = kb.TrainingJobKvsAll(m, kb.optim.create("Adagrad", m.parameters(), lr=0.2), kb.KgeLoss.create("bce"), e, r,
                                   math_mode=kb.lib.MATH_BF16)
        job.enable_graph_step(b, nnz_max)
        packed = [job.collate_packed(x) for x in batches]
        losses = []
        for i, pb in enumerate(packed):
            if prefetch and i + 1 < len(packed):
                job.prefetch_packed(packed[i + 1])
            losses.append(job.step(i, pb).avg_loss)
        res.append((losses, m.get_s_embedder().weight.detach().clone()))
    assert res[0][0] == res[1][0] and torch.equal(res[0][1], res[1][1])


@pytest.mark.parametrize("model", ["complex", "transe"])
@pytest.mark.parametrize("ties", ["rounded_mean_rank", "best_rank", "worst_rank"])
def test_ranking_edge_cases_all_ties_and_nan_scores(kb, model, ties):
    """Edge cases of _filter_and_rank / _get_ranks_and_num_ties (entity_ranking.py:469-551): every candidate tied with
    the true triple (all-zero tables), and NaN scores, which count as -inf (:500-507) -- also for the true triple."""
    g = kb.graph.synthetic_graph("toy", seed=5)
    e, r, d = g["num_entities"], g["num_relations"], 16
    gen = torch.Generator().manual_seed(7)
    valid = g["valid"][:40]
    for case in ("zeros", "nan"):
        if case == "zeros":
            ent, rel = torch.zeros(e, d), torch.zeros(r, d)
        else:
            if model == "complex":
                ent = torch.round(torch.randn(e, d, generator=gen) * 4) / 4    # dyadic: sums are exact in fp32 and TF32
                rel = torch.round(torch.randn(r, d, generator=gen) * 4) / 4
            else:
                # TransE's true score carries pairwise_distance's eps (transe.py:17, +1e-6 per coordinate) while the
                # candidates' scores do not: on dyadic tables a candidate at exactly the true distance is ordered by the
                # rounding of sum |diff + 1e-6|, i.e. by the summation order.  Generic weights keep that boundary empty.
                ent = torch.randn(e, d, generator=gen) * 0.5
                rel = torch.randn(r, d, generator=gen) * 0.5
            bad = torch.randperm(e, generator=gen)[:9]
            ent[bad] = float("nan")
            ent[int(valid[0, 2])] = float("nan")                               # a true object with a NaN score
        want, wranks = ko.entity_ranking(model, ent, rel, valid, [g["train"], g["valid"]], g["test"], batch_size=16,
                                         tie_handling=ties, hits_at_k=(1, 3, 10))
        for math_mode in ((kb.lib.MATH_FP32, kb.lib.MATH_TF32) if model == "complex" else (kb.lib.MATH_FP32,)):
            m = kb.KgeModel(model, e, r, d, math_mode=math_mode).cuda()
            with torch.no_grad():
                m.get_s_embedder().weight.copy_(ent)
                m.get_p_embedder().weight.copy_(rel)
            job = kb.EntityRankingJob(m, e, [g["train"], g["valid"]], g["test"], batch_size=16, tie_handling=ties,
                                      hits_at_k_s=(1, 3, 10), math_mode=math_mode)
            res = job.run(valid)
            for key, ref in wranks.items():
                np.testing.assert_array_equal(res["ranks"][key].cpu().numpy(), ref.numpy(),
                                              err_msg=f"{model} {ties} {case} {key} math={math_mode}")
            assert res["metrics"]["mean_reciprocal_rank_filtered"] == pytest.approx(
                want["_filt"]["mean_reciprocal_rank"], abs=1e-6)


def test_sampler_edge_cases_empty_batch(kb):
    smp = kb.KgeUniformSampler(50, 5, num_samples=(4, 0, 4), seed=1)
    empty = torch.zeros(0, 3, dtype=torch.long, device="cuda")
    assert smp.sample(empty, 0).shape == (0, 4)
    shared = kb.KgeUniformSampler(50, 5, num_samples=(4, 0, 4), shared=True, with_replacement=False, seed=1)
    assert shared.sample(empty, 2).shape == (0, 4)
    shared.check_status()
    with pytest.raises(ValueError):
        shared.sample(torch.zeros(3, 3, dtype=torch.long, device="cuda"), 0, num_samples=50)   # 51 distinct of 50


# ---------------------------------------------------------------------------------------------
# Reciprocal relations, Adam, the toy configuration with penalties, frequency drill-downs, segment backward
# (golden vectors tests/golden/aux2.npz / aux3.npz from the unmodified reference)
# ---------------------------------------------------------------------------------------------
def _close(got, ref, rtol, what):
    got = got.detach().cpu().double().numpy()
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, what
    assert np.abs(got - ref).max() <= rtol * max(np.abs(ref).max(), 1e-30), what


def test_reciprocal_relations_model_matches_reference_golden(kb, golden):
    """a15: ReciprocalRelationsModel (reciprocal_relations_model.py:56-106); fp32 tolerance 1e-5."""
    g = golden("aux2")
    s, p, o, sub = (T(g[k]).cuda() for k in ("recip.idx_s", "recip.idx_p", "recip.idx_o", "recip.subset"))
    for tag in g["recip.cases"]:
        tag = str(tag)
        base = tag.split(".")[1]
        ent, rel = g[tag + ".ent"], g[tag + ".rel"]
        m = kb.ReciprocalRelationsModel(base, ent.shape[0], rel.shape[0] // 2, ent.shape[1]).cuda()
        with torch.no_grad():
            m.get_s_embedder().weight.copy_(T(ent))
            m.get_p_embedder().weight.copy_(T(rel))
            _close(m.score_spo(s, p, o, "o").view(-1), g[tag + ".spo_o"], 1e-5, "spo o")
            _close(m.score_spo(s, p, o, "s").view(-1), g[tag + ".spo_s"], 1e-5, "spo s")
            _close(m.score_sp(s, p), g[tag + ".sp"], 1e-5, "sp")
            _close(m.score_po(p, o), g[tag + ".po"], 1e-5, "po")
            _close(m.score_sp_po(s, p, o), g[tag + ".sp_po"], 1e-5, "sp_po")
            _close(m.score_sp_po(s, p, o, sub), g[tag + ".sp_po_sub"], 1e-5, "sp_po subset")
        with pytest.raises(Exception):
            m.score_spo(s, p, o, None)
        job = kb.TrainingJob1vsAll(m, kb.optim.create("Adagrad", m.parameters(), lr=0.2), kb.KgeLoss.create("kl"),
                                   fused_path=False)
        res = job.step(0, {"triples": T(g[tag + ".b0.triples"])})
        assert res.avg_loss == pytest.approx(float(g[tag + ".b0.loss"]), rel=2e-5)
        _close(m.get_s_embedder().weight.grad, g[tag + ".b0.grad_ent"], 2e-5, "grad entity")
        _close(m.get_p_embedder().weight.grad, g[tag + ".b0.grad_rel"], 2e-5, "grad relation")


def test_adam_steps_match_reference_golden(kb, golden):
    """train.optimizer: Adam (util/optimizer.py:10-17) through kgeb_adam_dense."""
    g = golden("aux2")
    ent, rel = g["adam.ent0"], g["adam.rel0"]
    m = kb.KgeModel("distmult", ent.shape[0], rel.shape[0], ent.shape[1]).cuda()
    with torch.no_grad():
        m.get_s_embedder().weight.copy_(T(ent))
        m.get_p_embedder().weight.copy_(T(rel))
    job = kb.TrainingJob1vsAll(m, kb.optim.create("Adam", m.parameters(), lr=0.01), kb.KgeLoss.create("kl"), fused_path=False)
    for step in range(2):
        res = job.step(step, {"triples": T(g[f"adam.b{step}.triples"])})
        assert res.avg_loss == pytest.approx(float(g[f"adam.b{step}.loss"]), rel=2e-5)
        # Adam's first steps move every weight by ~lr * sign(g): compare in units of the learning rate
        for got, ref in ((m.get_s_embedder().weight, g[f"adam.b{step}.ent"]), (m.get_p_embedder().weight, g[f"adam.b{step}.rel"])):
            assert (got.detach().cpu() - T(ref)).abs().max().item() <= 0.01 * 2e-2


def test_toy_config_steps_with_penalty_match_reference_golden(kb, golden):
    """examples/toy-complex-train.yaml as written: ComplEx, KvsAll + KL, Lp penalty; run_epoch's body (train.py:309-376)
    on the autograd path (fp32) and on the captured step with the penalty folded into the Adagrad kernels."""
    g = golden("aux2")
    e, r = 53, 7
    reg = dict(entity_embedder=dict(regularize_weight=1e-2, regularize_p=2), relation_embedder=dict(regularize_weight=1e-2, regularize_p=2))
    nnz_max = max(len(g[f"toy.b{step}.label_coords"]) for step in range(2))
    for captured in (False, True):
        m = kb.KgeModel("complex", e, r, g["toy.ent0"].shape[1], **reg).cuda()
        with torch.no_grad():
            m.get_s_embedder().weight.copy_(T(g["toy.ent0"]))
            m.get_p_embedder().weight.copy_(T(g["toy.rel0"]))
        job = kb.TrainingJobKvsAll(m, kb.optim.create("Adagrad", m.parameters(), lr=0.2), kb.KgeLoss.create("kl"), e, r)
        if captured:
            job.enable_graph_step(len(g["toy.b0.queries"]), nnz_max)
        for step in range(2):
            pre = f"toy.b{step}"
            batch = {"queries": T(g[pre + ".queries"]), "label_coords": T(g[pre + ".label_coords"]),
                     "query_type_indexes": T(g[pre + ".query_type"])}
            res = job.step(step, batch)
            # the golden holds the reference's avg_loss: train.py:747 overwrites it per query type, so it is the value
            # of the last non-empty type, not the cost that was back-propagated
            assert res.avg_loss == pytest.approx(float(g[pre + ".loss"]), rel=2e-5), (captured, step)
            assert res.penalty == pytest.approx(float(g[pre + ".penalties"].sum()), rel=2e-5)
            # Adagrad's first steps: compare in units of the learning rate (see tests/test_gpu_parity.py)
            for got, ref in ((m.get_s_embedder().weight, g[pre + ".ent"]), (m.get_p_embedder().weight, g[pre + ".rel"])):
                assert (got.detach().cpu() - T(ref)).abs().max().item() <= 0.2 * 2e-3, (captured, step)


def test_argument_frequency_metrics_match_reference_golden(kb, golden):
    """eval.py:173-224 through the device kernels (kgeb_isin_sorted masks + kgeb_rank_hist + kgeb_rank_metrics); the
    host control flow is verified on the CPU in tests/test_host_logic.py."""
    g = golden("aux3")
    graph = {k: g[f"graph.{k}"] for k in ("train", "valid", "test")}
    ent, rel = g["ent"], g["rel"]
    e, r = ent.shape[0], rel.shape[0]
    m = kb.KgeModel("distmult", e, r, ent.shape[1]).cuda()
    with torch.no_grad():
        m.get_s_embedder().weight.copy_(T(ent))
        m.get_p_embedder().weight.copy_(T(rel))
    ev = kb.EntityRankingJob(m, e, [graph["train"], graph["valid"]], graph["test"], batch_size=16, hits_at_k_s=(1, 3, 10))
    got = ev.run(graph["valid"], relations_per_type=kb.index.relations_per_type(graph["train"], r),
                 frequency_percentiles=kb.index.frequency_percentiles(graph["train"], e, r))["metrics"]
    keys = [str(k) for k in g["keys"]]
    assert len(keys) > 150
    for k in keys:
        assert k in got and abs(got[k] - float(g[f"value.{k}"])) <= 1e-6, (k, got.get(k), float(g[f"value.{k}"]))


@pytest.mark.parametrize("model", ["distmult", "transe", "rotate"])
def test_segment_backward_of_negative_sampling_equals_default_path(kb, model):
    """FusedNegSamplingStepper(segment_bwd=True) (csrc/ns_segment.cu) against the default pairs_bwd + sorted-scatter path:
    same batches, same negatives; gradients are sums of the same rows in a different grouping -> equal to rounding."""
    g = kb.graph.synthetic_graph("toy", seed=1)
    e, r, d, b, n = g["num_entities"], g["num_relations"], 32, 64, 16
    torch.manual_seed(0)
    models = [kb.KgeModel(model, e, r, d).cuda() for _ in range(2)]
    models[1].load_state_dict(models[0].state_dict())
    jobs = []
    for m, seg in zip(models, (False, True)):
        opt = kb.optim.create("Adagrad", m.parameters(), lr=0.1, initial_accumulator_value=0.1)
        job = kb.TrainingJobNegativeSampling(m, opt, kb.KgeLoss.create("kl"))
        job.enable_graph_step(b, n, n, use_graph=True, segment_bwd=seg)
        jobs.append(job)
    gen = torch.Generator().manual_seed(3)
    for step in range(3):
        triples = T(g["train"][step * b:(step + 1) * b].astype(np.int64))
        negs = [torch.randint(0, e, (b, n), generator=gen), torch.zeros(b, 0, dtype=torch.long),
                torch.randint(0, e, (b, n), generator=gen)]
        negs[0][:, :4] = 3                       # a hub candidate: long segments
        a, c = (j.step(step, {"triples": triples, "negative_samples": negs}) for j in jobs)
        assert c.avg_loss == pytest.approx(a.avg_loss, rel=1e-6)
        for x, y in ((models[0].get_s_embedder().weight, models[1].get_s_embedder().weight),
                     (models[0].get_p_embedder().weight, models[1].get_p_embedder().weight)):
            assert (x - y).abs().max().item() <= 1e-5, (model, step)


@pytest.mark.parametrize("model,loss", [("rotate", "kl"), ("transe", "bce"), ("distmult", "kl"), ("complex", "bce")])
def test_fused_negative_sampling_slot_kernel(kb, model, loss):
    """kgeb_ns_fused (scores + loss + dQ + candidate gradient of a slot in one kernel) against the three-kernel slot it
    replaces: the same sums in the same order -> BIT-IDENTICAL tables after three steps; the opt-in vector-reduction form
    (deterministic=False) agrees to rounding."""
    g = kb.graph.synthetic_graph("toy", seed=1)
    e, r, d, b, n = g["num_entities"], g["num_relations"], 64, 64, 16
    l_norm = 2.0 if model == "transe" else 1.0
    torch.manual_seed(0)
    models = [kb.KgeModel(model, e, r, d, l_norm=l_norm).cuda() for _ in range(3)]
    for m in models[1:]:
        m.load_state_dict(models[0].state_dict())
    jobs = []
    for m, kw in zip(models, (dict(fused_slot=False), dict(fused_slot=True), dict(fused_slot=True, deterministic=False))):
        opt = kb.optim.create("Adagrad", m.parameters(), lr=0.1, initial_accumulator_value=0.1)
        job = kb.TrainingJobNegativeSampling(m, opt, kb.KgeLoss.create(loss))
        job.enable_graph_step(b, n, n, use_graph=True, **kw)
        jobs.append(job)
    gen = torch.Generator().manual_seed(3)
    for step in range(3):
        triples = T(g["train"][step * b:(step + 1) * b].astype(np.int64))
        negs = [torch.randint(0, e, (b, n), generator=gen), torch.zeros(b, 0, dtype=torch.long),
                torch.randint(0, e, (b, n), generator=gen)]
        negs[2][:, :3] = 7                       # a hub candidate: many additions to one row
        res = [j.step(step, {"triples": triples, "negative_samples": negs}) for j in jobs]
        assert res[1].avg_loss == res[0].avg_loss
        assert res[2].avg_loss == pytest.approx(res[0].avg_loss, rel=1e-6)
        for get in (lambda m: m.get_s_embedder().weight, lambda m: m.get_p_embedder().weight):
            assert torch.equal(get(models[1]), get(models[0])), (model, step)
            assert (get(models[2]) - get(models[0])).abs().max().item() <= 1e-5, (model, step)
