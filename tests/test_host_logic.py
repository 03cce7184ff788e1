"""CPU tests of the host-side logic: CSR / index construction against the reference's golden arrays, the synthetic
graph generator, loss-from-statistics algebra, shard partitioning and -- with a world_size-2 gloo group -- the
collective combination of per-shard row statistics (the multi-GPU path of SURVEY.md 8e).  No GPU needed."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import kgeb200 as kb
from oracle import kge_oracle as ko

T = torch.from_numpy


def test_kvsall_index_matches_reference_arrays(golden):
    g = golden("index")
    for split in ("train", "valid"):
        for key in ("sp", "po"):
            ix = kb.index.KvsAllIndex(g[f"index.{split}.triples"], key)
            np.testing.assert_array_equal(ix._keys.numpy(), g[f"index.{split}.{key}.keys"])
            np.testing.assert_array_equal(ix._values_offset.numpy(), g[f"index.{split}.{key}.offsets"])
            np.testing.assert_array_equal(ix._values.numpy(), g[f"index.{split}.{key}.values"])
    ix = kb.index.KvsAllIndex(g["index.train.triples"], "sp")
    k = tuple(int(x) for x in g["index.train.sp.keys"][3])
    off = g["index.train.sp.offsets"]
    np.testing.assert_array_equal(ix.get(k).numpy(), g["index.train.sp.values"][off[3]:off[4]])
    assert len(ix.get((10 ** 6, 0))) == 0     # default_factory list() in the reference (indexing.py:57-63)


def test_csr_from_coords_and_gather_rows():
    coords = torch.tensor([[2, 5], [0, 7], [2, 1], [0, 3], [3, 9]])
    off, col = kb.fused.csr_from_coords(coords, 5)
    assert off.tolist() == [0, 2, 2, 4, 5, 5]
    assert col.tolist() == [3, 7, 1, 5, 9]
    off, col = kb.fused.csr_from_coords(torch.zeros(0, 2, dtype=torch.long), 3)
    assert off.tolist() == [0, 0, 0, 0] and col.numel() == 0
    offsets = torch.tensor([0, 2, 2, 5])
    values = torch.tensor([10, 11, 20, 21, 22])
    o, v = kb.index.gather_csr_rows(offsets, values, torch.tensor([2, -1, 0, 1]), add=100)
    assert o.tolist() == [0, 3, 3, 5, 5] and v.tolist() == [120, 121, 122, 110, 111]
    mo, mc = kb.index.merge_sorted_csr([(o, v), (torch.tensor([0, 1, 1, 1, 2]), torch.tensor([5, 7]))], 4)
    assert mo.tolist() == [0, 4, 4, 6, 7] and mc.tolist() == [5, 120, 121, 122, 110, 111, 7]


def test_kvsall_rows_and_oracle_collate_agree(golden):
    g = golden("train")
    tag = "train.KvsAll.complex.bce.b0"
    q, qt = T(g[tag + ".queries"]), T(g[tag + ".query_type"])
    a, p, rc = kb.trainer.kvsall_rows(q, qt)
    sp = qt == 0
    assert torch.equal(a[sp], q[sp, 0]) and torch.equal(p[sp], q[sp, 1])      # sp_: (s, p)
    assert torch.equal(a[~sp], q[~sp, 1]) and torch.equal(p[~sp], q[~sp, 0])  # _po: keys are (p, o)
    assert rc.dtype == torch.int32 and set(rc.tolist()) <= {0, 1}
    # the reference's collate emits coordinates grouped by row with ascending labels: CSR == plain counting
    off, col = kb.fused.csr_from_coords(T(g[tag + ".label_coords"]), len(q))
    assert torch.equal(col, T(g[tag + ".label_coords"])[:, 1].long())


def test_synthetic_graph_generator():
    g = kb.graph.synthetic_graph("toy", seed=5)
    e, r = g["num_entities"], g["num_relations"]
    allt = np.concatenate([g["train"], g["valid"], g["test"]])
    assert g["train"].dtype == np.int32 and g["train"].shape == (4565, 3)
    assert len(np.unique(allt, axis=0)) == len(allt)                     # no duplicates, disjoint splits
    assert allt[:, [0, 2]].max() < e and allt[:, 1].max() < r and allt.min() >= 0
    sp = kb.index.KvsAllIndex(g["train"], "sp")
    assert len(sp) < len(g["train"])                                      # Zipf draws give multi-answer keys
    g2 = kb.graph.synthetic_graph("toy", seed=5)
    np.testing.assert_array_equal(g["train"], g2["train"])               # seeded


def test_rows_loss_algebra_matches_oracle_losses():
    gen = torch.Generator().manual_seed(0)
    b, e = 7, 31
    x = torch.randn(b, e, generator=gen) * 3
    cols = torch.stack([torch.randperm(e, generator=gen)[:3].sort().values for _ in range(b)])
    lab_off = torch.arange(0, 3 * b + 1, 3)
    y = torch.zeros(b, e).scatter_(1, cols, 1.0)
    # KL: statistics (max, sum exp, sum x, sum over labels)
    st = torch.stack((x.max(1).values, torch.exp(x - x.max(1, keepdim=True).values).sum(1), x.sum(1),
                      x.gather(1, cols).sum(1)), 1)
    rows, lse = kb.fused.rows_loss(st, lab_off, kb.lib.LOSS_KL, 0.0, e)
    assert rows.sum().item() == pytest.approx(ko.loss_kl(x, y).item(), rel=1e-5)
    torch.testing.assert_close(lse, torch.logsumexp(x, 1))
    # BCE with offset and label smoothing (train.py:715-721)
    off, ls = 0.3, 0.1
    xo = x + off
    st = torch.stack((torch.nn.functional.softplus(xo).sum(1), torch.zeros(b), xo.sum(1), xo.gather(1, cols).sum(1)), 1)
    rows, _ = kb.fused.rows_loss(st, lab_off, kb.lib.LOSS_BCE, ls, e)
    assert rows.sum().item() == pytest.approx(ko.loss_bce(x, (1 - ls) * y + 1.0 / e, off).item(), rel=1e-5)


def test_shard_partition_covers_entities():
    for e, w in ((14541, 8), (10, 4), (7, 8), (4_600_000, 8)):
        shards = [kb.fused.Shard.of_rank(e, r, w) for r in range(w)]
        assert shards[0].e_lo == 0 and shards[-1].e_hi == e
        for a, b in zip(shards, shards[1:]):
            assert a.e_hi == b.e_lo and a.e_lo <= a.e_hi
        assert not shards[0].distributed


def _gloo_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gen = torch.Generator().manual_seed(1)
        b, e = 9, 40
        x = torch.randn(b, e, generator=gen) * 4
        cols = torch.stack([torch.randperm(e, generator=gen)[:2].sort().values for _ in range(b)])
        sh = kb.fused.Shard.of_rank(e, rank, world, dist.group.WORLD)
        assert sh.distributed
        xs = x[:, sh.e_lo:sh.e_hi]
        in_sh = (cols >= sh.e_lo) & (cols < sh.e_hi)
        labdot = torch.where(in_sh, x.gather(1, cols), torch.zeros(b, 2)).sum(1)
        for loss in (kb.lib.LOSS_KL, kb.lib.LOSS_BCE):
            if loss == kb.lib.LOSS_KL:
                m = xs.max(1).values
                st = torch.stack((m, torch.exp(xs - m[:, None]).sum(1), xs.sum(1), labdot), 1)
            else:
                st = torch.stack((torch.nn.functional.softplus(xs).sum(1), torch.zeros(b), xs.sum(1), labdot), 1)
            full = kb.fused.combine_rowstats(st, loss, sh)
            rows, lse = kb.fused.rows_loss(full, torch.arange(0, 2 * b + 1, 2), loss, 0.0, e)
            if loss == kb.lib.LOSS_KL:
                want = torch.logsumexp(x, 1) - x.gather(1, cols).sum(1) / 2 - np.log(2.0)
                torch.testing.assert_close(lse, torch.logsumexp(x, 1), rtol=1e-5, atol=1e-5)
            else:
                want = torch.nn.functional.softplus(x).sum(1) - x.gather(1, cols).sum(1)
            torch.testing.assert_close(rows, want, rtol=1e-5, atol=1e-4)
        # integer rank counts add exactly across shards
        t = x[:, 0:1]
        cnt = torch.stack(((xs > t).sum(1), (xs == t).sum(1)), 1)
        dist.all_reduce(cnt)
        assert torch.equal(cnt, torch.stack(((x > t).sum(1), (x == t).sum(1)), 1))
        out.put((rank, "ok"))
    except Exception as ex:  # pragma: no cover
        out.put((rank, repr(ex)))
    finally:
        dist.destroy_process_group()


def test_sharded_statistics_combine_over_gloo_world2():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    res = dict(out.get() for _ in range(2))
    assert res == {0: "ok", 1: "ok"}, res


def test_p2p_slices_partition_the_table():
    """Owner slices of the peer-memory exchange (trainer.p2p_slice == slice_of in csrc/p2p.cu)."""
    from importlib import import_module
    tr = import_module("kge-1_b200.trainer")
    for numel in (0, 4, 12, 128 * 237, 14541 * 128, 1000):
        for world in (1, 2, 3, 4, 8, 16):
            cover = 0
            for r in range(world):
                lo, hi = tr.p2p_slice(numel, world, r)
                assert lo == cover and lo <= hi <= numel and lo % 4 == 0
                assert hi == numel or (hi - lo) % 4 == 0
                cover = hi
            assert cover == numel


def test_sampler_option_errors_match_reference():
    """kge/util/sampler.py:27-31, 46-50, 67-77: option validation happens before anything touches a device."""
    from importlib import import_module
    sm = import_module("kge-1_b200.sampler")
    import pytest as _pt
    with _pt.raises(ValueError, match="Without replacement"):
        sm.KgeUniformSampler(10, 3, shared=False, with_replacement=False, device="cpu")
    with _pt.raises(ValueError, match="Filtering is not supported"):
        sm.KgeUniformSampler(10, 3, shared=True, filter_positives=(True, False, False), device="cpu")
    with _pt.raises(ValueError):
        sm.KgeSampler.create("frequency", 10, 3)
    with _pt.raises(ValueError, match="no CPU path"):
        sm.KgeUniformSampler(10, 3, device="cpu")


def test_nan_cost_aborts_like_the_reference():
    """train.py:341-345: cost = avg_loss + penalty; NaN raises FloatingPointError when train.abort_on_nan."""
    from importlib import import_module
    import pytest as _pt
    jobs = import_module("kge-1_b200.jobs")
    job = object.__new__(jobs.TrainingJob)
    job.abort_on_nan = True
    ok = jobs.ProcessBatchResult(1.5, 4)
    ok.penalty = 0.25
    assert job._check_cost(ok) is ok
    bad = jobs.ProcessBatchResult(float("nan"), 4)
    with _pt.raises(FloatingPointError, match="Cost became nan"):
        job._check_cost(bad)
    pen = jobs.ProcessBatchResult(1.0, 4)
    pen.penalty = float("nan")
    with _pt.raises(FloatingPointError):
        job._check_cost(pen)
    job.abort_on_nan = False
    assert job._check_cost(bad) is bad


def _p2p_protocol_worker(rank, world, port, out):
    """The peer-memory exchange (csrc/p2p.cu) restated with gloo collectives on host tensors: every rank contributes a
    gradient, the slice owner sums the contributions in rank order, applies Adagrad to its slice of weights and state
    and the new weights are gathered -- the result must equal all-reduce(sum) + a full Adagrad step on every rank, and
    gather_owner_slices must restore the sharded state."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from importlib import import_module
        tr = import_module("kge-1_b200.trainer")
        n, lr, eps = 1003 * 4, 0.2, 1e-10
        gen = torch.Generator().manual_seed(5)
        w0 = torch.randn(n, generator=gen)
        grads = [torch.randn(n, generator=gen) * 0.1 for _ in range(world)]     # every rank can rebuild all of them
        # reference: all-reduce in rank order + full update everywhere
        g_sum = torch.zeros(n)
        for g in grads:
            g_sum = g_sum + g
        s_ref = g_sum * g_sum
        w_ref = w0 - lr * g_sum / (s_ref.sqrt() + eps)
        # protocol: owner slices
        w, state = w0.clone(), torch.zeros(n)
        gathered = [torch.empty(n) for _ in range(world)]
        dist.all_gather(gathered, grads[rank])                                   # "peer-mapped" gradient buffers
        lo, hi = tr.p2p_slice(n, world, rank)
        g = torch.zeros(hi - lo)
        for k in range(world):
            g = g + gathered[k][lo:hi]
        state[lo:hi] += g * g
        w[lo:hi] -= lr * g / (state[lo:hi].sqrt() + eps)
        staged = [torch.empty(n) for _ in range(world)]
        dist.all_gather(staged, w)                                               # owners push their slices
        for k in range(world):
            klo, khi = tr.p2p_slice(n, world, k)
            if k != rank:
                w[klo:khi] = staged[k][klo:khi]
        assert torch.equal(w, w_ref), "weights differ from all-reduce + full update"
        assert not torch.equal(state, s_ref) and torch.equal(state[lo:hi], s_ref[lo:hi])
        tr.gather_owner_slices(state, dist.group.WORLD)
        assert torch.equal(state, s_ref), "state shards not restored"
        out.put((rank, "ok"))
    except Exception as ex:  # pragma: no cover
        out.put((rank, repr(ex)))
    finally:
        dist.destroy_process_group()


def test_peer_memory_exchange_protocol_over_gloo_world2():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_p2p_protocol_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    res = dict(out.get() for _ in range(2))
    assert res == {0: "ok", 1: "ok"}, res


def test_entity_ranking_run_control_flow_with_cpu_stand_ins(golden, monkeypatch):
    """EntityRankingJob.run's host side (which ranks go to which histogram, metric names) with the device kernels
    replaced by torch-CPU stand-ins and the ranks taken from the oracle: the metrics must equal the reference's golden
    values for the head/tail, relation-type and argument-frequency drill-downs (eval.py:138-224)."""
    from importlib import import_module
    from oracle import kge_oracle as ko
    jobs, dm, ix = (import_module("kge-1_b200." + m) for m in ("jobs", "metrics", "index"))

    def rank_hist(ranks, num_entities, hist=None, mask=None, status=None):
        r = ranks if mask is None else ranks[mask.bool()]
        hist += torch.bincount(r, minlength=num_entities).float()
        return hist

    monkeypatch.setattr(dm, "rank_hist", rank_hist)
    monkeypatch.setattr(dm, "isin_sorted", lambda v, s: torch.isin(v.long(), s).to(torch.uint8))
    monkeypatch.setattr(dm, "rank_metrics", lambda hist, ks, suffix="": {
        k + suffix: v for k, v in ko.metrics_from_hist(hist, tuple(ks)).items()})

    class Model:
        training = False

        def eval(self):
            pass

    for name, base, head_tail in (("aux", "complex", True), ("aux3", "distmult", False)):
        g = golden(name)
        pre = "metrics." if name == "aux" else ""
        graph = {k: g[f"{pre}graph.{k}"] for k in ("train", "valid", "test")}
        ent, rel = torch.from_numpy(g[pre + "ent"]), torch.from_numpy(g[pre + "rel"])
        e, r = ent.shape[0], 7
        _, ranks = ko.entity_ranking(base, ent, rel, graph["valid"], [graph["train"], graph["valid"]], graph["test"],
                                     batch_size=16, hits_at_k=(1, 3, 10))
        job = object.__new__(jobs.EntityRankingJob)
        job.model, job.num_entities, job.batch_size, job.hits_at_k_s = Model(), e, 16, (1, 3, 10)
        job.test_indexes, job.device = object(), torch.device("cpu")
        pos = {"n": 0}

        def rank_batch(batch, pos=pos, ranks=ranks):
            lo = pos["n"]
            pos["n"] += len(batch)
            return {k: v[lo:lo + len(batch)] for k, v in ranks.items()}

        job.rank_batch = rank_batch
        rpt = ix.relations_per_type(graph["train"], r)
        fp = ix.frequency_percentiles(graph["train"], e, r) if name == "aux3" else None
        got = job.run(graph["valid"].astype(np.int64), head_and_tail=head_tail, relations_per_type=rpt,
                      frequency_percentiles=fp)["metrics"]
        keys = [str(k) for k in g[pre + "keys"]]
        for k in keys:
            want = float(g[f"{pre}value.{k}"])
            assert k in got and abs(got[k] - want) <= 1e-6, (name, k, got.get(k), want)


def _cpu_stand_ins(monkeypatch):
    """Replaces the kernel-backed ops used by kge-1_b200/model.py with oracle calls on host tensors, so that the model
    API's host logic (which embeddings go where, index arithmetic, views) can be checked without a GPU."""
    from importlib import import_module
    from oracle import kge_oracle as ko
    ops, lib = import_module("kge-1_b200.ops"), import_module("kge-1_b200.lib")

    class Query:
        def __init__(self, model, combine, a, p):
            self.model, self.combine, self.a, self.p = model, combine, a, p

    monkeypatch.setattr(ops, "gather_rows", lambda w, idx, sparse=False: w[idx.long()])
    monkeypatch.setattr(ops, "score_spo_emb", lambda model, ln, s, p, o: ko.score_emb(model, s, p, o, "spo", ln).view(-1))
    monkeypatch.setattr(ops, "score_spo_indexed",
                        lambda model, ln, ew, rw, s, p, o: ko.score_spo(model, ew, rw, s.long(), p.long(), o.long(), ln).view(-1))
    monkeypatch.setattr(ops, "query_build", lambda model, combine, a, p: Query(model, combine, a, p))
    monkeypatch.setattr(ops, "score_all", lambda kind, q, cand, math=0: (
        ko.score_emb(q.model, q.a, q.p, cand, "sp_") if q.combine == lib.SP_ else ko.score_emb(q.model, cand, q.p, q.a, "_po")))
    return ko


def test_model_api_host_logic_against_reference_golden(golden, monkeypatch):
    """KgeModel / ReciprocalRelationsModel (kge_model.py:620-746, reciprocal_relations_model.py:56-106): with the kernels
    replaced by oracle stand-ins the model API must reproduce the reference's scores bit for bit."""
    from importlib import import_module
    _cpu_stand_ins(monkeypatch)
    mdl = import_module("kge-1_b200.model")
    T = torch.from_numpy
    g = golden("scores")
    s, p, o, sub = (T(g[k]).long() for k in ("idx_s", "idx_p", "idx_o", "subset"))
    for tag in ("distmult", "complex", "cp", "simple", "rescal"):
        ent, rel = g[f"{tag}.ent"], g[f"{tag}.rel"]
        m = mdl.KgeModel(tag, ent.shape[0], rel.shape[0], ent.shape[1], relation_dim=rel.shape[1])
        with torch.no_grad():
            m.get_s_embedder().weight.copy_(T(ent))
            m.get_p_embedder().weight.copy_(T(rel))
            for got, key in ((m.score_spo(s, p, o), "spo"), (m.score_sp(s, p), "sp"), (m.score_po(p, o), "po"),
                             (m.score_sp(s, p, sub), "sp_sub"), (m.score_sp_po(s, p, o, sub), "sp_po_sub"),
                             (m.score_so(s, o), "so"), (m.score_spo(s.int(), p.int(), o.int()), "spo_i32")):
                np.testing.assert_array_equal(got.numpy().reshape(g[f"{tag}.{key}"].shape), g[f"{tag}.{key}"], err_msg=f"{tag} {key}")
    g = golden("aux2")
    s, p, o, sub = (T(g[k]).long() for k in ("recip.idx_s", "recip.idx_p", "recip.idx_o", "recip.subset"))
    for tag in g["recip.cases"]:
        tag = str(tag)
        ent, rel = g[tag + ".ent"], g[tag + ".rel"]
        m = mdl.ReciprocalRelationsModel(tag.split(".")[1], ent.shape[0], rel.shape[0] // 2, ent.shape[1])
        with torch.no_grad():
            m.get_s_embedder().weight.copy_(T(ent))
            m.get_p_embedder().weight.copy_(T(rel))
            for got, key in ((m.score_spo(s, p, o, "o"), "spo_o"), (m.score_spo(s, p, o, "s"), "spo_s"),
                             (m.score_sp(s, p), "sp"), (m.score_po(p, o), "po"), (m.score_sp_po(s, p, o), "sp_po"),
                             (m.score_sp_po(s, p, o, sub), "sp_po_sub")):
                np.testing.assert_array_equal(got.numpy().reshape(g[f"{tag}.{key}"].shape), g[f"{tag}.{key}"], err_msg=f"{tag} {key}")
        with pytest.raises(Exception, match="undirected"):
            m.score_spo(s, p, o, None)
        with pytest.raises(Exception, match="cannot score relations"):
            m.score_so(s, o)


def test_training_job_bodies_host_logic_against_reference_golden(golden, monkeypatch):
    """TrainingJob1vsAll / KvsAll / NegativeSampling._process_batch in their reference flow (train.py:679-756, 823-999,
    1032-1062) and TrainingJob.step (train.py:309-376) with the kernels replaced by oracle stand-ins: losses, gradients
    and post-Adagrad parameters of the reference's golden batches."""
    import ast
    from importlib import import_module
    _cpu_stand_ins(monkeypatch)
    mdl, jobs = import_module("kge-1_b200.model"), import_module("kge-1_b200.jobs")
    T = torch.from_numpy
    g = golden("train")
    e, r = 53, 7
    for tag in [str(x) for x in g["train.cases"]]:
        _, ttype, model, loss = tag.split(".")
        opts = dict(ast.literal_eval(str(g[tag + ".options"]))) if tag + ".options" in g else {}
        ent0, rel0 = g[tag + ".ent0"], g[tag + ".rel0"]
        m = mdl.KgeModel(model, e, r, ent0.shape[1], l_norm=float(opts.get(model + ".l_norm", 1.0)), relation_dim=rel0.shape[1])
        with torch.no_grad():
            m.get_s_embedder().weight.copy_(T(ent0))
            m.get_p_embedder().weight.copy_(T(rel0))
        opt = torch.optim.Adagrad(m.parameters(), lr=0.2)
        lossf = jobs.KgeLoss.create(loss, float(opts.get("train.loss_arg", float("nan"))))
        if ttype == "1vsAll":
            job = jobs.TrainingJob1vsAll(m, opt, lossf, fused_path=False)
        elif ttype == "KvsAll":
            job = jobs.TrainingJobKvsAll(m, opt, lossf, e, r, label_smoothing=float(opts.get("KvsAll.label_smoothing", 0.0)),
                                         fused_path=False)
        else:
            job = jobs.TrainingJobNegativeSampling(m, opt, lossf, fused_path=False)
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
            assert res.avg_loss == pytest.approx(float(g[pre + ".loss"]), rel=1e-6), tag
            np.testing.assert_allclose(m.get_s_embedder().weight.grad.numpy(), g[pre + ".grad_ent"], rtol=0, atol=1e-7, err_msg=tag)
            np.testing.assert_allclose(m.get_p_embedder().weight.grad.numpy(), g[pre + ".grad_rel"], rtol=0, atol=1e-7, err_msg=tag)
            np.testing.assert_allclose(m.get_s_embedder().weight.detach().numpy(), g[pre + ".ent"], rtol=0, atol=1e-6, err_msg=tag)
            np.testing.assert_allclose(m.get_p_embedder().weight.detach().numpy(), g[pre + ".rel"], rtol=0, atol=1e-6, err_msg=tag)


def _row_exchange_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        e, d = 37, 6
        gen = torch.Generator().manual_seed(5)
        table = torch.randn(e, d, generator=gen)                      # the same full table on every rank (the oracle)
        ex = kb.trainer.RowExchange(e, dist.group.WORLD)
        shard = table[ex.lo:ex.hi].clone()
        g2 = torch.Generator().manual_seed(100 + rank)
        ids = torch.randint(0, e, (23 + 5 * rank,), generator=g2)
        ids[:4] = 0                                                   # duplicates, rows of the first owner
        ids[4] = e - 1                                                # the last row (short last shard)
        plan = ex.plan(ids)
        rows = ex.fetch(plan, lambda w: shard[w])
        assert torch.equal(rows, table[ids]), "fetched rows"
        # gradient rows back to the owners: what arrives, summed per local row, equals index_add over ALL ranks' requests
        grads = torch.randn(ids.numel(), d, generator=g2)
        wanted, recv = ex.push(plan, grads)
        mine = torch.zeros(ex.hi - ex.lo, d).index_add_(0, wanted, recv)
        full = torch.zeros(e, d).index_add_(0, ids, grads)
        dist.all_reduce(full)
        torch.testing.assert_close(mine, full[ex.lo:ex.hi], rtol=1e-6, atol=1e-6)
        out.put((rank, "ok"))
    except Exception as exc:  # noqa: BLE001
        out.put((rank, repr(exc)))
    finally:
        dist.destroy_process_group()


def test_row_exchange_routes_requests_and_gradients_over_gloo_world2():
    """SURVEY.md 8e second row (negative sampling with sharded rows): RowExchange.plan / fetch / push -- requested rows come
    back in request order, gradient rows reach their owners; world_size-2 gloo."""
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    procs = [ctx.Process(target=_row_exchange_worker, args=(r, 2, 29633, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get() for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_captured_steppers_refuse_what_they_would_silently_drop():
    """a captured step reads the raw tables without autograd; reciprocal-relations models, embedder dropout,
    per-batch renormalisation and (outside the all-entity stepper's folded Lp term) penalties must raise, not be ignored.
    The refusal comes before anything touches a device, so it is checked here on CPU models."""
    import kgeb200 as kb
    e, r, d = 50, 5, 16
    mk = lambda m: kb.optim.create("Adagrad", m.parameters(), lr=0.1)   # noqa: E731
    cases = {
        "reciprocal": kb.model.ReciprocalRelationsModel("distmult", e, r, d),
        "dropout": kb.KgeModel("distmult", e, r, d, entity_embedder={"dropout": 0.2}),
        "normalize": kb.KgeModel("distmult", e, r, d, relation_embedder={"normalize_p": 2.0}),
    }
    for name, m in cases.items():
        with pytest.raises(NotImplementedError):
            kb.trainer.FusedAllEntityStepper(m, mk(m), 8, 8, kb.lib.LOSS_KL, 4, use_graph=False)
        with pytest.raises(NotImplementedError):
            kb.trainer.FusedNegSamplingStepper(m, mk(m), 4, 2, 2, kb.lib.LOSS_KL, use_graph=False)
    pen = kb.KgeModel("distmult", e, r, d, entity_embedder={"regularize_weight": 1e-3})
    with pytest.raises(NotImplementedError):
        kb.trainer.FusedNegSamplingStepper(pen, mk(pen), 4, 2, 2, kb.lib.LOSS_KL, use_graph=False)
    weighted = kb.KgeModel("distmult", e, r, d, entity_embedder={"regularize_weight": 1e-3, "regularize_weighted": True})
    with pytest.raises(NotImplementedError):
        kb.trainer.FusedAllEntityStepper(weighted, mk(weighted), 8, 8, kb.lib.LOSS_KL, 4, use_graph=False)
