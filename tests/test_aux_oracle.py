"""CPU checks of the oracles for the rows around the hot path: the Philox restatement against the
Random123 known-answer vectors, the penalty / metric restatements against golden vectors from the unmodified
reference (tests/golden/make_golden_aux.py -> aux.npz), and the reference's sampling contract on the oracle sampler."""
import os

import numpy as np
import pytest
import torch

from oracle import kge_oracle as ko
from oracle import sampler_oracle as so

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def aux():
    return np.load(os.path.join(GOLD, "aux.npz"), allow_pickle=False)


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = so.philox4x32_10(np.array([ctr], dtype=np.uint32), np.array(key, dtype=np.uint32))[0]
        assert tuple(int(x) for x in got) == want


def test_penalty_oracle_matches_reference(aux):
    triples = torch.from_numpy(aux["penalty.triples"])
    for tag in aux["penalty.cases"]:
        tag = str(tag)
        p = int(tag.split(".")[1][1:])
        weighted = tag.endswith(".weighted")
        ent = torch.from_numpy(aux[tag + ".ent"]).requires_grad_()
        rel = torch.from_numpy(aux[tag + ".rel"]).requires_grad_()
        vals = []
        for v in ko.model_penalties(ent, rel, triples, p, 0.05, 0.02, weighted):
            v.backward()
            vals.append(v.item())
        np.testing.assert_allclose(vals, aux[tag + ".values"], rtol=1e-6)
        np.testing.assert_allclose(ent.grad.numpy(), aux[tag + ".grad_ent"], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(rel.grad.numpy(), aux[tag + ".grad_rel"], rtol=1e-6, atol=1e-9)


def _oracle_grouped_metrics(aux):
    g = {k: aux[f"metrics.graph.{k}"] for k in ("train", "valid", "test")}
    ent, rel = torch.from_numpy(aux["metrics.ent"]), torch.from_numpy(aux["metrics.rel"])
    rpt = {str(t): aux[f"metrics.relations_of.{t}"].tolist() for t in aux["metrics.relation_types"]}
    _, ranks = ko.entity_ranking("complex", ent, rel, g["valid"], [g["train"], g["valid"]], g["test"], batch_size=16,
                                 hits_at_k=(1, 3, 10))
    out = {}
    suffix = {"_raw": "", "_filt": "_filtered", "_filt_test": "_filtered_with_test"}
    rels = torch.from_numpy(g["valid"][:, 1].astype(np.int64))
    for n, sfx in suffix.items():
        hs = ko.grouped_rank_histograms({"s": ranks["s" + n], "o": ranks["o" + n]}, rels, len(ent), rpt, True)
        for key, h in hs.items():
            for k, v in ko.metrics_from_hist(h, (1, 3, 10)).items():
                out[k + sfx + ("" if key == "all" else "_" + key)] = v
    return out


def test_grouped_metrics_oracle_matches_reference(aux):
    got = _oracle_grouped_metrics(aux)
    keys = [str(k) for k in aux["metrics.keys"]]
    assert len(keys) >= 60
    for k in keys:
        assert k in got, k
        assert abs(got[k] - float(aux[f"metrics.value.{k}"])) <= 1e-6, (k, got[k], float(aux[f"metrics.value.{k}"]))


def test_sampler_oracle_contract():
    vocab, b, n = 97, 40, 12
    rng = np.random.default_rng(0)
    pos = rng.integers(0, vocab, b)
    # uniform: range and rough uniformity (sampler.py:195-198)
    u = so.sample_uniform(7, 0, vocab, 20000)
    assert u.min() >= 0 and u.max() < vocab
    cnt = np.bincount(u, minlength=vocab)
    assert cnt.min() > 120 and cnt.max() < 300
    assert not np.array_equal(u[:100], so.sample_uniform(7, 1 << 20, vocab, 100))
    # shared WOR (sampler.py:200-255): n distinct values per row, own positive never present, rows differ from the
    # shared set in at most one position
    out, nd, shared = so.sample_shared(3, 0, vocab, pos, n, with_replacement=False)
    assert nd == n and len(set(shared.tolist())) == n + 1
    for i in range(b):
        assert len(set(out[i].tolist())) == n and pos[i] not in out[i]
        assert (out[i] != shared[:n]).sum() <= 1
    # shared WR: num_distinct <= n distinct values per row, the remaining columns are copies
    out, nd, shared = so.sample_shared(5, 0, 9, rng.integers(0, 9, b), 7, with_replacement=True)
    assert 1 <= nd <= 7
    for i in range(b):
        assert len(set(out[i, :nd].tolist())) == nd and set(out[i, nd:].tolist()) <= set(out[i, :nd].tolist())
    # filtering (sampler.py:148-176): no known positive survives
    index = {(int(a), int(c)): np.sort(rng.choice(vocab, 30, replace=False)) for a, c in zip(range(b), range(b))}
    neg = so.sample_uniform(11, 0, vocab, b * n).reshape(b, n)
    f = so.sample_filter(11, 0, vocab, index, np.arange(b), np.arange(b), neg)
    for i in range(b):
        assert not set(f[i].tolist()) & set(index[(i, i)].tolist())
        keep = ~np.isin(neg[i], index[(i, i)])
        assert np.array_equal(f[i][keep], neg[i][keep])


@pytest.fixture(scope="module")
def aux2():
    return np.load(os.path.join(GOLD, "aux2.npz"), allow_pickle=False)


def test_reciprocal_relations_oracle_matches_reference(aux2):
    """a15: ReciprocalRelationsModel (reciprocal_relations_model.py:56-106) over distmult / complex."""
    T = torch.from_numpy
    s, p, o, sub = (T(aux2[k]).long() for k in ("recip.idx_s", "recip.idx_p", "recip.idx_o", "recip.subset"))
    num_rel = 7
    for tag in aux2["recip.cases"]:
        tag = str(tag)
        base = tag.split(".")[1]
        ent, rel = T(aux2[tag + ".ent"]), T(aux2[tag + ".rel"])
        assert rel.shape[0] == 2 * num_rel
        eq = lambda got, key: np.testing.assert_array_equal(got.numpy(), aux2[f"{tag}.{key}"], err_msg=f"{tag} {key}")  # noqa: E731
        eq(ko.reciprocal_score_spo(base, ent, rel, num_rel, s, p, o, "o").view(-1), "spo_o")
        eq(ko.reciprocal_score_spo(base, ent, rel, num_rel, s, p, o, "s").view(-1), "spo_s")
        eq(ko.score_sp(base, ent, rel, s, p), "sp")
        eq(ko.reciprocal_score_po(base, ent, rel, num_rel, p, o), "po")
        eq(ko.reciprocal_score_sp_po(base, ent, rel, num_rel, s, p, o), "sp_po")
        eq(ko.reciprocal_score_sp_po(base, ent, rel, num_rel, s, p, o, sub), "sp_po_sub")
        with pytest.raises(Exception, match="undirected"):
            ko.reciprocal_score_spo(base, ent, rel, num_rel, s, p, o, None)
        prm = ko.Params(ent, rel)
        opt = ko.make_optimizer("Adagrad", prm, lr=0.2)
        lv = ko.batch_1vsall_reciprocal(base, prm, num_rel, T(aux2[tag + ".b0.triples"]), "kl")
        assert lv == pytest.approx(float(aux2[tag + ".b0.loss"]), rel=1e-6)
        ge, gr = prm.grads()
        np.testing.assert_allclose(ge.numpy(), aux2[tag + ".b0.grad_ent"], rtol=0, atol=1e-7)
        np.testing.assert_allclose(gr.numpy(), aux2[tag + ".b0.grad_rel"], rtol=0, atol=1e-7)
        opt.step()
        np.testing.assert_allclose(prm.ent.detach().numpy(), aux2[tag + ".b0.ent"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(prm.rel.detach().numpy(), aux2[tag + ".b0.rel"], rtol=0, atol=1e-6)


def test_adam_steps_oracle_matches_reference(aux2):
    """util/optimizer.py:10-17 with train.optimizer: Adam."""
    T = torch.from_numpy
    prm = ko.Params(T(aux2["adam.ent0"]), T(aux2["adam.rel0"]))
    opt = ko.make_optimizer("Adam", prm, lr=0.01)
    for step in range(2):
        opt.zero_grad()
        lv = ko.batch_1vsall("distmult", prm, T(aux2[f"adam.b{step}.triples"]), "kl")
        assert lv == pytest.approx(float(aux2[f"adam.b{step}.loss"]), rel=1e-6)
        ge, _ = prm.grads()
        np.testing.assert_allclose(ge.numpy(), aux2[f"adam.b{step}.grad_ent"], rtol=0, atol=1e-7)
        opt.step()
        np.testing.assert_allclose(prm.ent.detach().numpy(), aux2[f"adam.b{step}.ent"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(prm.rel.detach().numpy(), aux2[f"adam.b{step}.rel"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(opt.state[prm.ent]["exp_avg"].numpy(), aux2["adam.exp_avg_ent"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(opt.state[prm.ent]["exp_avg_sq"].numpy(), aux2["adam.exp_avg_sq_ent"], rtol=0, atol=1e-9)


def test_toy_config_steps_with_penalty_oracle_matches_reference(aux2):
    """examples/toy-complex-train.yaml as written (ComplEx, KvsAll + KL, Lp penalty): run_epoch's body (train.py:309-376)."""
    T = torch.from_numpy
    e, r = 53, 7
    prm = ko.Params(T(aux2["toy.ent0"]), T(aux2["toy.rel0"]))
    opt = ko.make_optimizer("Adagrad", prm, lr=0.2)
    assert [str(k) for k in aux2["toy.penalty_keys"]] == ["complex.entity_embedder.L2_penalty",
                                                          "complex.relation_embedder.L2_penalty",
                                                          "complex.entity_embedder.L2_penalty"]
    for step in range(2):
        pre = f"toy.b{step}"
        opt.zero_grad()
        lv, _ = ko.batch_kvsall("complex", prm, aux2[pre + ".queries"], aux2[pre + ".label_coords"], aux2[pre + ".query_type"],
                                e, r, "kl")
        # KvsAll batches carry no "triples": the unweighted penalty of the s-, p- and o-embedder (kge_model.py:598-606)
        vals = []
        for w, lam in ((prm.ent, 1e-2), (prm.rel, 1e-2), (prm.ent, 1e-2)):
            v = ko.lp_penalty(w, 2, lam)
            v.backward()
            vals.append(v.item())
        assert lv == pytest.approx(float(aux2[pre + ".loss"]), rel=1e-6)
        np.testing.assert_allclose(vals, aux2[pre + ".penalties"], rtol=1e-6)
        ge, gr = prm.grads()
        np.testing.assert_allclose(ge.numpy(), aux2[pre + ".grad_ent"], rtol=0, atol=1e-7)
        np.testing.assert_allclose(gr.numpy(), aux2[pre + ".grad_rel"], rtol=0, atol=1e-7)
        opt.step()
        np.testing.assert_allclose(prm.ent.detach().numpy(), aux2[pre + ".ent"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(prm.rel.detach().numpy(), aux2[pre + ".rel"], rtol=0, atol=1e-6)


@pytest.fixture(scope="module")
def aux3():
    return np.load(os.path.join(GOLD, "aux3.npz"), allow_pickle=False)


def test_drilldown_indexes_match_reference(aux3):
    """Host-side restatements in kge-1_b200/index.py against kge/indexing.py:142-263 (bit-exact sets)."""
    from importlib import import_module
    ix = import_module("kge-1_b200.index")
    train, e, r = aux3["graph.train"], 53, 7
    assert ix.relation_types(train, r) == [str(x) for x in aux3["relation_types"]]
    rpt = ix.relations_per_type(train, r)
    assert set(rpt) == {k.split(".", 1)[1] for k in aux3.files if k.startswith("relations_of.")}
    for t, rels in rpt.items():
        assert rels == aux3[f"relations_of.{t}"].tolist()
    fp = ix.frequency_percentiles(train, e, r)
    for arg, percs in fp.items():
        for perc, ids in percs.items():
            assert ids == aux3[f"percentile.{arg}.{perc}"].tolist(), (arg, perc)


def test_argument_frequency_metrics_oracle_matches_reference(aux3):
    """eval.py:173-224 + entity_ranking.py:370-381: per-relation-type and per-frequency-percentile metrics."""
    from importlib import import_module
    ix = import_module("kge-1_b200.index")
    g = {k: aux3[f"graph.{k}"] for k in ("train", "valid", "test")}
    e, r = 53, 7
    ent, rel = torch.from_numpy(aux3["ent"]), torch.from_numpy(aux3["rel"])
    _, ranks = ko.entity_ranking("distmult", ent, rel, g["valid"], [g["train"], g["valid"]], g["test"], batch_size=16,
                                 hits_at_k=(1, 3, 10))
    rpt, fp = ix.relations_per_type(g["train"], r), ix.frequency_percentiles(g["train"], e, r)
    vt = torch.from_numpy(g["valid"].astype(np.int64))
    got = {}
    for n, sfx in {"_raw": "", "_filt": "_filtered", "_filt_test": "_filtered_with_test"}.items():
        rk = {"s": ranks["s" + n], "o": ranks["o" + n]}
        hs = ko.grouped_rank_histograms(rk, vt[:, 1], e, rpt, False)
        hs.update(ko.frequency_histograms(rk, vt, e, fp))
        for key, h in hs.items():
            for k, v in ko.metrics_from_hist(h, (1, 3, 10)).items():
                got[k + sfx + ("" if key == "all" else "_" + key)] = v
    keys = [str(k) for k in aux3["keys"]]
    assert len(keys) > 150
    for k in keys:
        assert k in got, k
        assert abs(got[k] - float(aux3[f"value.{k}"])) <= 1e-6, (k, got[k], float(aux3[f"value.{k}"]))
