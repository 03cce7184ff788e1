"""The oracle (oracle/kge_oracle.py) against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py) and against the reference's tie-handling known-answer vector
(kge/config-default.yaml:392-406).  CPU only."""
import ast

import numpy as np
import pytest
import torch

from oracle import kge_oracle as ko

T = torch.from_numpy
TAGS = ["distmult", "complex", "cp", "simple", "rescal", "transe_l1", "transe_l2", "rotate_l1", "rotate_l2"]


def _model(tag):
    m = tag.split("_l")[0]
    ln = float(tag.split("_l")[1]) if "_l" in tag else 1.0
    return m, ln


@pytest.mark.parametrize("tag", TAGS)
def test_scores_match_reference(golden, tag):
    g = golden("scores")
    m, ln = _model(tag)
    ent, rel = T(g[f"{tag}.ent"]), T(g[f"{tag}.rel"])
    s, p, o, sub = (T(g[k]) for k in ("idx_s", "idx_p", "idx_o", "subset"))
    assert rel.shape[1] == ko.relation_dim(m, ent.shape[1])
    got = {
        "spo": ko.score_spo(m, ent, rel, s, p, o, ln),
        "sp": ko.score_sp(m, ent, rel, s, p, l_norm=ln),
        "po": ko.score_po(m, ent, rel, p, o, l_norm=ln),
        "sp_sub": ko.score_sp(m, ent, rel, s, p, sub, l_norm=ln),
        "sp_po_sub": ko.score_sp_po(m, ent, rel, s, p, o, sub, l_norm=ln),
        "so": ko.score_so(m, ent, rel, s, o, l_norm=ln),
        "spo_i32": ko.score_spo(m, ent, rel, s.int(), p.int(), o.int(), ln),
    }
    for k, v in got.items():
        np.testing.assert_array_equal(v.numpy(), g[f"{tag}.{k}"], err_msg=f"{tag}.{k}")


def test_losses_match_reference(golden):
    g = golden("losses")
    sc, idx, mat = T(g["loss.scores"]), T(g["loss.idx"]), T(g["loss.mat"])
    assert ko.loss_kl(sc, idx).item() == g["loss.kl.idx"]
    assert ko.loss_kl(sc, mat).item() == g["loss.kl.mat"]
    assert ko.loss_bce(sc, idx).item() == g["loss.bce.idx"]
    assert ko.loss_bce(sc, mat).item() == g["loss.bce.mat"]
    assert ko.loss_bce(sc, idx, 0.7).item() == g["loss.bce_off.idx"]
    assert ko.loss_bce(sc, mat, 0.7).item() == g["loss.bce_off.mat"]


def test_kvsall_index_and_coords_bit_exact(golden):
    g = golden("index")
    for split in ("train", "valid"):
        for key in ("sp", "po"):
            keys, off, vals = ko.kvsall_index(g[f"index.{split}.triples"], key)
            np.testing.assert_array_equal(keys, g[f"index.{split}.{key}.keys"])
            np.testing.assert_array_equal(off, g[f"index.{split}.{key}.offsets"])
            np.testing.assert_array_equal(vals, g[f"index.{split}.{key}.values"])
    tr = g["index.train.triples"]
    coords = ko.sp_po_coords(g["index.coords_batch"], 53, ko.CsrIndex(tr, "sp"), ko.CsrIndex(tr, "po"))
    np.testing.assert_array_equal(coords, g["index.coords"])


def _train_cases(golden):
    return [str(x) for x in golden("train")["train.cases"]]


def test_train_steps_match_reference(golden):
    g = golden("train")
    graph = g["train.graph.train"]
    e, r = 53, 7
    for tag in _train_cases(golden):
        _, ttype, model, loss = tag.split(".")
        opts = dict(ast.literal_eval(str(g[tag + ".options"]))) if tag + ".options" in g else {}
        ln = float(opts.get(model + ".l_norm", 1.0))
        off = float(opts.get("train.loss_arg", 0.0))
        ls = float(opts.get("KvsAll.label_smoothing", 0.0))
        prm = ko.Params(T(g[tag + ".ent0"]), T(g[tag + ".rel0"]))
        opt = ko.make_optimizer("Adagrad", prm, lr=0.2)
        for step in range(2):
            pre = f"{tag}.b{step}"
            opt.zero_grad()
            if ttype == "1vsAll":
                lv = ko.batch_1vsall(model, prm, T(g[pre + ".triples"]), loss, off, ln)
            elif ttype == "KvsAll":
                lv, _ = ko.batch_kvsall(model, prm, g[pre + ".queries"], g[pre + ".label_coords"],
                                        g[pre + ".query_type"], e, r, loss, off, ls, ln)
            else:
                negs = {slot: T(g[f"{pre}.neg{slot}"]) for slot in range(3)}
                lv = ko.batch_negative_sampling(model, prm, T(g[pre + ".triples"]), negs, loss, off, ln)
            ge, gr = prm.grads()
            assert lv == pytest.approx(float(g[pre + ".loss"]), rel=1e-6), tag
            np.testing.assert_allclose(ge.numpy(), g[pre + ".grad_ent"], rtol=0, atol=1e-7, err_msg=tag)
            np.testing.assert_allclose(gr.numpy(), g[pre + ".grad_rel"], rtol=0, atol=1e-7, err_msg=tag)
            opt.step()
            np.testing.assert_allclose(prm.ent.detach().numpy(), g[pre + ".ent"], rtol=0, atol=1e-6, err_msg=tag)
            np.testing.assert_allclose(prm.rel.detach().numpy(), g[pre + ".rel"], rtol=0, atol=1e-6, err_msg=tag)
    # the KvsAll collate restatement reproduces the reference's batches from example ids
    tr = graph
    idx = [ko.kvsall_index(tr, "sp"), ko.kvsall_index(tr, "po")]
    q, coords, qt = ko.kvsall_collate(list(range(0, len(idx[0][0]) + 5, 3)), idx)
    assert q.shape[1] == 2 and coords[:, 0].max() == len(q) - 1 and set(qt.tolist()) == {0, 1}


def test_entity_ranking_matches_reference(golden):
    g = golden("eval")
    graph = {k: g[f"eval.graph.{k}"] for k in ("train", "valid", "test")}
    for tag in [str(x) for x in g["eval.cases"]]:
        _, model, chunk, ties = tag.split(".")
        ent, rel = T(g[tag + ".ent"]), T(g[tag + ".rel"])
        metrics, ranks = ko.entity_ranking(
            model, ent, rel, graph["valid"], [graph["train"], graph["valid"]], graph["test"],
            batch_size=16, chunk_size=int(chunk[1:]), tie_handling=ties, hits_at_k=(1, 3, 10, 50))
        for nm in ("raw", "filt", "filt_test"):
            for d in "so":
                np.testing.assert_array_equal(ranks[f"{d}_{nm}"].numpy(), g[f"{tag}.{d}_ranks_{nm}"],
                                              err_msg=f"{tag} {d}_{nm}")
        assert metrics["_filt"]["mean_reciprocal_rank"] == pytest.approx(
            float(g[tag + ".metric.mean_reciprocal_rank_filtered"]), abs=1e-6)
        assert metrics["_raw"]["mean_reciprocal_rank"] == pytest.approx(
            float(g[tag + ".metric.mean_reciprocal_rank"]), abs=1e-6)
        assert metrics["_filt"]["hits_at_10"] == pytest.approx(
            float(g[tag + ".metric.hits_at_10_filtered"]), abs=1e-6)
        assert metrics["_filt_test"]["mean_reciprocal_rank"] == pytest.approx(
            float(g[tag + ".metric.mean_reciprocal_rank_filtered_with_test"]), abs=1e-6)


def test_tie_handling_known_answer(golden):
    """config-default.yaml:392-406: scores a:10 b:10 c:10 d:11 e:9, true a ->
    worst 4, best 2, rounded mean 3 (1-based)."""
    sc = torch.tensor([[10.0, 10.0, 10.0, 11.0, 9.0]])
    rank, ties = ko.ranks_and_ties(sc, torch.tensor([10.0]))
    np.testing.assert_array_equal([rank.item(), ties.item()], golden("eval")["eval.tie_example.rank_ties"])
    assert ko.final_ranks(rank, ties, "worst_rank").item() + 1 == 4
    assert ko.final_ranks(rank, ties, "best_rank").item() + 1 == 2
    assert ko.final_ranks(rank, ties, "rounded_mean_rank").item() + 1 == 3
