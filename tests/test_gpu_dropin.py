"""The drop-in, end to end: the reference's OWN job code (baseline/_ref, unmodified: Job.create -> TrainingJob*.run() with
its embedded EntityRankingJob) runs once with the reference's model on the CPU and once with the B200 plug-in model on
cuda, from the same seed.  Losses, parameters and ranking metrics must agree.

baseline/_ref is installed by baseline/install_ref.sh in the build container and travels with the gpurun snapshot;
/root/reference is never read here."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import ref_env  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_env.available(), reason="baseline/_ref is not installed")]


@pytest.fixture(scope="module")
def toy():
    import kgeb200 as kb
    g = kb.graph.synthetic_graph("toy", seed=1)
    graph = {k: g[k].astype(np.int32) for k in ("train", "valid", "test")}
    return graph, g["num_entities"], g["num_relations"]


def _run_pair(toy, model, options, plugin_args=None, dim=32):
    graph, e, r = toy
    jobs = []
    for plugin in (False, True):
        job = ref_env.make_job(model, graph, e, r, dim, options, device="cuda" if plugin else "cpu", plugin=plugin,
                               plugin_args=plugin_args, seed=0)
        torch.manual_seed(1)          # DataLoader shuffling
        job.run()
        jobs.append(job)
    return jobs


def _epoch_trace(job):
    return job.epoch_entries[-1]          # run_epoch's trace entry (train.py:436-460), kept by a post_epoch hook


# a non-zero initial accumulator: with sum = 0 the first Adagrad steps are lr * g / (|g| + 1e-10), which turn rounding-level
# differences of the two arithmetic paths into O(lr) differences wherever g ~ 0 (and the runs then drift apart)
BASE = {"train.max_epochs": 1, "valid.every": 1, "train.optimizer": "Adagrad", "train.optimizer_args": {"lr": 0.1, "initial_accumulator_value": 0.1},
        "eval.batch_size": 64, "train.batch_size": 512, "train.num_workers": 0, "eval.num_workers": 0}


def _compare(ref, new, loss_rtol, w_atol, mrr_atol=1e-3):
    a, c = _epoch_trace(ref), _epoch_trace(new)
    assert c["avg_loss"] == pytest.approx(a["avg_loss"], rel=loss_rtol), ("epoch avg_loss", a["avg_loss"], c["avg_loss"])
    assert c["avg_penalty"] == pytest.approx(a["avg_penalty"], rel=1e-4, abs=1e-12)
    for (ka, wa), (kc, wc) in zip(ref.model.state_dict().items(), new.model.state_dict().items()):
        assert ka == kc
        assert (wa - wc.cpu()).abs().max().item() <= w_atol, (ka, (wa - wc.cpu()).abs().max().item())
    va, vc = ref.valid_trace[-1], new.valid_trace[-1]
    for k in ("mean_reciprocal_rank", "mean_reciprocal_rank_filtered", "hits_at_1_filtered", "hits_at_10_filtered"):
        assert abs(va[k] - vc[k]) <= mrr_atol, (k, va[k], vc[k])
    assert abs(va["mean_rank_filtered"] - vc["mean_rank_filtered"]) <= 0.5
    assert set(va.keys()) == set(vc.keys())          # same metric names / trace fields


def test_reference_kvsall_job_with_fused_loss_matches_cpu_reference(toy):
    """examples/toy-complex-train.yaml's setting (ComplEx, KvsAll + KL, Lp penalty) through the reference's run():
    _process_batch as written (train.py:679-756) with LazyScores + FusedKgeLoss, B200Adagrad, fused evaluation."""
    opts = dict(BASE, **{"train.type": "KvsAll", "train.loss": "kl", "lookup_embedder.regularize_weight": 1e-3,
                         "entity_ranking.metrics_per.relation_type": True})
    ref, new = _run_pair(toy, "complex", opts)
    import kge.job as kj
    from kgeb200 import libkge_jobs
    assert type(new.valid_job) is kj.EntityRankingJob and new.valid_job._b200          # the rebound, fused evaluation job
    assert not getattr(ref.valid_job, "_b200", False)
    assert isinstance(new.loss, libkge_jobs.FusedKgeLoss) and type(new.optimizer).__name__ == "Adagrad"
    assert type(new.optimizer).__module__.endswith("optim") and not isinstance(ref.loss, libkge_jobs.FusedKgeLoss)
    _compare(ref, new, loss_rtol=2e-5, w_atol=0.1 * 2e-3)


def test_reference_1vsall_job_with_fused_loss_matches_cpu_reference(toy):
    opts = dict(BASE, **{"train.type": "1vsAll", "train.loss": "kl"})
    ref, new = _run_pair(toy, "distmult", opts)
    _compare(ref, new, loss_rtol=2e-5, w_atol=0.1 * 2e-3)


def test_reference_kvsall_job_with_bce_and_label_smoothing(toy):
    opts = dict(BASE, **{"train.type": "KvsAll", "train.loss": "bce", "KvsAll.label_smoothing": 0.1})
    ref, new = _run_pair(toy, "simple", opts)
    _compare(ref, new, loss_rtol=2e-5, w_atol=0.1 * 2e-3)


@pytest.mark.parametrize("train_type,model,loss", [("KvsAll", "complex", "kl"), ("1vsAll", "distmult", "kl"),
                                                   ("KvsAll", "distmult", "bce")])
def test_reference_job_with_captured_step_matches_cpu_reference(toy, train_type, model, loss):
    """b200.captured_step: job._process_batch replaced ON THE JOB OBJECT by the graph-captured stepper (fp32 math); the
    last, smaller batch of the epoch takes the reference's own body with the fused loss.  Penalty folded into Adagrad."""
    opts = dict(BASE, **{"train.type": train_type, "train.loss": loss, "lookup_embedder.regularize_weight": 1e-3})
    ref, new = _run_pair(toy, model, opts, plugin_args={"captured_step": True})
    assert getattr(new.model, "_b200_stepper", None) is not None, "the captured step never ran"
    _compare(ref, new, loss_rtol=5e-5, w_atol=0.1 * 2e-3)


def test_captured_step_bf16_tiles_train_the_reference_job(toy):
    """The fast arithmetic (bf16 tcgen05 tiles) behind the same reference job: looser, stated bounds."""
    opts = dict(BASE, **{"train.type": "KvsAll", "train.loss": "bce"})
    ref, new = _run_pair(toy, "complex", opts, plugin_args={"captured_step": True, "math": "bf16"}, dim=64)
    a, c = _epoch_trace(ref), _epoch_trace(new)
    assert c["avg_loss"] == pytest.approx(a["avg_loss"], rel=1e-2)
    assert abs(ref.valid_trace[-1]["mean_reciprocal_rank_filtered"] - new.valid_trace[-1]["mean_reciprocal_rank_filtered"]) <= 5e-3


def test_reference_negative_sampling_job_runs_on_plugin_model(toy):
    """TrainingJobNegativeSampling as written (train.py:823-999): score_spo is the fused gather+score kernel."""
    opts = dict(BASE, **{"train.type": "negative_sampling", "train.loss": "kl", "negative_sampling.num_samples.s": 8,
                         "negative_sampling.num_samples.o": 8, "train.batch_size": 256})
    ref, new = _run_pair(toy, "distmult", opts)
    # the CPU sampler draws from the global torch RNG in the collate: both runs see the same negatives
    _compare(ref, new, loss_rtol=5e-5, w_atol=0.1 * 2e-3)


def test_standalone_entity_ranking_job_matches_reference_metrics(toy):
    """EvaluationJob.create -> the rebound EntityRankingJob on identical weights: every metric of the reference's trace
    entry (incl. head/tail and relation-type drill-downs, filter_with_test) within 1e-6 / ranks equal."""
    graph, e, r = toy
    opts = {"job.type": "eval", "eval.split": "valid", "eval.batch_size": 50, "eval.num_workers": 0,
            "entity_ranking.metrics_per.head_and_tail": True, "entity_ranking.metrics_per.relation_type": True,
            "entity_ranking.metrics_per.argument_frequency": True, "entity_ranking.filter_with_test": True,
            "entity_ranking.tie_handling": "rounded_mean_rank"}
    for model in ("complex", "transe", "rotate"):
        ref = ref_env.make_job(model, graph, e, r, 32, opts, device="cpu", plugin=False, seed=3)
        new = ref_env.make_job(model, graph, e, r, 32, opts, device="cuda", plugin=True, seed=3)
        new.model.load_state_dict(ref.model.state_dict())
        ta, tc = ref.run(), new.run()
        keys = [k for k in ta if k.startswith(("mean_r", "hits_at"))]
        assert len(keys) > 100 and set(ta.keys()) == set(tc.keys())
        for k in keys:
            assert abs(ta[k] - tc[k]) <= (1e-3 if "mean_rank" not in k else 0.5), (model, k, ta[k], tc[k])
