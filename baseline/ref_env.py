"""Imports the UNMODIFIED reference from the git-ignored baseline/_ref/ (installed by baseline/install_ref.sh; it travels
to the GPU box with the gpurun snapshot).  The three import-time shims of SURVEY.md Appendix B let it load under Python
3.12 / torch 2.11; no source file of the reference is edited.

Used by `bench.py` (the CPU reference arm and the e2e leg through the reference's own job objects) and by
tests/test_gpu_dropin.py.  Nothing here touches /root/reference."""
import collections
import collections.abc
import os
import sys
import types

REF_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "kge", "job"))


def install(root: str = REF_ROOT):
    """Idempotent.  Returns the imported `kge` package."""
    if not os.path.isdir(os.path.join(root, "kge", "job")):
        raise ImportError(f"no reference install under {root} (run baseline/install_ref.sh in the build container)")
    collections.Mapping = collections.abc.Mapping                      # kge/config.py:56,59 (removed in Python 3.10)
    if "path" not in sys.modules:                                       # kge/misc.py:5,22,38 (package `path` absent)
        mod = types.ModuleType("path")

        class Path(str):
            def __enter__(self):
                self._old = os.getcwd()
                os.chdir(self)
                return self

            def __exit__(self, *a):
                os.chdir(self._old)

        mod.Path = Path
        sys.modules["path"] = mod
    for name, attrs in {"ax": ["Models"], "ax.core": ["ObservationFeatures"], "ax.modelbridge": [],
                        "ax.modelbridge.generation_strategy": ["GenerationStep", "GenerationStrategy"],
                        "ax.service": [], "ax.service.ax_client": ["AxClient"]}.items():   # kge/job/ax_search.py:3-9
        if name not in sys.modules:
            m = types.ModuleType(name)
            for a in attrs:
                setattr(m, a, object)
            sys.modules[name] = m
    sys.dont_write_bytecode = True
    if root not in sys.path:
        sys.path.insert(0, root)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import kge
        import kge.job     # noqa: F401
        import kge.model   # noqa: F401
    return kge


def _quiet(*args, **kwargs):
    """Config.log / Config.print replacement (module level: the config object is pickled into checkpoints)."""


def _keep_epoch_entry(job, entry):
    job.epoch_entries.append(dict(entry))


def make_job(model: str, graph: dict, num_entities: int, num_relations: int, dim: int, options: dict, device: str = "cpu",
             plugin: bool = False, plugin_args: dict = None, seed: int = 0, folder: str = None):
    """Builds a job of the reference through its own factories (Config -> Dataset -> Job.create) on an in-memory graph
    {"train" | "valid" | "test": int32 [N,3]} (SURVEY.md Appendix B recipe).  plugin=True selects the B200 plug-in model of
    `model` (kge-1_b200/libkge_plugin.py)."""
    import tempfile
    import torch
    install()
    from kge import Config, Dataset
    from kge.job import Job
    folder = folder or tempfile.mkdtemp(prefix="kgeb200_ref_")
    c = Config(folder=folder)
    c.set("job.device", device)
    for k, v in options.items():
        if k.startswith("train.optimizer"):
            c.set(k, v, create=True)
    if plugin:
        import kgeb200
        kgeb200.libkge_plugin.register()
        kgeb200.libkge_plugin.configure(c, model, **(plugin_args or {}))
    else:
        c.set("model", model)
        c._import(model)
    c.set("dataset.num_entities", num_entities)
    c.set("dataset.num_relations", num_relations)
    c.set("dataset.pickle", False)
    c.set("lookup_embedder.dim", dim)
    for k, v in options.items():
        if not k.startswith("train.optimizer"):
            c.set(k, v, create=True)
    ds = Dataset(c, folder=folder)
    for split, arr in graph.items():
        ds._triples[split] = torch.from_numpy(arr.copy())
    torch.manual_seed(seed)
    if c.get("job.type") == "eval":      # an evaluation job is handed its model (job.py:86-89)
        from kge.model import KgeModel
        job = Job.create(c, ds, model=KgeModel.create(c, ds))
    else:
        job = Job.create(c, ds)
        job.epoch_entries = []
        job.post_epoch_hooks.append(_keep_epoch_entry)
    c.log = _quiet
    c.print = _quiet
    return job
