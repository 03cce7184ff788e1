"""kge-1_b200: B200-native (sm_100a) scoring/embedding hot path of LibKGE behind the reference's model API.

The directory name follows the build contract; import it as `importlib.import_module("kge-1_b200")` or
through the `kgeb200` alias module at the repository root.
"""
from . import lib  # noqa: F401  (ctypes binding; loading the .so is deferred to first use)
from . import ops, fused, index, graph, optim, jobs, model, trainer, sampler, metrics, libkge_plugin  # noqa: F401
from .model import KgeModel, LookupEmbedder, RelationalScorer, ReciprocalRelationsModel  # noqa: F401
from .jobs import (KgeLoss, TrainingJob1vsAll, TrainingJobKvsAll, TrainingJobNegativeSampling,  # noqa: F401
                   EntityRankingJob)
from .sampler import KgeSampler, KgeUniformSampler  # noqa: F401
