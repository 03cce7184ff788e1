"""Import shims that let the UNMODIFIED reference (/root/reference, LibKGE fork) be imported
under Python 3.12 / torch 2.11 in the build container.

Only used by tests/golden/make_golden.py to generate the committed golden vectors.  The
reference tree does not exist on the GPU box, so nothing under tests/ imports this at test time.
"""
import collections
import collections.abc
import os
import sys
import types


def install(reference_root="/root/reference"):
    collections.Mapping = collections.abc.Mapping  # kge/config.py:56,59 (removed in py3.10)

    path_mod = types.ModuleType("path")  # kge/misc.py:5 ("path" package absent)

    class Path(str):
        def __enter__(self):
            self._old = os.getcwd()
            os.chdir(self)
            return self

        def __exit__(self, *a):
            os.chdir(self._old)

    path_mod.Path = Path
    sys.modules["path"] = path_mod

    stubs = {  # kge/job/ax_search.py:3-9 (ax-platform absent)
        "ax": ["Models"],
        "ax.core": ["ObservationFeatures"],
        "ax.modelbridge": [],
        "ax.modelbridge.generation_strategy": ["GenerationStep", "GenerationStrategy"],
        "ax.service": [],
        "ax.service.ax_client": ["AxClient"],
    }
    for name, attrs in stubs.items():
        m = types.ModuleType(name)
        for a in attrs:
            setattr(m, a, object)
        sys.modules[name] = m
    sys.dont_write_bytecode = True
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
