"""`import jamun` -> jamun_b200: lets the reference's Hydra `_target_` strings and pickled checkpoint
hyper-parameters (`jamun.model.Denoiser`, `jamun.model.arch.E3Conv`, `jamun.e3tools.nn.ConvBlock`, ...) resolve to
the B200 implementation of the walk-jump path.  Modules are aliased, not copied: `jamun.model is jamun_b200.model`.
"""
import importlib
import importlib.abc
import importlib.util
import sys

import jamun_b200 as _impl


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname == "jamun" or not fullname.startswith("jamun."):
            return None
        real = "jamun_b200" + fullname[len("jamun"):]
        try:
            if importlib.util.find_spec(real) is None:
                return None
        except ModuleNotFoundError:
            return None
        return importlib.util.spec_from_loader(fullname, self)

    def create_module(self, spec):
        return importlib.import_module("jamun_b200" + spec.name[len("jamun"):])

    def exec_module(self, module):
        pass


sys.meta_path.insert(0, _AliasFinder())
__version__ = _impl.__version__
__path__ = []  # submodules come from the finder above


def __getattr__(name):
    return getattr(_impl, name)
