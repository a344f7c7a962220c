# alias `deepspeed` -> `deepspeed_b200` for running the reference's unit tests against this framework
import importlib, importlib.abc, importlib.util, sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

class _Alias(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path=None, target=None):
        if name == "deepspeed" or name.startswith("deepspeed."):
            real = "deepspeed_b200" + name[len("deepspeed"):]
            try:
                if importlib.util.find_spec(real) is None:
                    return None
            except (ImportError, ValueError):
                return None
            return importlib.util.spec_from_loader(name, self, is_package=True)
        return None
    def create_module(self, spec):
        real = "deepspeed_b200" + spec.name[len("deepspeed"):]
        return importlib.import_module(real)
    def exec_module(self, module):
        pass

sys.meta_path.insert(0, _Alias())
