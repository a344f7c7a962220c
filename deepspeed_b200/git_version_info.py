"""Version / build info (reference ``git_version_info.py``): version, git hash + branch when run from a checkout, and the
native-op compatibility table used by ``ds_report``."""
import os
import subprocess

import torch


_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _version():
    import re
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "__init__.py")) as f:
        m = re.search(r'__version__\s*=\s*"([^"]+)"', f.read())
    return m.group(1) if m else "0.0.0"


version = _version()


def _git(*args):
    try:
        return subprocess.check_output(["git", "-C", _root, *args], stderr=subprocess.DEVNULL).decode().strip() or "[none]"
    except Exception:
        return "[none]"


git_hash = _git("rev-parse", "--short", "HEAD")
git_branch = _git("rev-parse", "--abbrev-ref", "HEAD")
accelerator_name = "cuda" if torch.cuda.is_available() else "cpu"
def _nccl_version():
    try:
        return ".".join(str(x) for x in torch.cuda.nccl.version()[:2])
    except Exception:
        return "0.0"


torch_info = {"version": torch.__version__, "bf16_support": True, "cuda_version": torch.version.cuda or "0.0",
              "nccl_version": _nccl_version(), "hip_version": "0.0"}


def _ops():
    try:
        from deepspeed_b200.op_builder import ALL_OPS
        return ALL_OPS
    except Exception:
        return {}


def _built(name, builder):
    try:
        b = builder() if isinstance(builder, type) else builder
        return bool(b.installed())
    except Exception:
        return False


def _compat(builder):
    try:
        b = builder() if isinstance(builder, type) else builder
        return bool(b.is_compatible())
    except Exception:
        return False


installed_ops = {n: _built(n, b) for n, b in _ops().items()}
compatible_ops = {n: _compat(b) for n, b in _ops().items()}
compatible_ops["deepspeed_not_implemented"] = False
