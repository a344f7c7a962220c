"""CUDA-graph feature for inference wrappers (reference ``model_implementations/features/cuda_graph.py``).

``CUDAGraph`` is the reference's abstract protocol.  ``GraphedCallable`` is the implementation the wrappers here share:
it keeps one captured graph per *input signature* (tensor shapes / dtypes + non-tensor argument values), so a pipeline
that alternates batch sizes (classifier-free guidance on/off, two CLIP calls per step) replays instead of re-capturing,
and it degrades to eager execution when there is no CUDA device.
"""
from abc import ABC, abstractmethod

import torch


class CUDAGraph(ABC):

    def __init__(self, enable_cuda_graph=False):
        super().__init__()
        self.enable_cuda_graph = enable_cuda_graph

    @abstractmethod
    def _create_cuda_graph(self):
        raise NotImplementedError

    @abstractmethod
    def _graph_replay(self):
        raise NotImplementedError


def _sig(x):
    if torch.is_tensor(x):
        return ("T", tuple(x.shape), x.dtype, x.device.type)
    if isinstance(x, (list, tuple)):
        return (type(x).__name__, tuple(_sig(v) for v in x))
    if isinstance(x, dict):
        return ("D", tuple((k, _sig(v)) for k, v in sorted(x.items())))
    try:
        hash(x)
        return ("V", x)
    except TypeError:
        return ("O", id(x))


def _copy_into(dst, src):
    if torch.is_tensor(dst):
        dst.copy_(src)
    elif isinstance(dst, (list, tuple)):
        for d, s in zip(dst, src):
            _copy_into(d, s)
    elif isinstance(dst, dict):
        for k in dst:
            _copy_into(dst[k], src[k])


def _clone_static(x):
    if torch.is_tensor(x):
        return x.clone()
    if isinstance(x, (list, tuple)):
        return type(x)(_clone_static(v) for v in x)
    if isinstance(x, dict):
        return {k: _clone_static(v) for k, v in x.items()}
    return x


class _Entry:
    __slots__ = ("graph", "args", "kwargs", "output")


class GraphedCallable:
    """``fn`` replayed from CUDA graphs keyed by input signature.  Outputs are the graph's static buffers: consume (or
    clone) them before the next call with the same signature."""

    def __init__(self, fn, enabled=True, warmup=3, max_graphs=8):
        self.fn, self.warmup, self.max_graphs = fn, warmup, max_graphs
        self.enabled = bool(enabled) and torch.cuda.is_available()
        self._entries = {}
        self.captures = 0
        self.replays = 0

    def _capture(self, args, kwargs):
        e = _Entry()
        e.args, e.kwargs = _clone_static(args), _clone_static(kwargs)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(self.warmup):  # allocate workspaces / pick algorithms outside the capture
                self.fn(*e.args, **e.kwargs)
        torch.cuda.current_stream().wait_stream(side)
        e.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(e.graph), torch.no_grad():
            e.output = self.fn(*e.args, **e.kwargs)
        self.captures += 1
        return e

    def __call__(self, *args, **kwargs):
        if not self.enabled:
            return self.fn(*args, **kwargs)
        key = (_sig(args), _sig(kwargs))
        e = self._entries.get(key)
        if e is None:
            if len(self._entries) >= self.max_graphs:
                self._entries.pop(next(iter(self._entries)))  # oldest signature
            e = self._entries[key] = self._capture(args, kwargs)
        _copy_into(e.args, args)
        _copy_into(e.kwargs, kwargs)
        e.graph.replay()
        self.replays += 1
        return e.output

    def reset(self):
        self._entries.clear()
