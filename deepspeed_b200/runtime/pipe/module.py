"""``PipelineModule``: a model expressed as a list of layers partitioned over pipeline stages.

Parity target: reference ``runtime/pipe/module.py`` (``LayerSpec :30``, ``TiedLayerSpec :77``,
``PipelineModule :86``, ``_partition_layers :393``, per-layer checkpoint files ``:571-620``).
"""
import collections
import glob
import os
import re as regex
from functools import partial

import torch
from torch import nn

from deepspeed_b200 import comm as dist
from deepspeed_b200.runtime.utils import partition_balanced, partition_uniform
from deepspeed_b200.utils.logging import logger
from .topology import PipeDataParallelTopology, PipelineParallelGrid


class PipelineError(Exception):
    pass


class LayerSpec:
    """Deferred layer construction: only the owning stage instantiates the module."""

    def __init__(self, typename, *module_args, **module_kwargs):
        self.typename = typename
        self.module_args = module_args
        self.module_kwargs = module_kwargs
        if not issubclass(typename, nn.Module):
            raise RuntimeError("LayerSpec only supports torch.nn.Module types.")
        self.global_rank = dist.get_rank() if dist.is_initialized() else -1

    def __repr__(self):
        from deepspeed_b200.runtime.utils import call_to_str
        return call_to_str(self.typename.__name__, *self.module_args, **self.module_kwargs)

    def build(self, log=False):
        if log:
            logger.info(f"RANK={self.global_rank} building {repr(self)}")
        return self.typename(*self.module_args, **self.module_kwargs)


class TiedLayerSpec(LayerSpec):

    def __init__(self, key, typename, *module_args, forward_fn=None, tied_weight_attr=("weight", ), **module_kwargs):
        super().__init__(typename, *module_args, **module_kwargs)
        self.key = key
        self.forward_fn = forward_fn
        self.tied_weight_attr = [tied_weight_attr] if isinstance(tied_weight_attr, str) else list(tied_weight_attr)


class PipelineModule(nn.Module):

    def __init__(self, layers, num_stages=None, topology=None, loss_fn=None, seed_layers=False, seed_fn=None,
                 base_seed=1234, partition_method="parameters", activation_checkpoint_interval=0,
                 activation_checkpoint_func=None, checkpointable_layers=None, dynamic_shape=False):
        super().__init__()
        if num_stages is None and topology is None:
            raise RuntimeError("must provide num_stages or topology")
        self.micro_offset = 0
        self.loss_fn = loss_fn
        self.checkpointable_layers = checkpointable_layers
        self.seed_layers, self.seed_fn, self.base_seed = seed_layers, seed_fn, base_seed
        self.dynamic_shape = dynamic_shape
        if not dist.is_initialized():
            dist.init_distributed()
        self.world_group = None
        self.global_rank = dist.get_rank()
        self.world_size = dist.get_world_size()
        self.local_rank = int(os.environ.get("LOCAL_RANK", 0))
        if topology:
            self._topo = topology
            self.num_stages = self._topo.get_dim("pipe")
        else:
            self.num_stages = num_stages
            if self.world_size % self.num_stages != 0:
                raise RuntimeError(f"num_stages ({self.num_stages}) must divide distributed world size ({self.world_size})")
            self._topo = PipeDataParallelTopology(self.num_stages, self.world_size // self.num_stages)
        self._grid = PipelineParallelGrid(topology=self._topo)
        self.stage_id = self._topo.get_coord(self.global_rank).pipe
        self._layer_specs = list(layers)
        self._num_layers = len(self._layer_specs)
        self._local_start = self._local_stop = 0
        self._partition_layers(method=partition_method)
        self.forward_funcs = []
        self.fwd_map = {}
        self.tied_modules = nn.ModuleDict()
        self.tied_weight_attrs = {}
        self._build()
        from deepspeed_b200.accelerator import get_accelerator
        self.to(get_accelerator().current_device_name())
        self.tied_comms = self._index_tied_modules()
        self._synchronize_tied_weights()
        self.activation_checkpoint_interval = activation_checkpoint_interval
        if activation_checkpoint_func is None:
            from deepspeed_b200.runtime.activation_checkpointing import checkpointing
            activation_checkpoint_func = checkpointing.non_reentrant_checkpoint
        self.activation_checkpoint_func = activation_checkpoint_func

    # ---- construction ---------------------------------------------------------------------------------
    def _build(self):
        specs = self._layer_specs
        for local_idx, layer in enumerate(specs[self._local_start:self._local_stop]):
            layer_idx = local_idx + self._local_start
            if self.seed_layers:
                (self.seed_fn or torch.manual_seed)(self.base_seed + layer_idx)
            if isinstance(layer, PipelineModule):
                raise NotImplementedError("RECURSIVE BUILD NOT YET IMPLEMENTED")
            elif isinstance(layer, nn.Module):
                name = str(layer_idx)
                self.forward_funcs.append(layer)
                self.fwd_map[name] = len(self.forward_funcs) - 1
                self.add_module(name, layer)
            elif isinstance(layer, TiedLayerSpec):
                if layer.key not in self.tied_modules:
                    self.tied_modules[layer.key] = layer.build()
                    self.tied_weight_attrs[layer.key] = layer.tied_weight_attr
                if layer.forward_fn is None:
                    self.forward_funcs.append(self.tied_modules[layer.key])
                else:
                    self.forward_funcs.append(partial(layer.forward_fn, self.tied_modules[layer.key]))
            elif isinstance(layer, LayerSpec):
                module = layer.build()
                name = str(layer_idx)
                self.forward_funcs.append(module)
                self.fwd_map[name] = len(self.forward_funcs) - 1
                self.add_module(name, module)
            else:
                self.forward_funcs.append(layer)  # plain callable
        for p in self.parameters():
            p.ds_pipe_replicated = False

    def _count_layer_params(self):
        counts = [0] * len(self._layer_specs)
        for idx, layer in enumerate(self._layer_specs):
            if isinstance(layer, LayerSpec):
                m = layer.build()
                counts[idx] = sum(p.numel() for p in m.parameters() if p.requires_grad)
            elif isinstance(layer, nn.Module):
                counts[idx] = sum(p.numel() for p in layer.parameters() if p.requires_grad)
        return counts

    def _find_layer_type(self, layername):
        idxs = []
        rx = regex.compile(layername, regex.IGNORECASE)
        for idx, layer in enumerate(self._layer_specs):
            name = layer.typename.__name__ if isinstance(layer, LayerSpec) else (
                layer.__class__.__name__ if isinstance(layer, nn.Module) else getattr(layer, "__name__", ""))
            if rx.search(name):
                idxs.append(idx)
        if not idxs:
            raise RuntimeError(f"Partitioning '{layername}' found no valid layers to partition.")
        return idxs

    def _partition_layers(self, method="uniform"):
        num_stages = self._topo.get_dim("pipe")
        stage_id = self._topo.get_coord(self.global_rank).pipe
        method = method.lower()
        if method == "uniform":
            self.parts = partition_uniform(len(self._layer_specs), num_stages)
        elif method in ("parameters", "best"):
            self.parts = partition_balanced(self._count_layer_params(), num_stages)
        elif method.startswith("type:"):
            hits = self._find_layer_type(method.split(":", 1)[1])
            w = [0] * len(self._layer_specs)
            for i in hits:
                w[i] = 1
            self.parts = partition_balanced(w, num_stages)
        elif method == "profile":
            raise NotImplementedError(f"Partitioning method {method} not implemented.")
        else:
            raise NotImplementedError(f"Partitioning method {method} not implemented.")
        if self.global_rank == 0:
            for s in range(num_stages):
                logger.info(f"stage={s} layers={self.parts[s + 1] - self.parts[s]} [{self.parts[s]}, {self.parts[s + 1]})")
        self._set_bounds(self.parts[stage_id], self.parts[stage_id + 1])

    def _set_bounds(self, start=None, stop=None):
        self._local_start, self._local_stop = start, stop

    # ---- forward -------------------------------------------------------------------------------------------
    def _is_checkpointable(self, funcs):
        if self.checkpointable_layers is not None:
            return all(f.__class__.__name__ in self.checkpointable_layers for f in funcs)
        params = [f.parameters() for f in funcs if isinstance(f, nn.Module)]
        return any(len(list(p)) > 0 for p in params)

    def forward(self, forward_input):
        self.micro_offset += 1

        def exec_range(start, end):

            def run(*inputs):
                x = inputs[0] if len(inputs) == 1 else inputs
                for idx, layer in enumerate(self.forward_funcs[start:end]):
                    x = layer(x)
                return x

            return run

        if self.activation_checkpoint_interval == 0:
            return exec_range(0, len(self.forward_funcs))(forward_input)
        x = forward_input
        n = len(self.forward_funcs)
        for start in range(0, n, self.activation_checkpoint_interval):
            end = min(start + self.activation_checkpoint_interval, n)
            funcs = self.forward_funcs[start:end]
            if not isinstance(x, tuple):
                x = (x, )
            if self._is_checkpointable(funcs):
                x = self.activation_checkpoint_func(exec_range(start, end), *x)
            else:
                x = exec_range(start, end)(*x)
        return x

    # ---- tied weights -------------------------------------------------------------------------------------
    def _index_tied_modules(self):
        tied = {}
        if self._topo.get_dim("pipe") == 1:
            return tied
        keys = sorted({s.key for s in self._layer_specs if isinstance(s, TiedLayerSpec)})
        for key in keys:
            stages = sorted({self.stage_owner(i) for i, s in enumerate(self._layer_specs)
                             if isinstance(s, TiedLayerSpec) and s.key == key})
            for dp in range(self._grid.data_parallel_size):
                for mp in range(self._grid.get_slice_parallel_world_size()):
                    ranks = []
                    for s in stages:
                        kw = {"data": dp}
                        if "model" in self._topo.get_axis_names():
                            kw["model"] = mp
                        ranks.append(self._grid.stage_to_global(stage_id=s, **kw))
                    ranks = sorted(set(ranks))
                    group = dist.new_group(ranks=ranks)
                    if self.global_rank in ranks and key in self.tied_modules:
                        tied[key] = {"ranks": ranks, "group": group, "weight_attr": self.tied_weight_attrs[key],
                                     "module": self.tied_modules[key]}
                        if self.global_rank != ranks[0]:
                            for p in self.tied_modules[key].parameters():
                                p.ds_pipe_replicated = True
        return tied

    @staticmethod
    def _get_attr(module, name):
        obj = module
        for part in name.split("."):
            obj = getattr(obj, part)
        return obj

    def _synchronize_tied_weights(self):
        for key, comm in self.tied_comms.items():
            for attr in comm["weight_attr"]:
                dist.broadcast(self._get_attr(comm["module"], attr).data, src=min(comm["ranks"]), group=comm["group"])

    def allreduce_tied_weight_gradients(self):
        for key, comm in self.tied_comms.items():
            for attr in comm["weight_attr"]:
                w = self._get_attr(comm["module"], attr)
                g = w.grad if w.grad is not None else getattr(w, "_ds_pending_grad", None)
                if g is not None:
                    dist.all_reduce(g, group=comm["group"])

    def get_tied_weights_and_groups(self):
        out = []
        for key, comm in self.tied_comms.items():
            for attr in comm["weight_attr"]:
                out.append((self._get_attr(comm["module"], attr), comm["group"]))
        return out

    # ---- info ------------------------------------------------------------------------------------------------
    def stage_owner(self, layer_idx):
        assert 0 <= layer_idx < self._num_layers
        for stage in range(self._topo.get_dim("pipe")):
            if self.parts[stage] <= layer_idx < self.parts[stage + 1]:
                return stage
        raise RuntimeError(f"Layer {layer_idx} not owned? parts={self.parts}")

    def topology(self):
        return self._topo

    def mpu(self):
        return self._grid

    def num_pipeline_stages(self):
        return self._topo.get_dim("pipe")

    def partitions(self):
        return self.parts

    # ---- per-layer checkpoints ----------------------------------------------------------------------------------
    def ckpt_prefix(self, checkpoints_path, tag):
        rank_name = "module"
        omit = ["data", "pipe"]
        rep = self._grid._topo.get_rank_repr(rank=self.global_rank, omit_axes=omit)
        if rep:
            rank_name += "-" + rep
        return os.path.join(checkpoints_path, str(tag), rank_name)

    def ckpt_layer_path(self, ckpt_dir, local_layer_idx):
        idx = local_layer_idx + self._local_start
        name = os.path.join(ckpt_dir, f"layer_{idx:02d}")
        rep = self._grid._topo.get_rank_repr(rank=self.global_rank, omit_axes=["data", "pipe"])
        if rep:
            name += f"-{rep}"
        return name + "-model_states.pt"

    def ckpt_layer_path_list(self, ckpt_dir, local_layer_idx):
        idx = local_layer_idx + self._local_start
        return sorted(glob.glob(os.path.join(ckpt_dir, f"layer_{idx:02d}-") + "*model_states.pt"))

    def save_state_dict(self, save_dir, checkpoint_engine, exclude_frozen_params=False):
        dp_rank = self._grid.data_parallel_id
        if dp_rank != 0:
            return
        os.makedirs(save_dir, exist_ok=True)
        for idx, layer in enumerate(self.forward_funcs):
            if not hasattr(layer, "state_dict"):
                continue
            sd = layer.state_dict()
            if exclude_frozen_params:
                frozen = {n for n, p in layer.named_parameters() if not p.requires_grad}
                sd = {k: v for k, v in sd.items() if k not in frozen}
            checkpoint_engine.save(collections.OrderedDict((k, v.detach().cpu().clone()) for k, v in sd.items()),
                                   self.ckpt_layer_path(save_dir, idx))

    def load_state_dir(self, load_dir, checkpoint_engine, strict=True):
        for idx, layer in enumerate(self.forward_funcs):
            if not hasattr(layer, "load_state_dict"):
                continue
            files = self.ckpt_layer_path_list(load_dir, idx)
            if not files:
                raise FileNotFoundError(f"no checkpoint file for layer {idx + self._local_start} in {load_dir}")
            sd = checkpoint_engine.load(files[0], map_location="cpu")
            layer.load_state_dict(sd, strict=strict)
        self._synchronize_tied_weights()

    def set_checkpoint_interval(self, interval):
        """Layers per activation-checkpoint segment (0 disables)."""
        assert interval >= 0
        self.activation_checkpoint_interval = interval

    @property
    def checkpoint_interval(self):
        return self.activation_checkpoint_interval

    def get_additional_losses(self):
        """Override to report ``{"name": value}`` extra losses from the last stage; ``None`` by default."""
        return None
