"""Load Megatron-style checkpoints saved with a different model-parallel degree (reference
``runtime/state_dict_factory.py``: ``SDLoaderFactory :23``, ``MegatronSDLoader :190``): merge N source shards into one
target shard or split one source shard into N target shards; QKV weights are merged/split per (q, k, v) third, row-parallel
weights along dim 1, column-parallel along dim 0."""
import copy
import json
import os
from abc import ABC, abstractmethod

import torch

AUTO_MODULE_KEY = "auto"


class SDLoaderFactory:

    @staticmethod
    def get_sd_loader_json(json_file, checkpoint_engine=None):
        if isinstance(json_file, str):
            with open(json_file) as f:
                data = json.load(f)
        else:
            assert isinstance(json_file, dict)
            data = json_file
        sd_type = data["type"]
        ckpt_list = data["checkpoints"]
        version = data.get("version", 1.0)
        if "BLOOM" in sd_type or "Bloom" in sd_type or sd_type.lower() == "ds_model":
            return data
        return SDLoaderFactory.get_sd_loader(ckpt_list, checkpoint_engine, sd_type, version)

    @staticmethod
    def get_sd_loader(ckpt_list, checkpoint_engine=None, sd_type="Megatron", version=None):
        if sd_type == "Megatron":
            return MegatronSDLoader(ckpt_list, version, checkpoint_engine)
        raise ValueError(f"{sd_type} checkpoint type is not supported")


class SDLoaderBase(ABC):

    def __init__(self, ckpt_list, version, checkpoint_engine=None):
        self.module_key = None
        self.ckpt_list = ckpt_list
        self.version = version
        self.checkpoint_engine = checkpoint_engine
        self.check_ckpt_list()

    def _load(self, path):
        if self.checkpoint_engine is not None:
            return self.checkpoint_engine.load(path, map_location="cpu")
        return torch.load(path, map_location="cpu", weights_only=False)

    def load(self, mp_world_size, mp_rank, module_key=AUTO_MODULE_KEY, is_pipe_parallel=False, quantize=False,
             quantize_bits=8, quantize_groups=64, mlp_extra_grouping=True):
        self.module_key = module_key
        n = len(self.ckpt_list)
        idx = mp_rank * n // mp_world_size
        if is_pipe_parallel and module_key is not None and mp_world_size != n:
            mp_world_size = n
            idx = 0
        load_path = self.ckpt_list[idx]
        merge = n > mp_world_size
        if n == mp_world_size:
            assert os.path.exists(load_path), load_path
            sd = self._load(load_path)
            all_scales = None
            if quantize:
                from deepspeed_b200.runtime.weight_quantizer import WeightQuantization
                q = WeightQuantization(mlp_extra_grouping=mlp_extra_grouping, mp_size=mp_world_size)
                sd_module, all_scales = q.sd_quantize_megatron(self.get_module(sd), quantize_bits, quantize_groups)
                self.set_module(sd, sd_module)
        elif merge:
            sd, all_scales = self.merge_state_dict(mp_world_size, mp_rank, quantize, quantize_bits, quantize_groups,
                                                   mlp_extra_grouping)
        else:
            sd, all_scales = self.split_state_dict(mp_world_size, mp_rank, quantize, quantize_bits, quantize_groups,
                                                   mlp_extra_grouping)
        return load_path, sd, (all_scales, merge)

    def get_merge_state_dicts(self, mp_world_size, mp_rank):
        n = len(self.ckpt_list)
        assert n % mp_world_size == 0, "Invalid checkpoints and world size for sd merge"
        per = n // mp_world_size
        return [self._load(p) for p in self.ckpt_list[per * mp_rank:per * (mp_rank + 1)]]

    def get_split_state_dict(self, mp_world_size, mp_rank):
        n = len(self.ckpt_list)
        assert mp_world_size % n == 0, "Invalid checkpoints and world size for sd split"
        num_to_split = mp_world_size // n
        sd = self._load(self.ckpt_list[mp_rank // num_to_split])
        return sd, num_to_split, mp_rank % num_to_split

    def _choose_module_key(self, sd):
        assert not ("module" in sd and "model" in sd), "checkpoint has both 'model' and 'module' keys"
        assert "module" in sd or "model" in sd, "checkpoint contains neither 'model' or 'module' keys"
        return "module" if "module" in sd else "model"

    def get_module(self, sd):
        if self.module_key is None:
            return sd
        if self.module_key == AUTO_MODULE_KEY:
            return sd[self._choose_module_key(sd)]
        return sd[self.module_key]

    def set_module(self, sd, module):
        if self.module_key is None:
            sd = module
        elif self.module_key == AUTO_MODULE_KEY:
            sd[self._choose_module_key(sd)] = module
        else:
            sd[self.module_key] = module
        return sd

    def check_ckpt_list(self):
        assert len(self.ckpt_list) > 0

    @abstractmethod
    def merge_state_dict(self, mp_world_size, mp_rank, quantize, quantize_bits, groups, mlp_extra_grouping):
        ...

    @abstractmethod
    def split_state_dict(self, mp_world_size, mp_rank, quantize, quantize_bits, groups, mlp_extra_grouping):
        ...

    @abstractmethod
    def sanity_check(self, ckpt_file_name):
        ...


class MegatronSDLoader(SDLoaderBase):
    _ROW = ("attention.dense.weight", "mlp.dense_4h_to_h.weight")            # split along dim 1 (input features)
    _COL = ("mlp.dense_h_to_4h.weight", "mlp.dense_h_to_4h.bias", "word_embeddings.weight", "final_linear.weight")
    _QKV = ("attention.query_key_value", )

    def merge_query_key_value(self, param_list, ckpt_ver):
        """Version 0: [3*np*hn, h] stored as (q|k|v) blocks per shard -> concatenate thirds; version >= 1: heads are the
        outer dimension, plain concatenation along dim 0."""
        if ckpt_ver == 0:
            thirds = [p.chunk(3, dim=0) for p in param_list]
            return torch.cat([torch.cat([t[i] for t in thirds], dim=0) for i in range(3)], dim=0)
        return torch.cat(param_list, dim=0)

    def split_query_key_value(self, param, num_to_split, offset, ckpt_ver):
        if ckpt_ver == 0:
            q, k, v = param.chunk(3, dim=0)
            return torch.cat([t.chunk(num_to_split, dim=0)[offset] for t in (q, k, v)], dim=0)
        return param.chunk(num_to_split, dim=0)[offset]

    def merge_state_dict(self, mp_world_size, mp_rank, quantize=False, quantize_bits=8, groups=64, mlp_extra_grouping=True):
        self.sanity_check(self.ckpt_list[0])
        sd_list = self.get_merge_state_dicts(mp_world_size, mp_rank)
        ds_sd = copy.deepcopy(sd_list[0])
        mods = [self.get_module(sd) for sd in sd_list]
        ver = self.get_checkpoint_version(ds_sd)
        new = {}
        quantizer = None
        if quantize:
            from deepspeed_b200.runtime.weight_quantizer import WeightQuantization
            quantizer = WeightQuantization(mlp_extra_grouping=mlp_extra_grouping, mp_size=mp_world_size)
        for key in mods[0].keys():
            vals = [m[key] for m in mods]
            if any(k in key for k in self._ROW):
                if quantizer:
                    vals = quantizer.Quantize(vals, quantize_bits, groups, key=key, merge_dim=1)
                new[key] = torch.cat(vals, dim=1)
            elif any(k in key for k in self._QKV):
                if quantizer and "weight" in key:
                    vals = quantizer.Quantize(vals, quantize_bits, groups, key=key)
                    new[key] = torch.cat(vals, dim=0)
                else:
                    new[key] = self.merge_query_key_value(vals, ver)
            elif any(k in key for k in self._COL):
                if quantizer and "mlp.dense_h_to_4h.weight" in key:
                    vals = quantizer.Quantize(vals, quantize_bits, groups, key=key)
                new[key] = torch.cat(vals, dim=0)
            else:
                new[key] = vals[0]
        scales = quantizer.merge_scales() if quantizer else None
        return self.set_module(ds_sd, new), scales

    def split_state_dict(self, mp_world_size, mp_rank, quantize=False, quantize_bits=8, groups=64, mlp_extra_grouping=True):
        sd, num_to_split, offset = self.get_split_state_dict(mp_world_size, mp_rank)
        ds_sd = copy.deepcopy(sd)
        mod = self.get_module(sd)
        ver = self.get_checkpoint_version(ds_sd)
        new = {}
        quantizer = None
        if quantize:
            from deepspeed_b200.runtime.weight_quantizer import WeightQuantization
            quantizer = WeightQuantization(mlp_extra_grouping=mlp_extra_grouping, mp_size=mp_world_size)
        for key, value in mod.items():
            if any(k in key for k in self._ROW):
                v = value.chunk(num_to_split, dim=1)[offset]
                if quantizer:
                    v = quantizer.Quantize([v], quantize_bits, groups, key)[0]
                new[key] = v
            elif any(k in key for k in self._QKV):
                v = self.split_query_key_value(value, num_to_split, offset, ver)
                if quantizer and "weight" in key:
                    v = quantizer.Quantize([v], quantize_bits, groups, key)[0]
                new[key] = v
            elif any(k in key for k in self._COL):
                v = value.chunk(num_to_split, dim=0)[offset]
                if quantizer and "mlp.dense_h_to_4h.weight" in key:
                    v = quantizer.Quantize([v], quantize_bits, groups, key)[0]
                new[key] = v
            else:
                new[key] = value
        scales = quantizer.merge_scales_split(num_to_split) if quantizer else None
        return self.set_module(ds_sd, new), scales

    def sanity_check(self, ckpt_file_name):
        needed = ["attention.dense.weight", "mlp.dense_4h_to_h.weight", "attention.query_key_value",
                  "mlp.dense_h_to_4h.weight", "mlp.dense_h_to_4h.bias"]
        sd = self._load(ckpt_file_name)
        keys = self.get_module(sd).keys() if self.module_key is not None else sd.keys()
        for n in needed:
            assert any(n in k for k in keys), f"key: {n} is not found in the checkpoint {ckpt_file_name}"

    def get_checkpoint_version(self, state_dict):
        return self.version if self.version is not None else state_dict.get("checkpoint_version", 0)
