"""Inspect a checkpoint directory (reference ``checkpoint/deepspeed_checkpoint.py:38 DeepSpeedCheckpoint``)."""
import glob
import os
import re

import torch


class DeepSpeedCheckpoint:

    def __init__(self, dir, tp_degree=None, pp_degree=None, dp_degree=None, final_layer_norm_idx=-1):
        self.dir = dir
        self.file_list = sorted(glob.glob(os.path.join(dir, "*.pt")))
        self.zero_files = [f for f in self.file_list if f.endswith("_optim_states.pt")]
        self.layer_files = [f for f in self.file_list if os.path.basename(f).startswith("layer_")]
        self.mp_rank_files = [f for f in self.file_list if re.search(r"mp_rank_\d+_model_states\.pt$", f)
                              and "layer_" not in os.path.basename(f)]
        dp, mp = set(), set()
        for f in self.zero_files:
            m = re.search(r"zero_pp_rank_(\d+)_mp_rank_(\d+)", os.path.basename(f))
            if m:
                dp.add(int(m.group(1)))
                mp.add(int(m.group(2)))
        self.original_dp_degree = len(dp) or 1
        self.original_tp_degree = len(mp) or max(1, len(set(re.search(r"mp_rank_(\d+)", os.path.basename(f)).group(1)
                                                            for f in self.mp_rank_files)))
        self.original_pp_degree = max(1, len(self.mp_rank_files) // max(1, self.original_tp_degree))
        self.tp_degree = tp_degree or self.original_tp_degree
        self.pp_degree = pp_degree or self.original_pp_degree
        self.dp_degree = dp_degree or self.original_dp_degree
        self.world_size = self.tp_degree * self.pp_degree * self.dp_degree
        self._model_state = None
        self.final_layer_norm_idx = final_layer_norm_idx
        self._maps = None

    def is_change_tp_degree(self):
        return self.tp_degree != self.original_tp_degree

    def is_change_pp_degree(self):
        return self.pp_degree != self.original_pp_degree

    def is_change_dp_degree(self):
        return self.dp_degree != self.original_dp_degree

    def show_file_list(self):
        for f in self.file_list:
            print(f)

    def _ms(self):
        if self._model_state is None:
            self._model_state = torch.load(self.mp_rank_files[0], map_location="cpu", weights_only=False)
        return self._model_state

    def get_iteration(self):
        return self._ms().get("global_steps", 0)

    def get_args(self):
        return self._ms().get("args")

    def get_checkpoint_info(self, key="checkpoint_info"):
        return self._ms().get(key)

    def get_zero_files(self):
        return self.zero_files

    def get_zero_checkpoint_state(self, pp_index=0, tp_index=0, dp_index=0):
        for f in self.zero_files:
            if re.search(rf"zero_pp_rank_{dp_index}_mp_rank_{tp_index:02d}", os.path.basename(f)):
                return torch.load(f, map_location="cpu", weights_only=False)
        raise FileNotFoundError((pp_index, tp_index, dp_index))

    # ---- pipeline layer files (``layer_<idx>-model_<tp>-model_states.pt``) --------------------------------------------
    # Keys are the ``layer_<idx>`` prefixes in file order: key 0 is the embedding, ``final_layer_norm_idx`` the final norm,
    # everything in between transformer layers (reference ``deepspeed_checkpoint.py:209-262``).  Maps are built on first use.
    def _layer_maps(self):
        if self._maps is not None:
            return self._maps
        from .reshape_meg_2d import reshape_meg_2d_parallel
        from .reshape_utils import get_files_with_prefix, partition_data
        keys = sorted({re.match(r"(layer_\d+)", os.path.basename(f)).group(1) for f in self.layer_files})
        per_tp = lambda key: {i: part for i, part in enumerate(
            partition_data(get_files_with_prefix(self.layer_files, key + "-"), self.tp_degree))} if keys else {}
        m = {"keys": keys, "embedding": per_tp(keys[0]) if keys else {},
             "final_norm": per_tp(keys[self.final_layer_norm_idx]) if keys else {}}
        body = keys[1:self.final_layer_norm_idx] if keys else []
        per_stage = max(1, len(body) // max(1, self.pp_degree))
        m["pp_to_transformer"] = {pp: body[pp * per_stage:(pp + 1) * per_stage] for pp in range(self.pp_degree)}
        files = {}
        for idx, key in enumerate(body):
            pp = min(idx // per_stage, self.pp_degree - 1)
            parts = partition_data(get_files_with_prefix(self.layer_files, key + "-"), self.tp_degree)
            for tp in range(self.tp_degree):
                files.setdefault((tp, pp), []).append(parts[tp])
        m["transformer_files"] = files
        m["2d"] = reshape_meg_2d_parallel(old_pp_degree=self.original_pp_degree, old_tp_degree=self.original_tp_degree,
                                          new_pp_degree=self.pp_degree, new_tp_degree=self.tp_degree) \
            if self.mp_rank_files else None
        self._maps = m
        return m

    @property
    def layer_keys(self):
        return self._layer_maps()["keys"]

    @property
    def tp_to_embedding_map(self):
        return self._layer_maps()["embedding"]

    @property
    def tp_to_final_norm_map(self):
        return self._layer_maps()["final_norm"]

    @property
    def pp_to_transformer_map(self):
        return self._layer_maps()["pp_to_transformer"]

    @property
    def transformer_file_map(self):
        return self._layer_maps()["transformer_files"]

    @staticmethod
    def _load(path):
        return torch.load(path, map_location="cpu", weights_only=False)

    @staticmethod
    def _merge_state_dicts(sd_list):
        """TP slices of one layer → one state dict (per-key merge rule: see ``reshape_utils.merge_state``)."""
        from .reshape_utils import merge_state
        out = sd_list[0]
        for sd in sd_list[1:]:
            out = merge_state(out, sd)
        return out

    @staticmethod
    def _dump_mapping(data_map, map_tag=None):
        if map_tag is not None:
            print(f"Dump mapping: {map_tag}")
        for k, v in data_map.items():
            print(f"{k} = {v}")

    def get_embedding_layer_id(self):
        return self.layer_keys[0]

    def get_final_norm_layer_id(self):
        return self.layer_keys[self.final_layer_norm_idx]

    def get_embedding_files(self, tp_index: int) -> list:
        assert tp_index in self.tp_to_embedding_map
        return self.tp_to_embedding_map[tp_index]

    def get_embedding_state(self, tp_index: int):
        return self._merge_state_dicts([self._load(f) for f in self.get_embedding_files(tp_index)])

    def get_final_norm_files(self, tp_index: int) -> list:
        assert tp_index in self.tp_to_final_norm_map
        return self.tp_to_final_norm_map[tp_index]

    def get_final_norm_state(self, tp_index: int):
        return self._load(self.get_final_norm_files(tp_index)[0])

    def get_pp_transformer_map(self, pp_index: int) -> list:
        assert pp_index < self.pp_degree
        return self.pp_to_transformer_map[pp_index]

    def get_transformer_state(self, tp_index: int, pp_index: int) -> list:
        assert tp_index < self.tp_degree and pp_index < self.pp_degree
        return [self._merge_state_dicts([self._load(f) for f in group])
                for group in self.transformer_file_map.get((tp_index, pp_index), [])]

    def get_2d_parallel_files(self, tp_index: int, pp_index: int) -> list:
        assert tp_index < self.tp_degree and pp_index < self.pp_degree
        idx = self._layer_maps()["2d"].get_data(pp_index=pp_index, tp_index=tp_index)
        return [self.mp_rank_files[i] for i in idx]

    def get_2d_parallel_state(self, tp_index: int, pp_index: int) -> dict:
        from .reshape_utils import merge_state
        merged = None
        for f in self.get_2d_parallel_files(tp_index=tp_index, pp_index=pp_index):
            sd = self._load(f)
            merged = sd if merged is None else merge_state(merged, sd)
        return merged

    def show_2d_mapping(self):
        print("reshaped 2d map ---- begin")
        for i in range(self.pp_degree):
            for j in range(self.tp_degree):
                print(f"[{i}, {j}] = {self.get_2d_parallel_files(pp_index=i, tp_index=j)}")
        print("reshaped 2d map ---- end")

    def show_tp_embedding_map(self):
        self._dump_mapping(self.tp_to_embedding_map, "tp_to_embedding_layers")

    def show_tp_final_norm_map(self):
        self._dump_mapping(self.tp_to_final_norm_map, "tp_to_final_norm_layers")

    def show_pp_transformer_map(self):
        self._dump_mapping(self.pp_to_transformer_map, "pp_to_transformer_layers")

    def show_transformer_file_map(self):
        self._dump_mapping(self.transformer_file_map, "rank_to_transformer_files")

    def validate_files(self):
        missing = [f for f in self.file_list if not os.path.isfile(f)]
        for f in missing:
            print(f"Error: {f} is not existent")
        return not missing
