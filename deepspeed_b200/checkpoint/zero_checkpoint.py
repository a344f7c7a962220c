"""Reshape-aware view of the ZeRO optimizer-state files of a checkpoint (reference ``checkpoint/zero_checkpoint.py``)."""
import torch

from .constants import BASE_OPTIMIZER_STATE, GROUP_PADDINGS, OPTIMIZER_STATE_DICT, PARTITION_COUNT
from .reshape_3d_utils import get_model_3d_descriptor, model_3d_desc
from .reshape_utils import basic_folder_validation, get_zero_files, merge_state

GROUP_STATE_KEY = "state"


class ZeROCheckpoint:

    def __init__(self, dir):
        basic_folder_validation(dir)
        self.dir = dir
        self.file_list = get_zero_files(dir)
        self.num_files = len(self.file_list)
        assert self.num_files > 0, f"No ZeRO files found in {dir}"
        self.src_3d = get_model_3d_descriptor(dir)
        self.reshape(model_3d_desc(self.src_3d.pp_degree, self.src_3d.tp_degree, self.src_3d.dp_degree))

    def reshape(self, target_3d_desc: model_3d_desc):
        self.target_3d = target_3d_desc
        self._3d_file_map = self.src_3d.reshape(target_3d_desc)

    def get_src_world_size(self):
        return self.src_3d.world_size()

    def get_src_tp_degree(self):
        return self.src_3d.tp_degree

    def get_src_pp_degree(self):
        return self.src_3d.pp_degree

    def get_src_dp_degree(self):
        return self.src_3d.dp_degree

    def get_file_indices_for_rank(self, pp_index, tp_index, dp_index):
        assert dp_index < len(self._3d_file_map), f"DP index {dp_index} >= DP degree {len(self._3d_file_map)}"
        return self._3d_file_map[dp_index].get_data(pp_index, tp_index)

    def get_files_for_rank(self, pp_index, tp_index, dp_index):
        return [self.file_list[i] for i in self.get_file_indices_for_rank(pp_index, tp_index, dp_index)]

    def get_state_for_rank(self, pp_index, tp_index, dp_index, keys_to_ignore=(), strip_tensor_paddings=True):
        """Load and merge every source shard the target rank inherits (flat partitions are concatenated in rank order)."""
        merged = None
        for path in self.get_files_for_rank(pp_index, tp_index, dp_index):
            sd = torch.load(path, map_location="cpu", weights_only=False)
            for k in keys_to_ignore:
                sd.pop(k, None)
            if strip_tensor_paddings:
                self._strip_tensor_paddings(sd)
            merged = sd if merged is None else merge_state(merged, sd)
            opt = merged.get(OPTIMIZER_STATE_DICT) or {}
            if opt.get(PARTITION_COUNT):
                opt[PARTITION_COUNT] = [self.target_3d.dp_degree] * len(opt[PARTITION_COUNT])
            if strip_tensor_paddings and opt.get(GROUP_PADDINGS):
                opt[GROUP_PADDINGS] = [0] * len(opt[GROUP_PADDINGS])
        return merged

    def print_3d_index_map(self, tag=None):
        if tag:
            print(f"3D index map: {tag}")
        for dp, grid in enumerate(self._3d_file_map):
            grid.print_data(f"dp = {dp}")

    def print_3d_file_map(self, tag=None):
        if tag:
            print(f"3D file map: {tag}")
        for dp, grid in enumerate(self._3d_file_map):
            for p in range(grid.pp_degree):
                for t in range(grid.tp_degree):
                    print(f"{p}, {t}, {dp} => {[self.file_list[i] for i in grid.get_data(p, t)]}")

    @staticmethod
    def _strip_tensor_paddings(sd):
        """Drop the alignment padding at the tail of each group's flat optimizer state so shards concatenate exactly."""
        opt = sd.get(OPTIMIZER_STATE_DICT) or {}
        states = (opt.get(BASE_OPTIMIZER_STATE) or {}).get(GROUP_STATE_KEY)
        pads = opt.get(GROUP_PADDINGS)
        if states is None or pads is None:
            return
        for key, group_state in states.items():
            pad = pads[key]
            if pad == 0:
                continue
            for name, val in group_state.items():
                if name != "step" and torch.is_tensor(val) and val.dim() > 0:
                    group_state[name] = val.narrow(0, 0, val.numel() - pad).clone()
