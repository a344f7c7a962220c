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
        self.original_pp_degree = 1
        self.tp_degree = tp_degree or self.original_tp_degree
        self.pp_degree = pp_degree or self.original_pp_degree
        self.dp_degree = dp_degree or self.original_dp_degree
        self.world_size = self.tp_degree * self.pp_degree * self.dp_degree
        self._model_state = None

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
