"""Cartesian process topologies for 3D parallelism.

Parity target: reference ``runtime/pipe/topology.py`` (``ProcessTopology :12``, ``PipeDataParallelTopology``,
``PipeModelDataParallelTopology :244``, ``PipelineParallelGrid :251``).  Pure rank <-> coordinate math (unit
tested without processes) plus the grid object that materialises process groups.
"""
import itertools
from collections import namedtuple
from typing import Dict, List

from deepspeed_b200 import comm as dist


class ProcessTopology:
    """Row-major mapping of ranks onto named axes; the LAST axis varies fastest."""

    def __init__(self, axes: List[str], dims: List[int]):
        assert len(axes) == len(dims)
        self.axes = list(axes)
        self.dims = list(dims)
        self.ProcessCoord = namedtuple("ProcessCoord", axes)
        self.mapping: Dict = {}
        for rank, coord in enumerate(itertools.product(*[range(d) for d in dims])):
            self.mapping[self.ProcessCoord(*coord)] = rank

    def get_rank(self, **coord_kwargs):
        if len(coord_kwargs) != len(self.axes):
            raise ValueError("get_rank() does not support slices. Use filter_match())")
        key = self.ProcessCoord(**coord_kwargs)
        assert key in self.mapping, f"key {coord_kwargs} invalid"
        return self.mapping[key]

    def get_axis_names(self):
        return self.axes

    def get_rank_repr(self, rank, omit_axes=("data", "pipe"), inner_sep="_", outer_sep="-"):
        """String such as ``model_00`` used in per-layer checkpoint file names."""
        omit = frozenset(omit_axes)
        axes = [a for a in self.axes if a not in omit]
        coord = self.get_coord(rank)
        return outer_sep.join(f"{a}{inner_sep}{getattr(coord, a):02d}" for a in axes)

    def get_dim(self, axis):
        if axis not in self.axes:
            return 0
        return self.dims[self.axes.index(axis)]

    def get_coord(self, rank):
        for coord, r in self.mapping.items():
            if r == rank:
                return coord
        raise ValueError(f"rank {rank} not found in topology.")

    def get_axis_comm_lists(self, axis):
        """Rank lists that differ only along ``axis`` (one communicator per list)."""
        if axis not in self.axes:
            return []
        others = [a for a in self.axes if a != axis]
        lists = []
        for coord in itertools.product(*[range(self.get_dim(a)) for a in others]):
            fixed = dict(zip(others, coord))
            lists.append([self.mapping[self.ProcessCoord(**fixed, **{axis: i})] for i in range(self.get_dim(axis))])
        return lists

    def filter_match(self, **filter_kwargs):
        return sorted(r for c, r in self.mapping.items() if all(getattr(c, k) == v for k, v in filter_kwargs.items()))

    def get_axis_list(self, axis, idx):
        ax = self.axes.index(axis)
        return sorted(r for c, r in self.mapping.items() if c[ax] == idx)

    def world_size(self):
        return len(self.mapping)

    def __str__(self):
        return str(self.mapping)


class PipeDataParallelTopology(ProcessTopology):
    """pipe x data: adjacent ranks are data-parallel replicas (gradient all-reduce stays on NVLink)."""

    def __init__(self, num_pp, num_dp):
        super().__init__(axes=["pipe", "data"], dims=[num_pp, num_dp])


class PipeModelDataParallelTopology(ProcessTopology):

    def __init__(self, num_pp, num_mp, num_dp):
        super().__init__(axes=["pipe", "data", "model"], dims=[num_pp, num_dp, num_mp])


class PipelineParallelGrid:
    """Process groups for a topology; doubles as a Megatron-style ``mpu`` object."""

    def __init__(self, topology=None, process_group=None):
        self.global_rank = dist.get_rank()
        self.world_size = dist.get_world_size()
        if topology is None:
            topology = PipeDataParallelTopology(1, self.world_size)
        self._topo = topology
        self.data_parallel_size = max(topology.get_dim("data"), 1)
        self.pipe_parallel_size = max(topology.get_dim("pipe"), 1)
        self.model_parallel_size = max(topology.get_dim("model"), 1)
        self.slice_parallel_size = self.model_parallel_size
        assert self.data_parallel_size * self.pipe_parallel_size * self.model_parallel_size == self.world_size
        coord = topology.get_coord(self.global_rank)
        self.stage_id = getattr(coord, "pipe", 0)
        self.data_parallel_id = getattr(coord, "data", 0)
        self.dp_group = self.dp_proc_group = None
        self.dp_groups = topology.get_axis_comm_lists("data")
        for ranks in self.dp_groups:
            g = dist.new_group(ranks=ranks)
            if self.global_rank in ranks:
                self.dp_group, self.dp_proc_group = ranks, g
        self.is_first_stage = self.stage_id == 0
        self.is_last_stage = self.stage_id == self.pipe_parallel_size - 1
        self.p2p_groups = self._build_p2p_groups()
        self.pp_group = self.pp_proc_group = None
        self.pipe_groups = topology.get_axis_comm_lists("pipe")
        for ranks in self.pipe_groups:
            g = dist.new_group(ranks=ranks)
            if self.global_rank in ranks:
                self.pp_group, self.pp_proc_group = ranks, g
        self.slice_group = self.slice_proc_group = None
        if self.model_parallel_size > 1:
            for ranks in topology.get_axis_comm_lists("model"):
                g = dist.new_group(ranks=ranks)
                if self.global_rank in ranks:
                    self.slice_group, self.slice_proc_group = ranks, g
        else:
            self.slice_group = [self.global_rank]
            self.slice_proc_group = dist.new_group(ranks=[self.global_rank]) if False else None

    def _build_p2p_groups(self):
        """(this rank, next-stage rank) pairs along every pipe communicator (wrap-around included)."""
        pairs = []
        for ranks in self._topo.get_axis_comm_lists("pipe"):
            for i, r in enumerate(ranks):
                pairs.append([r, ranks[(i + 1) % len(ranks)]])
        return pairs

    def get_stage_id(self):
        return self.stage_id

    def get_data_parallel_id(self):
        return self.data_parallel_id

    def stage_to_global(self, stage_id, **kwargs):
        me = self._topo.get_coord(self.global_rank)
        transform = me._replace(pipe=stage_id, **kwargs)._asdict()
        return self._topo.get_rank(**transform)

    def topology(self):
        return self._topo

    # ---- mpu API ------------------------------------------------------------------------------------
    def get_global_rank(self):
        return self.global_rank

    def get_pipe_parallel_rank(self):
        return self.stage_id

    def get_pipe_parallel_world_size(self):
        return self.pipe_parallel_size

    def get_pipe_parallel_group(self):
        return self.pp_proc_group

    def get_data_parallel_rank(self):
        return self.data_parallel_id

    def get_data_parallel_world_size(self):
        return self.data_parallel_size

    def get_data_parallel_group(self):
        return self.dp_proc_group

    def get_model_parallel_rank(self):
        return getattr(self._topo.get_coord(self.global_rank), "model", 0)

    def get_model_parallel_world_size(self):
        return self.model_parallel_size

    def get_model_parallel_group(self):
        return self.slice_proc_group

    get_slice_parallel_rank = get_model_parallel_rank
    get_slice_parallel_world_size = get_model_parallel_world_size
    get_slice_parallel_group = get_model_parallel_group
