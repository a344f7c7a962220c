"""MiCS: shard model states inside a *sub-group* of the data-parallel world and replicate across sub-groups.

Parity target: reference ``runtime/zero/mics.py`` (``MiCS_Init :64``, ``MiCS_Optimizer :361``) and
``mics_utils.py`` (group construction).  In this framework MiCS is the same ``ZeroShardedOptimizer`` with two
communicators: ``dp_group`` = the shard group (all-gather / reduce-scatter scope) and ``replica_group`` = ranks
holding the same shard index (gradient all-reduce scope) — i.e. exactly the hierarchical communication of the
paper, with the intra-group collectives riding NVSwitch.
"""
from dataclasses import dataclass
from typing import List

from deepspeed_b200 import comm as dist
from deepspeed_b200.runtime.zero.partition_parameters import Init


@dataclass
class MiCS_CommGroups:
    param_shard_group: object = None
    param_shard_size: int = -1
    param_shard_rank: int = -1
    param_repli_group: object = None
    param_repli_size: int = -1
    param_repli_rank: int = -1
    param_intra_node_group: object = None
    param_inter_node_shard_group: object = None
    shard_ranks: List[List[int]] = None
    repli_ranks: List[List[int]] = None


def mics_rank_layout(world_ranks: List[int], shard_size: int):
    """-> (shard groups, replica groups) as lists of global ranks."""
    n = len(world_ranks)
    assert n % shard_size == 0, f"DP world {n} is not divisible by mics_shard_size {shard_size}"
    shard = [world_ranks[i:i + shard_size] for i in range(0, n, shard_size)]
    repli = [[g[i] for g in shard] for i in range(shard_size)]
    return shard, repli


_cache = {}


def hierarchy_layout(shard_group: List[int], ndev_per_node: int):
    """Split one shard group that spans nodes into (intra-node groups, inter-node groups): intra = consecutive runs of
    ``ndev_per_node`` ranks; inter = ranks with the same local index across the nodes."""
    assert len(shard_group) % ndev_per_node == 0
    intra = [shard_group[i:i + ndev_per_node] for i in range(0, len(shard_group), ndev_per_node)]
    inter = [[g[i] for g in intra] for i in range(ndev_per_node)]
    return intra, inter


def _generate_mics_config(world_size, ndev_per_node, shard_size, pp_size=1):
    """Rank lists of a MiCS layout: ``{"shard_groups", "replicate_groups", "span_nodes"}`` (reference
    ``mics_utils._generate_mics_config``); with pipeline parallelism each stage owns a contiguous block of ranks."""
    assert world_size % pp_size == 0
    per_stage = world_size // pp_size
    assert per_stage % shard_size == 0, f"dp size {per_stage} is not divisible by the MiCS shard size {shard_size}"
    shard, repli = [], []
    for st in range(pp_size):
        s, r = mics_rank_layout(list(range(st * per_stage, (st + 1) * per_stage)), shard_size)
        shard += s
        repli += r
    return {"shard_groups": shard, "replicate_groups": repli, "span_nodes": max(1, shard_size // ndev_per_node)}


def scale_tensors(tensors, scale):
    for t in tensors:
        t.div_(scale)


def hierarchical_all_gather(output, shard, groups: "MiCS_CommGroups"):
    """All-gather ``shard`` over a shard group spanning nodes in two hops (reference ``MiCS_AllGatherCoalescedHandle`` /
    ``_hierarchical_all_gather_params``): first across nodes between same-local-index ranks (small messages on the slow
    fabric, all local ranks in parallel), then inside the node over NVLink.  ``output`` is ordered by shard rank."""
    import torch
    inter, intra = groups.param_inter_node_shard_group, groups.param_intra_node_group
    if inter is None or intra is None:
        return dist.all_gather_into_tensor(output, shard, group=groups.param_shard_group)
    n_nodes, n_local = dist.get_world_size(inter), dist.get_world_size(intra)
    n = shard.numel()
    stage1 = torch.empty(n_nodes * n, dtype=shard.dtype, device=shard.device)  # [node][n] for my local index
    dist.all_gather_into_tensor(stage1, shard.contiguous().view(-1), group=inter)
    stage2 = torch.empty(n_local * n_nodes * n, dtype=shard.dtype, device=shard.device)  # [local][node][n]
    dist.all_gather_into_tensor(stage2, stage1, group=intra)
    # shard rank = node * n_local + local
    output.view(n_nodes, n_local, n).copy_(stage2.view(n_local, n_nodes, n).transpose(0, 1))
    return None


def create_mics_comm_groups(shard_size, dp_group=None, hierarchical_allgather=False, mpu=None, ndev_per_node=None) -> MiCS_CommGroups:
    key = (shard_size, id(dp_group), bool(hierarchical_allgather), ndev_per_node)
    if key in _cache:
        return _cache[key]
    world = dist.get_world_size(dp_group)
    ranks = [dist.get_global_rank(dp_group, i) for i in range(world)] if dp_group is not None else list(range(world))
    shard, repli = mics_rank_layout(ranks, shard_size)
    me = dist.get_rank()
    g = MiCS_CommGroups(shard_ranks=shard, repli_ranks=repli)
    for rs in shard:
        h = dist.new_group(rs)
        if me in rs:
            g.param_shard_group, g.param_shard_size, g.param_shard_rank = h, len(rs), rs.index(me)
    for rs in repli:
        h = dist.new_group(rs)
        if me in rs:
            g.param_repli_group, g.param_repli_size, g.param_repli_rank = h, len(rs), rs.index(me)
    if hierarchical_allgather:
        import os
        import torch
        per_node = ndev_per_node or int(os.environ.get("LOCAL_WORLD_SIZE", 0)) or max(1, torch.cuda.device_count())
        if shard_size > per_node and shard_size % per_node == 0:  # only when a shard group really spans nodes
            for rs in shard:
                intra, inter = hierarchy_layout(rs, per_node)
                for sub in intra:
                    h = dist.new_group(sub)
                    if me in sub:
                        g.param_intra_node_group = h
                for sub in inter:
                    h = dist.new_group(sub)
                    if me in sub:
                        g.param_inter_node_shard_group = h
    _cache[key] = g
    return g


class MiCS_Init(Init):
    """``zero.Init`` whose partitioning scope is the MiCS shard group (parameters are sharded ``mics_shard_size``
    ways instead of over the whole DP world)."""

    def __init__(self, module=None, data_parallel_group=None, sequence_data_parallel_group=None, mem_efficient_linear=True,
                 remote_device=None, pin_memory=False, config_dict_or_path=None, config=None, enabled=True, dtype=None,
                 mpu=None):
        cfg = config_dict_or_path if config_dict_or_path is not None else config
        shard = -1
        if isinstance(cfg, dict):
            shard = cfg.get("zero_optimization", {}).get("mics_shard_size", -1)
        self.mics_comm_groups = None
        if enabled and shard and shard > 0 and dist.is_initialized():
            self.mics_comm_groups = create_mics_comm_groups(shard, data_parallel_group, mpu=mpu)
            data_parallel_group = self.mics_comm_groups.param_shard_group
        super().__init__(module=module, data_parallel_group=data_parallel_group, mem_efficient_linear=mem_efficient_linear,
                         remote_device=remote_device, pin_memory=pin_memory, config_dict_or_path=config_dict_or_path,
                         config=config, enabled=enabled, dtype=dtype, mpu=mpu)


def has_hierarchical_all_gather_groups(comm_groups: "MiCS_CommGroups") -> bool:
    """Two-level (intra-node, then inter-node) all-gather groups exist (reference ``mics.py:29``)."""
    return getattr(comm_groups, "param_intra_node_group", None) is not None and \
        getattr(comm_groups, "param_inter_node_shard_group", None) is not None


class MiCS_AllGatherCoalescedHandle:
    """Handle of a MiCS coalesced gather: identical to the ZeRO-3 handle, scoped to the shard group (the hierarchical
    variant issues its two stages inside ``hierarchical_all_gather`` and hands back a completed handle)."""

    def __new__(cls, allgather_handle, params, partitions, world_size):
        from .gather_handles import AllGatherCoalescedHandle
        return AllGatherCoalescedHandle(allgather_handle, params, partitions, world_size)


def _mics_offload_cls():
    from .parameter_offload import DeepSpeedZeRoOffload

    class MiCS_Offload(DeepSpeedZeRoOffload):
        """Parameter-offload manager whose partitioning scope is the MiCS shard group (reference ``mics.py:334``)."""

        def __init__(self, module, *args, ds_config=None, dp_process_group=None, mpu=None, **kw):
            shard = getattr(getattr(ds_config, "zero_config", None), "mics_shard_size", -1) if ds_config is not None else -1
            self.mics_comm_groups = None
            if shard and shard > 0 and dist.is_initialized():
                self.mics_comm_groups = create_mics_comm_groups(shard, dp_process_group, mpu=mpu)
                dp_process_group = self.mics_comm_groups.param_shard_group
            super().__init__(module, *args, ds_config=ds_config, dp_process_group=dp_process_group, mpu=mpu, **kw)

    return MiCS_Offload


def __getattr__(name):
    if name == "MiCS_Offload":  # built lazily: parameter_offload imports the sharded optimizer
        cls = _mics_offload_cls()
        globals()[name] = cls
        return cls
    raise AttributeError(name)


def MiCS_Optimizer(module, *, shard_size, dp_group=None, **kw):
    """Factory: the ZeRO-3 optimizer over (shard group, replica group)."""
    from deepspeed_b200.runtime.zero.sharded import ZeroShardedOptimizer
    g = create_mics_comm_groups(shard_size, dp_group)
    return ZeroShardedOptimizer(module, 3, dp_group=g.param_shard_group, replica_group=g.param_repli_group, **kw)
