"""Process-group registry: DP / TP / PP / SP / EP / expert-DP / hpZ groups.

Parity target: reference ``utils/groups.py`` (``_create_expert_and_data_parallel :236``,
``_create_expert_data_and_model_parallel :376``, ``_get_sequence_* :591-643``,
``_create_zero_param_parallel_group :650``, ``_get_local_all_to_all_group :490``).

Design: every layout is computed by one pure function, :func:`rank_layout`, which maps a world
size and the parallel degrees onto lists of ranks (testable without any process group), and the
registry simply materialises ``dist.new_group`` for each list.  Rank order on a node is
``[pp][dp][sp][tp]`` with tp fastest, so TP / SP / EP groups stay on NVSwitch-adjacent ranks.
"""
from typing import Dict, List, Optional

import torch.distributed as dist

_mpu = None
_groups: Dict[str, object] = {}
_ranks: Dict[str, List[int]] = {}
_expert_parallel_size: Dict[str, int] = {}
_tp_size = 1
_pp_size = 1
_sp_size = 1
mesh_device = None
expert_tensor_parallel_world_size = 1


def reset():
    global _mpu, _tp_size, _pp_size, _sp_size, mesh_device
    _groups.clear()
    _ranks.clear()
    _expert_parallel_size.clear()
    _mpu = None
    _tp_size = _pp_size = _sp_size = 1
    mesh_device = None


def _world():
    return dist.get_world_size() if dist.is_initialized() else 1


def _rank():
    return dist.get_rank() if dist.is_initialized() else 0


# --------------------------------------------------------------------------------------------
# pure layout math
# --------------------------------------------------------------------------------------------
def rank_layout(world: int, tp: int = 1, pp: int = 1, sp: int = 1) -> Dict[str, List[List[int]]]:
    """All groups for a (pp, dp, sp, tp) grid, tp fastest-varying.

    Returns dict with keys ``tp``, ``sp``, ``dp``, ``pp``, ``sdp`` (sequence x data, the ZeRO
    sharding group under Ulysses) and ``mp`` (= tp x pp slices sharing a data shard).
    """
    assert world % (tp * pp * sp) == 0, f"world {world} not divisible by tp*pp*sp = {tp * pp * sp}"
    dp = world // (tp * pp * sp)

    def rid(p, d, s, t):
        return ((p * dp + d) * sp + s) * tp + t

    out = {k: [] for k in ("tp", "sp", "dp", "pp", "sdp", "mp")}
    for p in range(pp):
        for d in range(dp):
            for s in range(sp):
                out["tp"].append([rid(p, d, s, t) for t in range(tp)])
    for p in range(pp):
        for d in range(dp):
            for t in range(tp):
                out["sp"].append([rid(p, d, s, t) for s in range(sp)])
    for p in range(pp):
        for s in range(sp):
            for t in range(tp):
                out["dp"].append([rid(p, d, s, t) for d in range(dp)])
    for d in range(dp):
        for s in range(sp):
            for t in range(tp):
                out["pp"].append([rid(p, d, s, t) for p in range(pp)])
    for p in range(pp):
        for t in range(tp):
            out["sdp"].append([rid(p, d, s, t) for d in range(dp) for s in range(sp)])
    for d in range(dp):
        for s in range(sp):
            out["mp"].append([rid(p, d, s, t) for p in range(pp) for t in range(tp)])
    return out


def expert_layout(dp_ranks: List[int], ep_size: int, data_before_expert: bool = False):
    """Split one DP group into expert-parallel and expert-data-parallel groups.

    Default ("E+D"): consecutive ``ep_size`` ranks form an EP group; ranks with the same offset
    form an expert-DP group.  ``data_before_expert`` ("D+E") swaps the roles (reference
    ``use_data_before_expert_parallel_``).
    """
    n = len(dp_ranks)
    assert n % ep_size == 0, f"dp size {n} not divisible by ep_size {ep_size}"
    ep_groups, edp_groups = [], []
    if not data_before_expert:
        for i in range(0, n, ep_size):
            ep_groups.append(dp_ranks[i:i + ep_size])
        for off in range(ep_size):
            edp_groups.append(dp_ranks[off::ep_size])
    else:
        stride = n // ep_size
        for off in range(stride):
            ep_groups.append(dp_ranks[off::stride])
        for i in range(0, n, stride):
            edp_groups.append(dp_ranks[i:i + stride])
    return ep_groups, edp_groups


# --------------------------------------------------------------------------------------------
# registry
# --------------------------------------------------------------------------------------------
def _register(name: str, rank_lists: List[List[int]]):
    """Create every group in ``rank_lists`` (collective call) and remember the one we are in."""
    me = _rank()
    for ranks in rank_lists:
        g = dist.new_group(ranks=ranks) if dist.is_initialized() else None
        if me in ranks:
            _groups[name] = g
            _ranks[name] = list(ranks)


def initialize(ep_size=1, mpu=None, tp_size: int = 1, pp_size: int = 1, sp_size: int = 1):
    """Build the (pp, dp, sp, tp) grid groups; ``mpu`` (Megatron-style) overrides tp/pp groups."""
    global _mpu, _tp_size, _pp_size, _sp_size
    _mpu = mpu
    if mpu is not None:
        return
    _tp_size, _pp_size, _sp_size = tp_size, pp_size, sp_size
    if tp_size * pp_size * sp_size == 1:
        return
    lay = rank_layout(_world(), tp_size, pp_size, sp_size)
    for k in ("tp", "sp", "dp", "pp", "sdp", "mp"):
        _register(k, lay[k])


def _mpu_call(*names):
    for n in names:
        if _mpu is not None and hasattr(_mpu, n):
            return getattr(_mpu, n)()
    return None


def _get_data_parallel_group():
    if _mpu is not None:
        return _mpu_call("get_data_parallel_group")
    if mesh_device is not None:
        return mesh_device.get_group(mesh_dim="data_parallel")
    return _groups.get("dp")  # None == WORLD


def _get_data_parallel_world_size():
    if _mpu is not None:
        return _mpu_call("get_data_parallel_world_size")
    if "dp" in _ranks:
        return len(_ranks["dp"])
    if mesh_device is not None:
        return dist.get_world_size(_get_data_parallel_group())
    return _world() // max(_tp_size * _pp_size * _sp_size, 1)


def _get_data_parallel_rank():
    if _mpu is not None:
        return _mpu_call("get_data_parallel_rank")
    if "dp" in _ranks:
        return _ranks["dp"].index(_rank())
    if mesh_device is not None:
        return dist.get_rank(_get_data_parallel_group())
    return _rank()


def _get_model_parallel_group():
    if _mpu is not None:
        return _mpu_call("get_model_parallel_group", "get_tensor_model_parallel_group")
    return _groups.get("tp")


def _get_model_parallel_world_size():
    if _mpu is not None:
        return _mpu_call("get_model_parallel_world_size", "get_tensor_model_parallel_world_size") or 1
    return len(_ranks["tp"]) if "tp" in _ranks else 1


def _get_model_parallel_rank():
    if _mpu is not None:
        return _mpu_call("get_model_parallel_rank", "get_tensor_model_parallel_rank") or 0
    return _ranks["tp"].index(_rank()) if "tp" in _ranks else 0


get_tensor_model_parallel_group = _get_model_parallel_group
get_tensor_model_parallel_world_size = _get_model_parallel_world_size
get_tensor_model_parallel_rank = _get_model_parallel_rank


def _get_pipe_parallel_group():
    return _groups.get("pp")


def _get_pipe_parallel_world_size():
    return len(_ranks["pp"]) if "pp" in _ranks else 1


def _get_pipe_parallel_rank():
    return _ranks["pp"].index(_rank()) if "pp" in _ranks else 0


# ---- sequence parallel ---------------------------------------------------------------------
def _get_sequence_parallel_group():
    if _mpu is not None and hasattr(_mpu, "get_sequence_parallel_group"):
        return _mpu.get_sequence_parallel_group()
    if mesh_device is not None:
        return mesh_device.get_group(mesh_dim="sequence_parallel")
    return _groups.get("sp")


def _get_sequence_parallel_world_size():
    if _mpu is not None and hasattr(_mpu, "get_sequence_parallel_world_size"):
        return _mpu.get_sequence_parallel_world_size()
    if mesh_device is not None:
        return dist.get_world_size(_get_sequence_parallel_group())
    return len(_ranks["sp"]) if "sp" in _ranks else 1


def _get_sequence_parallel_rank():
    if _mpu is not None and hasattr(_mpu, "get_sequence_parallel_rank"):
        return _mpu.get_sequence_parallel_rank()
    if mesh_device is not None:
        return dist.get_rank(_get_sequence_parallel_group())
    return _ranks["sp"].index(_rank()) if "sp" in _ranks else 0


def _get_sequence_data_parallel_group():
    if _mpu is not None and hasattr(_mpu, "get_sequence_data_parallel_group"):
        return _mpu.get_sequence_data_parallel_group()
    if mesh_device is not None:
        return None  # flattened (dp, sp) mesh == WORLD
    return _groups.get("sdp", _get_data_parallel_group())


def _get_sequence_data_parallel_world_size():
    if _mpu is not None and hasattr(_mpu, "get_sequence_data_parallel_world_size"):
        return _mpu.get_sequence_data_parallel_world_size()
    if "sdp" in _ranks:
        return len(_ranks["sdp"])
    if mesh_device is not None:
        return _world()
    return _get_data_parallel_world_size()


def _get_sequence_data_parallel_rank():
    if _mpu is not None and hasattr(_mpu, "get_sequence_data_parallel_rank"):
        return _mpu.get_sequence_data_parallel_rank()
    if "sdp" in _ranks:
        return _ranks["sdp"].index(_rank())
    return _get_data_parallel_rank()


# ---- expert parallel -----------------------------------------------------------------------
def _ep_name(ep_size):
    return f"ep_size_{ep_size}"


def _create_expert_and_data_parallel(expert_parallel_size_, use_data_before_expert_parallel_=False):
    """Create EP and expert-DP groups inside every DP group (reference: groups.py:236)."""
    name = _ep_name(expert_parallel_size_)
    if f"ep:{name}" in _groups or f"ep:{name}" in _ranks:
        return
    world = _world()
    if _mpu is None and "dp" not in _ranks:
        dp_lists = rank_layout(world, _tp_size, _pp_size, 1)["dp"] if _tp_size * _pp_size > 1 else [list(range(world))]
    elif _mpu is not None:
        tp = _get_model_parallel_world_size()
        dp_lists = rank_layout(world, tp, 1, 1)["dp"]
    else:
        dp_lists = rank_layout(world, _tp_size, _pp_size, _sp_size)["sdp"]
    ep_all, edp_all = [], []
    for dp_ranks in dp_lists:
        ep, edp = expert_layout(dp_ranks, expert_parallel_size_, use_data_before_expert_parallel_)
        ep_all += ep
        edp_all += edp
    _register(f"ep:{name}", ep_all)
    _register(f"edp:{name}", edp_all)
    _expert_parallel_size[name] = expert_parallel_size_


_create_expert_data_and_model_parallel = _create_expert_and_data_parallel


def _get_max_expert_size():
    assert _expert_parallel_size, "no expert groups have been created"
    return max(_expert_parallel_size.values())


def _get_max_expert_size_name():
    return _ep_name(_get_max_expert_size())


def _get_max_expert_parallel_group():
    return _get_expert_parallel_group(_get_max_expert_size_name())


def _get_expert_parallel_group(group_name):
    return _groups.get(f"ep:{group_name}")


def _get_expert_parallel_ranks(world_size, tensor_parallel_size_, expert_parallel_size_, pipeline_parallel_size_=1,
                               use_data_before_expert_parallel_=False):
    """Rank lists of every expert-parallel and expert-data-parallel group of an E + M (+ P) + D layout (pure function;
    reference ``utils/groups.py:307``).  Tensor-parallel ranks are adjacent, data-parallel ranks of one tensor slice are
    ``tensor_parallel_size_`` apart; an expert-parallel group takes ``expert_parallel_size_`` consecutive data-parallel
    ranks (E + D) or, with ``use_data_before_expert_parallel_``, every ``dp / ep``-th one (D + E); the ranks that hold the
    same experts form the expert-data-parallel groups.

    world 16, tensor 2, expert 4 -> EP [[0, 2, 4, 6], [8, 10, 12, 14], [1, 3, 5, 7], [9, 11, 13, 15]],
    EDP [[0, 8], [2, 10], [4, 12], [6, 14], [1, 9], [3, 11], [5, 13], [7, 15]]."""
    tp, pp, ep = int(tensor_parallel_size_), int(pipeline_parallel_size_), int(expert_parallel_size_)
    assert world_size % (tp * pp) == 0, f"{world_size} is not divisible by {tp * pp}"
    dp_world = world_size // (tp * pp)
    assert dp_world % ep == 0, f"{dp_world} is not divisible by {ep}"
    ep_groups, edp_groups = [], []
    pp_stride = world_size // pp
    for stage_start in range(0, world_size, pp_stride):
        for t in range(tp):
            dp_ranks = list(range(stage_start + t, stage_start + pp_stride, tp))
            if use_data_before_expert_parallel_:
                stride = dp_world // ep
                groups_here = [dp_ranks[i::stride] for i in range(stride)]
            else:
                groups_here = [dp_ranks[i:i + ep] for i in range(0, dp_world, ep)]
            ep_groups.extend(groups_here)
            edp_groups.extend([list(col) for col in zip(*groups_here)])
    return ep_groups, edp_groups


def _expert_parallel_ranks_of(group_name):
    """Ranks of the named, already created expert-parallel group."""
    return _ranks.get(f"ep:{group_name}", [_rank()])


def _get_expert_parallel_group_dict():
    return {k[3:]: v for k, v in _groups.items() if k.startswith("ep:")}


def _get_expert_data_parallel_group(group_name):
    return _groups.get(f"edp:{group_name}")


def _get_expert_data_parallel_group_dict():
    return {k[4:]: v for k, v in _groups.items() if k.startswith("edp:")}


def _get_expert_parallel_world_size(group_name):
    return len(_ranks.get(f"ep:{group_name}", [0]))


def _get_expert_data_parallel_world_size(group_name):
    return len(_ranks.get(f"edp:{group_name}", [0]))


def _get_expert_parallel_rank(group_name):
    return _ranks.get(f"ep:{group_name}", [_rank()]).index(_rank())


def _get_expert_data_parallel_rank(group_name):
    return _ranks.get(f"edp:{group_name}", [_rank()]).index(_rank())


def _get_expert_model_parallel_world_size():
    return expert_tensor_parallel_world_size


# ---- ZeRO++ hpZ secondary partition groups ---------------------------------------------------
def _create_zero_param_parallel_group(group_size: int):
    """Intra-node sub-groups of ``group_size`` ranks holding the hpZ secondary shard."""
    if "hpz" in _ranks:
        return
    world = _world()
    assert world % group_size == 0, f"world {world} not divisible by hpz partition size {group_size}"
    _register("hpz", [list(range(i, i + group_size)) for i in range(0, world, group_size)])


def _zero_param_parallel_is_initialized():
    return "hpz" in _ranks


def _get_zero_param_intra_parallel_group():
    return _groups.get("hpz")


def _get_zero_param_intra_parallel_group_ranks():
    return _ranks.get("hpz")


def _get_zero_param_intra_parallel_group_world_size():
    return len(_ranks.get("hpz", [0]))


def _get_zero_param_intra_parallel_rank_in_mygroup():
    return _ranks.get("hpz", [_rank()]).index(_rank())


# ---- qgZ two-hop all-to-all groups -----------------------------------------------------------
def _get_local_all_to_all_group(local_size: Optional[int] = None):
    """Groups for hierarchical all-to-all: ``local_<n>`` (intra-node) and ``global_<i>`` (one
    rank per node).  Reference: groups.py:490."""
    import os
    if any(k.startswith("a2a:") for k in _ranks):
        return {k[4:]: v for k, v in _groups.items() if k.startswith("a2a:")}
    world = _world()
    local = local_size or int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("LOCAL_SIZE", world)))
    local = min(local, world)
    nodes = world // local
    me = _rank()
    for n in range(nodes):
        ranks = list(range(n * local, (n + 1) * local))
        g = dist.new_group(ranks) if dist.is_initialized() else None
        if me in ranks:
            _groups[f"a2a:local_{n}"] = g
            _ranks[f"a2a:local_{n}"] = ranks
    if nodes > 1:
        for i in range(local):
            ranks = [i + n * local for n in range(nodes)]
            g = dist.new_group(ranks) if dist.is_initialized() else None
            if me in ranks:
                _groups[f"a2a:global_{i}"] = g
                _ranks[f"a2a:global_{i}"] = ranks
    return {k[4:]: v for k, v in _groups.items() if k.startswith("a2a:")}


# ---- broadcast source helpers ---------------------------------------------------------------
def _get_broadcast_src_rank():
    ranks = _ranks.get("dp")
    return ranks[0] if ranks else 0


def _get_expert_broadcast_src_rank(group_name):
    return _ranks.get(f"edp:{group_name}", [0])[0]


def _get_sequence_data_parallel_src_rank():
    ranks = _ranks.get("sdp")
    return ranks[0] if ranks else _get_broadcast_src_rank()


def _init_tp_mesh_device(tensor_model_parallel_size=1, data_parallel_size=None):
    """Create a (dp, tp) mesh and register its groups (reference: groups.py:80)."""
    global mesh_device, _tp_size
    world = _world()
    dp = data_parallel_size or world // tensor_model_parallel_size
    _tp_size = tensor_model_parallel_size
    lay = rank_layout(world, tp=tensor_model_parallel_size)
    _register("tp", lay["tp"])
    _register("dp", lay["dp"])
    return _groups.get("tp"), _groups.get("dp")


def ranks_of(name: str):
    return _ranks.get(name)


# ---- public accessors + explicit TP overrides (reference ``utils/groups.py``) ------------------------------------------
_tp_rank_override = None
_tp_world_override = None


def get_data_parallel_group():
    return _get_data_parallel_group()


def get_data_parallel_world_size():
    return _get_data_parallel_world_size()


def get_data_parallel_rank():
    return _get_data_parallel_rank()


def get_model_parallel_group():
    return _get_model_parallel_group()


def get_model_parallel_world_size():
    return _tp_world_override if _tp_world_override is not None else _get_model_parallel_world_size()


def get_model_parallel_rank():
    return _tp_rank_override if _tp_rank_override is not None else _get_model_parallel_rank()


def set_tensor_model_parallel_world_size(world_size):
    global _tp_world_override
    _tp_world_override = world_size


def set_tensor_model_parallel_rank(rank):
    global _tp_rank_override
    _tp_rank_override = rank


def get_tensor_model_parallel_src_rank():
    """Global rank of the first member of this rank's tensor-parallel group."""
    from deepspeed_b200 import comm as dist
    tp = get_model_parallel_world_size()
    return (dist.get_rank() // tp) * tp if dist.is_initialized() else 0


def __getattr__(name):
    if name == "mpu":
        return _mpu
    raise AttributeError(name)
