"""Pipeline x tensor parallel rank-grid contraction (reference ``checkpoint/reshape_meg_2d.py``).

A grid cell (pp, tp) of the *target* topology lists the source ranks whose shards it must merge.  Contracting tp merges
neighbouring tp columns inside a pipeline stage; contracting pp merges neighbouring stages inside a tp column.
"""
import numpy as np

from .reshape_utils import partition_data


class meg_2d_parallel_map:

    def __init__(self, pp_degree, tp_degree):
        self.pp_degree, self.tp_degree = pp_degree, tp_degree
        self.map = {}

    @staticmethod
    def _make_key(i, j):
        return f"{i},{j}"

    def _validate_indices(self, pp_index, tp_index):
        assert pp_index is None or 0 <= pp_index < self.pp_degree
        assert tp_index is None or 0 <= tp_index < self.tp_degree

    def simple_init(self):
        self.map = {self._make_key(p, t): [p * self.tp_degree + t] for p in range(self.pp_degree)
                    for t in range(self.tp_degree)}

    def add_data(self, pp_index, tp_index, data):
        self._validate_indices(pp_index, tp_index)
        assert isinstance(data, list)
        self.map.setdefault(self._make_key(pp_index, tp_index), []).extend(data)

    def get_data(self, pp_index=None, tp_index=None):
        self._validate_indices(pp_index, tp_index)
        pps = range(self.pp_degree) if pp_index is None else (pp_index, )
        tps = range(self.tp_degree) if tp_index is None else (tp_index, )
        return [r for p in pps for t in tps for r in self.map[self._make_key(p, t)]]

    def print_data(self, tag):
        print(tag)
        for k, v in self.map.items():
            print(f"{k} = {v}")


def _reshape_tp_dimension(old, new_tp_degree):
    new = meg_2d_parallel_map(old.pp_degree, new_tp_degree)
    for p in range(old.pp_degree):
        for t, ranks in enumerate(partition_data(old.get_data(pp_index=p), new_tp_degree)):
            new.add_data(p, t, ranks)
    return new


def _reshape_pp_dimension(old, new_pp_degree):
    new = meg_2d_parallel_map(new_pp_degree, old.tp_degree)
    for t in range(old.tp_degree):
        for p, ranks in enumerate(partition_data(old.get_data(tp_index=t), new_pp_degree)):
            new.add_data(p, t, ranks)
    return new


def reshape_meg_2d_parallel(old_pp_degree, old_tp_degree, new_pp_degree, new_tp_degree, verbose=False):
    assert new_pp_degree <= old_pp_degree and new_tp_degree <= old_tp_degree, "only contraction is supported"
    grid = meg_2d_parallel_map(old_pp_degree, old_tp_degree)
    grid.simple_init()
    if verbose:
        grid.print_data("original_2d_map:")
    if new_tp_degree != old_tp_degree:
        grid = _reshape_tp_dimension(grid, new_tp_degree)
        if verbose:
            grid.print_data("after_tp_reshape:")
    if new_pp_degree != old_pp_degree:
        grid = _reshape_pp_dimension(grid, new_pp_degree)
    if verbose:
        grid.print_data("final_2d_map:")
    return grid


def get_mpu_ranks(tp_size=1, pp_size=1, dp_size=1, virtual_pp_size=None):
    """Megatron rank layout (tp fastest, then dp, then pp): returns (tp groups, pp groups, dp groups)."""
    world = tp_size * pp_size * dp_size
    grid = np.arange(world).reshape(pp_size, dp_size, tp_size)
    tp_groups = [grid[p, d, :].tolist() for p in range(pp_size) for d in range(dp_size)]
    dp_groups = [grid[p, :, t].tolist() for p in range(pp_size) for t in range(tp_size)]
    pp_groups = [grid[:, d, t].tolist() for d in range(dp_size) for t in range(tp_size)]
    return tp_groups, pp_groups, dp_groups


def reshape(src, tgt):
    """Print the tp-then-pp contraction plan for ``[tp, pp, dp]`` source -> target (debug helper)."""
    (tp_s, pp_s, dp_s), (tp_t, pp_t, _) = src, tgt
    tp1, _, _ = get_mpu_ranks(tp_s, pp_s, dp_s)
    tp2, pp2, _ = get_mpu_ranks(tp_t, pp_s, dp_s)
    _, pp3, _ = get_mpu_ranks(tp_t, pp_t, dp_s)
    for a, b in zip(tp1, tp2):
        print(f"TP {a} => {b}")
    for a, b in zip(pp2, pp3):
        print(f"PP {a} => {b}")
