"""3-D (pp, tp, dp) checkpoint topology descriptor + contraction (reference ``checkpoint/reshape_3d_utils.py``)."""
from .constants import LAYER_FILE_PREFIX, MODEL_FILE_PREFIX
from .reshape_meg_2d import meg_2d_parallel_map, reshape_meg_2d_parallel
from .reshape_utils import get_files, get_files_with_prefix, get_zero_files, partition_data

PP_DIM, TP_DIM, DP_DIM = "PP", "TP", "DP"


class model_3d_desc:

    def __init__(self, pp_degree=1, tp_degree=1, dp_degree=1):
        self.pp_degree, self.tp_degree, self.dp_degree = pp_degree, tp_degree, dp_degree

    def _dims(self):
        return ((PP_DIM, self.pp_degree), (TP_DIM, self.tp_degree), (DP_DIM, self.dp_degree))

    def get_desc(self):
        return f"{PP_DIM},{TP_DIM},{DP_DIM} = ({self.pp_degree}, {self.tp_degree}, {self.dp_degree})"

    def world_size(self):
        return self.pp_degree * self.tp_degree * self.dp_degree

    def is_valid(self, pp_index, tp_index, dp_index):
        errs = [f"{name} indexing error: index {idx} >= degree {deg}"
                for idx, (name, deg) in zip((pp_index, tp_index, dp_index), self._dims()) if idx >= deg]
        return not errs, errs

    def can_reshape(self, target_3d_desc):
        errs = [f"Expansion reshape not supported - {name}: {mine} ---> {theirs}"
                for (name, mine), (_, theirs) in zip(self._dims(), target_3d_desc._dims()) if theirs > mine]
        return not errs, errs

    def reshape(self, target_3d_desc, verbose=False):
        """One ``meg_2d_parallel_map`` per target dp rank; cell (pp, tp) = source *file indices* to merge there."""
        ok, errs = self.can_reshape(target_3d_desc)
        assert ok, ",".join(errs)
        grid = reshape_meg_2d_parallel(self.pp_degree, self.tp_degree, target_3d_desc.pp_degree,
                                       target_3d_desc.tp_degree, verbose)
        flat = flatten_dp_dimension(grid, self.pp_degree * self.tp_degree, self.dp_degree)
        return unflatten_dp_dimension(flat, target_3d_desc.dp_degree)


def get_model_3d_descriptor(dir):
    """Infer (pp, tp, dp) from the file names of a checkpoint folder."""
    files = get_files(dir)
    n_zero = len(get_zero_files(dir))
    n_mp = len(get_files_with_prefix(files, MODEL_FILE_PREFIX))
    n_first_layer = len(get_files_with_prefix(files, f"{LAYER_FILE_PREFIX}01"))
    if n_first_layer > 0:  # pipeline checkpoints carry one layer_01 file per tp rank
        tp = n_first_layer
        pp = n_mp // tp
    else:
        tp, pp = n_mp, 1
    return model_3d_desc(pp, tp, max(1, n_zero // max(1, pp * tp)))


def flatten_dp_dimension(meg_2d_map, src_2d_size, dp_degree):
    """Expand each 2-D source rank r into its dp replicas r + k * (pp*tp)."""
    out = meg_2d_parallel_map(meg_2d_map.pp_degree, meg_2d_map.tp_degree)
    for p in range(meg_2d_map.pp_degree):
        for t in range(meg_2d_map.tp_degree):
            for r in meg_2d_map.get_data(p, t):
                out.add_data(p, t, [r + k * src_2d_size for k in range(dp_degree)])
    return out


def unflatten_dp_dimension(meg_2d_map, dp_degree):
    outs = [meg_2d_parallel_map(meg_2d_map.pp_degree, meg_2d_map.tp_degree) for _ in range(dp_degree)]
    for p in range(meg_2d_map.pp_degree):
        for t in range(meg_2d_map.tp_degree):
            for grid, ranks in zip(outs, partition_data(meg_2d_map.get_data(p, t), dp_degree)):
                grid.add_data(p, t, ranks)
    return outs
