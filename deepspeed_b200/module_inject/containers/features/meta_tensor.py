"""Mixin: the model was built on the ``meta`` device; weights arrive from a checkpoint state dict instead of the layer
(reference ``containers/features/meta_tensor.py``)."""
import torch


class MetaTensorContainer:
    is_meta = False

    def initialize_tensors(self, enable_training=False):
        super().initialize_tensors(enable_training)
        self.is_meta = any(t is not None and t.is_meta for t in (self.qkvw, self.dense_w, self._h4h_w, self._4hh_w))

    def apply_tensor_parallelism(self, mp_replace=None):
        if self.is_meta:
            return  # nothing to slice yet: load_params() slices what it reads
        super().apply_tensor_parallelism(mp_replace)

    def copy_data_to_new_module(self):
        if self.is_meta:
            return
        super().copy_data_to_new_module()

    def load_params(self, module, sd, weight_quantizer=None, mp_replace=None, prefix=""):
        """Fill the fused layer from ``sd`` using the policy's ``param_names`` (qkv pieces, o, mlp, norms), applying the
        same TP slicing as the live-weight path."""
        names = self.policy.param_names(prefix)
        get = lambda n: None if n is None else (torch.cat([sd[x] for x in n], 0) if isinstance(n, (list, tuple)) else sd.get(n))
        self.qkvw, self.qkvb, self.dense_w, self.dense_b = (get(names[k]) for k in ("qkvw", "qkvb", "dense_w", "dense_b"))
        self._h4h_w, self._h4h_b, self._4hh_w, self._4hh_b = (get(names[k]) for k in ("h4h_w", "h4h_b", "4hh_w", "4hh_b"))
        self.attn_nw, self.attn_nb, self.input_nw, self.input_nb = (get(names[k]) for k in ("attn_nw", "attn_nb", "input_nw",
                                                                                           "input_nb"))
        self.is_meta = False
        super().apply_tensor_parallelism(mp_replace)
        super().copy_data_to_new_module()
