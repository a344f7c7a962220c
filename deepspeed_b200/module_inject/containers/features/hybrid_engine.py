"""Mixin: refresh the fused inference layer from the (trained) original layer before a generation phase
(reference ``containers/features/hybrid_engine.py``)."""


class HybridEngineContainer:

    def initialize_tensors(self, enable_training=False):
        super().initialize_tensors(enable_training=enable_training)
        self._hybrid = enable_training

    def refresh(self):
        """Re-read every tensor through the policy and copy into the fused layer (weights changed by an optimizer step)."""
        self.initialize_tensors(enable_training=True)
        self.apply_tensor_parallelism()
        self.copy_data_to_new_module()
