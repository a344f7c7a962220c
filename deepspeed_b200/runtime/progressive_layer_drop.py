"""Progressive layer dropping (reference ``runtime/progressive_layer_drop.py``): keep probability
``theta(t) = (1 - theta) * exp(-gamma * t) + theta`` passed to the model's forward as kwargs."""
import numpy as np

from deepspeed_b200.utils.logging import log_dist


class ProgressiveLayerDrop:

    def __init__(self, theta=0.5, gamma=0.001):
        self.theta = theta
        self.gamma = gamma
        self.current_theta = 1.0
        log_dist(f"Enabled progressive layer dropping (theta = {self.theta})", ranks=[0])

    def get_state(self):
        return {"progressive_layer_drop": True, "pld_theta": self.get_theta()}

    def get_theta(self):
        return self.current_theta

    def update_state(self, global_step):
        self.current_theta = (1.0 - self.theta) * float(np.exp(-self.gamma * global_step)) + self.theta
