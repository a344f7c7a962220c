from .loss_scaler import CreateLossScaler, DynamicLossScaler, LossScaler  # noqa: F401
