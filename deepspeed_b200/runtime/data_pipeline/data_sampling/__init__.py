from .indexed_dataset import MMapIndexedDataset, MMapIndexedDatasetBuilder, make_dataset, make_builder  # noqa: F401
from .data_sampler import DeepSpeedDataSampler  # noqa: F401
from .data_analyzer import DataAnalyzer, DistributedDataAnalyzer  # noqa: F401
