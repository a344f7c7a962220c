"""Data loading helpers (reference: ``runtime/dataloader.py:17 RepeatingLoader``, ``:41 DeepSpeedDataLoader``)."""
from torch.utils.data import DataLoader, RandomSampler
from torch.utils.data.distributed import DistributedSampler


class RepeatingLoader:
    """Wrap an iterable so it restarts instead of raising ``StopIteration``."""

    def __init__(self, loader):
        self.loader = loader
        self.data_iter = iter(self.loader)

    def __iter__(self):
        return self

    def __next__(self):
        try:
            return next(self.data_iter)
        except StopIteration:
            self.data_iter = iter(self.loader)
            return next(self.data_iter)


class DeepSpeedDataLoader:

    def __init__(self, dataset, batch_size, pin_memory, local_rank, tput_timer, collate_fn=None,
                 num_local_io_workers=None, data_sampler=None, data_parallel_world_size=None, data_parallel_rank=None,
                 dataloader_drop_last=False, deepspeed_dataloader_config=None):
        self.deepspeed_dataloader_config = deepspeed_dataloader_config or {}
        self.tput_timer = tput_timer
        self.batch_size = batch_size
        self.curriculum_learning_enabled = bool(self.deepspeed_dataloader_config.get("curriculum_learning_enabled"))
        if self.curriculum_learning_enabled:
            from deepspeed_b200.runtime.data_pipeline.data_sampling.data_sampler import DeepSpeedDataSampler
            data_sampler = DeepSpeedDataSampler(self.deepspeed_dataloader_config["data_efficiency"], len(dataset),
                                                batch_size, data_parallel_rank, data_parallel_world_size,
                                                self.deepspeed_dataloader_config.get("data_parallel_group"),
                                                self.deepspeed_dataloader_config.get("gradient_accumulation_steps", 1),
                                                self.deepspeed_dataloader_config.get("global_rank", 0),
                                                drop_last=dataloader_drop_last)
            self.device_count = 1
            self.len = len(data_sampler) if hasattr(data_sampler, "__len__") else None
        elif local_rank >= 0:
            if data_sampler is None:
                data_sampler = DistributedSampler(dataset=dataset, num_replicas=data_parallel_world_size,
                                                  rank=data_parallel_rank)
            self.device_count = 1
        else:
            if data_sampler is None:
                data_sampler = RandomSampler(dataset)
            self.device_count = 1
        if num_local_io_workers is None:
            num_local_io_workers = 0
        self.num_local_io_workers = num_local_io_workers
        self.data_sampler = data_sampler
        self.dataset = dataset
        self.collate_fn = collate_fn
        self.pin_memory = pin_memory
        self.dataloader_drop_last = dataloader_drop_last
        self.data = None
        self.post_process_func = None
        if not self.curriculum_learning_enabled:
            from math import ceil
            n = len(self.data_sampler)
            self.len = n // batch_size if dataloader_drop_last else ceil(n / batch_size)

    def __iter__(self):
        self._create_dataloader()
        return self

    def __len__(self):
        return self.len

    def __next__(self):
        if self.tput_timer:
            self.tput_timer.start()
        if self.curriculum_learning_enabled:
            data = next(self.data_iterator)
            if self.post_process_func is not None:
                data = self.post_process_func(data, self.data_sampler.state_dict())
            return data
        return next(self.data)

    def _create_dataloader(self):
        import torch
        pin = self.pin_memory and torch.cuda.is_available()
        if self.curriculum_learning_enabled:
            self.dataloader = DataLoader(self.dataset, pin_memory=pin, batch_sampler=self.data_sampler,
                                         num_workers=self.num_local_io_workers,
                                         **({"collate_fn": self.collate_fn} if self.collate_fn else {}))
            self.data_iterator = iter(self.dataloader)
            return self.dataloader
        kw = dict(batch_size=self.batch_size, pin_memory=pin, sampler=self.data_sampler,
                  num_workers=self.num_local_io_workers, drop_last=self.dataloader_drop_last)
        if self.collate_fn is not None:
            kw["collate_fn"] = self.collate_fn
        self.dataloader = DataLoader(self.dataset, **kw)
        self.data = (x for x in self.dataloader)
        return self.dataloader
