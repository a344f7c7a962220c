"""Experiment monitors: TensorBoard / W&B / CSV / Comet fan-out.

Parity target: reference ``monitor/monitor.py:30 MonitorMaster`` + ``tensorboard.py``, ``wandb.py``,
``csv_monitor.py``, ``comet.py``.  Events are ``(name, value, step)`` tuples; only rank 0 writes.
Back-ends whose package is missing log a warning once and disable themselves.
"""
import csv
import os
from typing import List, Tuple

from deepspeed_b200.utils.logging import logger

Event = Tuple[str, float, int]


def _rank():
    try:
        import torch.distributed as dist
        return dist.get_rank() if dist.is_initialized() else int(os.environ.get("RANK", 0))
    except Exception:
        return 0


class Monitor:

    def __init__(self, config):
        self.config = config
        self.enabled = bool(getattr(config, "enabled", False))

    def write_events(self, event_list: List[Event]):
        raise NotImplementedError


class TensorBoardMonitor(Monitor):

    def __init__(self, config):
        super().__init__(config)
        self.summary_writer = None
        self.output_path, self.job_name = config.output_path, config.job_name
        if self.enabled and _rank() == 0:
            self.get_summary_writer()

    def get_summary_writer(self, base=os.path.join(os.path.expanduser("~"), "tensorboard")):
        """Create (once) the ``SummaryWriter`` under ``<output_path or base>/<job_name>``; rank 0 only."""
        if self.summary_writer is None and self.enabled and _rank() == 0:
            try:
                from torch.utils.tensorboard import SummaryWriter
                log_dir = os.path.join(self.output_path or base, self.job_name)
                os.makedirs(log_dir, exist_ok=True)
                self.summary_writer = SummaryWriter(log_dir=log_dir)
            except Exception as e:
                logger.warning(f"tensorboard monitor disabled: {e}")
                self.enabled = False
        return self.summary_writer

    def flush(self):
        if self.enabled and self.summary_writer is not None and _rank() == 0:
            self.summary_writer.flush()

    def write_events(self, event_list, flush=True):
        if self.summary_writer is None:
            return
        for name, value, step in event_list:
            self.summary_writer.add_scalar(name, value, step)
        if flush:
            self.summary_writer.flush()


class WandbMonitor(Monitor):

    def __init__(self, config):
        super().__init__(config)
        self._wandb = None
        if self.enabled and _rank() == 0:
            try:
                import wandb
                wandb.init(project=config.project, group=config.group, entity=config.team)
                self._wandb = wandb
            except Exception as e:
                logger.warning(f"wandb monitor disabled: {e}")
                self.enabled = False

    def write_events(self, event_list):
        if self._wandb is None:
            return
        for name, value, step in event_list:
            self._wandb.log({name: value}, step=step)


class csvMonitor(Monitor):

    def __init__(self, config):
        super().__init__(config)
        self.filenames = {}
        self.log_dir = None
        if self.enabled and _rank() == 0:
            base = config.output_path or os.path.join(os.path.expanduser("~"), "csv_monitor")
            self.log_dir = os.path.join(base, config.job_name)
            os.makedirs(self.log_dir, exist_ok=True)

    def write_events(self, event_list):
        if self.log_dir is None:
            return
        for name, value, step in event_list:
            # "Train/Samples/lr" -> Train_Samples_lr.csv
            header = name.split("/")[-1]
            fname = os.path.join(self.log_dir, name.replace("/", "_") + ".csv")
            new = fname not in self.filenames and not os.path.exists(fname)
            self.filenames[fname] = True
            with open(fname, "a+", newline="") as f:
                w = csv.writer(f)
                if new:
                    w.writerow(["step", header])
                w.writerow([step, value])


class CometMonitor(Monitor):

    def __init__(self, config):
        super().__init__(config)
        self.experiment = None
        self.samples_log_interval = getattr(config, "samples_log_interval", 100)
        if self.enabled and _rank() == 0:
            try:
                import comet_ml
                self.experiment = comet_ml.start(api_key=config.api_key, project=config.project,
                                                 workspace=config.workspace, experiment_key=config.experiment_key,
                                                 mode=config.mode, online=config.online)
                if config.experiment_name:
                    self.experiment.set_name(config.experiment_name)
            except Exception as e:
                logger.warning(f"comet monitor disabled: {e}")
                self.enabled = False

    @property
    def samples_log_interval(self):
        return getattr(self, "_samples_log_interval", 1)

    def write_events(self, event_list):
        if self.experiment is None:
            return
        for name, value, step in event_list:
            self.experiment.log_metric(name, value, step=step)


class MonitorMaster(Monitor):

    def __init__(self, monitor_config):
        self.enabled = bool(monitor_config.enabled)
        self.tb_monitor = self.wandb_monitor = self.csv_monitor = self.comet_monitor = None
        if _rank() != 0:
            return
        if monitor_config.tensorboard.enabled:
            self.tb_monitor = TensorBoardMonitor(monitor_config.tensorboard)
        if monitor_config.wandb.enabled:
            self.wandb_monitor = WandbMonitor(monitor_config.wandb)
        if monitor_config.csv_monitor.enabled:
            self.csv_monitor = csvMonitor(monitor_config.csv_monitor)
        if monitor_config.comet.enabled:
            self.comet_monitor = CometMonitor(monitor_config.comet)

    def write_events(self, event_list):
        if _rank() != 0:
            return
        for m in (self.tb_monitor, self.wandb_monitor, self.csv_monitor, self.comet_monitor):
            if m is not None and m.enabled:
                m.write_events(event_list)
