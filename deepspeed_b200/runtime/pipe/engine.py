"""``PipelineEngine``: executes a pipeline schedule over a ``PipelineModule``.

Parity target: reference ``runtime/pipe/engine.py:61`` (``train_batch :338``, ``eval_batch :427``, instruction
handlers ``:709-1216``, ``_exec_schedule :1408``).  The ZeRO / mixed-precision machinery is inherited from
:class:`DeepSpeedEngine`: each micro-batch backward feeds the sharded optimizer exactly like gradient
accumulation (``gas == micro_batches``), so data-parallel reduction overlaps the pipeline.
"""
import torch

from deepspeed_b200 import comm as dist
from deepspeed_b200.runtime.dataloader import RepeatingLoader
from deepspeed_b200.runtime.engine import DeepSpeedEngine, _to_device
from deepspeed_b200.utils.logging import log_dist
from . import p2p, schedule
from .module import PipelineError, PipelineModule


class PipelineEngine(DeepSpeedEngine):

    def __init__(self, has_bool_tensors=False, *super_args, **super_kwargs):
        super().__init__(*super_args, **super_kwargs)
        assert isinstance(self.module, PipelineModule), "model must base PipelineModule"
        assert self.zero_optimization_stage() < 3, "ZeRO-3 parameter partitioning is incompatible with pipeline parallelism"
        self.enable_backward_allreduce = False
        self.has_bool_tensors = has_bool_tensors
        self.pipeline_enable_backward_allreduce = True
        self.grid = self.module._grid
        self.global_rank = self.grid.get_global_rank()
        self.micro_batch_size = self.train_micro_batch_size_per_gpu()
        self.micro_batches = self.gradient_accumulation_steps()
        self.num_stages = self.grid.pipe_parallel_size
        self.stage_id = self.grid.get_stage_id()
        self.prev_stage, self.next_stage = self.stage_id - 1, self.stage_id + 1
        self.data_iterator = None
        self.batch_fn = None
        self.is_pipe_parallel = self.grid.pipe_parallel_size > 1
        self.is_data_parallel = self.grid.data_parallel_size > 1
        self.dynamic_shape = self.module.dynamic_shape
        p2p.init_process_groups(self.grid)
        self.pipe_buffers = {"inputs": {}, "labels": {}, "outputs": {}}
        self.loss = torch.tensor(0.0, device=self.device)
        self.total_loss = None
        self.agg_loss = torch.tensor(0.0, device=self.device)
        self.loss_model = self.module.loss_fn
        if self.optimizer is not None and hasattr(self.optimizer, "disable_fused_in_backward") \
                and list(self.module.get_tied_weights_and_groups()):
            # tied copies on different stages must see the SUM of their gradients before the update: the step cannot
            # be fused into backward (reference pipe/engine.py:284 _exec_reduce_tied_grads precedes the step)
            self.optimizer.disable_fused_in_backward("tied weights across pipeline stages")
        if self.training_data is not None:
            self._build_data_iter(self.training_data)
        log_dist(f"PipelineEngine: stages={self.num_stages} stage_id={self.stage_id} micro_batches={self.micro_batches} "
                 f"dp={self.grid.data_parallel_size}", ranks=[0])

    # ---- data -----------------------------------------------------------------------------------------------
    def _build_data_iter(self, dataset):
        sampler = torch.utils.data.distributed.DistributedSampler(dataset, num_replicas=self.dp_world_size,
                                                                  rank=self.mpu.get_data_parallel_rank(), shuffle=False)
        loader = self.deepspeed_io(dataset, data_sampler=sampler)
        self.set_dataloader(RepeatingLoader(loader))

    def set_dataloader(self, loader):
        if self.is_first_stage() or self.is_last_stage():
            self.training_dataloader = loader
            self.data_iterator = iter(self.training_dataloader)

    def set_dataiterator(self, iterator):
        if self.is_first_stage() or self.is_last_stage():
            self.training_dataloader = None
            self.data_iterator = iterator

    def set_batch_fn(self, fn):
        self.batch_fn = fn

    def is_first_stage(self):
        return self.stage_id == 0

    def is_last_stage(self):
        return self.stage_id == self.num_stages - 1

    def is_gradient_accumulation_boundary(self):
        return self._force_grad_boundary

    _force_grad_boundary = False

    # ---- public API -----------------------------------------------------------------------------------------------
    def train_batch(self, data_iter=None):
        if not torch._C.is_grad_enabled():
            raise RuntimeError("train_batch() requires gradients enabled. Use eval_batch() instead.")
        if data_iter is not None:
            self.set_dataiterator(data_iter)
        self.module.train()
        self.total_loss = None
        self._compute_loss = True
        sched = schedule.TrainSchedule(micro_batches=self.micro_batches, stages=self.num_stages, stage_id=self.stage_id)
        self._exec_schedule(sched)
        self.agg_train_loss = self._aggregate_total_loss()
        return self.agg_train_loss

    def eval_batch(self, data_iter, return_logits=False, compute_loss=True, reduce_output="avg", bcast_loss=True,
                   num_micro_batches=None):
        self.module.eval()
        self.eval_return_logits = return_logits
        self._compute_loss = compute_loss
        train_iterator = self.data_iterator
        self.set_dataiterator(data_iter)
        micro = self.micro_batches if num_micro_batches is None else num_micro_batches
        sched = schedule.InferenceSchedule(micro_batches=micro, stages=self.num_stages, stage_id=self.stage_id)
        self.fwd_outputs = []
        self.total_loss = None
        with torch.no_grad():
            self._exec_schedule(sched)
        if self.is_last_stage():
            out = self._reduce_outputs(self.fwd_outputs, reduce=reduce_output, micro_batches=micro)
        else:
            out = None
        if compute_loss and bcast_loss:
            out = self._bcast_pipe_scalar(out if out is not None else torch.tensor(0.0, device=self.device))
        self.set_dataiterator(train_iterator)
        if return_logits:
            return out, getattr(self, "_last_logits", None)
        return out

    def _reduce_outputs(self, outputs, reduce="avg", micro_batches=None):
        if reduce is None or not outputs:
            return outputs
        if torch.is_tensor(outputs[0]):
            total = sum(outputs)
        else:
            total = [sum(o[i] for o in outputs) for i in range(len(outputs[0]))]
        if reduce == "avg":
            total = total / (micro_batches or len(outputs)) if torch.is_tensor(total) else [t / micro_batches for t in total]
        if self.is_data_parallel and torch.is_tensor(total):
            dist.all_reduce(total, group=self.mpu.get_data_parallel_group())
            total = total / self.dp_world_size
        return total

    def _bcast_pipe_scalar(self, data, src_rank=None, dtype=torch.float32):
        if src_rank is None:
            src_rank = self.grid.stage_to_global(self.num_stages - 1)
        t = data.clone().detach().to(dtype).to(self.device) if self.global_rank == src_rank else \
            torch.tensor([0.0], dtype=dtype, device=self.device).reshape(())
        t = t.reshape(1)
        dist.broadcast(t, src=src_rank, group=self.mpu.get_pipe_parallel_group())
        return t.reshape(())

    def _aggregate_total_loss(self):
        if self.is_last_stage():
            loss = self.total_loss / self.micro_batches if self.total_loss is not None else torch.tensor(0.0, device=self.device)
            loss = loss.detach().clone().float()
            if self.is_data_parallel:
                dist.all_reduce(loss, group=self.mpu.get_data_parallel_group())
                loss = loss / self.dp_world_size
        else:
            loss = torch.tensor(0.0, device=self.device)
        if self.is_pipe_parallel:
            loss = self._bcast_pipe_scalar(loss)
        return loss

    # ---- instruction handlers ---------------------------------------------------------------------------------------
    def _next_batch(self):
        batch = next(self.data_iterator)
        if self.batch_fn:
            batch = self.batch_fn(batch)
        return _to_device(batch, self.device)

    def _exec_load_micro_batch(self, buffer_id):
        batch = self._next_batch()
        if self.is_first_stage():
            x = batch[0]
            if torch.is_tensor(x):
                x = x.clone().detach()
                x.requires_grad_(x.is_floating_point())
            else:
                x = tuple(t.clone().detach().requires_grad_(t.is_floating_point()) for t in x)
            self.pipe_buffers["inputs"][buffer_id] = x
        if self.is_last_stage():
            self.pipe_buffers["labels"][buffer_id] = batch[1]

    def _exec_forward_pass(self, buffer_id):
        inputs = self.pipe_buffers["inputs"][buffer_id]
        self.module.micro_offset = 0
        outputs = DeepSpeedEngine.forward(self, inputs)
        self.pipe_buffers["outputs"][buffer_id] = outputs
        if self.is_last_stage():
            if self._compute_loss and self.loss_model is not None:
                self.loss = self.loss_model(outputs, self.pipe_buffers["labels"][buffer_id])
            else:
                self.loss = outputs
                self._last_logits = outputs
            if torch.is_tensor(self.loss):
                if getattr(self, "fwd_outputs", None) is not None and not self.module.training:
                    self.fwd_outputs.append(self.loss.detach())
                self.total_loss = self.loss.detach().clone() if self.total_loss is None else self.total_loss + self.loss.detach()

    def _exec_backward_pass(self, buffer_id):
        zo = self.optimizer
        zo._in_backward = True
        if self.is_last_stage():
            scaled = self.loss / self.micro_batches
            zo.loss_scaler.backward(scaled.float())
        else:
            outputs = self.pipe_buffers["outputs"][buffer_id]
            grads = self.grad_layer
            if torch.is_tensor(outputs):
                torch.autograd.backward((outputs, ), (grads, ))
            else:
                outs = [t for t in outputs if t.is_floating_point() and t.requires_grad]
                gs = [g for t, g in zip(outputs, grads if isinstance(grads, (tuple, list)) else (grads, ))
                      if t.is_floating_point() and t.requires_grad]
                torch.autograd.backward(tuple(outs), tuple(gs))
        zo.end_backward()
        self.pipe_buffers["outputs"][buffer_id] = None
        self.grad_layer = None

    def _exec_send_activations(self, buffer_id):
        p2p.send_obj(self.pipe_buffers["outputs"][buffer_id], self.next_stage, "act", self.dynamic_shape)

    def _exec_recv_activations(self, buffer_id):
        x = p2p.recv_obj(self.prev_stage, "act", self.dynamic_shape)
        if torch.is_tensor(x):
            if x.is_floating_point():
                x.requires_grad_(True)
        else:
            for t in x:
                if t.is_floating_point():
                    t.requires_grad_(True)
        self.pipe_buffers["inputs"][buffer_id] = x

    def _exec_send_grads(self, buffer_id):
        inputs = self.pipe_buffers["inputs"][buffer_id]
        if torch.is_tensor(inputs):
            g = inputs.grad
        else:
            g = tuple(t.grad if t.grad is not None else torch.zeros_like(t) for t in inputs if t.is_floating_point())
        p2p.send_obj(g, self.prev_stage, "grad", self.dynamic_shape)
        self.pipe_buffers["inputs"][buffer_id] = None

    def _exec_recv_grads(self, buffer_id):
        self.grad_layer = p2p.recv_obj(self.next_stage, "grad", self.dynamic_shape)

    def _exec_send_act_recv_grad(self, send_buffer, recv_buffer):
        self.grad_layer = p2p.send_recv(self.pipe_buffers["outputs"][send_buffer], self.next_stage, "act", "grad",
                                        self.dynamic_shape)

    def _exec_send_grad_recv_act(self, send_buffer, recv_buffer):
        inputs = self.pipe_buffers["inputs"][send_buffer]
        if torch.is_tensor(inputs):
            g = inputs.grad
        else:
            g = tuple(t.grad if t.grad is not None else torch.zeros_like(t) for t in inputs if t.is_floating_point())
        self.pipe_buffers["inputs"][send_buffer] = None
        x = p2p.send_recv(g, self.prev_stage, "grad", "act", self.dynamic_shape)
        if torch.is_tensor(x):
            if x.is_floating_point():
                x.requires_grad_(True)
        else:
            for t in x:
                if t.is_floating_point():
                    t.requires_grad_(True)
        self.pipe_buffers["inputs"][recv_buffer] = x

    def _exec_reduce_tied_grads(self):
        # tied-weight gradients are summed across the owning stages inside the flat gradient arena path:
        # all-reduce the accumulated shard of each tied parameter over its tie group
        zo = self.optimizer
        for w, group in self.module.get_tied_weights_and_groups():
            if zo.grad_arena is None:
                continue
            rt, s = zo.unit_of_param[id(w)], zo.slot_of_param[id(w)]
            from deepspeed_b200.runtime.zero.units import param_fragments
            for (r, p0, a0, ln) in param_fragments(rt.u, s, zo.shard_world):
                if r == zo.shard_rank:
                    dist.all_reduce(zo.grad_arena[a0:a0 + ln], group=group)

    def _exec_reduce_grads(self):
        pass  # data-parallel reduction already happened unit-by-unit inside each micro-batch backward

    def _exec_optimizer_step(self, lr_kwargs=None):
        self._force_grad_boundary = True
        self._take_model_step(lr_kwargs)
        self._force_grad_boundary = False
        self.micro_steps += self.micro_batches

    _INSTRUCTION_MAP = {
        schedule.OptimizerStep: _exec_optimizer_step,
        schedule.ReduceGrads: _exec_reduce_grads,
        schedule.ReduceTiedGrads: _exec_reduce_tied_grads,
        schedule.LoadMicroBatch: _exec_load_micro_batch,
        schedule.ForwardPass: _exec_forward_pass,
        schedule.BackwardPass: _exec_backward_pass,
        schedule.SendActivation: _exec_send_activations,
        schedule.RecvActivation: _exec_recv_activations,
        schedule.SendGrad: _exec_send_grads,
        schedule.RecvGrad: _exec_recv_grads,
        schedule.SendActivationRecvGrad: _exec_send_act_recv_grad,
        schedule.SendGradRecvActivation: _exec_send_grad_recv_act,
    }

    def _exec_schedule(self, pipe_schedule):
        self.fwd_outputs = [] if not self.module.training else None
        flat = [cmd for step_cmds in pipe_schedule for cmd in step_cmds]
        for cmd in schedule.fuse_exchanges(flat):
            fn = self._INSTRUCTION_MAP.get(type(cmd))
            if fn is None:
                raise RuntimeError(f"{self.__class__.__name__} does not understand instruction {cmd!r}")
            fn(self, **cmd.kwargs)

    # ---- guards (reference: these entry points are illegal on a pipeline engine) ---------------------------------------
    def forward(self, *args, **kwargs):
        raise PipelineError("Only train_batch() is accessible in pipeline mode.")

    def backward(self, *args, **kwargs):
        raise PipelineError("Only train_batch() is accessible in pipeline mode.")

    def step(self, *args, **kwargs):
        raise PipelineError("Only train_batch() is accessible in pipeline mode.")

    # ---- checkpoint: per-layer module files -------------------------------------------------------------------------------
    def module_state_dict(self, exclude_frozen_parameters=False, **kw):
        assert getattr(self, "_curr_ckpt_path", None) is not None, "PipelineEngine expects module_state_dict() to be called from save_checkpoint()"
        self.module.save_state_dict(self._curr_ckpt_path, checkpoint_engine=self.checkpoint_engine,
                                    exclude_frozen_params=exclude_frozen_parameters)
        return None

    def load_module_state_dict(self, checkpoint, strict=True, custom_load_fn=None, fetch_z3_params=False):
        sd = checkpoint.get("module") if isinstance(checkpoint, dict) else None
        if sd is not None and not isinstance(sd, str):
            super().load_module_state_dict(checkpoint, strict)
            return
        self.module.load_state_dir(load_dir=self._curr_ckpt_path, strict=strict, checkpoint_engine=self.checkpoint_engine)

    def save_checkpoint(self, save_dir, tag=None, client_state=None, save_latest=True, exclude_frozen_parameters=False):
        import os
        tag = tag if tag is not None else f"global_step{self.global_steps}"
        self._curr_ckpt_path = os.path.join(save_dir, str(tag))
        return super().save_checkpoint(save_dir, tag, client_state, save_latest, exclude_frozen_parameters)

    def load_checkpoint(self, load_dir, tag=None, **kw):
        import os
        if tag is None and os.path.isfile(os.path.join(load_dir, "latest")):
            tag = open(os.path.join(load_dir, "latest")).read().strip()
        self._curr_ckpt_path = os.path.join(load_dir, str(tag))
        return super().load_checkpoint(load_dir, tag, **kw)

    # ---- small public knobs of the reference engine ---------------------------------------------------------------------
    has_attention_mask = False
    agg_additional_losses = None

    def set_has_attention_mask(self, value):
        """Declare that the last tensor of a stage's tuple output is an attention mask (never differentiated)."""
        assert isinstance(value, bool)
        self.has_attention_mask = value

    def reset_activation_shape(self):
        """Forget the negotiated activation / gradient shapes: the next send re-transmits the shape header (call when a
        curriculum changes the sequence length)."""
        p2p._meta_cache.clear()
        self.grad_layer = None

    def log_for_device(self, *msg):
        print(f"RANK={dist.get_rank()} PIPE-ID={self.stage_id} DATA-ID={self.grid.data_parallel_id} ::", *msg, flush=True)

    def tput_log(self, *msg):
        if self.global_rank == 0 and self.global_steps % self.steps_per_print() == 0:
            print(*msg)

    def mem_status(self, msg, print_rank=-1, reset_max=False):
        if print_rank not in (-1, self.global_rank) or not torch.cuda.is_available():
            return
        a, r = torch.cuda.memory_allocated() / 2**30, torch.cuda.memory_reserved() / 2**30
        m = torch.cuda.max_memory_allocated() / 2**30
        if reset_max:
            torch.cuda.reset_peak_memory_stats()
        print(f"RANK={self.global_rank} STAGE={self.stage_id} MEMSTATS {msg} allocated={a:.2f}GB reserved={r:.2f}GB peak={m:.2f}GB")

    def get_additional_losses(self):
        """Extra named losses the model reported on the last stage (``PipelineModule.get_additional_losses``)."""
        return self.agg_additional_losses



def is_even(number):
    return number % 2 == 0
