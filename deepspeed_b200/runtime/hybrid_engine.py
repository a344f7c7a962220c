"""Hybrid engine: one set of weights, two execution modes — ZeRO training and fused-kernel generation (RLHF).

Parity target: reference ``runtime/hybrid_engine.py:32 DeepSpeedHybridEngine`` (``generate :168``, LoRA
``fuse_lora_weight :128`` / ``unfuse_lora_weight :139``, ZeRO-3 gather around generation, inference containers).
Design here: ``generate`` gathers the (possibly ZeRO-3 sharded) parameters once, re-packs them into the ragged
fused-kernel model of ``inference/v2`` — only when an optimizer step changed them since the last pack — and runs
the same CUDA-graphed decode loop the inference engine uses.  The packed copy and KV cache are dropped after
generation when ``release_inference_cache`` is set, so training gets its HBM back.
"""
import time

import torch

from deepspeed_b200 import comm as dist
from deepspeed_b200.runtime.engine import DeepSpeedEngine
from deepspeed_b200.utils.logging import log_dist


class DeepSpeedHybridEngine(DeepSpeedEngine):
    inference_mp_group = None

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        hc = self._config.hybrid_engine
        self._hybrid_cfg = hc
        self._ragged = None
        self._packed_at_step = -1
        self._uid = 0
        self._in_eval = False
        self._lora_fused = False
        self._t_gather = self._t_generate = self._t_train = 0.0
        self._iters = 0
        self._total_batch_size = None
        self._generate_latency = self._training_latency = 0.0
        log_dist(f"DeepSpeedHybridEngine: max_out_tokens={hc.max_out_tokens} inference_tp_size={hc.inference_tp_size} "
                 f"release_inference_cache={hc.release_inference_cache}", ranks=[0])

    # ---- LoRA: W <- W + scale * B A before generation, undone afterwards ---------------------------------------
    def _lora_modules(self):
        for m in self.module.modules():
            if all(hasattr(m, a) for a in ("lora_right_weight", "lora_left_weight", "lora_scaling")):
                yield m, m.weight, m.lora_right_weight, m.lora_left_weight, m.lora_scaling
            elif all(hasattr(m, a) for a in ("lora_weight_1", "lora_weight_2", "lora_scaling_factor")) \
                    and not getattr(m, "disabled", False) and m.weight.is_floating_point():
                # deepspeed_b200.linear.LoRAOptimizedLinear: weight [out, in], A=[r,in], B=[out,r]
                yield m, m.weight, m.lora_weight_1.weight.t(), m.lora_weight_2.weight.t(), m.lora_scaling_factor

    @torch.no_grad()
    def fuse_lora_weight(self):
        if self._lora_fused:
            return
        for _, w, right, left, scale in self._lora_modules():
            w.data += scale * torch.matmul(right.to(w.dtype), left.to(w.dtype)).t()
        self._lora_fused = True

    @torch.no_grad()
    def unfuse_lora_weight(self):
        if not self._lora_fused:
            return
        for _, w, right, left, scale in self._lora_modules():
            w.data -= scale * torch.matmul(right.to(w.dtype), left.to(w.dtype)).t()
        self._lora_fused = False

    def unfuse_lora_weight_non_pinned(self):
        self.unfuse_lora_weight()

    # ---- packing -------------------------------------------------------------------------------------------------
    def _build_ragged(self):
        from deepspeed_b200.inference.v2.config_v2 import RaggedInferenceEngineConfig
        from deepspeed_b200.inference.v2.engine_v2 import InferenceEngineV2
        from deepspeed_b200.inference.v2.model_implementations import (RaggedTransformer, arch_from_hf_config, load_hf_weights)
        from deepspeed_b200.inference.v2.engine_factory import build_engine_from_model
        hc = self._hybrid_cfg
        max_ctx = max(int(hc.max_out_tokens), 64)
        ec = RaggedInferenceEngineConfig(state_manager={"max_context": max_ctx, "max_ragged_batch_size": max(4 * max_ctx, 2048),
                                                        "max_ragged_sequence_count": 256})
        if self.device.type != "cuda":
            ec.state_manager.memory_config.mode = type(ec.state_manager.memory_config.mode)("allocate")
            ec.state_manager.memory_config.size = 64
        else:
            # leave the training state room: size the KV pool explicitly instead of "all free HBM"
            ec.state_manager.memory_config.mode = type(ec.state_manager.memory_config.mode)("allocate")
            ec.state_manager.memory_config.size = max(64, (256 * max_ctx) // 128)
        dtype = self._model_dtype()
        hf_cfg = getattr(self.module, "config", None)
        group, tp, tp_rank = self._inference_tp()
        if hf_cfg is not None and hasattr(hf_cfg, "model_type"):
            # every rank holds the gathered full weights here; load_hf_weights keeps this rank's TP slice only
            model = RaggedTransformer(arch_from_hf_config(hf_cfg), group, tp, tp_rank, dtype, self.device)
            load_hf_weights(model, self.module.state_dict().get)
            return InferenceEngineV2(model, ec, tp_group=group) if tp > 1 else InferenceEngineV2(model, ec)
        return build_engine_from_model(self.module, ec, dtype=dtype, device=self.device,
                                       tp_override=(group, tp, tp_rank) if tp > 1 else None)

    def _inference_tp(self):
        """Generation-time tensor parallelism (reference ``hybrid_engine.py:83`` ``inference_mp_group``): consecutive
        ranks form one TP group; each keeps ``1 / tp`` of the packed inference weights and KV cache, so a generation
        batch is served by the whole group (prompts are all-gathered inside :meth:`generate`)."""
        tp = int(getattr(self._hybrid_cfg, "inference_tp_size", 1) or 1)
        world = dist.get_world_size() if dist.is_initialized() else 1
        if tp <= 1 or world == 1:
            return None, 1, 0
        assert world % tp == 0, f"inference_tp_size {tp} must divide the world size {world}"
        if DeepSpeedHybridEngine.inference_mp_group is None or getattr(self, "_tp_world", None) != (world, tp):
            rank = dist.get_rank()
            for first in range(0, world, tp):      # every rank creates every group (collective)
                ranks = list(range(first, first + tp))
                g = dist.new_group(ranks)
                if rank in ranks:
                    DeepSpeedHybridEngine.inference_mp_group = g
            self._tp_world = (world, tp)
        return DeepSpeedHybridEngine.inference_mp_group, tp, dist.get_rank() % tp

    def _gather_ctx(self):
        from deepspeed_b200.runtime.zero.partition_parameters import GatheredParameters
        if self.zero_optimization_partition_weights():
            return GatheredParameters(list(self.module.parameters()), modifier_rank=None)
        import contextlib
        return contextlib.nullcontext()

    @torch.no_grad()
    def generate(self, *inputs, **kwargs):
        from deepspeed_b200.inference.engine import ragged_generate
        t0 = time.time()
        was_training = self.module.training
        self.module.eval()
        stale = self._ragged is None or self._packed_at_step != self.global_steps
        if stale:
            with self._gather_ctx():
                self.fuse_lora_weight()
                try:
                    self._ragged = None
                    self._ragged = self._build_ragged()
                finally:
                    self.unfuse_lora_weight()
            self._packed_at_step = self.global_steps
        self._t_gather += time.time() - t0
        t1 = time.time()
        input_ids = kwargs.pop("input_ids", inputs[0] if inputs else None)
        eos_default = getattr(getattr(self.module, "config", None), "eos_token_id", None)
        group, tp, tp_rank = self._inference_tp()
        if tp > 1:
            # the TP group decodes one batch: concatenate the members' prompts (reference generate :174-190), decode,
            # hand every rank its own rows back
            assert not kwargs.get("do_sample", False), "inference_tp_size > 1 supports greedy decoding (ranks must agree)"
            ids = input_ids.contiguous()
            parts = [torch.empty_like(ids) for _ in range(tp)]
            dist.all_gather(parts, ids, group=group)
            mask = kwargs.get("attention_mask")
            if mask is not None:
                mparts = [torch.empty_like(mask) for _ in range(tp)]
                dist.all_gather(mparts, mask.contiguous(), group=group)
                kwargs["attention_mask"] = torch.cat(mparts, 0)
            bsz = ids.shape[0]
            input_ids = torch.cat(parts, 0)
        out, self._uid = ragged_generate(self._ragged, input_ids, self._uid, int(self._hybrid_cfg.max_out_tokens),
                                         eos_default=eos_default, **kwargs)
        if tp > 1:
            out = out[tp_rank * bsz:(tp_rank + 1) * bsz]
        self._generate_latency = time.time() - t1
        self._t_generate += self._generate_latency
        self._iters += 1
        if self._hybrid_cfg.release_inference_cache:
            self._ragged = None
            if torch.cuda.is_available():
                torch.cuda.empty_cache()
        if was_training:
            self.module.train()
        return out

    def eval(self):
        self._in_eval = True
        self.module.eval()
        return self

    def train(self, mode=True):
        self._in_eval = not mode
        self.module.train(mode)
        return self

    def step(self, *a, **k):
        t = time.time()
        out = super().step(*a, **k)
        self._training_latency = time.time() - t
        self._t_train += self._training_latency
        return out

    def get_latency_report(self):
        n = max(self._iters, 1)
        return {"generate_s": self._t_generate / n, "pack_gather_s": self._t_gather / n, "train_step_s": self._t_train / n}
