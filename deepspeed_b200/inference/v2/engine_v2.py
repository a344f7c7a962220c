"""Ragged continuous-batching engine (reference ``inference/v2/engine_v2.py:30 InferenceEngineV2``).

``put(uids, tokens)`` schedules one forward over a ragged batch (any mix of prefill chunks and decodes),
``query`` / ``can_schedule`` answer the scheduler's admission questions, ``flush`` releases a sequence.
"""
from typing import Iterable, List, Tuple

import torch

from deepspeed_b200 import comm as dist
from deepspeed_b200.accelerator import get_accelerator
from deepspeed_b200.utils.logging import logger
from .config_v2 import RaggedInferenceEngineConfig
from .ragged import DSStateManager, RaggedBatchWrapper, PlaceholderSequenceDescriptor
from .scheduling_utils import SchedulingError, SchedulingResult


class InferenceEngineV2:

    def __init__(self, model, engine_config: RaggedInferenceEngineConfig, tp_group=None):
        self._config = engine_config
        self._model = model
        self._tp_group = tp_group
        smc = engine_config.state_manager
        kv = model.kv_cache_config(max_context=smc.max_context)
        self._batch = RaggedBatchWrapper(smc, max_blocks_per_seq=kv[0].max_blocks_per_allocation_group,
                                         device=model.device)
        self._state_manager = DSStateManager(smc, kv, base_mp_group=tp_group, device=model.device)
        model.set_state_manager(self._state_manager)
        self._graphs = {}

    @property
    def free_blocks(self) -> torch.Tensor:
        return self._state_manager.free_blocks

    @property
    def n_kv_cache_groups(self) -> int:
        return self._state_manager.n_kv_cache_groups

    @property
    def model(self):
        return self._model

    def put(self, batch_uids: Iterable[int], batch_tokens: Iterable[torch.Tensor], do_checks: bool = True) -> torch.Tensor:
        batch_uids = list(batch_uids)
        batch_tokens = [t if (isinstance(t, torch.Tensor) and t.dim() == 1 and t.device.type == "cpu") else
                        torch.as_tensor(t).reshape(-1).cpu() for t in batch_tokens]
        if do_checks:
            res = self.can_schedule(batch_uids, [t.numel() for t in batch_tokens])
            if res != SchedulingResult.Success:
                raise SchedulingError(res)
        self._batch.clear()
        for uid, tokens in zip(batch_uids, batch_tokens):
            seq = self._state_manager.get_or_create_sequence(uid)
            if seq.host_kv is not None:
                self._state_manager.restore_sequence(uid)
            self._model.maybe_allocate_kv(seq, tokens.numel())
            seq.pre_forward(tokens.numel())
            self._batch.insert_sequence(seq, tokens, do_checks=do_checks)
        self._batch.finalize()
        logits = self._forward()
        for uid in batch_uids:
            seq = self._state_manager.get_sequence(uid)
            seq.post_forward()
            self._model.maybe_free_kv(seq)
        return logits

    def _forward(self):
        b = self._batch
        if (self._config.cuda_graph_decode and b.is_pure_decode and torch.cuda.is_available()
                and str(self._model.device).startswith("cuda")):
            return self._graphed_decode(b.current_sequences)
        return self._model.forward(b)

    def _graphed_decode(self, n):
        """Pure-decode steps have static shapes given the sequence count; the metadata lives at fixed device
        addresses (RaggedBatchWrapper), so one captured graph per batch size replays the whole layer stack."""
        ent = self._graphs.get(n)
        if ent is None:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):
                    self._model.forward(self._batch)
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self._model.forward(self._batch)
            ent = self._graphs[n] = (g, out)
        g, out = ent
        g.replay()
        return out.clone()

    def query(self, uid: int, max_request_tokens: int, max_request_blocks) -> Tuple[int, torch.Tensor]:
        seq = self._state_manager.get_sequence(uid)
        if seq is None:
            if self._state_manager.n_tracked_sequences >= self._config.state_manager.max_tracked_sequences:
                return 0, 0
            seq = PlaceholderSequenceDescriptor()
        return self._model.get_kv_requirements(seq, max_request_tokens, int(max_request_blocks))

    def can_schedule(self, uids: Iterable[int], lengths: Iterable[int]) -> SchedulingResult:
        uids, lengths = list(uids), list(lengths)
        smc = self._config.state_manager
        cur_seqs = self._state_manager.n_tracked_sequences
        free = self._state_manager.free_block_count(0)
        if len(uids) > smc.max_ragged_sequence_count:
            return SchedulingResult.BatchSequenceLimitExceeded
        batch_len = 0
        for uid, n in zip(uids, lengths):
            seq = self._state_manager.get_sequence(uid)
            if seq is None:
                cur_seqs += 1
                seq = PlaceholderSequenceDescriptor()
            if seq.seen_tokens + n > smc.max_context:
                return SchedulingResult.SequenceTokenLimitExceeded
            sched_len, sched_blocks = self._model.get_kv_requirements(seq, n, free)
            if sched_len != n:
                return SchedulingResult.KVCacheLimitExceeded
            batch_len += n
            free -= sched_blocks
        if cur_seqs > smc.max_tracked_sequences:
            return SchedulingResult.EngineSequenceLimitExceeded
        if batch_len > smc.max_ragged_batch_size:
            return SchedulingResult.BatchTokenLimitExceeded
        return SchedulingResult.Success

    def get_remaining_block_capacity(self, uid: int) -> int:
        seq = self._state_manager.get_sequence(uid)
        return 0 if seq is None else self._model.get_remaining_block_capacity(seq)

    def flush(self, uid: int) -> None:
        self._state_manager.flush_sequence(uid)

    def serialize(self, save_path: str) -> None:
        """Write this rank's already-sharded weights for fast reload (reference ``engine_v2.py:237``)."""
        import os
        os.makedirs(save_path, exist_ok=True)
        m = self._model
        blob = {"spec": m.spec.__dict__, "tp_size": m.tp_size, "globals": {k: getattr(m, k) for k in
                ("embed_w", "pos_w", "final_ln_w", "final_ln_b", "lm_head_w", "lm_head_b")},
                "layers": [{s: getattr(lw, s) for s in lw.__slots__} for lw in m.layers]}
        from .model_implementations import flat_model_helpers as F
        torch.save(blob, F.make_param_filename(save_path, m.tp_rank, m.tp_size))
        # side files (reference layout): per-rank tensor table + the model config, so tools can inspect a serialized model
        # without unpickling the weights
        table, off = {}, 0
        for name, t in m.flat_tensors().items():
            off = F.pad_to_aligned_offset(off)
            table[name] = {"offset": off, "shape": list(t.shape), "dtype": str(t.dtype)}
            off += t.numel() * t.element_size()
        with open(F.make_metadata_filename(save_path, m.tp_rank, m.tp_size), "w") as f:
            f.write(F.to_model_metadata(table, policy=type(m).__name__).model_dump_json())
        if m.tp_rank == 0:
            import json
            with open(F.make_model_config_filename(save_path), "w") as f:
                json.dump({"spec": {k: (v if isinstance(v, (int, float, str, bool, type(None), list, dict)) else str(v))
                                    for k, v in m.spec.__dict__.items()}, "tp_size": m.tp_size}, f)
