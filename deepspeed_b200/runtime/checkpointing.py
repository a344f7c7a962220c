"""Checkpoint save / load for :class:`DeepSpeedEngine`.

Parity target: reference ``runtime/engine.py:2801-3809`` -- same directory layout and file names
(``latest``, ``<tag>/mp_rank_XX_model_states.pt``, ``<tag>/zero_pp_rank_D_mp_rank_XX_model_states.pt``
for ZeRO-3, ``<tag>/[bf16_]zero_pp_rank_D_mp_rank_XX_optim_states.pt``), the same top-level keys in
the model-states dict, tag validation across ranks, pluggable checkpoint engines,
``save_16bit_model`` consolidation and a ``zero_to_fp32`` converter copied next to the shards.

The *inside* of the optimizer shard is this framework's arena layout (flat fp32 master + flat
optimizer states + the unit plan) -- ``deepspeed_b200/utils/zero_to_fp32.py`` and the universal
checkpoint converter understand it; ``ds_b200_layout`` in the model-states file records where every
parameter lives in the arenas so any DP degree can be reconstructed offline.
"""
import hashlib
import os
import shutil
from collections import OrderedDict

import torch

from deepspeed_b200 import comm as dist
from deepspeed_b200.utils import groups
from deepspeed_b200.utils.logging import log_dist, logger

LATEST = "latest"
DS_VERSION = "ds_version"


def _mp_rank():
    return groups._get_model_parallel_rank()


class CheckpointMixin:

    # ---- configuration --------------------------------------------------------------------------
    def _configure_checkpointing(self):
        from deepspeed_b200.runtime.checkpoint_engine import build_checkpoint_engine
        self.checkpoint_engine = build_checkpoint_engine(self._config)
        dp_rank = groups._get_sequence_data_parallel_rank() if self.sequence_parallel_size > 1 else \
            groups._get_data_parallel_rank()
        # who writes the (replicated) model states; every DP rank writes its own ZeRO shard
        self.save_non_zero_checkpoint = dp_rank == 0 or self.zero_optimization_partition_weights()
        # every DP rank writes its own ZeRO shard; an unsharded optimizer (stage 0) is replicated -> dp rank 0 only
        sharded = getattr(self.optimizer, "shard_world", 1) > 1 or self.zero_optimization_partition_weights()
        self.save_zero_checkpoint = self.optimizer is not None and (sharded or dp_rank == 0)
        self._optimizer_replicated = self.optimizer is not None and not sharded
        self._dp_rank_for_ckpt = dp_rank

    # ---- names ---------------------------------------------------------------------------------------
    def _ckpt_mp_rank(self):
        """The ``mp_rank`` of the file names.  Under pipeline parallelism the reference counts every rank that shares a
        data-parallel index as one "model" group (``pipe/topology.py:302-313 ds_model_rank``: stage-major, tensor-slice
        minor), so each pipeline stage writes its own model-states / optimizer files."""
        grid = getattr(self, "grid", None)
        if grid is not None and getattr(grid, "pipe_parallel_size", 1) > 1:
            return grid.get_stage_id() * max(1, grid.get_model_parallel_world_size()) + grid.get_model_parallel_rank()
        return _mp_rank()

    def _get_ckpt_name(self, checkpoints_path, tag, mp_placeholder=None):
        mp = f"{self._ckpt_mp_rank():02d}" if mp_placeholder is None else mp_placeholder
        if self.zero_optimization_partition_weights():
            name = f"zero_pp_rank_{self._dp_rank_for_ckpt}_mp_rank_{mp}_model_states.pt"
        else:
            name = f"mp_rank_{mp}_model_states.pt"
        return os.path.join(checkpoints_path, str(tag), name)

    def _get_zero_ckpt_prefix(self, dp_rank, bf16_mode):
        return f"{'bf16_' if bf16_mode else ''}zero_pp_rank_{dp_rank}_mp_rank_{self._ckpt_mp_rank():02d}"

    def _get_zero_ckpt_name(self, checkpoints_path, tag):
        bf16 = self.bfloat16_enabled()
        dp_rank = 0 if getattr(self, "_optimizer_replicated", False) else self._dp_rank_for_ckpt
        return os.path.join(checkpoints_path, str(tag), f"{self._get_zero_ckpt_prefix(dp_rank, bf16)}_optim_states.pt")

    def _get_optimizer_ckpt_name(self, checkpoints_path, tag, expp_rank):
        """Reference name of the per-expert-parallel-rank optimizer file of a non-ZeRO MoE run (engine.py:2844). Every
        optimizer here is a sharded one and goes through ``_get_zero_ckpt_name``; kept for tools that build the path."""
        return os.path.join(checkpoints_path, str(tag), f"expp_rank_{expp_rank}_mp_rank_{self._ckpt_mp_rank():02d}_optim_states.pt")

    def _get_expert_ckpt_name(self, checkpoints_path, layer_id, expert_id, tag, mp_placeholder=None):
        mp = f"{self._ckpt_mp_rank():02d}" if mp_placeholder is None else mp_placeholder
        return os.path.join(checkpoints_path, str(tag), f"layer_{layer_id}_expert_{expert_id}_mp_rank_{mp}_model_states.pt")

    # ---- tag validation -----------------------------------------------------------------------------
    def _checkpoint_tag_validation(self, tag):
        if not self._config.checkpoint_tag_validation_enabled or dist.get_world_size() == 1:
            return
        digest = hashlib.sha1(str(tag).encode()).digest()
        dev = self.device if dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor(list(digest), dtype=torch.uint8, device=dev).to(torch.int32)
        hi, lo = t.clone(), t.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        ok = bool(torch.equal(hi, t) and torch.equal(lo, t))
        msg = (f"[rank={dist.get_rank()}] The checkpoint tag name '{tag}' is not consistent across all ranks. "
               f"Including rank unique information in checkpoint tag could cause issues when restoring with "
               f"different world sizes.")
        if self._config.checkpoint_tag_validation_fail:
            assert ok, msg
        elif not ok:
            logger.warning(msg)

    # ---- layout metadata ----------------------------------------------------------------------------
    def _arena_layout(self):
        zo = self.optimizer
        units = []
        for u in zo.units:
            units.append({
                "name": u.name,
                "full_numel": u.full_numel,
                "shard_numel": u.shard_numel,
                "arena_offset": u.arena_offset,
                "slots": [(s.name, s.offset, s.numel, tuple(s.shape), s.group) for s in u.slots],
            })
        return {"shard_world": zo.shard_world, "arena_numel": zo.arena_numel, "units": units, "stage": zo.stage}

    def _get_zero_param_shapes(self):
        """List (per optimizer group) of ``OrderedDict name -> shape`` (reference engine.py:3536), in the order the
        optimizer shards flatten the parameters (``ref_layout.group_param_order``)."""
        shapes = []
        if self.optimizer is None:
            return shapes
        if hasattr(self.optimizer, "rts") and not hasattr(self.optimizer, "parts"):
            from deepspeed_b200.runtime.zero.ref_layout import group_param_order
            for lst in group_param_order(self.optimizer):
                shapes.append(OrderedDict((s.name, torch.Size(s.shape)) for _, s in lst))
            return shapes
        names = {id(p): n for n, p in self.module.named_parameters()}
        for g in self.optimizer.param_groups:
            od = OrderedDict()
            for p in g["params"]:
                od[names.get(id(p), "?")] = torch.Size(getattr(p, "ds_shape", p.shape))
            shapes.append(od)
        return shapes

    def _get_zero_frozen_param_attributes(self, fn):
        od = OrderedDict()
        for n, p in self.module.named_parameters():
            if not p.requires_grad:
                od[n] = fn(p)
        return od

    @torch.no_grad()
    def _frozen_param_fragment(self, p):
        """Reference ``_get_param_fragment_func`` (engine.py:3520): the whole tensor below stage 3, this DP rank's
        zero-padded ``ceil(numel / dp)`` piece at stage 3 (collective: every rank walks the parameters in the same order)."""
        from deepspeed_b200.runtime.zero.partition_parameters import is_zero_param, materialize_full
        if not (self.zero_optimization_partition_weights() and is_zero_param(p)):
            return p.detach().cpu().clone()
        flat = materialize_full(p).detach().reshape(-1)
        world, rank = self.seq_dp_world_size, self._dp_rank_for_ckpt
        per = -(-flat.numel() // world)
        piece = torch.zeros(per, dtype=flat.dtype)
        lo, hi = rank * per, min((rank + 1) * per, flat.numel())
        if hi > lo:
            piece[:hi - lo].copy_(flat[lo:hi])
        return piece

    @torch.no_grad()
    def _restore_frozen_params(self, ckpt, load_dir, tag, own_file):
        """Stage 3: rebuild every frozen parameter from the per-rank pieces saved next to the module state and write it
        into the unit arenas (below stage 3 the frozen values arrive through the module state dict)."""
        frags, shapes = ckpt.get("frozen_param_fragments"), ckpt.get("frozen_param_shapes")
        if not frags or not shapes or self.optimizer is None:
            return
        from deepspeed_b200.runtime.zero.partition_parameters import _owner
        named = dict(self.module.named_parameters())
        world = self.seq_dp_world_size
        same = own_file and (ckpt.get("dp_world_size") or 1) == world
        pieces = None
        if not same:  # written at another DP degree: read the pieces of every saved rank
            import glob
            import re
            own = os.path.basename(self._get_ckpt_name(load_dir, tag))
            pat = re.sub(r"zero_pp_rank_\d+_", "zero_pp_rank_*_", own)
            files = sorted(glob.glob(os.path.join(load_dir, str(tag), pat)),
                           key=lambda f: int(re.search(r"zero_pp_rank_(\d+)_", os.path.basename(f)).group(1)))
            pieces = [self.checkpoint_engine.load(f, map_location="cpu").get("frozen_param_fragments") or {} for f in files]
        for name, shape in shapes.items():
            p = named.get(name)
            if p is None or name not in frags:
                continue
            n = 1
            for d in shape:
                n *= int(d)
            if same:
                dev = self.device if dist.get_backend() == "nccl" else "cpu"
                local = frags[name].reshape(-1).to(dev)
                if world > 1:
                    out = torch.empty(local.numel() * world, dtype=local.dtype, device=dev)
                    dist.all_gather_into_tensor(out, local.contiguous(), group=self.seq_data_parallel_group)
                else:
                    out = local
            else:
                out = torch.cat([d[name].reshape(-1) for d in pieces])
            full = out[:n].view(tuple(shape))
            zo = _owner(p)
            if zo is not None:
                zo.set_full_hp_param(full, p)
            else:
                p.data.copy_(full.to(p.device, p.dtype))

    # ---- save ------------------------------------------------------------------------------------------
    def save_checkpoint(self, save_dir, tag=None, client_state=None, save_latest=True, exclude_frozen_parameters=False):
        client_state = client_state or {}
        if tag is None:
            tag = f"global_step{self.global_steps}"
        tag = str(tag)
        self._checkpoint_tag_validation(tag)
        ce = self.checkpoint_engine
        if dist.get_rank() == 0:
            ce.makedirs(os.path.join(save_dir, tag), exist_ok=True)
        dist.barrier()
        ce.create(tag)
        if self.has_moe_layers:
            self._save_moe_checkpoint(save_dir, tag, client_state, exclude_frozen_parameters)
        elif self.save_non_zero_checkpoint:
            self._save_checkpoint(save_dir, tag, client_state, exclude_frozen_parameters)
        if self.save_zero_checkpoint:
            self._save_zero_checkpoint(save_dir, tag)
        ce.commit(tag)
        dist.barrier()
        if save_latest and dist.get_rank() == 0:
            with open(os.path.join(save_dir, LATEST), "w") as f:
                f.write(tag)
        dist.barrier()
        return True

    def _model_state_payload(self, client_state, exclude_frozen_parameters=False, module_sd=None):
        from deepspeed_b200 import __version__
        zero3 = self.zero_optimization_partition_weights()
        if module_sd is None:
            module_sd = self.module_state_dict(exclude_frozen_parameters=exclude_frozen_parameters)
            if zero3:
                # placeholders only: real values live in the ZeRO shards
                module_sd = OrderedDict((k, (v if v.numel() else torch.empty(0, dtype=v.dtype)))
                                        for k, v in module_sd.items())
        # frozen parameters never reach the optimizer shards: the reference (engine.py:3465) keeps them next to the
        # module state, whole for stage 2 and as this rank's ceil(numel / dp) piece for stage 3
        save_frozen = (self.optimizer is not None and self.zero_optimization_partition_gradients()
                       and not exclude_frozen_parameters)
        state = dict(
            module=module_sd,
            buffer_names=[n for n, _ in self.module.named_buffers()],
            optimizer=None,
            param_shapes=self._get_zero_param_shapes() if self.optimizer is not None else None,
            frozen_param_shapes=self._get_zero_frozen_param_attributes(lambda p: torch.Size(
                getattr(p, "ds_shape", p.shape))) if save_frozen else None,
            frozen_param_fragments=self._get_zero_frozen_param_attributes(self._frozen_param_fragment)
            if save_frozen else None,
            shared_params=self._get_shared_params(),
            lr_scheduler=self.lr_scheduler.state_dict() if self.lr_scheduler is not None and hasattr(
                self.lr_scheduler, "state_dict") else None,
            data_sampler=self.training_dataloader.data_sampler.state_dict() if
            (self.training_dataloader is not None and getattr(self.training_dataloader, "curriculum_learning_enabled",
                                                              False)) else None,
            random_ltd=None,
            sparse_tensor_module_names=set(getattr(self, "sparse_tensor_module_names", ())),
            skipped_steps=self.skipped_steps,
            global_steps=self.global_steps,
            global_samples=self.global_samples,
            dp_world_size=self.seq_dp_world_size,
            mp_world_size=self.mp_world_size,
            ds_config=self._config._param_dict,
            ds_version=__version__,
            ds_b200_layout=self._arena_layout() if self.optimizer is not None else None,
        )
        state.update(client_state)
        return state

    def _get_shared_params(self):
        """Tied parameters: ``{alias_name: canonical_name}`` (reference engine.py:3575)."""
        names = {}
        for n, p in self.module.named_parameters(remove_duplicate=False):
            names.setdefault(id(p), []).append(n)
        # the canonical name of a tied group is the one the optimizer shards are written under (consolidation tools
        # resolve ``alias -> canonical`` against the flattened fp32 weights)
        in_shards = set()
        if self.optimizer is not None:
            for od in self._get_zero_param_shapes():
                in_shards.update(od.keys())
        shared = {}
        for group in names.values():
            if len(group) < 2:
                continue
            canon = next((n for n in group if n in in_shards), group[0])
            for n in group:
                if n != canon:
                    shared[n] = canon
        return shared

    def _save_checkpoint(self, save_dir, tag, client_state, exclude_frozen_parameters=False):
        path = self._get_ckpt_name(save_dir, tag)
        self.checkpoint_engine.save(self._model_state_payload(client_state, exclude_frozen_parameters), path)

    def _save_zero_checkpoint(self, save_path, tag):
        from deepspeed_b200 import __version__
        path = self._get_zero_ckpt_name(save_path, tag)
        layout = (self._config._param_dict.get("checkpoint") or {}).get("b200_shard_layout", "reference")
        try:
            osd = self.optimizer.state_dict(layout=layout)
        except TypeError:  # client / wrapped optimizers without the layout switch
            osd = self.optimizer.state_dict()
        zsd = dict(optimizer_state_dict=osd, ds_config=self._config._param_dict, ds_version=__version__)
        self.checkpoint_engine.save(zsd, path)
        if dist.get_rank() == 0:
            self._copy_recovery_script(save_path)

    def _copy_recovery_script(self, save_path):
        src = os.path.join(os.path.dirname(os.path.dirname(__file__)), "utils", "zero_to_fp32.py")
        if os.path.exists(src):
            dst = os.path.join(save_path, "zero_to_fp32.py")
            shutil.copyfile(src, dst)
            os.chmod(dst, os.stat(dst).st_mode | 0o111)

    def _save_moe_checkpoint(self, save_dir, tag, client_state, exclude_frozen_parameters=False):
        """Per-expert files + a non-expert model-states file (reference engine.py:3319)."""
        from deepspeed_b200.moe.layer import MoE
        full_sd = self.module_state_dict(exclude_frozen_parameters=exclude_frozen_parameters)
        expert_keys = set()
        moe_layer_id = 0
        for n_module, module in self.module.named_modules():
            if not isinstance(module, MoE):
                continue
            ep_rank = groups._get_expert_parallel_rank(module.expert_group_name)
            edp_rank = groups._get_expert_data_parallel_rank(module.expert_group_name)
            n_local = module.num_local_experts
            prefix = f"{n_module}.deepspeed_moe.experts.deepspeed_experts."
            for local_id in range(n_local):
                global_id = ep_rank * n_local + local_id
                sub = OrderedDict()
                for k, v in full_sd.items():
                    if k.startswith(f"{prefix}{local_id}."):
                        expert_keys.add(k)
                        sub[k.replace(f"{prefix}{local_id}.", f"{prefix}{global_id}.")] = v.detach().cpu().clone()
                if edp_rank == 0:
                    self.checkpoint_engine.save(sub, self._get_expert_ckpt_name(save_dir, moe_layer_id, global_id, tag))
            moe_layer_id += 1
        if self.save_non_zero_checkpoint:
            non_expert = OrderedDict((k, v) for k, v in full_sd.items() if k not in expert_keys)
            payload = self._model_state_payload(client_state, module_sd=non_expert)
            payload["num_experts"] = self.num_experts
            self.checkpoint_engine.save(payload, self._get_ckpt_name(save_dir, tag))

    # ---- load ---------------------------------------------------------------------------------------------
    def load_checkpoint(self, load_dir, tag=None, load_module_strict=True, load_optimizer_states=True,
                        load_lr_scheduler_states=True, load_module_only=False, custom_load_fn=None):
        if tag is None:
            latest = os.path.join(load_dir, LATEST)
            if os.path.isfile(latest):
                with open(latest) as f:
                    tag = f.read().strip()
            else:
                if self._config.load_universal_checkpoint:
                    raise ValueError(f"Invalid for universal checkpoint: {latest} does not exist")
                logger.warning(f"Unable to find latest file at {latest}, if trying to load latest checkpoint please "
                               f"ensure this file exists or pass an explicit checkpoint tag when loading a checkpoint.")
                return None, None
        if self._config.load_universal_checkpoint:
            from deepspeed_b200.checkpoint.universal_checkpoint import load_universal_into_engine
            return load_universal_into_engine(self, load_dir, tag, load_optimizer_states)
        path, client = self._load_checkpoint(load_dir, tag, load_module_strict, load_optimizer_states,
                                             load_lr_scheduler_states, load_module_only, custom_load_fn)
        if path is None:
            return None, None
        if self.optimizer is not None and (not load_module_only or self.zero_optimization_partition_weights()):
            # ZeRO-3 keeps the weights themselves in the per-rank shards: a module-only load still reads the fp32
            # partitions (optimizer moments are left alone)
            ok = self._load_zero_checkpoint(load_dir, tag, load_optimizer_states and not load_module_only)
            if not ok:
                return None, None
        return path, client

    def _load_checkpoint(self, load_dir, tag, strict, load_opt, load_sched, module_only, custom_load_fn):
        path = self._get_ckpt_name(load_dir, tag)
        own_file = os.path.exists(path)
        if not own_file:
            # ZeRO-3 files are per-DP-rank; fall back to rank-0 file for replicated metadata
            alt = path.replace(f"zero_pp_rank_{self._dp_rank_for_ckpt}_", "zero_pp_rank_0_")
            if os.path.exists(alt):
                path = alt
            else:
                logger.warning(f"Client provided checkpoint load path: {path} does not exist")
                return None, None
        ckpt = self.checkpoint_engine.load(path, map_location="cpu")
        if self.has_moe_layers:
            self._load_moe_experts(load_dir, tag, ckpt["module"])
        if not self.zero_optimization_partition_weights():
            self.load_module_state_dict(ckpt, strict=strict, custom_load_fn=custom_load_fn)
            if self.optimizer is not None:
                self._resync_arena_from_module()
        else:
            # buffers only; parameters come from the ZeRO shards
            bufs = {k: v for k, v in ckpt["module"].items() if k in ckpt.get("buffer_names", [])}
            if bufs:
                self.module.load_state_dict(bufs, strict=False)
            self._restore_frozen_params(ckpt, load_dir, tag, own_file=own_file)
        saved_sparse = ckpt.get("sparse_tensor_module_names", ckpt.get("csr_tensor_module_names"))
        if saved_sparse is not None:
            if strict:
                self.sparse_tensor_module_names = set(saved_sparse)
            else:  # keep what is sparse in both models; parameters new to this model keep their own setting
                mine, here = set(self.sparse_tensor_module_names), dict(self.module.named_parameters())
                keep = {n for n in mine if not (n in ckpt["module"] and n not in saved_sparse)}
                self.sparse_tensor_module_names = keep | {n for n in saved_sparse if n in here}
        self._loaded_param_shapes = ckpt.get("param_shapes")
        self.loaded_checkpoint_dp_world_size = ckpt.get("dp_world_size")
        self.loaded_checkpoint_mp_world_size = ckpt.get("mp_world_size")
        if not module_only:
            self.global_steps = ckpt.get("global_steps", 0)
            self.global_samples = ckpt.get("global_samples", self.global_steps * self.train_batch_size())
            self.skipped_steps = ckpt.get("skipped_steps", 0)
            if load_sched and self.lr_scheduler is not None and ckpt.get("lr_scheduler") is not None:
                self.lr_scheduler.load_state_dict(ckpt["lr_scheduler"])
        reserved = {"module", "buffer_names", "optimizer", "param_shapes", "frozen_param_shapes",
                    "frozen_param_fragments", "shared_params", "lr_scheduler", "data_sampler", "random_ltd",
                    "sparse_tensor_module_names", "skipped_steps", "global_steps", "global_samples", "dp_world_size",
                    "mp_world_size", "ds_config", "ds_version", "ds_b200_layout", "num_experts"}
        client_state = {k: v for k, v in ckpt.items() if k not in reserved}
        return path, client_state

    def _load_moe_experts(self, load_dir, tag, module_sd):
        from deepspeed_b200.moe.layer import MoE
        moe_layer_id = 0
        for n_module, module in self.module.named_modules():
            if not isinstance(module, MoE):
                continue
            ep_rank = groups._get_expert_parallel_rank(module.expert_group_name)
            n_local = module.num_local_experts
            prefix = f"{n_module}.deepspeed_moe.experts.deepspeed_experts."
            for local_id in range(n_local):
                global_id = ep_rank * n_local + local_id
                sub = self.checkpoint_engine.load(self._get_expert_ckpt_name(load_dir, moe_layer_id, global_id, tag),
                                                  map_location="cpu")
                for k, v in sub.items():
                    module_sd[k.replace(f"{prefix}{global_id}.", f"{prefix}{local_id}.")] = v
            moe_layer_id += 1

    @torch.no_grad()
    def _resync_arena_from_module(self):
        """After ``module.load_state_dict`` (stage <= 2) refresh the fp32 master from the parameters."""
        for zo in getattr(self.optimizer, "parts", None) or [self.optimizer]:  # every reduction domain (dense + experts)
            if zo.master is None:
                continue
            for rt in zo.rts:
                a = rt.u.arena_offset
                zo.master[a:a + rt.u.shard_numel].copy_(zo._lp_shard(rt.u))

    def _saved_zero_shard_files(self, load_dir, tag):
        """Optimizer shard files of every saved DP rank for this model-parallel rank, in DP-rank order."""
        import glob
        import re
        own = os.path.basename(self._get_zero_ckpt_name(load_dir, tag))
        pat = re.sub(r"zero_pp_rank_\d+_", "zero_pp_rank_*_", own)
        files = glob.glob(os.path.join(load_dir, str(tag), pat))
        key = lambda f: int(re.search(r"zero_pp_rank_(\d+)_", os.path.basename(f)).group(1))
        return sorted(files, key=key)

    def _load_zero_checkpoint(self, load_dir, tag, load_optimizer_states=True):
        path = self._get_zero_ckpt_name(load_dir, tag)
        zo = self.optimizer
        saved = self._saved_zero_shard_files(load_dir, tag)
        want = 1 if getattr(self, "_optimizer_replicated", False) else getattr(zo, "shard_world", 1)
        if saved and len(saved) != want:
            # the data-parallel degree changed since the checkpoint was written: reassemble every parameter from ALL saved
            # shards and re-partition (reference `elastic_checkpoint`; here for stages 1-3)
            from deepspeed_b200.runtime.zero import ref_layout
            shards = [self.checkpoint_engine.load(f, map_location="cpu")["optimizer_state_dict"] for f in saved]
            if not all(ref_layout.is_reference_layout(sd) for sd in shards):
                raise ValueError(f"the checkpoint under {load_dir}/{tag} was written by {len(saved)} data-parallel ranks in the "
                                 f"rank-local arena layout; only reference-layout shards (checkpoint.b200_shard_layout="
                                 f"'reference', the default) can be loaded at a different degree ({want}) -- or convert with "
                                 f"ds_to_universal")
            parts = zo.parts if hasattr(zo, "parts") else [zo]
            if len(parts) != 1:
                raise ValueError("loading at a different data-parallel degree is not supported for multi-domain (MoE / Twin-Flow) "
                                 "optimizers; convert with ds_to_universal")
            ref_layout.import_reference_state_elastic(
                parts[0], shards, load_optimizer_states=load_optimizer_states,
                load_from_fp32_weights=self._config.zero_config.load_from_fp32_weights
                or self.zero_optimization_partition_weights(), param_shapes=getattr(self, "_loaded_param_shapes", None))
            log_dist(f"loaded zero checkpoint written by {len(saved)} data-parallel ranks into {want} (elastic)", ranks=[0])
            return True
        if not os.path.exists(path):
            logger.warning(f"The following zero checkpoint path is missing: {path}; if the DP world size changed, "
                           f"convert with ds_to_universal and set checkpoint.load_universal")
            return False
        zsd = self.checkpoint_engine.load(path, map_location="cpu")
        kw = {}
        if getattr(self, "_loaded_param_shapes", None) is not None:
            kw["param_shapes"] = self._loaded_param_shapes  # parameter order of the shards (a stock checkpoint's own)
        self.optimizer.load_state_dict(zsd["optimizer_state_dict"], load_optimizer_states=load_optimizer_states,
                                       load_from_fp32_weights=self._config.zero_config.load_from_fp32_weights
                                       or self.zero_optimization_partition_weights(), **kw)
        log_dist(f"loaded zero checkpoint {path}", ranks=[0])
        return True

    # ---- 16-bit export ---------------------------------------------------------------------------------------
    def _zero3_consolidated_16bit_state_dict(self, exclude_frozen_parameters=False):
        """Gather every ZeRO-3 parameter unit by unit; rank 0 keeps CPU copies (reference :3693)."""
        zo = self.optimizer
        sd = OrderedDict()
        for rt in zo.rts:
            zo.gather_param_temp(rt.u.slots[0].param)
            if dist.get_rank() == 0:
                for s in rt.u.slots:
                    if exclude_frozen_parameters and not s.param.requires_grad:
                        continue
                    sd[s.name] = s.param.detach().cpu().clone()
            zo.release_param_temp(rt.u.slots[0].param)
        if dist.get_rank() == 0:
            for n, b in self.module.named_buffers():
                sd[n] = b.detach().cpu()
            for alias, canon in self._get_shared_params().items():
                if canon in sd:
                    sd[alias] = sd[canon]
        return sd if dist.get_rank() == 0 else None

    def autotp_size(self):
        return int(getattr(self._config.tensor_parallel_config, "autotp_size", 0) or 0)

    def _replace_module_consolidated_state_dict(self):
        """Tensor-parallel training (reference engine.py:3647): full, un-sharded 16-bit weights on the CPU of rank 0,
        ``None`` elsewhere. Collective -- every rank must call it; one layer is reassembled at a time."""
        from deepspeed_b200.module_inject.layers import GatherReplacedLayerParams, TensorParallel_Layer
        sd = OrderedDict() if dist.get_rank() == 0 else None

        def walk(module, prefix):
            is_tp = isinstance(module, TensorParallel_Layer)
            with GatherReplacedLayerParams(list(module.parameters(recurse=False)), module, enabled=is_tp):
                for name, p in module.named_parameters(recurse=False):
                    if sd is not None:
                        sd[prefix + name] = p.detach().cpu().clone()
            if sd is not None:
                for name, b in module.named_buffers(recurse=False):
                    if name not in module._non_persistent_buffers_set:
                        sd[prefix + name] = b.detach().cpu()
            for name, child in module.named_children():
                walk(child, prefix + name + ".")

        walk(self.module, "")
        return sd

    def _consolidated_16bit_state_dict(self, exclude_frozen_parameters=False):
        """Full 16-bit state dict of a model whose weights are partitioned (ZeRO-3 or tensor parallelism)."""
        if self.zero_optimization_partition_weights():
            return self._zero3_consolidated_16bit_state_dict(exclude_frozen_parameters)
        if self.autotp_size() > 1:
            return self._replace_module_consolidated_state_dict()
        raise ValueError("consolidated_16bit_state_dict is only applicable to cases where weights are partitioned, "
                         "including Zero Stage 3 and tensor parallelism.")

    def save_16bit_model(self, save_dir, save_filename="pytorch_model.bin", exclude_frozen_parameters=False):
        path = os.path.join(save_dir, save_filename)
        if self.zero_optimization_partition_weights():
            if not self._config.zero_config.gather_16bit_weights_on_model_save:
                logger.info(f"Did not save the model {path} because `stage3_gather_16bit_weights_on_model_save` is False")
                return False
            sd = self._zero3_consolidated_16bit_state_dict(exclude_frozen_parameters)
        elif self.autotp_size() > 1:
            sd = self._replace_module_consolidated_state_dict()  # re-assembled from the tensor-parallel shards
        else:
            sd = self.module_state_dict(exclude_frozen_parameters=exclude_frozen_parameters)
        if dist.get_rank() == 0:
            self.checkpoint_engine.makedirs(save_dir, exist_ok=True)
            logger.info(f"Saving model weights to {path}, tag: {save_filename}")
            self.checkpoint_engine.save(sd, path)
        dist.barrier()
        return True

    save_fp16_model = save_16bit_model
