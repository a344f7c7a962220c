"""Elastic batch-size planning: pick ONE global batch size that stays valid across as many GPU counts as
possible, so a job can shrink/grow without changing convergence-relevant hyper-parameters.

Behavioural parity with reference ``elasticity/elasticity.py`` (``compute_elastic_config :233``, v0.1 and v0.2
algorithms).  The search: candidate batch sizes are ``base * h`` for every base in ``micro_batches + [lcm]`` and
``h`` the largest *highly composite number* keeping the product under the cap (HCNs maximise divisor count →
maximise compatible world sizes); the candidate with the most compatible GPU counts wins, ties broken toward
larger (or smaller) batch.  The HCN table is generated on first use rather than hard-coded.
"""
import functools
import json
import math
import os
from typing import List

from packaging import version as pkg_version

from deepspeed_b200.utils.logging import logger
from . import constants as C
from .config import ElasticityConfig, ElasticityConfigError, ElasticityError, ElasticityIncompatibleWorldSize


@functools.lru_cache(maxsize=None)
def highly_composite_numbers(limit: int = 1_000_000) -> tuple:
    """All n <= limit with more divisors than every smaller number (sieve of divisor counts)."""
    counts = [0] * (limit + 1)
    for d in range(1, limit + 1):
        for m in range(d, limit + 1, d):
            counts[m] += 1
    out, best = [], 0
    for n in range(1, limit + 1):
        if counts[n] > best:
            best = counts[n]
            out.append(n)
    return tuple(out)


def _largest_hcn_at_most(v: int) -> int:
    limit = 1024
    while limit < v:
        limit *= 4
    hcn = highly_composite_numbers(min(limit, 1_000_000))
    best = 1
    for h in hcn:
        if h > v:
            break
        best = h
    return best


def _candidate_batches(bases: List[int], cap: int) -> List[int]:
    out = set()
    for b in bases:
        out.add(b if b >= cap else b * _largest_hcn_at_most(cap // b))
    return sorted(out)


def _valid_gpu_counts(batch: int, micro_batches: List[int], lo: int, hi: int) -> List[int]:
    ok = set()
    for mb in micro_batches:
        if batch % mb:
            continue
        slots = batch // mb  # number of micro-batches making up the batch: any divisor is a legal GPU count
        for g in range(1, int(math.isqrt(slots)) + 1):
            if slots % g == 0:
                for c in (g, slots // g):
                    if lo <= c <= hi:
                        ok.add(c)
    return sorted(ok)


def _plan_v01(micro_batches, cap, min_gpus=None, max_gpus=None, prefer_larger=True):
    min_gpus = min_gpus or 1
    max_gpus = max_gpus or cap // min(micro_batches)
    if any(mb > cap for mb in micro_batches):
        raise ValueError(f"All micro batches must be less than or equal to max_acceptable_batch_size: {cap}")
    lcm = functools.reduce(lambda a, b: a * b // math.gcd(a, b), micro_batches)
    best_batch, best_gpus = int(min(micro_batches)), None
    best_n = 0
    for cand in _candidate_batches(list(micro_batches) + [lcm], cap):
        gpus = _valid_gpu_counts(cand, micro_batches, min_gpus, max_gpus)
        better_tie = (cand > best_batch) if prefer_larger else (cand < best_batch)
        if len(gpus) > best_n or (len(gpus) == best_n and better_tie):
            best_n, best_gpus, best_batch = len(gpus), gpus, cand
    return best_batch, best_gpus


def _plan_v02(micro_batches, cap, current_num_gpus, min_gpus=None, max_gpus=None, prefer_larger=True,
              num_gpus_per_node=1, model_parallel_size=1):
    if num_gpus_per_node % model_parallel_size:
        raise ElasticityError(f"In Elasticity v0.2, number of GPUs per node:{num_gpus_per_node} should be divisible by "
                              f"model parallel size {model_parallel_size}")
    dp_per_node = num_gpus_per_node // model_parallel_size

    def pick_micro(batch):
        fits = [mb for mb in micro_batches if (batch // current_num_gpus) % mb == 0]
        if not fits:
            return None
        return max(fits) if prefer_larger else fits[0]

    batch, nodes = _plan_v01(micro_batches, int(cap / dp_per_node), int(min_gpus / num_gpus_per_node),
                             int(max_gpus / num_gpus_per_node), prefer_larger)
    batch = int(batch) * dp_per_node
    dp_sizes = [n * dp_per_node for n in (nodes or [])]
    if current_num_gpus // model_parallel_size in dp_sizes:
        return batch, dp_sizes, pick_micro(batch)
    cur_dp = (current_num_gpus / num_gpus_per_node) * dp_per_node
    cands = [math.floor(cap / float(mb * cur_dp)) * mb * cur_dp for mb in micro_batches]
    chosen = max(cands) if prefer_larger else min(cands)
    return chosen, [int(cur_dp)], pick_micro(chosen)


def elasticity_enabled(ds_config: dict) -> bool:
    return ds_config.get(C.ELASTICITY, {}).get(C.ENABLED, C.ENABLED_DEFAULT)


def ensure_immutable_elastic_config(runtime_elastic_config_dict: dict):
    """The scheduler that launched the job froze its elastic config in the environment; the runtime config must
    agree on the three fields that determine the plan."""
    frozen = os.environ.get(C.DEEPSPEED_ELASTICITY_CONFIG)
    if frozen is None:
        logger.warning("Unable to find DEEPSPEED_ELASTICITY_CONFIG environment variable, cannot guarantee resource "
                       "scheduler will scale this job using compatible GPU counts.")
        return
    sched = ElasticityConfig(json.loads(frozen))
    run = ElasticityConfig(runtime_elastic_config_dict)
    for attr, what in (("max_acceptable_batch_size", "max_acceptable_batch_size"), ("micro_batches", "micro_batches"),
                       ("version", "version")):
        if getattr(run, attr) != getattr(sched, attr):
            raise ElasticityConfigError(f"Environment variable {C.DEEPSPEED_ELASTICITY_CONFIG} {what} "
                                        f"{getattr(sched, attr)} does not match runtime {what} {getattr(run, attr)}")


def compute_elastic_config(ds_config: dict, target_deepspeed_version: str, world_size=0, return_microbatch=False):
    """Returns ``(final_batch_size, valid_gpus[, micro_batch])`` exactly like the reference API."""
    if not isinstance(ds_config, dict):
        raise ValueError(f"Expected ds_config to be a dictionary but received a {type(ds_config)}, containing: {ds_config}")
    if C.ELASTICITY not in ds_config:
        raise ElasticityError(f"'{C.ELASTICITY}' is missing from config json, please add it if running an elastic "
                              f"training job.")
    ed = ds_config[C.ELASTICITY]
    if not ed.get(C.ENABLED, C.ENABLED_DEFAULT):
        raise ElasticityError("Elasticity is disabled, please enable it ('enabled':true) if running an elastic "
                              "training job.")
    cfg = ElasticityConfig(ed)
    mp, per_node = cfg.model_parallel_size, cfg.num_gpus_per_node
    if mp > 1 and float(cfg.version) != 0.2:
        raise ElasticityConfigError(f"Elasticity V{cfg.version} does not support model-parallel training. Given "
                                    f"model-parallel size: {mp}")
    if float(cfg.version) > C.LATEST_ELASTICITY_VERSION:
        raise ElasticityConfigError(f"Attempting to run elasticity version {cfg.version} but runtime only supports up "
                                    f"to {C.LATEST_ELASTICITY_VERSION}")
    from deepspeed_b200.git_version_info import version as _own_version
    # schedulers pass either an UPSTREAM release number (>= 0.3.8 introduced elasticity) or this package's own version
    if str(target_deepspeed_version) != str(_own_version) and \
            pkg_version.parse(target_deepspeed_version) < pkg_version.parse(C.MINIMUM_DEEPSPEED_VERSION):
        raise ElasticityError(f"Unable to run elasticity on target deepspeed version of {target_deepspeed_version}, "
                              f"currently {C.MINIMUM_DEEPSPEED_VERSION}+ is required")
    micro = None
    if float(cfg.version) == 0.1:
        batch, gpus = _plan_v01(cfg.micro_batches, cfg.max_acceptable_batch_size, cfg.min_gpus, cfg.max_gpus,
                                cfg.prefer_larger_batch_size)
        batch = int(batch)
    elif float(cfg.version) == 0.2:
        if world_size != 0:
            cur = world_size
        elif os.getenv("WORLD_SIZE", "").isnumeric():
            cur = int(os.environ["WORLD_SIZE"])
        else:
            raise ElasticityConfigError("Elasticity V 0.2 needs WORLD_SIZE to compute valid batch size. Either give it "
                                        "as argument to function compute_elastic_config or set it as an environment "
                                        "variable. Value of WORLD_SIZE as environment variable is "
                                        f"{os.getenv('WORLD_SIZE')}")
        batch, gpus, micro = _plan_v02(cfg.micro_batches, cfg.max_acceptable_batch_size, cur, cfg.min_gpus, cfg.max_gpus,
                                       cfg.prefer_larger_batch_size, per_node, mp)
        batch = int(batch)
    else:
        raise NotImplementedError(f"Unable to find elastic logic for version: {cfg.version}")
    logger.info(f"Valid World Size (GPUs / Model Parallel Size): {gpus}")
    if world_size > 0:
        if world_size not in gpus:
            raise ElasticityIncompatibleWorldSize(f"World size ({world_size}) is not valid with the current list of "
                                                  f"valid GPU counts: {gpus}")
        if micro is None:
            fits = [mb for mb in sorted(set(cfg.micro_batches), reverse=True) if (batch // world_size) % mb == 0]
            assert fits, f"Unable to find divisible micro batch size world_size={world_size}, " \
                         f"final_batch_size={batch}, and micro_batches={cfg.micro_batches}."
            micro = fits[0]
        return batch, gpus, micro
    if return_microbatch:
        if float(cfg.version) == 0.2:
            return batch, gpus, micro
        fits = [mb for mb in sorted(set(cfg.micro_batches), reverse=True) if batch % mb == 0]
        return batch, gpus, (fits[0] if fits else None)
    return batch, gpus


# ---- reference names of the planning primitives ---------------------------------------------------------------------
def get_candidate_batch_sizes(base_list, max_acceptable_batch_size):
    return _candidate_batches(list(base_list), max_acceptable_batch_size)


def get_valid_gpus(batch_size, micro_batches, min_valid_gpus, max_valid_gpus):
    return _valid_gpu_counts(batch_size, list(micro_batches), min_valid_gpus, max_valid_gpus)


def get_best_candidates(candidate_batch_sizes, micro_batches, min_gpus, max_gpus, prefer_larger):
    """(batch size, valid gpu counts) of the candidate that admits the most GPU counts (ties by batch size preference)."""
    best_batch, best_gpus, best_n = int(min(micro_batches)), None, 0
    for cand in candidate_batch_sizes:
        gpus = get_valid_gpus(cand, micro_batches, min_gpus, max_gpus)
        tie = (cand > best_batch) if prefer_larger else (cand < best_batch)
        if len(gpus) > best_n or (len(gpus) == best_n and tie):
            best_n, best_gpus, best_batch = len(gpus), gpus, cand
    return best_batch, best_gpus
