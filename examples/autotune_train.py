"""Minimal training script the autotuner can drive: a small Llama on synthetic tokens.
    python -m deepspeed_b200.launcher.runner --autotuning tune --num_gpus 1 examples/autotune_train.py \\
        --deepspeed_config examples/autotune_ds_config.json
The engine's autotuning hooks measure steps [start_profile_step, end_profile_step) and exit; outside an autotuning
experiment the loop simply trains for ``--steps`` steps."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a source checkout
import torch

import deepspeed_b200 as ds
from deepspeed_b200.models.llama import LlamaForCausalLM, llama_config


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--seq", type=int, default=512)
    ap.add_argument("--local_rank", type=int, default=0)
    ap = ds.add_config_arguments(ap)
    args = ap.parse_args()
    ds.init_distributed()
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    cfg = llama_config("tiny", hidden_size=1024, intermediate_size=2816, num_attention_heads=8, num_key_value_heads=8,
                       vocab_size=8192, num_hidden_layers=4)
    torch.manual_seed(0)
    with torch.device(dev):
        model = LlamaForCausalLM(cfg)
    engine, *_ = ds.initialize(args=args, model=model, model_parameters=model.parameters())
    mb = engine.train_micro_batch_size_per_gpu()
    g = torch.Generator().manual_seed(ds.comm.get_rank())
    for _ in range(args.steps):
        ids = torch.randint(0, cfg.vocab_size, (mb, args.seq), generator=g).to(dev)
        loss = engine(ids, labels=ids)
        engine.backward(loss)
        engine.step()
    if ds.comm.get_rank() == 0:
        print(f"final loss {loss.item():.4f}")


if __name__ == "__main__":
    main()
