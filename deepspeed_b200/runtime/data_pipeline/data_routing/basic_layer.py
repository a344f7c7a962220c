"""Random layerwise token dropping wrapper (reference ``data_routing/basic_layer.py:14 RandomLayerTokenDrop``):
during training the wrapped layer only sees a random sorted subset of tokens; its outputs are scattered back into
the full sequence so skipped tokens pass through unchanged."""
import torch
from torch import nn

from deepspeed_b200.ops.random_ltd import GatherTokens, ScatterTokens, bert_sample_tokens, gpt_sample_tokens


class RandomLayerTokenDrop(nn.Module):

    def __init__(self, layer: nn.Module):
        super().__init__()
        self.random_ltd_layer = layer
        self.reserved_length = None
        self.random_ltd_scheduler = None
        self.max_length = None
        self.curr_seq = -1
        self.batch_first = False
        self.random_ltd_layer_id = 0
        self.random_ltd_num_layer = 1
        self.mask_name = None
        self.index_generator = gpt_sample_tokens
        self.model_type = "decoder"

    def init_config(self, config, scheduler, random_ltd_layer_id):
        from .. import constants as C
        self.random_ltd_scheduler = scheduler
        self.random_ltd_layer_id = random_ltd_layer_id
        self.max_length = scheduler.state[C.RANDOM_LTD_MAX_VALUE]
        self.mask_name = config[C.RANDOM_LTD_MODEL_MASK_NAME]
        self.micro_bs = config[C.RANDOM_LTD_MICRO_BATCH_SIZE]
        self.random_ltd_num_layer = scheduler.random_ltd_layer_num
        order = config[C.RANDOM_LTD_HIDDEN_STATE_ORDER]
        self.batch_first = order == "batch_seq_dim"
        if order not in ("batch_seq_dim", "seq_batch_dim"):
            raise NotImplementedError(f"hidden_state_order {order} is not supported")
        self.model_type = config[C.RANDOM_LTD_MODEL_TYPE]
        self.index_generator = bert_sample_tokens if self.model_type == "encoder" else gpt_sample_tokens

    def get_bsh(self, hidden_states):
        self.curr_seq, self.curr_micro_batch = hidden_states.size()[1], hidden_states.size()[0]

    def get_sbh(self, hidden_states):
        self.curr_seq, self.curr_micro_batch = hidden_states.size()[0], hidden_states.size()[1]

    def forward(self, hidden_states, **kwargs) -> torch.Tensor:
        if self.random_ltd_scheduler is not None:
            self.reserved_length = self.random_ltd_scheduler.get_current_seq()
            (self.get_bsh if self.batch_first else self.get_sbh)(hidden_states)
        if self.training and self.random_ltd_scheduler is not None and self.reserved_length < self.curr_seq:
            mask = kwargs.get(self.mask_name) if self.mask_name is not None else None
            if self.random_ltd_layer_id == 0:
                idx, part_mask = self.index_generator(self.reserved_length, self.curr_seq, self.curr_micro_batch,
                                                      self.random_ltd_num_layer, hidden_states.device, mask)
                self.random_ltd_scheduler.state["sample_idx"] = idx
                self.random_ltd_scheduler.state["attention_mask"] = part_mask
            else:
                idx = self.random_ltd_scheduler.state["sample_idx"]
                part_mask = self.random_ltd_scheduler.state["attention_mask"]
            my_idx = idx[self.random_ltd_layer_id]
            hidden_states, part = GatherTokens.apply(hidden_states, my_idx, self.batch_first)
            if self.mask_name is not None:
                if self.model_type == "encoder":
                    kwargs[self.mask_name] = part_mask[self.random_ltd_layer_id]
                elif part_mask is not None:
                    kwargs[self.mask_name] = part_mask
            out = self.random_ltd_layer(part, **kwargs)
            if isinstance(out, tuple):
                full = ScatterTokens.apply(hidden_states, out[0], my_idx, self.batch_first)
                return (full, ) + tuple(out[1:])
            return ScatterTokens.apply(hidden_states, out, my_idx, self.batch_first)
        return self.random_ltd_layer(hidden_states, **kwargs)
