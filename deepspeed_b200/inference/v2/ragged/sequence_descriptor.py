"""Per-sequence tracking state (reference ``ragged/sequence_descriptor.py``)."""
from typing import List, Tuple, Union

import torch


class BaseSequenceDescriptor:

    @property
    def seen_tokens(self) -> int:
        raise NotImplementedError

    @property
    def cur_allocated_blocks(self) -> int:
        raise NotImplementedError


class PlaceholderSequenceDescriptor(BaseSequenceDescriptor):

    def __init__(self, seen_tokens=0, cur_allocated_blocks=0, kv_blocks_ptr=0):
        self._seen_tokens, self._cur_allocated_blocks = seen_tokens, cur_allocated_blocks

    @property
    def seen_tokens(self):
        return self._seen_tokens

    @property
    def cur_allocated_blocks(self, cache_group: int = 0):
        return self._cur_allocated_blocks


class DSSequenceDescriptor(BaseSequenceDescriptor):

    def __init__(self, tracking_id: int, kv_cache_ids: Tuple[torch.Tensor, ...], max_context: int = -1):
        self._tracking_id = tracking_id
        self._kv_cache_ids = kv_cache_ids          # per cache group: int32 [max_blocks] host tensors
        self._kv_np = tuple(t.numpy() for t in kv_cache_ids)  # same memory, cheap host-side access
        self._blocks_per = [0 for _ in kv_cache_ids]
        self._seen_tokens = 0
        self._in_flight_tokens = 0
        self._max_context = max_context
        self.host_kv = None                         # offloaded blocks (state-manager offload)

    @property
    def seen_tokens(self) -> int:
        return self._seen_tokens

    @property
    def in_flight_tokens(self) -> int:
        return self._in_flight_tokens

    @property
    def max_context(self) -> int:
        return self._max_context

    @property
    def tracking_id(self):
        return self._tracking_id

    def cur_allocated_blocks_of(self, cache_group: int = 0) -> int:
        return self._blocks_per[cache_group]

    @property
    def cur_allocated_blocks(self) -> int:
        return self._blocks_per[0]

    def kv_cache_ids(self, cache_group: int = 0) -> torch.Tensor:
        return self._kv_cache_ids[cache_group]

    def kv_ids_np(self, cache_group: int = 0):
        return self._kv_np[cache_group]

    def all_block_ids(self, cache_group: int = 0) -> torch.Tensor:
        return self._kv_cache_ids[cache_group][:self._blocks_per[cache_group]]

    def pre_forward(self, num_tokens: int) -> None:
        self._in_flight_tokens = num_tokens

    def post_forward(self) -> None:
        self._seen_tokens += self._in_flight_tokens
        self._in_flight_tokens = 0

    def extend_kv_cache(self, new_ids: Union[List[torch.IntTensor], torch.IntTensor], cache_group: int = 0) -> None:
        if isinstance(new_ids, torch.Tensor):
            new_ids = [new_ids]
        ids = new_ids[0] if len(new_ids) == 1 else torch.cat(list(new_ids))
        n = ids.numel()
        s = self._blocks_per[cache_group]
        self._kv_np[cache_group][s:s + n] = ids.numpy() if ids.device.type == "cpu" else ids.cpu().numpy()
        self._blocks_per[cache_group] += n

    def free_kv_cache(self, free_ids, cache_group: int = 0) -> None:
        raise NotImplementedError("Partial KV-cache freeing is not supported (matches the reference).")
