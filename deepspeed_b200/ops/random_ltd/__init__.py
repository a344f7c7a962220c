from .dropping_utils import gpt_sample_tokens, bert_sample_tokens, GatherTokens, ScatterTokens  # noqa: F401
