"""Small helpers shared by the injection code (reference ``module_inject/utils.py``)."""
from deepspeed_b200.utils import log_dist

_POLICY_FOR_MODEL_TYPE = None


def _table():
    global _POLICY_FOR_MODEL_TYPE
    if _POLICY_FOR_MODEL_TYPE is None:
        from . import containers as C
        _POLICY_FOR_MODEL_TYPE = {
            "bert": C.HFBertLayerPolicy, "roberta": C.HFBertLayerPolicy, "distilbert": C.HFDistilBertLayerPolicy,
            "gpt2": C.HFGPT2LayerPolicy, "gptj": C.HFGPTJLayerPolicy, "gpt_neo": C.HFGPTNEOLayerPolicy,
            "gpt_neox": C.GPTNEOXLayerPolicy, "opt": C.HFOPTLayerPolicy, "bloom": C.BLOOMLayerPolicy, "llama": C.LLAMALayerPolicy,
            "mistral": C.LLAMALayerPolicy, "qwen2": C.LLAMALayerPolicy, "internlm": C.InternLMLayerPolicy,
            "clip": C.HFCLIPLayerPolicy, "megatron": C.MegatronLayerPolicy}
    return _POLICY_FOR_MODEL_TYPE


def policy_to_ds_container(**kwargs):
    """Build the container registered for ``kwargs["policy"]`` (reference helper of the same name)."""
    from .replace_policy import policy_to_ds_container as table
    policy = kwargs["policy"]
    cls = table.get(type(policy))
    if cls is None:
        log_dist(f"Policy type {type(policy)} not supported", [0])
        return None
    return cls(**kwargs)


def policy_for_model_type(model_type):
    return _table().get(model_type)
