"""Registry of injection policies (reference ``module_inject/replace_policy.py``)."""
from .containers import (BLOOMLayerPolicy, GPTNEOXLayerPolicy, HFBertLayerPolicy, HFCLIPLayerPolicy, HFDistilBertLayerPolicy,
                         HFGPT2LayerPolicy, HFGPTJLayerPolicy, HFGPTNEOLayerPolicy, HFOPTLayerPolicy, InternLMLayerPolicy,
                         LLAMA2LayerPolicy, LLAMALayerPolicy, MegatronLayerPolicy, MegatronMoELayerPolicy, UNetPolicy, VAEPolicy)
from .containers import (DS_BERTContainer, DS_BloomContainer, DS_CLIPContainer, DS_DistilBERTContainer, DS_GPT2Container,
                         DS_GPTJContainer, DS_GPTNEOContainer, DS_GPTNEOXContainer, DS_InternLMContainer, DS_LLAMA2Container,
                         DS_LLAMAContainer, DS_MegatronGPTContainer, DS_MegatronGPTMoEContainer, DS_OPTContainer)

# transformer-layer policies (checked in order) and the container each one builds
replace_policies = [HFBertLayerPolicy, HFGPTNEOLayerPolicy, GPTNEOXLayerPolicy, HFGPTJLayerPolicy, MegatronMoELayerPolicy,
                    MegatronLayerPolicy, HFGPT2LayerPolicy, BLOOMLayerPolicy, HFOPTLayerPolicy, HFCLIPLayerPolicy,
                    HFDistilBertLayerPolicy, LLAMALayerPolicy, LLAMA2LayerPolicy, InternLMLayerPolicy]
policy_to_ds_container = {
    HFBertLayerPolicy: DS_BERTContainer, HFGPTNEOLayerPolicy: DS_GPTNEOContainer, GPTNEOXLayerPolicy: DS_GPTNEOXContainer,
    HFGPTJLayerPolicy: DS_GPTJContainer, MegatronLayerPolicy: DS_MegatronGPTContainer,
    MegatronMoELayerPolicy: DS_MegatronGPTMoEContainer, HFGPT2LayerPolicy: DS_GPT2Container, BLOOMLayerPolicy: DS_BloomContainer,
    HFOPTLayerPolicy: DS_OPTContainer, HFCLIPLayerPolicy: DS_CLIPContainer, HFDistilBertLayerPolicy: DS_DistilBERTContainer,
    LLAMALayerPolicy: DS_LLAMAContainer, LLAMA2LayerPolicy: DS_LLAMA2Container, InternLMLayerPolicy: DS_InternLMContainer,
}
# whole-module policies (diffusers)
generic_policies = [UNetPolicy, VAEPolicy]


def policy_for(module):
    """The registered policy class that recognises ``module``: exact class (or listed alias) first, structural
    ``matches`` of the class-less policies (Megatron, Meta-llama, remote-code models) only when no class claims it."""
    for pol in replace_policies:
        classes = [c for c in [getattr(pol, "_orig_layer_class", None), *getattr(pol, "_also", [])] if c is not None]
        if any(type(module) is c for c in classes):
            return pol
    if type(module).__module__.startswith("transformers."):
        return None  # a Hugging Face layer no policy lists: do not guess by structure
    for pol in replace_policies:
        m = getattr(pol, "matches", None)
        if m is not None and getattr(pol, "_orig_layer_class", None) is None and m(module):
            return pol
    return None
