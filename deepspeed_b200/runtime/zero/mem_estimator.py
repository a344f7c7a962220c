"""Model-state memory estimators (reference ``runtime/zero/stage_1_and_2.py`` / ``stage3.py`` module-level
``estimate_zero{2,3}_model_states_mem_needs*``).  Byte accounting: bf16/fp16 params 2 B + grads 2 B, fp32 master 4 B
+ Adam moments 8 B (16 B total optimizer+master per parameter; 18 with fp16 copies)."""
import torch


def _model_numbers(model):
    total = sum(getattr(p, "ds_numel", p.numel()) for p in model.parameters())
    largest = 0
    for m in model.modules():
        n = sum(getattr(p, "ds_numel", p.numel()) for p in m.parameters(recurse=False))
        largest = max(largest, n)
    return total, largest


def estimate_zero2_model_states_mem_needs(total_params, num_gpus_per_node=1, num_nodes=1, cpu_offload=True,
                                          additional_buffer_factor=1.5):
    total_gpus = num_nodes * num_gpus_per_node
    if cpu_offload:
        gpu_mem = 2 * total_params
        cpu_mem = total_params * max(4 * total_gpus, 16) * additional_buffer_factor
    else:
        gpu_mem = 4 * total_params + int(16 * total_params / total_gpus)
        cpu_mem = total_params * 4 * num_gpus_per_node * additional_buffer_factor
    return int(cpu_mem), int(gpu_mem)


def estimate_zero3_model_states_mem_needs(total_params, largest_layer_params, num_gpus_per_node=1, num_nodes=1,
                                          cpu_offload=True, cpu_offload_params=True, zero_init=True,
                                          additional_buffer_factor=1.5):
    total_gpus = num_nodes * num_gpus_per_node
    gpus_factor = 1 / num_nodes
    largest_layer_memory = 4 * largest_layer_params
    if cpu_offload:
        if cpu_offload_params:
            gpu_mem = largest_layer_memory
            cpu_mem = total_params * 18 * gpus_factor * additional_buffer_factor if zero_init else \
                total_params * max(4 * num_gpus_per_node, 18 * gpus_factor) * additional_buffer_factor
        else:
            gpu_mem = largest_layer_memory + int(2 * total_params / total_gpus)
            cpu_mem = total_params * 16 * gpus_factor * additional_buffer_factor if zero_init else \
                total_params * max(4 * num_gpus_per_node, 16 * gpus_factor) * additional_buffer_factor
    else:
        gpu_mem = largest_layer_memory + int(18 * total_params / total_gpus)
        cpu_mem = largest_layer_params * 4 * num_gpus_per_node * additional_buffer_factor if zero_init else \
            total_params * 4 * num_gpus_per_node * additional_buffer_factor
    return int(cpu_mem), int(gpu_mem), largest_layer_memory


def _fmt(b):
    return f"{b / 2**30:7.2f}GB"


def estimate_zero2_model_states_mem_needs_all_cold(total_params, num_gpus_per_node=1, num_nodes=1,
                                                   additional_buffer_factor=1.5):
    print(f"Estimated memory needed for params, optim states and gradients for a:\nHW: Setup with {num_nodes} node"
          f"{'s' if num_nodes > 1 else ''}, {num_gpus_per_node} GPU{'s' if num_gpus_per_node > 1 else ''} per node.\n"
          f"SW: Model with {int(total_params / 1e6)}M total params.")
    print("  per CPU  |  per GPU |   Options")
    for off in (True, False):
        cpu, gpu = estimate_zero2_model_states_mem_needs(total_params, num_gpus_per_node, num_nodes, off,
                                                         additional_buffer_factor)
        print(f" {_fmt(cpu)} | {_fmt(gpu)} | offload_optimizer={'cpu' if off else 'none'}")


def estimate_zero2_model_states_mem_needs_all_live(model, num_gpus_per_node=1, num_nodes=1, additional_buffer_factor=1.5):
    total, _ = _model_numbers(model)
    estimate_zero2_model_states_mem_needs_all_cold(total, num_gpus_per_node, num_nodes, additional_buffer_factor)


def estimate_zero3_model_states_mem_needs_all_cold(total_params, largest_layer_params, num_gpus_per_node=1, num_nodes=1,
                                                   additional_buffer_factor=1.5):
    print(f"Estimated memory needed for params, optim states and gradients for a:\nHW: Setup with {num_nodes} node"
          f"{'s' if num_nodes > 1 else ''}, {num_gpus_per_node} GPU{'s' if num_gpus_per_node > 1 else ''} per node.\n"
          f"SW: Model with {int(total_params / 1e6)}M total params, {int(largest_layer_params / 1e6)}M largest layer "
          f"params.")
    print("  per CPU  |  per GPU |   Options")
    for off_p, off_o in ((True, True), (False, True), (False, False)):
        if off_p and not off_o:
            continue
        for zi in (True, False):
            cpu, gpu, _ = estimate_zero3_model_states_mem_needs(total_params, largest_layer_params, num_gpus_per_node,
                                                                num_nodes, off_o, off_p, zi, additional_buffer_factor)
            print(f" {_fmt(cpu)} | {_fmt(gpu)} | offload_param={'cpu' if off_p else 'none'}, "
                  f"offload_optimizer={'cpu' if off_o else 'none'}, zero_init={int(zi)}")


def estimate_zero3_model_states_mem_needs_all_live(model, num_gpus_per_node=1, num_nodes=1, additional_buffer_factor=1.5):
    total, largest = _model_numbers(model)
    estimate_zero3_model_states_mem_needs_all_cold(total, largest, num_gpus_per_node, num_nodes, additional_buffer_factor)
