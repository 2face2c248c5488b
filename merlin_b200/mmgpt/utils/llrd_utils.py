"""Layer-wise LR decay multipliers + AdamW param grouping — mirror of mmgpt/utils/llrd_utils.py:4-79.

Same rules: parameters whose name ends in ".bias" or that are 1-D get no weight decay; with `llrd` the
CLIP encoder layer i is scaled by 0.9 ** (23 - i - 1) and every other `vision_model` parameter by 0.1
(llm_llrd: decoder layer i by 0.931 ** (32 - i - 1)).  Instead of returning torch param-group dicts this
returns per-parameter (lr_mult, weight_decay), which the fused AdamW kernel consumes as per-chunk tables
over the flat parameter arena; `get_param_groups` reproduces the reference's group list for inspection.
"""
import re


def vit_lr_scale_func(key):
    if "vision_model.encoder.layers" in key:
        in_pp_layer = int(re.findall(r"layers\.(\d+)\.", key)[0])
        return 0.9 ** (23 - in_pp_layer - 1)
    elif "vision_model" in key:
        return 0.1
    return 1


def llm_lr_scale_func(key):
    if "model.layers" in key:
        in_pp_layer = int(re.findall(r"layers\.(\d+)\.", key)[0])
        return 0.931 ** (32 - in_pp_layer - 1)
    return 1


def param_hyper(name, shape, scale_lr_cond, wd):
    """(lr multiplier, weight decay) of one parameter under the reference's grouping rules."""
    no_wd = name.endswith(".bias") or len(shape) == 1
    lr_mult = scale_lr_cond(name) if scale_lr_cond is not None else 1
    return float(lr_mult), (0.0 if no_wd else float(wd))


def get_param_groups(named_shapes, no_weight_decay_cond, scale_lr_cond, lr, wd):
    """Reference group order (llrd_utils.py:25-79): wd & unscaled, wd & scaled (one group per multiplier),
    no-wd & unscaled, no-wd & scaled.  `named_shapes`: iterable of (name, shape).  Each group carries the
    parameter names instead of tensors."""
    wd_no_scale, wd_scale, no_wd_no_scale, no_wd_scale = [], {}, [], {}
    for name, shape in named_shapes:
        if no_weight_decay_cond is not None:
            no_wd = no_weight_decay_cond(name, shape)
        else:
            no_wd = name.endswith(".bias") or len(shape) == 1
        lr_mult = scale_lr_cond(name) if scale_lr_cond is not None else 1
        scale_lr = lr_mult != 1
        if not no_wd and not scale_lr:
            wd_no_scale.append(name)
        elif not no_wd and scale_lr:
            wd_scale.setdefault(lr_mult, []).append(name)
        elif no_wd and not scale_lr:
            no_wd_no_scale.append(name)
        else:
            no_wd_scale.setdefault(lr_mult, []).append(name)
    groups = []
    if wd_no_scale:
        groups.append({"params": wd_no_scale, "weight_decay": wd, "lr": lr})
    for lr_mult, names in wd_scale.items():
        groups.append({"params": names, "weight_decay": wd, "lr": lr * lr_mult})
    if no_wd_no_scale:
        groups.append({"params": no_wd_no_scale, "weight_decay": 0.0, "lr": lr})
    for lr_mult, names in no_wd_scale.items():
        groups.append({"params": names, "weight_decay": 0.0, "lr": lr * lr_mult})
    return groups
