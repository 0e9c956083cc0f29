"""Layer-wise learning-rate-decay parameter groups (reference audiomae/util/lr_decay.py:15-75, BEiT's scheme).

Same grouping rule and group keys as the reference, so ``lr_sched.adjust_learning_rate`` style code (which writes
``group["lr"] = lr * group["lr_scale"]``, util/lr_sched.py:9-21) drives either ``torch.optim.AdamW`` or
``tpat.optim.FusedAdamW`` built on these groups."""
from typing import Iterable, List


def get_layer_id_for_vit(name: str, num_layers: int) -> int:
    """cls_token / pos_embed / patch_embed -> 0, blocks.i -> i + 1, everything else (fc_norm, head) -> num_layers."""
    if name in ("cls_token", "pos_embed") or name.startswith("patch_embed"):
        return 0
    if name.startswith("blocks"):
        return int(name.split(".")[1]) + 1
    return num_layers


def param_groups_lrd(model, weight_decay: float = 0.05, no_weight_decay_list: Iterable[str] = (), layer_decay: float = 0.75
                     ) -> List[dict]:
    no_decay = set(no_weight_decay_list)
    num_layers = len(model.blocks) + 1
    layer_scales = [layer_decay ** (num_layers - i) for i in range(num_layers + 1)]
    groups = {}
    for n, p in model.named_parameters():
        if not p.requires_grad:
            continue
        decay = not (p.ndim == 1 or n in no_decay)          # no decay: all 1-D parameters and the model's own list
        layer_id = get_layer_id_for_vit(n, num_layers)
        key = "layer_%d_%s" % (layer_id, "decay" if decay else "no_decay")
        if key not in groups:
            groups[key] = {"lr_scale": layer_scales[layer_id], "weight_decay": weight_decay if decay else 0.0, "params": []}
        groups[key]["params"].append(p)
    return list(groups.values())
