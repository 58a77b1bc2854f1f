"""param_groups_weight_decay (utils/mim_vit.py:126): no decay for biases and 1-d parameters."""


def param_groups_weight_decay(model, weight_decay=1e-5, no_weight_decay_list=()):
    decay, no_decay = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        (no_decay if p.ndim <= 1 or name.endswith(".bias") or name in no_weight_decay_list else decay).append(p)
    return [{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": weight_decay}]
