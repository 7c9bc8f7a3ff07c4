"""Learning-rate schedules the training scripts pick by name (drop-in for /root/reference/utils/lrschedule.py): host arithmetic on
Python floats; ``FusedAdam`` picks a changed ``param_groups[0]["lr"]`` up into its device-side scalar before the next step."""
import math


def noam_learning_rate_decay(init_lr, global_step, warmup_steps=2000):
    """Noam scheme: linear warm-up for ``warmup_steps`` steps, then 1/sqrt(step) decay."""
    warm = float(warmup_steps)
    step = global_step + 1.0
    return init_lr * math.sqrt(warm) * min(step * warm ** -1.5, step ** -0.5)


def step_learning_rate_decay(init_lr, global_step, anneal_rate=0.98, anneal_interval=50000):
    return init_lr * anneal_rate ** (global_step // anneal_interval)


def cyclic_cosine_annealing(init_lr, global_step, T, M):
    """Cyclic cosine annealing (snapshot ensembles): T total iterations, M cycles."""
    period = T // M
    return init_lr / 2.0 * (math.cos(math.pi * ((global_step - 1) % period) / period) + 1.0)
