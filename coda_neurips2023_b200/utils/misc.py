"""Small loss helpers (reference utils/misc.py:25-36)."""
import torch


def huber_loss(error, delta: float = 1.0):
    """0.5 x^2 for |x| <= delta, 0.5 delta^2 + delta (|x| - delta) beyond."""
    abs_error = torch.abs(error)
    quadratic = torch.clamp(abs_error, max=delta)
    linear = abs_error - quadratic
    return 0.5 * quadratic ** 2 + delta * linear
