"""Discrepancies between predicted and measured profiles (mentflow/loss.py:7-17).

These act on (B,) or (Bx, By) tensors -- O(K*B) work per step, after the particle
reduction -- and stay differentiable torch expressions so that any user-supplied
discrepancy function works unchanged."""
import torch


def mean_absolute_error(pred: torch.Tensor, targ: torch.Tensor) -> torch.Tensor:
    return torch.mean(torch.abs(pred - targ))


def mean_square_error(pred: torch.Tensor, targ: torch.Tensor) -> torch.Tensor:
    return torch.mean(torch.square(pred - targ))


def kl_divergence(pred: torch.Tensor, targ: torch.Tensor, pad: float = 1.0e-12) -> torch.Tensor:
    """sum targ * (log targ - log(pred + pad)) / pred.shape[0]  with 0 log 0 = 0
    (= F.kl_div(log(pred+pad), targ, 'batchmean'), loss.py:15-17)."""
    return torch.sum(torch.xlogy(targ, targ) - targ * torch.log(pred + pad)) / pred.shape[0]


def kl_divergence_batched(pred: torch.Tensor, targ: torch.Tensor, pad: float = 1.0e-12) -> torch.Tensor:
    """(K,) KL of K stacked profiles in one expression (first axis = profile index)."""
    k = pred.shape[0]
    terms = torch.xlogy(targ, targ) - targ * torch.log(pred + pad)
    return terms.reshape(k, -1).sum(dim=1) / pred.shape[1]
