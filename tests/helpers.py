"""Shared helpers for the test-suite (oracle side)."""
import torch

from oracle import emap_oracle as O
from tests.conftest import load_golden


def oracle_params(pert: bool, multires: int = 10) -> O.UDFParams:
    name = "net_init_state" if multires == 10 else "net_init_state_mr6"
    p = O.UDFParams.from_state_dict(load_golden(name), multires=multires)
    return O.perturbed_params(p) if pert else p


def oracle_scalars() -> O.ScalarParams:
    s = load_golden("scalars")
    return O.ScalarParams(s["variance"].clone(), s["beta_raw"].clone(), s["gamma_raw"].clone())


def maxdiff(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.detach().double() - b.detach().double()).abs().max())


def relerr(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / (max|b| + tiny): error relative to the tensor's scale."""
    return maxdiff(a, b) / (float(b.double().abs().max()) + 1e-30)
