"""Shared helpers: seeded inputs / weights identical to oracle/pin_against_reference.py, golden loading."""
import json
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def seeded(shape, seed, scale=1.0):
    return torch.randn(tuple(shape), generator=torch.Generator().manual_seed(seed)) * scale


def schema(name):
    with open(os.path.join(GOLD, f"schema_{name}.json")) as f:
        return {k: tuple(v) for k, v in json.load(f).items()}


def golden(name):
    return torch.load(os.path.join(GOLD, name), map_location="cpu", weights_only=False)


def err_stats(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    diff = (got - ref)
    rel_l2 = (diff.norm() / ref.norm().clamp_min(1e-12)).item()
    return dict(max_abs=diff.abs().max().item(), max_ref=ref.abs().max().item(), rel_l2=rel_l2)
