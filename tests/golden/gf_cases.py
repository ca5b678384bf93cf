"""Seeded guided-filter cases shared by the golden generator (make_golden_guided_filter.py) and the tests."""
import torch

CASES = (("a", (2, 3, 20, 22), 4), ("b", (1, 3, 40, 36), 16), ("c", (1, 1, 12, 40), 5))


def inputs(shape, r):
    g = torch.Generator().manual_seed(1000 + r)
    return (torch.rand(*shape, generator=g, dtype=torch.float64), torch.rand(*shape, generator=g, dtype=torch.float64),
            torch.randn(*shape, generator=g, dtype=torch.float64))


def fast_inputs():
    g = torch.Generator().manual_seed(5)
    return (torch.rand(1, 3, 20, 24, generator=g, dtype=torch.float64), torch.rand(1, 3, 20, 24, generator=g, dtype=torch.float64),
            torch.rand(1, 3, 40, 48, generator=g, dtype=torch.float64))
