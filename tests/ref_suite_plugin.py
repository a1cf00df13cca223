"""pytest plugin used by tests/test_reference_suite.py: makes the reference's own test file import the B200 modules and
build its tensors on the GPU.  `from src.models.two_tower import ...` resolves to b200rec.two_tower (sys.modules is
consulted before the path the test file inserts), and torch's default device is CUDA, so the bare torch.randn(...) /
nn.Module constructions of the reference tests (written for CPU) produce CUDA tensors — SURVEY.md section 7: run them
on the GPU box rather than adding a CPU path to the product."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    import torch
    import b200rec.two_tower as tt
    for name in ("src", "src.models"):
        mod = types.ModuleType(name)
        mod.__path__ = []
        sys.modules[name] = mod
    sys.modules["src.models.two_tower"] = tt
    sys.modules["src.models"].two_tower = tt
    torch.set_default_device("cuda")
