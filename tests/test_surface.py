"""CPU: the package mirrors the reference's call surface for the path -- same names, same
parameter order, same defaults.  tests/golden/surface.json holds the reference's signatures (written by
tests/golden/make_golden.py --only surface from the unmodified reference).  Allowed differences, both documented in
the docstrings: extra trailing keyword parameters (``device``), and ``torch_dtype=None`` standing for upstream's
``torch.float32`` default (torch is imported lazily here)."""
import importlib
import inspect
import json
import os

import pytest

from conftest import GOLDEN

with open(os.path.join(GOLDEN, 'surface.json')) as _fh:
    SURFACE = json.load(_fh)
CASES = [(m, name) for m in sorted(SURFACE) for name in sorted(SURFACE[m])]


@pytest.mark.parametrize('module,name', CASES)
def test_same_parameters_and_defaults(module, name):
    obj = importlib.import_module(f'mc3d_b200.{module}')
    for part in name.split('.'):
        assert hasattr(obj, part), f'{module}.{name} is missing'
        obj = getattr(obj, part)
    mine = [[p.name, None if p.default is inspect.Parameter.empty else repr(p.default)]
            for p in inspect.signature(obj).parameters.values()]
    want = SURFACE[module][name]
    assert [p[0] for p in mine[:len(want)]] == [p[0] for p in want], (mine, want)
    for (pname, got), (_, ref) in zip(mine, want):
        if pname == 'torch_dtype' and ref == 'torch.float32':
            assert got == 'None'                                   # resolved to torch.float32 inside the function
        else:
            assert got == ref, (pname, got, ref)
    for pname, default in mine[len(want):]:
        assert default is not None, f'extra parameter {pname} of {module}.{name} must be optional'
