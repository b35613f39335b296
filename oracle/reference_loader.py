"""Load the REAL reference code from ``/root/reference`` (build container only).

TEST INFRASTRUCTURE ONLY.  Used by ``tests/golden/gen_golden.py`` to produce the
committed fixtures and by the ``-m "not gpu"`` tests that cross-check the oracle
against the live reference when the checkout is present.  ``/root/reference``
does not exist on the GPU box: nothing in the ``-m gpu`` tests, ``smoke()`` or
``bench.py`` calls into this module.

Nothing is copied: modules are compiled from the sources where they lie.

* ``common/controller6.py:4`` imports ``isaacgym.gymapi`` and never uses it -> a
  stub module is placed in ``sys.modules`` for the duration of the import.
* ``strip_prints=True`` drops ``print(`` statement lines before compiling (28 in
  ``secondary_control_vecenv.py``, 6 in ``world2pixel``).  Needed for N > ~1e4:
  the debug print at ``common/secondary_control_vecenv.py:174`` builds an
  (N,N,1) array.  Arithmetic lines are untouched.
* ``examples/franka_cube_ik_osc.py`` cannot be imported (module level creates a
  simulator, :88-148); its pure-torch functions are AST-extracted.
"""
from __future__ import annotations

import ast
import contextlib
import io
import os
import re
import sys
import types

REFERENCE_ROOT = os.environ.get("B200CTL_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "common", "controller6.py"))


_PRINT_LINE = re.compile(r"^\s*print\(.*\)\s*$")


def _filtered_source(path: str, strip_prints: bool) -> str:
    with open(path, encoding="utf-8") as fh:
        lines = fh.read().split("\n")
    if strip_prints:
        # a print statement line is replaced by `pass` at the same indent so blocks stay valid
        lines = [re.sub(r"print\(.*$", "pass", ln) if _PRINT_LINE.match(ln) else ln for ln in lines]
    return "\n".join(lines)


@contextlib.contextmanager
def _isaacgym_stub():
    had = "isaacgym" in sys.modules
    if not had:
        stub = types.ModuleType("isaacgym")
        stub.gymapi = object()
        sys.modules["isaacgym"] = stub
    try:
        yield
    finally:
        if not had:
            sys.modules.pop("isaacgym", None)


def load_common(name: str, strip_prints: bool = False) -> types.ModuleType:
    """Compile ``common/<name>.py`` from the reference checkout into a fresh module."""
    path = os.path.join(REFERENCE_ROOT, "common", name + ".py")
    mod = types.ModuleType("reference_" + name + ("_quiet" if strip_prints else ""))
    mod.__file__ = path
    code = compile(_filtered_source(path, strip_prints), path, "exec")
    with _isaacgym_stub():
        exec(code, mod.__dict__)   # noqa: S102 - executing the reference is the point
    return mod


@contextlib.contextmanager
def quiet():
    """Swallow the reference's debug prints."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def extract_functions(rel_path: str, names, namespace: dict) -> dict:
    """AST-extract top-level ``def``s from a reference script and exec them in ``namespace``."""
    path = os.path.join(REFERENCE_ROOT, rel_path)
    with open(path, encoding="utf-8") as fh:
        tree = ast.parse(fh.read(), filename=path)
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in set(names)]
    missing = set(names) - {n.name for n in wanted}
    if missing:
        raise KeyError(f"{rel_path}: no top-level def for {sorted(missing)}")
    module = ast.Module(body=wanted, type_ignores=[])
    exec(compile(module, path, "exec"), namespace)   # noqa: S102
    return namespace


def franka_namespace(dtype, **globals_):
    """Namespace for ``control_ik`` / ``control_osc`` / ``orientation_error`` of
    ``examples/franka_cube_ik_osc.py:34-79`` with restated ``isaacgym.torch_utils``
    quaternion helpers (un-installable dependency; see ``oracle/franka.py``)."""
    import numpy as np
    import torch
    from . import franka as _f
    ns = {"torch": torch, "np": np, "math": __import__("math"), "device": "cpu",
          "quat_mul": _f.quat_mul, "quat_conjugate": _f.quat_conjugate}
    ns.update(globals_)
    extract_functions("examples/franka_cube_ik_osc.py",
                      ["orientation_error", "control_ik", "control_osc"], ns)
    return ns
