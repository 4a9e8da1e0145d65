"""Import the UNMODIFIED reference (read-only at /root/reference) in the build
container.  Test infrastructure only; used by oracle/make_golden.py to pin the
oracle.  /root/reference does not exist on the GPU box, so nothing under
tests/ -m gpu, smoke() or bench.py imports this module.

emcee, corner and matplotlib are imported at module top by the reference
(inference.py:15-19, functions.py:10,13) but never touched on the log-prob
path; they are absent from this image, so empty stand-in modules are
registered before the import."""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("CHA1_REFERENCE_ROOT", "/root/reference")


def load():
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError(f"reference tree not present at {REF_ROOT}")
    for name in ("emcee", "corner", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import inference  # noqa: E402  (the reference's top-level module)
    from spectral_simulator import classes, functions, constants  # noqa: E402
    return types.SimpleNamespace(inference=inference, classes=classes, functions=functions,
                                 constants=constants)


def load_tmc1():
    """scripts/MCMC/TMC1_four_component.py as a module (it is a script, not a package)."""
    load()
    path = os.path.join(REF_ROOT, "scripts", "MCMC", "TMC1_four_component.py")
    spec = importlib.util.spec_from_file_location("ref_tmc1_four_component", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
