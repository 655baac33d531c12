"""Import the UNMODIFIED reference solver from /root/reference (dev container only).

TEST INFRASTRUCTURE - not product code.  Only `oracle/make_golden.py` and the
oracle-pinning tests (which skip when /root/reference is absent, i.e. on the GPU
box) use this module.

The reference's `source/admm.py:6-11` and `source/parafac_epc.py:3-9` import
tensorly 0.4.5 and musco-pytorch 1.0.6 at module top; neither is installed here
(and there is no network).  Registering empty stub modules lets
`admm_iteration`, `init_factors('random'|'svd')`, `squared_relative_diff` and
`quantize_tensor` be imported and run unmodified; the `parafac`/`parafac-epc`
branches hit the stubs and raise.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ADMMQ_REFERENCE_ROOT", "/root/reference")
# Where /root/reference does not exist (the GPU box): the byte-compiled build product of oracle/build_ref.py
COMPILED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def reference_root():
    """Directory to import the reference's `source` package from: the sources themselves in the dev container, else
    the sourceless bytecode package oracle/_ref (made from those sources by oracle/build_ref.py), else None."""
    if os.path.isfile(os.path.join(REFERENCE_ROOT, "source", "admm.py")):
        return REFERENCE_ROOT
    if os.path.isfile(os.path.join(COMPILED_ROOT, "source", "admm.pyc.bin")):
        return COMPILED_ROOT
    return None


def reference_available() -> bool:
    return reference_root() is not None


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def _missing(*_a, **_k):
    raise RuntimeError("third-party routine (tensorly/musco) is not available in this container")


def _load_compiled():
    """The reference's `source` package from the bytecode files of oracle/build_ref.py: a package module plus the four
    hot-path modules executed in dependency order (source.admm imports the other three)."""
    import marshal
    pkg = types.ModuleType("source")
    pkg.__path__ = [os.path.join(COMPILED_ROOT, "source")]
    sys.modules["source"] = pkg
    mods = {}
    for name in ("utils", "quantization", "parafac_epc", "admm"):
        with open(os.path.join(COMPILED_ROOT, "source", name + ".pyc.bin"), "rb") as f:
            code = marshal.loads(f.read()[16:])      # 16-byte .pyc header: magic, flags, mtime, size
        mod = types.ModuleType("source." + name)
        mod.__package__ = "source"
        mod.__file__ = code.co_filename
        sys.modules["source." + name] = mod
        exec(code, mod.__dict__)
        setattr(pkg, name, mod)
        mods[name] = mod
    return mods["admm"], mods["quantization"], mods["utils"]


def import_reference():
    """Return a namespace with the reference's hot-path callables."""
    root = reference_root()
    if root is None:
        raise FileNotFoundError(f"{REFERENCE_ROOT} (and no oracle/_ref build product)")
    saved_path = list(sys.path)
    saved_mods = {k: v for k, v in sys.modules.items() if k == "source" or k.startswith("source.")}
    for k in saved_mods:
        del sys.modules[k]
    try:
        for name in ("tensorly", "tensorly.decomposition", "tensorly.decomposition.candecomp_parafac",
                     "tensorly.kruskal_tensor", "musco", "musco.pytorch", "musco.pytorch.compressor",
                     "musco.pytorch.compressor.decompose", "musco.pytorch.compressor.decompose.cpd",
                     "musco.pytorch.compressor.decompose.cpd.lib_anc"):
            if name not in sys.modules:
                _stub(name)
        sys.modules["tensorly"].set_backend = lambda *_a, **_k: None
        sys.modules["tensorly.decomposition"].parafac = _missing
        sys.modules["tensorly.decomposition.candecomp_parafac"].initialize_factors = _missing
        sys.modules["tensorly.kruskal_tensor"].KruskalTensor = _missing
        sys.modules["tensorly.kruskal_tensor"].kruskal_to_tensor = _missing
        sys.modules["musco.pytorch.compressor.decompose.cpd.lib_anc"].cp_anc = _missing
        if root == COMPILED_ROOT:
            admm, quant, utils = _load_compiled()
        else:
            sys.path.insert(0, root)
            admm = importlib.import_module("source.admm")
            quant = importlib.import_module("source.quantization")
            utils = importlib.import_module("source.utils")
        ns = types.SimpleNamespace(
            admm_iteration=admm.admm_iteration,
            init_factors=admm.init_factors,
            squared_relative_diff=admm.squared_relative_diff,
            quantize_tensor=quant.quantize_tensor,
            quantize_tensor_mse=quant.quantize_tensor_mse,
            min_max_quantize=quant.min_max_quantize,
            unfold=utils.unfold,
            root=root,
        )
        return ns
    finally:
        # leave no trace of the reference's `source` package: the product has its own
        for k in [k for k in sys.modules if k == "source" or k.startswith("source.")]:
            del sys.modules[k]
        sys.modules.update(saved_mods)
        sys.path[:] = saved_path
