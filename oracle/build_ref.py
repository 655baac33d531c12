"""Byte-compile the UNMODIFIED reference's hot-path modules into oracle/_ref/ (git-ignored, travels to the GPU box).

TEST / MEASUREMENT INFRASTRUCTURE - not product code.  /root/reference does not exist on the GPU box, and its sources
are never copied into this repository.  What travels is a BUILD PRODUCT made from the sources where they lie, like a
compiled C reference would be: CPython bytecode (`py_compile`; stored as `<module>.pyc.bin` because snapshot tools
commonly drop `*.pyc`, loaded by oracle/ref_import.py with `marshal`) of

    source/admm.py  source/quantization.py  source/utils.py  source/parafac_epc.py

The GPU box runs the same image (same CPython magic number), so `oracle/ref_import.py` can import the package from
oracle/_ref when /root/reference is absent: `bench.py --impl reference` and the `cpu_baseline` leg then time the
reference's OWN functions (`cpu_baseline.kind = "reference"`), and the eager reference-on-B200 arm runs them on CUDA
tensors.  Usage:  python oracle/build_ref.py   (also called by __graft_entry__.build() when /root/reference exists)
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("ADMMQ_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
MODULES = ("admm", "quantization", "utils", "parafac_epc")


def build(verbose=False):
    src_dir = os.path.join(REFERENCE_ROOT, "source")
    if not os.path.isdir(src_dir):
        return None
    out_pkg = os.path.join(OUT, "source")
    os.makedirs(out_pkg, exist_ok=True)
    made = []
    for f in os.listdir(out_pkg):   # stale products of an earlier layout
        if f.endswith(".pyc"):
            os.remove(os.path.join(out_pkg, f))
    for m in MODULES:
        made.append(py_compile.compile(os.path.join(src_dir, m + ".py"), cfile=os.path.join(out_pkg, m + ".pyc.bin"),
                                       doraise=True, optimize=0))
    with open(os.path.join(OUT, "PROVENANCE.txt"), "w") as f:
        f.write(f"py_compile of {src_dir}/{{{','.join(MODULES)}}}.py by oracle/build_ref.py; python {sys.version.split()[0]}\n")
    if verbose:
        print("\n".join(made))
    return OUT


if __name__ == "__main__":
    print(build(verbose=True))
