"""Recipe for oracle/_ref/: the UNMODIFIED reference hot-path modules, so that the CPU baseline of bench.py can time the
reference itself (`cpu_baseline.kind = "reference"`) on a GPU box, where /root/reference does not exist.

The reference is pure Python without a build system; "building" it means placing its two hot-path files,
source_code/filters_and_operators.py and source_code/stylization_layers.py, byte for byte under oracle/_ref/ (git-ignored,
not gpurun-ignored: it travels with the snapshot like the built .so files; it never enters the history).  They import
MONAI, which this image lacks; oracle/monai_shim provides the few MONAI 0.5 classes they use.
Run in the build container:  python oracle/build_ref.py        (__graft_entry__.build() does, when /root/reference exists)
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MVTB_REFERENCE", "/root/reference")
FILES = ["source_code/filters_and_operators.py", "source_code/stylization_layers.py"]
OUT = os.path.join(HERE, "_ref")


def build(verbose=True):
    if not os.path.isdir(REF):
        return False
    os.makedirs(OUT, exist_ok=True)
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(OUT, os.path.basename(rel))
        shutil.copyfile(src, dst)
        if verbose:
            print("oracle/_ref/%s  sha256 %s" % (os.path.basename(rel), hashlib.sha256(open(dst, "rb").read()).hexdigest()[:16]))
    return True


def import_reference():
    """(filters_and_operators, stylization_layers) of the unmodified reference from oracle/_ref, or None if absent."""
    if not os.path.exists(os.path.join(OUT, "filters_and_operators.py")):
        return None
    import importlib.util
    shim = os.path.join(HERE, "monai_shim")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    mods = []
    for name in ("filters_and_operators", "stylization_layers"):
        spec = importlib.util.spec_from_file_location("mvtb_reference_" + name, os.path.join(OUT, name + ".py"))
        m = importlib.util.module_from_spec(spec)
        if name == "stylization_layers":                       # it does `from filters_and_operators import ...`
            saved = sys.modules.get("filters_and_operators")
            sys.modules["filters_and_operators"] = mods[0]
            try:
                spec.loader.exec_module(m)
            finally:
                if saved is not None:
                    sys.modules["filters_and_operators"] = saved
                else:
                    sys.modules.pop("filters_and_operators", None)
        else:
            spec.loader.exec_module(m)
        mods.append(m)
    return tuple(mods)


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
