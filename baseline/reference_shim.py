"""Import the UNMODIFIED reference package from baseline/_ref (git-ignored; installed there by __graft_entry__.build() with
`pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>`; it travels to the GPU box
with the gpurun snapshot).  The reference imports matplotlib and `Quaternion` at module import
(policy_transportation/models/gaussian_process.py:11-13, transportation/policy_transportation.py:9); both are absent from this
image, so empty stand-ins are registered first.  Nothing of the reference's arithmetic is touched: the classes
returned here are the reference's own, running on the installed numpy / scipy / scikit-learn.

Used by bench.py (`--impl reference` and the `cpu_baseline` block) and by nothing in the product package."""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available():
    return os.path.isdir(os.path.join(REF_DIR, "policy_transportation"))


def import_reference():
    """Returns the reference's `policy_transportation` package (GaussianProcess, AffineTransform, GaussianProcessTransportation)."""
    if not available():
        raise ImportError(f"{REF_DIR}/policy_transportation is missing: run __graft_entry__.build() in the build container")
    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.cm", "mpl_toolkits", "mpl_toolkits.mplot3d", "Quaternion"]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["mpl_toolkits.mplot3d"].Axes3D = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import policy_transportation as pt
    return pt
