"""Recipe for oracle/_ref/: the UNMODIFIED reference solver module, for the `--impl reference` arm of bench.py.

The reference's hot path lives in one dependency-free file, wot/ot/optimal_transport.py (imports: logging, numpy).
`import wot` itself needs anndata / h5py / POT / matplotlib, which are absent from the image, so the file is taken
by path: this script copies it byte for byte from the read-only checkout into oracle/_ref/ (git-ignored, NOT
gpurun-ignored: it travels to the GPU box, where /root/reference does not exist) and records its SHA-256.
Nothing under oracle/_ref/ is ever imported by the product package; bench.py's reference arm loads it with
importlib and calls its public functions (compute_transport_matrix, optimal_transport_duality_gap) unchanged.
The cost matrix of the reference (ot_model.py:249-252: sklearn pairwise_distances 'sqeuclidean', n_jobs=-1, divided
by np.median) is a call into scikit-learn, which bench.py makes directly.

Run by __graft_entry__.build() whenever /root/reference is present; `python oracle/build_ref.py` by hand.
"""
from __future__ import annotations

import hashlib
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("WOT_REFERENCE", "/root/reference")
SOURCES = {"wot/ot/optimal_transport.py": "ref_optimal_transport.py"}
OUT = os.path.join(HERE, "_ref")


def build(ref_root=REF_ROOT, out=OUT):
    """Returns the output directory, or None when the reference checkout is not available (GPU box)."""
    if not os.path.isdir(ref_root):
        return out if os.path.exists(os.path.join(out, "ref_optimal_transport.py")) else None
    os.makedirs(out, exist_ok=True)
    lines = []
    for rel, name in SOURCES.items():
        src = os.path.join(ref_root, rel)
        dst = os.path.join(out, name)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as fh:
            lines.append("%s  %s  <- %s\n" % (hashlib.sha256(fh.read()).hexdigest(), name, rel))
    with open(os.path.join(out, "SOURCE.txt"), "w") as fh:
        fh.write("byte-for-byte copies made by oracle/build_ref.py from the reference checkout (sha256, file, origin)\n")
        fh.writelines(lines)
    return out


def load(out=OUT):
    """The reference's wot.ot.optimal_transport module object, or None if oracle/_ref has not been built."""
    import importlib.util
    path = os.path.join(out, "ref_optimal_transport.py")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("wot_reference_optimal_transport", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build())
