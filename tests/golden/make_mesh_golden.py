"""Parse the reference's cylinder meshes (examples/input/tetra_cyl.k: 2 513 nodes / 12 197 tetrahedra, node valence
3..38; cyl_hex.k: 6 479 nodes / 5 760 hexahedra in an unstructured O-grid numbering) with weldformfem_b200.deck.read_k
and commit node coordinates + connectivity as tests/golden/meshes/*.npz: /root/reference does not exist on the GPU box.

    python tests/golden/make_mesh_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from weldformfem_b200 import deck  # noqa: E402

SRC = "/root/reference/examples/input"


def main():
    for name in ("tetra_cyl", "cyl_hex"):
        x, el = deck.read_k(os.path.join(SRC, name + ".k"))
        out = os.path.join(HERE, "meshes", name + ".npz")
        np.savez_compressed(out, x=np.ascontiguousarray(x, dtype=np.float64), elnod=np.ascontiguousarray(el, dtype=np.uint32))
        print(name, x.shape, el.shape, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
