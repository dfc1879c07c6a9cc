"""Copy the reference's OWN example decks (inputs, not source code) next to the tests, because /root/reference does not
exist on the GPU box:  examples/input/{Compression_tetra, Contact_Compression_tetra, Contact_Compression_axisymm_quad}.json
verbatim and the mesh they name, tetra_cyl.k, gzip-compressed (1.2 MB of text -> 0.2 MB).

    python tests/golden/fetch_ref_decks.py
"""
import gzip
import os
import shutil

SRC = "/root/reference/examples/input"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_decks")
DECKS = ["Compression_tetra.json", "Contact_Compression_tetra.json", "Contact_Compression_axisymm_quad.json"]


def main():
    os.makedirs(DST, exist_ok=True)
    for f in DECKS:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
    with open(os.path.join(SRC, "tetra_cyl.k"), "rb") as fi, gzip.GzipFile(os.path.join(DST, "tetra_cyl.k.gz"), "wb", mtime=0) as fo:
        shutil.copyfileobj(fi, fo)
    for f in sorted(os.listdir(DST)):
        print(f, os.path.getsize(os.path.join(DST, f)))


if __name__ == "__main__":
    main()
