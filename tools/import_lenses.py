"""Re-emit the tabular lens prescriptions used by the benchmarks in this repository's own layout.

The prescriptions are published optical designs (radius, thickness, index, [V-number,] clear aperture per
surface, millimetres, front surface first).  This script reads the numeric fields of each table under
<reference>/lenses_tabular and writes them, numeric token for numeric token, to zoic_b200/data/lenses/ with
this repository's header, so that the GPU box (which has no reference tree) can run every configuration.
tests/test_host_setup.py::test_shipped_lens_tables_parse_like_the_reference_originals checks that both spellings parse to
identical element tables.

usage: python tools/import_lenses.py [/root/reference]
"""
import os
import re
import sys

NAMES = {
    "F_1.25_PETZVAL.dat": ("petzval_f1.25.dat", "Petzval, f/1.25"),
    "F_1.6_PETZVAL.dat": ("petzval_f1.6.dat", "Petzval, f/1.6"),
    "F_2.0_DOUBLE_GAUSS.dat": ("double_gauss_f2.0.dat", "Double Gauss, f/2.0, 22 deg half field"),
    "F_2.5_HFOV_TRIPLET.dat": ("triplet_f2.5.dat", "Cooke-type triplet, f/2.5"),
    "F_2.8_MORI_USP.dat": ("mori_f2.8.dat", "Mori (US patent) wide angle, f/2.8"),
    "F_2.8_TESSAR.dat": ("tessar_f2.8.dat", "Tessar, f/2.8"),
    "F_4.0_FISHEYE_MULLER.dat": ("fisheye_muller_f4.0.dat", "Muller fisheye, f/4.0"),
    "F_5.0_TELEPHOTO.dat": ("telephoto_f5.0.dat", "Telephoto, f/5.0"),
}


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    src_dir = os.path.join(ref, "lenses_tabular")
    dst_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "zoic_b200", "data", "lenses")
    os.makedirs(dst_dir, exist_ok=True)
    for src, (dst, title) in NAMES.items():
        rows = []
        for line in open(os.path.join(src_dir, src)):
            line = line.rstrip("\r\n")
            if not line or line.startswith("#"):
                continue
            rows.append([t for t in re.split(r"[\t,;: ]", line) if t])
        ncol = len(rows[0])
        assert all(len(r) == ncol for r in rows) and ncol in (4, 5), src
        cols = "radius\tthickness\tior\taperture" if ncol == 4 else "radius\tthickness\tior\tvnumber\taperture"
        with open(os.path.join(dst_dir, dst), "w") as f:
            f.write("# zoic_b200 tabular lens prescription: %s\n" % title)
            f.write("# units mm; one surface per line, front (object side) surface first;\n")
            f.write("# radius 0 marks the aperture stop; ior 0 means air\n")
            f.write("# %s\n" % cols)
            for r in rows:
                f.write("\t".join(r) + "\n")
        print(dst, len(rows), "surfaces,", ncol, "columns")


if __name__ == "__main__":
    main()
