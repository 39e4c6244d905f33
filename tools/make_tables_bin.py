#!/usr/bin/env python
"""etx_tracer_b200/data/{color_tables,spectra}.npz -> etx_tracer_b200/data/tables.bin, the table file the C++ scene loader (csrc/scene_loader.cpp) reads.

Layout (little endian): u32 magic 'ETXT' (0x54585445), u32 entry_count; per entry u16 name_len, name, u32 dtype (0 = f32, 1 = u32, 2 = u8), u32 count, data.
Names are "<npz stem>/<array name>".  The black-body fixtures of spectra.npz are test data and stay out.  Needs only the npz files (no reference tree).
"""
import os
import struct
import sys

import numpy as np

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "etx_tracer_b200", "data")


def main():
    entries = []
    for stem in ("color_tables", "spectra", "bluenoise"):
        z = np.load(os.path.join(DATA, stem + ".npz"))
        for key in z.files:
            if key.startswith("blackbody_") or key.startswith("nblackbody_"):
                continue
            a = z[key]
            if a.dtype == np.float32:
                entries.append((f"{stem}/{key}", 0, np.ascontiguousarray(a).reshape(-1)))
            elif a.dtype == np.uint32:
                entries.append((f"{stem}/{key}", 1, np.ascontiguousarray(a).reshape(-1)))
            elif a.dtype == np.uint8:
                entries.append((f"{stem}/{key}", 2, np.ascontiguousarray(a).reshape(-1)))
            else:
                raise SystemExit(f"{stem}/{key}: dtype {a.dtype} is not stored")
    out = os.path.join(DATA, "tables.bin")
    with open(out + ".tmp", "wb") as f:
        f.write(struct.pack("<II", 0x54585445, len(entries)))
        for name, dt, a in entries:
            nb = name.encode()
            f.write(struct.pack("<H", len(nb)) + nb + struct.pack("<II", dt, a.size) + a.tobytes())
    os.replace(out + ".tmp", out)
    print(out, os.path.getsize(out), "bytes,", len(entries), "tables")


if __name__ == "__main__":
    sys.exit(main())
