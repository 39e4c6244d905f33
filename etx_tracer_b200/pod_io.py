"""Scene / Camera PODs <-> one .npz file.

The C ABI takes the reference's `Scene` / `Camera` PODs as they lie in host memory (render/shared/scene.hxx:22-65, camera.hxx:8-39): 528 /
176 bytes whose ArrayViews point into arrays owned by whoever built them (the reference's SceneRepresentation, scenes.py, a host
application).  `dump()` follows that pointer graph (Appendix B of SURVEY.md: images -> pixels, row / column distributions; media -> density
grids) and writes every array as raw bytes; `load()` rebuilds the graph over numpy arrays and returns an object with the `.scene` /
`.camera` / `.width` / `.height` / `.triangle_count` surface api.GPUVCM and the oracle read.  A scene prepared where the reference's loader
exists can so be rendered where it does not (tools/dump_reference_scene.py -> tests/golden/).
"""
import ctypes as C

import numpy as np

from . import structs as S

_TOP = (("vertices", S.VERTEX), ("triangles", S.TRIANGLE), ("triangle_to_emitter", np.dtype(np.uint32)), ("materials", S.MATERIAL),
        ("emitter_profiles", S.EMITTER_PROFILE), ("emitter_instances", S.EMITTER), ("images", S.IMAGE), ("mediums", S.MEDIUM),
        ("spectrums", S.SPECTRUM))


def _read(addr, count, dtype):
    dtype = np.dtype(dtype)
    n = int(count) * dtype.itemsize
    if (n == 0) or (int(addr) == 0):
        return np.zeros(0, dtype=dtype)
    return np.frombuffer((C.c_char * n).from_address(int(addr)), dtype=dtype).copy()


def dump(path, scene_obj):
    """scene_obj: anything with `.scene` (numpy S.SCENE[1], host pointers inside) and `.camera` (numpy S.CAMERA[1])."""
    sc = scene_obj.scene
    out = {"scene": np.frombuffer(sc.tobytes(), dtype=np.uint8), "camera": np.frombuffer(scene_obj.camera.tobytes(), dtype=np.uint8)}
    arrays = {}
    for name, dt in _TOP:
        arrays[name] = _read(sc[name]["a"][0], sc[name]["count"][0], dt)
        out[name] = arrays[name].view(np.uint8)
    # a Distribution over n items owns n + 1 entries (the closing {0, 0, 1} one, distribution_builder.hxx:8-14,52) whatever `values.count` says
    # (the reference's loader writes n, scenes.py n + 1): always carry n + 1 and keep the count field as it was
    out["emitters_distribution"] = _read(sc["emitters_distribution"]["values"]["a"][0], int(sc["emitter_instances"]["count"][0]) + 1, S.DIST_ENTRY).view(np.uint8)
    for i, im in enumerate(arrays["images"]):
        px_bytes = int(im["isize"][0]) * int(im["isize"][1]) * (16 if int(im["format"]) == 1 else 4)
        out[f"image{i}_pixels"] = _read(im["pixels"]["a"], px_bytes, np.uint8)
        has_table = int(im["y_distribution"]["values"]["a"]) != 0
        out[f"image{i}_ydist"] = _read(im["y_distribution"]["values"]["a"], (int(im["isize"][1]) + 1) if has_table else 0, S.DIST_ENTRY).view(np.uint8)
        rows = _read(im["x_distributions"]["a"], im["x_distributions"]["count"], S.DISTRIBUTION)
        out[f"image{i}_xdist_rows"] = rows.view(np.uint8)
        flat = [_read(r["values"]["a"], int(im["isize"][0]) + 1, S.DIST_ENTRY) for r in rows]
        out[f"image{i}_xdist_values"] = (np.concatenate(flat) if flat else np.zeros(0, dtype=S.DIST_ENTRY)).view(np.uint8)
    for i, md in enumerate(arrays["mediums"]):
        out[f"medium{i}_density"] = _read(md["density"]["a"], md["density"]["count"], np.float32).view(np.uint8)
    np.savez_compressed(path, **out)


class LoadedScene:
    """Owns the arrays a dumped Scene POD points to."""

    def __init__(self, path):
        z = np.load(path)
        self._keep = []
        self.scene = np.frombuffer(z["scene"].tobytes(), dtype=S.SCENE).copy()
        self.camera = np.frombuffer(z["camera"].tobytes(), dtype=S.CAMERA).copy()
        self.name = "pod:" + str(path)

        def own(raw, dt):
            a = np.frombuffer(raw.tobytes(), dtype=dt).copy()
            self._keep.append(a)
            return a

        def point(view, arr, count=None):
            view["a"] = arr.ctypes.data if arr.size else 0
            view["count"] = arr.shape[0] if count is None else count

        top = {name: own(z[name], dt) for name, dt in _TOP}
        images, mediums = top["images"], top["mediums"]
        for i in range(images.shape[0]):
            im = images[i:i + 1]
            px = own(z[f"image{i}_pixels"], np.uint8)
            im["pixels"]["a"] = px.ctypes.data if px.size else 0  # count keeps the loader's pixel count
            yd = own(z[f"image{i}_ydist"], S.DIST_ENTRY)
            point(im["y_distribution"]["values"], yd, int(im["y_distribution"]["values"]["count"][0]))
            rows = own(z[f"image{i}_xdist_rows"], S.DISTRIBUTION)
            vals = own(z[f"image{i}_xdist_values"], S.DIST_ENTRY)
            at = 0
            per_row = int(im["isize"][0][0]) + 1
            for r in range(rows.shape[0]):
                rows[r]["values"]["a"] = vals[at:at + per_row].ctypes.data
                at += per_row
            point(im["x_distributions"], rows)
        for i in range(mediums.shape[0]):
            d = own(z[f"medium{i}_density"], np.float32)
            point(mediums[i:i + 1]["density"], d)
        for name, _ in _TOP:
            point(self.scene[name], top[name])
        ed = own(z["emitters_distribution"], S.DIST_ENTRY)
        point(self.scene["emitters_distribution"]["values"], ed, int(self.scene["emitters_distribution"]["values"]["count"][0]))

    @property
    def width(self):
        return int(self.camera["film_size"][0][0])

    @property
    def height(self):
        return int(self.camera["film_size"][0][1])

    @property
    def triangle_count(self):
        return int(self.scene["triangles"]["count"][0])


def load(path):
    return LoadedScene(path)
