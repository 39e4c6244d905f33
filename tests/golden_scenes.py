"""Scenes behind the committed oracle renders tests/golden/oracle_*_24.npz (generator: tools/make_golden.py).
name -> (scene factory, iterations, options factory or None)."""
from etx_tracer_b200 import scenes, structs as S


def _connect_only():
    o = S.default_vcm_options()
    o["options"] = S.VCM_CONNECT_ONLY
    return o


SCENES = {
    "oracle_c3_24.npz": (lambda: scenes.procedural_room(24, 14, env_size=(256, 128)), 2, None),
    "oracle_c4_24.npz": (lambda: scenes.sss_dragon(24, 24), 2, None),
    "oracle_c5_24.npz": (lambda: scenes.cloud_box(24, 24, grid=64), 3, _connect_only),
    "oracle_vmf_24.npz": (lambda: scenes.material_box("diffuse_vmf_mid", 24, 24, spectral=True), 3, None),
}
