"""Generates tests/golden/*.npz from the UNMODIFIED reference compiled into oracle/_ref (oracle/build_ref.sh).

Run in the authoring container (needs /root/reference):   python tests/golden/make_golden.py
Each fixture holds, for one small scene of tests/conftest.py::SMALL_SCENES, the reference's own float colour buffer
(RayTracer::render's return value), primary-ray hit ids / t, the PPM bytes and the exact ray counts.  The scene files
themselves are regenerated deterministically by the scenes module at test time (closed-form geometry, no RNG); their
SHA-256 is stored so a generator change cannot silently invalidate a fixture.
"""
import hashlib
import importlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import binding as ob  # noqa: E402
from conftest import PKG, SMALL_SCENES  # noqa: E402

sc = importlib.import_module(PKG + ".scenes")


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    with tempfile.TemporaryDirectory() as d:
        for name, spec in SMALL_SCENES.items():
            scene = sc.CONFIGS[spec["builder"]](**spec["kw"])
            if "textures" in scene:
                for t in scene["textures"]:
                    if t["type"] == "bitmap":
                        sc.write_png_rgb(d + t["file_path"], sc.pattern_bitmap(64))
            path = sc.write_crtscene(os.path.join(d, name + ".crtscene"), scene)
            sha = hashlib.sha256(open(path, "rb").read()).hexdigest()
            ref = ob.run_reference(name + ".crtscene", d, os.path.join(d, name), textured=spec["tex"])
            ppm = ob.read_ppm_p3(ref["ppm_path"])
            rays = ref["rays"]
            np.savez_compressed(os.path.join(out_dir, name + ".npz"), rgb=ref["rgb"], hits=ref["hits"], ppm=ppm,
                                rays=np.array([rays["primary"], rays["shadow"], rays["reflection"], rays["refraction"]], np.int64),
                                scene_sha256=np.array(sha))
            print(name, ref["width"], ref["height"], rays, sha[:12])


if __name__ == "__main__":
    main()
