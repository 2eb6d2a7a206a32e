#!/usr/bin/env python
"""Build miniworld_b200/assets/pack_v1.npz from a Miniworld resource directory.

    python tools/pack_assets.py --src /root/reference/miniworld

The pack holds decoded RGB8 texels for the texture families the in-scope levels use and
decoded per-face-vertex arrays for the ball / key meshes (Apache-2.0 assets of
Farama-Foundation/Miniworld; data only, no code).  Identical mesh geometry (ball_red ==
ball_blue ...) is stored once and referenced by name.
"""
import argparse
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

TEXTURES = ["concrete", "floor_tiles_bw", "concrete_tiles", "brick_wall", "asphalt", "cinder_blocks", "slime",
            "logo_mila"] + ["chars/ch_0x%d" % ord(c) for c in
                            "0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz"]   # every TextFrame glyph
MESHES = ["ball_%s" % c for c in ("blue", "green", "grey", "purple", "red", "yellow")] + \
         ["key_%s" % c for c in ("blue", "green", "grey", "purple", "red", "yellow")] + \
         ["building", "cone", "medkit", "duckie"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", required=True)
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                                                  "miniworld_b200", "assets", "pack_v1.npz"))
    args = ap.parse_args()
    from PIL import Image
    from miniworld_b200.assets import parse_obj
    out = {}
    for name in TEXTURES:
        for i in range(1, 10):
            p = os.path.join(args.src, "textures", "%s_%d.png" % (name, i))
            if not os.path.exists(p):
                break
            with Image.open(p) as im:
                out["tex/%s_%d" % (name, i)] = np.asarray(im.convert("RGB"))
    geoms = {}
    for name in MESHES:
        d = parse_obj(os.path.join(args.src, "meshes", name + ".obj"))
        h = hashlib.sha1(d["verts"].tobytes() + d["norms"].tobytes() + d["texcs"].tobytes()).hexdigest()[:12]
        if h not in geoms:
            geoms[h] = name.split("_")[0]
            for k in ("verts", "norms", "texcs", "min_coords", "max_coords"):
                out["meshgeom/%s/%s" % (geoms[h], k)] = d[k]
            tri_tex = np.full(len(d["verts"]), -1, np.int32)
            ntex = 0
            for start, end, tex_path in d["chunks"]:
                if tex_path is not None:
                    with Image.open(tex_path) as im:
                        out["meshtex/%s/%d" % (geoms[h], ntex)] = np.asarray(im.convert("RGB"))
                    tri_tex[start:end] = -2 - ntex
                    ntex += 1
            if ntex:
                out["meshgeom/%s/tri_tex" % geoms[h]] = tri_tex
        out["mesh/%s/geom" % name] = np.array(geoms[h])
        out["mesh/%s/colors" % name] = d["colors"]
    np.savez_compressed(args.out, **out)
    print("wrote", args.out, os.path.getsize(args.out), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
