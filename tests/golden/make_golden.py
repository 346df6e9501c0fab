#!/usr/bin/env python
"""Generates the committed golden fixtures with the CPU oracle (the reference itself cannot run here and ships
no golden images: SURVEY 4, 8c).  Run from the repo root:  python tests/golden/make_golden.py [name ...]

  converged_tunnel_160x90_4096spp_b8.npz   tunnel.gltf --sky, file camera, frames 0..511 x 8 spp, max_bounces 8
  converged_cornell_96x96_4096spp_b32.npz  cornell.gltf, file camera, frames 0..511 x 8 spp, max_bounces 32
  hitids_*.npz                             primary-hit (instance, primitive) maps + the oracle's edge/tie list
  ao_duck_160x90_f0-3.npz                  4-ray-ao on Duck.gltf (ToyCar.glb is missing upstream), frames 0..3
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from oracle import camera as ocam  # noqa: E402
from helpers import load_blue_noise, oracle_camera, oracle_scene  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def converged(name, w, h, sky, mb, frames=512):
    fs, osc = oracle_scene(name)
    cam = oracle_camera(fs, name, w, h)
    acc = np.zeros((h, w, 4), dtype=np.float32)
    st = oracle.OrcStats()
    t0 = time.time()
    rgba = None
    for f in range(frames):
        rgba, _ = osc.pathtrace_frame(ocam.scene_uniforms(cam, w, h, f), w, h, acc, 0, sky, 8, mb, st)
    path = os.path.join(OUT, "converged_%s_%dx%d_%dspp_b%d.npz" % (name, w, h, frames * 8, mb))
    np.savez_compressed(path, accum=acc[..., :3].copy(), rgba8=rgba, frames=frames, spp=8, max_bounces=mb, sky=int(sky),
                        rays=int(st.rays), paths=int(st.paths), capped=int(st.capped), emissive=int(st.emissive))
    print("%s: %.0f s, %d rays, %.2f rays/path" % (path, time.time() - t0, st.rays, st.rays / st.paths))


def hitids(name, w, h):
    fs, osc = oracle_scene(name)
    u = ocam.scene_uniforms(oracle_camera(fs, name, w, h), w, h, 0)
    rgba, ids, bary_t, flags = osc.debug(u, w, h)
    path = os.path.join(OUT, "hitids_%s_%dx%d.npz" % (name, w, h))
    inst = np.where(ids[..., 0] == oracle.MISS, 255, ids[..., 0]).astype(np.uint8)
    prim = np.where(ids[..., 1] == oracle.MISS, 65535, ids[..., 1]).astype(np.uint16)
    np.savez_compressed(path, inst=inst, prim=prim, flags=flags, rgba8=rgba)
    print("%s: listed %.3f %%" % (path, 100 * (flags != 0).mean()))


def ao(w, h, frames=4):
    fs, osc = oracle_scene("Duck")
    cam = oracle_camera(fs, "Duck_ao", w, h)
    blue = load_blue_noise()
    img = np.zeros((h, w, 4), dtype=np.float32)
    st = oracle.OrcStats()
    for f in range(frames):
        osc.ao_frame(ocam.scene_uniforms(cam, w, h, f), w, h, img, blue, 0, st)
    path = os.path.join(OUT, "ao_duck_%dx%d_f0-%d.npz" % (w, h, frames - 1))
    np.savez_compressed(path, image=img[..., :3].copy(), rays=int(st.rays), paths=int(st.paths))
    print(path, st.rays, st.paths)


if __name__ == "__main__":
    what = sys.argv[1:] or ["hitids", "ao", "cornell", "tunnel"]
    if "hitids" in what:
        hitids("cornell", 256, 256)
        hitids("Duck", 450, 300)
        hitids("tunnel", 480, 270)
    if "ao" in what:
        ao(160, 90)
    if "cornell" in what:
        converged("cornell", 96, 96, False, 32)
    if "tunnel" in what:
        converged("tunnel", 160, 90, True, 8)
