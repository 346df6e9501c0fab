#!/usr/bin/env python
"""Screenshot-layer capture (binary PPM, P6) -> tests/golden/ref_*.npz.
usage: ppm_to_fixture.py ids   capture.ppm out.npz
       ppm_to_fixture.py frame capture.ppm out.npz <frames accumulated> <sky 0|1>"""
import sys

import numpy as np


def read_ppm(path):
    data = open(path, "rb").read()
    tokens, pos = [], 0
    while len(tokens) < 4:  # magic, width, height, maxval (comments allowed between them)
        while data[pos:pos + 1].isspace():
            pos += 1
        if data[pos:pos + 1] == b"#":
            pos = data.index(b"\n", pos) + 1
            continue
        end = pos
        while not data[end:end + 1].isspace():
            end += 1
        tokens.append(data[pos:end])
        pos = end
    if tokens[0] != b"P6" or int(tokens[3]) != 255:
        raise SystemExit("expected an 8-bit binary PPM")
    w, h = int(tokens[1]), int(tokens[2])
    return np.frombuffer(data, np.uint8, w * h * 3, pos + 1).reshape(h, w, 3).copy()


def main():
    kind, src, dst = sys.argv[1:4]
    rgb = read_ppm(src)
    if kind == "ids":
        r, g, b = (rgb[..., k].astype(np.uint32) for k in range(3))
        prim = r | (g << 8) | ((b & 15) << 16)
        inst = b >> 4
        miss = np.all(rgb == 255, axis=2)
        ids = np.stack([np.where(miss, 0xFFFFFFFF, inst), np.where(miss, 0xFFFFFFFF, prim)], axis=2).astype(np.uint32)
        # an sRGB swapchain would have re-encoded the bytes: ids of neighbouring pixels would then not form runs of valid values
        if (~miss).any() and prim[~miss].max() >= (1 << 20):
            raise SystemExit("capture does not decode to ids: was the swapchain an sRGB format?")
        np.savez_compressed(dst, ids=ids)
    else:
        np.savez_compressed(dst, rgba8=np.concatenate([rgb, np.full(rgb.shape[:2] + (1,), 255, np.uint8)], axis=2),
                            frames=int(sys.argv[4]), sky=int(sys.argv[5]))
    print("wrote", dst)


if __name__ == "__main__":
    main()
