"""File formats on either side of the hot path, as the reference's drivers read/write them.

.flo   : "PIEH", int32 W, int32 H, H rows of W interleaved (u,v) float32 LE
         (ARAP/deformation/src/main.cpp:53-75, main.h:7-8; reader ARAP/warping/src/main.cpp:228-274)
cstr   : text, n then n x "x1 y1 x2 y2" ints (ARAP/deformation/src/main.cpp:26-50; para_gen.py:476-479)
PNG    : via PIL here (tests/bench only); the CLI binaries carry their own zlib-based codec.
"""
from __future__ import annotations

import numpy as np

FLO_TAG = b"PIEH"


def write_flo(path: str, flow: np.ndarray) -> None:
    flow = np.ascontiguousarray(flow, dtype="<f4")
    H, W, two = flow.shape
    assert two == 2
    with open(path, "wb") as f:
        f.write(FLO_TAG)
        f.write(np.asarray([W, H], dtype="<i4").tobytes())
        f.write(flow.tobytes())


def read_flo(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        d = f.read()
    if d[:4] != FLO_TAG:
        raise ValueError(f"{path}: wrong .flo tag {d[:4]!r}")
    W, H = np.frombuffer(d[4:12], dtype="<i4")
    if not (1 <= W <= 99999 and 1 <= H <= 99999):
        raise ValueError(f"{path}: illegal size {W}x{H}")
    body = np.frombuffer(d[12:], dtype="<f4")
    if body.size != 2 * W * H:
        raise ValueError(f"{path}: expected {2 * W * H} floats, found {body.size}")
    return body.reshape(H, W, 2).copy()


def write_constraints(path: str, matches: np.ndarray) -> None:
    with open(path, "w") as f:
        f.write(f"{len(matches)}\n")
        for x1, y1, x2, y2 in np.asarray(matches).reshape(-1, 4):
            f.write(f"{int(x1)}\t{int(y1)}\t{int(x2)}\t{int(y2)}\n")


def read_constraints(path: str) -> np.ndarray:
    with open(path) as f:
        tok = f.read().split()
    n = int(tok[0])
    vals = [int(t) for t in tok[1:1 + 4 * n]]
    if len(vals) != 4 * n:
        raise ValueError(f"{path}: expected {4 * n} integers")
    return np.asarray(vals, dtype=np.int32).reshape(n, 4)


def read_png_rgb(path: str) -> np.ndarray:
    from PIL import Image
    return np.asarray(Image.open(path).convert("RGB"), dtype=np.uint8).copy()


def read_png_mask_red(path: str) -> np.ndarray:
    """Red channel of the mask PNG -- the only channel the reference looks at (CombinedSolver.h:213)."""
    return read_png_rgb(path)[..., 0].copy()


def write_png(path: str, img: np.ndarray) -> None:
    from PIL import Image
    Image.fromarray(np.ascontiguousarray(img)).save(path)
