"""Host-side mirror of how the reference's Python drivers hand work to the solver (para_gen.py:178-214, 441-445,
560-567; run_arap.py:10-15, 65-78): list files of 6-path lines, one solver process per GPU selected with
CUDA_VISIBLE_DEVICES, no communication between GPUs (the solves do not couple).

Nothing here computes: the work is done by the arap_deform binary (arap_flow_b200/bin) or, in-process, by
libarapb200 through arap_flow_b200.lib.
"""
from __future__ import annotations

import os
import subprocess
import sys
import tempfile
import time
from typing import List, Sequence, Tuple

Item = Tuple[str, str, str, str, str, str]  # rgb, mask, cstr, flo_out, wrgb_out, wmask_out

_HERE = os.path.dirname(os.path.abspath(__file__))
ARAP_BIN = os.path.join(_HERE, "bin", "arap_deform")
WARP_BIN = os.path.join(_HERE, "bin", "warp_image")
PLAN = os.path.join(_HERE, "arap_plan.t")


def read_list_file(path: str) -> List[Item]:
    """One whitespace-separated 6-tuple per line (ARAP/deformation/src/main.cpp:182-193); short lines are skipped."""
    items = []
    with open(path) as f:
        for line in f:
            tok = line.split()
            if len(tok) >= 6:
                items.append(tuple(tok[:6]))
    return items


def write_list_file(path: str, items: Sequence[Item]) -> None:
    with open(path, "w") as f:
        for it in items:
            f.write(" ".join(it) + "\n")


def shard(items: Sequence[Item], rank: int, world: int) -> List[Item]:
    """Static round-robin of independent (pair, segment) units over `world` GPUs.  The reference
    uses a dynamic free-GPU queue (para_gen.py:441-445); with equal-cost units the static split is equivalent."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return [it for i, it in enumerate(items) if i % world == rank]


def do_arap(items: Sequence[Item], gpu: int, tmp_dir: str, arap_bin: str = ARAP_BIN, plan: str = PLAN, batch: int = 9):
    """para_gen.do_arap (para_gen.py:178-200): write a temporary list file, run the solver binary on one GPU, assert
    a zero exit code, always remove the list file.  Returns the elapsed seconds."""
    os.makedirs(tmp_dir, exist_ok=True)
    fd, list_path = tempfile.mkstemp(prefix=f"gpu-{gpu}_", suffix=".txt", dir=tmp_dir)
    os.close(fd)
    t0 = time.time()
    try:
        write_list_file(list_path, items)
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(gpu), ARAP_PLAN=plan, ARAP_BATCH=str(batch))
        rc = subprocess.call([arap_bin, list_path], env=env, stdout=subprocess.DEVNULL)
        assert rc == 0, f"{arap_bin} failed with code {rc} on GPU {gpu}"
    finally:
        os.remove(list_path)
    return time.time() - t0


def run_sharded(items: Sequence[Item], gpus: Sequence[int], tmp_dir: str, **kw) -> float:
    """One solver process per GPU, all at once, each on its shard; returns the wall-clock seconds of the slowest."""
    from multiprocessing.pool import ThreadPool
    t0 = time.time()
    with ThreadPool(len(gpus)) as pool:
        pool.starmap(lambda r, g: do_arap(shard(items, r, len(gpus)), g, tmp_dir, **kw) if shard(items, r, len(gpus)) else 0.0,
                     list(enumerate(gpus)))
    return time.time() - t0


def main(argv=None):
    """python -m arap_flow_b200.driver LISTFILE --gpu 0 1 2 3   (the --gpu flag of para_gen.py:611-640)"""
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("listfile")
    ap.add_argument("--gpu", type=int, nargs="+", default=[0])
    ap.add_argument("--tmp", default=os.path.join(tempfile.gettempdir(), "arapb200"))
    ap.add_argument("--batch", type=int, default=8)
    a = ap.parse_args(argv)
    items = read_list_file(a.listfile)
    dt = run_sharded(items, a.gpu, a.tmp, batch=a.batch)
    print(f"{len(items)} items on {len(a.gpu)} GPU(s) in {dt:.2f} s")


if __name__ == "__main__":
    main()
