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


class Server:
    """One resident `arap_deform --serve SPOOL` worker on one GPU: the CUDA context, the plan and the
    device buffers outlive the dispatches, so a dispatch of a handful of pairs no longer pays 0.6-4 s of start-up.
    para_gen.py keeps calling the binary with a list file (para_gen.py:190-195); with ARAP_SERVER set that call is a thin
    client of this worker."""

    def __init__(self, gpu: int, spool: str, arap_bin: str = ARAP_BIN, plan: str = PLAN, batch: int = 8, timeout: float = 120.0,
                 warm=None):
        """warm = (W, H): build the plan and the buffers for that image size at start-up (para_gen's --size), so that the
        first dispatch already runs at steady state."""
        self.gpu, self.spool = gpu, spool
        os.makedirs(spool, exist_ok=True)
        for f in os.listdir(spool):                       # leftovers of a previous worker
            os.remove(os.path.join(spool, f))
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(gpu), ARAP_PLAN=plan, ARAP_BATCH=str(batch))
        env.pop("ARAP_SERVER", None)
        cmd = [arap_bin, "--serve", spool] + (["--warm", "%dx%d" % tuple(warm)] if warm else [])
        self.proc = subprocess.Popen(cmd, env=env, stdout=subprocess.DEVNULL)
        t0 = time.time()
        while not os.path.exists(os.path.join(spool, "ready")):
            if self.proc.poll() is not None:
                raise RuntimeError(f"arap_deform --serve exited with code {self.proc.returncode}")
            if time.time() - t0 > timeout:
                self.proc.kill()
                raise RuntimeError("arap_deform --serve did not come up")
            time.sleep(0.005)

    def close(self):
        if self.proc and self.proc.poll() is None:
            open(os.path.join(self.spool, "stop"), "w").close()
            try:
                self.proc.wait(timeout=30)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        self.proc = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def do_arap(items: Sequence[Item], gpu: int, tmp_dir: str, arap_bin: str = ARAP_BIN, plan: str = PLAN, batch: int = 8,
            server: str = None):
    """para_gen.do_arap (para_gen.py:178-200): write a temporary list file, run the solver binary on one GPU, assert
    a zero exit code, always remove the list file.  Returns the elapsed seconds.  `server` = spool directory of a
    running Server for that GPU: the binary then only hands the list over (ARAP_SERVER)."""
    os.makedirs(tmp_dir, exist_ok=True)
    fd, list_path = tempfile.mkstemp(prefix=f"gpu-{gpu}_", suffix=".txt", dir=tmp_dir)
    os.close(fd)
    t0 = time.time()
    try:
        write_list_file(list_path, items)
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(gpu), ARAP_PLAN=plan, ARAP_BATCH=str(batch))
        if server:
            env["ARAP_SERVER"] = server
        rc = subprocess.call([arap_bin, list_path], env=env, stdout=subprocess.DEVNULL)
        assert rc == 0, f"{arap_bin} failed with code {rc} on GPU {gpu}"
    finally:
        os.remove(list_path)
    return time.time() - t0


def run_sharded(items: Sequence[Item], gpus: Sequence[int], tmp_dir: str, servers: Sequence[str] = None, **kw) -> float:
    """One solver process per GPU, all at once, each on its shard; returns the wall-clock seconds of the slowest.
    `servers` = one spool directory per entry of `gpus` (resident workers already running)."""
    from multiprocessing.pool import ThreadPool
    t0 = time.time()

    def one(r, g):
        mine = shard(items, r, len(gpus))
        return do_arap(mine, g, tmp_dir, server=servers[r] if servers else None, **kw) if mine else 0.0

    with ThreadPool(len(gpus)) as pool:
        pool.starmap(one, list(enumerate(gpus)))
    return time.time() - t0


def run_dispatches(dispatches: Sequence[Sequence[Item]], gpus: Sequence[int], tmp_dir: str, servers: Sequence[str] = None, **kw) -> float:
    """para_gen's dynamic dispatch (para_gen.py:441-445, 560-567): a queue of free GPUs; every dispatch (a short list of
    6-tuples) goes to the next free GPU as its own do_arap call.  Returns the wall-clock seconds until the last finishes."""
    import queue
    import threading
    free = queue.Queue()
    for r, g in enumerate(gpus):
        free.put((r, g))
    t0 = time.time()
    threads = []

    def work(r, g, d):
        try:
            do_arap(d, g, tmp_dir, server=servers[r] if servers else None, **kw)
        finally:
            free.put((r, g))

    for d in dispatches:
        r, g = free.get()
        th = threading.Thread(target=work, args=(r, g, d))
        th.start()
        threads.append(th)
    for th in threads:
        th.join()
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
