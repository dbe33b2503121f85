"""Batch sharding over independent replicas (one process per GPU).

Images are independent (inference-mode batch norm, per-image decode), so the
N-GPU path has no data-path collective: every rank runs the whole path on a
contiguous slice of the batch and rank 0 gathers the small per-image result
tables in image order.  ``torch.distributed`` is used only for
that gather (and, in bench.py, for the barrier and the max-reduction of the
elapsed time); the backend is NCCL on GPUs and gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple


def shard_bounds(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice ``[start, stop)`` of ``n_items`` for ``rank``: the first ``n_items % world`` ranks get one
    extra item, so the slices differ by at most one image and concatenate back in order."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} of {world}")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_in_order(local: Sequence, dst: int = 0, group=None) -> Optional[List]:
    """Gathers every rank's list of per-image results on ``dst`` and concatenates them in rank (= image) order.
    Returns the full list on ``dst``, ``None`` elsewhere.  Works without an initialised process group
    (single replica)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return list(local)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bucket = [None] * world if rank == dst else None
    dist.gather_object(list(local), bucket, dst=dst, group=group)
    if rank != dst:
        return None
    out: List = []
    for part in bucket:
        out.extend(part)
    return out


def run_sharded(items: Sequence, infer: Callable[[Sequence], Sequence], dst: int = 0, group=None) -> Optional[List]:
    """Runs ``infer`` (a callable mapping a batch of items to one result per item, e.g. ``Engine.infer_host``)
    on this rank's slice of ``items`` and gathers all results on ``dst`` in the original order."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    lo, hi = shard_bounds(len(items), world, rank)
    local = list(infer(items[lo:hi])) if hi > lo else []
    if len(local) != hi - lo:
        raise RuntimeError(f"infer returned {len(local)} results for {hi - lo} items")
    return gather_in_order(local, dst, group)


# ---------------------------------------------------------------------------
# host-side result gather through shared memory (one node, one process per GPU)
# ---------------------------------------------------------------------------
_ROWS_PER_IMAGE = (128, 128, 16)      # joint, connection and person rows per image the default tables are sized for


class SharedResultRing:
    """Result tables of ``world`` ranks in ONE POSIX shared-memory segment, ``slots`` steps deep.

    The batch-sharded job has no data-path collective: every rank decodes its slice of the batch
    and rank 0 needs the small per-image tables of all of them, in image order.  Instead of serialising them through
    a socket, every rank's device->host download lands DIRECTLY in its region of this segment (the region is
    page-locked with ``cudaHostRegister`` and handed to ``Engine.infer_host_begin(tables=...)``); "gather" is then a
    per-(slot, rank) sequence flag that rank 0 polls, and the gathered result is a list of zero-copy views in rank
    (= image) order.  Ranks run ahead of rank 0 by at most ``slots`` steps.

    Layout: header ``int64[2 + slots * world]`` (magic, steps consumed by rank 0, published step + 1 per slot and
    rank), then per slot and rank the five tables (counts int32 [b,32]; joints f64 [b*128,6]; limbs f64 [b*128,5];
    persons f64 [b*16,16]; keypoints f32 [b*16,42]) with ``b`` = that rank's slice of the global batch."""

    MAGIC = 0x4C574F5052494E47      # "LWOPRING"

    def __init__(self, name: str, global_batch: int, world: int, rank: int, slots: int = 4, counts_width: int = 32,
                 rows_scale: int = 1, pin: bool = True, timeout_s: float = 120.0):
        import time
        from multiprocessing import resource_tracker, shared_memory
        import numpy as np
        self.name, self.world, self.rank, self.slots = name, world, rank, slots
        self.global_batch, self.counts_width = global_batch, counts_width
        self.bounds = [shard_bounds(global_batch, world, r) for r in range(world)]
        self._rows = tuple(r * rows_scale for r in _ROWS_PER_IMAGE)
        self._hdr_words = 2 + slots * world
        off = ((self._hdr_words * 8 + 4095) // 4096) * 4096
        self._region = []                       # [slot][rank] -> (offset, batch)
        for s in range(slots):
            row = []
            for r in range(world):
                b = self.bounds[r][1] - self.bounds[r][0]
                row.append((off, b))
                off += self._region_bytes(b)
            self._region.append(row)
        self.nbytes = off
        self._registered = []
        if rank == 0:
            try:
                stale = shared_memory.SharedMemory(name=name)
                stale.close()
                stale.unlink()
            except FileNotFoundError:
                pass
            self.shm = shared_memory.SharedMemory(name=name, create=True, size=self.nbytes)
            self.hdr = np.ndarray((self._hdr_words,), dtype=np.int64, buffer=self.shm.buf)
            self.hdr[:] = 0
            self.hdr[0] = self.MAGIC
        else:
            t0 = time.time()
            while True:
                try:
                    self.shm = shared_memory.SharedMemory(name=name)
                    if self.shm.size >= self.nbytes:
                        hdr = np.ndarray((self._hdr_words,), dtype=np.int64, buffer=self.shm.buf)
                        if hdr[0] == self.MAGIC:
                            self.hdr = hdr
                            break
                    self.shm.close()
                except FileNotFoundError:
                    pass
                if time.time() - t0 > timeout_s:
                    raise TimeoutError(f"shared result ring {name} did not appear")
                time.sleep(0.01)
            try:        # the creator unlinks; keep this process's resource tracker from doing it (and warning) at exit
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        self._tables = [[self._make_tables(s, r) for r in range(world)] for s in range(slots)]
        if pin:
            self._pin_own_regions()

    # -- layout ------------------------------------------------------------------------------------------------
    def _region_bytes(self, b: int) -> int:
        nj, nl, npers = (b * r for r in self._rows)
        n = b * self.counts_width * 4 + nj * 6 * 8 + nl * 5 * 8 + npers * 16 * 8 + npers * 42 * 4
        return ((n + 4095) // 4096) * 4096

    def _make_tables(self, slot: int, rank: int):
        import numpy as np
        off, b = self._region[slot][rank]
        nj, nl, npers = (b * r for r in self._rows)
        buf = self.shm.buf
        out = {"batch": b, "external": True, "_offset": off, "_nbytes": self._region_bytes(b)}
        for key, shape, dt in (("joints", (nj, 6), np.float64), ("limbs", (nl, 5), np.float64),
                               ("persons", (npers, 16), np.float64), ("counts", (b, self.counts_width), np.int32),
                               ("keypoints", (npers, 42), np.float32)):
            out[key + "_np"] = np.ndarray(shape, dtype=dt, buffer=buf, offset=off)
            off += int(np.prod(shape)) * np.dtype(dt).itemsize
        return out

    def _pin_own_regions(self) -> None:
        """Page-lock this rank's regions so the device->host copies into them are asynchronous DMA."""
        try:
            import torch
            if not torch.cuda.is_available():
                return
            import ctypes
            base = ctypes.addressof(ctypes.c_char.from_buffer(self.shm.buf))
            rt = torch.cuda.cudart()
            for s in range(self.slots):
                t = self._tables[s][self.rank]
                ptr, n = base + t["_offset"], t["_nbytes"]
                err = rt.cudaHostRegister(ptr, n, 0)
                if int(err) != 0:
                    raise RuntimeError(f"cudaHostRegister failed: {err}")
                self._registered.append(ptr)
        except ImportError:
            pass

    # -- producer side (every rank) -----------------------------------------------------------------------------
    def tables(self, step: int):
        """Host result buffers of THIS rank for ``step`` (torch views for ``Engine.infer_host_begin(tables=...)``)."""
        import torch
        t = self._tables[step % self.slots][self.rank]
        if "counts" not in t:
            for key in ("counts", "joints", "limbs", "persons", "keypoints"):
                t[key] = torch.from_numpy(t[key + "_np"])
        return t

    def wait_slot_free(self, step: int, timeout_s: float = 120.0) -> None:
        """Block until rank 0 has consumed the step that last used this step's slot."""
        import time
        t0 = time.time()
        while int(self.hdr[1]) < step - self.slots + 1:
            if time.time() - t0 > timeout_s:
                raise TimeoutError("rank 0 stopped consuming results")
            time.sleep(0)

    def publish(self, step: int) -> None:
        """This rank's results of ``step`` are complete in host memory."""
        self.hdr[2 + (step % self.slots) * self.world + self.rank] = step + 1

    # -- consumer side (rank 0) ---------------------------------------------------------------------------------
    def ready(self, step: int) -> bool:
        base = 2 + (step % self.slots) * self.world
        return bool((self.hdr[base:base + self.world] == step + 1).all())

    def gather(self, step: int, timeout_s: float = 120.0):
        """Per-rank result views of ``step`` in rank (= image) order: a list of dicts with the numpy tables
        (``counts``, ``joints``, ``limbs``, ``persons``, ``keypoints``) and ``batch``; valid until :meth:`release`."""
        import time
        t0 = time.time()
        while not self.ready(step):
            if time.time() - t0 > timeout_s:
                raise TimeoutError(f"results of step {step} did not arrive from every rank")
            time.sleep(0)
        out = []
        for r in range(self.world):
            t = self._tables[step % self.slots][r]
            out.append({"batch": t["batch"], "counts": t["counts_np"], "joints": t["joints_np"], "limbs": t["limbs_np"],
                        "persons": t["persons_np"], "keypoints": t["keypoints_np"]})
        return out

    def release(self, step: int) -> None:
        self.hdr[1] = step + 1

    def close(self) -> None:
        try:
            import torch
            if self._registered:
                rt = torch.cuda.cudart()
                for ptr in self._registered:
                    rt.cudaHostUnregister(ptr)
                self._registered = []
        except Exception:
            pass
        self.hdr = None
        for row in self._tables:
            for t in row:
                t.clear()
        self._tables = []
        try:
            self.shm.close()
        except BufferError:
            pass
        if self.rank == 0:
            try:
                self.shm.unlink()
            except FileNotFoundError:
                pass


def pack_host_results(results, tables) -> None:
    """Writes per-image decode results (dicts with ``joint_list`` (N,6), ``limbs`` 13 x (m,5), ``persons`` (P,16),
    ``joints`` (P,42)) into host result tables in the packed layout ``lwop_infer_host`` produces (``include/lwop.h``:
    ``lwop_packed_tables``).  The CUDA path fills these tables by DMA; this is the same layout written from the host,
    for producers that are not the CUDA library (the CPU tests of the multi-rank gather)."""
    import numpy as np
    counts, joints, limbs = tables["counts_np"], tables["joints_np"], tables["limbs_np"]
    persons, keypoints = tables["persons_np"], tables["keypoints_np"]
    counts[:] = 0
    oj = ol = op = 0
    for b, r in enumerate(results):
        jl = np.asarray(r["joint_list"], dtype=np.float64).reshape(-1, 6)
        pa = np.asarray(r["persons"], dtype=np.float64).reshape(-1, 16)
        kp = np.asarray(r["joints"], dtype=np.float32).reshape(-1, 42)
        counts[b, 0], counts[b, 1] = jl.shape[0], pa.shape[0]
        for c in range(14):
            counts[b, 4 + c] = int((jl[:, 5] == c).sum()) if jl.size else 0
        joints[oj:oj + jl.shape[0]] = jl
        oj += jl.shape[0]
        persons[op:op + pa.shape[0]] = pa
        keypoints[op:op + pa.shape[0]] = kp
        op += pa.shape[0]
        for t in range(13):
            lt = np.asarray(r["limbs"][t], dtype=np.float64).reshape(-1, 5)
            counts[b, 18 + t] = lt.shape[0]
            counts[b, 3] += lt.shape[0]
            limbs[ol:ol + lt.shape[0]] = lt
            ol += lt.shape[0]


def batch_result_from_tables(view):
    """Zero-copy :class:`~lightweight_openpose_b200.engine.BatchResult` over one rank's gathered tables."""
    from .engine import BatchResult
    return BatchResult(view["batch"], view["counts"], view["joints"], view["limbs"], view["persons"], view["keypoints"],
                       copy=False)
