"""Host-side engine: owns one ``lwop_handle`` per (GPU, frame size), binds the
checkpoint and exposes forward / decode / whole-path calls on torch tensors or
host arrays.  PyTorch is used for device memory and streams only; all compute is
in ``_lwop.so``.

Reference protocol this replaces (reference ``test&eval/model_test.py:39-49,63-71``):
build graph -> ``Saver.restore`` -> per frame ``sess.run([cpm, paf])`` -> ``decode_pose``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib, checkpoint, graph

_MODES = {"bf16": _lib.MODE_BF16, "fp32": _lib.MODE_FP32_CHECK, "fp32_check": _lib.MODE_FP32_CHECK,
          "bf16_direct": _lib.MODE_BF16_DIRECT}


def _torch():
    import torch
    return torch


def require_gpu(device: int = 0) -> None:
    torch = _torch()
    if not torch.cuda.is_available():
        raise _lib.LwopError(_lib.ERR_NO_DEVICE, "no CUDA device: the lwop path has no CPU fallback")
    major, _ = torch.cuda.get_device_capability(device)
    if major != 10:
        raise _lib.LwopError(_lib.ERR_NO_DEVICE, "lwop kernels are built for sm_100a (B200) only")


def validate_host_images(images_host, height: int, width: int, max_batch: int):
    """Host frames for the whole-path calls -> contiguous CPU tensor ``[B, H, W, 3]`` uint8 or float32.
    The C side reads ``B*H*W*3`` elements of the stated type from a raw pointer, so everything is checked here:
    shape against the engine, batch against its capacity, contiguity, dtype.  float64 / float16 frames (the
    reference's own ``img / 255.`` is float64, ``test&eval/model_test.py:65``) are cast to float32, which is what
    the reference's float32 placeholder does to its feed; any other dtype raises."""
    torch = _torch()
    if isinstance(images_host, np.ndarray):
        images_host = torch.from_numpy(np.ascontiguousarray(images_host))
    if not isinstance(images_host, torch.Tensor):
        raise ValueError("images_host must be a numpy array or a CPU torch tensor")
    if images_host.device.type != "cpu":
        raise ValueError("images_host must live in host memory (use forward() for CUDA tensors)")
    if images_host.dim() != 4 or tuple(images_host.shape[1:]) != (height, width, 3):
        raise ValueError(f"engine built for frames [B,{height},{width},3], got {tuple(images_host.shape)}")
    if not 1 <= images_host.shape[0] <= max_batch:
        raise ValueError(f"batch {images_host.shape[0]} outside 1..{max_batch}")
    if images_host.dtype in (torch.float64, torch.float16, torch.bfloat16):
        images_host = images_host.to(torch.float32)
    elif images_host.dtype not in (torch.uint8, torch.float32):
        raise ValueError(f"frames must be uint8 or floating point, got {images_host.dtype}")
    if not images_host.is_contiguous():
        images_host = images_host.contiguous()
    return images_host


class DecodeResult:
    """Per-image decode output in the reference's types (pose_decode.py:516-517)."""
    __slots__ = ("joint_list", "person_to_joint_assoc", "joints", "limbs", "counts")

    def __init__(self, joint_list, persons, joints, limbs, counts):
        self.joint_list = joint_list
        self.person_to_joint_assoc = persons
        self.joints = joints
        self.limbs = limbs
        self.counts = counts


class BatchResult:
    """Decode results of a batch, backed by the packed host arrays the C call filled.
    ``len()``, indexing and iteration give per-image :class:`DecodeResult` objects,
    built lazily (slicing + copying the packed rows) so that the hot path does not
    pay for Python objects nobody looks at."""

    def __init__(self, batch: int, counts: np.ndarray, joints: np.ndarray, limbs: np.ndarray,
                 persons: np.ndarray, keypoints: np.ndarray, copy: bool = True):
        self.batch = batch
        c = np.array(counts[:batch], copy=True)
        self.counts = c
        nj, npers, nconn = int(c[:, 0].sum()), int(c[:, 1].sum()), int(c[:, 3].sum())
        take = (lambda a, n: np.array(a[:n], copy=True)) if copy else (lambda a, n: a[:n])
        self.joints, self.limbs = take(joints, nj), take(limbs, nconn)
        self.persons, self.keypoints = take(persons, npers), take(keypoints, npers)
        self._oj = np.concatenate([[0], np.cumsum(c[:, 0])])
        self._op = np.concatenate([[0], np.cumsum(c[:, 1])])
        self._oc = np.concatenate([[0], np.cumsum(c[:, 3])])

    def __len__(self) -> int:
        return self.batch

    def __getitem__(self, b):
        if isinstance(b, slice):
            return [self[i] for i in range(*b.indices(self.batch))]
        if b < 0:
            b += self.batch
        if not 0 <= b < self.batch:
            raise IndexError(b)
        c = self.counts[b]
        j0, j1, p0, p1 = self._oj[b], self._oj[b + 1], self._op[b], self._op[b + 1]
        jl = self.joints[j0:j1].copy() if j1 > j0 else np.array([])          # shape (0,) when empty (:481)
        pa = self.persons[p0:p1].copy() if p1 > p0 else np.array([])         # np.array([]) (:396)
        kp = self.keypoints[p0:p1].copy().reshape(-1, 42)
        lim = []
        o = int(self._oc[b])
        for t in range(_lib.NUM_LIMBS):
            n = int(c[18 + t])
            lim.append(self.limbs[o:o + n].copy())
            o += n
        return DecodeResult(jl, pa, kp, lim, c.copy())

    def __iter__(self):
        return (self[b] for b in range(self.batch))

    @property
    def nbytes(self) -> int:
        """Bytes the device->host copies of this result moved."""
        return int(self.counts.nbytes + self.joints.nbytes + self.limbs.nbytes + self.persons.nbytes
                   + self.keypoints.nbytes)


def _split_packed(batch: int, counts: np.ndarray, joints: np.ndarray, limbs: np.ndarray,
                  persons: np.ndarray, keypoints: np.ndarray) -> BatchResult:
    return BatchResult(batch, counts, joints, limbs, persons, keypoints)


class Engine:
    """One per (device, H, W).  Not thread-safe: serialise calls per engine."""

    def __init__(self, height: int, width: int, max_batch: int = 1, device: int = 0, use_graph: bool = True,
                 fuse_separable: bool = True, stream=None):
        """``stream``: a ``torch.cuda.Stream`` all calls of this engine are issued on (None = the caller's current
        stream at call time).  Engines with their own streams run concurrently (:class:`EnginePool`)."""
        require_gpu(device)
        self.lib = _lib.load()
        self.device = device
        self.height, self.width, self.max_batch = height, width, max_batch
        self.map_h, self.map_w = graph.output_hw(height, width)
        h = C.c_void_p()
        flags = (0 if use_graph else _lib.FLAG_NO_GRAPH) | (0 if fuse_separable else _lib.FLAG_NO_FUSE_SEP)
        _lib.check(self.lib.lwop_create(device, max_batch, height, width, flags, C.byref(h)))
        self.handle = h
        self.stream = stream
        self.weights_source: Optional[str] = None
        self._tables = None
        self._host = None
        self._out = {}
        self._async_tables = []
        self._inflight = {}
        self.last_d2h_bytes = 0

    def close(self) -> None:
        if getattr(self, "handle", None):
            self.lib.lwop_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- weights ---------------------------------------------------------------
    def load_variables(self, variables: Dict[str, np.ndarray], source: str = "variables") -> None:
        blob = graph.pack_blob(graph.fold_all(variables))
        n = self.lib.lwop_weights_num_floats()
        if blob.size != n:
            raise ValueError(f"packed blob has {blob.size} floats, library expects {n}")
        _lib.check(self.lib.lwop_load_weights(self.handle, blob.ctypes.data_as(C.c_void_p), blob.size), self.handle)
        self.weights_source = source

    def load_checkpoint(self, prefix: Optional[str] = None, seed: int = 61236) -> str:
        variables, source = checkpoint.load_or_synthesize(prefix, seed)
        self.load_variables(variables, source)
        return source

    # -- forward -----------------------------------------------------------------
    def _stream(self) -> int:
        if self.stream is not None:
            return self.stream.cuda_stream
        return _torch().cuda.current_stream(self.device).cuda_stream

    def forward(self, images, mode: str = "bf16", return_stage1: bool = False):
        """images: CUDA tensor [B,H,W,3] float32 in [0,1] or uint8.  Returns
        ``(heat [B,h,w,14], paf [B,h,w,26])`` float32 CUDA tensors (+ stage-1 maps).
        The returned tensors are engine-owned and reused by the next call with the
        same batch size."""
        torch = _torch()
        if images.device.type != "cuda" or not images.is_contiguous():
            raise ValueError("images must be a contiguous CUDA tensor")
        B, H, W, Cc = images.shape
        if (H, W, Cc) != (self.height, self.width, 3) or B > self.max_batch:
            raise ValueError(f"engine built for [<= {self.max_batch},{self.height},{self.width},3], got {tuple(images.shape)}")
        heat, paf, h1, p1 = self._outputs(B, images.device)
        m = _MODES[mode]
        if images.dtype == torch.uint8:
            if return_stage1:
                raise ValueError("stage-1 outputs need float32 input")
            _lib.check(self.lib.lwop_forward_u8(self.handle, images.data_ptr(), B, m, heat.data_ptr(), paf.data_ptr(),
                                                self._stream()), self.handle)
            return heat, paf
        if images.dtype != torch.float32:
            raise ValueError("images must be float32 or uint8")
        if return_stage1:
            _lib.check(self.lib.lwop_forward(self.handle, images.data_ptr(), B, m, heat.data_ptr(), paf.data_ptr(),
                                             h1.data_ptr(), p1.data_ptr(), self._stream()), self.handle)
            return heat, paf, h1, p1
        _lib.check(self.lib.lwop_forward(self.handle, images.data_ptr(), B, m, heat.data_ptr(), paf.data_ptr(),
                                         None, None, self._stream()), self.handle)
        return heat, paf

    def _outputs(self, batch: int, dev):
        """Engine-owned output maps per batch size (stable pointers keep the CUDA graph
        valid across calls).  They are overwritten by the next forward of the same
        batch size: clone what must outlive it."""
        torch = _torch()
        bufs = self._out.get(batch)
        if bufs is None:
            shape_h = (batch, self.map_h, self.map_w, _lib.NUM_JOINTS)
            shape_p = (batch, self.map_h, self.map_w, _lib.NUM_PAFS)
            bufs = (torch.empty(shape_h, dtype=torch.float32, device=dev), torch.empty(shape_p, dtype=torch.float32, device=dev),
                    torch.empty(shape_h, dtype=torch.float32, device=dev), torch.empty(shape_p, dtype=torch.float32, device=dev))
            self._out[batch] = bufs
        return bufs

    # -- decode ------------------------------------------------------------------
    def _device_tables(self):
        if self._tables is None:
            torch = _torch()
            B = self.max_batch
            dev = torch.device("cuda", self.device)
            t = {
                "counts": torch.zeros((B, _lib.COUNTS), dtype=torch.int32, device=dev),
                "joints": torch.empty((B, _lib.MAX_JOINTS, 6), dtype=torch.float64, device=dev),
                "limbs": torch.empty((B, _lib.NUM_LIMBS, _lib.MAX_PEAKS, 5), dtype=torch.float64, device=dev),
                "persons": torch.empty((B, _lib.MAX_PERSONS, 16), dtype=torch.float64, device=dev),
                "keypoints": torch.empty((B, _lib.MAX_PERSONS, 42), dtype=torch.float32, device=dev),
            }
            s = _lib.DecodeTables(t["counts"].data_ptr(), t["joints"].data_ptr(), t["limbs"].data_ptr(),
                                  t["persons"].data_ptr(), t["keypoints"].data_ptr())
            self._tables = (t, s)
        return self._tables

    def _new_host_tables(self, batch: int, scale: int = 1):
        """Pinned host result buffers.  The asynchronous path downloads them at full capacity, so the defaults are
        sized for typical frames (<= 128 joints / connections, 16 persons per frame) and grown on overflow."""
        torch = _torch()
        nj, nl, npers = batch * 128 * scale, batch * 128 * scale, batch * 16 * scale
        return {
            "batch": batch,
            "counts": torch.zeros((batch, _lib.COUNTS), dtype=torch.int32).pin_memory(),
            "joints": torch.empty((nj, 6), dtype=torch.float64).pin_memory(),
            "limbs": torch.empty((nl, 5), dtype=torch.float64).pin_memory(),
            "persons": torch.empty((npers, 16), dtype=torch.float64).pin_memory(),
            "keypoints": torch.empty((npers, 42), dtype=torch.float32).pin_memory(),
        }

    def _host_tables(self, batch: int):
        """Pinned host result buffers of the synchronous calls, grown on demand."""
        if self._host is None or self._host["batch"] < batch:
            self._host = self._new_host_tables(batch)
        return self._host

    def _packed_struct(self, hst) -> _lib.PackedTables:
        return _lib.PackedTables(hst["counts"].data_ptr(), hst["joints"].data_ptr(), hst["limbs"].data_ptr(),
                                 hst["persons"].data_ptr(), hst["keypoints"].data_ptr(),
                                 hst["joints"].shape[0], hst["limbs"].shape[0], hst["persons"].shape[0])

    @staticmethod
    def _table_bytes(hst, batch: int) -> int:
        return int(batch * _lib.COUNTS * 4 + hst["joints"].numel() * 8 + hst["limbs"].numel() * 8
                   + hst["persons"].numel() * 8 + hst["keypoints"].numel() * 4)

    def _grow_host(self, hst, msg: str) -> bool:
        """Overflow of the *host* buffers: quadruple them and let the caller retry."""
        if "host result buffers too small" not in msg:
            return False
        torch = _torch()
        for k, w, dt in (("joints", 6, torch.float64), ("limbs", 5, torch.float64),
                         ("persons", 16, torch.float64), ("keypoints", 42, torch.float32)):
            hst[k] = torch.empty((hst[k].shape[0] * 4, w), dtype=dt).pin_memory()
        return True

    def decode_device(self, heat, paf, img_hw: Optional[Tuple[int, int]] = None, thre1: float = 0.1,
                      thre2: float = 0.0) -> None:
        """Runs the decode kernels on CUDA tensors ``heat [B,h,w,14]``, ``paf [B,h,w,26]``
        (float32, contiguous); results stay in the engine's device tables."""
        B, h, w, _ = heat.shape
        if B > self.max_batch:
            raise ValueError("batch exceeds engine capacity")
        H, W = img_hw if img_hw is not None else (self.height, self.width)
        _, s = self._device_tables()
        _lib.check(self.lib.lwop_decode(self.handle, heat.data_ptr(), paf.data_ptr(), B, h, w, H, W,
                                        float(np.float32(thre1)), float(thre2), C.byref(s), self._stream()),
                   self.handle)

    def forward_decode(self, images, mode: str = "bf16", thre1: float = 0.1, thre2: float = 0.0):
        """Forward + decode chain as ONE stream operation (one CUDA graph in bf16 mode): ``images`` as in
        :meth:`forward`; returns the engine-owned maps, the decode results stay in the engine's device tables
        (:meth:`fetch`)."""
        torch = _torch()
        if images.device.type != "cuda" or not images.is_contiguous():
            raise ValueError("images must be a contiguous CUDA tensor")
        B, H, W, Cc = images.shape
        if (H, W, Cc) != (self.height, self.width, 3) or B > self.max_batch:
            raise ValueError(f"engine built for [<= {self.max_batch},{self.height},{self.width},3], got {tuple(images.shape)}")
        if images.dtype not in (torch.uint8, torch.float32):
            raise ValueError("images must be float32 or uint8")
        heat, paf, _, _ = self._outputs(B, images.device)
        _, s = self._device_tables()
        _lib.check(self.lib.lwop_forward_decode(self.handle, images.data_ptr(), int(images.dtype == torch.uint8), B,
                                                _MODES[mode], float(np.float32(thre1)), float(thre2), heat.data_ptr(),
                                                paf.data_ptr(), C.byref(s), self._stream()), self.handle)
        return heat, paf

    def draw_poses(self, frames, tables=None, out=None):
        """Canvases of ``plot_pose`` (reference pose_decode.py:431-436) drawn on the device: ``frames`` uint8 CUDA
        ``[B,H,W,3]`` (any H, W); ``tables`` = dict of device tables (``counts``, ``joints``, ``persons``; default: the
        engine's own, i.e. the result of the last :meth:`decode_device` / :meth:`forward_decode`).  Returns a new uint8
        tensor (or ``out``, which may be ``frames`` itself), byte-identical to cv2's drawing."""
        torch = _torch()
        if frames.device.type != "cuda" or frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3 \
                or not frames.is_contiguous():
            raise ValueError("frames must be a contiguous uint8 CUDA tensor [B,H,W,3]")
        B, H, W, _ = frames.shape
        if tables is None:
            if B > self.max_batch:
                raise ValueError("batch exceeds engine capacity")
            _, s = self._device_tables()
        else:
            for k, dt, tail in (("counts", torch.int32, (_lib.COUNTS,)), ("joints", torch.float64, (_lib.MAX_JOINTS, 6)),
                                ("persons", torch.float64, (_lib.MAX_PERSONS, 16))):
                t = tables[k]
                if t.device != frames.device or t.dtype != dt or tuple(t.shape[1:]) != tail or t.shape[0] < B \
                        or not t.is_contiguous():
                    raise ValueError(f"tables['{k}'] must be a contiguous {dt} CUDA tensor [>= {B}, {tail}]")
            s = _lib.DecodeTables(tables["counts"].data_ptr(), tables["joints"].data_ptr(), None,
                                  tables["persons"].data_ptr(), None)
        if out is None:
            out = torch.empty_like(frames)
        elif out.shape != frames.shape or out.dtype != torch.uint8 or out.device != frames.device or not out.is_contiguous():
            raise ValueError("out must match frames")
        _lib.check(self.lib.lwop_draw_poses(self.handle, frames.data_ptr(), B, H, W, C.byref(s), out.data_ptr(),
                                            self._stream()), self.handle)
        return out

    # -- decode stages, separately callable ------------------------------------------
    def nms_device(self, heat, factor: float, thre1: float, mode: int = 0):
        """Peak search + refinement on ``heat [B,h,w,14]`` (CUDA float32).  Returns the device peak table
        ``(peaks [B,14,MAX_PEAKS,4], peak_counts [B,16], counts [B,COUNTS])``."""
        torch = _torch()
        B, h, w, _ = heat.shape
        dev = heat.device
        peaks = torch.zeros((B, _lib.NUM_JOINTS, _lib.MAX_PEAKS, 4), dtype=torch.float32, device=dev)
        pc = torch.zeros((B, 16), dtype=torch.int32, device=dev)
        counts = torch.zeros((B, _lib.COUNTS), dtype=torch.int32, device=dev)
        _lib.check(self.lib.lwop_nms(self.handle, heat.data_ptr(), B, h, w, float(factor), float(np.float32(thre1)),
                                     int(mode), peaks.data_ptr(), pc.data_ptr(), counts.data_ptr(), self._stream()),
                   self.handle)
        return peaks, pc, counts

    def connect_device(self, paf, upsampled: bool, map_hw, img_hw, thre2: float, peaks, pc, counts,
                       num_intermed_pts: int = 10, max_dist_thresh: float = 1.0):
        """find_connected_joints on the device peak table; returns ``limbs [B,13,MAX_PEAKS,5]`` (float64)."""
        torch = _torch()
        B = paf.shape[0]
        limbs = torch.zeros((B, _lib.NUM_LIMBS, _lib.MAX_PEAKS, 5), dtype=torch.float64, device=paf.device)
        _lib.check(self.lib.lwop_find_connected_joints(self.handle, paf.data_ptr(), int(bool(upsampled)), B,
                                                       int(map_hw[0]), int(map_hw[1]), int(img_hw[0]), int(img_hw[1]),
                                                       float(thre2), int(num_intermed_pts), float(max_dist_thresh),
                                                       peaks.data_ptr(), pc.data_ptr(), limbs.data_ptr(),
                                                       counts.data_ptr(), self._stream()), self.handle)
        return limbs

    def group_device(self, peaks, pc, limbs, counts):
        """Joint list + grouping + keypoint table from device peak / limb tables; returns device tensors
        ``(joints [B,MAX_JOINTS,6], persons [B,MAX_PERSONS,16], keypoints [B,MAX_PERSONS,42])``; ``counts`` is updated."""
        torch = _torch()
        B = peaks.shape[0]
        if B > self.max_batch:
            raise ValueError("batch exceeds engine capacity")
        dev = peaks.device
        joints = torch.zeros((B, _lib.MAX_JOINTS, 6), dtype=torch.float64, device=dev)
        persons = torch.zeros((B, _lib.MAX_PERSONS, 16), dtype=torch.float64, device=dev)
        kp = torch.zeros((B, _lib.MAX_PERSONS, 42), dtype=torch.float32, device=dev)
        t = _lib.DecodeTables(counts.data_ptr(), joints.data_ptr(), limbs.data_ptr(), persons.data_ptr(), kp.data_ptr())
        _lib.check(self.lib.lwop_group_limbs(self.handle, peaks.data_ptr(), pc.data_ptr(), B, C.byref(t), self._stream()),
                   self.handle)
        return joints, persons, kp

    def fetch(self, batch: int) -> BatchResult:
        _, s = self._device_tables()
        hst = self._host_tables(batch)
        while True:
            p = self._packed_struct(hst)
            rc = self.lib.lwop_decode_fetch(self.handle, C.byref(s), batch, C.byref(p), self._stream())
            if rc == _lib.ERR_OVERFLOW:
                msg = self.lib.lwop_last_error(self.handle).decode()
                if self._grow_host(hst, msg):
                    continue
            _lib.check(rc, self.handle)
            break
        return _split_packed(batch, hst["counts"].numpy(), hst["joints"].numpy(), hst["limbs"].numpy(),
                             hst["persons"].numpy(), hst["keypoints"].numpy())

    def decode(self, heat, paf, img_hw: Optional[Tuple[int, int]] = None, thre1: float = 0.1,
               thre2: float = 0.0) -> BatchResult:
        self.decode_device(heat, paf, img_hw, thre1, thre2)
        return self.fetch(heat.shape[0])

    # -- whole path from host memory --------------------------------------------
    def infer_host(self, images_host, mode: str = "bf16", thre1: float = 0.1, thre2: float = 0.0,
                   want_maps: bool = False):
        """images_host: pinned (or pageable) CPU tensor / ndarray [B,H,W,3] float32 or uint8.
        One C call: H2D, forward, decode, D2H.  Returns a :class:`BatchResult`
        (and the host maps when ``want_maps``)."""
        torch = _torch()
        images_host = validate_host_images(images_host, self.height, self.width, self.max_batch)
        B = images_host.shape[0]
        is_u8 = images_host.dtype == torch.uint8
        hst = self._host_tables(B)
        heat_h = paf_h = None
        if want_maps:
            heat_h = torch.empty((B, self.map_h, self.map_w, _lib.NUM_JOINTS), dtype=torch.float32).pin_memory()
            paf_h = torch.empty((B, self.map_h, self.map_w, _lib.NUM_PAFS), dtype=torch.float32).pin_memory()
        while True:
            p = self._packed_struct(hst)
            rc = self.lib.lwop_infer_host(self.handle, images_host.data_ptr(), int(is_u8), B, _MODES[mode],
                                          float(np.float32(thre1)), float(thre2), C.byref(p),
                                          heat_h.data_ptr() if want_maps else None,
                                          paf_h.data_ptr() if want_maps else None, self._stream())
            if rc == _lib.ERR_OVERFLOW:
                msg = self.lib.lwop_last_error(self.handle).decode()
                if self._grow_host(hst, msg):
                    continue
            _lib.check(rc, self.handle)
            break
        self.last_d2h_bytes = self._table_bytes(hst, B)
        res = _split_packed(B, hst["counts"].numpy(), hst["joints"].numpy(), hst["limbs"].numpy(),
                            hst["persons"].numpy(), hst["keypoints"].numpy())
        if want_maps:
            return res, heat_h.numpy(), paf_h.numpy()
        return res

    def infer_host_begin(self, images_host, mode: str = "bf16", thre1: float = 0.1, thre2: float = 0.0,
                         tables=None) -> int:
        """Asynchronous form: enqueues upload, forward, decode and the download of the result tables for this
        batch and returns a ticket at once; :meth:`infer_host_end` waits for it.  Keeping two batches in flight
        (begin k+1 before end k) overlaps the upload of one with the kernels of the other.  ``images_host``
        must be pinned and stay alive until the matching ``infer_host_end``.  ``tables``: caller-owned host result
        buffers (a dict shaped like :meth:`_new_host_tables` returns, e.g. a slice of a shared-memory ring the results
        of several ranks are gathered in); default: engine-owned pinned buffers."""
        torch = _torch()
        images_host = validate_host_images(images_host, self.height, self.width, self.max_batch)
        B = images_host.shape[0]
        is_u8 = images_host.dtype == torch.uint8
        if tables is not None:
            if tables["batch"] < B:
                raise ValueError("result tables are smaller than the batch")
            hst = tables
        else:
            free = [t for t in self._async_tables if t["batch"] >= B and not t.get("busy")]
            hst = free[0] if free else self._new_host_tables(B)
            if not free:
                self._async_tables.append(hst)
        p = self._packed_struct(hst)
        ticket = C.c_int(-1)
        rc = self.lib.lwop_infer_host_begin(self.handle, images_host.data_ptr(), int(is_u8), B, _MODES[mode],
                                            float(np.float32(thre1)), float(thre2), C.byref(p), None, None,
                                            self._stream(), C.byref(ticket))
        _lib.check(rc, self.handle)
        hst["busy"] = True
        self._inflight[ticket.value] = (B, hst, images_host, mode, thre1, thre2)
        return ticket.value

    def infer_host_end(self, ticket: int) -> "BatchResult":
        B, hst, images_host, mode, thre1, thre2 = self._inflight.pop(ticket)
        rc = self.lib.lwop_infer_host_end(self.handle, ticket)
        hst["busy"] = False
        if rc == _lib.ERR_OVERFLOW and not hst.get("external"):
            msg = self.lib.lwop_last_error(self.handle).decode()
            if self._grow_host(hst, msg):
                # this batch needs larger result buffers: run it again synchronously with the grown ones
                self._host = hst
                return self.infer_host(images_host, mode, thre1, thre2)
        _lib.check(rc, self.handle)
        self.last_d2h_bytes = self._table_bytes(hst, B)
        return BatchResult(B, hst["counts"].numpy(), hst["joints"].numpy(), hst["limbs"].numpy(),
                           hst["persons"].numpy(), hst["keypoints"].numpy(), copy=not hst.get("external"))

    def step_descs(self):
        """[(layer, is_depthwise, in_h, in_w, cin, cout, ksize, stride, dilation, kernel), ...] with kernel
        0 = CUDA cores, 1 = tcgen05 GEMM, 2 = fused depthwise+pointwise (runs at this step), 3 = absorbed."""
        n = self.lib.lwop_num_steps(self.handle)
        out = []
        buf = (C.c_int * 10)()
        for i in range(n):
            _lib.check(self.lib.lwop_step_desc(self.handle, i, buf), self.handle)
            out.append(tuple(int(v) for v in buf))
        return out

    def profile_forward(self, images, mode: str = "bf16"):
        """One eager (graph-less) forward with CUDA events around every kernel; returns the per-step ms."""
        torch = _torch()
        B = images.shape[0]
        heat = torch.empty((B, self.map_h, self.map_w, _lib.NUM_JOINTS), dtype=torch.float32, device=images.device)
        paf = torch.empty((B, self.map_h, self.map_w, _lib.NUM_PAFS), dtype=torch.float32, device=images.device)
        n = self.lib.lwop_num_steps(self.handle)
        ms = (C.c_float * n)()
        _lib.check(self.lib.lwop_profile_forward(self.handle, images.data_ptr(), B, _MODES[mode], heat.data_ptr(),
                                                 paf.data_ptr(), ms, self._stream()), self.handle)
        return [float(v) for v in ms]

    def set_option(self, name: str, value: int) -> None:
        _lib.check(self.lib.lwop_set_option(self.handle, name.encode(), int(value)), self.handle)

    def launch_count(self, reset: bool = False) -> int:
        return int(self.lib.lwop_launch_count(self.handle, int(reset)))


class EnginePool:
    """``lanes`` engines (one ``lwop_handle`` each: own activation arena, CUDA graphs, copy / decode streams) on ONE
    device, each on its own CUDA stream, fed round-robin.

    Why: every forward kernel is a persistent grid with a fill and a drain, the decode chain is a few latency-bound
    kernels, and both leave SMs idle at small batch.  With several batches in flight on independent streams the
    hardware fills those gaps with the other lanes' kernels (measured on B200, 32 frames per batch: 24.2k images/s
    with one lane, 30.5k with three).  Results come back in submission order.

    ``submit`` / ``result`` is the host-buffer path (upload, forward, decode, download per batch);
    ``step_device`` / ``join`` the device-resident one."""

    def __init__(self, height: int, width: int, max_batch: int, lanes: int = 3, device: int = 0, depth: int = 1):
        """``depth``: host batches in flight per lane (1 or 2; with 2 the upload of a lane's next batch overlaps the
        kernels of its current one inside the handle's two-slot pipeline)."""
        torch = _torch()
        require_gpu(device)
        if lanes < 1:
            raise ValueError("lanes must be >= 1")
        if depth not in (1, 2):
            raise ValueError("depth must be 1 or 2")
        self.device = device
        self.capacity = lanes * depth
        self.engines = [Engine(height, width, max_batch=max_batch, device=device,
                               stream=torch.cuda.Stream(device=device)) for _ in range(lanes)]
        # one upload queue for the pool: with a private copy stream per engine the uploads of every batch in flight run
        # concurrently and share PCIe, so the FIRST batch lands only when nearly all have (a long pipeline fill);
        # through one stream they land in submission order, each at the full rate
        self._copy_stream = torch.cuda.Stream(device=device)
        if os.environ.get("LWOP_POOL_SHARED_COPY", "1") != "0":
            for e in self.engines:
                e.set_option("copy_stream", self._copy_stream.cuda_stream)
        self._next = 0
        self._pending: List[Tuple[int, int]] = []          # (lane, ticket) in submission order
        self._events = [torch.cuda.Event() for _ in range(lanes)]

    def __len__(self) -> int:
        return len(self.engines)

    def close(self) -> None:
        for e in self.engines:
            e.close()

    def load_variables(self, variables: Dict[str, np.ndarray], source: str = "variables") -> None:
        for e in self.engines:
            e.load_variables(variables, source)

    def load_checkpoint(self, prefix: Optional[str] = None, seed: int = 61236) -> str:
        variables, source = checkpoint.load_or_synthesize(prefix, seed)
        self.load_variables(variables, source)
        return source

    @property
    def in_flight(self) -> int:
        return len(self._pending)

    def submit(self, images_host, mode: str = "bf16", thre1: float = 0.1, thre2: float = 0.0, tables=None) -> None:
        """Enqueue one batch of host frames on the next lane (at most ``depth`` batches per lane in flight: call
        :meth:`result` first when ``in_flight == pool.capacity``)."""
        if len(self._pending) >= self.capacity:
            raise RuntimeError("every lane is busy: collect a result() first")
        lane = self._next
        self._next = (self._next + 1) % len(self.engines)
        ticket = self.engines[lane].infer_host_begin(images_host, mode, thre1, thre2, tables=tables)
        self._pending.append((lane, ticket))

    def result(self) -> "BatchResult":
        """Wait for the oldest submitted batch."""
        lane, ticket = self._pending.pop(0)
        return self.engines[lane].infer_host_end(ticket)

    def map(self, batches, mode: str = "bf16", thre1: float = 0.1, thre2: float = 0.0):
        """Generator over the results of an iterable of host batches, lanes kept full, submission order."""
        for images in batches:
            if self.in_flight == self.capacity:
                yield self.result()
            self.submit(images, mode, thre1, thre2)
        while self.in_flight:
            yield self.result()

    def step_device(self, images, mode: str = "bf16", thre1: float = 0.1, thre2: float = 0.0) -> int:
        """Forward + decode of a device-resident batch on the next lane's stream (one CUDA graph launch); the
        results stay in that lane's device tables.  Returns the lane index."""
        lane = self._next
        self._next = (self._next + 1) % len(self.engines)
        e = self.engines[lane]
        e.forward_decode(images, mode, thre1, thre2)
        return lane

    def fork(self) -> None:
        """Make every lane's stream wait for the work already queued on the caller's current stream."""
        torch = _torch()
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        for e in self.engines:
            e.stream.wait_event(ev)

    def join(self) -> None:
        """Make the caller's current stream wait for everything queued on the lanes."""
        torch = _torch()
        cur = torch.cuda.current_stream(self.device)
        for ev, e in zip(self._events, self.engines):
            ev.record(e.stream)
            cur.wait_event(ev)


# ---------------------------------------------------------------------------
# process-wide engine cache used by the reference-shaped API in ``.src``
# ---------------------------------------------------------------------------
_engines: Dict[Tuple[int, int, int], Engine] = {}
_engines_lock = threading.RLock()
_variables: Optional[Dict[str, np.ndarray]] = None
_variables_source: Optional[str] = None


def set_variables(variables: Dict[str, np.ndarray], source: str = "variables") -> None:
    """Bind the network weights for the reference-shaped API (replaces Saver.restore)."""
    global _variables, _variables_source
    _variables, _variables_source = variables, source
    with _engines_lock:
        if graph.fits_executor(variables):
            for e in _engines.values():
                e.load_variables(variables, source)
        else:
            # head widths other than the executor's 14 / 26: only the op-level (eager) graph can run these variables;
            # the cached whole-network engines would keep the previous weights, so they go
            for e in _engines.values():
                e.close()
            _engines.clear()


def load_weights(prefix: Optional[str] = None, seed: int = 61236) -> str:
    variables, source = checkpoint.load_or_synthesize(prefix, seed)
    set_variables(variables, source)
    return source


def get_variables() -> Dict[str, np.ndarray]:
    if _variables is None:
        load_weights()
    return _variables


def get_engine(height: int, width: int, batch: int, device: int = 0) -> Engine:
    key = (device, height, width)
    with _engines_lock:
        e = _engines.get(key)
        if e is None or e.max_batch < batch:
            if e is not None:
                e.close()
            e = Engine(height, width, max_batch=max(batch, 1), device=device)
            e.load_variables(get_variables(), _variables_source or "variables")
            _engines[key] = e
        return e
