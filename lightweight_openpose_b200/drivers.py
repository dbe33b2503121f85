"""Batched replacements for the reference's inference drivers.

``test&eval/model_test.py:81-102`` (image folder / video loop) and
``test&eval/model_json.py:63-112`` (AI-Challenger prediction JSON) run, per frame::

    img_data = cv2.cvtColor(img_ori, cv2.COLOR_BGR2RGB); img_data = cv2.resize(img_data, (S, S)); img = img_data / 255.
    heatmap, paf = sess.run([cpm, paf], {input_img: [img]})
    canvas, joint_list, person_to_joint_assoc, joints = decode_pose(img_data, params, heatmap[0], paf[0])

Here the frames of a batch are uploaded as uint8 BGR (4x less PCIe traffic than float32), and colour swap, resize,
``/ 255.``, the network and the decode all run on the GPU; only the small result tables come back.  The canvas
(``plot_pose`` drawing) is produced on request: on the host with cv2 from the kept frame (``keep_frames=True``), as in
``pose_decode.decode_pose``, or by the CUDA rasteriser straight from the device-resident result tables
(``draw=True``; the same bytes).
"""
from __future__ import annotations

import ctypes as C
import json
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib, engine
from .src import pose_decode

DEFAULT_PARAMS = {"thre1": 0.1, "thre2": 0.0}          # test&eval/model_test.py:27-28, model_json.py:32-33


def _torch():
    import torch
    return torch


class FrameResult:
    """What one iteration of the reference loop produces (model_test.py:68-71)."""
    __slots__ = ("joint_list", "person_to_joint_assoc", "joints", "img_data", "orig_hw", "canvas_data")

    def __init__(self, res: "engine.DecodeResult", img_data: Optional[np.ndarray], orig_hw: Tuple[int, int],
                 canvas_data: Optional[np.ndarray] = None):
        self.joint_list = res.joint_list
        self.person_to_joint_assoc = res.person_to_joint_assoc
        self.joints = res.joints
        self.img_data = img_data            # resized RGB uint8 frame (only when requested)
        self.orig_hw = orig_hw
        self.canvas_data = canvas_data      # canvas rasterised on the device (only with draw=True)

    def canvas(self) -> np.ndarray:
        """The drawn frame of ``plot_pose`` (pose_decode.py:431-436); needs ``draw=True`` (drawn on the device) or
        ``keep_frames=True`` (drawn here with cv2)."""
        if self.canvas_data is not None:
            return self.canvas_data
        if self.img_data is None:
            raise ValueError("run with draw=True or keep_frames=True to get canvases")
        if not self.person_to_joint_assoc.size:
            return self.img_data.copy()
        return pose_decode.plot_pose(self.img_data, self.joint_list, self.person_to_joint_assoc)[0]


class PoseRunner:
    """Owns one engine for a network input size; ``run_frames`` is the batched loop body."""

    def __init__(self, input_size: Tuple[int, int] = (256, 256), params: Optional[Dict[str, float]] = None,
                 max_batch: int = 64, device: int = 0, checkpoint: Optional[str] = None):
        self.H, self.W = int(input_size[0]), int(input_size[1])
        self.params = dict(DEFAULT_PARAMS if params is None else params)
        self.device = device
        self.max_batch = max_batch
        self.engine = engine.Engine(self.H, self.W, max_batch=max_batch, device=device)
        self.weights_source = self.engine.load_checkpoint(checkpoint)
        # network-size frames per batch size: the engine keys its CUDA graphs on the input pointer, so the staging
        # tensor is owned here and reused (a fresh tensor per call would re-capture a graph whenever the allocator
        # hands out a new address)
        self._staging = {}

    def preprocess(self, frames_bgr: np.ndarray):
        """uint8 BGR ``[B, Hs, Ws, 3]`` (host) -> uint8 RGB ``[B, H, W, 3]`` CUDA tensor at the network size."""
        torch = _torch()
        frames_bgr = np.ascontiguousarray(frames_bgr)
        if frames_bgr.dtype != np.uint8 or frames_bgr.ndim != 4 or frames_bgr.shape[-1] != 3:
            raise ValueError("frames must be uint8 [B, H, W, 3] (BGR, as cv2.imread returns them)")
        B, Hs, Ws, _ = frames_bgr.shape
        dev = torch.device("cuda", self.device)
        src = torch.from_numpy(frames_bgr).to(dev, non_blocking=True)
        dst = self._staging.get(B)
        if dst is None:
            dst = self._staging[B] = torch.empty((B, self.H, self.W, 3), dtype=torch.uint8, device=dev)
        e = self.engine
        _lib.check(e.lib.lwop_preprocess_bgr_u8(e.handle, src.data_ptr(), B, Hs, Ws, dst.data_ptr(), self.H, self.W,
                                                e._stream()), e.handle)
        return dst

    def run_frames(self, frames_bgr: Sequence[np.ndarray], keep_frames: bool = False, draw: bool = False) -> List[FrameResult]:
        """``frames_bgr``: uint8 BGR frames as ``cv2.imread`` / ``VideoCapture.read`` return them (any sizes).
        Frames of equal size are batched together; results come back in input order.  ``draw=True`` rasterises the
        canvases of ``plot_pose`` on the device from the decode tables (``lwop_draw_poses``) and downloads them."""
        out: List[Optional[FrameResult]] = [None] * len(frames_bgr)
        by_size: Dict[Tuple[int, int], List[int]] = {}
        for i, f in enumerate(frames_bgr):
            by_size.setdefault((f.shape[0], f.shape[1]), []).append(i)
        for (hs, ws), idx in by_size.items():
            for k in range(0, len(idx), self.max_batch):
                part = idx[k:k + self.max_batch]
                batch = np.stack([frames_bgr[i] for i in part])
                rgb = self.preprocess(batch)            # valid until the next preprocess() of the same batch size
                self.engine.forward_decode(rgb, "bf16", self.params["thre1"], self.params["thre2"])
                canvases = self.engine.draw_poses(rgb).cpu().numpy() if draw else None
                res = self.engine.fetch(len(part))
                frames_host = rgb.cpu().numpy() if keep_frames else None
                for j, i in enumerate(part):
                    out[i] = FrameResult(res[j], frames_host[j] if keep_frames else None, (hs, ws),
                                         canvases[j] if draw else None)
        return out            # type: ignore[return-value]


def keypoint_annotations(runner: PoseRunner, labels: Iterable[dict], load_image: Callable[[str], np.ndarray],
                         batch: Optional[int] = None) -> List[dict]:
    """The prediction list of ``model_json.py:63-103``: every label dict gets ``keypoint_annotations`` =
    ``{'human1': [x, y, v] * 14, ...}`` with the coordinates rescaled from the network frame to the original image
    (``factorX = w_orig / w_in``, ``factorY = h_orig / h_in``, :79-85).  ``load_image(image_id)`` returns the BGR frame."""
    labels = list(labels)
    batch = batch or runner.max_batch
    predictions: List[dict] = []
    for k in range(0, len(labels), batch):
        part = labels[k:k + batch]
        frames = [load_image(lb["image_id"]) for lb in part]
        for lb, fr in zip(part, runner.run_frames(frames)):
            factor_y, factor_x = fr.orig_hw[0] / runner.H, fr.orig_hw[1] / runner.W
            kps = {}
            for human, joint in enumerate(fr.joints, start=1):
                joint = np.array(joint, dtype=np.float32, copy=True)        # float32 row, scaled in place like :82-85
                joint[0::3] *= factor_x
                joint[1::3] *= factor_y
                kps["human" + str(human)] = joint.tolist()
            pred = dict(lb)
            pred["keypoint_annotations"] = kps
            predictions.append(pred)
    return predictions


def write_predictions(predictions: List[dict], path: str) -> None:
    """``json.dump(predictions, fw)`` of model_json.py:111-112."""
    with open(path, "w") as fw:
        json.dump(predictions, fw)
