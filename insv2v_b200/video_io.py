"""Frame I/O around the edit: the reference's `LoveuTgveVideoDataset.load_frames`
(dataset/loveu_tgve_dataset.py:44-58) and `save_tensor_to_gif` / `save_tensor_to_images`
(misc_utils/image_utils.py:127-132,233-241) with the per-pixel work on the GPU.

Decode and resize stay on the host with the reference's own calls (`cv2.VideoCapture`, `cv2.resize`: there is no
NVDEC path without the network-installed codecs, and cv2's fixed-point bilinear resize is what the reference's pixels
are), but each frame lands in ONE pinned uint8 staging buffer, crosses PCIe once as bytes (4x fewer than the fp32
tensors the reference ships with `.cuda()`), and BGR->RGB + ToTensor + Normalize is a single kernel whose output is
bit-identical to the reference's torchvision transform. The way back is symmetric: one kernel to uint8 HWC, one D2H
copy of bytes, then the GIF / JPEG encoders on the host."""
import csv
import os

import numpy as np
import torch

from . import lib as _lib
from . import ops


def frames_u8_to_tensor(frames_u8, bgr=True, device=None):
    """uint8 [n, h, w, 3] (numpy or torch, host or device) -> fp32 [n, 3, h, w] in [-1, 1] on the GPU."""
    t = torch.from_numpy(frames_u8) if isinstance(frames_u8, np.ndarray) else frames_u8
    if t.dtype != torch.uint8 or t.dim() != 4 or t.shape[-1] != 3:
        raise ValueError(f"expected uint8 [n, h, w, 3], got {t.dtype} {tuple(t.shape)}")
    if not t.is_cuda:
        dev = torch.device("cuda") if device is None else torch.device(device)
        t = (t if t.is_pinned() else t.pin_memory()).to(dev, non_blocking=True)
    t = t.contiguous()
    n, h, w, _ = t.shape
    out = torch.empty((n, 3, h, w), dtype=torch.float32, device=t.device)
    _lib.check(_lib.load().ivv_frames_u8_to_f32(ops._p(t), ops._p(out), n, h * w, int(bgr), ops._s()),
               "ivv_frames_u8_to_f32")
    ops._count()
    return out


def tensor_to_frames_u8(images):
    """[n, 3, h, w] or [1, n, 3, h, w] fp32 / fp16 in [-1, 1] on the GPU -> uint8 [n, h, w, 3] (RGB) on the GPU."""
    if images.dim() == 5:
        images = images.squeeze(0)
    if not images.is_cuda or images.dtype not in (torch.float32, torch.float16) or images.dim() != 4 or images.shape[1] != 3:
        raise ValueError(f"expected a CUDA fp32/fp16 [n, 3, h, w] tensor, got {images.dtype} {tuple(images.shape)}")
    images = images.contiguous()
    n, _, h, w = images.shape
    out = torch.empty((n, h, w, 3), dtype=torch.uint8, device=images.device)
    _lib.check(_lib.load().ivv_frames_to_u8(ops._p(images), int(images.dtype == torch.float32), ops._p(out), n, h * w,
                                            ops._s()), "ivv_frames_to_u8")
    ops._count()
    return out


def load_video_frames(video_path, image_size, device="cuda", max_frames=None):
    """LoveuTgveVideoDataset.load_frames: every frame of the file, cv2.resize to image_size (w, h), -> fp32
    [n, 3, h, w] in [-1, 1] on `device`."""
    import cv2
    cap = cv2.VideoCapture(video_path)
    if not cap.isOpened():
        raise FileNotFoundError(f"cannot open video {video_path}")
    frames = []
    while True:
        ret, frame = cap.read()
        if not ret:
            break
        frames.append(cv2.resize(frame, tuple(image_size)))
        if max_frames is not None and len(frames) >= max_frames:
            break
    cap.release()
    if not frames:
        raise ValueError(f"{video_path}: no frames decoded")
    staging = torch.empty((len(frames),) + frames[0].shape, dtype=torch.uint8).pin_memory()
    np.stack(frames, axis=0, out=staging.numpy())
    return frames_u8_to_tensor(staging, bgr=True, device=device)


class LoveuTgveVideoDataset(torch.utils.data.Dataset):
    """Same constructor, index contract and item keys as the reference class (dataset/loveu_tgve_dataset.py:9-83);
    `item['frames']` is already on the GPU."""

    def __init__(self, root_dir, image_size=(480, 480), device="cuda"):
        self.root_dir, self.image_size, self.device = root_dir, image_size, device
        self.data = {}
        with open(os.path.join(root_dir, "LOVEU-TGVE-2023_Dataset.csv"), "r") as file:
            reader = csv.reader(file)
            next(reader, None)  # skip the headers
            for row in reader:
                if len(row[0]) == 0:
                    continue
                if row[0].endswith("Videos:"):
                    dataset_type = row[0].split(" ")[0]
                    self.source_folder = (dataset_type if dataset_type == "DAVIS" else dataset_type.lower()) + \
                        "_480p/480p_videos"
                elif len(row) > 1:
                    self.data[row[0]] = {"video_name": row[0], "original": row[1], "style": row[2], "object": row[3],
                                         "background": row[4], "multiple": row[5], "source_folder": self.source_folder}

    def _path(self, video_name, source_folder):
        return os.path.join(self.root_dir, source_folder, f"{video_name}.mp4")

    def load_frames(self, video_name, source_folder):
        return load_video_frames(self._path(video_name, source_folder), self.image_size, self.device)

    def load_fps(self, video_name, source_folder):
        import cv2
        cap = cv2.VideoCapture(self._path(video_name, source_folder))
        fps = cap.get(cv2.CAP_PROP_FPS)
        cap.release()
        return fps

    def __len__(self):
        return len(self.data)

    def __getitem__(self, idx):
        video_name = idx if isinstance(idx, str) else list(self.data.keys())[idx]
        item = self.data[video_name].copy()
        item["frames"] = self.load_frames(video_name, item["source_folder"])
        item["fps"] = self.load_fps(video_name, item["source_folder"])
        return item


def _to_host_u8(images):
    u8 = tensor_to_frames_u8(images)
    host = torch.empty(u8.shape, dtype=torch.uint8).pin_memory()
    host.copy_(u8, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return host.numpy()


def save_tensor_to_gif(images, filename, fps):
    """misc_utils/image_utils.py:233-235 + images_to_gif (:127-132); encoder: PIL (imageio is not a dependency here)."""
    from PIL import Image
    frames = [Image.fromarray(f) for f in _to_host_u8(images)]
    os.makedirs(os.path.dirname(filename) or ".", exist_ok=True)
    frames[0].save(filename, save_all=True, append_images=frames[1:], duration=int(round(1000.0 / fps)), loop=0)


def save_tensor_to_images(images, output_dir):
    """misc_utils/image_utils.py:237-241: one '{i:03d}.jpg' per frame."""
    import cv2
    os.makedirs(output_dir, exist_ok=True)
    for i, f in enumerate(_to_host_u8(images)):
        cv2.imwrite(f"{output_dir}/{i:03d}.jpg", cv2.cvtColor(f, cv2.COLOR_RGB2BGR))
