"""Host side of the data path in front of the GPU preprocessing.

The reference feeds one GPU from 8 + 4 DataLoader worker processes that decode, resize, normalise and stack every clip on
the CPU (Code/dataset/two_stream_dataset.py:72-105, 491-537; Code/run_helper/test_helper.py:414-417) and then copy fp32
clips -- five times the bytes of the frames they are cut from -- to the device.  Here the host only DECODES:

* `read_flo`       Middlebury .flo reader (reference Code/utils/flowlib.py:589-611), failing loudly on a bad magic number
* `decode_frame`   JPEG/PNG -> BGR uint8 via cv2 (libjpeg-turbo, the decoder family of the reference's TurboJPEG wrapper)
* `VideoLoader`    decode threads fill PINNED staging buffers chunk by chunk; each chunk is copied to the device on a side
                   stream while the next one is being decoded (double buffering) and preprocessed there by
                   `preprocess_frames` / `preprocess_flow` (bit-exact with the reference loaders) straight into the
                   video's frame / flow tensors -- the inputs of `VideoScorer.score_video`.
* `gpu_jpeg=True`  decodes on the device instead (nvJPEG through torchvision.io.decode_jpeg): only the compressed bytes
                   cross PCIe.  nvJPEG's IDCT rounds differently from libjpeg-turbo (+-1..2 grey levels), so this mode is
                   NOT bit-exact with the reference loader and is reported separately.
"""
from __future__ import annotations

import glob
import os
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

FLO_MAGIC = 202021.25


def read_flo(path: str) -> np.ndarray:
    """.flo (Middlebury) -> float32 [h, w, 2]; the payload is returned as stored (the loader's scalings run on the GPU)."""
    with open(path, "rb") as f:
        head = np.fromfile(f, np.float32, count=1)
        if head.size != 1 or float(head[0]) != FLO_MAGIC:
            raise ValueError("%s: not a .flo file (magic %r)" % (path, head))
        w = int(np.fromfile(f, np.int32, count=1)[0])
        h = int(np.fromfile(f, np.int32, count=1)[0])
        if w <= 0 or h <= 0:
            raise ValueError("%s: bad .flo size %dx%d" % (path, w, h))
        data = np.fromfile(f, np.float32, count=2 * w * h)
        if data.size != 2 * w * h:
            raise ValueError("%s: truncated .flo payload (%d of %d values)" % (path, data.size, 2 * w * h))
    return data.reshape(h, w, 2)


def write_flo(path: str, flow: np.ndarray) -> None:
    """float32 [h, w, 2] -> .flo (for tests and synthetic datasets)."""
    flow = np.ascontiguousarray(flow, dtype=np.float32)
    h, w, c = flow.shape
    assert c == 2
    with open(path, "wb") as f:
        np.array([FLO_MAGIC], np.float32).tofile(f)
        np.array([w, h], np.int32).tofile(f)
        flow.tofile(f)


def decode_frame(path: str) -> np.ndarray:
    """Image file -> BGR uint8 [h, w, 3] (what cv2.imread / the reference's TurboJPEG wrapper return)."""
    import cv2
    img = cv2.imread(path, cv2.IMREAD_COLOR)
    if img is None:
        raise ValueError("%s: cannot decode image" % path)
    return img


def list_frames(folder: str) -> List[str]:
    """Sorted file list of one sub-video folder (reference two_stream_dataset.py `setup`: glob + sort)."""
    return sorted(glob.glob(os.path.join(folder, "*")))


class VideoLoader:
    """Decoded sub-videos on the device, produced with decode / upload / preprocess overlapped.

        loader = VideoLoader(device, size=(256, 256), chunk=64)
        frames, flows = loader.load(rgb_paths, flo_paths)          # [T,3,H,W], [T',2,H,W] float32 on `device`
        for name, frames, flows in loader.iter_dataset(rgb_root, op_root): ...   # next video prefetched meanwhile

    All frames of a video must share one native size (true for ped2 / avenue / shanghaitech)."""

    def __init__(self, device, size: Tuple[int, int] = (256, 256), chunk: int = 64, decode_threads: int = 8,
                 gpu_jpeg: bool = False):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ammc_b200: VideoLoader feeds the GPU preprocessing kernels; device must be CUDA")
        self.size = (int(size[0]), int(size[1]))
        self.chunk = int(chunk)
        self.gpu_jpeg = bool(gpu_jpeg)
        self.pool = ThreadPoolExecutor(max_workers=max(1, int(decode_threads)))      # per-file decodes
        self.chunk_pool = ThreadPoolExecutor(max_workers=1)                          # one chunk ahead (never nests in `pool`)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._pinned = {}            # (kind, slot, shape) -> pinned staging tensor

    # -- staging ----------------------------------------------------------------------------------
    def _staging(self, kind: str, slot: int, shape, dtype) -> torch.Tensor:
        key = (kind, slot)
        t = self._pinned.get(key)
        if t is None or tuple(t.shape[1:]) != tuple(shape[1:]) or t.shape[0] < shape[0] or t.dtype != dtype:
            t = torch.empty((max(shape[0], self.chunk),) + tuple(shape[1:]), dtype=dtype).pin_memory()
            self._pinned[key] = t
        return t[: shape[0]]

    def _decode_chunk_host(self, paths: Sequence[str], kind: str, slot: int) -> torch.Tensor:
        """Decode `paths` in the pool straight into a pinned staging buffer; returns the (pinned) chunk tensor."""
        first = decode_frame(paths[0]) if kind == "rgb" else read_flo(paths[0])
        buf = self._staging(kind, slot, (len(paths),) + first.shape, torch.uint8 if kind == "rgb" else torch.float32)
        view = buf.numpy()
        view[0] = first

        def work(i):
            arr = decode_frame(paths[i]) if kind == "rgb" else read_flo(paths[i])
            if arr.shape != first.shape:
                raise ValueError("%s: size %s differs from the video's %s" % (paths[i], arr.shape, first.shape))
            view[i] = arr
        list(self.pool.map(work, range(1, len(paths))))
        return buf

    def _decode_chunk_gpu(self, paths: Sequence[str]) -> torch.Tensor:
        """nvJPEG: compressed bytes -> BGR uint8 [n, h, w, 3] on the device (not bit-exact with libjpeg-turbo)."""
        import torchvision.io as tvio
        datas = list(self.pool.map(lambda p: torch.from_numpy(np.fromfile(p, np.uint8)), paths))
        imgs = tvio.decode_jpeg(datas, device=self.device, mode=tvio.ImageReadMode.RGB)
        return torch.stack(imgs).permute(0, 2, 3, 1).flip(-1).contiguous()

    # -- one video --------------------------------------------------------------------------------
    def _load_kind(self, paths: Sequence[str], kind: str, out: Optional[torch.Tensor]) -> torch.Tensor:
        from .preprocess import preprocess_flow, preprocess_frames
        W, H = self.size
        n = len(paths)
        ch = 3 if kind == "rgb" else 2
        out = torch.empty((n, ch, H, W), dtype=torch.float32, device=self.device) if out is None else out
        if n == 0:
            return out
        chunks = [(i, min(n, i + self.chunk)) for i in range(0, n, self.chunk)]
        cur = torch.cuda.current_stream(self.device)
        self.copy_stream.wait_stream(cur)
        done = [None, None]                                      # event per staging slot: its H2D copy has finished
        nxt = None
        if not (self.gpu_jpeg and kind == "rgb"):
            nxt = self.chunk_pool.submit(self._decode_chunk_host, paths[chunks[0][0]:chunks[0][1]], kind, 0)
        for ci, (a, b) in enumerate(chunks):
            slot = ci & 1
            if self.gpu_jpeg and kind == "rgb":
                with torch.cuda.stream(self.copy_stream):
                    dev_chunk = self._decode_chunk_gpu(paths[a:b])
            else:
                host = nxt.result()
                if ci + 1 < len(chunks):                         # decode the next chunk into the other slot meanwhile
                    na, nb = chunks[ci + 1]
                    if done[slot ^ 1] is not None:
                        done[slot ^ 1].synchronize()             # its previous contents must have left the host
                    nxt = self.chunk_pool.submit(self._decode_chunk_host, paths[na:nb], kind, slot ^ 1)
                with torch.cuda.stream(self.copy_stream):
                    dev_chunk = host.to(self.device, non_blocking=True)
                    done[slot] = torch.cuda.Event()
                    done[slot].record(self.copy_stream)
            with torch.cuda.stream(self.copy_stream):
                res = (preprocess_frames if kind == "rgb" else preprocess_flow)(dev_chunk, self.size)
                out[a:b].copy_(res, non_blocking=True)
                dev_chunk.record_stream(self.copy_stream)
        cur.wait_stream(self.copy_stream)
        out.record_stream(cur)
        return out

    def load(self, rgb_paths: Sequence[str], flo_paths: Sequence[str]) -> Tuple[torch.Tensor, torch.Tensor]:
        """One sub-video: (frames [T,3,H,W], flows [T',2,H,W]) on the device, ready for `VideoScorer.score_video`."""
        frames = self._load_kind(list(rgb_paths), "rgb", None)
        flows = self._load_kind(list(flo_paths), "op", None)
        return frames, flows

    # -- a dataset --------------------------------------------------------------------------------
    def iter_dataset(self, rgb_root: str, op_root: str, videos: Optional[Sequence[str]] = None
                     ) -> Iterator[Tuple[str, torch.Tensor, torch.Tensor]]:
        """Yield (video name, frames, flows) for every sub-folder of `rgb_root` (sorted, like the reference's
        `sorted(os.listdir(...))`, test_helper.py:405-408); the next video is loaded by a helper thread while the caller
        scores the current one."""
        names = sorted(os.listdir(rgb_root)) if videos is None else list(videos)
        names = [v for v in names if os.path.isdir(os.path.join(rgb_root, v))]
        if not names:
            return
        result = {}

        def fetch(i):
            with torch.cuda.device(self.device):
                try:
                    result[i] = self.load(list_frames(os.path.join(rgb_root, names[i])),
                                          list_frames(os.path.join(op_root, names[i])))
                    torch.cuda.current_stream(self.device).synchronize()
                except BaseException as e:       # re-raised in the consumer
                    result[i] = e

        th = threading.Thread(target=fetch, args=(0,))
        th.start()
        for i, name in enumerate(names):
            th.join()
            item = result.pop(i)
            if i + 1 < len(names):
                th = threading.Thread(target=fetch, args=(i + 1,))
                th.start()
            if isinstance(item, BaseException):
                raise item
            yield name, item[0], item[1]

    def close(self):
        self.chunk_pool.shutdown(wait=True)
        self.pool.shutdown(wait=True)
