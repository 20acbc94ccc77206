"""Packed feature shards + asynchronous loader.

The engine consumes PRECOMPUTED feature sequences (north_star).  The reference reads them one tiny file at a time on
the training thread -- one `np.load` per clip per step for WavLM (train.py:150-171; written by
create_wavlm_audio_feat.py:30-33 as `<video>/<clip>.npy`) -- which cannot feed an engine that consumes 13 k windows/s
(3.5 GB/s of bf16 features).  A shard packs whole windows contiguously:

    <name>.jmtshard = 4 KiB header (JSON, space padded) + visual (W, Cv, T) bf16 + audio (W, T, Ca) bf16
                      + labels_v (W, T) fp32 + labels_a (W, T) fp32          (C order, little endian)

so a batch of windows is ONE contiguous byte range per array: the loader memory-maps the shard, copies batch i+1 into
pinned staging buffers on a worker thread while batch i trains, and issues the host->device copies on a side stream
(double-buffered device tensors, CUDA events for hand-off) -- what bench.py's `e2e` leg times.
"""
from __future__ import annotations

import json
import threading
from typing import Iterator, Optional, Sequence, Tuple

import numpy as np
import torch

HEADER_BYTES = 4096
MAGIC = "jmtshard-v1"


def _bf16_bits(x: np.ndarray) -> np.ndarray:
    """float32 -> bf16 bit patterns (uint16), round-to-nearest-even like torch.Tensor.to(torch.bfloat16)."""
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)


def write_shard(path: str, visual: np.ndarray, audio: np.ndarray, labels_v: np.ndarray, labels_a: np.ndarray):
    """visual (W, Cv, T), audio (W, T, Ca) float32 (stored as bf16); labels (W, T) float32 (-5 = ignore sentinel)."""
    W, Cv, T = visual.shape
    assert audio.shape[:2] == (W, T) and labels_v.shape == (W, T) and labels_a.shape == (W, T)
    hdr = {"magic": MAGIC, "windows": int(W), "seq_len": int(T), "visual_dim": int(Cv), "audio_dim": int(audio.shape[2]),
           "visual_layout": "W,C,T", "audio_layout": "W,T,C", "feature_dtype": "bf16", "label_dtype": "f32"}
    raw = json.dumps(hdr).encode()
    assert len(raw) < HEADER_BYTES
    with open(path, "wb") as f:
        f.write(raw.ljust(HEADER_BYTES, b" "))
        f.write(_bf16_bits(visual).tobytes())
        f.write(_bf16_bits(audio).tobytes())
        f.write(np.ascontiguousarray(labels_v, dtype="<f4").tobytes())
        f.write(np.ascontiguousarray(labels_a, dtype="<f4").tobytes())


def read_clip_features(root: str, video: str, clip_ids: Sequence, prev: Optional[np.ndarray] = None):
    """One window of per-clip feature vectors from the reference's on-disk layout `<root>/<video>/<clip>.npy`
    (create_wavlm_audio_feat.py:30-33 writes one 1-D vector per clip; train.py:150-171 reads them back one `np.load` per clip
    per step).  Returns ((T, D) float32, last vector).  A missing file repeats the most recently loaded vector, which is what
    the reference's loop does: `feat_numpy` (train.py:157-159) survives across clips, windows AND batches, so `prev` carries
    the last vector of the previous window in; only a missing file before ANY clip was ever loaded fails (unbound
    variable in the reference, FileNotFoundError here)."""
    import os
    rows = []
    for c in clip_ids:
        f = os.path.join(root, str(video), f"{c}.npy")
        if os.path.exists(f):
            prev = np.load(f).astype(np.float32, copy=False).reshape(-1)
        elif prev is None:
            raise FileNotFoundError(f"{f}: no feature file and no previously loaded clip to repeat "
                                    "(the reference raises UnboundLocalError here, train.py:157-171)")
        rows.append(prev)
    return np.stack(rows, axis=0), prev


def pack_npy_tree(path: str, audio_root: str, windows: Sequence[Tuple[str, Sequence]], visual: np.ndarray,
                  labels_v: np.ndarray, labels_a: np.ndarray):
    """Pack W windows into one shard: window i takes its audio from the per-clip tree (`windows[i] = (video, clip_ids)`,
    see read_clip_features; the last loaded vector carries over from window to window as in train.py:150-171) and its visual
    features / labels from row i of the given arrays (visual (W, Cv, T) as the TCN consumes it, labels (W, T) with -5 = ignore)."""
    mats, prev = [], None
    for v, ids in windows:
        m, prev = read_clip_features(audio_root, v, ids, prev)
        mats.append(m)
    audio = np.stack(mats, axis=0)
    if audio.shape[:2] != (visual.shape[0], visual.shape[2]):
        raise ValueError(f"audio windows {audio.shape[:2]} do not match visual (W, T) = {(visual.shape[0], visual.shape[2])}")
    write_shard(path, visual, audio, labels_v, labels_a)


class Shard:
    """Memory-mapped view of one shard: .visual / .audio are uint16 (bf16 bit patterns), .labels_v / .labels_a float32."""

    def __init__(self, path: str):
        with open(path, "rb") as f:
            hdr = json.loads(f.read(HEADER_BYTES).decode().strip())
        if hdr.get("magic") != MAGIC:
            raise ValueError(f"{path}: not a {MAGIC} file")
        self.header = hdr
        W, T, Cv, Ca = hdr["windows"], hdr["seq_len"], hdr["visual_dim"], hdr["audio_dim"]
        off = HEADER_BYTES
        self.visual = np.memmap(path, dtype=np.uint16, mode="r", offset=off, shape=(W, Cv, T)); off += W * Cv * T * 2
        self.audio = np.memmap(path, dtype=np.uint16, mode="r", offset=off, shape=(W, T, Ca)); off += W * T * Ca * 2
        self.labels_v = np.memmap(path, dtype="<f4", mode="r", offset=off, shape=(W, T)); off += W * T * 4
        self.labels_a = np.memmap(path, dtype="<f4", mode="r", offset=off, shape=(W, T))
        self.windows = W


class FeatureShardLoader:
    """Iterates (audio (B,T,Ca) bf16, visual (B,Cv,T) bf16, labels_v (B,T) f32, labels_a (B,T) f32) CUDA tensors over a list
    of shards; rank r of `world` takes windows [r*W/world, (r+1)*W/world) of every shard (SURVEY 8e batch sharding).
    The tensors of batch i are valid until batch i+2 is requested (two device buffer sets)."""

    def __init__(self, paths: Sequence[str], batch: int, device=None, rank: int = 0, world: int = 1, drop_last: bool = True):
        self.shards = [Shard(p) for p in paths]
        self.batch, self.drop_last = batch, drop_last
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.rank, self.world = rank, world
        h = self.shards[0].header
        T, Cv, Ca = h["seq_len"], h["visual_dim"], h["audio_dim"]
        mk = lambda shape, dt: torch.empty(shape, dtype=dt)          # noqa: E731
        self.host = [(mk((batch, T, Ca), torch.bfloat16).pin_memory(), mk((batch, Cv, T), torch.bfloat16).pin_memory(),
                      mk((batch, T), torch.float32).pin_memory(), mk((batch, T), torch.float32).pin_memory()) for _ in range(2)]
        self.dev = [tuple(torch.empty_like(t, device=self.device) for t in hs) for hs in self.host]
        self.stream = torch.cuda.Stream(device=self.device)
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.copied = [torch.cuda.Event() for _ in range(2)]      # H2D out of the pinned set has completed
        self.consumed = [torch.cuda.Event() for _ in range(2)]    # the consumer's work queued so far (reads of dev[slot]) is done

    def _plan(self):
        for s in self.shards:
            lo, hi = (self.rank * s.windows) // self.world, ((self.rank + 1) * s.windows) // self.world
            for w0 in range(lo, hi, self.batch):
                n = min(self.batch, hi - w0)
                if n < self.batch and self.drop_last:
                    break
                yield s, w0, n

    def _stage(self, slot: int, s: Shard, w0: int, n: int):
        """worker thread: shard (page cache / disk) -> pinned host buffers of `slot`."""
        a, v, lv, la = self.host[slot]
        a.view(torch.int16).numpy().view(np.uint16)[:n] = s.audio[w0:w0 + n]
        v.view(torch.int16).numpy().view(np.uint16)[:n] = s.visual[w0:w0 + n]
        lv.numpy()[:n] = s.labels_v[w0:w0 + n]
        la.numpy()[:n] = s.labels_a[w0:w0 + n]

    def _stage_guarded(self, box: dict, slot: int, s: Shard, w0: int, n: int):
        try:
            self._stage(slot, s, w0, n)
        except BaseException as e:       # re-raised on the consumer thread after join(): never yield stale pinned data
            box["error"] = e

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]]:
        plan = list(self._plan())
        worker: Optional[threading.Thread] = None
        box: dict = {}
        if plan:
            worker = threading.Thread(target=self._stage_guarded, args=(box, 0, *plan[0]))
            worker.start()
        for i, (s, w0, n) in enumerate(plan):
            slot = i % 2
            worker.join()                                            # batch i is in the pinned set `slot`
            if "error" in box:
                raise RuntimeError("FeatureShardLoader: staging a batch failed") from box["error"]
            cur = torch.cuda.current_stream(self.device)
            # write-after-read on the device buffers: dev[slot] was handed out as batch i-2 and everything the consumer
            # queued on its stream up to now (that step's kernels; the host may run several steps ahead of the GPU) must
            # have finished reading it before the copy engine overwrites it
            self.consumed[slot].record(cur)
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(self.consumed[slot])
                for d, h in zip(self.dev[slot], self.host[slot]):
                    d[:n].copy_(h[:n], non_blocking=True)
                self.copied[slot].record(self.stream)
                self.ready[slot].record(self.stream)
            if i + 1 < len(plan):                                    # stage batch i+1 while batch i copies / trains
                nslot = (i + 1) % 2
                self.copied[nslot].synchronize()                     # its previous H2D (batch i-1) has drained
                worker = threading.Thread(target=self._stage_guarded, args=(box, nslot, *plan[i + 1]))
                worker.start()
            cur.wait_event(self.ready[slot])
            yield tuple(d[:n] for d in self.dev[slot])
