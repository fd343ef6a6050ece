"""Packed feature-map shards: the data format either side of the BMU path.

The reference stores one ``.npy`` per feature map, at most 1000 per folder, indexed by a TinyDB
JSON file (``generate_fmap_dataset.py:42-72``), and reads them back one ``np.load`` at a time in
DataLoader workers (``dataset_loader/feature_map_dataset.py:22-42``).  At 1e8-1e9 patches/s that
reader, not the GPU, bounds tokenisation and pruning.  This module packs the same data into a few
large, memory-mappable shards and streams them into pinned batches:

    shard file  = 64-byte header | n * C*H*W float32, C-contiguous (the reference's dtype and layout)
    header      = b"SOMFMAP1" | u32 version | u32 C | u32 H | u32 W | u64 n | zero padding to 64 bytes

* ``convert_reference_dataset`` reads the reference layout (TinyDB json -> ``fmap_path`` entries, in
  document order) and writes shards; ``image_path`` is kept in a side ``index.json``.
* ``ShardReader`` memory-maps shards and yields ``(lo, hi, pinned_tensor_view)`` batches that feed
  ``HostTokenizer.tokenize`` / ``bmu_histogram`` directly.  Pure host code: numpy + torch pinned memory.
"""
import json
import os
import struct

import numpy as np
import torch

MAGIC = b"SOMFMAP1"
HEADER_BYTES = 64
VERSION = 1


def _header(c, h, w, n):
    head = MAGIC + struct.pack("<IIIIQ", VERSION, c, h, w, n)
    return head + b"\0" * (HEADER_BYTES - len(head))


def read_header(path):
    with open(path, "rb") as f:
        head = f.read(HEADER_BYTES)
    if len(head) != HEADER_BYTES or head[:8] != MAGIC:
        raise ValueError(f"{path}: not a feature-map shard")
    version, c, h, w, n = struct.unpack("<IIIIQ", head[8:32])
    if version != VERSION:
        raise ValueError(f"{path}: shard version {version}, expected {VERSION}")
    expect = HEADER_BYTES + n * c * h * w * 4
    if os.path.getsize(path) != expect:
        raise ValueError(f"{path}: size {os.path.getsize(path)} != {expect} implied by the header")
    return c, h, w, n


def write_shard(path, fmaps):
    """fmaps: (n, C, H, W) float32 array-like.  Returns n."""
    arr = np.ascontiguousarray(np.asarray(fmaps), dtype=np.float32)
    if arr.ndim != 4:
        raise ValueError("write_shard expects (n, C, H, W)")
    n, c, h, w = arr.shape
    with open(path, "wb") as f:
        f.write(_header(c, h, w, n))
        f.write(arr.tobytes(order="C"))
    return n


def reference_entries(db_json_path):
    """The documents of the reference's TinyDB file in document-id order (generate_fmap_dataset.py
    :62-72 inserts them in file_index order).  TinyDB's on-disk form is
    {"_default": {"1": {...}, "2": {...}}}; no tinydb import is needed to read it."""
    with open(db_json_path, "r") as f:
        db = json.load(f)
    table = db.get("_default", db)
    if not isinstance(table, dict) or len(table) == 0:
        raise Exception("No data found.")          # same message as feature_map_dataset.py:29
    return [table[k] for k in sorted(table, key=lambda s: int(s))]


def convert_reference_dataset(db_json_path, out_dir, fmaps_per_shard=65536, path_root=None):
    """Pack the reference's per-file dataset into shards.  ``path_root`` re-bases relative
    ``fmap_path`` entries.  Returns the list of shard paths; writes ``index.json`` beside them."""
    entries = reference_entries(db_json_path)
    os.makedirs(out_dir, exist_ok=True)
    shards, index, batch = [], [], []

    def flush():
        if not batch:
            return
        path = os.path.join(out_dir, f"fmaps_{len(shards):05d}.shard")
        write_shard(path, np.stack(batch))
        shards.append(path)
        batch.clear()

    for i, doc in enumerate(entries):
        p = doc["fmap_path"]
        if path_root is not None and not os.path.isabs(p):
            p = os.path.join(path_root, p)
        with open(p, "rb") as f:
            fmap = np.load(f, allow_pickle=False)           # as feature_map_dataset.py:38-39
        batch.append(np.asarray(fmap, dtype=np.float32))    # .float() of :42
        index.append({"i": i, "shard": len(shards), "row": len(batch) - 1,
                      "image_path": doc.get("image_path")})
        if len(batch) == fmaps_per_shard:
            flush()
    flush()
    with open(os.path.join(out_dir, "index.json"), "w") as f:
        json.dump({"shards": [os.path.basename(s) for s in shards], "entries": index}, f)
    return shards


class ShardReader:
    """Sequential reader over shards: memory-mapped, double-buffered pinned staging."""

    def __init__(self, shard_paths, batch_fmaps=4096, pin=True, depth=2):
        if isinstance(shard_paths, str):
            shard_paths = [shard_paths]
        if len(shard_paths) == 0:
            raise Exception("No data found.")
        self.paths = list(shard_paths)
        self.shape = None
        self.counts = []
        for p in self.paths:
            c, h, w, n = read_header(p)
            if self.shape is None:
                self.shape = (c, h, w)
            elif self.shape != (c, h, w):
                raise ValueError(f"{p}: shape {(c, h, w)} differs from {self.shape}")
            self.counts.append(n)
        self.batch = int(batch_fmaps)
        self.pin = bool(pin) and torch.cuda.is_available()
        self.depth = int(depth)
        self._bufs = None
        self._busy = {}               # staging slot -> CUDA event of the last asynchronous consumer
        self._last_slot = None

    def mark_in_flight(self, event):
        """Tell the reader that the batch yielded LAST is still being read asynchronously (e.g. by a non-blocking
        host-to-device copy); ``event`` is a ``torch.cuda.Event`` recorded after that work.  The staging slot is not
        refilled before the event has completed.  Consumers that finish with a batch before asking for the next one
        (``HostTokenizer.tokenize(..., sync=True)``) need not call this."""
        if self._last_slot is not None:
            self._busy[self._last_slot] = event

    def __len__(self):
        return sum(self.counts)

    def _map(self, i):
        c, h, w = self.shape
        return np.memmap(self.paths[i], dtype=np.float32, mode="r", offset=HEADER_BYTES,
                         shape=(self.counts[i], c, h, w))

    def read_all(self):
        """All feature maps as one (N, C, H, W) float32 tensor (pinned when CUDA is present)."""
        c, h, w = self.shape
        out = torch.empty(len(self), c, h, w, dtype=torch.float32, pin_memory=self.pin)
        lo = 0
        for i, n in enumerate(self.counts):
            out[lo:lo + n].numpy()[...] = self._map(i)
            lo += n
        return out

    def batches(self):
        """Yields (lo, hi, tensor) with tensor a (hi - lo, C, H, W) float32 view of a staging buffer
        that stays valid until ``depth`` further batches have been produced."""
        c, h, w = self.shape
        if self._bufs is None:
            self._bufs = [torch.empty(self.batch, c, h, w, dtype=torch.float32, pin_memory=self.pin)
                          for _ in range(self.depth)]
        k, lo = 0, 0
        for i, n in enumerate(self.counts):
            mm = self._map(i)
            for a in range(0, n, self.batch):
                b = min(n, a + self.batch)
                slot = k % self.depth
                ev = self._busy.pop(slot, None)
                if ev is not None:
                    ev.synchronize()              # an earlier DMA out of this pinned slot is still pending
                buf = self._bufs[slot]
                buf[:b - a].numpy()[...] = mm[a:b]
                self._last_slot = slot
                yield lo + a, lo + b, buf[:b - a]
                k += 1
            lo += n
