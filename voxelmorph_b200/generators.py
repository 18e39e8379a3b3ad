"""Host-side data feed for the training loop ("next" row N1 of SURVEY.md section 8(f)).

Same generator names, arguments and yield contract as the reference (voxelmorph/generators.py:9-143: `volgen` yields a
tuple of channel-last numpy batches, `scan_to_scan` / `scan_to_atlas` yield `(invols, outvols)` lists), and the same
sequence of draws from `np.random`, so `scripts/torch/train.py:200-201` consumes it unchanged.  What differs is the cost
per step, which bounds the unmodified loop once a GPU step takes milliseconds:

* every file is decoded ONCE (`VolumeCache`): the reference re-opens and re-inflates the `.npz` on every draw
  (py/utils.py:69-129 -> ~0.19 s per 160x192x224 volume);
* volumes are kept as float32, C-contiguous, feature axis included, in page-locked memory when CUDA is present: the
  `.float()` of train.py:200 is a no-op and `.to(device)` is one DMA from pinned memory instead of a pageable staged copy
  preceded by a float64 -> float32 pass;
* a batch of one is a zero-copy view (no `np.concatenate`);
* the all-zero "target flow" that `Grad` ignores (generators.py:96-99) is float32 instead of float64: half the bytes
  for train.py:201 to move, and no cast kernel;
* `Prefetcher` runs any generator one or more batches ahead on a background thread.

Nothing here touches the GPU; it is covered by CPU tests against the reference generators (tests/test_generators.py).
"""
import glob
import os
import queue
import threading
from collections import OrderedDict

import numpy as np

__all__ = ["VolumeCache", "load_volfile", "volgen", "scan_to_scan", "scan_to_atlas", "semisupervised", "Prefetcher"]


def _cuda_ready():
    """True once THIS PROCESS already has a CUDA context.  The data feed must never create one itself: the reference's
    train.py draws its first batch (scripts/torch/train.py:113) BEFORE it sets CUDA_VISIBLE_DEVICES (train.py:125), and the CUDA
    runtime reads that variable when it initialises — a feed that page-locks memory at the first draw (or merely asks
    torch.cuda.is_available()) pins every rank of a multi-GPU run to device 0."""
    try:
        import torch
        return torch.cuda.is_initialized()
    except Exception:  # noqa: BLE001
        return False


def _pinned_empty(shape, dtype):
    """Page-locked numpy array when this process already uses CUDA, plain numpy otherwise.  Returns (array, owner)."""
    try:
        if _cuda_ready():
            import torch
            t = torch.empty(tuple(shape), dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
            return t.numpy(), t
    except Exception:  # noqa: BLE001 - pinning is an optimisation, never a requirement
        pass
    return np.empty(shape, dtype=dtype), None


def _pad_centered(vol, shape):
    """Zero-pad to `shape`, content centred with floor((target - size) / 2) leading zeros (py/utils.py:235-247)."""
    if vol.shape == tuple(shape):
        return vol
    if len(shape) != vol.ndim or any(p < v for p, v in zip(shape, vol.shape)):
        raise ValueError("pad_shape %r cannot hold a volume of shape %r" % (tuple(shape), vol.shape))
    out = np.zeros(shape, dtype=vol.dtype)
    lead = [int((p - v) / 2) for p, v in zip(shape, vol.shape)]
    out[tuple(slice(o, o + n) for o, n in zip(lead, vol.shape))] = vol
    return out


def _decode(src, np_var):
    """File name or preloaded array -> numpy array (py/utils.py:92-117).  Preloaded arrays are passed through."""
    if isinstance(src, os.PathLike):
        src = os.fspath(src)
    if not isinstance(src, str):
        return np.asarray(src)
    if not os.path.isfile(src):
        raise ValueError("'%s' is not a file." % src)
    if src.endswith((".nii", ".nii.gz", ".mgz")):
        try:
            import nibabel as nib
        except ImportError as e:
            raise ValueError("loading %s needs nibabel, which is not installed" % src) from e
        return np.squeeze(nib.load(src).dataobj)
    if src.endswith(".npy"):
        return np.load(src)
    if src.endswith(".npz"):
        with np.load(src) as z:
            keys = list(z.keys())
            return z[keys[0]] if len(keys) == 1 else z[np_var]
    raise ValueError("unknown filetype for %s" % src)


def _resize_nearest(vol, factor):
    """Nearest-neighbour zoom of every axis but the trailing feature axis (py/utils.py:250-262, scipy zoom order 0)."""
    if factor == 1:
        return vol
    from scipy import ndimage
    return ndimage.zoom(vol, [factor] * (vol.ndim - 1) + [1], order=0)


class VolumeCache:
    """Decode-once store: (source, np_var, pad_shape, resize_factor, add_feat_axis) -> ready-to-batch array.

    Floating-point volumes are stored as float32 (what train.py:200 converts to anyway); integer volumes (label maps)
    keep their dtype.  `max_bytes` bounds the cache (least recently used entries are dropped); `pin=None` page-locks the
    volumes once the process has a CUDA context (never creating one: see _cuda_ready).
    """

    def __init__(self, max_bytes=None, pin=None):
        self.max_bytes = max_bytes
        self.pin = pin
        self._items = OrderedDict()
        self._owners = {}
        self._bytes = 0
        self._lock = threading.Lock()
        self.hits = self.misses = 0

    def get(self, src, np_var="vol", pad_shape=None, resize_factor=1, add_feat_axis=True):
        if not isinstance(src, (str, os.PathLike)):
            # preloaded arrays are never cached (an id()-keyed entry could outlive its array and be served to a later
            # array that reuses the address, or go stale when the array is modified in place): like the reference
            # (py/utils.py:95-99) they are passed through, with the same padding / axis handling
            vol = np.asarray(src)
            if pad_shape:
                vol = _pad_centered(vol, pad_shape)
            if add_feat_axis:
                vol = vol[..., np.newaxis]
            return _resize_nearest(vol, resize_factor)
        key = (os.fspath(src), np_var, None if pad_shape is None else tuple(pad_shape), resize_factor, bool(add_feat_axis))
        with self._lock:
            hit = self._items.get(key)
            if hit is not None:
                self._items.move_to_end(key)
                self.hits += 1
                if (self.pin is not False and key not in self._owners and hit.dtype == np.float32 and _cuda_ready()):
                    # decoded before the process had a CUDA context (see _cuda_ready): page-lock it now, once
                    buf, owner = _pinned_empty(hit.shape, hit.dtype)
                    if owner is not None:
                        np.copyto(buf, hit)
                        buf.setflags(write=False)
                        self._items[key] = buf
                        self._owners[key] = owner
                        hit = buf
                return hit
        vol = _decode(src, np_var)
        if pad_shape:
            vol = _pad_centered(vol, pad_shape)
        if add_feat_axis:
            vol = vol[..., np.newaxis]
        vol = _resize_nearest(vol, resize_factor)
        dtype = np.float32 if np.issubdtype(vol.dtype, np.floating) else vol.dtype
        owner = None
        if self.pin is not False and dtype == np.float32:
            buf, owner = _pinned_empty(vol.shape, dtype)
            np.copyto(buf, vol, casting="same_kind")
            vol = buf
        else:
            vol = np.ascontiguousarray(vol, dtype=dtype)
        vol.setflags(write=False)
        with self._lock:
            self.misses += 1
            self._items[key] = vol
            if owner is not None:
                self._owners[key] = owner
            self._bytes += vol.nbytes
            while self.max_bytes is not None and self._bytes > self.max_bytes and len(self._items) > 1:
                k, v = self._items.popitem(last=False)
                self._owners.pop(k, None)
                self._bytes -= v.nbytes
        return vol

    def __len__(self):
        return len(self._items)


def _default_cache_bytes():
    """Bound of the process-wide cache: VXM_B200_CACHE_GB, else a quarter of the host's RAM, at most 16 GiB
    (page-locking more than that can destabilise the host)."""
    env = os.environ.get("VXM_B200_CACHE_GB")
    if env:
        return int(float(env) * (1 << 30))
    try:
        ram = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES")
    except (ValueError, OSError, AttributeError):
        ram = 32 << 30
    return int(min(16 << 30, ram // 4))


_default_cache = VolumeCache(max_bytes=_default_cache_bytes())


def load_volfile(filename, np_var="vol", add_batch_axis=False, add_feat_axis=False, pad_shape=None, resize_factor=1, cache=None):
    """Reference-compatible loader (py/utils.py:69-129 without `ret_affine`) that decodes each file once."""
    vol = (_default_cache if cache is None else cache).get(filename, np_var, pad_shape, resize_factor, add_feat_axis)
    return vol[np.newaxis, ...] if add_batch_axis else vol


def _expand_names(vol_names):
    if isinstance(vol_names, str):
        if os.path.isdir(vol_names):
            vol_names = os.path.join(vol_names, "*")
        vol_names = glob.glob(vol_names)
    return vol_names


def _batch(items):
    """Stack volumes along a new leading axis; one volume is returned as a view."""
    if len(items) == 1:
        return items[0][np.newaxis, ...]
    out, owner = _pinned_empty((len(items),) + items[0].shape, items[0].dtype) if items[0].dtype == np.float32 else \
        (np.empty((len(items),) + items[0].shape, items[0].dtype), None)
    for i, v in enumerate(items):
        out[i] = v
    if owner is not None:
        out = _Owned(out, owner)
    return out


class _Owned(np.ndarray):
    """ndarray view that keeps the page-locked torch storage it aliases alive."""

    def __new__(cls, arr, owner):
        obj = arr.view(cls)
        obj._owner = owner
        return obj

    def __array_finalize__(self, obj):
        self._owner = getattr(obj, "_owner", None)


def volgen(vol_names, batch_size=1, segs=None, np_var="vol", pad_shape=None, resize_factor=1, add_feat_axis=True, cache=None):
    """Random volume batches; arguments and yields as reference generators.py:9-68 (`cache`: a VolumeCache to share)."""
    vol_names = _expand_names(vol_names)
    if isinstance(segs, list) and len(segs) != len(vol_names):
        raise ValueError("Number of image files must match number of seg files.")
    cache = _default_cache if cache is None else cache
    opts = dict(pad_shape=pad_shape, resize_factor=resize_factor, add_feat_axis=add_feat_axis)
    while True:
        indices = np.random.randint(len(vol_names), size=batch_size)   # same draw as generators.py:50
        vols = [_batch([cache.get(vol_names[i], np_var, **opts) for i in indices])]
        if segs is True:        # npz files carrying a 'seg' variable next to 'vol'
            vols.append(_batch([cache.get(vol_names[i], "seg", **opts) for i in indices]))
        elif isinstance(segs, list):
            vols.append(_batch([cache.get(segs[i], np_var, **opts) for i in indices]))
        yield tuple(vols)


def _zero_flow(batch_size, vol_shape, dtype):
    z = np.zeros((batch_size,) + tuple(vol_shape) + (len(vol_shape),), dtype=dtype)
    z.setflags(write=False)
    return z


def scan_to_scan(vol_names, bidir=False, batch_size=1, prob_same=0, no_warp=False, zeros_dtype=np.float32, **kwargs):
    """Scan-to-scan pairs; arguments, yields and random draws as reference generators.py:71-107."""
    zeros = None
    gen = volgen(vol_names, batch_size=batch_size, **kwargs)
    while True:
        scan1 = next(gen)[0]
        scan2 = next(gen)[0]
        if prob_same > 0 and np.random.rand() < prob_same:
            if np.random.rand() > 0.5:
                scan1 = scan2
            else:
                scan2 = scan1
        if not no_warp and zeros is None:
            zeros = _zero_flow(batch_size, scan1.shape[1:-1], zeros_dtype)
        invols = [scan1, scan2]
        outvols = [scan2, scan1] if bidir else [scan2]
        if not no_warp:
            outvols.append(zeros)
        yield (invols, outvols)


def scan_to_atlas(vol_names, atlas, bidir=False, batch_size=1, no_warp=False, segs=None, zeros_dtype=np.float32, **kwargs):
    """Scan-to-atlas pairs; arguments and yields as reference generators.py:110-143 (`atlas`: (1, *vol, 1) array)."""
    atlas = np.asarray(atlas)
    zeros = _zero_flow(batch_size, atlas.shape[1:-1], zeros_dtype)
    atlas = np.repeat(atlas.astype(np.float32, copy=False), batch_size, axis=0)
    gen = volgen(vol_names, batch_size=batch_size, segs=segs, **kwargs)
    while True:
        res = next(gen)
        scan = res[0]
        invols = [scan, atlas]
        if not segs:
            outvols = [atlas, scan] if bidir else [atlas]
        else:
            outvols = [res[1], scan] if bidir else [res[1]]
        if not no_warp:
            outvols.append(zeros)
        yield (invols, outvols)


def semisupervised(vol_names, seg_names, labels, atlas_file=None, downsize=2, prob_dtype=np.float32, zeros_dtype=np.float32, cache=None):
    """Semi-supervised pairs ("next" row N2; reference generators.py:146-194): yields
    ([src_vol, trg_vol, src_prob_seg], [trg_vol, zeros, trg_prob_seg]) with the label maps turned into one-hot
    probability maps over `labels` and sub-sampled by `downsize` per axis.  Batch size is 1, 3-D only, like the
    reference; the draws from np.random follow its order (one volgen draw for the source, one for the target unless an
    atlas file supplies it).  The one-hot is built on the sub-sampled label map in one comparison against the label
    vector (float32 instead of float64: a 30-label map at 80x96x112 is 103 MB rather than 206 MB per segmentation)."""
    labels = np.asarray(labels)
    cache = _default_cache if cache is None else cache
    gen = volgen(vol_names, segs=seg_names, np_var="vol", cache=cache)

    def prob_seg(seg):
        if seg.shape[0] != 1 or seg.ndim != 5:
            raise ValueError("semisupervised: segmentations must be (1, D, H, W, 1) label maps (batch size 1, 3-D)")
        sub = seg[0, ::downsize, ::downsize, ::downsize, 0]
        return (sub[..., np.newaxis] == labels).astype(prob_dtype)[np.newaxis]

    trg_vol = trg_seg = None
    if atlas_file:
        trg_vol = cache.get(atlas_file, "vol", add_feat_axis=True)[np.newaxis]
        trg_seg = prob_seg(cache.get(atlas_file, "seg", add_feat_axis=True)[np.newaxis])
    zeros = None
    while True:
        src_vol, src_lab = next(gen)
        src_seg = prob_seg(src_lab)
        if not atlas_file:
            trg_vol, trg_lab = next(gen)
            trg_seg = prob_seg(trg_lab)
        if zeros is None:
            zeros = _zero_flow(1, src_vol.shape[1:-1], zeros_dtype)
        yield ([src_vol, trg_vol, src_seg], [trg_vol, zeros, trg_seg])


class Prefetcher:
    """Iterate `gen` on a background thread, `depth` items ahead (decode / batch assembly overlaps the GPU step).

    The random draws still happen in generator order, on the worker thread.  Exceptions raised by the generator are
    re-raised by `next()`; `close()` stops the worker.
    """

    _END = object()

    def __init__(self, gen, depth=2):
        self._q = queue.Queue(maxsize=max(1, depth))
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, args=(gen,), daemon=True)
        self._thread.start()

    def _run(self, gen):
        try:
            for item in gen:
                while not self._stop.is_set():
                    try:
                        self._q.put(item, timeout=0.1)
                        break
                    except queue.Full:
                        continue
                if self._stop.is_set():
                    return
            self._q.put(self._END)
        except BaseException as e:  # noqa: BLE001 - handed to the consumer
            self._q.put(e)

    def __iter__(self):
        return self

    def __next__(self):
        item = self._q.get()
        if item is self._END:
            raise StopIteration
        if isinstance(item, BaseException):
            raise item
        return item

    def close(self):
        self._stop.set()
