"""Minimal GeoTIFF reader / writer for the rasters on the path (NumPy only).

The reference reads and writes its rasters through rasterio/GDAL
(``TreeDetection/prediction.py:61``, ``postprocessing.py:781-800``, ``merging.py:56-75``);
neither is available here.  Supported: classic (non-Big) TIFF, little or big endian,
strips or tiles, chunky or planar configuration, uint8 / uint16 / int16 / float32 samples,
no compression, deflate (zlib) or LZW (libtreedet's td_tiff_lzw_decode, strips / tiles decoded on a
thread pool), predictors 1, 2 (horizontal differencing) and 3 (floating point), the GeoTIFF tags
ModelPixelScale (33550),
ModelTiepoint (33922), ModelTransformation (34264), GeoKeyDirectory (34735, EPSG code of the
projected / geographic CRS) and GDAL_NODATA (42113).  That covers the bundled
``data/nDSM/324125317.tif`` and everything this package writes.
"""
from __future__ import annotations

import os
import struct
import zlib
from dataclasses import dataclass

import numpy as np

_TYPES = {1: ("B", 1), 2: ("c", 1), 3: ("H", 2), 4: ("I", 4), 5: ("II", 8), 6: ("b", 1), 7: ("B", 1), 8: ("h", 2),
          9: ("i", 4), 10: ("ii", 8), 11: ("f", 4), 12: ("d", 8), 16: ("Q", 8), 17: ("q", 8), 18: ("Q", 8)}


def _tiff_kind(buf):
    """(byte order, is BigTIFF) of a TIFF header, or None: classic TIFF (magic 42, 32-bit offsets) or BigTIFF
    (magic 43, 64-bit offsets -- what GDAL writes for rasters beyond 4 GB, e.g. a merged county mosaic)."""
    if len(buf) < 16 or buf[:2] not in (b"II", b"MM"):
        return None
    bo = "<" if buf[:2] == b"II" else ">"
    magic = struct.unpack(bo + "H", buf[2:4])[0]
    if magic == 42:
        return bo, False
    if magic == 43 and struct.unpack(bo + "HH", buf[4:8]) == (8, 0):
        return bo, True
    return None


@dataclass
class GeoInfo:
    width: int
    height: int
    count: int
    dtype: np.dtype
    transform: tuple            # (a, b, c, d, e, f)
    epsg: int | None
    nodata: float | None

    @property
    def bounds(self):
        a, b, c, d, e, f = self.transform
        return (c, f + e * self.height, c + a * self.width, f)


def _read_ifd(buf, bo):
    big = struct.unpack(bo + "H", buf[2:4])[0] == 43
    if big:      # BigTIFF: 64-bit first-IFD offset, entry count, value counts and value offsets; 20-byte entries
        off = struct.unpack(bo + "Q", buf[8:16])[0]
        n = struct.unpack(bo + "Q", buf[off:off + 8])[0]
        first, esize, inline, cfmt, ofmt = off + 8, 20, 8, "Q", "Q"
    else:
        off = struct.unpack(bo + "I", buf[4:8])[0]
        n = struct.unpack(bo + "H", buf[off:off + 2])[0]
        first, esize, inline, cfmt, ofmt = off + 2, 12, 4, "I", "I"
    tags = {}
    for i in range(n):
        e = buf[first + esize * i: first + esize * (i + 1)]
        tag, typ, cnt = struct.unpack(bo + "HH" + cfmt, e[:esize - inline])
        fmt, size = _TYPES.get(typ, ("B", 1))
        total = size * cnt
        if total <= inline:
            raw = e[esize - inline:esize - inline + total]
        else:
            p = struct.unpack(bo + ofmt, e[esize - inline:esize])[0]
            raw = buf[p:p + total]
        if typ == 2:
            val = raw.split(b"\x00")[0].decode("latin1")
        elif typ in (5, 10):
            v = struct.unpack(bo + fmt[0] * (2 * cnt), raw)
            val = tuple(v[2 * k] / v[2 * k + 1] if v[2 * k + 1] else 0.0 for k in range(cnt))
        else:
            val = struct.unpack(bo + fmt * cnt, raw)
        tags[tag] = val
    return tags


def _info_from_tags(t):
    width, height = int(t[256][0]), int(t[257][0])
    spp = int(t.get(277, (1,))[0])
    bits = int(t.get(258, (8,))[0])
    fmt = int(t.get(339, (1,))[0])
    dtype = {(8, 1): np.uint8, (16, 1): np.uint16, (16, 2): np.int16, (32, 3): np.float32, (32, 1): np.uint32,
             (32, 2): np.int32, (64, 3): np.float64}.get((bits, fmt))
    if dtype is None:
        raise ValueError(f"unsupported sample format: {bits} bits, format {fmt}")
    if 34264 in t:
        m = t[34264]
        transform = (m[0], m[1], m[3], m[4], m[5], m[7])
    elif 33550 in t and 33922 in t:
        sx, sy = t[33550][0], t[33550][1]
        i, j, _, x, y, _ = t[33922][:6]
        transform = (sx, 0.0, x - i * sx, 0.0, -sy, y + j * sy)
    else:
        transform = (1.0, 0.0, 0.0, 0.0, 1.0, 0.0)
    epsg = None
    if 34735 in t:
        g = t[34735]
        for k in range(1, g[3] + 1):
            key, loc, cnt, val = g[4 * k: 4 * k + 4]
            if key in (3072, 2048) and loc == 0 and (epsg is None or key == 3072):
                epsg = int(val)
    nodata = None
    if 42113 in t:
        try:
            nodata = float(str(t[42113]).strip())
        except ValueError:
            nodata = None
    return GeoInfo(width, height, spp, np.dtype(dtype), tuple(float(v) for v in transform), epsg, nodata)


def _lzw(raw: bytes, cap: int) -> bytes:
    """One LZW strip / tile through libtreedet (host function td_tiff_lzw_decode)."""
    import ctypes as C
    from . import _lib
    dst = C.create_string_buffer(max(cap, 1))
    n = _lib.lib().td_tiff_lzw_decode(raw, len(raw), dst, cap)
    if n < 0:
        _lib.check(int(n), "td_tiff_lzw_decode")
    return dst.raw[:n]


def read_info(path) -> GeoInfo:
    import mmap
    with open(path, "rb") as f:
        buf = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)   # only the pages of the header / IFD are touched
    try:
        kind = _tiff_kind(buf)
        if kind is None:
            raise ValueError("not a TIFF / BigTIFF file")
        return _info_from_tags(_read_ifd(buf, kind[0]))
    finally:
        buf.close()


def read(path, window=None, out=None):
    """Returns (array (bands, h, w), GeoInfo).  window = (col_off, row_off, w, h): only the strips / tiles
    that intersect it are decoded (rasterio's windowed read, prediction.py:164, postprocessing.py:781-800).
    ``out``: optional destination of shape (bands, h, w) and the raster's dtype -- e.g. the NumPy view of a
    PINNED torch tensor, so that the pixels go from the page cache to the buffer the H2D copy reads with a
    single copy.  The file is memory mapped; uncompressed data is never staged anywhere else."""
    import mmap
    with open(path, "rb") as f:
        buf = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
        try:
            return _read_mapped(buf, window, out, fd=f.fileno())
        finally:
            try:
                buf.close()
            except BufferError:      # a zero-copy view is still alive (never the case for the returned array)
                pass


_DEVICE_LZW_MAX_CHUNK = 1 << 20     # td_tiff_lzw_decode_batch: decoded bytes per strip / tile


def device_decodable(path):
    """True when :func:`read_device` can decode the raster on the GPU: LZW strips / tiles of at most 1 MiB,
    8-bit samples (predictor 1 or 2, up to 4 interleaved bands or any number of planes) or 32-bit samples
    (predictor 1)."""
    import mmap
    with open(path, "rb") as f:
        buf = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
    try:
        return _device_plan(buf) is not None
    except Exception:
        return False
    finally:
        buf.close()


def _device_plan(buf):
    kind = _tiff_kind(buf)
    if kind is None:
        return None
    bo = kind[0]
    t = _read_ifd(buf, bo)
    info = _info_from_tags(t)
    comp, pred, planar = int(t.get(259, (1,))[0]), int(t.get(317, (1,))[0]), int(t.get(284, (1,))[0])
    size = info.dtype.itemsize
    if comp != 5 or size not in (1, 4) or pred not in (1, 2) or (pred == 2 and size != 1) or (size == 4 and bo != "<"):
        return None
    W, H, C = info.width, info.height, info.count
    if planar != 2 and C > 4:
        return None
    spp = 1 if planar == 2 else C
    tiled = 322 in t
    if tiled:
        chunk_rows, chunk_cols = int(t[323][0]), int(t[322][0])
        offs, cnts = t[324], t[325]
    else:
        chunk_rows, chunk_cols = min(int(t.get(278, (H,))[0]), H), W
        offs, cnts = t[273], t[279]
    chunk_bytes = chunk_rows * chunk_cols * spp * size
    if chunk_bytes > _DEVICE_LZW_MAX_CHUNK:
        return None
    nx, ny = (W + chunk_cols - 1) // chunk_cols, (H + chunk_rows - 1) // chunk_rows
    n = nx * ny * (C if planar == 2 else 1)
    if len(offs) != n or len(cnts) != n:
        return None
    rows = np.full(ny, chunk_rows, dtype=np.int64)
    if not tiled:
        rows[-1] = H - (ny - 1) * chunk_rows
    dst_len = np.tile(np.repeat(rows * chunk_cols * spp * size, nx), C if planar == 2 else 1).astype(np.int32)
    return {"info": info, "pred": pred, "planar": planar, "chunk_rows": chunk_rows, "chunk_cols": chunk_cols,
            "chunk_bytes": chunk_bytes, "n": n, "src_pos": np.asarray(offs, dtype=np.int64),
            "src_len": np.asarray(cnts, dtype=np.int32), "dst_len": dst_len}


_dev_scratch = {}


def _scratch(device, name, nbytes, pinned=False):
    """byte buffers re-used across images (pinned host staging of the compressed file, device copies, decoded chunks)"""
    import torch
    key = (str(device), name, pinned)
    t = _dev_scratch.get(key)
    if t is None or t.numel() < nbytes:
        cap = int(nbytes * 1.25) + 4096
        t = torch.empty((cap,), dtype=torch.uint8, pin_memory=True) if pinned else \
            torch.empty((cap,), dtype=torch.uint8, device=device)
        _dev_scratch[key] = t
    return t


def read_device(path, device, out=None, slot=0):
    """The raster decoded ON THE DEVICE: the file travels over PCIe still LZW-compressed, every strip / tile is one
    warp of ``td_tiff_lzw_decode_batch``, ``td_tiff_place_chunks`` undoes the predictor and writes the planar
    (bands, H, W) tensor.  Returns (tensor, GeoInfo, status), or None when the file needs the host reader (see
    :func:`device_decodable`).  Work is enqueued on the current stream; ``out``: optional destination tensor;
    the call returns once the compressed bytes have left the (single, re-used) pinned staging buffer -- the decode
    itself is still running; ``slot`` (0..7) selects the status word, so that several reads may be in flight.  The
    third return value is a one-element int32 device tensor: non-zero
    once the stream has run means a corrupt or oversized LZW stream (TD_ERR_*)."""
    import torch

    from . import _lib
    import time
    t_begin = time.perf_counter()
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        import mmap
        buf = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
        try:
            plan = _device_plan(buf)
        finally:
            buf.close()
        if plan is None:
            return None
        # the compressed file -> pinned staging (parallel pread) -> device
        host = _scratch(device, "file", size, pinned=True)
        mv = memoryview(host.numpy())[:size]
        step = 8 << 20

        def fill(lo):
            done = lo
            hi = min(lo + step, size)
            while done < hi:
                got = os.preadv(f.fileno(), [mv[done:hi]], done)
                if got <= 0:
                    raise ValueError("short read")
                done += got
        if size > step:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=8) as ex:
                list(ex.map(fill, range(0, size, step)))
        else:
            fill(0)
    t_read = time.perf_counter()
    info = plan["info"]
    n = plan["n"]
    # the compressed bytes travel on a copy stream of their own, into one of two device buffers, so that the copy of
    # this raster overlaps the decode kernels of the previous one (same compute stream) instead of queueing behind them
    state = _dev_scratch.setdefault((str(device), "lzw_state"), {"copy": torch.cuda.Stream(device=device), "k": 0,
                                                                  "done": [None, None]})
    half = state["k"] & 1
    state["k"] += 1
    cur = torch.cuda.current_stream(device)
    meta = torch.from_numpy(np.concatenate([plan["src_pos"].view(np.int32), plan["src_len"], plan["dst_len"]]))
    meta_pin = _scratch(device, "meta", meta.numel() * 4, pinned=True)[:meta.numel() * 4].view(torch.int32)
    meta_pin.copy_(meta)
    with torch.cuda.stream(state["copy"]):
        dev_file = _scratch(device, f"file{half}", size)
        meta_dev = _scratch(device, f"meta{half}", meta.numel() * 4 + 64)[:meta.numel() * 4].view(torch.int32)
        if state["done"][half] is not None:
            state["copy"].wait_event(state["done"][half])      # the decode that read this half last
        dev_file[:size].copy_(host[:size], non_blocking=True)
        meta_dev.copy_(meta_pin, non_blocking=True)
        copied = state["copy"].record_event()
    cur.wait_event(copied)
    dev_file.record_stream(cur)
    meta_dev.record_stream(cur)
    src_pos = meta_dev[:2 * n].view(torch.int64)
    src_len, dst_len = meta_dev[2 * n:3 * n], meta_dev[3 * n:4 * n]
    stride = plan["chunk_bytes"]
    decoded = _scratch(device, "decoded", n * stride)
    status = _scratch(device, f"status{slot}", 64)[:4].view(torch.int32)
    tdt = {np.dtype(np.uint8): torch.uint8, np.dtype(np.float32): torch.float32}[info.dtype]
    if out is None:
        out = torch.empty((info.count, info.height, info.width), dtype=tdt, device=device)
    elif tuple(out.shape) != (info.count, info.height, info.width) or out.dtype != tdt or not out.is_cuda:
        raise ValueError("read_device: out does not match the raster")
    st = cur.cuda_stream
    _lib.call("td_tiff_lzw_decode_batch", dev_file.data_ptr(), src_pos.data_ptr(), src_len.data_ptr(), n,
              decoded.data_ptr(), stride, dst_len.data_ptr(), None, status.data_ptr(), st)
    _lib.call("td_tiff_place_chunks", decoded.data_ptr(), stride, n, out.data_ptr(), info.count, info.height, info.width,
              info.dtype.itemsize, plan["planar"], plan["chunk_rows"], plan["chunk_cols"], plan["pred"], st)
    state["done"][half] = cur.record_event()
    copied.synchronize()        # the pinned staging (one set, re-used by the next call) has been read
    if os.environ.get("TREEDET_TRACE_TIFF"):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        torch.cuda.synchronize()
        t_sync = time.perf_counter()
        ev[0].record()
        _lib.call("td_tiff_lzw_decode_batch", dev_file.data_ptr(), src_pos.data_ptr(), src_len.data_ptr(), n,
                  decoded.data_ptr(), stride, dst_len.data_ptr(), None, status.data_ptr(), st)
        ev[1].record()
        torch.cuda.synchronize()
        print(f"read_device {os.path.basename(path)}: file -> pinned {t_read - t_begin:.3f} s, enqueue + H2D + kernels "
              f"{t_sync - t_read:.3f} s, decode kernel alone {ev[0].elapsed_time(ev[1]):.1f} ms, {n} chunks")
    return out, info, status


_io_pool = None


def _pool():
    global _io_pool
    if _io_pool is None:
        from concurrent.futures import ThreadPoolExecutor
        _io_pool = ThreadPoolExecutor(max_workers=_io_threads())
    return _io_pool


def _io_threads():
    """Threads that copy file pages into the staging buffers (``TREEDET_IO_THREADS``, default 8)."""
    try:
        return max(1, int(os.environ.get("TREEDET_IO_THREADS", "8")))
    except ValueError:
        return 8


def _pread_into(fd, mv, file_off):
    done = 0
    while done < len(mv):
        got = os.preadv(fd, [mv[done:]], file_off + done)
        if got <= 0:
            raise ValueError("truncated TIFF")
        done += got


def read_device_plain(path, device, out=None, slot=0, piece=32 << 20, probe=False):
    """An UNCOMPRESSED planar (or single-band) strip raster straight to the device: the file is read in pieces of
    32 MiB into a two-slot pinned ring (8 threads of ``pread``) and every piece is copied to its place in the
    (bands, H, W) device tensor while the next one is read -- no raster-sized pinned staging buffer (pinning 1 GB
    costs ~0.5 s), and the H2D copies hide behind the reads.  Returns (tensor, GeoInfo) or None when the layout
    needs the host reader (compression, chunky pixels, tiles, a predictor, big-endian samples).  ``probe``: only
    answer whether the layout qualifies (True / None)."""
    import mmap

    import torch
    with open(path, "rb") as f:
        buf = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
        try:
            kind = _tiff_kind(buf)
            if kind is None:
                return None
            bo = kind[0]
            t = _read_ifd(buf, bo)
            info = _info_from_tags(t)
        finally:
            buf.close()
        comp, pred, planar = int(t.get(259, (1,))[0]), int(t.get(317, (1,))[0]), int(t.get(284, (1,))[0])
        W, H, C = info.width, info.height, info.count
        tdt = {np.dtype(np.uint8): torch.uint8, np.dtype(np.float32): torch.float32}.get(info.dtype)
        if comp != 1 or pred != 1 or 322 in t or (planar != 2 and C > 1) or tdt is None or \
                (bo != "<" and info.dtype.itemsize > 1):
            return None
        rps = min(int(t.get(278, (H,))[0]), H)
        ns = (H + rps - 1) // rps
        offs, cnts = t[273], t[279]
        if len(offs) != ns * C:
            return None
        row_bytes = W * info.dtype.itemsize
        # runs of strips that are consecutive in the file AND in the destination
        runs = []          # [file offset, bytes, destination byte offset]
        for k in range(ns * C):
            p, s_ = divmod(k, ns)
            nbytes = min(rps, H - s_ * rps) * row_bytes
            if cnts[k] < nbytes:
                return None
            dst = (p * H + s_ * rps) * row_bytes
            if runs and runs[-1][0] + runs[-1][1] == offs[k] and runs[-1][2] + runs[-1][1] == dst:
                runs[-1][1] += nbytes
            else:
                runs.append([offs[k], nbytes, dst])
        if probe:
            return True
        if out is None:
            out = torch.empty((C, H, W), dtype=tdt, device=device)
        elif tuple(out.shape) != (C, H, W) or out.dtype != tdt or not out.is_cuda or not out.is_contiguous():
            raise ValueError("read_device_plain: out does not match the raster")
        flat = out.view(-1).view(torch.uint8)
        ring = _scratch(device, f"ring{slot}", 2 * piece, pinned=True)
        ring_np = memoryview(ring.numpy())
        events = [None, None]
        pool, k = _pool(), 0
        for file_off, nbytes, dst in runs:
            for lo in range(0, nbytes, piece):
                n = min(piece, nbytes - lo)
                half = k & 1
                if events[half] is not None:
                    events[half].synchronize()           # the copy out of this half of the ring is done
                base = half * piece
                step = max(1 << 20, -(-n // _io_threads()))
                list(pool.map(lambda a: _pread_into(f.fileno(), ring_np[base + a:base + min(a + step, n)],
                                                    file_off + lo + a), range(0, n, step)))
                flat[dst + lo:dst + lo + n].copy_(ring[base:base + n], non_blocking=True)
                events[half] = torch.cuda.current_stream(device).record_event()
                k += 1
        for ev in events:                                # the ring may be re-used by the next call
            if ev is not None:
                ev.synchronize()
    return out, info


def _read_mapped(buf, window, out, fd=None):
    kind = _tiff_kind(buf)
    if kind is None:
        raise ValueError("not a TIFF / BigTIFF file")
    bo = kind[0]
    t = _read_ifd(buf, bo)
    info = _info_from_tags(t)
    comp = int(t.get(259, (1,))[0])
    if comp not in (1, 5, 8, 32946):
        raise ValueError(f"unsupported TIFF compression {comp}")
    pred = int(t.get(317, (1,))[0])
    if pred not in (1, 2, 3) or (pred == 3 and info.dtype != np.float32):
        raise ValueError(f"unsupported TIFF predictor {pred}")
    planar = int(t.get(284, (1,))[0])
    W, H, C = info.width, info.height, info.count
    dt = info.dtype.newbyteorder(bo)
    if window is None:
        c0, r0, w, h = 0, 0, W, H
    else:
        c0, r0, w, h = (int(v) for v in window)
        if c0 < 0 or r0 < 0 or w <= 0 or h <= 0 or c0 + w > W or r0 + h > H:
            raise ValueError(f"window {window} outside the {W}x{H} raster")
    if out is None:
        out = np.empty((C, h, w), dtype=info.dtype)
    elif tuple(out.shape) != (C, h, w) or out.dtype != info.dtype:
        raise ValueError(f"out has shape {out.shape} / {out.dtype}, the read needs {(C, h, w)} / {info.dtype}")
    spp = 1 if planar == 2 else C              # samples per pixel inside one chunk
    tiled = 322 in t
    if tiled:
        chunk_rows, chunk_cols = int(t[323][0]), int(t[322][0])
        all_offs, all_cnts = t[324], t[325]
    else:
        chunk_rows, chunk_cols = int(t.get(278, (H,))[0]), W
        all_offs, all_cnts = t[273], t[279]
    chunk_rows = min(chunk_rows, H) if not tiled else chunk_rows
    chunk_bytes = chunk_rows * chunk_cols * spp * info.dtype.itemsize
    nx = (W + chunk_cols - 1) // chunk_cols
    ny = (H + chunk_rows - 1) // chunk_rows
    per_plane = nx * ny
    # chunks that intersect the window
    js = range(r0 // chunk_rows, (r0 + h - 1) // chunk_rows + 1)
    is_ = range(c0 // chunk_cols, (c0 + w - 1) // chunk_cols + 1)
    need = [(p, j, i) for p in range(C if planar == 2 else 1) for j in js for i in is_]

    def unpredict(raw, rows):
        """undo the TIFF predictor of one decoded chunk (rows x chunk_cols x spp samples)"""
        if pred == 1:
            return raw
        if pred == 2:      # horizontal differencing per sample, modulo the sample width
            a = np.frombuffer(raw, dtype=dt)[:rows * chunk_cols * spp].reshape(rows, chunk_cols, spp)
            return np.cumsum(a, axis=1, dtype=info.dtype).astype(dt).tobytes()
        # predictor 3: bytes of every row are plane-separated (most significant first) and differenced
        n = chunk_cols * spp
        b = np.frombuffer(raw, dtype=np.uint8)[:rows * n * 4].reshape(rows, 4 * n)
        b = np.cumsum(b.reshape(rows, 4, n).reshape(rows, 4 * n).astype(np.uint8), axis=1, dtype=np.uint8)
        b = b.reshape(rows, 4, n)[:, ::-1, :] if bo == "<" else b.reshape(rows, 4, n)
        return np.ascontiguousarray(b.transpose(0, 2, 1)).tobytes()

    def rows_of(j):
        return chunk_rows if tiled else min(chunk_rows, H - j * chunk_rows)

    def decode(pji):
        p, j, i = pji
        k = p * per_plane + j * nx + i
        off, cnt = all_offs[k], all_cnts[k]
        if comp == 1:
            # zero-copy view of the mapped file: the only copy is the one into `out`
            return np.frombuffer(buf, dtype=dt, count=rows_of(j) * chunk_cols * spp, offset=off)
        raw = buf[off:off + cnt]
        raw = _lzw(raw, chunk_bytes) if comp == 5 else zlib.decompress(raw)
        return np.frombuffer(unpredict(raw, rows_of(j)), dtype=dt)

    def place(pji):
        """decode one chunk and copy its intersection with the window into ``out``"""
        p, j, i = pji
        rr = rows_of(j)
        y0, x0 = j * chunk_rows, i * chunk_cols
        # intersection of the chunk with the window, in chunk and in output coordinates
        ya, yb = max(y0, r0), min(y0 + min(rr, H - y0), r0 + h)
        xa, xb = max(x0, c0), min(x0 + min(chunk_cols, W - x0), c0 + w)
        if ya >= yb or xa >= xb:
            return
        if direct and xa == x0 and xb - xa == chunk_cols:
            # uncompressed planar rows of full width: the kernel copies them from the page cache straight into
            # `out` (pread: no page faults on a mapping, the GIL is released)
            dst = out[p, ya - r0:yb - r0]
            k = p * per_plane + j * nx + i
            row_bytes = chunk_cols * info.dtype.itemsize
            mv, off, done = memoryview(dst).cast("B"), all_offs[k] + (ya - y0) * row_bytes, 0
            while done < len(mv):
                got = os.preadv(fd, [mv[done:]], off + done)
                if got <= 0:
                    raise ValueError("truncated TIFF strip")
                done += got
            return
        a = decode(pji)
        if planar == 2:
            out[p, ya - r0:yb - r0, xa - c0:xb - c0] = a[:rr * chunk_cols].reshape(rr, chunk_cols)[ya - y0:yb - y0,
                                                                                                  xa - x0:xb - x0]
        else:
            out[:, ya - r0:yb - r0, xa - c0:xb - c0] = \
                a[:rr * chunk_cols * C].reshape(rr, chunk_cols, C)[ya - y0:yb - y0, xa - x0:xb - x0].transpose(2, 0, 1)

    direct = (fd is not None and comp == 1 and planar == 2 and pred == 1 and dt == info.dtype and w == W and
              out.flags.c_contiguous and hasattr(os, "preadv"))
    # chunks are independent and land in disjoint parts of `out`; zlib, the LZW decoder and NumPy's copy loops all
    # release the GIL, so a few threads multiply the decode / copy rate (one thread moves ~3 GB/s into pinned memory)
    total_bytes = len(need) * chunk_bytes
    if len(need) > 4 and (comp != 1 or total_bytes >= (64 << 20)):
        from concurrent.futures import ThreadPoolExecutor
        workers = min(16 if comp != 1 else 8, os.cpu_count() or 1, len(need))
        with ThreadPoolExecutor(max_workers=workers) as ex:
            list(ex.map(place, need))
    else:
        for pji in need:
            place(pji)
    if window is not None:
        a, b, c, d, e, f = info.transform
        info = GeoInfo(w, h, C, info.dtype, (a, b, a * c0 + b * r0 + c, d, e, d * c0 + e * r0 + f), info.epsg, info.nodata)
    return out, info


def _lzw_encode(raw: bytes) -> bytes:
    import ctypes as C
    from . import _lib
    cap = len(raw) * 3 // 2 + 16
    dst = C.create_string_buffer(cap)
    n = _lib.lib().td_tiff_lzw_encode(raw, len(raw), dst, cap)
    if n < 0:
        _lib.check(int(n), "td_tiff_lzw_encode")
    return dst.raw[:n]


def write(path, array, transform, epsg=None, nodata=None, compression=None, predictor=1, bigtiff=None):
    """Little-endian, planar (band-sequential) strips -- one strip per band row block, so that a device tensor's
    (bands, H, W) layout is written without a transpose.  ``compression``: None (strips of 4 MiB) or ``"lzw"``
    (td_tiff_lzw_encode on a thread pool, strips of at most 128 KiB so that a raster has thousands of independent
    streams; ``predictor`` 2 = horizontal differencing, 8-bit samples only).  ``bigtiff``: None = BigTIFF (64-bit
    offsets) only when the file would pass 4 GB, True / False force the layout."""
    arr = np.ascontiguousarray(array)
    if arr.ndim == 2:
        arr = arr[None]
    C, H, W = arr.shape
    dt = arr.dtype
    fmt = {"u": 1, "i": 2, "f": 3}[dt.kind]
    bits = dt.itemsize * 8
    if compression not in (None, "lzw"):
        raise ValueError(f"unsupported compression {compression!r}")
    if predictor not in (1, 2) or (predictor == 2 and (compression is None or dt.itemsize != 1)):
        raise ValueError("predictor 2 needs LZW compression and 8-bit samples")
    strip_target = (1 << 22) if compression is None else (1 << 17)
    rps = max(1, min(H, strip_target // max(1, W * dt.itemsize)))
    ns = (H + rps - 1) // rps
    a, b, c, d, e, f = transform
    entries = []
    extra = bytearray()

    def add(tag, typ, values):
        fmtc, size = _TYPES[typ]
        if typ == 2:
            raw = values.encode("latin1") + b"\x00"
            cnt = len(raw)
        else:
            cnt = len(values)
            raw = struct.pack("<" + fmtc * cnt, *values)
        entries.append((tag, typ, cnt, raw))

    add(256, 4, [W]); add(257, 4, [H]); add(258, 3, [bits] * C); add(259, 3, [1 if compression is None else 5])
    add(262, 3, [1]); add(277, 3, [C]); add(278, 4, [rps]); add(284, 3, [2]); add(339, 3, [fmt] * C)
    if predictor != 1:
        add(317, 3, [predictor])
    if C > 1:
        add(338, 3, [0] * (C - 1))
    add(33550, 12, [abs(a), abs(e), 0.0])
    add(33922, 12, [0.0, 0.0, 0.0, c, f, 0.0])
    if epsg is not None:
        add(34735, 3, [1, 1, 0, 3, 1024, 0, 1, 1, 1025, 0, 1, 1, 3072, 0, 1, int(epsg)])
    if nodata is not None:
        add(42113, 2, repr(float(nodata)))
    strip_bytes = [min(rps, H - s * rps) * W * dt.itemsize for s in range(ns)] * C
    strips = None
    if compression == "lzw":
        le = arr.astype(dt.newbyteorder("<"), copy=False)

        def pack(k):
            p, s_ = divmod(k, ns)
            block = le[p, s_ * rps:min((s_ + 1) * rps, H)]
            if predictor == 2:
                block = np.diff(block, axis=1, prepend=np.zeros((block.shape[0], 1), dtype=block.dtype))
            return _lzw_encode(np.ascontiguousarray(block).tobytes())
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
            strips = list(ex.map(pack, range(ns * C)))
        strip_bytes = [len(b) for b in strips]
    payload = sum(strip_bytes)
    big = bool(bigtiff) if bigtiff is not None else payload + (1 << 20) + 16 * ns * C >= (1 << 32)
    otyp = 16 if big else 4                    # LONG8 / LONG strip offsets and byte counts
    add(273, otyp, [0] * (ns * C)); add(279, otyp, strip_bytes)
    entries.sort(key=lambda x: x[0])
    # classic: 8-byte header, 2-byte entry count, 12-byte entries (4 bytes inline), 4-byte next-IFD offset;
    # BigTIFF: 16-byte header, 8-byte count, 20-byte entries (8 bytes inline), 8-byte next-IFD offset
    hdr, ncnt, esize, inline, ofmt = (16, 8, 20, 8, "<Q") if big else (8, 2, 12, 4, "<I")
    ifd_off = hdr
    ifd_size = ncnt + esize * len(entries) + inline
    extra_off = ifd_off + ifd_size
    # layout: header | IFD | out-of-line values | pixel data
    pos = extra_off
    placed = []
    for tag, typ, cnt, raw in entries:
        if len(raw) <= inline:
            placed.append((tag, typ, cnt, raw.ljust(inline, b"\x00"), None))
        else:
            placed.append((tag, typ, cnt, struct.pack(ofmt, pos), raw))
            pos += len(raw) + (len(raw) & 1)
    data_off = (pos + 15) & ~15
    offsets, o = [], data_off
    for sb in strip_bytes:
        offsets.append(o); o += sb
    if o >= (1 << 32) and not big:
        raise ValueError("raster too large for classic TIFF (bigtiff=False)")
    with open(path, "wb") as fh:
        if big:
            fh.write(b"II" + struct.pack("<HHHQ", 43, 8, 0, ifd_off))
            fh.write(struct.pack("<Q", len(placed)))
        else:
            fh.write(b"II" + struct.pack("<HI", 42, ifd_off))
            fh.write(struct.pack("<H", len(placed)))
        blob = bytearray()
        for tag, typ, cnt, val, raw in placed:
            if tag == 273:
                raw = struct.pack("<" + _TYPES[otyp][0] * len(offsets), *offsets)
                if len(raw) <= inline:
                    val, raw = raw.ljust(inline, b"\x00"), None
            fh.write(struct.pack("<HHQ" if big else "<HHI", tag, typ, cnt) + val)
            if raw is not None:
                blob += raw + (b"\x00" if len(raw) & 1 else b"")
        fh.write(struct.pack(ofmt, 0))
        fh.write(bytes(blob))
        fh.write(b"\x00" * (data_off - extra_off - len(blob)))
        if strips is None:
            fh.write(arr.astype(dt.newbyteorder("<"), copy=False).tobytes())
        else:
            for b in strips:
                fh.write(b)
