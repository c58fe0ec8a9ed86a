"""Diagnostic: LZW GeoTIFF -> device raster, GPU decoder (geotiff.read_device) against the host reader + H2D.
usage: python scripts/time_tiff_device.py [size_px] [predictor]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from PIL import Image  # noqa: E402

from treedetection_b200 import geotiff, synth  # noqa: E402

Image.MAX_IMAGE_PIXELS = None
size = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
pred = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sc = synth.make_scene(seed=1234, size_px=size, px=0.2, ndsm_px=0.2, density_per_km2=2500.0)
dev = torch.device("cuda", 0)
torch.zeros(1, device=dev)
for name, arr, p in (("rgbi", sc.rgbi, pred), ("ndsm", sc.ndsm, None)):
    path = f"/dev/shm/_lzw_{name}.tif"
    t = time.time()
    if os.environ.get("TIFF_WRITER", "own") == "pil":      # libtiff: chunky pixels
        img = Image.fromarray(np.ascontiguousarray(arr.transpose(1, 2, 0))) if arr.ndim == 3 else Image.fromarray(arr)
        img.save(path, format="TIFF", compression="tiff_lzw", tiffinfo={317: p} if p else {})
    else:                                                  # this package's writer: planar strips of <= 128 KiB
        geotiff.write(path, arr, (0.2, 0.0, 412000.0, 0.0, -0.2, 5318000.0), epsg=25832, compression="lzw",
                      predictor=p or 1)
    raw = arr.nbytes
    print(f"{name}: {raw / 1e6:.0f} MB raw -> {os.path.getsize(path) / 1e6:.0f} MB LZW (predictor {p}), written in "
          f"{time.time() - t:.1f} s, device decodable: {geotiff.device_decodable(path)}")
    shape = (arr.shape if arr.ndim == 3 else (1,) + arr.shape)
    pinned = torch.empty(shape, dtype=torch.uint8 if arr.dtype == np.uint8 else torch.float32, pin_memory=True)
    out = torch.empty(shape, dtype=pinned.dtype, device=dev)
    for k in range(3):
        t = time.time()
        geotiff.read(path, out=pinned.numpy())
        t1 = time.time()
        out.copy_(pinned, non_blocking=True)
        torch.cuda.synchronize()
        print(f"  host reader {t1 - t:.3f} s + H2D {time.time() - t1:.3f} s")
    ref = out.clone()
    for k in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t = time.time()
        a.record()
        got, info, status = geotiff.read_device(path, dev, out=out, slot=k & 1)
        b.record()
        torch.cuda.synchronize()
        print(f"  device reader {time.time() - t:.3f} s wall (device side {a.elapsed_time(b):.1f} ms), status "
              f"{int(status.item())}, equal {bool(torch.equal(out, ref))}, differing "
              f"{int((out != ref).sum().item())}")
    os.remove(path)
