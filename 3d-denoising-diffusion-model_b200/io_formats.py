"""File formats either side of the path: 3-D TIFF in (scripts/test.py:192-199),
`.npz` (`arr_0`, (H,W,Z)) + float32 `.tif` ((Z,H,W)) out (scripts/test.py:168-179), and the README's `.npz`
input (`arr_0` = (2,X,Y,Z) low/high pair divided by 4, image_datasets.py:592-603).  `tifffile` is used when it
is installed; otherwise a minimal baseline-TIFF codec (uncompressed, strips, one sample per pixel) is enough
for the volumes this model consumes and produces."""
from __future__ import annotations

import struct

import numpy as np

_TYPES = {1: "B", 3: "H", 4: "I", 16: "Q"}
_DTYPES = {(1, 8): np.uint8, (1, 16): np.uint16, (1, 32): np.uint32, (2, 8): np.int8, (2, 16): np.int16,
           (2, 32): np.int32, (3, 32): np.float32, (3, 64): np.float64}


def _read_tiff(path):
    data = open(path, "rb").read()
    bo = {b"II": "<", b"MM": ">"}[data[:2]]
    if struct.unpack(bo + "H", data[2:4])[0] != 42:
        raise ValueError("only classic TIFF is supported without tifffile")
    off = struct.unpack(bo + "I", data[4:8])[0]
    pages = []
    while off:
        n = struct.unpack(bo + "H", data[off:off + 2])[0]
        tags = {}
        for k in range(n):
            e = off + 2 + 12 * k
            tag, typ, cnt = struct.unpack(bo + "HHI", data[e:e + 8])
            if typ not in _TYPES:
                continue
            size = struct.calcsize(_TYPES[typ]) * cnt
            src = data[e + 8:e + 8 + size] if size <= 4 else None
            if src is None:
                p = struct.unpack(bo + "I", data[e + 8:e + 12])[0]
                src = data[p:p + size]
            tags[tag] = struct.unpack(bo + _TYPES[typ] * cnt, src)
        if tags.get(259, (1,))[0] != 1:
            raise ValueError("compressed TIFF needs tifffile")
        w, h = tags[256][0], tags[257][0]
        bits, fmt = tags.get(258, (8,))[0], tags.get(339, (1,))[0]
        dt = np.dtype(_DTYPES[(fmt, bits)]).newbyteorder(bo)
        raw = b"".join(data[o:o + c] for o, c in zip(tags[273], tags[279]))
        pages.append(np.frombuffer(raw, dtype=dt, count=w * h).reshape(h, w))
        off = struct.unpack(bo + "I", data[off + 2 + 12 * n:off + 6 + 12 * n])[0]
    return np.stack(pages).astype(pages[0].dtype.newbyteorder("="))


def _write_tiff(path, vol):
    vol = np.ascontiguousarray(vol, dtype=np.float32)
    if vol.ndim == 2:
        vol = vol[None]
    z, h, w = vol.shape
    out = bytearray(b"II" + struct.pack("<HI", 42, 8))
    page_bytes = h * w * 4
    ifd_size = 2 + 12 * 9 + 4
    for k in range(z):
        ifd_off = len(out)
        data_off = ifd_off + ifd_size
        nxt = data_off + page_bytes if k + 1 < z else 0
        entries = [(256, 4, w), (257, 4, h), (258, 3, 32), (259, 3, 1), (262, 3, 1), (273, 4, data_off),
                   (277, 3, 1), (279, 4, page_bytes), (339, 3, 3)]
        out += struct.pack("<H", len(entries))
        for tag, typ, val in entries:
            out += struct.pack("<HHI", tag, typ, 1) + (struct.pack("<I", val) if typ == 4 else struct.pack("<HH", val, 0))
        out += struct.pack("<I", nxt)
        out += vol[k].astype("<f4").tobytes()
    open(path, "wb").write(bytes(out))


def read_volume(path):
    """-> float32 (D,H,W) low-dose volume, un-normalised like scripts/test.py:201-203."""
    if path.endswith((".tif", ".tiff")):
        try:
            import tifffile
            vol = tifffile.imread(path)
        except ImportError:
            vol = _read_tiff(path)
        if vol.ndim == 4 and vol.shape[0] == 1:
            vol = vol[0]
        return np.asarray(vol, dtype=np.float32)
    if path.endswith(".npz"):
        arr = np.load(path)["arr_0"]
        if arr.ndim == 4 and arr.shape[0] == 2:  # (low, high) pair in (X,Y,Z), /4 (image_datasets.py:592-603)
            arr = arr[0] / 4.0
            return np.ascontiguousarray(np.transpose(arr, (2, 0, 1)), dtype=np.float32)
        return np.asarray(arr, dtype=np.float32)
    if path.endswith(".npy"):
        return np.asarray(np.load(path), dtype=np.float32)
    raise ValueError("Unsupported file type")  # scripts/test.py:187-190


def write_result(out_path_npz, arr_hwz):
    """scripts/test.py:168-179: np.savez(arr_0 = (H,W,Z)) and a float32 TIFF in (Z,H,W)."""
    arr = np.asarray(arr_hwz, dtype=np.float32)
    np.savez(out_path_npz, arr)
    tif_path = out_path_npz.replace(".npz", ".tif")
    zhw = arr.transpose(2, 0, 1)
    try:
        import tifffile
        tifffile.imwrite(tif_path, zhw.astype(np.float32))
    except ImportError:
        _write_tiff(tif_path, zhw)
    return tif_path
