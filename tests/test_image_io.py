"""Image writers of the C-ABI (include/oclr_abi.h: oclr_write_bmp / _ppm16 / _png16) -- host code, runs without a GPU.
Reference behaviour: source/util/writebmp.cpp:124-177 (layout, low-byte cast) and source/render.cpp:1381-1383 (value / 256)."""
import struct
import zlib

import numpy as np
import pytest

from opencl_render_b200 import api


def _planes(w, h, seed=0):
    rng = np.random.default_rng(seed)
    return tuple(rng.integers(0, 65536, (h, w), dtype=np.uint16) for _ in range(3))


@pytest.mark.parametrize("w,h", [(5, 3), (8, 2), (1, 1), (7, 9)])
@pytest.mark.parametrize("cast", [False, True])
def test_bmp_layout_and_conversion(tmp_path, w, h, cast):
    r, g, b = _planes(w, h, w * 31 + h)
    path = tmp_path / "img.bmp"
    api.write_image(path, (r, g, b), bmp_reference_cast=cast)
    data = path.read_bytes()
    row = (3 * w + 3) // 4 * 4
    assert data[:2] == b"BM" and len(data) == 54 + row * h
    size, off = struct.unpack_from("<I4xI", data, 2)
    hdr, bw, bh, planes, bpp = struct.unpack_from("<IiiHH", data, 14)
    # writebmp3s writes bfSize = 54 + 3*w*h (row padding not counted) and leaves biSizeImage 0 (writebmp.cpp:129, 146-152); the cast
    # mode reproduces that header byte for byte, the corrected mode writes the true sizes
    want_size = 54 + 3 * w * h if cast else len(data)
    assert (size, off, hdr, bw, bh, planes, bpp) == (want_size, 54, 40, w, h, 1, 24)
    assert struct.unpack_from("<I", data, 34)[0] == (0 if cast else row * h)
    if cast:
        head = bytearray(54)
        head[0:2] = b"BM"
        struct.pack_into("<I", head, 2, 54 + 3 * w * h)
        head[10] = 54
        head[14] = 40
        struct.pack_into("<ii", head, 18, w, h)
        head[26], head[28] = 1, 24
        assert data[:54] == bytes(head)
    conv = (lambda a: (a & 0xFF).astype(np.uint8)) if cast else (lambda a: (a // 256).astype(np.uint8))
    for j in range(h):          # file rows are bottom-up, pixels BGR (writebmp.cpp:136-141, 167-170)
        line = np.frombuffer(data, np.uint8, 3 * w, 54 + row * j).reshape(w, 3)
        src = h - 1 - j
        assert np.array_equal(line[:, 2], conv(r[src])) and np.array_equal(line[:, 1], conv(g[src])) and np.array_equal(line[:, 0], conv(b[src]))
        assert data[54 + row * j + 3 * w: 54 + row * (j + 1)] == b"\0" * (row - 3 * w)


def test_ppm16_round_trip(tmp_path):
    r, g, b = _planes(13, 6, 5)
    path = tmp_path / "img.ppm"
    api.write_image(path, (r, g, b))
    data = path.read_bytes()
    head = b"P6\n13 6\n65535\n"
    assert data.startswith(head)
    px = np.frombuffer(data[len(head):], ">u2").reshape(6, 13, 3)
    assert np.array_equal(px[..., 0], r) and np.array_equal(px[..., 1], g) and np.array_equal(px[..., 2], b)


def test_png16_decodes_to_the_planes(tmp_path):
    r, g, b = _planes(17, 11, 9)
    path = tmp_path / "img.png"
    api.write_image(path, (r, g, b))
    data = path.read_bytes()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, chunks = 8, []
    while pos < len(data):      # every chunk's CRC must hold
        n, kind = struct.unpack_from(">I4s", data, pos)
        body = data[pos + 8: pos + 8 + n]
        assert struct.unpack_from(">I", data, pos + 8 + n)[0] == zlib.crc32(kind + body)
        chunks.append((kind, body))
        pos += 12 + n
    assert [k for k, _ in chunks] == [b"IHDR", b"IDAT", b"IEND"]
    assert struct.unpack(">IIBBBBB", chunks[0][1]) == (17, 11, 16, 2, 0, 0, 0)
    raw = zlib.decompress(chunks[1][1])
    rows = np.frombuffer(raw, np.uint8).reshape(11, 1 + 17 * 6)
    assert not rows[:, 0].any()                                     # filter type 0 on every row
    px = rows[:, 1:].copy().view(">u2").reshape(11, 17, 3)
    assert np.array_equal(px[..., 0], r) and np.array_equal(px[..., 1], g) and np.array_equal(px[..., 2], b)
    PIL = pytest.importorskip("PIL.Image")
    im = PIL.open(path)
    assert im.size == (17, 11)


def test_writers_fail_loudly(tmp_path):
    r, g, b = _planes(4, 4)
    with pytest.raises(api.OclrError):
        api.write_image(tmp_path / "no_such_dir" / "img.bmp", (r, g, b))
    with pytest.raises(ValueError):
        api.write_image(tmp_path / "img.tiff", (r, g, b))
