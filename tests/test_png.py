"""Native PNG decoder of the indexing feed (csrc/png.cpp, sprc_png_decode_files) against Pillow — the decoder behind the
reference's `PIL.Image.open(path)` + `convert("RGB")` (src/data_utils.py:91-105,167-186,253-270).  Byte work: bit-exact.
CPU only (the decoder is host code; no kernel is launched)."""
import os
import struct
import sys
import zlib

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PIL = pytest.importorskip("PIL")
from PIL import Image  # noqa: E402

from sprc_b200.preprocess import PIL_MODES, PngBatchDecoder  # noqa: E402


def _pil_rgb(path):
    return np.asarray(Image.open(path).convert("RGB"))


def _smooth(rng, h, w, c):
    """Photo-like content (smooth gradients + noise) so that the encoder picks all five scanline filters."""
    y, x = np.mgrid[0:h, 0:w]
    chans = []
    for k in range(c):
        a, b, ph = rng.uniform(0.01, 0.2, 3)
        v = 127 + 90 * np.sin(a * x + ph) * np.cos(b * y) + rng.normal(0, 6, (h, w))
        chans.append(np.clip(v, 0, 255))
    return np.stack(chans, -1).astype(np.uint8)


def _write_png(path, w, h, depth, color, rows_bytes, plte=None, extra=(), filters=None, idat_split=3):
    """Hand-written PNG (chunks + zlib) for cases Pillow's writer cannot produce: low bit depths, chosen filters."""
    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)

    stride = len(rows_bytes[0])
    bpp = max(1, depth * {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[color] // 8)
    raw = bytearray()
    prev = bytes(stride)
    for y, row in enumerate(rows_bytes):
        f = filters[y % len(filters)] if filters else 0
        out = bytearray(stride)
        for i in range(stride):
            a = row[i - bpp] if i >= bpp else 0
            b = prev[i]
            c = prev[i - bpp] if i >= bpp else 0
            if f == 0:
                pred = 0
            elif f == 1:
                pred = a
            elif f == 2:
                pred = b
            elif f == 3:
                pred = (a + b) >> 1
            else:
                p = a + b - c
                pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
                pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
            out[i] = (row[i] - pred) & 0xFF
        raw += bytes([f]) + out
        prev = row
    comp = zlib.compress(bytes(raw), 6)
    parts = [comp[i * len(comp) // idat_split:(i + 1) * len(comp) // idat_split] for i in range(idat_split)]
    data = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, color, 0, 0, 0))
    for t, d in extra:
        data += chunk(t, d)
    if plte is not None:
        data += chunk(b"PLTE", plte)
    for p in parts:
        data += chunk(b"IDAT", p)
    data += chunk(b"IEND", b"")
    with open(path, "wb") as f:
        f.write(data)


@pytest.fixture(scope="module")
def dec():
    return PngBatchDecoder(threads=4, arena_bytes=1 << 20, pin=False)   # small arena: the growth path runs too


def test_pillow_written_pngs_all_modes_bit_exact(tmp_path, dec):
    rng = np.random.default_rng(0)
    paths, modes = [], []
    for i in range(40):
        h, w = int(rng.integers(1, 300)), int(rng.integers(1, 400))
        mode = ("RGB", "RGB", "RGB", "L", "RGBA", "LA", "P", "1")[i % 8]
        if mode == "RGB":
            im = Image.fromarray(_smooth(rng, h, w, 3), "RGB")
        elif mode == "L":
            im = Image.fromarray(_smooth(rng, h, w, 1)[..., 0], "L")
        elif mode == "RGBA":
            im = Image.fromarray(_smooth(rng, h, w, 4), "RGBA")
        elif mode == "LA":
            im = Image.fromarray(_smooth(rng, h, w, 2), "LA")
        elif mode == "P":
            im = Image.fromarray(_smooth(rng, h, w, 3), "RGB").quantize(int(rng.integers(2, 257)))
        else:
            im = Image.fromarray(_smooth(rng, h, w, 1)[..., 0] > 127)
        p = str(tmp_path / f"img{i}_{mode}.png")
        im.save(p, compress_level=int(rng.integers(0, 10)), optimize=bool(i % 3 == 0))
        paths.append(p)
        modes.append(Image.open(p).mode)
    b = dec.decode(paths)
    assert b.status.tolist() == [0] * len(paths)
    for i, p in enumerate(paths):
        want = _pil_rgb(p)
        assert (int(b.wh[i, 0]), int(b.wh[i, 1])) == (want.shape[1], want.shape[0])
        assert PIL_MODES[int(b.wh[i, 2])] == modes[i], (p, modes[i])
        assert np.array_equal(b.image(i), want), p


def test_hand_written_pngs_every_filter_low_bit_depths_and_split_idat(tmp_path, dec):
    rng = np.random.default_rng(1)
    paths = []
    # RGB8 / RGBA8 / gray8 / gray+alpha8 with every filter type cycling over the rows, IDAT in 1..5 chunks
    for k, (color, ch) in enumerate(((2, 3), (6, 4), (0, 1), (4, 2))):
        h, w = 37 + k, 53 + 3 * k
        px = _smooth(rng, h, w, ch)
        rows = [bytes(px[y].reshape(-1)) for y in range(h)]
        p = str(tmp_path / f"hand_c{color}.png")
        _write_png(p, w, h, 8, color, rows, filters=[0, 1, 2, 3, 4, 4, 3, 1], idat_split=k + 1,
                   extra=[(b"gAMA", struct.pack(">I", 45455)), (b"tEXt", b"Comment\x00hello")])
        paths.append(p)
    # 1/2/4-bit gray and palette (widths that do not fill the last byte)
    for depth in (1, 2, 4, 8):
        for color in (0, 3):
            h, w = 19, 45 + depth
            idx = rng.integers(0, 1 << depth, (h, w), dtype=np.uint8)
            per = 8 // depth
            rows = []
            for y in range(h):
                line = bytearray((w * depth + 7) // 8)
                for x in range(w):
                    line[x // per] |= int(idx[y, x]) << (8 - depth - (x % per) * depth)
                rows.append(bytes(line))
            plte = bytes(rng.integers(0, 256, 3 * (1 << depth), dtype=np.uint8)) if color == 3 else None
            extra = [(b"tRNS", bytes([0, 128]))] if color == 3 else []
            p = str(tmp_path / f"hand_d{depth}_c{color}.png")
            _write_png(p, w, h, depth, color, rows, plte=plte, filters=[0, 2, 1, 4, 3],
                       extra=[], idat_split=2)
            if extra:   # tRNS must follow PLTE: written by a second call with the chunk order PLTE, tRNS, IDAT
                data = open(p, "rb").read()
                i = data.index(b"IDAT") - 4
                t = b"tRNS" + extra[0][1]
                data = data[:i] + struct.pack(">I", len(extra[0][1])) + t + struct.pack(">I", zlib.crc32(t) & 0xFFFFFFFF) + data[i:]
                open(p, "wb").write(data)
            paths.append(p)
    b = dec.decode(paths)
    assert b.status.tolist() == [0] * len(paths), b.status.tolist()
    for i, p in enumerate(paths):
        assert PIL_MODES[int(b.wh[i, 2])] == Image.open(p).mode, p
        assert np.array_equal(b.image(i), _pil_rgb(p)), p


def test_files_the_decoder_leaves_to_pillow_and_broken_files(tmp_path, dec):
    rng = np.random.default_rng(2)
    good = str(tmp_path / "good.png")
    Image.fromarray(_smooth(rng, 40, 50, 3), "RGB").save(good)
    # 16-bit gray, a JPEG, an interlaced PNG: status 1 (Pillow decodes them)
    p16 = str(tmp_path / "g16.png")
    Image.fromarray((rng.integers(0, 65535, (20, 30))).astype(np.uint16)).save(p16)
    jpg = str(tmp_path / "photo.jpg")
    Image.fromarray(_smooth(rng, 40, 50, 3), "RGB").save(jpg)
    inter = str(tmp_path / "adam7.png")
    data = bytearray(open(good, "rb").read())
    data[28] = 1   # IHDR interlace byte (the CRC is now wrong too, but the header check comes first)
    open(inter, "wb").write(bytes(data))
    # corrupt: flipped byte inside IDAT (CRC), truncated file; unreadable: missing file
    bad_crc = str(tmp_path / "badcrc.png")
    data = bytearray(open(good, "rb").read())
    data[data.index(b"IDAT") + 20] ^= 0xFF
    open(bad_crc, "wb").write(bytes(data))
    trunc = str(tmp_path / "trunc.png")
    open(trunc, "wb").write(open(good, "rb").read()[:200])
    missing = str(tmp_path / "nope.png")
    paths = [good, p16, jpg, inter, bad_crc, trunc, missing, good]
    b = dec.decode(paths)
    assert b.status.tolist() == [0, 1, 1, 1, 2, 2, 3, 0]
    assert np.array_equal(b.image(0), _pil_rgb(good)) and np.array_equal(b.image(7), _pil_rgb(good))
    # what the statuses promise: Pillow reads the status-1 files, and refuses the corrupt ones
    for p in (p16, jpg):
        Image.open(p).convert("RGB")
    for p in (bad_crc, trunc):
        with pytest.raises(Exception):
            Image.open(p).convert("RGB")


def test_c_abi_contract(dec):
    from sprc_b200 import _lib as L

    lib = L.load()
    assert lib.sprc_png_decode_files(None, None, 0, 1, None, 0, None, None, None) != 0       # null arguments refused
    b = dec.decode([])                                                                        # an empty batch is fine
    assert b.status.size == 0 and int(b.offsets[0]) == 0


def test_large_batch_thread_counts_agree(tmp_path):
    rng = np.random.default_rng(3)
    paths = []
    for i in range(24):
        p = str(tmp_path / f"b{i}.png")
        Image.fromarray(_smooth(rng, int(rng.integers(100, 200)), int(rng.integers(100, 260)), 3), "RGB").save(p)
        paths.append(p)
    outs = []
    for th in (1, 3, 8):
        b = PngBatchDecoder(threads=th, arena_bytes=4 << 20, pin=False).decode(paths)
        assert b.status.tolist() == [0] * len(paths)
        outs.append(torch.cat([torch.from_numpy(b.image(i).copy()).flatten() for i in range(len(paths))]))
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    for i, p in enumerate(paths):
        assert np.array_equal(b.image(i), _pil_rgb(p))


# ---------------------------------------------------------------------------------------------------
# the decoder's own DEFLATE inflater (csrc/inflate.h) against zlib
# ---------------------------------------------------------------------------------------------------
def _inflate(lib, comp: bytes, n_out: int):
    out = np.zeros(max(n_out, 1), dtype=np.uint8)
    rc = lib.sprc_op_inflate_zlib(comp, len(comp), out.ctypes.data, n_out)
    return rc, out[:n_out].tobytes()


def _payloads(rng):
    yield b""
    yield b"a"
    yield b"abc" * 5
    yield bytes(1000)                                             # one long run (distance 1)
    yield bytes(rng.integers(0, 256, 70000, dtype=np.uint8))      # incompressible: stored blocks at level 0, literals else
    yield bytes(rng.integers(0, 4, 50000, dtype=np.uint8))        # tiny alphabet: short codes, many matches
    yield (b"0123456789abcdef" * 5000)[:70001]                    # long matches at distance 16
    yield bytes((np.arange(100000) % 251).astype(np.uint8))       # distance 251 matches of maximal length
    text = (b"the quick brown fox jumps over the lazy dog; " * 300)
    yield text
    img = _smooth(rng, 120, 160, 3)                               # filtered photo-like scanlines (what a PNG holds)
    yield bytes(np.diff(img.astype(np.int16), axis=1, prepend=0).astype(np.uint8).reshape(-1))
    big = rng.integers(0, 256, 300000, dtype=np.uint8)
    big[100000:200000] = big[0:100000]                            # matches at distance 100000 > 32768: not representable,
    yield bytes(big)                                              # the compressor must fall back to literals
    yield bytes(rng.choice(np.arange(256, dtype=np.uint8), 200000,
                           p=np.r_[np.full(8, 0.1), np.full(248, 0.2 / 248)]))   # skewed: code lengths up to 15


def test_inflate_equals_zlib_all_levels_and_strategies():
    from sprc_b200 import _lib as L

    lib = L.load()
    rng = np.random.default_rng(5)
    n = 0
    for data in _payloads(rng):
        for level in (0, 1, 3, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED):
                for wbits in (15, 9):
                    c = zlib.compressobj(level, zlib.DEFLATED, wbits, 9, strategy)
                    comp = c.compress(data) + c.flush()
                    rc, out = _inflate(lib, comp, len(data))
                    assert rc == 0, (len(data), level, strategy, wbits, rc)
                    assert out == data, (len(data), level, strategy, wbits)
                    n += 1
        # several deflate blocks in one stream (Z_FULL_FLUSH emits an empty stored block between them)
        c = zlib.compressobj(6)
        third = len(data) // 3
        comp = c.compress(data[:third]) + c.flush(zlib.Z_FULL_FLUSH) + c.compress(data[third:]) + c.flush()
        rc, out = _inflate(lib, comp, len(data))
        assert rc == 0 and out == data
    assert n >= 500


def test_inflate_refuses_what_it_must():
    from sprc_b200 import _lib as L

    lib = L.load()
    rng = np.random.default_rng(6)
    data = bytes(_smooth(rng, 64, 96, 3).reshape(-1))
    comp = zlib.compress(data, 6)
    assert _inflate(lib, comp, len(data))[0] == 0
    assert _inflate(lib, comp, len(data) - 1)[0] != 0            # output longer than the buffer
    assert _inflate(lib, comp, len(data) + 1)[0] != 0            # output shorter than the caller expects
    assert _inflate(lib, comp[:-1], len(data))[0] != 0           # trailer cut
    assert _inflate(lib, comp[:len(comp) // 2], len(data))[0] != 0
    bad = bytearray(comp)
    bad[-1] ^= 1
    assert _inflate(lib, bytes(bad), len(data))[0] != 0          # Adler-32 mismatch
    assert _inflate(lib, b"\x78\x9c" + b"\x07" + bytes(8), 0)[0] != 0   # reserved block type 3
    assert _inflate(lib, comp + b"trailing", len(data))[0] == 0  # bytes after the trailer are tolerated
    # random corruption anywhere in the stream: never a crash, and whenever the decoder says 0 the bytes are zlib's
    n_ok = 0
    for i in range(400):
        b = bytearray(comp)
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(2, len(b)))] ^= 1 << int(rng.integers(0, 8))
        rc, out = _inflate(lib, bytes(b), len(data))
        if rc == 0:
            n_ok += 1
            assert zlib.decompress(bytes(b)) == out
    assert n_ok < 20
