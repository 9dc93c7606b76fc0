"""PNG encoder oracle (numpy + plain Python, CPU).  TEST INFRASTRUCTURE - see oracle/__init__.py.

The reference writes its BEV inputs and targets with ``cv2.imwrite(path, image)``
(generating-dataset/generating_train_bev.py:215, :224, :229) - 8-bit PNG, colour images in
OpenCV's BGR convention - and reads them back with ``cv2.imread(..., cv2.IMREAD_UNCHANGED)``
(deeplab_v3_baseline/dataset/dataset.py:83-90).  PNG is lossless, so the contract of this row
is *decode-exactness*: ``cv2.imdecode(encode(image)) == image``; the compressed bytes themselves are
an encoder's choice (libpng's own strategy is not part of the reference's behaviour).

This file states the byte stream the CUDA encoder (csrc/lv_png.cu) must produce, so that the GPU
output can ALSO be compared byte for byte:

  signature, IHDR, one IDAT holding a zlib stream (header 78 01), IEND.
  scanlines: filter type 0; colour images are written R,G,B from the array's B,G,R (= cv2.imwrite).
  deflate: ONE fixed-Huffman block (BFINAL=1, BTYPE=01).  Every scanline is tokenised on its own:
           each run of R equal bytes = literal + distance-1 matches of min(rest, 258) while rest >= 3
           + `rest` literals.  End-of-block, zero padding to a byte.
  Adler-32 of the filtered scanlines, CRC-32 of every chunk.
"""
import struct
import zlib

import numpy as np

LEN_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
LEN_EXTRA = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0]


def _rev(code, n):
    r = 0
    for _ in range(n):
        r = (r << 1) | (code & 1)
        code >>= 1
    return r


def literal_token(v):
    """(value, nbits) of literal byte v in the fixed Huffman code, ready for LSB-first packing
    (RFC 1951 3.2.6: Huffman codes are packed most-significant bit first, hence the reversal)."""
    if v < 144:
        return _rev(0x30 + v, 8), 8
    return _rev(0x190 + (v - 144), 9), 9


def match_token(length):
    """(value, nbits) of a match of `length` (3..258) at distance 1: length symbol, extra bits
    (LSB first), 5-bit distance code 0."""
    idx = max(i for i, b in enumerate(LEN_BASE) if b <= length)
    if length == 258:
        idx = 28
    sym = 257 + idx
    if sym < 280:
        code, n = _rev(sym - 256, 7), 7
    else:
        code, n = _rev(0xC0 + (sym - 280), 8), 8
    extra = length - LEN_BASE[idx]
    value = code | (extra << n)
    n += LEN_EXTRA[idx]
    return value, n + 5          # the distance code of distance 1 is 00000


def row_tokens(row):
    """Tokens of one filtered scanline (1-D uint8 array incl. the filter byte)."""
    toks = []
    n = len(row)
    starts = np.flatnonzero(np.concatenate(([True], row[1:] != row[:-1])))
    ends = np.concatenate((starts[1:], [n]))
    for s, e in zip(starts, ends):
        v = int(row[s])
        toks.append(literal_token(v))
        rest = int(e - s) - 1
        while rest >= 3:
            take = min(rest, 258)
            toks.append(match_token(take))
            rest -= take
        for _ in range(rest):
            toks.append(literal_token(v))
    return toks


def filtered_scanlines(img, swap_rb=True):
    """(H, 1 + W*ch) uint8: filter byte 0 + the pixels, channels reversed for 3-channel images
    (cv2.imwrite stores B,G,R arrays as R,G,B)."""
    img = np.asarray(img)
    assert img.dtype == np.uint8 and img.ndim in (2, 3)
    if img.ndim == 3 and img.shape[2] == 1:
        img = img[:, :, 0]
    if img.ndim == 3:
        assert img.shape[2] == 3, "1 or 3 channels"
        body = (img[:, :, ::-1] if swap_rb else img).reshape(img.shape[0], -1)
    else:
        body = img
    return np.concatenate((np.zeros((img.shape[0], 1), np.uint8), body), axis=1)


def deflate_fixed(rows):
    """bytes of the single fixed-Huffman deflate block over the filtered scanlines."""
    toks = [(3, 3)]                      # BFINAL = 1, BTYPE = 01 (LSB first: 1, 1, 0)
    for r in rows:
        toks.extend(row_tokens(r))
    toks.append((0, 7))                  # end of block (symbol 256 = 0000000)
    vals = np.array([t[0] for t in toks], dtype=np.uint64)
    lens = np.array([t[1] for t in toks], dtype=np.int64)
    pos = np.concatenate(([0], np.cumsum(lens)))
    total = int(pos[-1])
    bits = np.zeros((total + 7) // 8 * 8, dtype=np.uint8)
    for j in range(int(lens.max())):
        m = lens > j
        bits[pos[:-1][m] + j] = ((vals[m] >> np.uint64(j)) & np.uint64(1)).astype(np.uint8)
    return np.packbits(bits, bitorder="little").tobytes()


def chunk(kind, data):
    return struct.pack(">I", len(data)) + kind + data + struct.pack(">I", zlib.crc32(kind + data) & 0xffffffff)


def encode_png(img, swap_rb=True):
    """uint8 (H,W) or (H,W,3) -> PNG file bytes (see the module docstring for the exact stream)."""
    rows = filtered_scanlines(img, swap_rb)
    h = rows.shape[0]
    colour = np.asarray(img).ndim == 3 and np.asarray(img).shape[2] == 3
    w = (rows.shape[1] - 1) // (3 if colour else 1)
    ihdr = struct.pack(">IIBBBBB", w, h, 8, 2 if colour else 0, 0, 0, 0)
    z = b"\x78\x01" + deflate_fixed(rows) + struct.pack(">I", zlib.adler32(rows.tobytes()) & 0xffffffff)
    return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr) + chunk(b"IDAT", z) + chunk(b"IEND", b"")


def decode_png_zlib(png):
    """Independent decoder used by the tests beside cv2: parses the chunks (checking every CRC),
    inflates with zlib, undoes filter 0.  Returns (H, W, ch) uint8 in FILE channel order (R,G,B)."""
    assert png[:8] == b"\x89PNG\r\n\x1a\n"
    p, idat, hdr = 8, b"", None
    while p < len(png):
        n, = struct.unpack(">I", png[p:p + 4])
        kind, data = png[p + 4:p + 8], png[p + 8:p + 8 + n]
        crc, = struct.unpack(">I", png[p + 8 + n:p + 12 + n])
        assert crc == (zlib.crc32(kind + data) & 0xffffffff), "bad CRC in %r" % kind
        if kind == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", data)
        elif kind == b"IDAT":
            idat += data
        p += 12 + n
    w, h, depth, ctype = hdr[:4]
    assert depth == 8 and ctype in (0, 2)
    ch = 3 if ctype == 2 else 1
    raw = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, 1 + w * ch)
    assert not raw[:, 0].any(), "only filter type 0 is produced"
    return raw[:, 1:].reshape(h, w, ch)
