"""Minimal PNG -> RGBA8 decoder standing in for the browser half of loadTexture (main.js:375-395:
Image -> offscreen canvas drawImage -> getImageData).  Handles what a texture needs: non-interlaced
8-bit greyscale / RGB / palette / grey+alpha / RGBA, optional tRNS.  The reference's earth.png and
mars.png are 256x128 8-bit palette images with PLTE only, so `getImageData`
returns the palette colours with alpha 255 and no colour management applies.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

_SIG = b"\x89PNG\r\n\x1a\n"


def _unfilter(raw: bytes, h: int, stride: int, bpp: int) -> np.ndarray:
    out = np.zeros((h, stride), np.uint8)
    pos = 0
    prev = np.zeros(stride, np.int32)
    for y in range(h):
        ft = raw[pos]
        line = np.frombuffer(raw, np.uint8, stride, pos + 1).astype(np.int32)
        pos += 1 + stride
        if ft == 0:
            cur = line
        elif ft == 2:  # Up
            cur = (line + prev) & 255
        elif ft in (1, 3, 4):  # Sub / Average / Paeth need a left-to-right scan
            cur = np.zeros(stride, np.int32)
            for i in range(stride):
                a = cur[i - bpp] if i >= bpp else 0
                b = prev[i]
                if ft == 1:
                    p = a
                elif ft == 3:
                    p = (a + b) >> 1
                else:
                    c = prev[i - bpp] if i >= bpp else 0
                    pa, pb, pc = abs(b - c), abs(a - c), abs(a + b - 2 * c)
                    p = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                cur[i] = (line[i] + p) & 255
        else:
            raise ValueError(f"bad PNG filter type {ft}")
        out[y] = cur
        prev = cur
    return out


def decode_rgba8(src):
    """src: path or bytes.  Returns (width, height, uint8 array H x W x 4)."""
    data = src if isinstance(src, (bytes, bytearray)) else open(src, "rb").read()
    if data[:8] != _SIG:
        raise ValueError("not a PNG")
    pos = 8
    ihdr = plte = trns = None
    idat = []
    while pos < len(data):
        (ln,) = struct.unpack(">I", data[pos:pos + 4])
        typ = data[pos + 4:pos + 8]
        body = data[pos + 8:pos + 8 + ln]
        pos += 12 + ln
        if typ == b"IHDR":
            ihdr = struct.unpack(">IIBBBBB", body)
        elif typ == b"PLTE":
            plte = np.frombuffer(body, np.uint8).reshape(-1, 3)
        elif typ == b"tRNS":
            trns = np.frombuffer(body, np.uint8)
        elif typ == b"IDAT":
            idat.append(body)
        elif typ == b"IEND":
            break
    if ihdr is None:
        raise ValueError("PNG without IHDR")
    w, h, depth, ctype, _comp, _filt, interlace = ihdr
    if depth != 8 or interlace != 0:
        raise ValueError("only 8-bit non-interlaced PNGs are supported")
    channels = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}.get(ctype)
    if channels is None:
        raise ValueError(f"bad PNG colour type {ctype}")
    stride = w * channels
    px = _unfilter(zlib.decompress(b"".join(idat)), h, stride, channels).reshape(h, w, channels)
    rgba = np.empty((h, w, 4), np.uint8)
    rgba[..., 3] = 255
    if ctype == 3:
        if plte is None:
            raise ValueError("palette PNG without PLTE")
        idx = px[..., 0]
        rgba[..., :3] = plte[idx]
        if trns is not None:
            a = np.full(256, 255, np.uint8)
            a[:len(trns)] = trns
            rgba[..., 3] = a[idx]
    elif ctype == 0:
        rgba[..., :3] = px
    elif ctype == 2:
        rgba[..., :3] = px
    elif ctype == 4:
        rgba[..., :3] = px[..., :1]
        rgba[..., 3] = px[..., 1]
    else:
        rgba[...] = px
    return w, h, rgba
