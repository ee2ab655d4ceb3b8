"""Minimal pure-Python FLAC decoder (test infrastructure): enough of the format for the reference's bundled
SEAME utterance (16 kHz mono 16-bit; CONSTANT / VERBATIM / FIXED / LPC subframes, partitioned Rice residuals).
No soundfile / ffmpeg / librosa exists in this environment."""
from __future__ import annotations

import numpy as np


class _Bits:
    def __init__(self, data: bytes, pos: int = 0):
        self.d, self.p, self.acc, self.n = data, pos, 0, 0

    def read(self, k: int) -> int:
        while self.n < k:
            self.acc = (self.acc << 8) | self.d[self.p]
            self.p += 1
            self.n += 8
        self.n -= k
        v = (self.acc >> self.n) & ((1 << k) - 1)
        self.acc &= (1 << self.n) - 1
        return v

    def read_signed(self, k: int) -> int:
        v = self.read(k)
        return v - (1 << k) if v >> (k - 1) else v

    def unary(self) -> int:
        c = 0
        while self.read(1) == 0:
            c += 1
        return c

    def align(self):
        self.n -= self.n % 8
        self.acc &= (1 << self.n) - 1

    def byte_pos(self) -> int:
        return self.p - self.n // 8


def _residual(b: _Bits, blocksize: int, order: int):
    method = b.read(2)
    pbits = 4 if method == 0 else 5
    esc = (1 << pbits) - 1
    porder = b.read(4)
    nparts = 1 << porder
    out = []
    for part in range(nparts):
        n = (blocksize >> porder) - (order if part == 0 else 0)
        k = b.read(pbits)
        if k == esc:
            raw = b.read(5)
            out.extend(b.read_signed(raw) if raw else 0 for _ in range(n))
        else:
            for _ in range(n):
                q = b.unary()
                v = (q << k) | (b.read(k) if k else 0)
                out.append((v >> 1) ^ -(v & 1))
    return out


_FIXED = {0: [], 1: [1], 2: [2, -1], 3: [3, -3, 1], 4: [4, -6, 4, -1]}


def _subframe(b: _Bits, blocksize: int, bps: int):
    assert b.read(1) == 0
    typ = b.read(6)
    wasted = 0
    if b.read(1):
        wasted = b.unary() + 1
        bps -= wasted
    if typ == 0:
        s = [b.read_signed(bps)] * blocksize
    elif typ == 1:
        s = [b.read_signed(bps) for _ in range(blocksize)]
    elif 8 <= typ <= 12:
        order = typ - 8
        s = [b.read_signed(bps) for _ in range(order)]
        res = _residual(b, blocksize, order)
        c = _FIXED[order]
        for r in res:
            s.append(r + sum(ci * s[-1 - i] for i, ci in enumerate(c)))
    elif typ >= 32:
        order = (typ & 31) + 1
        s = [b.read_signed(bps) for _ in range(order)]
        prec = b.read(4) + 1
        shift = b.read_signed(5)
        coef = [b.read_signed(prec) for _ in range(order)]
        res = _residual(b, blocksize, order)
        for r in res:
            s.append(r + (sum(ci * s[-1 - i] for i, ci in enumerate(coef)) >> shift))
    else:
        raise ValueError(f"reserved subframe type {typ}")
    return [x << wasted for x in s] if wasted else s


def decode_flac(path: str):
    data = open(path, "rb").read()
    assert data[:4] == b"fLaC"
    pos = 4
    sr = ch = bps = total = None
    while True:
        last, typ = data[pos] >> 7, data[pos] & 0x7F
        ln = int.from_bytes(data[pos + 1: pos + 4], "big")
        body = data[pos + 4: pos + 4 + ln]
        if typ == 0:
            x = int.from_bytes(body[10:18], "big")
            sr, ch, bps, total = x >> 44, ((x >> 41) & 7) + 1, ((x >> 36) & 31) + 1, x & ((1 << 36) - 1)
        pos += 4 + ln
        if last:
            break
    chans = [[] for _ in range(ch)]
    while pos < len(data) - 2 and sum(len(c) for c in chans) // ch < total:
        b = _Bits(data, pos)
        assert b.read(14) == 0x3FFE, "lost frame sync"
        b.read(1)
        b.read(1)
        bs_code, sr_code, ch_code, ss_code = b.read(4), b.read(4), b.read(4), b.read(3)
        b.read(1)
        first = b.read(8)  # UTF-8 coded frame/sample number
        extra = 0
        while first & 0x80 and first & (0x40 >> extra):
            extra += 1
        if first & 0x80:
            for _ in range(extra):
                b.read(8)
        if bs_code == 1:
            blocksize = 192
        elif 2 <= bs_code <= 5:
            blocksize = 576 << (bs_code - 2)
        elif bs_code == 6:
            blocksize = b.read(8) + 1
        elif bs_code == 7:
            blocksize = b.read(16) + 1
        else:
            blocksize = 256 << (bs_code - 8)
        if sr_code == 12:
            b.read(8)
        elif sr_code in (13, 14):
            b.read(16)
        b.read(8)  # header CRC-8
        fbps = {0: bps, 1: 8, 2: 12, 4: 16, 5: 20, 6: 24}[ss_code]
        if ch_code < 8:
            subs = [_subframe(b, blocksize, fbps) for _ in range(ch_code + 1)]
        else:  # stereo decorrelation
            side_ch = {8: 1, 9: 0, 10: 1}[ch_code]
            subs = [_subframe(b, blocksize, fbps + (1 if i == side_ch else 0)) for i in range(2)]
            a0, a1 = np.array(subs[0], dtype=np.int64), np.array(subs[1], dtype=np.int64)
            if ch_code == 8:
                subs = [a0, a0 - a1]
            elif ch_code == 9:
                subs = [a0 + a1, a1]
            else:
                mid = (a0 << 1) | (a1 & 1)
                subs = [(mid + a1) >> 1, (mid - a1) >> 1]
        b.align()
        b.read(16)  # frame CRC-16
        for c, s in zip(chans, subs):
            c.extend(int(v) for v in s)
        pos = b.byte_pos()
    pcm = np.array(chans, dtype=np.int64)[:, :total]
    return pcm, sr, bps


if __name__ == "__main__":
    import sys

    pcm, sr, bps = decode_flac(sys.argv[1])
    print(pcm.shape, sr, bps, pcm.min(), pcm.max(), float(np.abs(pcm).mean()))
