"""TEST INFRASTRUCTURE — hand-assembles small HDF5 files, byte by byte from the
published format specification, that use format features the library's own writer
never emits (superblock v2, version-2 object headers with link messages, chunked
layout with a B-tree v1 chunk index, deflate + shuffle filters, big-endian and 64-bit
integer types, compact layout).  They widen the reader's coverage beyond the
libhdf5-written fixture in tests/golden/; checksums of the v2 structures are written
as zero (the reader does not verify them)."""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


def _dtype_msg(dt: np.dtype) -> bytes:
    be = dt.byteorder == ">"
    if dt.kind == "f":
        d = dt.itemsize == 8
        return (bytes([0x11, 0x20 | (1 if be else 0), 63 if d else 31, 0]) + struct.pack("<I", dt.itemsize)
                + struct.pack("<HHBBBBI", 0, dt.itemsize * 8, 52 if d else 23, 11 if d else 8, 0,
                              52 if d else 23, 1023 if d else 127))
    signed = dt.kind == "i"
    return (bytes([0x10, (8 if signed else 0) | (1 if be else 0), 0, 0]) + struct.pack("<I", dt.itemsize)
            + struct.pack("<HH", 0, dt.itemsize * 8))


def _space_msg(n: int) -> bytes:  # dataspace v2, simple, rank 1
    return bytes([2, 1, 0, 1]) + struct.pack("<Q", n)


def _ohdr_v2(msgs) -> bytes:
    body = b"".join(bytes([t]) + struct.pack("<H", len(b)) + bytes([0]) + b for t, b in msgs)
    assert len(body) < 65536
    return b"OHDR" + bytes([2, 0x01]) + struct.pack("<H", len(body) + 4) + body + b"\0\0\0\0"


class Crafter:
    """superblock v2 file; datasets are linked from the root group by link messages"""

    def __init__(self):
        self.blob = bytearray(48)  # superblock v2 placeholder
        self.links = []

    def _append(self, b: bytes, align=8) -> int:
        while len(self.blob) % align:
            self.blob.append(0)
        a = len(self.blob)
        self.blob += b
        return a

    def _finish_dataset(self, name, arr, layout_msg, extra=()):
        msgs = [(0x01, _space_msg(arr.size)), (0x03, _dtype_msg(arr.dtype)), *extra, (0x08, layout_msg)]
        self.links.append((name, self._append(_ohdr_v2(msgs))))

    def contiguous(self, name, arr):
        a = self._append(arr.tobytes())
        self._finish_dataset(name, arr, bytes([3, 1]) + struct.pack("<QQ", a, arr.nbytes))

    def compact(self, name, arr):
        raw = arr.tobytes()
        self._finish_dataset(name, arr, bytes([3, 0]) + struct.pack("<H", len(raw)) + raw)

    def chunked(self, name, arr, chunk, deflate=False, shuffle=False, skip_chunks=()):
        es = arr.dtype.itemsize
        recs = []
        for k, c0 in enumerate(range(0, arr.size, chunk)):
            if k in skip_chunks:
                continue  # never-written chunk: reads as the fill value (0)
            piece = np.zeros(chunk, arr.dtype)
            seg = arr[c0:c0 + chunk]
            piece[:seg.size] = seg
            raw = piece.tobytes()
            if shuffle:
                raw = np.frombuffer(raw, np.uint8).reshape(chunk, es).T.tobytes()
            if deflate:
                raw = zlib.compress(raw, 6)
            recs.append((c0, self._append(raw), len(raw)))
        # B-tree v1, node type 1 (raw data chunks), leaves of <= 4 entries + one level-1 root
        def node(level, entries, last_key_offset):
            out = b"TREE" + bytes([1, level]) + struct.pack("<HQQ", len(entries), UNDEF, UNDEF)
            for off, addr, size in entries:
                out += struct.pack("<IIQQ", size, 0, off, 0) + struct.pack("<Q", addr)
            out += struct.pack("<IIQQ", 0, 0, last_key_offset, 0)
            return out
        end = (arr.size + chunk - 1) // chunk * chunk
        if not recs:
            root = UNDEF
        else:
            leaves = []
            for i in range(0, len(recs), 4):
                grp = recs[i:i + 4]
                leaves.append((grp[0][0], self._append(node(0, grp, grp[-1][0] + chunk)), 0))
            root = self._append(node(1, leaves, end)) if len(leaves) > 1 else leaves[0][1]
        extra = []
        filt = []
        if shuffle:
            filt.append(struct.pack("<HHH", 2, 0, 1) + struct.pack("<I", es))
        if deflate:
            filt.append(struct.pack("<HHH", 1, 0, 1) + struct.pack("<I", 6))
        if filt:
            extra.append((0x0B, bytes([2, len(filt)]) + b"".join(filt)))
        layout = bytes([3, 2, 2]) + struct.pack("<Q", root) + struct.pack("<II", chunk, es)
        self._finish_dataset(name, arr, layout, extra)

    def save(self, path):
        msgs = []
        for name, addr in self.links:
            nb = name.encode()
            msgs.append((0x06, bytes([1, 0, len(nb)]) + nb + struct.pack("<Q", addr)))
        root = self._append(_ohdr_v2(msgs))
        sb = (b"\x89HDF\r\n\x1a\n" + bytes([2, 8, 8, 0])
              + struct.pack("<QQQQ", 0, UNDEF, len(self.blob), root) + b"\0\0\0\0")
        self.blob[:48] = sb
        with open(path, "wb") as f:
            f.write(self.blob)
