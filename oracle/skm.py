"""TEST INFRASTRUCTURE ONLY: pure-Python reader for `.skm` / `.skd` sketch databases.

Restates MultiSketch::load_metadata / read_sketch_data / read_sketch_data_block
(reference src/sketch/multisketch.rs:90-103, 167-210; src/sketch/sketch_datafile.rs:159-194).
The `.skm` container is a snappy *frame* stream (crate `snap` 1.1) around one CBOR map
(crate `ciborium` 0.2, serde field names); neither crate is vendored in the reference, so
the decoders below follow the published formats (Google snappy framing_format.txt,
RFC 8949) and are pinned by the four `.skm` fixtures the reference ships.
The product has its own C++ reader (sketchlib.rust_b200/csrc/host/skm_reader.cpp); this
file is only used to check it and to drive the oracle.
"""
import struct

import numpy as np

BBITS = 14


def _snappy_block(buf: bytes) -> bytes:
    pos, n, shift = 0, 0, 0
    while True:  # uvarint uncompressed length
        b = buf[pos]
        pos += 1
        n |= (b & 0x7F) << shift
        shift += 7
        if not b & 0x80:
            break
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:  # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 2], "little")
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        for _ in range(ln):  # may overlap
            out.append(out[-off])
    assert len(out) == n, "snappy: bad uncompressed length"
    return bytes(out)


def snappy_frame_decode(raw: bytes) -> bytes:
    pos, out = 0, bytearray()
    while pos < len(raw):
        ctype = raw[pos]
        ln = int.from_bytes(raw[pos + 1:pos + 4], "little")
        body = raw[pos + 4:pos + 4 + ln]
        pos += 4 + ln
        if ctype == 0xFF:
            assert body == b"sNaPpY"
        elif ctype == 0x00:
            out += _snappy_block(body[4:])  # 4-byte masked CRC32C first
        elif ctype == 0x01:
            out += body[4:]
        elif 0x80 <= ctype <= 0xFE:
            continue
        else:
            raise ValueError(f"snappy frame: reserved unskippable chunk {ctype:#x}")
    return bytes(out)


def cbor_decode(buf: bytes):
    def item(pos):
        ib = buf[pos]
        pos += 1
        major, info = ib >> 5, ib & 31
        if info < 24:
            val = info
        elif info in (24, 25, 26, 27):
            nb = 1 << (info - 24)
            val = int.from_bytes(buf[pos:pos + nb], "big")
            if major == 7:
                raw = buf[pos:pos + nb]
            pos += nb
        elif info == 31:
            val = None  # indefinite
        else:
            raise ValueError("cbor: bad additional info")
        if major == 0:
            return val, pos
        if major == 1:
            return -1 - val, pos
        if major in (2, 3):
            if val is None:
                chunks = []
                while buf[pos] != 0xFF:
                    c, pos = item(pos)
                    chunks.append(c)
                pos += 1
                return (b"".join(chunks) if major == 2 else "".join(chunks)), pos
            data = buf[pos:pos + val]
            pos += val
            return (bytes(data) if major == 2 else data.decode("utf-8")), pos
        if major == 4:
            arr = []
            if val is None:
                while buf[pos] != 0xFF:
                    v, pos = item(pos)
                    arr.append(v)
                return arr, pos + 1
            for _ in range(val):
                v, pos = item(pos)
                arr.append(v)
            return arr, pos
        if major == 5:
            d = {}
            if val is None:
                while buf[pos] != 0xFF:
                    k, pos = item(pos)
                    v, pos = item(pos)
                    d[k] = v
                return d, pos + 1
            for _ in range(val):
                k, pos = item(pos)
                v, pos = item(pos)
                d[k] = v
            return d, pos
        if major == 6:
            return item(pos)  # ignore tags
        # major 7
        if info == 20:
            return False, pos
        if info == 21:
            return True, pos
        if info in (22, 23):
            return None, pos
        if info == 25:
            return float(np.frombuffer(raw, dtype=">f2")[0]), pos
        if info == 26:
            return struct.unpack(">f", raw)[0], pos
        if info == 27:
            return struct.unpack(">d", raw)[0], pos
        raise ValueError("cbor: unsupported simple value")

    val, _ = item(0)
    return val


class MultiSketch:
    """Mirror of the reference's MultiSketch as the distance path uses it."""

    def __init__(self, prefix: str):
        for ext in (".skm", ".skd"):  # strip_sketch_extension, src/utils.rs:9-15
            if prefix.endswith(ext):
                prefix = prefix[: -len(ext)]
        self.prefix = prefix
        with open(prefix + ".skm", "rb") as fh:
            meta = cbor_decode(snappy_frame_decode(fh.read()))
        self.meta = meta
        self.sketch_size = int(meta["sketch_size"])
        self.sketchsize64 = int(meta.get("sketchsize64", 0))
        if self.sketchsize64 == 0:  # legacy fix-up, multisketch.rs:97-100
            self.sketchsize64 = self.sketch_size
            self.sketch_size *= 64
        self.kmer_lengths = [int(k) for k in meta["kmer_lengths"]]
        self.kmer_stride = int(meta["kmer_stride"])
        self.sample_stride = int(meta["sample_stride"])
        self.sketch_metadata = meta["sketch_metadata"]
        self.name_map = meta["name_map"]
        self.block_reindex = None
        self.bins = None

    def read_sketch_data(self):  # multisketch.rs:167-184
        data = np.fromfile(self.prefix + ".skd", dtype="<u8")
        self.bins = np.ascontiguousarray(data[: self.sample_stride * len(self.sketch_metadata)])
        return self

    def read_sketch_data_block(self, names):  # multisketch.rs:187-210
        data = np.memmap(self.prefix + ".skd", dtype="<u8", mode="r")
        reindex, rows = [], []
        for name in names:
            if name not in self.name_map:
                raise KeyError(f"Could not find requested sample {name} in sketch metadata")
            sidx = self.name_map[name]
            reindex.append(sidx)
            fidx = self.sketch_metadata[sidx]["index"]
            rows.append(np.array(data[fidx * self.sample_stride:(fidx + 1) * self.sample_stride]))
        self.block_reindex = reindex
        self.bins = np.ascontiguousarray(np.concatenate(rows))
        return self

    def number_samples_loaded(self):
        return len(self.block_reindex) if self.block_reindex is not None else len(self.sketch_metadata)

    def sketch_name(self, idx):
        if self.block_reindex is not None:
            return self.sketch_metadata[self.block_reindex[idx]]["name"]
        return self.sketch_metadata[idx]["name"]

    def names(self):
        return [self.sketch_name(i) for i in range(self.number_samples_loaded())]
