"""The reference's seven on-disk index files: Snappy frame streams around bincode `Item{version=5, data}`
(/root/reference/src/index/indexing.rs:111-207, src/index/versioned_index.rs).  Host-only: no GPU needed.
The files are decoded here by an independent Python reader of the two published formats (Snappy framing, bincode 1.x
fixed-int little-endian) and re-encoded with *compressed* chunks (pyarrow's raw Snappy codec) to exercise the decoder."""
import os
import struct

import numpy as np
import pytest

from mapad_b200 import api
from helpers import random_genome

SUFFIXES = ["tbw", "tle", "toc", "trt", "tsa", "tpi", "tos"]


def _crc32c_table():
    t = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        t.append(c)
    return t


_T = _crc32c_table()


def crc32c(b):
    c = 0xFFFFFFFF
    for x in b:
        c = _T[(c ^ x) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked(c):
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def unframe(raw):
    assert raw[:10] == b"\xff\x06\x00\x00sNaPpY"
    out, p = bytearray(), 10
    while p < len(raw):
        typ = raw[p]
        ln = raw[p + 1] | (raw[p + 2] << 8) | (raw[p + 3] << 16)
        body = raw[p + 4:p + 4 + ln]
        p += 4 + ln
        assert typ == 0x01, "writer emits uncompressed chunks"
        crc = struct.unpack("<I", body[:4])[0]
        assert len(body) - 4 <= 65536
        assert masked(crc32c(body[4:])) == crc
        out += body[4:]
    return bytes(out)


def frame(payload, compress, chunk=65536):
    import pyarrow as pa
    out = bytearray(b"\xff\x06\x00\x00sNaPpY")
    for off in range(0, len(payload), chunk):
        d = payload[off:off + chunk]
        crc = struct.pack("<I", masked(crc32c(d)))
        if compress:
            body = crc + pa.compress(d, codec="snappy", asbytes=True)
            typ = 0x00
        else:
            body = crc + d
            typ = 0x01
        out += bytes([typ]) + struct.pack("<I", len(body))[:3] + body
        if off == 0:
            out += b"\xfe\x03\x00\x00pad"  # a padding chunk must be skipped
    return bytes(out)


@pytest.fixture(scope="module")
def small_index(tmp_path_factory):
    rng = np.random.default_rng(5)
    g1 = random_genome(70_000, rng)
    g2 = random_genome(9_000, rng)
    g2 = g2[:3000] + "N" * 40 + g2[3040:5000] + "RYN" + g2[5003:]
    ix = api.Index.build([("chrA", g1), ("chrB extra words", g2)], seed=7)
    d = tmp_path_factory.mktemp("idx")
    prefix = str(d / "ref.fa")
    ix.save(prefix)
    return ix, prefix


def _same(a, b):
    assert a["n"] == b["n"] and a["less"] == b["less"] and a["sentinel_rows"] == b["sentinel_rows"]
    assert a["sa_rate"] == b["sa_rate"] and a["contigs"] == b["contigs"]
    for k in ("bwt", "sa_sample", "extra_rows", "orig_pos", "orig_sym"):
        assert np.array_equal(a[k], b[k]), k


def test_roundtrip(small_index):
    ix, prefix = small_index
    for s in SUFFIXES:
        assert os.path.getsize(prefix + "." + s) > 10
    _same(ix.arrays(), api.Index.load(prefix).arrays())


def test_file_contents_independent_decoder(small_index):
    ix, prefix = small_index
    a = ix.arrays()
    n = a["n"]
    u64 = lambda b, o: struct.unpack_from("<Q", b, o)[0]

    b = unframe(open(prefix + ".tbw", "rb").read())
    assert b[0] == 5 and u64(b, 1) == n and np.array_equal(np.frombuffer(b, np.uint8, n, 9), a["bwt"]) and len(b) == 9 + n

    b = unframe(open(prefix + ".tle", "rb").read())
    assert b[0] == 5 and u64(b, 1) == 7 and len(b) == 9 + 56
    less = np.frombuffer(b, "<u8", 7, 9)
    cnt = np.bincount(a["bwt"], minlength=6)
    assert np.array_equal(less, np.concatenate([[0], np.cumsum(cnt)]))  # bio::data_structures::bwt::less

    b = unframe(open(prefix + ".toc", "rb").read())
    assert b[0] == 5 and u64(b, 1) == 6
    o = 9
    rows = (n - 1) // 128 + 1
    for c in range(6):  # Occ::new(bwt, 128, alphabet): occ[c][i] = #c in bwt[..=128 i]
        assert u64(b, o) == rows
        col = np.frombuffer(b, "<u8", rows, o + 8)
        want = np.cumsum(a["bwt"] == c)[::128]
        assert np.array_equal(col, want), c
        o += 8 + 8 * rows
    assert struct.unpack_from("<I", b, o)[0] == 128 and len(b) == o + 4

    b = unframe(open(prefix + ".trt", "rb").read())
    assert b[0] == 5 and u64(b, 1) == 6 and len(b) == 9 + 6 * 9
    assert [(u64(b, 9 + 9 * i), b[17 + 9 * i]) for i in range(6)] == [(ord(c), i) for i, c in enumerate("$ACGTX")]

    b = unframe(open(prefix + ".tsa", "rb").read())
    ns = u64(b, 1)
    assert b[0] == 5 and ns == (n + 31) // 32 and np.array_equal(np.frombuffer(b, "<u8", ns, 9), a["sa_sample"])
    o = 9 + 8 * ns
    assert u64(b, o) == 32
    ne = u64(b, o + 8)
    assert ne == len(a["extra_rows"]) and np.array_equal(np.frombuffer(b, "<u8", 2 * ne, o + 16).reshape(-1, 2), a["extra_rows"])
    assert b[o + 16 + 16 * ne] == 0 and len(b) == o + 17 + 16 * ne

    b = unframe(open(prefix + ".tpi", "rb").read())
    assert b[0] == 5 and u64(b, 1) == 2
    o, got = 9, []
    for _ in range(2):
        s, e, l = u64(b, o), u64(b, o + 8), u64(b, o + 16)
        got.append((b[o + 24:o + 24 + l].decode(), s, e))
        o += 24 + l
    assert got == a["contigs"] and len(b) == o

    b = unframe(open(prefix + ".tos", "rb").read())
    no = u64(b, 1)
    assert b[0] == 5 and no == len(a["orig_pos"]) == 3 and len(b) == 9 + 9 * no
    assert [(u64(b, 9 + 9 * i), b[17 + 9 * i]) for i in range(no)] == list(zip(a["orig_pos"].tolist(), a["orig_sym"].tolist()))
    assert bytes(a["orig_sym"]) == b"RYN"


def test_reads_compressed_chunks_and_padding(small_index, tmp_path):
    pytest.importorskip("pyarrow")
    ix, prefix = small_index
    dst = str(tmp_path / "c.fa")
    for s in SUFFIXES:
        payload = unframe(open(prefix + "." + s, "rb").read())
        open(dst + "." + s, "wb").write(frame(payload, compress=True, chunk=40_000))
    assert os.path.getsize(dst + ".tsa") != os.path.getsize(prefix + ".tsa")
    _same(ix.arrays(), api.Index.load(dst).arrays())


def test_version_mismatch_and_corruption(small_index, tmp_path):
    ix, prefix = small_index
    import shutil

    def clone(name):
        dst = str(tmp_path / name)
        for s in SUFFIXES:
            shutil.copy(prefix + "." + s, dst + "." + s)
        return dst

    d = clone("v")
    payload = bytearray(unframe(open(d + ".tsa", "rb").read()))
    payload[0] = 4
    open(d + ".tsa", "wb").write(frame(bytes(payload), compress=False))
    with pytest.raises(api.MapadError) as e:
        api.Index.load(d)
    assert e.value.code == -5  # MAPAD_EINDEX <- Error::IndexVersionMismatch

    d = clone("crc")
    raw = bytearray(open(d + ".tbw", "rb").read())
    raw[5000] ^= 1
    open(d + ".tbw", "wb").write(raw)
    with pytest.raises(api.MapadError) as e:
        api.Index.load(d)
    assert e.value.code == -6  # MAPAD_EIO

    d = clone("less")
    payload = bytearray(unframe(open(d + ".tle", "rb").read()))
    payload[9 + 8 * 2] ^= 1
    open(d + ".tle", "wb").write(frame(bytes(payload), compress=False))
    with pytest.raises(api.MapadError) as e:
        api.Index.load(d)
    assert e.value.code == -5

    with pytest.raises(api.MapadError) as e:
        api.Index.load(str(tmp_path / "missing"))
    assert e.value.code == -6
