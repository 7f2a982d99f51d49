"""Minimal BAM reader for the tests (BGZF = concatenated gzip members; records per the SAM/BAM spec)."""
import gzip
import struct


def read_bam(path):
    data = gzip.open(path, "rb").read()
    assert data[:4] == b"BAM\x01"
    l_text, = struct.unpack_from("<i", data, 4)
    text = data[8:8 + l_text].decode()
    p = 8 + l_text
    n_ref, = struct.unpack_from("<i", data, p); p += 4
    refs = []
    for _ in range(n_ref):
        l_name, = struct.unpack_from("<i", data, p); p += 4
        name = data[p:p + l_name - 1].decode(); p += l_name
        l_ref, = struct.unpack_from("<i", data, p); p += 4
        refs.append((name, l_ref))
    recs = []
    while p < len(data):
        bs, = struct.unpack_from("<i", data, p); p += 4
        ref_id, pos, l_rn, mapq, bin_, n_cig, flag, l_seq, nref, npos, tlen = struct.unpack_from("<iiBBHHHiiii", data, p)
        q = p + 32
        name = data[q:q + l_rn - 1].decode(); q += l_rn
        cigar = "".join("%d%s" % (v >> 4, "MIDNSHP=X"[v & 15]) for v in struct.unpack_from("<%dI" % n_cig, data, q)); q += 4 * n_cig
        sb = data[q:q + (l_seq + 1) // 2]; q += (l_seq + 1) // 2
        seq = "".join("=ACMGRSVTWYHKDBN"[(sb[i // 2] >> (4 if i % 2 == 0 else 0)) & 15] for i in range(l_seq))
        qual = bytes(data[q:q + l_seq]); q += l_seq
        tags = {}
        end = p + bs
        while q < end:
            tag = data[q:q + 2].decode(); ty = chr(data[q + 2]); q += 3
            if ty == "Z":
                e = data.index(b"\x00", q); tags[tag] = data[q:e].decode(); q = e + 1
            elif ty == "i":
                tags[tag], = struct.unpack_from("<i", data, q); q += 4
            elif ty == "f":
                tags[tag], = struct.unpack_from("<f", data, q); q += 4
            elif ty == "A":
                tags[tag] = chr(data[q]); q += 1
            elif ty in "cCsSI":
                fmt = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "I": "<I"}[ty]
                tags[tag], = struct.unpack_from(fmt, data, q); q += struct.calcsize(fmt)
            elif ty == "H":
                e = data.index(b"\x00", q); tags[tag] = ("H", data[q:e].decode()); q = e + 1
            elif ty == "B":
                sub = chr(data[q]); cnt, = struct.unpack_from("<I", data, q + 1); q += 5
                fmt = "<%d%s" % (cnt, {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I", "f": "f"}[sub])
                tags[tag] = (sub, list(struct.unpack_from(fmt, data, q))); q += struct.calcsize(fmt)
            else:
                raise ValueError(ty)
        tag_order = list(tags)
        recs.append(dict(tag_order=tag_order, name=name, flag=flag, ref_id=ref_id, pos=pos, mapq=mapq, cigar=cigar, seq=seq, qual=qual, tags=tags))
        p = end
    return text, refs, recs


def bgzf_block(payload):
    """One BGZF block (SAM spec 4.1): gzip member with the BC extra field."""
    import zlib
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = co.compress(payload) + co.flush()
    bsize = len(body) + 25
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize) + body +
            struct.pack("<II", zlib.crc32(payload) & 0xFFFFFFFF, len(payload)))


def write_bam(path, header_text, refs, records, block=40_000):
    """records: dicts with name, flag, seq, qual (bytes of raw Phred or None), aux (raw bytes); all unplaced."""
    out = bytearray(b"BAM\x01" + struct.pack("<i", len(header_text)) + header_text.encode() + struct.pack("<i", len(refs)))
    for name, ln in refs:
        out += struct.pack("<i", len(name) + 1) + name.encode() + b"\x00" + struct.pack("<i", ln)
    for r in records:
        name = r["name"].encode() + b"\x00"
        seq = r["seq"]
        codes = ["=ACMGRSVTWYHKDBN".index(c) for c in seq]
        if len(codes) % 2:
            codes.append(0)
        packed = bytes((codes[i] << 4) | codes[i + 1] for i in range(0, len(codes), 2))
        qual = r["qual"] if r["qual"] is not None else b"\xff" * len(seq)
        body = struct.pack("<iiBBHHHiiii", -1, -1, len(name), 0, 4680, 0, r["flag"], len(seq), -1, -1, 0) + name + packed + qual + r.get("aux", b"")
        out += struct.pack("<i", len(body)) + body
    with open(path, "wb") as f:
        for o in range(0, len(out), block):
            f.write(bgzf_block(bytes(out[o:o + block])))
        f.write(bgzf_block(b""))
