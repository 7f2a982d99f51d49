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
            else:
                raise ValueError(ty)
        recs.append(dict(name=name, flag=flag, ref_id=ref_id, pos=pos, mapq=mapq, cigar=cigar, seq=seq, qual=qual, tags=tags))
        p = end
    return text, refs, recs
