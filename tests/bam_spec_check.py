"""An independent check of the BAM files the stage writes (test infrastructure).

It shares no code with anchored_fusion_b200/bam.py: the BGZF container is unpacked by Python's own gzip
module (every BGZF block must be a complete gzip member; the file must end in the 28-byte EOF block) and by
walking the BC extra fields, and the decompressed stream is checked field by field against the SAM/BAM
specification (SAMv1 section 4.2): magic, header text vs reference list, block sizes, name / CIGAR / sequence /
quality lengths, CIGAR consistency with l_seq, the UCSC bin of every record recomputed from POS and the CIGAR's
reference span, flag consistency, coordinate order.  Returns the parsed records for further assertions."""
import gzip
import struct

EOF_BLOCK = bytes([0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0])


def _reg2bin(beg, end):
    """SAMv1 5.3, written out from the specification's C code."""
    end -= 1
    if beg >> 14 == end >> 14:
        return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def check_bam(path, expect_sorted=True):
    blob = open(path, "rb").read()
    assert blob.endswith(EOF_BLOCK), "no BGZF EOF marker"
    # BGZF framing: a chain of blocks, each with the BC subfield and a consistent BSIZE
    off, n_blocks, total_isize = 0, 0, 0
    while off < len(blob):
        assert blob[off:off + 4] == b"\x1f\x8b\x08\x04", "block %d: not a gzip member with FEXTRA" % n_blocks
        xlen = struct.unpack_from("<H", blob, off + 10)[0]
        extra = blob[off + 12: off + 12 + xlen]
        bsize, p = None, 0
        while p + 4 <= len(extra):
            si1, si2, slen = extra[p], extra[p + 1], struct.unpack_from("<H", extra, p + 2)[0]
            if (si1, si2) == (66, 67):
                assert slen == 2
                bsize = struct.unpack_from("<H", extra, p + 4)[0] + 1
            p += 4 + slen
        assert bsize is not None and off + bsize <= len(blob), "block %d: BSIZE missing or beyond the file" % n_blocks
        isize = struct.unpack_from("<I", blob, off + bsize - 4)[0]
        assert isize <= 65536
        total_isize += isize
        off += bsize
        n_blocks += 1
    raw = gzip.decompress(blob)                      # stdlib gzip: every block a member, CRC-32 and ISIZE verified
    assert len(raw) == total_isize
    assert raw[:4] == b"BAM\x01", "bad magic"
    l_text = struct.unpack_from("<i", raw, 4)[0]
    text = raw[8:8 + l_text].decode()
    p = 8 + l_text
    n_ref = struct.unpack_from("<i", raw, p)[0]
    p += 4
    refs = []
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", raw, p)[0]
        name = raw[p + 4: p + 4 + l_name]
        assert name.endswith(b"\0") and b"\0" not in name[:-1]
        l_ref = struct.unpack_from("<i", raw, p + 4 + l_name)[0]
        refs.append((name[:-1].decode(), l_ref))
        p += 8 + l_name
    sq = [tuple(f.split(":", 1)[1] for f in line.split("\t")[1:3]) for line in text.splitlines() if line.startswith("@SQ")]
    assert [(n, int(ln)) for n, ln in sq] == refs, "@SQ lines and the binary reference list differ"
    hd = [line for line in text.splitlines() if line.startswith("@HD")]
    assert len(hd) == 1 and text.startswith("@HD")
    so = dict(f.split(":", 1) for f in hd[0].split("\t")[1:]).get("SO")
    records, last = [], (-1, -1)
    while p < len(raw):
        block_size = struct.unpack_from("<i", raw, p)[0]
        assert block_size >= 32 and p + 4 + block_size <= len(raw), "record block_size runs past the stream"
        ref_id, pos, l_read_name, mapq, bin_, n_cigar, flag, l_seq, next_ref, next_pos, tlen = struct.unpack_from("<iiBBHHHiiii", raw, p + 4)
        q = p + 36
        name = raw[q: q + l_read_name]
        assert l_read_name >= 2 and name.endswith(b"\0") and b"\0" not in name[:-1], "read name is not NUL-terminated"
        q += l_read_name
        cigar = [(v >> 4, "MIDNSHP=X"[v & 15]) for v in struct.unpack_from("<%dI" % n_cigar, raw, q)]
        q += 4 * n_cigar
        seq_bytes = raw[q: q + (l_seq + 1) // 2]
        q += (l_seq + 1) // 2
        qual = raw[q: q + l_seq]
        q += l_seq
        assert q <= p + 4 + block_size, "fixed fields + variable fields exceed block_size"
        assert -1 <= ref_id < n_ref and -1 <= next_ref < n_ref
        assert all(n > 0 for n, _ in cigar)
        if cigar:
            assert sum(n for n, op in cigar if op in "MIS=X") == l_seq, "CIGAR does not cover the sequence"
        ref_span = sum(n for n, op in cigar if op in "MDN=X")
        unmapped = bool(flag & 0x4)
        if unmapped:
            assert not cigar and mapq == 0
        else:
            assert cigar and ref_id >= 0 and 0 <= pos and pos + ref_span <= refs[ref_id][1], "alignment runs off the reference"
        assert bin_ == _reg2bin(pos, pos + (ref_span if ref_span else 1)), "bin field differs from reg2bin(POS, end)"
        assert l_seq == 0 or all(b != 0xFF for b in qual) or all(b == 0xFF for b in qual)
        assert all(b <= 93 or b == 0xFF for b in qual)
        if l_seq & 1:
            assert seq_bytes[-1] & 0xF == 0, "padding nibble of an odd-length sequence must be 0"
        if flag & 0x1:
            assert bool(flag & 0x40) != bool(flag & 0x80), "paired read must be first or last"
        if expect_sorted and so == "coordinate":
            key = (ref_id if ref_id >= 0 else 1 << 30, pos)
            assert key >= last, "records are not in coordinate order"
            last = key
        seq = "".join("=ACMGRSVTWYHKDBN"[b >> 4] + "=ACMGRSVTWYHKDBN"[b & 15] for b in seq_bytes)[:l_seq]
        records.append({"qname": name[:-1].decode(), "flag": flag, "ref_id": ref_id, "pos": pos + 1, "mapq": mapq,
                        "cigar": "".join("%d%s" % c for c in cigar) or "*", "next_pos": next_pos + 1, "seq": seq,
                        "qual": "".join(chr(b + 33) for b in qual) if qual and qual[0] != 0xFF else "*"})
        p += 4 + block_size
    assert p == len(raw)
    return {"blocks": n_blocks, "header": text, "refs": refs, "records": records}
