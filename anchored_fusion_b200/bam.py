"""Minimal BAM (BGZF) writer and reader for the anchoring stage's output files.

The reference's stage hands over BAM files written by samtools (Anchored_Fusion.py:182,194)
and every later stage opens them with `samtools view` (functions.py:708).  samtools is not
needed to PRODUCE them: this module writes spec-conformant BAM so the files are a drop-in.
The reader exists for the tests (round trip) and for environments without samtools.
"""
import struct
import zlib

import numpy as np

_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
_CIGAR_OPS = "MIDNSHP=X"
_SEQ_CODE = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
_SEQ_DECODE = "=ACMGRSVTWYHKDBN"


def _reg2bin(beg, end):
    end -= 1
    for shift, offset in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return offset + (beg >> shift)
    return 0


_SEQ_TABLE = bytes(_SEQ_CODE.get(chr(b).upper(), 15) for b in range(256))      # ASCII base -> 4-bit BAM code
_QUAL_TABLE = bytes((b - 33) & 0xFF for b in range(256))                        # phred+33 -> phred


class BamWriter:
    def __init__(self, path, ref_name, ref_len, sort_order="coordinate", program="anchored_fusion_b200"):
        self._fh = open(path, "wb")
        self._buf = bytearray()
        text = "@HD\tVN:1.6\tSO:%s\n@SQ\tSN:%s\tLN:%d\n@PG\tID:%s\tPN:%s\n" % (sort_order, ref_name, ref_len, program, program)
        tb, nb = text.encode(), ref_name.encode() + b"\0"
        self._put(b"BAM\1" + struct.pack("<i", len(tb)) + tb + struct.pack("<i", 1) +
                  struct.pack("<i", len(nb)) + nb + struct.pack("<i", ref_len))

    def _put(self, data):
        self._buf += data
        while len(self._buf) >= 0xFF00:
            self._flush_block(bytes(self._buf[:0xFF00]))
            del self._buf[:0xFF00]

    def _flush_block(self, data):
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        comp = co.compress(data) + co.flush()
        bsize = len(comp) + 25
        self._fh.write(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", bsize) + comp +
                       struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))

    def write(self, qname, flag, pos, mapq, cigar, seq, qual, next_pos=None, mapped=True):
        """pos / next_pos are 1-based (SAM convention); cigar is [(len, op_char), ...]; qual is the
        FASTQ quality string (phred+33) or None."""
        name = qname.encode() + b"\0"
        ref_span = sum(n for n, op in cigar if op in "MDN=X") if mapped else 1
        pos0 = pos - 1
        nxt = (next_pos - 1) if next_pos is not None else pos0
        ops = b"".join(struct.pack("<I", (n << 4) | _CIGAR_OPS.index(op)) for n, op in cigar)
        codes = np.frombuffer(seq.encode().translate(_SEQ_TABLE) + (b"\0" if len(seq) & 1 else b""), dtype=np.uint8)
        packed = ((codes[0::2] << 4) | codes[1::2]).tobytes()
        q = qual.encode().translate(_QUAL_TABLE) if qual is not None else b"\xff" * len(seq)
        body = struct.pack("<iiBBHHHIiii", 0, pos0, len(name), mapq, _reg2bin(pos0, pos0 + max(ref_span, 1)),
                           len(cigar), flag, len(seq), 0, nxt, 0) + name + ops + packed + q
        self._put(struct.pack("<i", len(body)) + body)

    def close(self):
        if self._buf:
            self._flush_block(bytes(self._buf))
            self._buf = bytearray()
        self._fh.write(_EOF)
        self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def read_bam(path):
    """Returns (header_text, [(ref_name, ref_len)], [record dict...]) -- enough of BAM for tests and
    for `samtools view`-less environments (fields: qname flag rname pos mapq cigar pnext seq qual)."""
    raw = bytearray()
    with open(path, "rb") as fh:
        data = fh.read()
    off = 0
    while off < len(data):
        xlen = struct.unpack_from("<H", data, off + 10)[0]
        bsize = struct.unpack_from("<H", data, off + 16)[0] + 1
        comp = data[off + 12 + xlen: off + bsize - 8]
        raw += zlib.decompress(comp, -15) if comp else b""
        off += bsize
    assert raw[:4] == b"BAM\1"
    l_text = struct.unpack_from("<i", raw, 4)[0]
    text = raw[8: 8 + l_text].decode()
    p = 8 + l_text
    n_ref = struct.unpack_from("<i", raw, p)[0]
    p += 4
    refs = []
    for _ in range(n_ref):
        ln = struct.unpack_from("<i", raw, p)[0]
        name = raw[p + 4: p + 4 + ln - 1].decode()
        refs.append((name, struct.unpack_from("<i", raw, p + 4 + ln)[0]))
        p += 8 + ln
    recs = []
    while p < len(raw):
        size = struct.unpack_from("<i", raw, p)[0]
        refid, pos0, l_name, mapq, _bin, n_cig, flag, l_seq, _nref, nxt, _tlen = struct.unpack_from("<iiBBHHHIiii", raw, p + 4)
        q = p + 36
        qname = raw[q: q + l_name - 1].decode()
        q += l_name
        cig = ""
        for i in range(n_cig):
            v = struct.unpack_from("<I", raw, q + 4 * i)[0]
            cig += "%d%s" % (v >> 4, _CIGAR_OPS[v & 15])
        q += 4 * n_cig
        sb = raw[q: q + (l_seq + 1) // 2]
        seq = "".join(_SEQ_DECODE[b >> 4] + _SEQ_DECODE[b & 15] for b in sb)[:l_seq]
        q += (l_seq + 1) // 2
        qual = "".join(chr(b + 33) for b in raw[q: q + l_seq])
        recs.append({"qname": qname, "flag": flag, "rname": refs[refid][0] if refid >= 0 else "*", "pos": pos0 + 1,
                     "mapq": mapq, "cigar": cig or "*", "pnext": nxt + 1, "seq": seq, "qual": qual})
        p += 4 + size
    return text, refs, recs


def sam_line(r):
    """SAM text of a read_bam record, 11 mandatory columns (what `samtools view` prints)."""
    return "\t".join([r["qname"], str(r["flag"]), r["rname"], str(r["pos"]), str(r["mapq"]), r["cigar"], "=",
                      str(r["pnext"]), "0", r["seq"], r["qual"]]) + "\n"
