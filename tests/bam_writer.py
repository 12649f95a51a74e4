"""A minimal BAM writer for tests that need hand-made records (TEST INFRASTRUCTURE ONLY): every record goes into a BGZF block of its
own, so the virtual offset of record i is simply `block_start[i] << 16`.  [ext] SAM/BAM spec 4.2; SURVEY.md Appendix B."""
import struct
import zlib

_OPS = {c: i for i, c in enumerate("MIDNSHP=X")}


def _bgzf_block(payload: bytes) -> bytes:
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = co.compress(payload) + co.flush()
    blen = len(body) + 26
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", blen - 1) + body
            + struct.pack("<II", zlib.crc32(payload) & 0xFFFFFFFF, len(payload)))


def parse_cigar(s):
    out, n = [], ""
    for ch in s:
        if ch.isdigit():
            n += ch
        else:
            out.append((_OPS[ch], int(n))); n = ""
    return out


def record(tid, pos, cigar, flag=0, mapq=60, mtid=-1, mpos=-1, tlen=0, qname="r"):
    cig = parse_cigar(cigar) if isinstance(cigar, str) else cigar
    l_seq = sum(n for op, n in cig if op in (0, 1, 4, 7, 8))
    qn = qname.encode() + b"\0"
    core = struct.pack("<iiBBHHHiiii", tid, pos, len(qn), mapq, 4680, len(cig), flag, l_seq, mtid, mpos, tlen)
    body = core + qn + b"".join(struct.pack("<I", (n << 4) | op) for op, n in cig) + b"\x11" * ((l_seq + 1) // 2) + b"\x1e" * l_seq
    return struct.pack("<i", len(body)) + body


def write_bam(path, contigs, records):
    """contigs: [(name, length)]; records: bytes from record(), already coordinate sorted.  Returns (first_record_voff, [voff per
    record], end_voff)."""
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in contigs)
    hdr = b"BAM\1" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(contigs))
    for n, l in contigs:
        hdr += struct.pack("<i", len(n) + 1) + n.encode() + b"\0" + struct.pack("<i", l)
    out = bytearray(_bgzf_block(hdr))
    voffs = []
    for r in records:
        voffs.append(len(out) << 16)
        out += _bgzf_block(r)
    end = len(out) << 16
    out += bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
    with open(path, "wb") as f:
        f.write(out)
    return (voffs[0] if voffs else end), voffs, end
