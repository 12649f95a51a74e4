"""Synthetic BGZF members that exercise every corner of a DEFLATE decoder (shared by the CPU model test and the GPU
inflate test): stored / fixed / dynamic blocks, several DEFLATE blocks and empty stored blocks inside one member,
15-bit Huffman codes, dist=1 runs, 258-byte matches, 32 KB distances, maximum ISIZE."""
import zlib

import numpy as np


def member(payload: bytes, deflated: bytes) -> bytes:
    blen = len(deflated) + 26
    assert blen <= 65536, blen
    hdr = b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\0\xff" + b"\x06\0" + b"BC\x02\0" + (blen - 1).to_bytes(2, "little")
    return hdr + deflated + zlib.crc32(payload).to_bytes(4, "little") + len(payload).to_bytes(4, "little")


def deflate(p: bytes, level) -> bytes:
    if level == "stored":
        co = zlib.compressobj(0, zlib.DEFLATED, -15)
    elif level == "fixed":
        co = zlib.compressobj(6, zlib.DEFLATED, -15, 9, zlib.Z_FIXED)
    elif level == "huffman":
        co = zlib.compressobj(6, zlib.DEFLATED, -15, 9, zlib.Z_HUFFMAN_ONLY)
    elif level == "rle":
        co = zlib.compressobj(6, zlib.DEFLATED, -15, 9, zlib.Z_RLE)
    else:
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
    return co.compress(p) + co.flush()


def bgzf(payloads, level) -> bytes:
    return b"".join(member(p, deflate(p, level)) for p in payloads)


def basic_payloads(seed=7, n_blocks=300, max_len=65280):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n_blocks):
        n = int(rng.integers(1, max_len))
        kind = i % 5
        if kind == 0:
            p = rng.integers(0, 256, n, dtype=np.uint8).tobytes()                      # incompressible
        elif kind == 1:
            p = bytes([int(rng.integers(0, 256))]) * n                                  # dist=1 runs
        elif kind == 2:
            p = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), n).tobytes()        # tiny alphabet
        elif kind == 3:
            w = rng.integers(0, 256, 37, dtype=np.uint8).tobytes(); p = (w * (n // 37 + 1))[:n]   # period 37
        else:
            words = [rng.integers(97, 123, int(rng.integers(2, 12)), dtype=np.uint8).tobytes() for _ in range(200)]
            p = b" ".join(words[int(j)] for j in rng.integers(0, 200, n // 5 + 1))[:n]   # text-like
        out.append(p)
    return out


def deep_tree_payload(rng, n):
    """Byte counts following the Fibonacci numbers (the worst case for Huffman depth): zlib has to limit the code
    lengths to 15 bits, so 15-bit codes are present."""
    fib = [1, 1]
    while sum(fib) < n:
        fib.append(fib[-1] + fib[-2])
    vals = rng.permutation(256)[:len(fib)].astype(np.uint8)
    a = np.repeat(vals, fib)[:n]
    rng.shuffle(a)
    return a.tobytes()


def tricky_members(seed=11):
    """(payload, member) pairs built by hand-driving zlib."""
    rng = np.random.default_rng(seed)
    cases = []
    # 15-bit codes, literal-only (Z_HUFFMAN_ONLY) and with matches
    for n in (60000, 3000, 64000):
        p = deep_tree_payload(rng, n)
        cases.append((p, member(p, deflate(p, "huffman"))))
        cases.append((p, member(p, deflate(p, 9))))
    # several DEFLATE blocks in one member, with empty stored blocks (sync flush) and a full flush in between
    parts = [deep_tree_payload(rng, 9000), rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 20000).tobytes(), b"", bytes(5000),
             rng.integers(0, 256, 3000, dtype=np.uint8).tobytes(), b"x" * 7000]
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    d = b""
    for i, part in enumerate(parts):
        d += co.compress(part) + co.flush(zlib.Z_SYNC_FLUSH if i % 2 == 0 else zlib.Z_FULL_FLUSH)
    d += co.flush()
    p = b"".join(parts)
    cases.append((p, member(p, d)))
    # level-0 member made of several stored blocks, odd lengths (byte re-alignment after every block)
    co = zlib.compressobj(0, zlib.DEFLATED, -15)
    parts = [rng.integers(0, 256, k, dtype=np.uint8).tobytes() for k in (1, 2, 3, 5, 4093, 17, 30000)]
    d = b"".join(co.compress(x) + co.flush(zlib.Z_SYNC_FLUSH) for x in parts) + co.flush()
    p = b"".join(parts)
    cases.append((p, member(p, d)))
    # stored blocks mixed with Huffman blocks
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    a = rng.integers(0, 256, 20001, dtype=np.uint8).tobytes(); b = b"GATTACA" * 3000
    d = co.compress(a) + co.flush(zlib.Z_FULL_FLUSH) + co.compress(b) + co.flush(zlib.Z_FULL_FLUSH) + co.compress(a[:777]) + co.flush()
    p = a + b + a[:777]
    cases.append((p, member(p, d)))
    # 258-byte matches at distance 1, 2, 3, 31, 32, 33 and the maximum distance 32768
    for per in (1, 2, 3, 31, 32, 33, 255, 256, 257, 258, 259):
        w = rng.integers(0, 256, per, dtype=np.uint8).tobytes()
        p = (w * (65536 // per + 1))[:65536 - per]
        cases.append((p, member(p, deflate(p, 9))))
    w = rng.integers(0, 256, 32768, dtype=np.uint8).tobytes()
    p = w[:30000] + w[:300]            # distance 30000
    cases.append((p, member(p, deflate(p, 9))))
    w = rng.integers(0, 4, 32768, dtype=np.uint8).tobytes()
    p = w + w[:32768]                  # ISIZE = 65536, distance 32768
    cases.append((p, member(p, deflate(p, 9))))
    # tiny members: 1 byte, fixed Huffman, and the BGZF EOF marker (ISIZE 0)
    for p in (b"a", b"ab", b"abc" * 5, b""):
        cases.append((p, member(p, deflate(p, 6))))
    # literal runs longer than 510 between matches (skip tokens) followed by matches
    lit = rng.integers(0, 256, 700, dtype=np.uint8).tobytes()
    p = lit + lit[:50] + rng.integers(0, 256, 1500, dtype=np.uint8).tobytes() + lit[100:400]
    cases.append((p, member(p, deflate(p, 9))))
    return cases
