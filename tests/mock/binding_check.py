"""Run by tests/test_cli_on_mock.py in a process of its own: mosdepth_b200's ctypes binding pointed at the TEST DOUBLE of the C ABI
(argv[1]) instead of the CUDA library, exercising Context.per_base_bgzf (pointer-to-pointer outputs, msd_bgzf_member layout)."""
import gzip
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import mosdepth_b200 as m  # noqa: E402

m.LIB_PATH = Path(sys.argv[1]); m._lib = None
ctx = m.Context(0)
hdr = ctx.open(sys.argv[2])
ctx.set_filters(mode=m.MODE_DEFAULT)
ctx.run()
EOF_BLOCK = bytes([0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 0x42, 0x43, 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0])
seen = 0
for tid, name in enumerate(hdr.names):
    rc, n = ctx.contig_records(tid)
    if rc == -2:
        continue
    data, mem, (s, e, d) = ctx.per_base_bgzf(tid, name)
    s2, e2, d2 = ctx.per_base_runs(tid)
    assert (s == s2).all() and (e == e2).all() and (d == d2).all()
    want = "".join(f"{name}\t{a}\t{b}\t{c}\n" for a, b, c in zip(s.tolist(), e.tolist(), d.tolist())).encode()
    assert gzip.decompress(data + EOF_BLOCK) == want
    assert int(mem["offset"][0]) == 0 and int(mem["first_run"][0]) == 0 and int(mem["isize"].sum()) == len(want)
    assert int(mem["offset"][-1]) + int(mem["csize"][-1]) == len(data)
    assert all(int(mem["offset"][k]) + int(mem["csize"][k]) == int(mem["offset"][k + 1]) for k in range(len(mem) - 1))
    quants = [0, 1, 50, 100, 2**63 - 1]; labels = ["0:1", "1:50", "50:100", "100:inf"]
    zdata, zmem, (qs, qe, qb) = ctx.quantized_bgzf(tid, quants, name, labels)
    qs2, qe2, qb2 = ctx.quantized_runs(tid, quants)
    assert (qs == qs2).all() and (qe == qe2).all() and (qb == qb2).all()
    qwant = "".join(f"{name}\t{a}\t{b}\t{labels[c]}\n" for a, b, c in zip(qs.tolist(), qe.tolist(), qb.tolist())).encode()
    assert gzip.decompress(zdata + EOF_BLOCK) == qwant and int(zmem["isize"].sum()) == len(qwant)
    seen += 1
assert seen >= 1
print("ok", seen)
