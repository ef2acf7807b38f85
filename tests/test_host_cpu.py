"""CPU-side checks: the C-ABI library loads and exports every symbol include/rbcuda.h declares
(no compute without a GPU), the host packing mirrors the reference's parser, and the product
path fails loudly without a device instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest

import gen
import orc
from rustybam_b200 import bamstats, bed, build, capi, hostlib
from rustybam_b200.paf import Paf, ReferencePanic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _built():
    build.build_cuda()


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "rbcuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(rb_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    lib = ctypes.CDLL(capi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name


def test_rust_binding_declares_only_what_the_header_and_library_have():
    # rust/src/ffi.rs is the binding a maintainer adds to the reference (no Rust toolchain here: checked as text)
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "rbcuda.h")).read(), flags=re.S)
    declared = set(re.findall(r"\b(rb_[a-z_0-9]+)\s*\(", hdr))
    ffi = open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read()
    bound = set(re.findall(r"pub fn (rb_[a-z_0-9]+)", ffi))
    assert bound and bound <= declared, bound - declared
    assert {"rb_ctx_create", "rb_liftover", "rb_stats", "rb_free_lift_out", "rb_free_stats_out", "rb_last_error"} <= bound
    lib = ctypes.CDLL(capi.LIB_PATH)
    for name in bound:
        assert hasattr(lib, name), name
    for const, val in (("RB_WANT_TEXT", 1), ("RB_WANT_NUMERIC", 2), ("RB_WANT_QBED", 4), ("RB_WANT_STATS_TEXT", 8)):
        assert re.search(rf"#define {const} {val}u", open(os.path.join(ROOT, "include", "rbcuda.h")).read())
        assert re.search(rf"pub const {const}: u32 = {val};", ffi)
    assert os.path.exists(os.path.join(ROOT, "rust", "build.rs"))


def test_no_device_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.RbError) as e:
        capi.Context(0)
    assert e.value.code == capi.RB_ERR_NO_DEVICE


def test_product_does_not_reference_the_oracle():
    for d, _, files in os.walk(os.path.join(ROOT, "rustybam_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(d, f), errors="ignore").read()
                assert "rb_oracle" not in src and "liborc" not in src and "import orc" not in src, f


def test_paf_pack_layout():
    p = Paf.from_text(orc.golden_paf())
    assert len(p) == 249 and p.skipped == 0
    r = p.pack()
    assert r.n_rec == 249 and r.cigar_nbytes == 2049269
    assert int(r.cigar_off[0]) == 0 and int(r.cigar_off[-1]) == r.cigar_nbytes
    assert bytes(r.cigar[:8]) == b"2395=14D"
    assert r.names[int(r.t_id[0])] == b"chr20"


def test_paf_parser_mirrors_reference_quirks():
    with pytest.raises(ReferencePanic):
        Paf.from_text(b"a\tb\n")  # < 12 columns -> assert (paf.rs:381)
    with pytest.raises(ReferencePanic):
        Paf.from_text(b"\n")
    p = Paf.from_text(b"Q 10 x 10 + T 40 12 20 3 9 60 cg:Z:8=\n")  # non-numeric column -> skipped (paf.rs:73)
    assert len(p) == 0 and p.skipped == 1
    p = Paf.from_text(b"Q 10 2 10 + T 40 12 20 3 9 60 cs:Z::8 cg:Z:8= cg:Z:4=\n")  # first cg wins, any whitespace splits
    assert p.cigars == [b"8="]
    with pytest.raises(ReferencePanic):
        Paf.from_text(b"Q 10 2 10 + T 40 12 20 3 9 60 notatag\n")
    # a line that would be skipped (bad numeric column) still panics when its cg tag is malformed: the reference parses the
    # CIGAR first (paf.rs:386-399, .expect) and only then the numbers (paf.rs:401-417) — same in the C++ host and the oracle
    for bad in (b"Q 10 x 10 + T 40 12 20 3 9 60 cg:Z:8=3\n", b"Q 10 x 10 + T 40 12 20 3 9 60 cg:Z:8Z\n", b"Q 10 x 10 + T 40 12 20 3 9 60 cg:Z:3=2H3=\n"):
        with pytest.raises(ReferencePanic):
            Paf.from_text(bad)
        with pytest.raises(hostlib.HostPanic):
            hostlib.HostPaf.from_text(bad)
        with pytest.raises(orc.ReferencePanic):
            orc.run_stats(bad)
    ok = b"Q 10 x 10 + T 40 12 20 3 9 60 cg:Z:2S6=\n"
    assert len(Paf.from_text(ok)) == 0 and hostlib.HostPaf.from_text(ok).skipped == 1 and orc.run_stats(ok).count(b"\n") == 1


def test_bed_parser_matches_oracle():
    texts = [orc.golden_bed(), b"chr1\t2\t2000\n", b"chr1\t0\t1000\tid\n", b"#c\nchr1\t5\t9\tA\textra\nchr2\t1\t2\tB\textra\nchr3\t1\t2\n",
             b"chr1\t5\t9\r\nchr1\tx\t9\nchr1\t10\t12\n"]
    for t in texts:
        got = [(r.name.decode(), r.st, r.en, r.id.decode()) for r in bed.parse_bed_text(t)]
        assert got == orc.parse_bed(t)


def test_window_packing_sorts_and_keeps_bed_rows():
    paf_text, contigs = gen.random_paf(3)
    recs = Paf.from_text(paf_text).pack()
    rg = bed.parse_bed_text(gen.random_bed(5, contigs, 40))
    w = bed.pack_windows(rg, recs.name_index)
    key = list(zip(w.t_id.tolist(), w.st.tolist()))
    assert key == sorted(key)
    assert sorted(w.bed_row.tolist()) == list(range(w.n_win))


def test_fast_bed_packer_equals_the_two_step_form():
    """Windows::pack_text (one parallel pass over the BED text, no per-row strings: what `rb liftover` runs on a 3 M-row BED)
    == Windows::pack(parse_bed_text(..)): rows, bed_row numbers (counted over every row that parses, bed.rs:140-194) and ids."""
    import ctypes as C

    def arrs(w):
        n = w.n_win
        a = [np.ctypeslib.as_array(getattr(w.c, k), shape=(n,)).copy() if n else np.zeros(0) for k in ("t_id", "st", "en", "bed_row")]
        ids = None
        if w.c.ids_off:
            off = np.ctypeslib.as_array(w.c.ids_off, shape=(n + 1,)).copy()
            ids = (off, bytes(np.ctypeslib.as_array(w.c.ids, shape=(int(off[-1]),))) if off[-1] else b"")
        return a, ids
    for seed in range(12):
        paf_text, contigs = gen.random_paf(seed)
        hp = hostlib.HostPaf.from_text(paf_text)
        beds = [gen.random_bed(seed, contigs, 50), gen.random_bed(seed, contigs, 50, with_ids=False), gen.tiling_bed(contigs, 33, with_ids=bool(seed & 1)),
                b"#c\nchr1\t5\t9\tA\textra\nchr2\t1\t2\tB\textra\nchr3\t1\t2\n", b"chr1\t5\t9\r\nchr1\tx\t9\nchr1\t10\t12\n\n\nchrZ\t1\t2\nchr1\t1\t3\n", b""]
        for bed_text in beds:
            (a, i1), (b, i2) = arrs(hp.windows_from_bed_text(bed_text)), arrs(hp.windows_from_bed_text_slow(bed_text))
            assert all(np.array_equal(x, y) for x, y in zip(a, b))
            assert (i1 is None) == (i2 is None) and (i1 is None or (np.array_equal(i1[0], i2[0]) and i1[1] == i2[1]))
    big = hostlib.HostPaf.synth(scale=0.05)
    bed_text = big.tiling_bed_text(100)  # 1.5 M rows: the threaded path
    (a, _), (b, _) = arrs(big.windows_from_bed_text(bed_text)), arrs(big.windows_from_bed_text_slow(bed_text))
    assert len(a[0]) > 1_000_000 and all(np.array_equal(x, y) for x, y in zip(a, b))
    # the threaded path on an awkward text: ids, \r\n endings, comments and blank lines in between, rows out of order, a row
    # that does not parse, a contig the PAF does not have — thread ranges cut lines wherever they fall
    import random
    rng = random.Random(5)
    rows = bed_text.splitlines()[:400_000]
    rng.shuffle(rows)
    out = []
    for k, r in enumerate(rows):
        out.append(r + b"\tid%d" % k + (b"\r\n" if k % 3 == 0 else b"\n"))
        if k % 1000 == 0:
            out.append(b"#comment\n\n")
        if k % 7777 == 0:
            out.append(b"chrNOPE\t1\t2\tx\nchr1\tx\t2\ty\n")
    awkward = b"".join(out)
    assert len(awkward) > (8 << 20)
    (a, i1), (b, i2) = arrs(big.windows_from_bed_text(awkward)), arrs(big.windows_from_bed_text_slow(awkward))
    assert len(a[0]) == 400_000 and all(np.array_equal(x, y) for x, y in zip(a, b))
    assert np.array_equal(i1[0], i2[0]) and i1[1] == i2[1]


def test_fmt_f32_matches_oracle():
    rng = np.random.default_rng(3)
    eq = rng.integers(0, 2**31, 3000, dtype=np.uint64)
    tot = eq + rng.integers(0, 2**22, 3000, dtype=np.uint64)
    vals = (np.float32(100.0) * eq.astype(np.float32)) / np.maximum(tot, 1).astype(np.float32)
    for v in list(vals) + [np.float32("nan"), np.float32(0), np.float32(100), np.float32(1e-7), np.float32(50)]:
        assert bamstats.fmt_f32(v) == orc.fmt_f32(float(v)), v
    # exact ties between the two shortest candidates round UP in Rust's Display (numpy / printf: to even):
    # 100 * 2049 / 12800 = 16.0078125 is reachable (2049 '=' bases, 10751 'X' bases)
    assert bamstats.fmt_f32(np.float32(16.0078125)) == "16.007813" == orc.fmt_f32(16.0078125)
    for k in range(1, 4000, 7):  # odd multiples of 2^-7 in [16, 32): 8 significant digits, the 9th is an exact 5
        v = np.float32(16.0 + (2 * k + 1) / 128.0)
        assert bamstats.fmt_f32(v) == orc.fmt_f32(float(v)), v


def test_name_interning_scales():
    """A read-level PAF interns one name per record: 200 k distinct query names must parse in seconds, not minutes."""
    import time
    n = 200_000
    lines = b"".join(b"q%07d\t100\t0\t10\t+\tchr1\t1000\t0\t10\t10\t10\t60\tcg:Z:10=\n" % (n - i) for i in range(n))
    t0 = time.time()
    paf = hostlib.HostPaf.from_text(lines)
    assert paf.n_rec == n and time.time() - t0 < 20
    assert paf.find_name("q%07d" % 5) >= 0 and paf.find_name("nope") == -1


def _write_bgzf(path, data: bytes, block=60000):
    """Minimal BGZF writer (SAM spec 4.1): gzip members with a 'BC' extra field, ISIZE trailer, empty EOF block."""
    import struct
    import zlib
    with open(path, "wb") as f:
        for off in list(range(0, len(data), block)) + [None]:
            chunk = b"" if off is None else data[off:off + block]
            co = zlib.compressobj(6, zlib.DEFLATED, -15)
            cdata = co.compress(chunk) + co.flush()
            bsize = len(cdata) + 25  # 12 header + 6 extra + cdata + 8 trailer - 1
            f.write(b"\x1f\x8b\x08\x04" + b"\0" * 4 + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize))
            f.write(cdata + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))


def test_paf_from_file_maps_plain_files_and_inflates_the_rest(tmp_path):
    # Paf::from_file: an uncompressed file is mmap'ed and parsed in place, .gz / .bgz go through the inflating reader; every
    # form packs to the same records.  Plus the edge shapes of the line scan: no trailing newline, CRLF, an empty file.
    import gzip
    from rustybam_b200 import hostlib
    text = orc.golden_paf()
    want = hostlib.HostPaf.from_text(text)
    (tmp_path / "a.paf").write_bytes(text)
    (tmp_path / "a.paf.gz").write_bytes(gzip.compress(text))
    _write_bgzf(tmp_path / "a.paf.bgz", text)
    for name in ("a.paf", "a.paf.gz", "a.paf.bgz"):
        got = hostlib.HostPaf.from_file(str(tmp_path / name))
        assert got.n_rec == want.n_rec == 249 and got.text() == want.text()
        got.close()
    lines = text.splitlines()[:5]
    (tmp_path / "b.paf").write_bytes(b"\r\n".join(lines))  # CRLF between lines, nothing after the last one
    got = hostlib.HostPaf.from_file(str(tmp_path / "b.paf"))
    assert got.n_rec == 5 and got.text() == hostlib.HostPaf.from_text(b"\n".join(lines) + b"\n").text()
    got.close()
    (tmp_path / "c.paf").write_bytes(b"")
    got = hostlib.HostPaf.from_file(str(tmp_path / "c.paf"))
    assert got.n_rec == 0
    got.close()
    with pytest.raises(hostlib.HostPanic):
        hostlib.HostPaf.from_file(str(tmp_path / "missing.paf"))
    want.close()


def test_bgzf_detection_needs_no_device(tmp_path):
    # rb_is_bgzf (the switch in front of rb_inflate_bgzf, myio.rs:41-64): BGZF blocks yes, a plain gzip member or text no
    import gzip
    from rustybam_b200 import capi
    lib = capi.load()
    text = orc.golden_paf()[:100000]
    p = tmp_path / "a.bgz"
    _write_bgzf(p, text)
    z = p.read_bytes()
    assert lib.rb_is_bgzf(z, len(z)) == 1
    g = gzip.compress(text)
    assert lib.rb_is_bgzf(g, len(g)) == 0 and lib.rb_is_bgzf(text, len(text)) == 0 and lib.rb_is_bgzf(b"", 0) == 0
    assert lib.rb_is_bgzf(z, 10) == 0  # shorter than a block header


def test_host_reader_inflates_bgzf_blocks_in_parallel(tmp_path):
    # myio.rs:33-40: .paf, .paf.gz and .paf.bgz hold the same lines; the .bgz reader works block by block on all threads
    import gzip
    from rustybam_b200 import hostlib
    text = orc.golden_paf()
    bgz, gz, plain = tmp_path / "a.paf.bgz", tmp_path / "a.paf.gz", tmp_path / "a.paf"
    _write_bgzf(bgz, text)
    gz.write_bytes(gzip.compress(text))
    plain.write_bytes(text)
    assert hostlib.read_all(str(bgz)) == text
    assert hostlib.read_all(str(gz)) == text
    assert hostlib.read_all(str(plain)) == text
    (tmp_path / "b.paf.bgz").write_bytes(gzip.compress(text))  # a plain gzip stream behind a .bgz name still reads
    assert hostlib.read_all(str(tmp_path / "b.paf.bgz")) == text
