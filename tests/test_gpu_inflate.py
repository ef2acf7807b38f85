"""BGZF inflate on the device (rb_inflate_bgzf; reference: myio.rs:41-64) against zlib, byte for byte.
Needs a real B200 (pytest -m gpu)."""
import os
import random
import struct
import subprocess
import zlib

import pytest

import gen
import orc
from rustybam_b200 import capi
from rustybam_b200.capi import RbError

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bgzf(data: bytes, block=60000, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, eof=True) -> bytes:
    """Minimal BGZF writer (SAM spec 4.1): gzip members with a 'BC' extra field, CRC-32 + ISIZE trailer, empty EOF block."""
    out = []
    for off in list(range(0, len(data), block)) + ([None] if eof else []):
        chunk = b"" if off is None else data[off:off + block]
        co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
        cdata = co.compress(chunk) + co.flush()
        bsize = len(cdata) + 25  # 12 header + 6 extra + cdata + 8 trailer - 1
        assert bsize < 65536
        out.append(b"\x1f\x8b\x08\x04" + b"\0" * 4 + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize))
        out.append(cdata + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
    return b"".join(out)


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


def test_bundled_paf_every_block_type(ctx):
    text = orc.golden_paf()  # 2.05 MB of PAF text
    for level, strategy, block in ((6, zlib.Z_DEFAULT_STRATEGY, 60000), (1, zlib.Z_DEFAULT_STRATEGY, 65280), (9, zlib.Z_DEFAULT_STRATEGY, 4000),
                                   (0, zlib.Z_DEFAULT_STRATEGY, 60000),   # stored blocks
                                   (6, zlib.Z_FIXED, 50000),              # fixed Huffman codes
                                   (6, zlib.Z_HUFFMAN_ONLY, 30000), (6, zlib.Z_RLE, 30000)):
        z = bgzf(text, block, level, strategy)
        assert capi.load().rb_is_bgzf(z, len(z)) == 1
        assert ctx.inflate_bgzf(z) == text


def test_random_payloads(ctx):
    rng = random.Random(5)
    for k in range(12):
        kind = k % 4
        n = rng.choice([1, 2, 7, 300, 70000, 400000])
        if kind == 0:
            data = bytes(rng.getrandbits(8) for _ in range(min(n, 70000)))           # incompressible
        elif kind == 1:
            data = bytes(rng.choice(b"0123456789=XID") for _ in range(n))             # CIGAR-like
        elif kind == 2:
            data = (b"abc" * (n // 3 + 1))[:n]                                        # overlapping matches (dist < len)
        else:
            data = b"\0" * n                                                          # runs: dist 1, len 258
        for block in (rng.choice([1, 17, 1000]), 65280):
            z = bgzf(data, block, rng.choice([1, 6, 9]), eof=bool(k & 1))
            assert ctx.inflate_bgzf(z) == data


def test_many_blocks_and_empty_inputs(ctx):
    text = gen.random_paf(3, n_contigs=4, recs_per_contig=40, max_ops=3000)[0]
    z = bgzf(text, 512, 6)  # thousands of small blocks: more threads than one SM holds
    assert ctx.inflate_bgzf(z) == text
    assert ctx.inflate_bgzf(bgzf(b"")) == b""
    assert ctx.inflate_bgzf(b"") == b""


def test_corrupt_input_is_refused(ctx):
    text = orc.golden_paf()[:200000]
    z = bytearray(bgzf(text, 60000, 6))
    import gzip
    with pytest.raises(RbError):
        ctx.inflate_bgzf(gzip.compress(text))      # a plain gzip member is not BGZF (rb_is_bgzf says so up front)
    assert capi.load().rb_is_bgzf(gzip.compress(text), 30) == 0
    bad = bytearray(z)
    bad[len(bad) // 3] ^= 0x5A                     # a flipped byte inside some block's DEFLATE data: CRC / structure check
    with pytest.raises(RbError):
        ctx.inflate_bgzf(bytes(bad))
    bad = bytearray(z)
    first_bsize = struct.unpack_from("<H", z, 16)[0] + 1
    bad[first_bsize - 8] ^= 1                      # the first block's CRC-32
    with pytest.raises(RbError):
        ctx.inflate_bgzf(bytes(bad))
    with pytest.raises(RbError):
        ctx.inflate_bgzf(bytes(z[:-5]))            # truncated file
    assert ctx.inflate_bgzf(bytes(z)) == text      # the context still works


def test_random_corruptions_never_pass_silently(ctx):
    """A flipped byte anywhere in the file either does not matter (gzip header fields nobody reads) or is refused: structure
    checks while decoding, then ISIZE and CRC-32 of the block.  (tools/sanitize.sh runs this file under memcheck: a corrupt stream
    must not write or read out of bounds either.)"""
    text = orc.golden_paf()[:300000]
    z = bgzf(text, 20000, 6)
    rng = random.Random(11)
    refused = same = 0
    for _ in range(150):
        bad = bytearray(z)
        k = rng.randrange(len(bad))
        bad[k] ^= 1 << rng.randrange(8)
        try:
            got = ctx.inflate_bgzf(bytes(bad))
        except RbError:
            refused += 1
            continue
        assert got == text, "corruption at byte %d went through" % k
        same += 1
    assert refused > 100 and ctx.inflate_bgzf(z) == text


def test_cli_reads_bgz_through_the_device(tmp_path):
    rb = os.path.join(ROOT, "rustybam_b200", "rb")
    paf = orc.golden_paf()
    (tmp_path / "a.paf").write_bytes(paf)
    (tmp_path / "a.paf.bgz").write_bytes(bgzf(paf))
    want = subprocess.run([rb, "stats", "--paf", str(tmp_path / "a.paf")], capture_output=True, check=True).stdout
    env = dict(os.environ, RB_TIMING="1")
    got = subprocess.run([rb, "stats", "--paf", str(tmp_path / "a.paf.bgz")], capture_output=True, check=True, env=env)
    assert got.stdout == want == orc.run_stats(paf)
    assert b"gpu_inflate" in got.stderr            # the .bgz went through rb_inflate_bgzf, not the host's zlib
    host = subprocess.run([rb, "stats", "--paf", str(tmp_path / "a.paf.bgz")], capture_output=True, check=True, env=dict(env, RB_GPU_INFLATE="0"))
    assert host.stdout == want and b"gpu_inflate" not in host.stderr
