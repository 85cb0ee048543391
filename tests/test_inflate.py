"""inflate_fast.h (the BGZF block decoder of the native BAM reader) against zlib: stored / fixed / dynamic blocks,
long codes (subtables), overlapping and far matches, tiny and empty streams, truncated and corrupted input."""
import ctypes as C
import os
import random
import zlib

import numpy as np
import pytest

from himut_b200 import bamdec


def _inflate(comp, n):
    lib = C.CDLL(bamdec.lib_path())
    lib.hm_inflate_raw_test.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t]
    out = np.zeros(max(n, 1), np.uint8)
    rc = lib.hm_inflate_raw_test(comp, len(comp), out.ctypes.data_as(C.c_void_p), n)
    return rc, out[:n].tobytes()


def _raw_deflate(data, level, strategy=zlib.Z_DEFAULT_STRATEGY):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    return c.compress(data) + c.flush()


def _payloads():
    rnd = random.Random(7)
    yield b""
    yield b"a"
    yield b"abc" * 5
    yield bytes(1000)                                        # dist-1 runs
    yield bytes(rnd.randrange(256) for _ in range(70_000))   # incompressible: stored blocks / literal-only trees
    yield bytes(rnd.choice(b"ACGT") for _ in range(65_000))  # 4-symbol alphabet: short codes
    q = bytearray()
    while len(q) < 64_000:                                    # quality-string like: long runs with exceptions
        q += bytes([93]) * rnd.randrange(1, 60) + bytes([rnd.randrange(1, 93)])
    yield bytes(q[:64_000])
    words = [bytes(rnd.randrange(256) for _ in range(rnd.randrange(3, 40))) for _ in range(300)]
    yield b"".join(rnd.choice(words) for _ in range(4000))[:65_280]   # many distinct lengths / distances
    skew = bytes(min(255, int(rnd.expovariate(0.03))) for _ in range(65_000))  # skewed alphabet: codes longer than 10 bits
    yield skew


@pytest.mark.parametrize("level", [0, 1, 6, 9])
@pytest.mark.parametrize("strategy", [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE])
def test_matches_zlib(level, strategy):
    for data in _payloads():
        comp = _raw_deflate(data, level, strategy)
        assert zlib.decompress(comp, -15) == data
        rc, got = _inflate(comp, len(data))
        assert rc == 0 and got == data, (level, strategy, len(data))


def test_rejects_bad_input():
    data = bytes(random.Random(3).choice(b"ACGTN") for _ in range(20_000))
    comp = _raw_deflate(data, 6)
    assert _inflate(comp, len(data))[0] == 0
    assert _inflate(comp, len(data) - 1)[0] != 0         # output size is part of the contract
    assert _inflate(comp, len(data) + 1)[0] != 0
    assert _inflate(comp[: len(comp) // 2], len(data))[0] != 0   # truncated
    rnd = random.Random(4)
    for _ in range(200):                                  # corrupted: must not crash; a wrong answer is caught by the CRC
        bad = bytearray(comp)
        bad[rnd.randrange(len(bad))] ^= 1 << rnd.randrange(8)
        rc, got = _inflate(bytes(bad), len(data))
        assert rc != 0 or len(got) == len(data)


def test_bam_blocks_decode_identically_with_both_decoders(tmp_path, monkeypatch):
    from himut_b200 import synth
    import cases
    d = synth.generate(400_000, seed=81)
    digests = []
    for level in (1, 6):
        path = str(tmp_path / ("l%d.bam" % level))
        bamdec.write_batch_bam(path, "chr1", 400_000, d.batch, level=level)
        for zl in (False, True):
            if zl:
                monkeypatch.setenv("HIMUT_B200_ZLIB_INFLATE", "1")
            else:
                monkeypatch.delenv("HIMUT_B200_ZLIB_INFLATE", raising=False)
            nb = bamdec.NativeBam(path, threads=2)
            digests.append(cases.batch_digest(nb.read_batch("chr1", 0, 400_000)))
    assert len(set(digests)) == 1 and digests[0] == cases.batch_digest(d.batch)


def test_block_crc_equals_zlib():
    """every inflated BGZF block is checked with block_crc (carry-less-multiply folding + zlib for the tail,
    csrc/crc32_clmul.h): it must be zlib's crc32 for every length and content"""
    lib = bamdec.load()
    lib.hm_crc32_test.argtypes = [C.c_char_p, C.c_size_t]
    lib.hm_crc32_test.restype = C.c_uint32
    rnd = random.Random(5)
    lengths = list(range(0, 200)) + [rnd.randrange(0, 70_000) for _ in range(200)] + [65536, 65280, 4096, 4097]
    for n in lengths:
        data = rnd.randbytes(n)
        assert lib.hm_crc32_test(data, n) == zlib.crc32(data), n
    for data in (b"\x00" * 1000, b"\xff" * 4097, bytes(range(256)) * 37):
        assert lib.hm_crc32_test(data, len(data)) == zlib.crc32(data)


def _inflate2(c0, n0, c1, n1):
    lib = C.CDLL(bamdec.lib_path())
    vp = C.c_void_p
    lib.hm_inflate_raw2_test.argtypes = [C.c_char_p, C.c_size_t, vp, C.c_size_t, C.c_char_p, C.c_size_t, vp, C.c_size_t]
    o0, o1 = np.zeros(max(n0, 1), np.uint8), np.zeros(max(n1, 1), np.uint8)
    rc = lib.hm_inflate_raw2_test(c0, len(c0), o0.ctypes.data_as(vp), n0, c1, len(c1), o1.ctypes.data_as(vp), n1)
    return rc, o0[:n0].tobytes(), o1[:n1].tobytes()


@pytest.mark.parametrize("level", [1, 6])
def test_two_streams_in_one_loop_match_zlib(level):
    """the BGZF workers inflate two blocks at a time (hm_inflate_raw2: both streams advance in one loop); every
    pairing of the payload kinds, in both orders, so the streams end, stall on rare codes and leave their fast
    regions at different times"""
    items = [(d, _raw_deflate(d, level)) for d in _payloads()]
    items.append((items[6][0], _raw_deflate(items[6][0], 6, zlib.Z_FIXED)))
    for d0, c0 in items:
        for d1, c1 in items:
            rc, g0, g1 = _inflate2(c0, len(d0), c1, len(d1))
            assert rc == 0 and g0 == d0 and g1 == d1, (len(d0), len(d1))


def test_two_streams_fail_independently():
    rnd = random.Random(8)
    good = bytes(rnd.choice(b"ACGT") for _ in range(30_000))
    other = bytes([93] * 10 + [20]) * 3000
    cg, co = _raw_deflate(good, 6), _raw_deflate(other, 1)
    assert _inflate2(cg, len(good), co, len(other) - 1)[0] == 2      # second stream: wrong size
    assert _inflate2(cg[:100], len(good), co, len(other))[0] == 1    # first stream: truncated
    rc, g0, g1 = _inflate2(cg, len(good), co[: len(co) // 2], len(other))
    assert rc == 2 and g0 == good
    for _ in range(200):                                              # corrupted: no crash; wrong bytes are the CRC's job
        bad = bytearray(co)
        bad[rnd.randrange(len(bad))] ^= 1 << rnd.randrange(8)
        rc, g0, _ = _inflate2(cg, len(good), bytes(bad), len(other))
        assert (rc & 1) == 0 and g0 == good


def test_sanitizer_fuzz_session(tmp_path):
    """tools/fuzz_inflate.c under AddressSanitizer + UBSan: intact streams decode exactly (single and pair decoder),
    damaged ones never touch memory outside their exact-size buffers"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "fz")
    cc = subprocess.run(["gcc", "-O1", "-g", "-std=c11", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                         "-I", os.path.join(root, "himut_b200", "csrc"), "-o", exe, os.path.join(root, "tools", "fuzz_inflate.c"), "-lz"],
                        capture_output=True, text=True)
    if cc.returncode != 0:
        pytest.skip("no sanitizer runtime here: " + cc.stderr[-200:])
    run = subprocess.run([exe, "400"], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stdout[-500:] + run.stderr[-2000:]
    assert "iterations 3200" in run.stdout
