"""CPU: the reader's own gzip decoder (flagger_b200/csrc/hfg_inflate.c) against zlib on generated inputs: every compression
level and strategy (the reference writes Huffman-only streams, submodules/ptBlock/ptBlock.c:2271), stored / fixed / dynamic
blocks, piece boundaries inside matches, multi-member files, header fields, trailing zeros, and damaged files, which must
fail with a message (CRC-32 and length of every member are checked)."""
import ctypes as C
import gzip
import os
import zlib

import numpy as np
import pytest

from flagger_b200 import api


def gunzip(path, cap, piece=0):
    out = np.zeros(max(cap, 1), np.uint8)
    n = C.c_size_t(0)
    err = C.create_string_buffer(256)
    f = api.lib().hfg_debug_gunzip
    f.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_size_t, C.c_char_p, C.c_size_t]
    rc = f(str(path).encode(), out.ctypes.data, out.size, C.byref(n), piece, err, 256)
    return rc, bytes(out[:n.value]), err.value.decode()


def gz_bytes(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY):
    co = zlib.compressobj(level, zlib.DEFLATED, 31, 8, strategy)
    return co.compress(data) + co.flush()


def samples():
    rng = np.random.default_rng(3)
    text = b"".join(b"%d\t%d\t%d\t%d\t0\t1\t0\n" % (i * 37, i * 37 + 36, 30 + i % 17, 27 + i % 15) for i in range(60000))
    return {
        "empty": b"",
        "one": b"x",
        "text": text,
        "zeros": bytes(300000),
        "random": rng.integers(0, 256, 200000, dtype=np.uint8).tobytes(),
        "runs": b"".join(bytes([int(b)]) * int(n) for b, n in zip(rng.integers(0, 256, 4000), rng.integers(1, 600, 4000))),
        "far": (rng.integers(0, 256, 40000, dtype=np.uint8).tobytes()) * 6,  # matches at distances up to 32768
    }


@pytest.mark.parametrize("name", list(samples()))
def test_inflate_matches_zlib_for_every_level_and_strategy(tmp_path, name):
    data = samples()[name]
    strategies = (zlib.Z_DEFAULT_STRATEGY, zlib.Z_HUFFMAN_ONLY, zlib.Z_FILTERED, zlib.Z_RLE, zlib.Z_FIXED)
    for level in (0, 1, 6, 9):
        for strategy in strategies:
            p = tmp_path / f"{name}_{level}_{strategy}.gz"
            p.write_bytes(gz_bytes(data, level, strategy))
            for piece in (0, 4096, 70001):
                rc, out, err = gunzip(p, len(data) + 1024, piece)
                assert rc == 0, (level, strategy, piece, err)
                assert out == data, (level, strategy, piece)


def test_inflate_members_header_fields_and_padding(tmp_path):
    s = samples()
    p = tmp_path / "multi.gz"
    with open(p, "wb") as f:
        f.write(gz_bytes(s["text"], 6, zlib.Z_HUFFMAN_ONLY))
        f.write(gz_bytes(b"", 6))
        f.write(gz_bytes(s["far"], 9))
        f.write(bytes(37))  # zero padding behind the last member, as tape / block writers leave it
    want = s["text"] + s["far"]
    for piece in (0, 5000):
        rc, out, err = gunzip(p, len(want) + 10, piece)
        assert rc == 0 and out == want, err
    q = tmp_path / "named.gz"  # FNAME + mtime through Python's gzip module
    with gzip.GzipFile(filename="some_name.cov", mode="wb", fileobj=open(q, "wb"), mtime=12345) as g:
        g.write(s["runs"])
    rc, out, err = gunzip(q, len(s["runs"]) + 10)
    assert rc == 0 and out == s["runs"], err
    # FEXTRA, FCOMMENT and FHCRC by hand around a raw deflate stream
    raw = zlib.compress(s["text"], 6)[2:-4]
    hdr = bytes([0x1f, 0x8b, 8, 4 | 8 | 16 | 2, 0, 0, 0, 0, 0, 3]) + (5).to_bytes(2, "little") + b"extra" + b"name\0" + b"comment\0"
    hdr += (zlib.crc32(hdr) & 0xffff).to_bytes(2, "little")
    r = tmp_path / "fields.gz"
    r.write_bytes(hdr + raw + zlib.crc32(s["text"]).to_bytes(4, "little") + (len(s["text"]) & 0xffffffff).to_bytes(4, "little"))
    rc, out, err = gunzip(r, len(s["text"]) + 10)
    assert rc == 0 and out == s["text"], err


def test_inflate_rejects_damage(tmp_path):
    data = samples()["text"]
    good = gz_bytes(data, 6, zlib.Z_HUFFMAN_ONLY)
    rng = np.random.default_rng(11)
    cases = {"truncated": good[: len(good) // 2], "no_trailer": good[:-8], "bad_crc": good[:-8] + bytes(4) + good[-4:],
             "bad_len": good[:-4] + bytes(4)}
    for k in range(12):  # single flipped bytes anywhere in the stream
        pos = int(rng.integers(10, len(good) - 8))
        b = bytearray(good)
        b[pos] ^= 1 << int(rng.integers(0, 8))
        cases[f"flip{k}"] = bytes(b)
    for name, blob in cases.items():
        p = tmp_path / f"{name}.gz"
        p.write_bytes(blob)
        rc, out, err = gunzip(p, len(data) * 2 + 1024)
        assert rc != 0 and err, name  # never silently wrong
    p = tmp_path / "plain.txt"
    p.write_bytes(data)
    rc, _, err = gunzip(p, 10)
    assert rc != 0 and "not a readable gzip file" in err
    # garbage BEHIND a complete, CRC-checked member ends the stream quietly, as zlib's gzread (the reference's reader) does
    for name, tail in (("garbage_after", b"garbage"), ("zeros_then_garbage", bytes(5) + b"xyz"), ("short", b"\x1f")):
        p = tmp_path / f"{name}.gz"
        p.write_bytes(good + tail)
        rc, out, err = gunzip(p, len(data) * 2 + 1024)
        assert rc == 0 and out == data, (name, err)
        with gzip.open(p, "rb") as f:  # the zlib family agrees on the content it returns
            try:
                assert f.read(len(data)) == data
            except OSError:
                pass


def test_cov_reader_uses_the_decoder_and_zlib_gives_the_same(tmp_path, monkeypatch):
    from flagger_b200 import binfmt
    p = tmp_path / "h.cov"
    binfmt.write_random_rle_cov(str(p), [30011, 8200, 5], seed=4, n_regions=2)
    q = tmp_path / "h.cov.gz"
    q.write_bytes(gz_bytes(p.read_bytes(), 6, zlib.Z_HUFFMAN_ONLY))  # what gzopen(path, "w6h") writes
    a, ha = binfmt.read_cov_native(str(p), 7000, 1000)
    b, hb = binfmt.read_cov_native(str(q), 7000, 1000)
    monkeypatch.setenv("HFG_ZLIB_INFLATE", "1")
    c, hc = binfmt.read_cov_native(str(q), 7000, 1000)
    for other, h in ((b, hb), (c, hc)):
        assert np.array_equal(a.cov, other.cov) and np.array_equal(a.chunks, other.chunks)
        assert np.array_equal(ha["annotation_flag"], h["annotation_flag"]) and np.array_equal(a.truth, other.truth)
    bad = tmp_path / "bad.cov.gz"
    blob = bytearray(q.read_bytes())
    blob[len(blob) // 2] ^= 0x10
    bad.write_bytes(bytes(blob))
    monkeypatch.delenv("HFG_ZLIB_INFLATE")
    with pytest.raises(ValueError):
        binfmt.read_cov_native(str(bad), 7000, 1000)
