"""CPU tests of the two pieces of the BAM ingest that run on the device but are plain sequential functions compiled for
host and device alike: the DEFLATE decoder (fslr_b200/csrc/inflate.cuh) against zlib, in its sequential form and in the
warp-cooperative control flow emulated with a one-lane warp, and the record-boundary finder (bam_chain.cuh) against the
true record chain of synthetic BAM streams."""
import ctypes
import os
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "_build")


def _lib(name, src):
    os.makedirs(BUILD, exist_ok=True)
    so = os.path.join(BUILD, name)
    deps = [os.path.join(ROOT, "tests", src), os.path.join(ROOT, "fslr_b200", "csrc", "inflate.cuh"),
            os.path.join(ROOT, "fslr_b200", "csrc", "bam_chain.cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", "-o", so, deps[0]])
    return ctypes.CDLL(so)


def _streams():
    rng = np.random.default_rng(7)
    for trial in range(40):
        L = int(rng.integers(0, 65536)) if trial else 0
        yield [rng.integers(0, 256, L, dtype=np.uint8).tobytes(), "".join("ACGT"[i] for i in rng.integers(0, 4, L)).encode(),
               bytes([65]) * L, (b"abcabcabd" * 8000)[:L], rng.integers(0, 4, L, dtype=np.uint8).tobytes()][trial % 5]


@pytest.mark.parametrize("variant", ["sequential", "warp_flow"])
def test_inflate_matches_zlib(variant):
    lib = _lib("libhost_inflate.so", "host_inflate_harness.cpp") if variant == "sequential" else \
        _lib("libhost_inflate_warp.so", "host_inflate_warp_harness.cpp")
    fn = lib.host_inflate if variant == "sequential" else lib.host_inflate_warp
    fn.argtypes = [ctypes.c_char_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_longlong]
    for data in _streams():
        for level, strat in ((0, 0), (1, 0), (6, 0), (9, 0), (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE)):
            c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strat)
            comp = c.compress(data) + c.flush()
            out = np.zeros(len(data) + 1, np.uint8)
            assert fn(comp, len(comp), out.ctypes.data, len(data)) == 0
            assert out[:len(data)].tobytes() == data
    data = b"hello world, hello world" * 50
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = c.compress(data) + c.flush()
    out = np.zeros(len(data) + 8, np.uint8)
    assert all(fn(comp[:k], k, out.ctypes.data, len(data)) != 0 for k in (0, 1, 5, len(comp) - 1))      # truncated input
    assert fn(comp, len(comp), out.ctypes.data, len(data) - 1) != 0 and fn(comp, len(comp), out.ctypes.data, len(data) + 1) != 0
    bad = bytearray(comp); bad[0] |= 6                                                                   # block type 3
    assert fn(bytes(bad), len(bad), out.ctypes.data, len(data)) != 0


@pytest.mark.parametrize("seed,kw", [(5, {}), (6, dict(name_style="prefix", with_seq_on_supp=True)), (7, dict(p_unmapped=0.3))])
def test_tile_chain_finds_every_record(seed, kw):
    from fslr_b200 import mapping_info as mi, synth_bam as sb
    lib = _lib("libhost_inflate.so", "host_inflate_harness.cpp")
    lib.host_tile_chain.restype = ctypes.c_longlong
    lib.host_tile_chain.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong)]
    refs, recs, _ = sb.make_alignments(1500, seed=seed, **kw)
    u = mi.inflate_bgzf(sb.bam_bytes(refs, [sb.encode_record(*r) for r in recs]))
    _, first = mi.parse_bam_header(u)
    found = ctypes.c_longlong()
    assert lib.host_tile_chain(u.ctypes.data, len(u), first, len(refs), ctypes.byref(found)) == 0
    assert found.value == len(recs)
