"""CPU-only tests: the C-ABI library loads and exports what include/mgpu.h declares, fails loudly
without a GPU, and the index writer's bytes round-trip through the (independent) oracle reader."""
import ctypes as C
import os
import re
import struct

import pytest

import helpers
import manticoresearch_b200.mgpu as M
from conftest import has_gpu

ROOT = helpers.ROOT


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "mgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mgpu_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = M.lib()
    declared = _declared_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "libmgpu.so does not export %s" % name
    assert sorted(M.EXPORTED_SYMBOLS) == declared
    assert lib.mgpu_abi_version() == 1


def test_writer_library_is_separate_and_host_only():
    """libmgpu_writer.so exports what include/mgpu_writer.h declares and does not depend on CUDA: processes that only build
    index files (the CPU arm of bench.py) never map the GPU library"""
    import subprocess
    from manticoresearch_b200 import build as B
    w = M.writer_lib()
    src = open(os.path.join(ROOT, "include", "mgpu_writer.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    declared = sorted(set(re.findall(r"\b(mgpu_[a-z_0-9]+)\s*\(", src)))
    assert declared == sorted(M.WRITER_EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(w, name)
    assert w.mgpu_writer_abi_version() == 1
    needed = subprocess.run(["readelf", "-d", B.WRITER_LIB], capture_output=True, text=True).stdout
    assert "libcuda" not in needed and "libmgpu.so" not in needed
    lib = M.lib()
    for name in ("mgpu_build_index", "mgpu_build_synthetic"):
        assert not hasattr(lib, name), "the writer moved out of libmgpu.so"


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_open_fails_loudly_without_gpu(golden_indexes):
    with pytest.raises(M.MgpuError) as e:
        M.Index(golden_indexes["test_019"])
    assert e.value.code == M.MGPU_E_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_header_is_format_v62(golden_indexes):
    raw = open(golden_indexes["test_019"] + ".sph", "rb").read()
    magic, version, nfields = struct.unpack_from("<III", raw, 0)
    assert magic == 0x58485053 and version == 62 and nfields == 2
    for ext in ("spd", "spp", "spe", "spi"):
        assert open(golden_indexes["test_019"] + "." + ext, "rb").read(1) == b"\x01"   # dummy first byte, src/sphinx.cpp:8404-8409


def test_synthetic_index_roundtrip(tmp_path):
    """every posting the oracle decodes from the written files equals the corpus definition"""
    prefix = str(tmp_path / "rt")
    p = M.SynthParams(3000, vocab=5000, threads=3)
    M.build_synthetic(prefix, p)
    # ground truth straight from the corpus definition
    postings = {}
    for d in range(3000):
        for f in range(2):
            n = M.synth_field_len(p, d, f)
            for k in range(n):
                t = M.synth_token(p, d, f, k)
                postings.setdefault(t, {}).setdefault(d, []).append((f << 24) | (k + 1) | ((1 << 23) if k == n - 1 else 0))
    idx = helpers.OracleIndex(prefix)
    try:
        assert idx.total_docs == 3000
        for t in list(range(0, 60)) + list(range(100, 5000, 97)):
            w = M.synth_keyword(t)
            exp = postings.get(t)
            got = idx.decode_doclist(w)
            if exp is None:
                assert got is None
                continue
            rowid, hits, fields, pos = got
            assert list(rowid) == sorted(exp.keys())
            for i, d in enumerate(sorted(exp.keys())):
                hl = sorted(exp[d])
                assert hits[i] == len(hl)
                fm = 0
                for h in hl:
                    fm |= 1 << (h >> 24)
                assert fields[i] == fm
                assert idx.decode_hitlist(w, pos[i]) == hl
    finally:
        idx.close()


def test_ctypes_mirrors_match_the_header(tmp_path):
    """sizeof / last-field offset of every public struct as gcc sees include/mgpu.h == the ctypes mirror the tests and bench.py use
    (a stale mirror would read garbage stats or mis-marshal queries without failing loudly)"""
    import subprocess
    pairs = [("mgpu_xqkeyword", M.c_xqkeyword), ("mgpu_xqnode", M.c_xqnode), ("mgpu_sortkey", M.c_sortkey), ("mgpu_filter", M.c_filter),
             ("mgpu_query", M.c_query), ("mgpu_wordstat", M.c_wordstat), ("mgpu_result", M.c_result), ("mgpu_batch_stats", M.c_batch_stats), ("mgpu_sharded_stats", M.c_sharded_stats),
             ("mgpu_build_doc_input", M.c_build_doc_input), ("mgpu_synth_params", M.SynthParams), ("mgpu_parser_settings", M.c_parser_settings)]
    src = tmp_path / "sizes.c"
    body = "".join('printf("%%s %%zu %%zu\\n", "%s", sizeof(%s), offsetof(%s, %s));\n' % (c, c, c, py._fields_[-1][0]) for c, py in pairs)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "mgpu.h"\n#include "mgpu_writer.h"\nint main(void){\n' + body + "return 0;}\n")
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    got = {l.split()[0]: (int(l.split()[1]), int(l.split()[2])) for l in out if l.strip()}
    for c, py in pairs:
        assert got[c] == (C.sizeof(py), getattr(py, py._fields_[-1][0]).offset), (c, got[c], C.sizeof(py))


def test_same_corpus_any_thread_count(tmp_path):
    """the parallel builder is deterministic: 1 thread and 5 threads write identical bytes"""
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    M.build_synthetic(a, M.SynthParams(2000, vocab=3000, threads=1))
    M.build_synthetic(b, M.SynthParams(2000, vocab=3000, threads=5))
    for ext in ("spd", "spp", "spe", "spi", "spa", "sph"):
        assert open(a + "." + ext, "rb").read() == open(b + "." + ext, "rb").read(), ext


def test_shard_builder_is_a_rowid_range_of_the_corpus(tmp_path):
    """docs [first_doc, first_doc+n) written as a shard carry the same postings as that range of the full index"""
    full, shard = str(tmp_path / "full"), str(tmp_path / "shard")
    M.build_synthetic(full, M.SynthParams(1200, vocab=2000, threads=2))
    M.build_synthetic(shard, M.SynthParams(500, first_doc=700, vocab=2000, threads=2))
    fi, si = helpers.OracleIndex(full), helpers.OracleIndex(shard)
    try:
        for t in (0, 1, 5, 50, 300, 999):
            w = M.synth_keyword(t)
            f, s = fi.decode_doclist(w), si.decode_doclist(w)
            if f is None:
                assert s is None
                continue
            rows = [(int(r), int(h), int(m)) for r, h, m in zip(f[0], f[1], f[2]) if 700 <= r < 1200]
            got = [] if s is None else [(int(r) + 700, int(h), int(m)) for r, h, m in zip(s[0], s[1], s[2])]
            assert rows == got
    finally:
        fi.close()
        si.close()


def test_index_check_passes_what_the_writer_and_the_reference_wrote(tmp_path, golden_indexes, golden_indexes_crc):
    """mgpu_index_check (DiskIndexChecker_c, src/indexcheck.cpp): no failures on the product writer's indexes (synthetic with skiplists,
    golden corpora in both dictionary forms, plain hit format) nor on the index the reference's own writer produced"""
    prefix = str(tmp_path / "chk")
    M.build_synthetic(prefix, M.SynthParams(4000, vocab=3000, threads=2))
    assert M.check_index(prefix) == (0, "")
    for name in ("test_019", "test_116", "test_055"):
        assert M.check_index(golden_indexes[name]) == (0, ""), name
        assert M.check_index(golden_indexes_crc[name]) == (0, ""), name
    docs = [{"id": 5 + i, "fields": [[("a", 1), ("b", 2)], [("a", 1)] * 0 + [("c", p + 1) for p in range(1 + i % 4)]], "attrs": [i]} for i in range(200)]
    plain = str(tmp_path / "plain")
    M.build_index(plain, ["title", "body"], docs, attr_names=["n"], hit_format_inline=False)
    assert M.check_index(plain) == (0, "")
    ref = os.path.join(ROOT, "tests", "golden", "ref_index", "index.0")
    assert M.check_index(ref) == (0, "")


def test_index_check_finds_corruption(tmp_path):
    """each kind of damage is reported in the reference's words, and never crashes the checker"""
    import shutil
    prefix = str(tmp_path / "ok")
    M.build_synthetic(prefix, M.SynthParams(3000, vocab=2000, threads=1))

    def damaged(ext, fn):
        dst = str(tmp_path / ("bad_" + ext))
        for e in ("sph", "spi", "spd", "spp", "spe", "spa", "spm"):
            shutil.copy(prefix + "." + e, dst + "." + e)
        raw = bytearray(open(dst + "." + ext, "rb").read())
        fn(raw)
        open(dst + "." + ext, "wb").write(bytes(raw))
        return M.check_index(dst)

    def flip(pos, val=None):
        def f(raw):
            raw[pos] = (raw[pos] ^ 0x15) if val is None else val
        return f

    n, rep = damaged("spe", flip(40))
    assert n > 0 and "skiplist" in rep
    n, rep = damaged("spd", flip(1000))
    assert n > 0
    n, rep = damaged("spp", lambda raw: raw.__setitem__(slice(5000, 5004), b"\0\0\0\0"))     # hitlists end early
    assert n > 0 and ("hit" in rep)
    n, rep = damaged("spm", lambda raw: raw.extend(b"\0\0\0\0"))
    assert n == 1 and "dead row map" in rep
    n, rep = damaged("spa", lambda raw: raw.__setitem__(slice(16, 24), raw[0:8]))      # row 1 gets row 0's document id
    assert n >= 1 and "duplicate of docid" in rep
    n, rep = damaged("spd", lambda raw: raw.__delitem__(slice(len(raw) // 2, len(raw))))
    assert n > 0
    n, rep = damaged("spi", lambda raw: raw.__setitem__(slice(2, 6), b"zzzz"))       # the first keyword now sorts behind the second
    assert n > 0 and "word order decreased" in rep
    with pytest.raises(M.MgpuError):
        M.check_index(str(tmp_path / "nosuchindex"))
