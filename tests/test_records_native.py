"""Native record formatter (include/o2a_records.h) against the Python statement of the reference's text
(records.chain_records = GenomicRangesAlignmentChain.to_record / to_bed12, genomicranges.py:1192-1301)."""
import os
import re
import struct

import numpy as np

from conftest import ROOT


def test_records_library_exports_every_declared_symbol():
    from ortho2align_b200 import build, records_native as R
    build.build_ingest()
    with open(os.path.join(ROOT, "include", "o2a_records.h")) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    names = sorted(set(re.findall(r"\b(o2a_(?:records|best)_[a-z_0-9]+)\s*\(", src)))
    assert set(names) == set(R.EXPORTS)
    for n in names:
        assert hasattr(R.lib(), n)


def test_repr_double_is_pythons_and_numpys():
    """str(np.float64(x)) is what the reference writes for p/q-values, repr(float) for float scores: both equal ours."""
    from ortho2align_b200 import records_native as R
    rng = np.random.default_rng(5)
    bits = rng.integers(0, 1 << 63, 60_000, dtype=np.int64).astype(np.uint64) | (rng.integers(0, 2, 60_000).astype(np.uint64) << np.uint64(63))
    xs = [struct.unpack("<d", struct.pack("<Q", int(b)))[0] for b in bits]               # every exponent, subnormals, nan, inf
    xs += (10.0 ** rng.uniform(-30, 30, 20_000) * rng.choice([-1, 1], 20_000)).tolist()
    xs += rng.random(20_000).tolist() + (rng.random(5_000) * 1e-4).tolist() + (rng.random(5_000) * 1e17).tolist()
    xs += [0.0, -0.0, 1e-4, 9.999999999999999e-05, 1e16, 9999999999999998.0, 1e15, 5e-324, 2.2250738585072014e-308,
           1.7976931348623157e308, 0.05, 1.0, 100.0, 123456.0, 1e22, 1e23, float("inf"), float("-inf")]
    xs += [float(k) for k in range(-50, 50)] + [k / 8 for k in range(-40, 40)]
    for x in xs:
        got = R.repr_double(x)
        assert got == repr(x), (x.hex() if x == x else "nan", got, repr(x))
        assert got == str(np.float64(x))


def _random_chain(rng, int_scores):
    n = int(rng.choice([1, 1, 2, 3, 8, 40]))
    reverse = bool(rng.integers(0, 2))
    qs = np.sort(rng.integers(1, 2_000_000_000, n)).astype(np.int64)
    qe = qs + rng.integers(1, 200, n)
    ss = rng.integers(1, 2_000_000_000, n).astype(np.int64)
    se = ss - rng.integers(1, 200, n) if reverse else ss + rng.integers(1, 200, n)
    if rng.random() < 0.1:
        qe[0] = qs[0]                                   # zero-length first HSP: strand flag False
    score, is_int = [], []
    for _ in range(n):
        if int_scores or rng.random() < 0.6:
            score.append(int(rng.integers(-5, 300))); is_int.append(1)
        else:
            score.append(float(rng.choice([rng.random() * 100, 1e16 * rng.random(), 0.1, 1e-7 * rng.random(), 12.5]))); is_int.append(0)
    pv = (rng.random(n) * rng.choice([1.0, 1e-3, 1e-9, 1e-300])).tolist()
    if rng.random() < 0.1:
        pv[0] = 0.0
    return qs, np.minimum(qe, 2**31 - 1), ss, np.clip(se, -2**31, 2**31 - 1), score, is_int, pv


def test_native_records_match_the_python_statement():
    from ortho2align_b200 import records as P
    from ortho2align_b200 import records_native as R
    rng = np.random.default_rng(11)
    chains, want = [], [[], [], [], []]
    strings, qplus = [], []
    for c in range(3000):
        qs, qe, ss, se, score, is_int, pv = _random_chain(rng, int_scores=(c % 3 == 0))
        strand = str(rng.choice(["+", "-", ".", "*"]))
        qchrom = rng.choice(["chr1", "chrUn_gl000220", 7, "χρ1"])               # str() of whatever the JSON held
        qchrom = int(qchrom) if str(qchrom).isdigit() else str(qchrom)
        name, schrom = str(rng.choice(["lnc1", "ENST000.1|x", "chr1:5-9+"])), str(rng.choice(["chr2", "scaffold_12"]))
        hsps = [{"qstart": int(a), "qend": int(b), "sstart": int(x), "send": int(y), "score": s}
                for a, b, x, y, s in zip(qs, qe, ss, se, score)]
        rec = P.chain_records(hsps, P.as_pvals(pv), P.Range(qchrom, 0, 1, strand, name), schrom, name)
        for k in range(4):
            want[k].append(rec[k] + "\n")
        chains.append((qs, qe, ss, se, score, is_int, pv))
        strings += [str(qchrom), name, schrom]
        qplus.append(strand in ("+", "."))
    off = np.zeros(len(chains) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(c[0]) for c in chains])
    cat = lambda k, dt: np.concatenate([np.asarray(c[k], dtype=dt) for c in chains])
    for threads in (1, 5):
        got = R.format_records(off, cat(0, np.int32), cat(1, np.int32), cat(2, np.int32), cat(3, np.int32), cat(4, np.float64),
                               cat(5, np.uint8), cat(6, np.float64), np.array(qplus, dtype=np.uint8), strings, threads=threads)
        for k in range(4):
            assert got[k].decode("utf-8") == "".join(want[k]), f"file {k}, {threads} threads"
    assert R.format_records(np.zeros(1, np.int64), *[np.zeros(0, np.int32)] * 4, np.zeros(0), np.zeros(0, np.uint8), np.zeros(0),
                            np.zeros(0, np.uint8), []) == (b"", b"", b"", b"")
