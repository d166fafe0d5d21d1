"""The native TSV formatter (psb_format_rows) against utils.format_output applied row by row as the
CLI's result loop does, and against lines produced by the reference's own format_output
(tests/golden/host_goldens.json, oracle/gen_golden_host.py)."""
import json
import os

import numpy as np

from conftest import GOLDEN
from pyseer_b200 import _lib
from pyseer_b200 import classes as var_obj
from pyseer_b200.engine import Results, notes_from_flags
from pyseer_b200.model import seer_from_row
from pyseer_b200.utils import format_output, format_table


def _table(rng, n, nb):
    r = Results()
    vals = lambda: np.where(rng.uniform(size=n) < 0.1, np.nan, rng.normal(size=n) * 10.0 ** rng.randint(-250, 6, n))
    r.af, r.prep, r.pvalue = np.abs(vals()), np.abs(vals()), np.abs(vals())
    r.beta, r.bse, r.extra = vals(), np.abs(vals()), vals()
    r.betas = rng.normal(size=(n, nb)) * 10.0 ** rng.randint(-20, 3, (n, nb))
    r.carriers = r.missing = np.zeros(n, dtype=np.int32)
    kinds = rng.randint(0, 7, n)
    f = np.zeros(n, dtype=np.uint32)
    f[kinds == 0] = _lib.F_AF_FILTER | _lib.F_PREFILTER
    f[kinds == 1] = _lib.F_PREFILTER_FAILED | _lib.F_PREFILTER
    f[kinds == 2] = _lib.F_LRT_FAILED | _lib.F_FILTER | _lib.F_TESTED
    f[kinds == 3] = _lib.F_BAD_CHISQ | _lib.F_FIRTH_FAIL | _lib.F_FILTER | _lib.F_TESTED
    f[kinds == 4] = _lib.F_TESTED
    f[kinds == 5] = _lib.F_TESTED | _lib.F_HIGH_BSE | _lib.F_FIRTH_USED
    f[kinds == 6] = _lib.F_MISSING_DATA | _lib.F_FILTER | _lib.F_TESTED
    r.flags = f
    return r


def _python_lines(r, names, model, block_size, print_filtered):
    """The CLI's result loop (pyseer_b200/__main__.py), one tuple and one format_output per row."""
    nan = np.nan
    out, pre, tested, printed = [], 0, 0, 0
    n = len(names)
    for b0 in range(0, n, block_size):
        idx = list(range(b0, min(b0 + block_size, n)))
        if model == 'lmm':
            idx = [j for j in idx if r.flags[j] & _lib.F_PREFILTER] + \
                  [j for j in idx if not (r.flags[j] & _lib.F_PREFILTER)]
        for j in idx:
            f = int(r.flags[j])
            if f & _lib.F_PREFILTER:
                pre += 1
                if not print_filtered:
                    continue
            else:
                tested += 1
                if (f & _lib.F_FILTER) and not print_filtered:
                    continue
            notes = notes_from_flags(f)
            if model == 'lmm':
                if f & _lib.F_PREFILTER:
                    item = var_obj.LMM(names[j], None, r.af[j], r.prep[j], nan, nan, nan, nan, None, [], [],
                                       notes, True, False)
                elif f & _lib.F_FILTER:
                    item = var_obj.LMM(names[j], None, r.af[j], r.prep[j], r.pvalue[j], nan, nan, nan, None,
                                       [], [], notes, False, True)
                else:
                    item = var_obj.LMM(names[j], None, r.af[j], r.prep[j], r.pvalue[j], r.beta[j], r.bse[j],
                                       r.extra[j], None, [], [], notes, False, False)
            else:
                item = seer_from_row(r, j, names[j], None, r.af[j], [], [])
            printed += 1
            out.append(format_output(item, None, model, False))
    return out, pre, tested, printed


def _same(a, b):
    """Equal lines; the notes field (a set in the reference) is compared as a set."""
    fa, fb = a.split('\t'), b.split('\t')
    return fa[:-1] == fb[:-1] and set(fa[-1].split(',')) == set(fb[-1].split(','))


def test_native_formatter_matches_result_loop():
    rng = np.random.RandomState(3)
    for model, nb in (('seer', 11), ('seer', 0), ('lmm', 0)):
        for block_size, pf in ((1, False), (7, True), (3000, False), (50, True)):
            n = 5003
            r = _table(rng, n, nb)
            names = ['K%d_%s' % (i, 'ACGT' * (i % 9)) for i in range(n)]
            want, pre, tested, printed = _python_lines(r, names, model, block_size, pf)
            text, c0, c1, c2 = format_table(r, names, model, block_size, pf)
            # ranges of whole blocks formatted by several threads give the same bytes
            assert format_table(r, names, model, block_size, pf, threads=5) == (text, c0, c1, c2)
            got = text.decode().split('\n')
            assert got[-1] == '' and len(got) - 1 == len(want) == printed
            assert (c0, c1, c2) == (pre, tested, printed)
            for a, b in zip(got, want):
                assert _same(a, b), (model, nb, a, b)


def test_native_formatter_matches_reference_lines():
    """Numbers, blanks and field order against pyseer.utils.format_output itself."""
    with open(os.path.join(GOLDEN, 'host_goldens.json')) as fh:
        cases = json.load(fh)['format_output']
    bit = {s: b for b, s in _lib.NOTE_BITS}
    num = lambda x: float(x) if isinstance(x, str) else (np.nan if x is None else x)
    for model, key, extra in (('seer', 'seer|nolin|nosamples', 'intercept'), ('lmm', 'lmm|nolin|nosamples', 'frac_h2')):
        for case in cases:
            f = case['fields']
            r = Results()
            for col, src in (('af', 'af'), ('prep', 'prep'), ('pvalue', 'pvalue'), ('beta', 'kbeta'),
                             ('bse', 'bse'), ('extra', extra)):
                setattr(r, col, np.array([num(f[src])]))
            r.betas = np.array([[num(b) for b in f['betas']]], dtype=float).reshape(1, -1)
            flags = _lib.F_TESTED
            for s in f['notes']:
                flags |= bit[s]
            flags &= ~_lib.F_FIRTH_FAIL          # that note blanks the row in the Seer tuple
            r.flags = np.array([flags], dtype=np.uint32)
            want = case['lines'][key]
            if 'firth-fail' in f['notes']:
                continue
            text, _, _, printed = format_table(r, [f['kmer']], model, 1, True)
            assert printed == 1
            assert _same(text.decode().rstrip('\n'), want), (model, f['kmer'])


def test_fast_number_formatting_equals_printf():
    """fmt_num's fast path (scale into [100, 1000) by a table of powers of ten, round, hand the cases
    within 1e-6 of a tie to snprintf) against '%.2E' on ties of exact binary fractions, decimal half-way
    cases, powers of ten, both ends of the fast range, subnormals and signed zeros."""
    rng = np.random.RandomState(1)
    vals = [0.0, -0.0, 1.0, -1.0, 1.125, 1.135, 1.145, 1.115, 9.995, 9.985, 99.95e-7, 9.995e22, 9.995e-23, 999.5,
            99.95, 1e-5, 1e5, 1e22, 1e23, 1e-22, 1e-23, 5e-324, 2.2e-308, 1e-300, 1e300, 1.7976931348623157e308,
            1e-290, 1e290, 1e-291, 9.994999999999999, 9.995000000000001, 0.001005, 1.005, 1.015, 1.025, 2.675,
            123456789.0, 0.000123456]
    vals += list((rng.randint(1, 2000, 4000) / 8.0) * 10.0 ** rng.randint(-30, 30, 4000))
    vals += list(rng.randint(100, 1000, 3000) / 100.0 + 0.005)
    vals += list(rng.normal(size=20000) * 10.0 ** rng.randint(-300, 300, 20000))
    edge = np.array([1.125, 9.995, 99.95, 2.5e-7, 1e-290, 1e290])
    vals += list(np.nextafter(edge, np.inf)) + list(np.nextafter(edge, -np.inf))
    v = np.array(vals, dtype=float)
    n = len(v)
    r = Results()
    r.af, r.prep, r.pvalue, r.beta, r.bse, r.extra = v.copy(), np.abs(v), np.abs(v), v.copy(), np.abs(v), v.copy()
    r.flags = np.full(n, _lib.F_TESTED, dtype=np.uint32)
    r.betas = None
    r.carriers = r.missing = np.zeros(n, dtype=np.int32)
    text, _, _, _ = format_table(r, ['v%d' % i for i in range(n)], 'lmm', 3000, True, threads=4)
    lines = text.decode().split('\n')[:-1]
    assert len(lines) == n
    for i, line in enumerate(lines):
        f = line.split('\t')
        assert f[1] == '%.2E' % v[i] and f[5] == '%.2E' % abs(v[i]), (repr(v[i]), f)


def test_native_formatter_lineage_column():
    """psb_format_rows_lineage: the lineage column of --lineage runs (utils.py:93-97) sits between the
    coefficients and the notes, holds the lineage's name or NA; against format_output row by row."""
    rng = np.random.RandomState(5)
    lineage_dict = ['MDS1', 'MDS2', 'BAPS_cluster_3', 'x']
    for model, nb in (('seer', 4), ('seer', 0), ('lmm', 0)):
        n = 2003
        r = _table(rng, n, nb)
        names = ['K%d' % i for i in range(n)]
        lin = rng.randint(-1, len(lineage_dict), n).astype(np.int32)
        text, c0, c1, c2 = format_table(r, names, model, 50, True, threads=3, lineage=lin,
                                        lineage_names=lineage_dict)
        got = text.decode().split('\n')[:-1]
        plain, p0, p1, p2 = format_table(r, names, model, 50, True, threads=3)
        base = plain.decode().split('\n')[:-1]
        assert len(got) == len(base) == c2 and (c0, c1, c2) == (p0, p1, p2)
        # the rows come out in the formatter's order: recover each row's index from its name
        for a, b in zip(got, base):
            fa, fb = a.split('\t'), b.split('\t')
            j = int(fa[0][1:])
            want = lineage_dict[lin[j]] if lin[j] >= 0 else 'NA'
            assert fa[:-2] == fb[:-1] and fa[-2] == want and fa[-1] == fb[-1], (model, a, b)
        # and one tuple through the reference-mirroring format_output
        j = int(got[0].split('\t')[0][1:])
        item = var_obj.LMM(names[j], None, r.af[j], r.prep[j], np.nan, np.nan, np.nan, np.nan,
                           int(lin[j]) if lin[j] >= 0 else None, [], [], notes_from_flags(int(r.flags[j])), True, False)
        if model == 'lmm' and (int(r.flags[j]) & _lib.F_PREFILTER):
            assert _same(got[0], format_output(item, lineage_dict, 'lmm', False))


def test_similarity_matrix_writer_equals_pandas():
    """similarity.write_matrix (psb_format_matrix) against DataFrame.to_csv(sep='\\t'), the reference's
    output call (similarity.py:118-120): integer counts incl. 0 and 12-digit values, one sample and many;
    non-integer matrices and names a csv writer would quote go through pandas itself."""
    import io
    import pandas as pd
    from pyseer_b200.similarity import write_matrix
    rng = np.random.RandomState(0)
    for n in (1, 2, 65, 300):
        K = rng.randint(0, 200000, (n, n)).astype(float)
        K[0, 0] = 0
        K[-1, 0] = 999999999999.0
        names = ['s%d_x' % i for i in range(n)]
        a, b = io.StringIO(), io.StringIO()
        write_matrix(K, names, a)
        pd.DataFrame(K, index=names, columns=names).to_csv(b, sep='\t')
        assert a.getvalue() == b.getvalue()
    for K, names in ((rng.uniform(size=(5, 5)), list('abcde')), (np.ones((3, 3)), ['a b', 'c,d', 'e"f'])):
        a, b = io.StringIO(), io.StringIO()
        write_matrix(K, names, a)
        pd.DataFrame(K, index=names, columns=names).to_csv(b, sep='\t')
        assert a.getvalue() == b.getvalue()
