"""Burden regions, host side: the packed-row union rule (oracle.input_oracle.burden_union, the
NumPy checker of psb_submit_burden) against the reference's dictionary semantics restated in
oracle/input_oracle.py (input.py:395-411, 457-502), and the VcfReader region reader on the
reference's own VCF / burden fixtures."""
import os

import numpy as np

from conftest import GOLDEN


def _pack_states(states, n):
    from pyseer_b200.engine import words_per_row
    W = words_per_row(n)
    st = np.zeros((len(states), W * 32), dtype=np.int8)
    st[:, :n] = np.array(states, dtype=np.int8).reshape(len(states), n)
    bits = np.ascontiguousarray(np.packbits(st == 1, axis=1, bitorder='little').view('<u4'))
    miss = np.ascontiguousarray(np.packbits(st == 2, axis=1, bitorder='little').view('<u4'))
    return bits, miss


def _random_records(rng, n_rec, n, diploid):
    calls = ['0', '1', '.'] if not diploid else ['0/0', '0/1', '1/0', '1|1', './.', './0', '0/.', './1', '1/.']
    p = [0.8, 0.1, 0.1] if not diploid else [0.6, 0.05, 0.05, 0.05, 0.05, 0.05, 0.05, 0.05, 0.05]
    return [list(rng.choice(calls, size=n, p=p)) for _ in range(n_rec)]


def test_union_rule_matches_reference_dictionary():
    from oracle import input_oracle
    from oracle.input_oracle import burden_union as burden_union_host
    from pyseer_b200.engine import unpack_rows
    rng = np.random.RandomState(3)
    for diploid in (False, True):
        n, n_rec, n_reg = 37, 40, 60
        records = _random_records(rng, n_rec, n, diploid)
        # per-record rows: the dictionary of the record alone
        rec_states = [input_oracle.state_vector(input_oracle.read_vcf_var(r, {}), n) for r in records]
        vbits, vmiss = _pack_states(rec_states, n)
        offsets, members = [0], []
        for r in range(n_reg):
            k = rng.randint(0, 8) if r % 7 else 0            # some empty regions
            members += list(rng.randint(0, n_rec, size=k))   # repeats allowed
            offsets.append(len(members))
        bits, miss = burden_union_host(vbits, vmiss, np.array(offsets), np.array(members, dtype=np.int32))
        for r in range(n_reg):
            mem = members[offsets[r]:offsets[r + 1]]
            want = input_oracle.state_vector(input_oracle.burden_region(records, mem), n)
            got = unpack_rows(bits[r:r + 1], n)[0] + 2 * unpack_rows(miss[r:r + 1], n)[0]
            assert list(got) == want, (diploid, r)


def test_vcf_reader_burden_regions():
    """VcfReader on the reference fixtures (tests/variants50.vcf.gz, tests/burden_regions.txt):
    every region row equals the dictionary built record by record as input.py:395-411 does."""
    import gzip
    import re
    import pandas as pd
    from oracle import input_oracle
    from pyseer_b200.input import VcfReader
    from pyseer_b200.engine import unpack_rows
    vcf = os.path.join(GOLDEN, 'variants50.vcf.gz')
    regions = os.path.join(GOLDEN, 'burden_regions_multiple.txt')
    with gzip.open(vcf, 'rt') as fh:
        lines = [l.rstrip('\n').split('\t') for l in fh if not l.startswith('##')]
    header, recs = lines[0], lines[1:]
    samples = header[9:]
    p = pd.Series(np.zeros(len(samples)), index=samples)
    reader = VcfReader(vcf, p, regions, reducer=input_oracle.burden_union)
    batches = list(reader.batches(1000))
    assert len(batches) == 1
    b = batches[0]
    n = len(samples)
    with open(regions) as rf:
        specs = [l.rstrip().split() for l in rf]
    assert b.names == [s[0] for s in specs]
    for r, (name, spec) in enumerate(specs):
        d = {}
        for sec in spec.split(','):
            mt = re.match(r'^(.+):(\d+)-(\d+)$', sec)
            lo, hi = int(mt.group(2)) - 1, int(mt.group(3))
            for f in recs:
                start = int(f[1]) - 1
                if f[0] != mt.group(1) or not (start < hi and start + len(f[3]) > lo):
                    continue
                if len(f[4].split(',')) > 1 or (f[6] not in ('.', '', 'PASS') and 'PASS' not in f[6].split(';')):
                    continue
                gi = f[8].split(':').index('GT')
                input_oracle.read_vcf_var([c.split(':')[gi] for c in f[9:]], d)
        want = input_oracle.state_vector(d, n)
        got = unpack_rows(b.bits[r:r + 1], n)[0].astype(int)
        if b.missing is not None:
            got = got + 2 * unpack_rows(b.missing[r:r + 1], n)[0]
        assert list(got) == want, name
