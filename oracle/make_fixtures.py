"""Derives the small end-to-end fixtures under tests/golden/ from the reference's test data
(run in the build container, where /root/reference is mounted): the VCF restricted to the 50
phenotyped samples of subset.pheno, the burden region files, and the reference's own baseline
logs for the CLI cases the GPU path covers.  pyseer intersects samples before any computation,
so results on the restricted files equal those on the full files."""
import gzip
import os
import shutil
import sys

import pandas as pd

REF = '/root/reference/tests'
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def main():
    keep = set(pd.read_csv(os.path.join(REF, 'subset.pheno'), sep='\t', index_col=0).index.astype(str))
    with gzip.open(os.path.join(REF, 'variants.vcf.gz'), 'rt') as fin, \
            gzip.open(os.path.join(OUT, 'variants50.vcf.gz'), 'wt') as fout:
        cols = None
        for line in fin:
            if line.startswith('##'):
                if line.startswith(('##fileformat', '##FILTER', '##FORMAT=<ID=GT', '##contig')):
                    fout.write(line)
                continue
            f = line.rstrip('\n').split('\t')
            if line.startswith('#CHROM'):
                cols = list(range(9)) + [i for i in range(9, len(f)) if f[i] in keep]
            else:
                f[7] = '.'                      # INFO is not used by the reader
                # keep only the GT sub-field
                gi = f[8].split(':').index('GT')
                for i in cols[9:]:
                    f[i] = f[i].split(':')[gi]
                f[8] = 'GT'
            fout.write('\t'.join(f[i] for i in cols) + '\n')
    with open(os.path.join(REF, 'lineage_clusters.txt')) as fin, \
            open(os.path.join(OUT, 'lineage_clusters50.txt'), 'w') as fout:
        for line in fin:
            if line.split()[0] in keep:
                fout.write(line)
    for name in ('burden_regions.txt', 'burden_regions_multiple.txt'):
        shutil.copy(os.path.join(REF, name), os.path.join(OUT, name))
    os.makedirs(os.path.join(OUT, 'baseline'), exist_ok=True)
    for case in sys.argv[1:]:
        for ext in ('log', 'err'):
            shutil.copy(os.path.join(REF, 'baseline', '%s.%s' % (case, ext)),
                        os.path.join(OUT, 'baseline', '%s.%s' % (case, ext)))


if __name__ == '__main__':
    main()
