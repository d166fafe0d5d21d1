#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/.  TEST INFRASTRUCTURE.

Run in the build container only (it reads /root/reference, which does not exist on
the GPU box):

    python oracle/gen_golden.py

What it writes
  unit_tests_data.npz   the n=100 arrays of reference tests/unit_tests_data/*.txt
  lmm_fixture.npz       tests/subset.pheno, similarity_subset.tsv.gz, covariates.txt and
                        the 50x50 sub-blocks of distances.tsv.gz / similarity.tsv.gz
  kmers.gz, kmers_head.txt  the reference's k-mer fixture (config C1)
  baseline/*.log|err    reference CLI outputs for the invocations we replay
  reference_goldens.json  constants asserted by the reference's own unit tests
                        (tests/model_test.py, tests/lmm_test.py), with line cites
  lmm_ref_*.npz         outputs of the UNMODIFIED reference module
                        pyseer.fastlmm.lmm_cov (imported from /root/reference) on seeded
                        synthetic inputs with interior h2 (the reference's own LMM
                        goldens all have h2 = 0)
"""
import gzip
import json
import os
import shutil
import sys

import numpy as np
import pandas as pd
from scipy import stats

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def unit_tests_data():
    d = os.path.join(REF, 'tests', 'unit_tests_data')
    arrs = {}
    for name in ['p_binary', 'p_continuous', 'k', 'm', 'cov', 'lin', 'firth_vars']:
        arrs[name] = np.loadtxt(os.path.join(d, name + '.txt'))
    np.savez_compressed(os.path.join(OUT, 'unit_tests_data.npz'), **arrs)


def lmm_fixture():
    t = os.path.join(REF, 'tests')
    ph = pd.read_csv(os.path.join(t, 'subset.pheno'), index_col=0, sep='\t')
    ph.index = ph.index.astype(str)
    sim = pd.read_csv(os.path.join(t, 'similarity_subset.tsv.gz'), index_col=0, sep='\t')
    sim.index = sim.index.astype(str)
    cov = pd.read_csv(os.path.join(t, 'covariates.txt'), index_col=0, sep='\t')
    cov.index = cov.index.astype(str)
    dist = pd.read_csv(os.path.join(t, 'distances.tsv.gz'), index_col=0, sep='\t')
    dist.index = dist.index.astype(str)
    simfull = pd.read_csv(os.path.join(t, 'similarity.tsv.gz'), index_col=0, sep='\t')
    simfull.index = simfull.index.astype(str)
    s = list(ph.index)
    np.savez_compressed(
        os.path.join(OUT, 'lmm_fixture.npz'),
        samples=np.array(s),
        pheno_binary=ph['binary'].values.astype(float),
        pheno_continuous=ph['continuous'].values.astype(float),
        sim_subset_names=np.array(list(sim.index)),
        sim_subset=sim.values,
        cov_names=np.array(list(cov.index)),
        cov_quantitative=cov['quantitative'].values.astype(float),
        cov_categorical=cov['categorical'].values.astype(str),
        dist50=dist.loc[s, s].values,
        sim50=simfull.loc[s, s].values,
    )
    # text copies so the CLI can be replayed on the GPU box (small: 50 samples)
    ph.to_csv(os.path.join(OUT, 'subset.pheno'), sep='\t')
    shutil.copy(os.path.join(t, 'covariates.txt'), os.path.join(OUT, 'covariates.txt'))
    dist.loc[s, s].to_csv(os.path.join(OUT, 'distances50.tsv'), sep='\t')
    simfull.loc[s, s].to_csv(os.path.join(OUT, 'similarity50.tsv'), sep='\t')
    shutil.copy(os.path.join(t, 'kmers.gz'), os.path.join(OUT, 'kmers.gz'))
    os.makedirs(os.path.join(OUT, 'baseline'), exist_ok=True)
    for n in ['1', '5', '6', '9', '15', '20', '25', '28', '29']:
        for ext in ['log', 'err']:
            shutil.copy(os.path.join(t, 'baseline', '%s.%s' % (n, ext)),
                        os.path.join(OUT, 'baseline', '%s.%s' % (n, ext)))
    shutil.copy(os.path.join(t, 'presence_absence.Rtab.gz'),
                os.path.join(OUT, 'presence_absence.Rtab.gz'))
    for n in ['14', '24']:
        for ext in ['log', 'err']:
            shutil.copy(os.path.join(t, 'baseline', '%s.%s' % (n, ext)),
                        os.path.join(OUT, 'baseline', '%s.%s' % (n, ext)))


def reference_goldens():
    """Constants the reference's own tests assert (file:line in the key comments)."""
    g = {
        # tests/model_test.py:84-116
        'prefilter_binary_p': 0.5365065578449575,
        'prefilter_binary_bad_p': 1.4919966396986922e-19,
        'prefilter_cont_p': 0.29623810011571716,
        'prefilter_cont_binaryp_p': 8.6308642007939013e-30,
        # tests/model_test.py:119-172
        'null_binary_params': [-1.41572498, 0.35847998, -0.03014792, 2.46252819,
                               0.96908425, -0.20952455, -0.27988125, 0.36798503,
                               -0.03278285, -1.34132024, 0.844149],
        'null_binary_firth': -57.884527394557985,
        'null_binary_cov_params': [-0.87072948, 0.26456701, 0.03485904, 2.80243184,
                                   1.086393, -0.3882244, -0.46883396, 0.61387846,
                                   0.09962477, -1.45376984, 0.93929299, 0.07927743,
                                   -1.54631396, 0.1098796],
        'null_binary_cov_firth': -55.60790630835098,
        'null_cont_params': [0.65572473, -0.16129649, 0.03417796, -0.08011702,
                             0.10902641, 0.00599514, -0.09081684, -0.13653787,
                             0.17798003, -0.16793408, 0.12959982],
        'null_cont_cov_params': [0.49070237, -0.17284083, 0.00710691, -0.11784811,
                                 0.07352861, 0.01219004, -0.04772721, -0.17089199,
                                 0.18198025, -0.17141095, 0.11330439, 0.08887165,
                                 0.20304982, 0.13802362],
        # tests/model_test.py:175-195
        'lineage_index': 2,
        # tests/model_test.py:198-234
        'firth_likelihood': 97.13375906431875,
        'firth_intercept': 0.13954805021495864,
        'firth_kbeta': -0.31901219992017243,
        'firth_beta': [1.9588025, 0.7251749, -0.5605268, -0.5396909, 0.0594742,
                       -0.2001795, -1.4873298, 0.5050208],
        'firth_bse': 2.848207537910185,
        'firth_ll': -58.249948818380204,
        # tests/model_test.py:237-388 (binary)
        'fe_binary': {'prep': 0.5365065578449575, 'pvalue': 1,
                      'kbeta': -0.668215625696782, 'bse': 0.47087488598995186,
                      'intercept': -1.29962042280822,
                      'betas': [0.42265596, 0.10078512, 2.77587593, 0.94439244,
                                -0.13846857, -0.14140035, 0.38328562, -0.1986484,
                                -1.51779346, 0.94618541]},
        'fe_binary_cov': {'prep': 0.5365065578449575, 'pvalue': 1,
                          'kbeta': -0.7082070719359966, 'bse': 0.4852518061533321,
                          'intercept': -0.809194818156449,
                          'betas': [0.325464, 0.16147301, 3.17003634, 1.05383182,
                                    -0.31762591, -0.32545411, 0.65876263, -0.07939636,
                                    -1.61743885, 1.04396837, 0.13034889, -1.59225167,
                                    0.1938934]},
        # tests/model_test.py:390-517 (continuous)
        'fe_cont': {'prep': 0.29623810011571716, 'pvalue': 0.4694146479961355,
                    'kbeta': -0.043638262259610316, 'bse': 0.06006023185402142,
                    'intercept': 0.6655803214920781,
                    'betas': [-0.1560651, 0.04372272, -0.06398297, 0.10658197,
                              0.01046428, -0.08089156, -0.13733075, 0.16774866,
                              -0.17746121, 0.13386466]},
        'fe_cont_cov': {'prep': 0.29623810011571716, 'pvalue': 0.4039092383440829,
                        'kbeta': -0.04946894010582922, 'bse': 0.05897268709495734,
                        'intercept': 0.49957867277580303,
                        'betas': [-0.16730353, 0.01750906, -0.09994545, 0.07018266,
                                  0.01718979, -0.03593312, -0.17211066, 0.17065225,
                                  -0.18230721, 0.11787759, 0.09058623, 0.20484901,
                                  0.14072312]},
        # tests/lmm_test.py:66-133
        'lmm_nLL': 35.7033778, 'lmm_h2': 0.0, 'lmm_cov_nLL': 34.554038607321814,
        # tests/lmm_test.py:136-392
        'lmm_fit': {'prep': 0.28252075514059294, 'pvalue': 0.2920532220978148,
                    'kbeta': 0.1513687600644123, 'bse': 0.1420853593711293,
                    'frac_h2': 0.1519818397711344},
        'lmm_fit_badchisq': {'prep': 0.2544505826463333, 'pvalue': 0.263519965703956,
                             'kbeta': 0.2666666666666663, 'bse': 0.2357022603955158,
                             'frac_h2': 0.16116459280507586},
        'lmm_fit_cont_prep': 0.2937152511367835,
        'lmm_lineage_index': 0,
        # tests/input_test.py:839-849
        'hash_pattern_k': 'gwi2uQb68G5LfLr7qJuVpw==\n',
    }
    with open(os.path.join(OUT, 'reference_goldens.json'), 'w') as f:
        json.dump(g, f, indent=1, sort_keys=True)


def lmm_reference_vectors():
    """Run the unmodified reference FaST-LMM module on seeded inputs."""
    sys.path.insert(0, REF)
    from pyseer.fastlmm.lmm_cov import LMM as lmm_cov

    def one(tag, n, nsnp, ncov, seed, binary):
        rng = np.random.RandomState(seed)
        af0 = rng.uniform(0.05, 0.95, size=2 * n)
        G0 = (rng.uniform(size=(n, 2 * n)) < af0).astype(float)
        K = G0.dot(G0.T)
        g = G0.dot(rng.normal(size=2 * n))
        g = (g - g.mean()) / g.std()
        y = np.sqrt(0.5) * g + np.sqrt(0.5) * rng.normal(size=n)
        if binary:
            y = (y > np.median(y)).astype(float)
        cov = rng.normal(size=(n, ncov))
        covar = np.c_[cov, np.ones((n, 1))] if ncov else np.ones((n, 1))
        afs = rng.uniform(0.02, 0.98, size=nsnp)
        snps = (rng.uniform(size=(n, nsnp)) < afs).astype(float)
        # a few structured / planted columns
        snps[:, 0] = (y > np.percentile(y, 70)).astype(float)   # strongly associated
        snps[:, 1] = G0[:, 0]
        snps[:, 2] = 0.0          # constant column -> rotate() zeroes it
        snps[:, 3] = 1.0
        Kin = K.copy()
        factor = float(n) / np.diag(Kin).sum()
        Kin *= factor
        Kpass = Kin.copy()
        lmm = lmm_cov(X=covar, Y=y.reshape(-1, 1), K=Kpass, G=None, inplace=True)
        res = lmm.findH2()
        h2 = res['h2']
        out = lmm.nLLeval(h2=h2, dof=None, scale=1.0, penalty=0.0, snps=snps.copy())
        beta = out['beta']
        with np.errstate(all='ignore'):
            chi2 = beta * beta / out['variance_beta']
            pv = stats.f.sf(chi2, 1, lmm.U.shape[0] - (lmm.linreg.D + 1))[:, 0]
        np.savez_compressed(
            os.path.join(OUT, 'lmm_ref_%s.npz' % tag),
            K=Kin, y=y, cov=cov, snps=snps.astype(np.uint8), h2=np.array([h2]),
            nLL=np.asarray(res['nLL']).reshape(-1),
            U=lmm.U, S=lmm.S,
            beta=beta[:, 0], variance_beta=out['variance_beta'][:, 0],
            frac=out['fraction_variance_explained_beta'][:, 0], p_values=pv)
        print(tag, 'h2=%.6f' % h2, 'nLL=%.6f' % float(np.asarray(res['nLL']).reshape(-1)[0]),
              'min p=%.3g' % np.nanmin(pv))

    one('interior_cont', 120, 64, 0, 11, False)
    one('interior_cov', 150, 64, 2, 12, False)
    one('interior_binary', 130, 64, 1, 13, True)


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    unit_tests_data()
    lmm_fixture()
    reference_goldens()
    lmm_reference_vectors()
    print('golden fixtures written to', OUT)
